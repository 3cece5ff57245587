#!/usr/bin/env python
"""config #5 alone (tools/bench_configs.py:config5_two_pass) -- quick A/B of host-thread counts."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import soundswallower_b200 as ssb
import bench_configs as bc
m = ssb.AcousticModel(os.path.join(ROOT, "soundswallower_b200/model/en-us"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
r = bc.config5_two_pass(ssb, m, total_utts=n)
print(json.dumps({k: r[k] for k in ("utts", "aligned", "wall_s", "host_threads")}), r["utts"] * 10 / r["wall_s"])
