"""Host planning time of one config-#2 batch (tools; prints stats()['plan_us'])."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import soundswallower_b200 as ssb
g = np.load(bench.GOLDEN)
m = ssb.AcousticModel(bench.MODEL, device=0)
U = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
base, chain = bench.config2_template(g)
feat = np.tile(base[None], (U, 1, 1)).reshape(-1, m.blk)
frame_off = np.arange(U + 1, dtype=np.int64) * bench.FRAMES
phone_off = np.arange(U + 1, dtype=np.int64) * len(chain["ssid"])
flat = {k: np.tile(chain[k], U) for k in ("ssid", "tmat", "sf", "ef")}
b = ssb.StateAlignBatch(m)
for i in range(3):
    t0 = time.perf_counter()
    b.upload_raw(feat, frame_off, phone_off, flat["ssid"], flat["tmat"], flat["sf"], flat["ef"])
    print("upload %.1f ms, plan %.2f ms" % (1e3 * (time.perf_counter() - t0), b.stats()["plan_us"] / 1e3))
