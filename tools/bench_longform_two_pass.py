#!/usr/bin/env python
"""BASELINE config #4 from text: ONE utterance of about an hour (fr-fr, 714 repetitions of the
sentence + silence padding, 2856 words) and its transcript through ssb_align_texts -- grammar,
first pass in the reference's default mode, chains, second pass, JSON.  The first pass is one
warp walking 360 000 frames (beam search is sequential in time); invariants only at this size
(tests/test_gpu_parity.py::test_long_two_pass_alignment_from_text is the bit-exact check on 280
words).  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import soundswallower_b200 as ssb  # noqa: E402
from bench_longform import build  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 714
    hmm = os.path.join(ROOT, "soundswallower_b200/model/fr-fr")
    g = np.load(os.path.join(ROOT, "tests/golden/align_fr-fr.npz"))
    x = build(g, reps, 504)[0]
    text = " ".join(["avance de dix mètres"] * reps)
    m = ssb.AcousticModel(hmm)
    lx = ssb.Lexicon(m, hmmdir=hmm)
    t0 = time.perf_counter()
    ta = ssb.TextAlignment(m, lx, [x], [text], align_level=1)
    t1 = time.perf_counter()
    rv, hyp, nfr = ta.status(0)
    wd, ph = ta.entries(0, "words"), ta.entries(0, "phones")
    j = ta.json(0, align_level=1)
    t2 = time.perf_counter()
    ok = (rv == 0 and wd[0, 1] == 0 and (wd[1:, 1] == wd[:-1, 1] + wd[:-1, 2]).all()
          and wd[-1, 1] + wd[-1, 2] == len(x) and (ph[1:, 1] == ph[:-1, 1] + ph[:-1, 2]).all())
    real = [w for w in json.loads(j)["w"] if not w["t"].startswith("<")]
    print(json.dumps({"workload": "config#4 from text: fr-fr, 1 utterance, %d frames, %d-word transcript"
                                  % (len(x), 4 * reps),
                      "rv": rv, "invariants_ok": bool(ok), "words_aligned": len(real), "phones": int(len(ph)),
                      "align_texts_s": t1 - t0, "json_s": t2 - t1,
                      "audio_s_per_s": len(x) / 100.0 / (t2 - t0), "pass1_kernel_ms": ta.kernel_ms()}))


if __name__ == "__main__":
    main()
