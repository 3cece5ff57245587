#!/usr/bin/env python
"""Randomised stress of the aligner against the oracle (not collected by pytest; run on a GPU box:
`python tests/stress_align.py [iterations]`).  Every iteration draws a ragged batch (random chains,
word windows feasible and not, unwindowed chains, flags carried in) and a random combination of the
kernel switches -- chain cutting, lanes per utterance, K1 kernel -- and compares rv / best score /
every state entry with the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import soundswallower_b200 as ssb  # noqa: E402
from conftest import model_dir, model_features, random_chain  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402


def main():
    n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    models = {l: (ssb.AcousticModel(model_dir(l), device=0), Oracle(model_dir(l))) for l in ("en-us", "fr-fr")}
    n_checked = n_fail = 0
    for it in range(n_iter):
        rs = np.random.RandomState(1000 + it)
        lang = ("en-us", "fr-fr")[it % 2]
        m, o = models[lang]
        env = {"SSB_K3_CUT": rs.choice(["all", "0", ""]), "SSB_K3_LANES": rs.choice(["8", "16", "32", ""]),
               "SSB_K1": rs.choice(["ft", "tc2", ""])}
        for k, v in env.items():
            if v:
                os.environ[k] = v
            else:
                os.environ.pop(k, None)
        arrays = o.model_arrays()
        n_utts = int(rs.randint(1, 48))
        feats, chains, init = [], [], []
        for u in range(n_utts):
            T = int(rs.randint(30, 260))
            feats.append(model_features(rs, arrays, T))
            chains.append(random_chain(rs, o, int(rs.randint(1, 30)), T, windowed=rs.rand() < 0.75))
            init.append(sorted(set(int(x) for x in rs.randint(0, m.n_sen, size=rs.randint(0, 4)))))
        use_init = rs.rand() < 0.4
        b = ssb.StateAlignBatch(m)
        b.upload(feats, chains, init_active=init if use_init else None)
        b.run()
        res = b.per_utt(b.download())
        b.close()
        for u, (f, c, r) in enumerate(zip(feats, chains, res)):
            w = o.state_align(f, c["ssid"], c["tmat"], c["sf"], c["ef"], init_active=init[u] if use_init else None)
            ok = r["rv"] == w["rv"]
            if ok and w["rv"] == 0:
                ok = r["best_score"] == w["best_score"] and all(np.array_equal(r[k], w[k]) for k in ("start", "dur", "score"))
            n_checked += 1
            if not ok:
                n_fail += 1
                print("MISMATCH iteration %d utterance %d env %s lang %s" % (it, u, env, lang))
    print("stress_align: %d utterances over %d iterations, %d mismatches" % (n_checked, n_iter, n_fail))
    return 1 if n_fail else 0


if __name__ == "__main__":
    sys.exit(main())
