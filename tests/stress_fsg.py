#!/usr/bin/env python
"""Randomised stress of the grammar search in the reference's default mode (active lists, scoring
inside the search kernel) against the oracle -- not collected by pytest; run on a GPU box:
`python tests/stress_fsg.py [iterations]`.  Random noise levels, truncations, tie-heavy scalings
and synthetic features, two grammars, both top-N kernels for the first pass; the whole history
table, evaluation counts, final flags, exit and segmentation must equal the oracle's."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import soundswallower_b200 as ssb  # noqa: E402
from conftest import GOLDEN, model_dir, model_features  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from test_oracle_fsg import graph_of  # noqa: E402


def main():
    n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    n_checked = n_fail = 0
    for lang in ("en-us", "fr-fr"):
        m, o = ssb.AcousticModel(model_dir(lang), device=0), Oracle(model_dir(lang))
        g = np.load(os.path.join(GOLDEN, "fsg_%s.npz" % lang))
        feat = np.load(os.path.join(GOLDEN, "align_%s.npz" % lang))["feat"]
        graphs = [graph_of(g, "align"), graph_of(g, "jsgf")]
        for it in range(n_iter):
            rs = np.random.RandomState(7000 + it)
            k1 = rs.choice(["ft", "tc2", ""])
            if k1:
                os.environ["SSB_K1"] = k1
            else:
                os.environ.pop("SSB_K1", None)
            feats, ug = [], []
            for u in range(int(rs.randint(1, 24))):
                k = int(rs.randint(0, 6))
                if k == 0:
                    f = (feat + rs.normal(0, rs.uniform(0, 0.5), feat.shape)).astype(np.float32)
                elif k == 1:
                    f = feat[:int(rs.randint(1, len(feat)))]
                elif k == 2:
                    f = feat.copy()
                    f[::int(rs.randint(2, 9))] *= np.float32(rs.choice([30, 300, 3000]))
                elif k == 3:
                    f = np.concatenate([feat, feat[int(rs.randint(0, 100)):]])
                elif k == 4:
                    f = model_features(rs, o.model_arrays(), int(rs.randint(1, 120)))
                else:
                    f = (feat * np.float32(rs.uniform(0.5, 2.0))).astype(np.float32)
                feats.append(np.ascontiguousarray(f, np.float32))
                ug.append(int(rs.randint(0, 2)))
            res = ssb.fsg_batch(m, feats, graphs, utt_graph=ug, want_hist=True, compallsen=False)
            for u, (f, r) in enumerate(zip(feats, res)):
                w = o.fsg_search_active(graphs[ug[u]], f)
                ok = (r["rv"] == w["rv"] == 0 and np.array_equal(r["hist"], w["hist"])
                      and r["n_hmm_eval"] == w["n_hmm_eval"] and r["n_sen_eval"] == w["n_sen_eval"]
                      and np.array_equal(r["active"], w["active"]) and r["exit"] == w["exit"])
                if ok and w["exit"] > 0:
                    ok = r["hyp_score"] == w["hyp_score"] and np.array_equal(r["segs"], w["segs"])
                n_checked += 1
                if not ok:
                    n_fail += 1
                    print("MISMATCH %s iteration %d utterance %d K1=%r" % (lang, it, u, k1))
        m.close()
    print("stress_fsg: %d utterances, %d mismatches" % (n_checked, n_fail))
    return 1 if n_fail else 0


if __name__ == "__main__":
    sys.exit(main())
