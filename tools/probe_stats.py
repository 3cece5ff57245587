"""GPU: statistics of the tensor-core screening on the golden audio (how tight is eps, how
many survivors are re-scored exactly)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import soundswallower_b200 as ssb
from oracle.oracle import Oracle

for lang in ("en-us", "fr-fr"):
    g = np.load("tests/golden/align_%s.npz" % lang)
    m = ssb.AcousticModel(os.path.join("soundswallower_b200/model", lang))
    o = Oracle(os.path.join("soundswallower_b200/model", lang))
    feat = g["feat"]
    cw, sc, approx, eps, cnt = ssb.tc_probe(m, [feat])
    a = o.model_arrays()
    mean, var, det = (a[k].astype(np.float64) for k in ("mean", "var", "det"))
    x = feat.astype(np.float64).reshape(feat.shape[0], 1, 3, 1, 13)
    exact = det[None] - (((x - mean[None]) ** 2) * var[None]).sum(-1)
    err = np.abs(approx - exact)
    print(lang, "frames", len(feat), "exact evals per scan step: %.2f" % (cnt["exact_evals"] / cnt["scan_steps"]))
    print("  slow-path steps: %d of %d; hot densities: %d" % (cnt["slow_steps"], cnt["scan_steps"], cnt["hot"].sum()))
    print("  eps (regular) percentiles (raw units) 10/50/90/99:", np.percentile(cnt["eps_regular"], [10, 50, 90, 99]).round(0))
    print("  err/eps max %.3f  median %.4f" % ((err / eps).max(), np.median(err / eps)))
    eps = cnt["eps_regular"]
    # how many densities lie within eps of the 4th best (the intrinsic survivor count)
    srt = np.sort(exact, -1)[..., ::-1]
    fourth = srt[..., 3]
    for mult in (0.0, 0.25, 1.0):
        n = (exact >= (fourth - mult * eps)[..., None]).sum(-1)
        print("  densities >= 4th best - %.2f eps: mean %.2f  p99 %d" % (mult, n.mean(), np.percentile(n, 99)))
    gap = (srt[..., 3] - srt[..., 4])
    print("  gap 4th-5th best percentiles 10/50/90:", np.percentile(gap, [10, 50, 90]).round(0))
