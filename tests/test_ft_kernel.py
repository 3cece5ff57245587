"""K1 frame-tiled (csrc/gmm_scan_ft.cu): tcgen05 screening with the A tile in tensor memory,
quantised decisions on the common path, queued exact evaluations otherwise.

Checked against the oracle (ref: src/ptm_mgau.c:63-253):
  * the screening error bound holds for every (frame, codebook, stream, density);
  * in `exact` mode the lists are the reference's raw top-N, bit for bit;
  * in production mode the codewords (in order) and scores >> 10 are the reference's --
    everything the mixing stage reads (ref: src/ptm_mgau.c:276-285, 372-389).
"""
import numpy as np
import pytest

import soundswallower_b200 as ssb
from test_gpu_parity import _exact_dist64, model_features

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _ft(monkeypatch):
    monkeypatch.setenv("SSB_K1", "ft")


def _feats(rs, o, g):
    return [g["feat"][:96], model_features(rs, o.model_arrays(), 64, noise=1.5),
            (g["feat"][100:140] * 3.0).astype(np.float32), g["feat"], g["feat"][:129], g["feat"][:1]]


@pytest.mark.parametrize("lang", ["en-us", "fr-fr"])
def test_ft_error_bound_and_exact_lists(models, oracles, golden, lang):
    m, o, g = models(lang), oracles(lang), golden[lang]
    feats = _feats(np.random.RandomState(17), o, g)
    cw, sc, approx, eps, cnt = ssb.tc_probe(m, feats)
    feat = np.concatenate(feats)
    exact = _exact_dist64(o.model_arrays(), feat)
    ratio = np.abs(approx.astype(np.float64) - exact) / eps
    assert np.isfinite(approx).all() and (eps > 0).all()
    assert ratio.max() <= 1.0, ratio.max()
    assert np.median(cnt["eps_regular"]) < 64
    off = 0
    for f in feats:
        ocw, osc = o.topn_all(f)
        assert np.array_equal(sc[off:off + len(f)], osc) and np.array_equal(cw[off:off + len(f)], ocw)
        off += len(f)
    assert cnt["scan_steps"] == len(feat) * m.n_mgau * m.n_feat
    assert cnt["exact_evals"] < 0.10 * cnt["scan_steps"] * m.n_density


@pytest.mark.parametrize("lang", ["en-us", "fr-fr"])
def test_ft_quantised_lists(models, oracles, golden, lang, monkeypatch):
    """Production mode: no exact evaluation where the screening scores decide."""
    m, o, g = models(lang), oracles(lang), golden[lang]
    monkeypatch.setenv("SSB_FT_EXACT", "0")
    feats = _feats(np.random.RandomState(5), o, g)
    cw, sc, approx, eps, cnt = ssb.tc_probe(m, feats)
    off = 0
    for f in feats:
        ocw, osc = o.topn_all(f)
        assert np.array_equal(cw[off:off + len(f)], ocw)
        assert np.array_equal(sc[off:off + len(f)] >> 10, osc >> 10)
        off += len(f)
    # most steps are decided without any exact evaluation
    assert cnt["exact_evals"] < 2.0 * cnt["scan_steps"]


def test_ft_topn_batch_equals_fp32_kernel(models, oracles, golden, monkeypatch):
    m, o, g = models("en-us"), oracles("en-us"), golden["en-us"]
    rs = np.random.RandomState(23)
    feats = [g["feat"][:50]] + [model_features(rs, o.model_arrays(), int(rs.randint(1, 300))) for _ in range(40)]
    cw_ft, sc_ft = ssb.topn_batch(m, feats)
    monkeypatch.setenv("SSB_K1", "fp32")
    cw_fp, sc_fp = ssb.topn_batch(m, feats)
    for a, b in zip(sc_ft, sc_fp):
        assert np.array_equal(a, b)
    for a, b in zip(cw_ft, cw_fp):
        assert np.array_equal(a, b)
