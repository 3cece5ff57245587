"""Semi-continuous acoustic models (s2_semi_mgau): one codebook, per-stream
normaliser = its own best, optional top-N beam, 4-bit clustered weights selected by senone
parity, no normalisation over senones.

No bundled model is semi-continuous, so the models are synthetic (tests/model_variants.py:
the en-us mdef/transitions, Gaussians drawn from the en-us codebooks, seeded mixture weights);
tests/golden/semi_en-us.npz holds what the compiled reference computed on them
(tools/make_golden.py --semi): weight-table digests, the dense senone-score matrix, and the
chain alignments (windowed / unwindowed / compallsen) of the goforward utterance.
"""
import hashlib
import os

import numpy as np
import pytest

import model_variants as mv
from conftest import GOLDEN, chain_from_golden, model_dir


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def semi_golden():
    return np.load(os.path.join(GOLDEN, "semi_en-us.npz"))


@pytest.fixture(scope="module")
def semi_dirs(tmp_path_factory, golden):
    root = str(tmp_path_factory.mktemp("semi"))
    out = {}
    for tag, kw, beam in mv.SEMI_CASES:
        d = os.path.join(root, tag)
        mv.write_semi_model(model_dir("en-us"), d, int(golden["en-us"]["dims"][4]), **kw)
        out[tag] = (d, beam)
    return out


TAGS = [c[0] for c in mv.SEMI_CASES]
MODES = ["win", "nowin", "win_call"]


def check_alignment(r, g, key):
    st = g[key + "states"]
    assert r["rv"] == int(g[key + "rv"])
    assert r["best_score"] == int(g[key + "best"])
    assert np.array_equal(r["start"], st[:, 1]) and np.array_equal(r["dur"], st[:, 2])
    assert np.array_equal(r["score"], st[:, 3])


# ------------------------------------------------------------------ oracle (CPU)
@pytest.mark.parametrize("tag", TAGS)
def test_oracle_semi_scores(semi_golden, semi_dirs, golden, tag):
    from oracle.oracle import Oracle
    d, beam = semi_dirs[tag]
    o = Oracle(d)
    assert o.n_mgau == 1
    if beam:
        o.set_topn_beam(beam)
    a = o.model_arrays()
    assert sha(a["mixw"]) == str(semi_golden[tag + "_mixw_sha"])
    assert sha(a["det"]) == str(semi_golden[tag + "_det_sha"])
    feat = golden["en-us"]["feat"]
    dense = o.score_all(feat)
    assert np.array_equal(dense[[0, 1, 100, len(feat) - 1]], semi_golden[tag + "_senscr_rows"])
    assert sha(dense) == str(semi_golden[tag + "_senscr_sha"])


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("tag", TAGS)
def test_oracle_semi_alignment(semi_golden, semi_dirs, golden, tag, mode):
    from oracle.oracle import Oracle
    d, beam = semi_dirs[tag]
    o = Oracle(d)
    if beam:
        o.set_topn_beam(beam)
    g = golden["en-us"]
    chain = chain_from_golden(g, windows=mode != "nowin")
    r = o.state_align(g["feat"], chain["ssid"], chain["tmat"], chain["sf"], chain["ef"],
                      compallsen=mode.endswith("call"), want_tokens=True, want_senscr=True)
    key = "%s_%s_" % (tag, mode)
    check_alignment(r, semi_golden, key)
    assert sha(r["tokens"]) == str(semi_golden[key + "tokens_sha"])
    chain_sen = o.model_arrays()["sseq"][chain["ssid"]].reshape(-1)
    assert np.array_equal(r["senscr"][:, chain_sen], semi_golden[key + "chain_scr"])


# ------------------------------------------------------------------ product, host side
@pytest.mark.parametrize("tag", TAGS)
def test_product_loads_semi_model(semi_golden, semi_dirs, tag):
    import soundswallower_b200 as ssb
    d, beam = semi_dirs[tag]
    m = ssb.AcousticModel(d, device=-1, topn_beam=beam)
    assert m.kind == 1 and m.n_mgau == 1
    a = m.arrays()
    assert sha(a["mixw"]) == str(semi_golden[tag + "_mixw_sha"])
    assert sha(a["det"]) == str(semi_golden[tag + "_det_sha"])
    assert not a["sen2cb"].any()
    m.close()
    ptm = ssb.AcousticModel(model_dir("en-us"), device=-1)
    assert ptm.kind == 0
    ptm.close()


def test_continuous_models_are_declined(tmp_path, golden):
    """Neither one codebook per CI phone nor a single one: the reference falls through to
    ms_mgau; this library declines like ptm_mgau_init / s2_semi_mgau_init do."""
    import soundswallower_b200 as ssb
    src = model_dir("en-us")
    d = str(tmp_path / "cont")
    mv.write_semi_model(src, d, int(golden["en-us"]["dims"][4]), n_density=8)
    n_mgau, n_feat, nd, featlen, mean = mv.read_gauden(os.path.join(d, "means"))
    arr = np.tile(mean.reshape(1, n_feat, nd, featlen[0]), (5, 1, 1, 1))
    mv.write_gauden(os.path.join(d, "means"), arr, featlen)
    _, _, _, _, var = mv.read_gauden(os.path.join(d, "variances"))
    mv.write_gauden(os.path.join(d, "variances"), np.tile(var.reshape(1, n_feat, nd, featlen[0]),
                                                          (5, 1, 1, 1)), featlen)
    with pytest.raises(ssb.SsbError):
        ssb.AcousticModel(d, device=-1)


# ------------------------------------------------------------------ GPU parity
@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_gpu_semi_dense_scores(semi_golden, semi_dirs, golden, tag):
    import soundswallower_b200 as ssb
    d, beam = semi_dirs[tag]
    m = ssb.AcousticModel(d, device=0, topn_beam=beam)
    feat = golden["en-us"]["feat"]
    dense = ssb.score_batch(m, [feat, feat[:7]])
    assert np.array_equal(dense[0][[0, 1, 100, len(feat) - 1]], semi_golden[tag + "_senscr_rows"])
    assert sha(dense[0]) == str(semi_golden[tag + "_senscr_sha"])
    assert np.array_equal(dense[1], dense[0][:7])
    m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("tag", TAGS)
def test_gpu_semi_alignment(semi_golden, semi_dirs, golden, tag, mode):
    import soundswallower_b200 as ssb
    d, beam = semi_dirs[tag]
    m = ssb.AcousticModel(d, device=0, topn_beam=beam)
    g = golden["en-us"]
    chain = chain_from_golden(g, windows=mode != "nowin")
    r = ssb.align_batch(m, [g["feat"]], [chain], compallsen=mode.endswith("call"),
                        want_chain_scr=True)[0]
    key = "%s_%s_" % (tag, mode)
    check_alignment(dict(rv=r["rv"], best_score=r["best_score"], start=r["start"], dur=r["dur"],
                         score=r["score"]), semi_golden, key)
    assert np.array_equal(r["chain_scr"], semi_golden[key + "chain_scr"])
    m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_gpu_semi_vtable_frame_eval(semi_dirs, golden, tag):
    """mgau_t drop-in on a semi-continuous model, frame by frame with active lists, against
    the oracle (which is pinned to the reference above)."""
    import soundswallower_b200 as ssb
    from oracle.oracle import Oracle
    d, beam = semi_dirs[tag]
    m = ssb.AcousticModel(d, device=0, topn_beam=beam)
    o = Oracle(d)
    if beam:
        o.set_topn_beam(beam)
    mg = ssb.PtmMgau(m)
    rs = np.random.RandomState(11)
    feat = golden["en-us"]["feat"][:24]
    p = o.new_ptm()
    for t in range(len(feat)):
        call = t % 5 == 4
        act = np.unique(rs.randint(0, m.n_sen, rs.randint(1, 400)))
        lst = ssb.flags2list(act, m.n_sen)
        got = mg.frame_eval(feat[t], t, senone_active=lst, compallsen=call)
        want = o.frame_eval(p, feat[t], t, active=lst, compallsen=call)
        assert np.array_equal(got, want), t
    o.free_ptm(p)
    m.close()
