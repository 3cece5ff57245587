"""Fully continuous acoustic models (ms_mgau / ms_gauden / ms_senone): one
codebook per senone, stateless float top-N, (int + 1023) >> 10, table log-add, int16 clamps,
best score subtracted over the evaluated senones.

No bundled model is continuous, so the models are synthetic (tests/model_variants.py: the
en-us mdef/transitions, one 39-dimensional stream, Gaussians drawn from the en-us codebooks,
seeded float mixture weights); tests/golden/cont_en-us.npz holds what the compiled reference
computed on them (tools/make_golden.py --cont).
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
def cont_golden():
    return np.load(os.path.join(GOLDEN, "cont_en-us.npz"))


@pytest.fixture(scope="module")
def cont_dirs(tmp_path_factory, golden):
    root = str(tmp_path_factory.mktemp("cont"))
    out = {}
    for tag, kw in mv.CONT_CASES:
        d = os.path.join(root, tag)
        mv.write_cont_model(model_dir("en-us"), d, int(golden["en-us"]["dims"][4]), **kw)
        out[tag] = d
    return out


TAGS = [c[0] for c in mv.CONT_CASES]
MODES = ["win", "nowin", "win_call"]


def check_alignment(r, g, key):
    st = g[key + "states"]
    assert r["rv"] == int(g[key + "rv"])
    assert r["best_score"] == int(g[key + "best"])
    assert np.array_equal(r["start"], st[:, 1]) and np.array_equal(r["dur"], st[:, 2])
    assert np.array_equal(r["score"], st[:, 3])


# ------------------------------------------------------------------ oracle (CPU)
@pytest.mark.parametrize("tag", TAGS)
def test_oracle_cont_scores(cont_golden, cont_dirs, golden, tag):
    from oracle.oracle import Oracle
    o = Oracle(cont_dirs[tag])
    assert o.n_mgau == o.n_sen
    a = o.model_arrays()
    assert sha(a["mixw"]) == str(cont_golden[tag + "_mixw_sha"])
    assert sha(a["det"]) == str(cont_golden[tag + "_det_sha"])
    feat = golden["en-us"]["feat"]
    dense = o.score_all(feat)
    assert np.array_equal(dense[[0, 1, 100, len(feat) - 1]], cont_golden[tag + "_senscr_rows"])
    assert sha(dense) == str(cont_golden[tag + "_senscr_sha"])


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("tag", TAGS)
def test_oracle_cont_alignment(cont_golden, cont_dirs, golden, tag, mode):
    from oracle.oracle import Oracle
    o = Oracle(cont_dirs[tag])
    g = golden["en-us"]
    chain = chain_from_golden(g, windows=mode != "nowin")
    r = o.state_align(g["feat"], chain["ssid"], chain["tmat"], chain["sf"], chain["ef"],
                      compallsen=mode.endswith("call"), want_tokens=True, want_senscr=True)
    key = "%s_%s_" % (tag, mode)
    check_alignment(r, cont_golden, key)
    assert sha(r["tokens"]) == str(cont_golden[key + "tokens_sha"])
    if mode == "win_call":
        chain_sen = o.model_arrays()["sseq"][chain["ssid"]].reshape(-1)
        assert np.array_equal(r["senscr"][:, chain_sen], cont_golden[key + "chain_scr"])


# ------------------------------------------------------------------ product, host side
@pytest.mark.parametrize("tag", TAGS)
def test_product_loads_cont_model(cont_golden, cont_dirs, tag):
    import soundswallower_b200 as ssb
    m = ssb.AcousticModel(cont_dirs[tag], device=-1)
    assert m.kind == 2 and m.n_mgau == m.n_sen and m.n_feat == 1 and m.blk == 39
    a = m.arrays()
    assert sha(a["mixw"]) == str(cont_golden[tag + "_mixw_sha"])   # [sen][feat][density]
    assert sha(a["det"]) == str(cont_golden[tag + "_det_sha"])
    m.close()


def test_explicit_senmgau_map_is_declined(cont_dirs, tmp_path):
    import shutil
    import soundswallower_b200 as ssb
    d = str(tmp_path / "mapped")
    shutil.copytree(cont_dirs["cont3"], d, symlinks=True)
    open(os.path.join(d, "senmgau"), "wb").write(b"s3\nendhdr\n")
    with pytest.raises(ssb.SsbError, match="senmgau"):
        ssb.AcousticModel(d, device=-1)


# ------------------------------------------------------------------ GPU parity
@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_gpu_cont_dense_scores(cont_golden, cont_dirs, golden, tag):
    import soundswallower_b200 as ssb
    m = ssb.AcousticModel(cont_dirs[tag], device=0)
    feat = golden["en-us"]["feat"]
    dense = ssb.score_batch(m, [feat, feat[:7], feat[:0]])
    assert np.array_equal(dense[0][[0, 1, 100, len(feat) - 1]], cont_golden[tag + "_senscr_rows"])
    assert sha(dense[0]) == str(cont_golden[tag + "_senscr_sha"])
    assert np.array_equal(dense[1], dense[0][:7]) and len(dense[2]) == 0
    m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("tag", TAGS)
def test_gpu_cont_alignment(cont_golden, cont_dirs, golden, tag, mode):
    import soundswallower_b200 as ssb
    m = ssb.AcousticModel(cont_dirs[tag], device=0)
    g = golden["en-us"]
    chain = chain_from_golden(g, windows=mode != "nowin")
    r = ssb.align_batch(m, [g["feat"], g["feat"][:150]], [chain, chain],
                        compallsen=mode.endswith("call"), want_chain_scr=True, want_tokens=True)[0]
    key = "%s_%s_" % (tag, mode)
    check_alignment(dict(rv=r["rv"], best_score=r["best_score"], start=r["start"], dur=r["dur"],
                         score=r["score"]), cont_golden, key)
    assert sha(np.ascontiguousarray(r["tokens"], np.int32)) == str(cont_golden[key + "tokens_sha"])
    if mode == "win_call":
        assert np.array_equal(r["chain_scr"], cont_golden[key + "chain_scr"])
    m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_gpu_cont_vtable_frame_eval(cont_dirs, golden, tag):
    """mgau_t drop-in ("ms") frame by frame against the oracle.  With an active list only the
    listed senones are written (the rest of the buffer keeps earlier values, as in the
    reference): the listed entries are compared."""
    import soundswallower_b200 as ssb
    from oracle.oracle import Oracle
    m = ssb.AcousticModel(cont_dirs[tag], device=0)
    o = Oracle(cont_dirs[tag])
    mg = ssb.PtmMgau(m)
    rs = np.random.RandomState(12)
    feat = golden["en-us"]["feat"][:16]
    p = o.new_ptm()
    for t in range(len(feat)):
        call = t % 5 == 4
        act = np.unique(rs.randint(0, m.n_sen, rs.randint(1, 400)))
        lst = ssb.flags2list(act, m.n_sen)
        listed = np.cumsum(lst.astype(np.int64))
        got = mg.frame_eval(feat[t], t, senone_active=lst, compallsen=call)
        want = o.frame_eval(p, feat[t], t, active=lst, compallsen=call)
        if call:
            assert np.array_equal(got, want), t
        else:
            assert np.array_equal(got[listed], want[listed]), t
    o.free_ptm(p)
    m.close()


@pytest.mark.gpu
def test_gpu_grammar_search_on_semi_and_continuous_scores(cont_dirs, golden, tmp_path):
    """The grammar search (first pass) is scorer-agnostic: with a semi-continuous and with a
    continuous model its history table equals the oracle's search over the oracle's scores."""
    import soundswallower_b200 as ssb
    from oracle.oracle import Oracle
    from test_oracle_fsg import graph_of
    fg = np.load(os.path.join(GOLDEN, "fsg_en-us.npz"))
    feat = golden["en-us"]["feat"]
    semi = str(tmp_path / "semi")
    mv.write_semi_model(model_dir("en-us"), semi, int(golden["en-us"]["dims"][4]), n_density=64, seed=21)
    for d in (cont_dirs["cont8"], semi):
        m, o = ssb.AcousticModel(d, device=0), Oracle(d)
        for name in ("align", "jsgf"):
            graph = graph_of(fg, name)
            r = ssb.fsg_batch(m, [feat, feat[:120]], [graph], want_hist=True)
            for f, got in zip((feat, feat[:120]), r):
                w = o.fsg_search(graph, o.score_all(f))
                assert got["rv"] == w["rv"] == 0
                assert np.array_equal(got["hist"], w["hist"])
                assert got["n_hmm_eval"] == w["n_hmm_eval"]
        m.close()
