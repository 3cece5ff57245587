"""The oracle (oracle/ss_oracle.c, our plain-C restatement) against the golden vectors
generated from the unmodified reference (tools/make_golden.py), against SURVEY.md
Appendix B, and -- where oracle/_ref/libssref.so is present -- against the reference
itself on fresh inputs.  CPU only."""
import hashlib

import numpy as np
import pytest

import os

from conftest import GOLDEN, chain_from_golden, model_dir, model_features, random_chain


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("lang", ["en-us", "fr-fr"])
def test_model_tables_match_reference(oracles, golden, lang):
    o, g = oracles(lang), golden[lang]
    assert [o.n_mgau, o.n_feat, o.n_density, o.veclen, o.n_sen, o.n_sseq, o.n_emit, o.n_tmat,
            o.n_ciphone, o.n_phone, o.sil] == g["dims"].tolist()
    for k, v in o.model_arrays().items():
        assert sha(v) == str(g["model_sha_" + k]), k


def test_logadd_table_is_survey_appendix_b(oracles):
    rv, lut = oracles("en-us").logadd_table8(1.0001, 10)
    want = [7, 6, 6, 5, 5, 5, 4, 4, 4, 3, 3, 3, 3, 2, 2, 2, 2, 2] + [1] * 11
    assert lut[:29].tolist() == want and not lut[29:].any()


def test_en_us_shapes_are_survey_section_8(oracles):
    o = oracles("en-us")
    assert (o.n_mgau, o.n_feat, o.n_density, o.veclen, o.n_sen, o.n_emit) == (42, 3, 128, 13, 5126, 3)
    f = oracles("fr-fr")
    assert (f.n_mgau, f.n_sen, f.n_sseq) == (36, 2108, 7011)


@pytest.mark.parametrize("lang", ["en-us", "fr-fr"])
def test_dense_scores_match_reference(oracles, golden, lang):
    o, g = oracles(lang), golden[lang]
    d = o.score_all(g["feat"])
    assert sha(d) == str(g["senscr_sha"])
    assert np.array_equal(d[g["senscr_rows"]], g["senscr_sample"])
    assert (d.min(1) == 0).all()


def test_dense_scores_survey_sha(oracles, golden):
    d = oracles("en-us").score_all(golden["en-us"]["feat"])
    assert d.shape == (278, 5126)
    assert sha(d) == "4129ae8da103a1aa3a6b44487596c1bb3563a84cc5929fc82c1fc9068951f26f"
    assert d[0, :12].tolist() == [33, 64, 27, 92, 77, 34, 94, 104, 157, 80, 93, 161]
    assert int(d[100].argmin()) == 4948


@pytest.mark.parametrize("lang", ["en-us", "fr-fr"])
def test_topn_head_matches_reference(oracles, golden, lang):
    o, g = oracles(lang), golden[lang]
    p = o.new_ptm()
    try:
        for t in range(g["topn_norm_head"].shape[0]):
            _, tn = o.frame_eval(p, g["feat"][t], t, compallsen=True, want_topn=True)
            assert np.array_equal(tn, g["topn_norm_head"][t])
            o.set_frame_idx(p, t + 1)  # acmod_advance (ref: src/acmod.c:760)
    finally:
        o.free_ptm(p)


@pytest.mark.parametrize("lang", ["en-us", "fr-fr"])
@pytest.mark.parametrize("mode", ["win", "nowin", "win_call"])
def test_state_align_matches_reference(oracles, golden, lang, mode):
    o, g = oracles(lang), golden[lang]
    chain = chain_from_golden(g, windows=mode != "nowin")
    r = o.state_align(g["feat"], chain["ssid"], chain["tmat"], chain["sf"], chain["ef"],
                      compallsen=mode.endswith("call"), want_tokens=True, want_senscr=True)
    st = g[mode + "_states"]
    assert r["rv"] == int(g[mode + "_rv"]) == 0
    assert r["best_score"] == int(g[mode + "_best"])
    assert np.array_equal(r["start"], st[:, 1])
    assert np.array_equal(r["dur"], st[:, 2])
    assert np.array_equal(r["score"], st[:, 3])
    assert sha(r["tokens"]) == str(g[mode + "_tokens_sha"])
    assert sha(r["senscr"]) == str(g[mode + "_senscr_sha"])
    ps, pd, pc = o.propagate(r["start"], r["dur"], r["score"])
    ph = g[mode + "_phones"]
    assert np.array_equal(ps, ph[:, 3]) and np.array_equal(pd, ph[:, 4]) and np.array_equal(pc, ph[:, 5])


def test_two_pass_alignment_is_survey_appendix_b(golden):
    g = golden["en-us"]
    assert int(g["hyp_score"]) == -2761 and int(g["n_frames"]) == 279
    assert g["segs"][:, 1:].tolist() == [[0, 45, -230, -337], [46, 63, -133, 0], [64, 116, -369, 0],
                                        [117, 152, -468, 0], [153, 210, -563, 0], [211, 277, -324, -337]]
    st = g["states"]
    assert st[:4, :4].tolist() == [[96, 0, 44, 0], [97, 44, 1, -37], [98, 45, 1, -30], [2085, 46, 3, -18]]
    assert st[-1, :4].tolist() == [98, 273, 5, -62]
    assert int(golden["fr-fr"]["hyp_score"]) == -4236
    # the CLI's second pass equals the windowed chain alignment
    assert np.array_equal(g["states"], g["win_states"])


def test_hmm_eval_known_answers(oracles, synthetic):
    o, s = oracles("en-us"), synthetic
    tp = o.model_arrays()["tp"]
    for i in range(len(s["hmm_best"])):
        senscr = np.zeros(o.n_sen, np.int16)
        senscr[s["hmm_senid"][i]] = s["hmm_senscr3"][i]
        # duplicate senone ids inside one HMM would make the scatter ambiguous
        if len(set(s["hmm_senid"][i].tolist())) < 3:
            continue
        best, st = o.hmm_eval(3, tp[s["hmm_tmat"][i]], s["hmm_senid"][i], senscr, s["hmm_st_in"][i])
        assert best == int(s["hmm_best"][i]), i
        assert np.array_equal(st, s["hmm_st_out"][i]), i


def test_synthetic_scores(oracles, synthetic):
    o = oracles("en-us")
    for u in range(4):
        d = o.score_all(synthetic["feat%d" % u])
        assert sha(d) == str(synthetic["senscr_sha%d" % u])


def test_flags2list_bridges_gaps(oracles):
    o = oracles("en-us")
    lst = o.flags2list([3, 10, 700, 701, 5000])
    # 10 -> 700 is a gap of 690 = 255 + 255 + 180; 701 -> 5000 is 16 * 255 + 219
    assert lst[:2].tolist() == [3, 7] and lst[2:5].tolist() == [255, 255, 180] and lst[5] == 1
    assert int(lst.astype(np.int64).sum()) == 5000
    import soundswallower_b200 as ssb
    assert np.array_equal(ssb.flags2list([3, 10, 700, 701, 5000]), lst)


# ---- live comparison with the compiled reference, where it is available -----------------
def _ref(lang, **kw):
    from oracle import refshim
    if not refshim.available():
        pytest.skip("oracle/_ref/libssref.so not built")
    return refshim.Ref(model_dir(lang), **kw)


def test_oracle_vs_reference_random_features(oracles):
    o = oracles("en-us")
    ref = _ref("en-us", compallsen=True)
    rs = np.random.RandomState(7)
    x = model_features(rs, o.model_arrays(), 25)
    assert np.array_equal(o.score_all(x), ref.score_all(x))
    ref.close()


def test_oracle_vs_reference_frontend_golden(golden):
    """The committed feature fixture is what the reference's frontend produces here."""
    import os
    from conftest import DATA
    ref = _ref("en-us")
    pcm = np.fromfile(os.path.join(DATA, "goforward.raw"), np.int16)
    assert np.array_equal(ref.features_from_pcm(pcm), golden["en-us"]["feat"])
    ref.close()


@pytest.mark.parametrize("E", [5, 3])
def test_hmm_eval_on_given_transition_matrices(oracles, E):
    """hmm_vit_eval_5st_lr (and _3st_lr) known answers of the reference on random left-to-right
    transition matrices with impossible arcs, clamps and ties (tests/golden/hmm5.npz): the
    bundled models only have 3-state HMMs."""
    o = oracles("en-us")
    g = np.load(os.path.join(GOLDEN, "hmm5.npz"))
    for i in range(len(g["best%d" % E])):
        best, st = o.hmm_eval(E, g["tp%d" % E][i], np.arange(E, dtype=np.uint16), g["senscr%d" % E][i],
                              g["st_in%d" % E][i])
        assert best == int(g["best%d" % E][i]), i
        assert np.array_equal(st, g["st_out%d" % E][i]), i
