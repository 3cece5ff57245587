"""Parity of the CUDA path (through the C ABI) with the oracle and the golden vectors.
Integer work is compared bit for bit."""
import hashlib
import os

import numpy as np
import pytest

import soundswallower_b200 as ssb
from conftest import chain_from_golden, model_features, random_chain

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ------------------------------------------------------------------ K1: Gaussian top-N
@pytest.mark.parametrize("lang", ["en-us", "fr-fr"])
def test_topn_matches_oracle_golden_audio(models, oracles, golden, lang):
    m, o, g = models(lang), oracles(lang), golden[lang]
    cw, sc = ssb.topn_batch(m, [g["feat"]])
    ocw, osc = o.topn_all(g["feat"])
    assert np.array_equal(sc[0], osc)
    assert np.array_equal(cw[0], ocw)


def test_topn_ragged_batch(models, oracles):
    m, o = models("en-us"), oracles("en-us")
    rs = np.random.RandomState(3)
    arrays = o.model_arrays()
    lens = [1, 0, 17, 5, 64, 2, 33] + [3] * 130  # > one CTA of utterances, an empty one
    feats = [model_features(rs, arrays, T) for T in lens]
    cw, sc = ssb.topn_batch(m, feats)
    for u in (0, 1, 2, 4, 6, 7, 100, len(lens) - 1):
        ocw, osc = o.topn_all(feats[u])
        assert np.array_equal(sc[u], osc) and np.array_equal(cw[u], ocw), u


@pytest.mark.parametrize("topn", [1, 2, 3])
def test_topn_other_n(models, oracles, golden, topn):
    from oracle.oracle import Oracle
    from conftest import model_dir
    m = models("en-us", topn=topn)
    o = Oracle(model_dir("en-us"), topn=topn)
    feat = golden["en-us"]["feat"][:40]
    cw, sc = ssb.topn_batch(m, [feat])
    ocw, osc = o.topn_all(feat)
    assert np.array_equal(sc[0], osc) and np.array_equal(cw[0], ocw)
    assert np.array_equal(ssb.score_batch(m, [feat])[0], o.score_all(feat))


# ------------------------------------------------------------------ K1+K2 dense scores
@pytest.mark.parametrize("lang", ["en-us", "fr-fr"])
def test_dense_scores_golden(models, golden, lang):
    m, g = models(lang), golden[lang]
    d = ssb.score_batch(m, [g["feat"]])[0]
    assert sha(d) == str(g["senscr_sha"])
    assert np.array_equal(d[g["senscr_rows"]], g["senscr_sample"])


def test_dense_scores_survey_sha(models, golden):
    d = ssb.score_batch(models("en-us"), [golden["en-us"]["feat"]])[0]
    assert sha(d) == "4129ae8da103a1aa3a6b44487596c1bb3563a84cc5929fc82c1fc9068951f26f"


def test_dense_scores_synthetic(models, synthetic):
    m = models("en-us")
    feats = [synthetic["feat%d" % u] for u in range(4)]
    out = ssb.score_batch(m, feats)
    for u in range(4):
        assert sha(out[u]) == str(synthetic["senscr_sha%d" % u]), u
    # batch composition must not matter
    out2 = ssb.score_batch(m, feats[::-1])
    for u in range(4):
        assert np.array_equal(out2[3 - u], out[u])


def test_dense_scores_spanning_slabs(models, oracles):
    """More frames than one dense slab (32768): only properties + spot rows vs the oracle."""
    m, o = models("fr-fr"), oracles("fr-fr")
    rs = np.random.RandomState(5)
    base = model_features(rs, o.model_arrays(), 700)
    feats = [base[rs.randint(0, 600):][:100] for _ in range(340)]  # 34000 frames
    out = ssb.score_batch(m, feats)
    for u in (0, 327, 328, 339):
        assert np.array_equal(out[u], o.score_all(feats[u])), u
    assert all((d.min(1) == 0).all() for d in out)


# ------------------------------------------------------------------ mgau vtable (one frame per call)
def test_mgau_vtable_compallsen_and_rewind(models, oracles, golden):
    m, o, g = models("en-us"), oracles("en-us"), golden["en-us"]
    mg = ssb.PtmMgau(m)
    assert mg.name == "ptm" and mg.frame_idx == 0
    p = o.new_ptm()
    feat = g["feat"]
    for t in range(12):
        a = mg.frame_eval(feat[t], t, compallsen=True)
        b = o.frame_eval(p, feat[t], t, compallsen=True)
        assert np.array_equal(a, b), t
        mg.frame_idx = t + 1
        o.set_frame_idx(p, t + 1)
    # re-scoring a frame that is already in the history does not recompute the codebooks
    a = mg.frame_eval(feat[11], 11, compallsen=True)
    b = o.frame_eval(p, feat[11], 11, compallsen=True)
    assert np.array_equal(a, b)
    # acmod_rewind: frame_idx back to 0, history kept (ref: src/acmod.c:730-751)
    mg.frame_idx = 0
    o.set_frame_idx(p, 0)
    for t in range(3):
        assert np.array_equal(mg.frame_eval(feat[t], t, compallsen=True),
                              o.frame_eval(p, feat[t], t, compallsen=True))
        mg.frame_idx = t + 1
        o.set_frame_idx(p, t + 1)
    mg.reset()
    o.reset_ptm(p)
    o.set_frame_idx(p, 0)
    assert np.array_equal(mg.frame_eval(feat[5], 0, compallsen=True),
                          o.frame_eval(p, feat[5], 0, compallsen=True))
    o.free_ptm(p)
    mg.close()


def test_mgau_vtable_active_lists(models, oracles, golden):
    m, o, g = models("en-us"), oracles("en-us"), golden["en-us"]
    mg = ssb.PtmMgau(m)
    p = o.new_ptm()
    rs = np.random.RandomState(9)
    feat = g["feat"]
    sen = set()
    for t in range(20):
        # a growing active set with gaps above 255 (bridging entries)
        sen |= set(int(x) for x in rs.randint(0, m.n_sen, 6))
        lst = ssb.flags2list(sen)
        assert np.array_equal(lst, o.flags2list(sorted(sen)))
        a = mg.frame_eval(feat[t], t, senone_active=lst, compallsen=False)
        b = o.frame_eval(p, feat[t], t, active=lst, compallsen=False)
        assert np.array_equal(a, b), t
        mg.frame_idx = t + 1
        o.set_frame_idx(p, t + 1)
    # empty active list: everything 0 - INT_MAX truncated, as the reference computes it
    a = mg.frame_eval(feat[20], 20, senone_active=np.zeros(0, np.uint8), compallsen=False)
    b = o.frame_eval(p, feat[20], 20, active=np.zeros(0, np.uint8), compallsen=False)
    assert np.array_equal(a, b)
    o.free_ptm(p)
    mg.close()


# ------------------------------------------------------------------ hmm_vit_eval
def test_hmm_vit_eval_known_answers(models, synthetic):
    m, s = models("en-us"), synthetic
    for i in range(0, len(s["hmm_best"]), 3):
        if len(set(s["hmm_senid"][i].tolist())) < 3:
            continue
        senscr = np.zeros(m.n_sen, np.int16)
        senscr[s["hmm_senid"][i]] = s["hmm_senscr3"][i]
        best, st = ssb.hmm_vit_eval(m, s["hmm_tmat"][i], s["hmm_senid"][i], senscr, s["hmm_st_in"][i])
        assert best == int(s["hmm_best"][i]), i
        assert np.array_equal(st, s["hmm_st_out"][i]), i


@pytest.mark.parametrize("E", [5, 3])
def test_hmm_step_on_given_transition_matrices(models, E):
    """hmm_step5 / hmm_step3 (the device evaluators K3 and K4 are built on) against the
    reference's hmm_vit_eval_5st_lr / _3st_lr known answers on random left-to-right transition
    matrices (tests/golden/hmm5.npz): no bundled model has 5-state HMMs."""
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "hmm5.npz"))
    best, st = ssb.hmm_vit_eval_tp(models("en-us"), g["tp%d" % E], g["senscr%d" % E], g["st_in%d" % E])
    assert np.array_equal(best, g["best%d" % E])
    assert np.array_equal(st, g["st_out%d" % E])


# ------------------------------------------------------------------ chain alignment
def _check_against_golden(r, g, mode):
    st = g[mode + "_states"]
    assert r["rv"] == 0 and r["best_score"] == int(g[mode + "_best"])
    assert np.array_equal(r["start"], st[:, 1])
    assert np.array_equal(r["dur"], st[:, 2])
    assert np.array_equal(r["score"], st[:, 3])
    assert np.array_equal(r["chain_scr"], g[mode + "_chain_scr"])
    assert sha(r["tokens"]) == str(g[mode + "_tokens_sha"])


@pytest.mark.parametrize("k1", ["", "tc2"])
@pytest.mark.parametrize("lang", ["en-us", "fr-fr"])
@pytest.mark.parametrize("mode", ["win", "nowin", "win_call"])
def test_align_golden(models, golden, lang, mode, k1, monkeypatch):
    if k1:
        monkeypatch.setenv("SSB_K1", k1)
    """Windowed / unwindowed, default (active-list) / compallsen scoring, vs the reference's
    own state_align_search run on the same features (tools/make_golden.py)."""
    m, g = models(lang), golden[lang]
    chain = chain_from_golden(g, windows=mode != "nowin")
    r = ssb.align_batch(m, [g["feat"]], [chain], compallsen=mode.endswith("call"),
                        want_chain_scr=True, want_tokens=True)[0]
    _check_against_golden(r, g, mode)
    ps, pd, pc = ssb.propagate(r["start"], r["dur"], r["score"], m.n_emit)
    ph = g[mode + "_phones"]
    assert np.array_equal(ps, ph[:, 3]) and np.array_equal(pd, ph[:, 4]) and np.array_equal(pc, ph[:, 5])


def test_align_reproduces_survey_appendix_b(models, golden):
    m, g = models("en-us"), golden["en-us"]
    r = ssb.align_batch(m, [g["feat"]], [chain_from_golden(g)])[0]
    sen = m.arrays()["sseq"][g["phones"][:, 1]].reshape(-1)
    got = ["%d:%d+%d:%d" % (sen[i], r["start"][i], r["dur"][i], r["score"][i]) for i in range(len(sen))]
    assert got[:6] == ["96:0+44:0", "97:44+1:-37", "98:45+1:-30", "2085:46+3:-18", "2115:49+3:-18",
                       "2138:52+2:-13"]
    assert got[-3:] == ["96:211+61:-296", "97:272+1:-21", "98:273+5:-62"]


def _random_batch(rs, o, n_utts, max_T=60, max_ph=14):
    arrays = o.model_arrays()
    feats, chains = [], []
    for u in range(n_utts):
        T = int(rs.randint(1, max_T))
        npn = int(rs.randint(1, max_ph))
        feats.append(model_features(rs, arrays, T))
        chains.append(random_chain(rs, o, npn, T, windowed=u % 3 != 0))
    return feats, chains


@pytest.mark.parametrize("k1", ["", "tc2"])
@pytest.mark.parametrize("compallsen", [False, True])
def test_align_random_ragged_batch(models, oracles, compallsen, k1, monkeypatch):
    # (k1: the kernel a small batch gets by default -- frame-tiled -- and the big batches' tc2)
    if k1:
        monkeypatch.setenv("SSB_K1", k1)
    m, o = models("en-us"), oracles("en-us")
    rs = np.random.RandomState(21 + compallsen)
    feats, chains = _random_batch(rs, o, 40)
    # edge cases: zero frames, one frame, chain longer than the utterance
    feats[3] = feats[3][:0]
    feats[5] = feats[5][:1]
    chains[7] = random_chain(rs, o, 30, feats[7].shape[0], windowed=False)
    res = ssb.align_batch(m, feats, chains, compallsen=compallsen, want_chain_scr=True,
                          want_tokens=True)
    n_ok = 0
    for u, (f, c, r) in enumerate(zip(feats, chains, res)):
        w = o.state_align(f, c["ssid"], c["tmat"], c["sf"], c["ef"], compallsen=compallsen,
                          want_tokens=True, want_senscr=True)
        assert r["rv"] == w["rv"], u
        if f.shape[0]:
            assert r["best_score"] == w["best_score"], u
            sen = o.model_arrays()["sseq"][c["ssid"]].reshape(-1)
            assert np.array_equal(r["chain_scr"], w["senscr"][:, sen]), u
            assert np.array_equal(r["tokens"], w["tokens"]), u
        if w["rv"] == 0:
            n_ok += 1
            assert np.array_equal(r["start"], w["start"]) and np.array_equal(r["dur"], w["dur"]), u
            assert np.array_equal(r["score"], w["score"]), u
    assert n_ok >= 10


@pytest.mark.parametrize("compallsen", [False, True])
def test_pipeline_equals_oracle_and_single_batch(models, oracles, compallsen):
    """ssb_pipeline_align: chunks of whole utterances through 3 lanes (own stream + host thread
    each) give, utterance by utterance, the oracle's answer and the one-batch answer, debug
    outputs (chain scores, token stacks) and init_active slices included; lanes are reused
    across calls."""
    m, o = models("en-us"), oracles("en-us")
    rs = np.random.RandomState(77 + compallsen)
    feats, chains = _random_batch(rs, o, 37)
    feats[0] = feats[0][:0]
    feats[36] = feats[36][:1]
    init = [sorted(set(int(x) for x in rs.randint(0, m.n_sen, size=rs.randint(0, 5))))
            for _ in range(37)]
    one = ssb.align_batch(m, feats, chains, init_active=init, compallsen=compallsen,
                          want_chain_scr=True, want_tokens=True)
    pipe = ssb.AlignPipeline(m, n_lanes=3, chunk_frames=150)
    for rep in range(2):
        pipe.upload(feats, chains, init_active=init, compallsen=compallsen)
        got = pipe.per_utt(pipe.align(want_chain_scr=True, want_tokens=True))
        assert pipe.n_chunks() >= 6 and pipe.n_launches() >= 4 * pipe.n_chunks()
        for u, (a, b) in enumerate(zip(got, one)):
            assert a["rv"] == b["rv"] and a["best_score"] == b["best_score"], u
            assert a["n_renorm"] == b["n_renorm"], u
            for k in ("start", "dur", "score", "chain_scr", "tokens"):
                assert np.array_equal(a[k], b[k]), (u, k)
    pipe.close()
    n_ok = 0
    for u, (f, c, r) in enumerate(zip(feats, chains, got)):
        w = o.state_align(f, c["ssid"], c["tmat"], c["sf"], c["ef"], compallsen=compallsen,
                          init_active=init[u], want_tokens=True)
        assert r["rv"] == w["rv"], u
        if f.shape[0]:
            assert np.array_equal(r["tokens"], w["tokens"]), u
        if w["rv"] == 0:
            n_ok += 1
            assert np.array_equal(r["start"], w["start"]) and np.array_equal(r["dur"], w["dur"]), u
            assert np.array_equal(r["score"], w["score"]), u
    assert n_ok >= 8


def test_pipeline_stream_of_batches(models, oracles):
    """ssb_pipeline_submit / collect: three different batches in flight through two lanes, whole
    batches as chunks, kernels kept apart; every batch equals its own one-shot result."""
    m, o = models("en-us"), oracles("en-us")
    rs = np.random.RandomState(9)
    batches = [_random_batch(rs, o, n) for n in (11, 1, 17, 6)]
    pipe = ssb.AlignPipeline(m, n_lanes=2, chunk_frames=1 << 40, overlap_kernels=False)
    tickets = []
    for feats, chains in batches:
        pipe.upload(feats, chains)
        tickets.append((pipe.submit(want_chain_scr=True), pipe.n_utts, pipe.frame_off, pipe.phone_off))
    for (t, n_utts, fo, po), (feats, chains) in zip(tickets, batches):
        res = pipe.collect(t)
        assert pipe.n_chunks() == 1
        pipe.n_utts, pipe.frame_off, pipe.phone_off = n_utts, fo, po
        got = pipe.per_utt(res)
        one = ssb.align_batch(m, feats, chains, want_chain_scr=True)
        for u, (a, b) in enumerate(zip(got, one)):
            assert a["rv"] == b["rv"] and a["best_score"] == b["best_score"], u
            for k in ("start", "dur", "score", "chain_scr"):
                assert np.array_equal(a[k], b[k]), (u, k)
    # an error in one batch is reported by its collect and does not poison the next one
    feats, chains = batches[0]
    bad = [dict(c) for c in chains]
    bad[3] = dict(bad[3], ssid=bad[3]["ssid"] + 10 ** 6)
    pipe.upload(feats, bad)
    t_bad = pipe.submit()
    pipe.upload(feats, chains)
    t_ok = pipe.submit()
    with pytest.raises(ssb.SsbError, match="out of range"):
        pipe.collect(t_bad)
    assert pipe.collect(t_ok)["rv"].shape == (len(feats),)
    pipe.close()


def test_align_batch_routes_large_batches_through_pipeline(models, oracles, monkeypatch):
    m, o = models("en-us"), oracles("en-us")
    rs = np.random.RandomState(5)
    feats, chains = _random_batch(rs, o, 20)
    one = ssb.align_batch(m, feats, chains)
    monkeypatch.setenv("SSB_PIPE_CHUNK_FRAMES", "64")
    monkeypatch.setenv("SSB_PIPE_LANES", "2")
    got = ssb.align_batch(m, feats, chains)
    for a, b in zip(got, one):
        assert a["rv"] == b["rv"] and a["best_score"] == b["best_score"]
        for k in ("start", "dur", "score"):
            assert np.array_equal(a[k], b[k])


def test_align_init_active_senones(models, oracles, golden):
    """Pass 2 starts with whatever pass 1 left flagged (ref: src/state_align_search.c:186-188)."""
    m, o, g = models("en-us"), oracles("en-us"), golden["en-us"]
    chain = chain_from_golden(g)
    left = [5, 300, 301, 2999, 5100]
    r = ssb.align_batch(m, [g["feat"]], [chain], init_active=[left], want_chain_scr=True)[0]
    w = o.state_align(g["feat"], chain["ssid"], chain["tmat"], chain["sf"], chain["ef"],
                      init_active=left, want_senscr=True)
    sen = o.model_arrays()["sseq"][chain["ssid"]].reshape(-1)
    assert np.array_equal(r["chain_scr"], w["senscr"][:, sen])
    assert np.array_equal(r["start"], w["start"]) and np.array_equal(r["score"], w["score"])


def test_align_states_off_path_keep_caller_values(models, golden):
    m, g = models("en-us"), golden["en-us"]
    chain = chain_from_golden(g)
    b = ssb.StateAlignBatch(m)
    # second utterance cannot reach its final state: 2 frames for 18 phones
    b.upload([g["feat"], g["feat"][:2]], [chain, chain_from_golden(g, windows=False)])
    b.run()
    ns = 18 * 3
    init = (np.full(2 * ns, 777, np.int32), np.full(2 * ns, 888, np.int32), np.full(2 * ns, 999, np.int32))
    res = b.per_utt(b.download(init=init))
    assert res[0]["rv"] == 0 and np.array_equal(res[0]["dur"], g["win_states"][:, 2])
    assert res[1]["rv"] == -1 and (res[1]["dur"] == 888).all() and (res[1]["start"] == 777).all()
    # gmm_topn_tc2, senone_mix, chain_viterbi, backtrace (+ pack_features with SSB_K1_PACK=1,
    # + topn_fixup when segmented)
    # (the frame-tiled K1 always runs the tie fix-up: 5 launches)
    # (small batches take the frame-tiled K1 by default)
    ft = os.environ.get("SSB_K1", "") in ("", "ft") and not os.environ.get("SSB_K1_SEG")
    assert b.n_launches() == (5 if ft else 4 + bool(os.environ.get("SSB_K1_SEG")) + (os.environ.get("SSB_K1_PACK") == "1"))
    ms = b.kernel_ms()
    assert ms["total"] > 0
    st = b.stats()
    assert st["frames"] == 280 and st["state_frames"] == 280 * ns
    b.close()


def test_align_empty_batch(models):
    m = models("en-us")
    assert ssb.align_batch(m, [], []) == []
    assert ssb.score_batch(m, []) == []


def test_align_rejects_bad_chains(models, golden):
    m, g = models("en-us"), golden["en-us"]
    chain = chain_from_golden(g)
    bad = dict(chain, ssid=chain["ssid"].copy())
    bad["ssid"][3] = m.n_sseq
    with pytest.raises(ssb.SsbError, match="out of range"):
        ssb.align_batch(m, [g["feat"]], [bad])
    bad = dict(chain, ef=chain["ef"].copy())
    bad["ef"][4] = 10
    with pytest.raises(ssb.SsbError, match="must not decrease"):
        ssb.align_batch(m, [g["feat"]], [bad])


# ------------------------------------------------------------------ BASELINE-size properties
def test_config2_shape_properties(models, golden):
    """BASELINE config #2 shape (1000-frame utterances, 52 phones / 156 states), 512 utterances:
    size-independent invariants + identical utterances give identical alignments wherever
    they sit in the batch + spot parity."""
    from bench import make_config2_batch
    m, g = models("en-us"), golden["en-us"]
    feats, chains = make_config2_batch(g, n_utts=512, noise=0.05, seed=1234)
    feats[400] = feats[7].copy()
    b = ssb.StateAlignBatch(m)
    b.upload(feats, chains)
    b.run()
    res = b.per_utt(b.download())
    b.close()
    for u, r in enumerate(res):
        assert r["rv"] == 0, u
        on = r["dur"] > 0
        start, dur = r["start"][on], r["dur"][on]
        assert start[0] == 0 and (start[1:] == start[:-1] + dur[:-1]).all(), u  # tiles, monotone
        assert start[-1] + dur[-1] == 1000, u
        # every state lies inside its word window
        sf = np.repeat(chains[u]["sf"], 3)[on]
        ef = np.repeat(chains[u]["ef"], 3)[on]
        assert (start >= sf).all() and (start + dur <= ef).all(), u
    assert np.array_equal(res[400]["start"], res[7]["start"])
    assert np.array_equal(res[400]["score"], res[7]["score"])


def test_config2_spot_parity_with_oracle(models, oracles, golden):
    from bench import make_config2_batch
    m, o, g = models("en-us"), oracles("en-us"), golden["en-us"]
    feats, chains = make_config2_batch(g, n_utts=160, noise=0.05, seed=99)
    res = ssb.align_batch(m, feats, chains)
    for u in (0, 31, 32, 127, 128, 159):
        c = chains[u]
        w = o.state_align(feats[u], c["ssid"], c["tmat"], c["sf"], c["ef"])
        assert w["rv"] == 0 == res[u]["rv"]
        assert np.array_equal(res[u]["start"], w["start"]) and np.array_equal(res[u]["dur"], w["dur"])
        assert np.array_equal(res[u]["score"], w["score"]) and res[u]["best_score"] == w["best_score"]


# ------------------------------------------------------------------ tensor-core screening
def _exact_dist64(arrays, feat):
    """det - sum (x - mu)^2 v in float64: [T][mgau][feat][density]."""
    mean, var, det = (arrays[k].astype(np.float64) for k in ("mean", "var", "det"))
    x = feat.astype(np.float64).reshape(feat.shape[0], 1, mean.shape[1], 1, mean.shape[3])
    return det[None] - (((x - mean[None]) ** 2) * var[None]).sum(-1)


@pytest.mark.parametrize("lang", ["en-us", "fr-fr"])
def test_tc_screening_error_bound_holds(models, oracles, golden, lang):
    """|TF32 GEMM score - exact score| <= eps for EVERY (frame, codebook, stream, density):
    this inequality is what makes skipping non-survivors exact."""
    m, o, g = models(lang), oracles(lang), golden[lang]
    rs = np.random.RandomState(17)
    feats = [g["feat"][:96], model_features(rs, o.model_arrays(), 64, noise=1.5),
             (g["feat"][100:140] * 3.0).astype(np.float32)]  # far-off frames: large magnitudes
    cw, sc, approx, eps, cnt = ssb.tc_probe(m, feats)
    feat = np.concatenate(feats)
    exact = _exact_dist64(o.model_arrays(), feat)
    err = np.abs(approx.astype(np.float64) - exact)
    ratio = err / eps
    assert np.isfinite(approx).all() and (eps > 0).all()
    assert ratio.max() <= 1.0, ratio.max()
    # the bound is not vacuous: a few output units (1024 raw) at most on real audio
    assert np.median(cnt["eps_regular"][:96]) < 1024
    assert cnt["hot"].mean() < 0.05  # hot densities are the exception
    # and the probe's top-N are the oracle's
    off = 0
    for f in feats:
        ocw, osc = o.topn_all(f)
        assert np.array_equal(sc[off:off + len(f)], osc) and np.array_equal(cw[off:off + len(f)], ocw)
        off += len(f)
    # screening really prunes: far fewer exact evaluations than a full scan
    assert cnt["scan_steps"] == len(feat) * m.n_mgau * m.n_feat
    assert cnt["exact_evals"] < 0.25 * cnt["scan_steps"] * m.n_density


def test_tc_and_fp32_kernels_agree(models, oracles, golden, monkeypatch):
    """SSB_K1=fp32 selects the plain CUDA-core scan; both kernels give identical lists."""
    m, o, g = models("en-us"), oracles("en-us"), golden["en-us"]
    rs = np.random.RandomState(23)
    feats = [g["feat"][:50]] + [model_features(rs, o.model_arrays(), int(rs.randint(1, 30))) for _ in range(140)]
    cw_tc, sc_tc = ssb.topn_batch(m, feats)
    monkeypatch.setenv("SSB_K1", "fp32")
    cw_fp, sc_fp = ssb.topn_batch(m, feats)
    for a, b in zip(sc_tc, sc_fp):
        assert np.array_equal(a, b)
    for a, b in zip(cw_tc, cw_fp):
        assert np.array_equal(a, b)


# ------------------------------------------------------------------ long chains (config #4 shape)
def _tiled_fr(golden, reps):
    """goforward_fr tiled `reps` times: features, chain (14 phones per sentence + final SIL),
    word windows shifted per repetition (the long-form read-along shape of BASELINE config #4)."""
    g = golden["fr-fr"]
    feat, words, phones = g["feat"], g["words"], g["phones"]
    T1 = feat.shape[0]
    ssid, tmat, ws, wd = [], [], [], []
    for k in range(reps):
        for i in range(len(phones) - 1):
            w = int(phones[i, 6])
            s, d = int(words[w, 1]) + k * T1, int(words[w, 2])
            if w == 0 and k > 0:
                s = int(words[-1, 1]) + (k - 1) * T1
                d = k * T1 + int(words[0, 2]) - s
            ssid.append(int(phones[i, 1])); tmat.append(int(phones[i, 2])); ws.append(s); wd.append(d)
    s = int(words[-1, 1]) + (reps - 1) * T1
    ssid.append(int(phones[-1, 1])); tmat.append(int(phones[-1, 2])); ws.append(s); wd.append(reps * T1 - s)
    sf, ef = ssb.windows(np.array(ws, np.int32), np.array(wd, np.int32))
    x = np.concatenate([feat] * reps)
    rs = np.random.RandomState(777)
    x = (x + rs.normal(0, 0.05, x.shape)).astype(np.float32)
    return x, dict(ssid=np.array(ssid, np.int32), tmat=np.array(tmat, np.int32), sf=sf, ef=ef)


@pytest.mark.parametrize("windowed", [True, False])
def test_long_chain_multiwarp(models, oracles, golden, windowed):
    """169 phones / 507 states (> 128 phones: the multi-warp Viterbi kernel), 2868 frames."""
    m, o = models("fr-fr"), oracles("fr-fr")
    x, chain = _tiled_fr(golden, 12)
    if not windowed:
        chain = dict(chain, sf=chain["sf"] * 0, ef=chain["ef"] * 0 + ssb.INT_MAX)
    r = ssb.align_batch(m, [x], [chain], want_chain_scr=True)[0]
    w = o.state_align(x, chain["ssid"], chain["tmat"], chain["sf"], chain["ef"], want_senscr=True)
    sen = o.model_arrays()["sseq"][chain["ssid"]].reshape(-1)
    assert np.array_equal(r["chain_scr"], w["senscr"][:, sen])
    assert r["rv"] == w["rv"] == 0 and r["best_score"] == w["best_score"]
    for k in ("start", "dur", "score"):
        assert np.array_equal(r[k], w[k]), k
    on = r["dur"] > 0
    assert r["start"][on][0] == 0 and (r["start"][on][1:] == (r["start"][on] + r["dur"][on])[:-1]).all()


def test_very_long_chain_state_in_global_memory(models, oracles):
    """6600 phones: more HMM state than fits in shared memory (spill path), 4 warps, windows of a
    few frames per word.  Random triphones: the alignment may or may not reach the final state;
    either way every score and the verdict must equal the oracle's."""
    m, o = models("fr-fr"), oracles("fr-fr")
    rs = np.random.RandomState(4242)
    n_ph, T = 6600, 7000
    ssid_t, tmat_t, _ = o.phone_table()
    pid = rs.randint(0, len(ssid_t), n_ph)
    word_of = np.arange(n_ph) // 3
    n_words = int(word_of[-1]) + 1
    edges = (np.arange(n_words + 1) * T) // n_words
    sf, ef = ssb.windows(edges[:-1][word_of].astype(np.int32), np.diff(edges)[word_of].astype(np.int32))
    chain = dict(ssid=ssid_t[pid].astype(np.int32), tmat=tmat_t[pid].astype(np.int32), sf=sf, ef=ef)
    x = model_features(rs, o.model_arrays(), T)
    r = ssb.align_batch(m, [x], [chain])[0]
    w = o.state_align(x, chain["ssid"], chain["tmat"], chain["sf"], chain["ef"])
    assert r["rv"] == w["rv"] and r["best_score"] == w["best_score"] and r["n_renorm"] == w["n_renorm"]
    if w["rv"] == 0:
        for k in ("start", "dur", "score"):
            assert np.array_equal(r[k], w[k]), k


def test_long_form_five_minute_prefix(models, oracles, golden):
    """BASELINE config #4 shape on the prefix the CPU oracle can hold: fr-fr,
    30 114 frames (5 min), 126 repetitions = 1765 phones / 5295 states, word windows.
    Bit-exact against the oracle, plus the invariants used for the full hour."""
    m, o = models("fr-fr"), oracles("fr-fr")
    x, chain = _tiled_fr(golden, 126)
    assert x.shape[0] > 30000 and len(chain["ssid"]) == 1765
    r = ssb.align_batch(m, [x], [chain])[0]
    w = o.state_align(x, chain["ssid"], chain["tmat"], chain["sf"], chain["ef"])
    assert r["rv"] == w["rv"] == 0 and r["best_score"] == w["best_score"]
    assert r["n_renorm"] == w["n_renorm"]
    for k in ("start", "dur", "score"):
        assert np.array_equal(r[k], w[k]), k
    on = r["dur"] > 0
    start, dur = r["start"][on], r["dur"][on]
    assert start[0] == 0 and (start[1:] == start[:-1] + dur[:-1]).all() and start[-1] + dur[-1] == x.shape[0]
    sf, ef = np.repeat(chain["sf"], 3)[on], np.repeat(chain["ef"], 3)[on]
    assert (start >= sf).all() and (start + dur <= ef).all()


# ---- K1 over time: long utterances of small batches are scored in independent segments and
# ---- the tie steps replayed afterwards (csrc/topn_fixup.cu); $SSB_K1_SEG forces it everywhere
@pytest.mark.parametrize("seg", [5, 37])
def test_segmented_topn_equals_oracle(models, oracles, golden, monkeypatch, seg):
    m, o = models("en-us"), oracles("en-us")
    feat = golden["en-us"]["feat"]
    rs = np.random.RandomState(seg)
    feats = [feat, feat[:101], (feat[50:] * np.float32(300)).astype(np.float32)]
    f4 = feat.copy()
    f4[::7] *= np.float32(3000)         # tie steps sprinkled over segment boundaries
    feats.append(f4)
    monkeypatch.setenv("SSB_K1_SEG", str(seg))
    cw, sc = ssb.topn_batch(m, feats)
    dense = ssb.score_batch(m, feats)
    for u, f in enumerate(feats):
        wcw, wsc = o.topn_all(f)
        assert np.array_equal(cw[u], wcw) and np.array_equal(sc[u], wsc), u
        assert np.array_equal(dense[u], o.score_all(f)), u


@pytest.mark.parametrize("seg", [5, 37])
def test_segmented_aligner_equals_oracle(models, oracles, golden, monkeypatch, seg):
    """Active sets that grow over time: the segment that a codebook's first scan falls into, and
    the carried list that eval_topn has been re-sorting since the utterance began."""
    m, o, g = models("en-us"), oracles("en-us"), golden["en-us"]
    chain = chain_from_golden(g)
    nowin = dict(chain, sf=chain["sf"] * 0, ef=chain["ef"] * 0 + ssb.INT_MAX)
    f_tie = g["feat"].copy()
    f_tie[::5] *= np.float32(300)
    f_all = (g["feat"] * np.float32(3000)).astype(np.float32)
    cases = [(g["feat"], chain), (g["feat"], nowin), (f_tie, chain), (f_tie, nowin), (f_all, nowin)]
    monkeypatch.setenv("SSB_K1_SEG", str(seg))
    res = ssb.align_batch(m, [c[0] for c in cases], [c[1] for c in cases], want_chain_scr=True)
    for u, ((f, c), r) in enumerate(zip(cases, res)):
        w = o.state_align(f, c["ssid"], c["tmat"], c["sf"], c["ef"], want_senscr=True)
        sen = o.model_arrays()["sseq"][c["ssid"]].reshape(-1)
        assert np.array_equal(r["chain_scr"], w["senscr"][:, sen]), u
        assert r["rv"] == w["rv"] and r["best_score"] == w["best_score"], u
        if w["rv"] == 0:
            assert np.array_equal(r["start"], w["start"]) and np.array_equal(r["score"], w["score"]), u


def test_long_two_pass_alignment_from_text(models, oracles, golden):
    """Both passes on one long utterance through ssb_align_texts: fr-fr, 70 repetitions of the
    sentence = 16 730 frames, 280 words (more than one max_seg round), a 281-state alignment
    grammar, K1 scored in segments in both passes.  Against the oracle running the reference's
    own sequence: default-mode grammar search, pass 1's words and windows, alignment_populate,
    the aligner starting from the flags and the top-N lists pass 1 left."""
    m, o = models("fr-fr"), oracles("fr-fr")
    from conftest import model_dir
    lx = ssb.Lexicon(m, hmmdir=model_dir("fr-fr"))
    reps = 70
    x, _chain = _tiled_fr(golden, reps)
    text = " ".join(["avance de dix mètres"] * reps)
    ta = ssb.TextAlignment(m, lx, [x], [text], align_level=2)
    rv, hyp_score, n_frames = ta.status(0)
    assert rv == 0 and n_frames == len(x) + 1 and ta.hyp(0) == text
    G = lx.align_graph(text)
    p1 = o.fsg_search_active(G, x, cap=1 << 18)
    assert p1["exit"] > 0 and p1["hyp_score"] == hyp_score
    seg = ta.entries(0, "seg")
    assert len(seg) == len(p1["segs"]) > 256
    assert np.array_equal(seg[:, 1:], p1["segs"][:, 1:])
    words = p1["segs"][G["link"][p1["segs"][:, 0], 3] >= 0]
    wids = G["dict_wid"][G["link"][words[:, 0], 3]]
    c = lx.populate(wids, words[:, 1], words[:, 2] - words[:, 1] + 1)
    left = [int(w * 32 + b) for w, v in enumerate(p1["active"]) for b in range(32) if (int(v) >> b) & 1]
    w2 = o.state_align(x, c["ssid"], c["tmat"], c["sf"], c["ef"], init_active=left, init_topn=p1["carried"])
    assert w2["rv"] == 0
    st = ta.entries(0, "states")
    assert np.array_equal(st[:, 1], w2["start"]) and np.array_equal(st[:, 2], w2["dur"])
    assert np.array_equal(st[:, 3], w2["score"])
    wd = ta.entries(0, "words")
    assert len(wd) == len(wids) and wd[0, 1] == 0 and wd[-1, 1] + wd[-1, 2] == len(x)
