"""K4 (FSG token passing) through the C ABI against the golden history tables of the reference
(tests/golden/fsg_*.npz) and against the oracle on ragged / multi-grammar batches."""
import os

import numpy as np
import pytest

import soundswallower_b200 as ssb
from conftest import DATA, GOLDEN, model_features
from test_oracle_fsg import graph_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fsg_golden():
    return {lang: np.load(os.path.join(GOLDEN, "fsg_%s.npz" % lang)) for lang in ("en-us", "fr-fr")}


@pytest.mark.parametrize("lang", ["en-us", "fr-fr"])
@pytest.mark.parametrize("name", ["align", "jsgf"])
def test_fsg_golden(models, golden, fsg_golden, lang, name):
    m, g = models(lang), fsg_golden[lang]
    r = ssb.fsg_batch(m, [golden[lang]["feat"]], [graph_of(g, name)], want_hist=True)[0]
    assert r["rv"] == 0
    assert np.array_equal(r["hist"], g[name + "_hist"])
    assert r["n_hmm_eval"] == int(g[name + "_n_hmm_eval"])
    assert r["hyp_score"] == int(g[name + "_hyp_score"])
    assert np.array_equal(r["segs"][:, 1:], g[name + "_segs"][:, 1:])
    assert r["n_launches"] >= 4 and r["kernel_ms"]["fsg_search"] > 0


def test_fsg_ragged_two_grammars(models, oracles, golden, fsg_golden):
    """Utterances of different lengths searching different graphs in one batch, some of which
    do not match their grammar; every history table equals the oracle's."""
    m, o = models("en-us"), oracles("en-us")
    g, feat = fsg_golden["en-us"], golden["en-us"]["feat"]
    graphs = [graph_of(g, "align"), graph_of(g, "jsgf")]
    rs = np.random.RandomState(31)
    feats, ug = [], []
    for u in range(37):
        k = int(rs.randint(0, 4))
        if k == 0:
            f = feat
        elif k == 1:
            f = feat[:int(rs.randint(1, 278))]           # truncated: may not reach the final state
        elif k == 2:
            f = (feat + rs.normal(0, 0.3, feat.shape)).astype(np.float32)
        else:
            f = model_features(rs, o.model_arrays(), int(rs.randint(1, 80)))  # noise
        feats.append(f)
        ug.append(u % 2)
    feats[5] = feats[5][:0]  # an empty utterance
    res = ssb.fsg_batch(m, feats, graphs, utt_graph=ug, want_hist=True)
    n_match = 0
    for u, (f, r) in enumerate(zip(feats, res)):
        w = o.fsg_search(graphs[ug[u]], o.score_all(f) if len(f) else np.zeros((0, o.n_sen), np.int16))
        assert r["rv"] == w["rv"] == 0, u
        assert np.array_equal(r["hist"], w["hist"]), u
        assert r["n_hmm_eval"] == w["n_hmm_eval"], u
        assert r["exit"] == w["exit"], u
        if w["exit"] > 0:
            n_match += 1
            assert r["hyp_score"] == w["hyp_score"], u
            assert np.array_equal(r["segs"], w["segs"]), u
    assert n_match >= 10


def test_fsg_history_overflow_and_bad_graph(models, golden, fsg_golden):
    m, g = models("en-us"), fsg_golden["en-us"]
    feat = golden["en-us"]["feat"]
    # the raw call reports the overflow (rv == -2); the public wrapper repeats with a larger table
    # (the reference's history is unbounded) and gives the unlimited answer
    r = ssb._fsg_batch_once(m, [feat[:60]], [graph_of(g, "align")], hist_cap=16)[0]
    assert r["rv"] == -2
    r = ssb.fsg_batch(m, [feat[:60]], [graph_of(g, "align")], hist_cap=16, want_hist=True)[0]
    w = ssb.fsg_batch(m, [feat[:60]], [graph_of(g, "align")], want_hist=True)[0]
    assert r["rv"] == 0 and np.array_equal(r["hist"], w["hist"]) and r["hyp_score"] == w["hyp_score"]
    # a segmentation longer than max_seg: asked again, not silently empty
    r = ssb.fsg_batch(m, [feat], [graph_of(g, "align")], max_seg=2)[0]
    assert r["n_seg"] == len(r["segs"]) > 2
    bad = dict(graph_of(g, "align"))
    bad["pnode"] = bad["pnode"].copy()
    bad["pnode"][3, 0] = m.n_sseq + 5
    with pytest.raises(ssb.SsbError, match="malformed"):
        ssb.fsg_batch(m, [feat[:10]], [bad])


def test_config3_shape(models, golden, fsg_golden):
    """BASELINE config #3 shape at reduced width: 256 utterances x 279 frames on goforward.gram;
    identical inputs give identical results wherever they sit, noisy copies keep the words."""
    m, g = models("en-us"), fsg_golden["en-us"]
    feat = golden["en-us"]["feat"]
    rs = np.random.RandomState(8)
    feats = [feat] + [(feat + rs.normal(0, 0.05, feat.shape)).astype(np.float32) for _ in range(255)]
    feats[200] = feats[0]
    res = ssb.fsg_batch(m, feats, [graph_of(g, "jsgf")])
    assert np.array_equal(res[0]["segs"][:, 1:], g["jsgf_segs"][:, 1:])
    assert np.array_equal(res[200]["segs"], res[0]["segs"]) and res[200]["hyp_score"] == res[0]["hyp_score"]
    words = [s for s in res[0]["segs"][:, 0]]
    for r in res:
        assert r["rv"] == 0 and r["exit"] > 0
        assert [s for s in r["segs"][:, 0]] == words            # same link sequence (same words)
        sf, ef = r["segs"][:, 1], r["segs"][:, 2]
        assert sf[0] == 0 and ef[-1] == 277 and (sf[1:] >= ef[:-1]).all()


@pytest.mark.parametrize("lang", ["en-us", "fr-fr"])
def test_two_pass_alignment_on_gpu_equals_cli(models, golden, fsg_golden, lang):
    """BASELINE config #1 on the GPU: pass 1 = FSG search of the alignment grammar (word
    boundaries), pass 2 = chain Viterbi inside those windows.  The state segmentation must be the
    one the reference's CLI produces (tests/golden/align_*.npz `states`).
    Only the word -> phone expansion (alignment_populate, host graph preparation) comes from
    the fixture."""
    m, g, fg = models(lang), golden[lang], fsg_golden[lang]
    feat = g["feat"]
    p1 = ssb.fsg_batch(m, [feat], [graph_of(fg, "align")])[0]
    assert p1["rv"] == 0 and p1["exit"] > 0
    segs = p1["segs"]
    # decoder_alignment: one alignment word per segment, start = sf, duration = ef - sf + 1
    # (ref: src/decoder.c:753-768)
    w_start, w_dur = segs[:, 1], segs[:, 2] - segs[:, 1] + 1
    assert np.array_equal(w_start, g["words"][:, 1]) and np.array_equal(w_dur, g["words"][:, 2])
    parent = g["phones"][:, 6]
    sf, ef = ssb.windows(w_start[parent], w_dur[parent])
    chain = dict(ssid=g["phones"][:, 1].astype(np.int32), tmat=g["phones"][:, 2].astype(np.int32),
                 sf=sf, ef=ef)
    p2 = ssb.align_batch(m, [feat], [chain])[0]
    st = g["states"]  # the reference's full two-pass result
    assert p2["rv"] == 0
    assert np.array_equal(p2["start"], st[:, 1]) and np.array_equal(p2["dur"], st[:, 2])
    assert np.array_equal(p2["score"], st[:, 3])
    ps, pd, pc = ssb.propagate(p2["start"], p2["dur"], p2["score"], m.n_emit)
    assert np.array_equal(ps, g["phones"][:, 3]) and np.array_equal(pd, g["phones"][:, 4])
    assert np.array_equal(pc, g["phones"][:, 5])


# ---- the reference's default mode: active lists, scoring inside the search kernel
from test_oracle_fsg import ACTIVE_CASES, TEXT, active_case  # noqa: E402
from conftest import model_dir  # noqa: E402


@pytest.fixture(scope="module")
def active_golden():
    return {lang: np.load(os.path.join(GOLDEN, "fsg_active_%s.npz" % lang)) for lang in ("en-us", "fr-fr")}


@pytest.mark.parametrize("lang", ["en-us", "fr-fr"])
@pytest.mark.parametrize("name", ACTIVE_CASES)
def test_fsg_active_lists_golden(models, golden, fsg_golden, active_golden, lang, name):
    """History table, counters, hypothesis score (App. B: -2761 / -4236), segmentation and the
    flags left in acmod, all equal to the unmodified reference run with its default config."""
    m, a = models(lang), active_golden[lang]
    G, feat = active_case(lang, name, golden[lang]["feat"], fsg_golden[lang], a)
    r = ssb.fsg_batch(m, [feat], [G], want_hist=True, compallsen=False)[0]
    assert r["rv"] == 0
    assert np.array_equal(r["hist"], a[name + "_hist"])
    assert r["n_hmm_eval"] == int(a[name + "_n_hmm_eval"])
    assert r["n_sen_eval"] == int(a[name + "_n_sen_eval"])
    assert np.array_equal(r["active"], a[name + "_active"])
    assert r["hyp_score"] == int(a[name + "_hyp_score"])
    assert np.array_equal(r["segs"][:, 1:], a[name + "_segs"][:, 1:])


@pytest.mark.parametrize("lang", ["en-us", "fr-fr"])
@pytest.mark.parametrize("name", ["align", "mid46", "odd", "wide", "noisy", "tiled"])
def test_second_pass_inherits_the_scorers_lists(models, oracles, golden, fsg_golden, active_golden, lang, name):
    """ptm_mgau's top-N lists are not reset between the passes: the grammar search exports the
    lists the reference is left carrying (history slot 1 = after the last odd frame, re-sorted by
    eval_topn on every frame since the codebook's last scan) and the aligner starts from them.
    On en-us mid46 / odd / wide the reference's state scores come out only this way."""
    m, o, a = models(lang), oracles(lang), active_golden[lang]
    lx = ssb.Lexicon(m, hmmdir=model_dir(lang))
    G, feat = active_case(lang, name, golden[lang]["feat"], fsg_golden[lang], a)
    p1 = ssb.fsg_batch(m, [feat, feat[:-3]], [G], compallsen=False)
    for f, r in zip((feat, feat[:-3]), p1):     # both frame-count parities
        assert np.array_equal(r["carried"], o.fsg_search_active(G, f)["carried"].reshape(-1, 4))
    segs = a[name + "_segs"]
    segs = segs[segs[:, 0] >= 0]
    chain = lx.populate(segs[:, 0], segs[:, 1], segs[:, 2] - segs[:, 1] + 1)
    left = [int(w * 32 + b) for w, x in enumerate(p1[0]["active"]) for b in range(32) if (int(x) >> b) & 1]
    want = a[name + "_p2_states"]
    r = ssb.align_batch(m, [feat], [chain], init_active=[left], init_topn=[p1[0]["carried"]])[0]
    assert r["rv"] == 0 and np.array_equal(np.stack([r["start"], r["dur"], r["score"]], 1), want[:, 1:4])
    if lang == "en-us" and name in ("mid46", "odd", "wide"):
        r0 = ssb.align_batch(m, [feat], [chain], init_active=[left])[0]
        assert not np.array_equal(r0["score"], want[:, 3])
        # dense first pass (compallsen = yes) exports its lists too: K1's own of the last odd frame
        d1 = ssb.fsg_batch(m, [feat], [G], compallsen=True)[0]
        assert d1["carried"].shape == (m.n_mgau * m.n_feat, 4)


@pytest.mark.parametrize("seg", [5, 37])
def test_carried_lists_with_segmented_topn(models, oracles, golden, fsg_golden, active_golden, monkeypatch, seg):
    """The same hand-over when K1 scores segments independently and the fix-up replays the ties."""
    m, a = models("en-us"), active_golden["en-us"]
    lx = ssb.Lexicon(m, hmmdir=model_dir("en-us"))
    monkeypatch.setenv("SSB_K1_SEG", str(seg))
    for name in ("mid46", "odd", "wide"):
        G, feat = active_case("en-us", name, golden["en-us"]["feat"], fsg_golden["en-us"], a)
        p1 = ssb.fsg_batch(m, [feat], [G], compallsen=False)[0]
        assert p1["hyp_score"] == int(a[name + "_hyp_score"])
        segs = a[name + "_segs"]
        segs = segs[segs[:, 0] >= 0]
        chain = lx.populate(segs[:, 0], segs[:, 1], segs[:, 2] - segs[:, 1] + 1)
        left = [int(w * 32 + b) for w, x in enumerate(p1["active"]) for b in range(32) if (int(x) >> b) & 1]
        r = ssb.align_batch(m, [feat], [chain], init_active=[left], init_topn=[p1["carried"]])[0]
        assert np.array_equal(np.stack([r["start"], r["dur"], r["score"]], 1), a[name + "_p2_states"][:, 1:4]), name


@pytest.mark.parametrize("lang,name", [("fr-fr", "trunc"), ("en-us", "noisy"), ("fr-fr", "tiled")])
def test_second_pass_starts_from_the_flags_pass_one_left(models, golden, fsg_golden, active_golden, lang, name):
    """decoder_alignment after a default-mode first pass: the aligner never clears acmod's flags
    (ref: src/state_align_search.c:186-188), so the senones of pass 1's last frame stay active
    throughout pass 2 and move its normalisers.  fr-fr/trunc leaves 29 of them."""
    m, a = models(lang), active_golden[lang]
    lx = ssb.Lexicon(m, hmmdir=model_dir(lang))
    G, feat = active_case(lang, name, golden[lang]["feat"], fsg_golden[lang], a)
    p1 = ssb.fsg_batch(m, [feat], [G], compallsen=False)[0]
    segs = a[name + "_segs"]
    segs = segs[segs[:, 0] >= 0]
    chain = lx.populate(segs[:, 0], segs[:, 1], segs[:, 2] - segs[:, 1] + 1)
    left = [int(w * 32 + b) for w, x in enumerate(p1["active"]) for b in range(32) if (int(x) >> b) & 1]
    r = ssb.align_batch(m, [feat], [chain], init_active=[left])[0]
    want = a[name + "_p2_states"]
    assert r["rv"] == int(a[name + "_p2_rv"]) == 0
    assert np.array_equal(np.stack([r["start"], r["dur"], r["score"]], 1), want[:, 1:4])
    if name == "trunc":   # the flags matter: without them the scores differ
        r0 = ssb.align_batch(m, [feat], [chain])[0]
        assert len(left) == 29 and not np.array_equal(r0["score"], want[:, 3])


def test_fsg_active_lists_ragged_batch_vs_oracle(models, oracles, golden, fsg_golden):
    """Many utterances, two grammars, every length: the data-dependent active sets, the bridging
    entries of the delta list and the tie replays all have to agree with the oracle."""
    m, o = models("en-us"), oracles("en-us")
    g, feat = fsg_golden["en-us"], golden["en-us"]["feat"]
    graphs = [graph_of(g, "align"), graph_of(g, "jsgf")]
    rs = np.random.RandomState(77)
    feats, ug = [], []
    for u in range(48):
        k = int(rs.randint(0, 4))
        if k == 0:
            f = (feat + rs.normal(0, 0.05, feat.shape)).astype(np.float32)
        elif k == 1:
            f = feat[:int(rs.randint(1, 278))]
        elif k == 2:
            f = (feat + rs.normal(0, 0.3, feat.shape)).astype(np.float32)
        else:
            f = model_features(rs, o.model_arrays(), int(rs.randint(1, 80)))
        feats.append(f)
        ug.append(u % 2)
    feats[7] = feats[7][:0]
    res = ssb.fsg_batch(m, feats, graphs, utt_graph=ug, want_hist=True, compallsen=False)
    n_match = 0
    for u, (f, r) in enumerate(zip(feats, res)):
        w = o.fsg_search_active(graphs[ug[u]], f if len(f) else np.zeros((0, 39), np.float32))
        assert r["rv"] == w["rv"] == 0, u
        assert np.array_equal(r["hist"], w["hist"]), u
        assert r["n_hmm_eval"] == w["n_hmm_eval"] and r["n_sen_eval"] == w["n_sen_eval"], u
        assert np.array_equal(r["active"], w["active"]), u
        assert r["exit"] == w["exit"], u
        if w["exit"] > 0:
            n_match += 1
            assert r["hyp_score"] == w["hyp_score"] and np.array_equal(r["segs"], w["segs"]), u
    assert n_match >= 10


def test_fsg_active_lists_need_the_ptm_tensor_core_path(models, golden, fsg_golden, monkeypatch):
    m, g = models("en-us"), fsg_golden["en-us"]
    monkeypatch.setenv("SSB_K1", "fp32")
    with pytest.raises(ssb.SsbError, match="active lists need"):
        ssb.fsg_batch(m, [golden["en-us"]["feat"][:20]], [graph_of(g, "align")], compallsen=False)


def _tie_heavy(feat, kind):
    """Inputs on which integer Gaussian scores tie massively: far from every mean the float
    distances are coarser than 1 (or clamp at INT32_MIN), so the reference's carried-list rules
    (eval_topn's stable re-sort on every frame, eval_cb's newcomer-before-equals) decide."""
    f = feat.copy()
    if kind == "x300":
        f *= np.float32(300)
    elif kind == "x3000":
        f *= np.float32(3000)
    elif kind == "every5th":
        f[::5] *= np.float32(300)
    elif kind == "bursts":
        f[40:60] *= np.float32(3000)
        f[120:123] *= np.float32(100)
        f[200:] *= np.float32(30)
    return f


@pytest.mark.parametrize("kind", ["x300", "x3000", "every5th", "bursts"])
def test_fsg_active_lists_when_scores_tie(models, oracles, golden, fsg_golden, kind):
    m, o = models("en-us"), oracles("en-us")
    g, feat = fsg_golden["en-us"], golden["en-us"]["feat"]
    graphs = [graph_of(g, "align"), graph_of(g, "jsgf")]
    feats = [_tie_heavy(feat, kind), _tie_heavy(feat[:150], kind), _tie_heavy(feat[30:], kind)]
    res = ssb.fsg_batch(m, feats, graphs, utt_graph=[0, 1, 0], want_hist=True, compallsen=False)
    for u, (f, r) in enumerate(zip(feats, res)):
        w = o.fsg_search_active(graphs[[0, 1, 0][u]], f)
        assert r["rv"] == w["rv"] == 0, u
        assert np.array_equal(r["hist"], w["hist"]), u
        assert r["n_sen_eval"] == w["n_sen_eval"] and np.array_equal(r["active"], w["active"]), u
        assert r["exit"] == w["exit"] and (w["exit"] <= 0 or r["hyp_score"] == w["hyp_score"]), u


@pytest.mark.parametrize("kind", ["x300", "x3000", "every5th", "bursts"])
def test_dense_and_aligner_when_scores_tie(models, oracles, golden, kind):
    """The same inputs through K1's own slow path: dense scores (every codebook scanned on
    every frame) and the aligner's planned active sets (codebooks join over time)."""
    from conftest import chain_from_golden
    m, o, g = models("en-us"), oracles("en-us"), golden["en-us"]
    f = _tie_heavy(g["feat"], kind)
    assert np.array_equal(ssb.score_batch(m, [f[:120]])[0], o.score_all(f[:120]))
    chain = chain_from_golden(g)
    for c in (chain, dict(chain, sf=chain["sf"] * 0, ef=chain["ef"] * 0 + ssb.INT_MAX)):
        r = ssb.align_batch(m, [f], [c], want_chain_scr=True)[0]
        w = o.state_align(f, c["ssid"], c["tmat"], c["sf"], c["ef"], want_senscr=True)
        sen = o.model_arrays()["sseq"][c["ssid"]].reshape(-1)
        assert np.array_equal(r["chain_scr"], w["senscr"][:, sen])
        assert r["rv"] == w["rv"] and r["best_score"] == w["best_score"]
        if w["rv"] == 0:
            assert np.array_equal(r["start"], w["start"]) and np.array_equal(r["score"], w["score"])


@pytest.mark.gpu
def test_fsg_file_grammar_built_in_tree_decodes_like_the_reference(models, golden):
    """Config #3's kind of grammar without the reference: goforward.fsg's transition list ->
    ssb_fsg_build -> K4 on dense scores; history table, HMM evaluations, hypothesis score and
    segmentation equal the reference's on decoder_set_fsg(fsg_model_readfile(...))."""
    g = np.load(os.path.join(GOLDEN, "fsg_file_en-us.npz"))
    m = models("en-us")
    lx = ssb.Lexicon(m, hmmdir=model_dir("en-us"))
    graph = lx.fsg_graph(*ssb.read_fsg_file(os.path.join(DATA, "goforward.fsg")))
    r = ssb.fsg_batch(m, [golden["en-us"]["feat"]], [graph], want_hist=True, compallsen=True)[0]
    assert r["rv"] == 0
    assert np.array_equal(r["hist"], g["file_hist"])
    assert r["n_hmm_eval"] == int(g["file_n_hmm_eval"]) and r["hyp_score"] == int(g["file_hyp_score"])
    assert np.array_equal(r["segs"][:, 1:], g["file_segs"][:, 1:])
    lx.close()
