"""The search-module drop-in (ssb_search_t == search_module_t): start / step / finish / hyp /
seg_iter through the object's own vtable, against the reference's results in tests/golden/."""
import ctypes as C
import os

import numpy as np
import pytest

import soundswallower_b200 as ssb
from soundswallower_b200 import _lib
from conftest import model_dir

TEXT = {"en-us": "go forward ten meters", "fr-fr": "avance de dix mètres"}


@pytest.fixture(scope="module")
def host():
    out = {}
    for lang in ("en-us", "fr-fr"):
        m = ssb.AcousticModel(model_dir(lang), device=-1)
        out[lang] = (m, ssb.Lexicon(m, hmmdir=model_dir(lang)))
    return out


def test_struct_layout_is_search_module_t():
    """LP64 offsets of search_module_t / seg_iter_t (ref: search_module.h:89-113, 165-174)."""
    want = dict(vt=0, type=8, name=16, config=24, acmod=32, dict=40, d2p=48, hyp_str=56, dag=64,
                last_link=72, post=80, n_words=84, start_wid=88, silence_wid=92, finish_wid=96)
    for k, off in want.items():
        assert getattr(_lib.SearchBase, k).offset == off, k
    want = dict(vt=0, search=8, word=16, sf=24, ef=28, ascr=32, lscr=36, prob=40)
    for k, off in want.items():
        assert getattr(_lib.SegIter, k).offset == off, k
    names = [f[0] for f in _lib.SearchFuncs._fields_]
    assert names == ["start", "step", "finish", "reinit", "free", "lattice", "hyp", "prob", "seg_iter"]


def test_host_side_protocol(host):
    m, lx = host["en-us"]
    s = ssb.fsg_search(m, lx, TEXT["en-us"], name="pass1")
    assert (s.type, s.name) == ("fsg", "pass1")
    assert s.base.n_words == len(lx)
    assert [s.base.start_wid, s.base.finish_wid, s.base.silence_wid] == \
        [lx.wordid(w) for w in ("<s>", "</s>", "<sil>")]
    assert s.vt.lattice and s.vt.prob                        # fsg_search_lattice / fsg_search_prob
    assert s.start() == 0
    assert s.step(0) == -1 and "not been fed" in _lib.last_error()
    s.feed(np.zeros((3, m.blk), np.float32))
    assert s.step(0) == 1 and s.step(1) == 1                 # fsg_search_step returns 1
    assert s.step(5) == -1 and "out of order" in _lib.last_error()
    assert s.hyp() == (None, 0) and s.seg() == []            # no hypothesis before finish
    assert s.finish() == -1 and "no CPU compute path" in _lib.last_error()
    s.close()
    with pytest.raises(ssb.SsbError, match="Unknown word"):
        ssb.fsg_search(m, lx, "go xyzzyq")


def test_aligner_entries_before_the_search(host):
    """alignment_populate's phone and state entries inherit the word's window
    (ref: src/ps_alignment.c:167-170, 237-240); hyp = real words, seg = all words."""
    m, lx = host["en-us"]
    wids = [lx.wordid(w) for w in ("<sil>", "go", "the(2)", "<sil>")]
    a = ssb.state_align_search(m, lx, wids, [0, 10, 20, 50], [10, 10, 30, 5])
    assert a.type == "state_align"
    assert not a.vt.lattice and not a.vt.prob                # NULL in the reference's vtable too
    w, p, st = a.alignment("words"), a.alignment("phones"), a.alignment("states")
    assert w[:, 0].tolist() == wids and w[:, 1].tolist() == [0, 10, 20, 50]
    c = lx.populate(wids)
    assert np.array_equal(p[:, 0], c["ci"]) and np.array_equal(p[:, 4], c["parent"])
    assert np.array_equal(p[:, 1], w[p[:, 4], 1]) and np.array_equal(p[:, 2], w[p[:, 4], 2])
    E = m.n_emit
    assert len(st) == E * len(p)
    assert np.array_equal(st[:, 0], m.arrays()["sseq"][c["ssid"]].reshape(-1))
    assert np.array_equal(st[:, 1], np.repeat(p[:, 1], E)) and np.array_equal(st[:, 4], np.repeat(np.arange(len(p)), E))
    assert a.hyp() == ("go the", 0)                           # base string of the alternate
    assert a.seg() == [("<sil>", 0, 9, 0, 0), ("go", 10, 19, 0, 0), ("the(2)", 20, 49, 0, 0),
                       ("<sil>", 50, 54, 0, 0)]
    assert a.start() == 0
    a.feed(np.zeros((2, m.blk), np.float32))
    assert a.step(0) == 0                                     # state_align_search_step returns 0
    a.close()


@pytest.mark.gpu
@pytest.mark.parametrize("lang", ["en-us", "fr-fr"])
def test_two_pass_through_the_vtables(models, golden, lang):
    """decoder_set_align_text -> search_module_forward -> hyp / seg_iter, then
    decoder_alignment -> state_align_search -> alignment entries: the reference's CLI result."""
    m, g = models(lang), golden[lang]
    # the reference's default mode (compallsen = no): App. B's -2761 / -4236
    fg = np.load(os.path.join(os.path.dirname(__file__), "golden", "fsg_active_%s.npz" % lang))
    lx = ssb.Lexicon(m, hmmdir=model_dir(lang))
    feat = g["feat"]
    p1 = ssb.fsg_search(m, lx, TEXT[lang])
    assert p1.start() == 0
    assert p1.forward(feat) == len(feat)
    assert p1.finish() == 0
    hyp, score = p1.hyp()
    assert hyp == TEXT[lang] and score == int(fg["align_hyp_score"]) == {"en-us": -2761, "fr-fr": -4236}[lang]
    seg = p1.seg()
    want = fg["align_segs"]
    assert [lx.wordid(s[0]) for s in seg] == want[:, 0].tolist()
    assert np.array_equal(np.array([s[1:] for s in seg], np.int32), want[:, 1:])
    # pass 2 on pass 1's words and windows
    wids = [lx.wordid(s[0]) for s in seg]
    p2 = ssb.state_align_search(m, lx, wids, [s[1] for s in seg], [s[2] - s[1] + 1 for s in seg])
    assert np.array_equal(p1.final_active(), fg["align_active"])
    p2.set_init_active(p1.final_active())      # the acmod the two searches share
    p2.set_init_topn(p1.final_topn())          # ... and the scorer's lists
    assert p2.start() == 0 and p2.forward(feat) == len(feat) and p2.finish() == 0
    assert np.array_equal(p2.alignment("words")[:, :4], g["words"])
    ph = p2.alignment("phones")
    assert np.array_equal(ph[:, 0], g["phones"][:, 0]) and np.array_equal(ph[:, 1:4], g["phones"][:, 3:6])
    assert np.array_equal(ph[:, 4], g["phones"][:, 6])
    assert np.array_equal(p2.alignment("states"), g["states"])
    hyp2, score2 = p2.hyp()
    assert hyp2 == TEXT[lang] and score2 == int(g["words"][-1, 3])   # the last word's score
    assert p2.seg() == [(lx.wordstr(int(w[0])), int(w[1]), int(w[1] + w[2] - 1), int(w[3]), 0)
                        for w in g["words"]]
    p1.close()
    p2.close()


@pytest.mark.gpu
@pytest.mark.parametrize("lang", ["en-us", "fr-fr"])
def test_hypothesis_between_steps(models, golden, lang):
    """hyp / seg_iter while the utterance is running: what the reference's decoder_hyp /
    decoder_seg_iter return between two search steps (find_exit with final = FALSE, ref:
    src/fsg_search.c:853-960), then the final result after finish()."""
    m, g = models(lang), golden[lang]
    pg = np.load(os.path.join(os.path.dirname(__file__), "golden", "fsg_partial.npz"))
    lx = ssb.Lexicon(m, hmmdir=model_dir(lang))
    feat = g["feat"]
    p1 = ssb.fsg_search(m, lx, TEXT[lang])
    assert p1.start() == 0
    assert p1.hyp()[0] is None and p1.seg() == []              # nothing yet
    p1.feed(feat)
    t = 0
    for k, stop in enumerate(pg["%s_stops" % lang].tolist()):
        while t < stop:
            assert p1.step(t) == 1
            t += 1
        hyp, score = p1.hyp()
        want = str(pg["%s_hyp" % lang][k])
        assert (hyp or "") == want, (stop, hyp, want)
        if want:
            assert score == int(pg["%s_score" % lang][k]), stop
        n = int(pg["%s_nseg" % lang][k])
        seg = p1.seg()
        assert [lx.wordid(s[0]) for s in seg] == pg["%s_segs" % lang][k, :n, 0].tolist(), stop
        assert np.array_equal(np.array([s[1:] for s in seg], np.int32).reshape(-1, 4),
                              pg["%s_segs" % lang][k, :n, 1:]), stop
    assert t == len(feat) and p1.finish() == 0
    hyp, score = p1.hyp()
    assert hyp == TEXT[lang] and score == {"en-us": -2761, "fr-fr": -4236}[lang]
    p1.close()


@pytest.mark.gpu
def test_feature_source_callback_equals_feed(models, golden):
    """The part acmod plays: step(frame_idx) pulls the frame through the callback."""
    m, g = models("en-us"), golden["en-us"]
    lx = ssb.Lexicon(m, hmmdir=model_dir("en-us"))
    feat = g["feat"]
    asked = []

    def source(t):
        asked.append(t)
        return feat[t] if t < len(feat) else None
    a = ssb.fsg_search(m, lx, TEXT["en-us"], source=source)
    a.start()
    for t in range(len(feat)):
        assert a.step(t) == 1
    assert a.step(len(feat)) == -1 and "no frame" in _lib.last_error()
    assert a.finish() == 0
    b = ssb.fsg_search(m, lx, TEXT["en-us"])
    b.start(); b.forward(feat); b.finish()
    assert asked == list(range(len(feat) + 1))
    assert a.hyp() == b.hyp() and a.seg() == b.seg()
    # a second utterance on the same object (the decoder reuses its search)
    a.start()
    for t in range(150):
        a.step(t)
    a.finish()
    c = ssb.fsg_search(m, lx, TEXT["en-us"])
    c.start(); c.forward(feat[:150]); c.finish()
    assert a.hyp() == c.hyp() and a.seg() == c.seg()


@pytest.mark.gpu
def test_aligner_reports_failure_like_the_reference(models, golden):
    """A chain that cannot reach its final state: finish() = -1 with the reference's message."""
    m, g = models("en-us"), golden["en-us"]
    lx = ssb.Lexicon(m, hmmdir=model_dir("en-us"))
    wids = [lx.wordid(w) for w in "go forward ten meters".split()] * 4
    a = ssb.state_align_search(m, lx, wids)
    a.start(); a.forward(g["feat"][:20])
    assert a.finish() == -1 and "Failed to reach final state" in _lib.last_error()


@pytest.mark.gpu
def test_two_pass_hands_over_the_scorers_lists(models, golden):
    """en-us "mid46": three frames whose Gaussian scores all clamp sit where `go` is first scanned
    in pass 2, so the reference's state scores depend on the lists pass 1 left in the scorer."""
    m, g = models("en-us"), golden["en-us"]
    a = np.load(os.path.join(os.path.dirname(__file__), "golden", "fsg_active_en-us.npz"))
    lx = ssb.Lexicon(m, hmmdir=model_dir("en-us"))
    feat = g["feat"].copy()
    feat[46:49] *= np.float32(3000)
    want = a["mid46_p2_states"]
    # search objects
    p1 = ssb.fsg_search(m, lx, TEXT["en-us"])
    p1.start(); p1.forward(feat); p1.finish()
    seg = [s for s in p1.seg() if s[0] != "(NULL)"]
    p2 = ssb.state_align_search(m, lx, [lx.wordid(s[0]) for s in seg], [s[1] for s in seg],
                                [s[2] - s[1] + 1 for s in seg])
    p2.set_init_active(p1.final_active())
    p2.set_init_topn(p1.final_topn())
    p2.start(); p2.forward(feat); assert p2.finish() == 0
    assert np.array_equal(p2.alignment("states"), want)
    # align_texts (Python) and ssb_align_texts (C)
    r = ssb.align_texts(m, lx, [feat], [TEXT["en-us"]])[0]
    assert np.array_equal(r["states"][:, :4], want[:, :4])
    ta = ssb.TextAlignment(m, lx, [feat], [TEXT["en-us"]], align_level=2)
    assert np.array_equal(ta.entries(0, "states"), want)
