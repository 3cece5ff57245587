"""Host-side graph preparation (first half): dictionary, triphone lookup and
alignment_populate.  tests/golden/lexicon.npz holds word ids and phone chains the compiled
reference produced (tools/make_golden.py --lexicon)."""
import os

import numpy as np
import pytest

import soundswallower_b200 as ssb
from conftest import DATA, GOLDEN, chain_from_golden, model_dir


@pytest.fixture(scope="module")
def lex_golden():
    return np.load(os.path.join(GOLDEN, "lexicon.npz"))


@pytest.fixture(scope="module")
def lexicons():
    out = {}
    for lang in ("en-us", "fr-fr"):
        m = ssb.AcousticModel(model_dir(lang), device=-1)
        out[lang] = (m, ssb.Lexicon(m, hmmdir=model_dir(lang)))
    return out


@pytest.mark.parametrize("lang", ["en-us", "fr-fr"])
def test_word_ids_match_reference(lex_golden, lexicons, lang):
    _, lx = lexicons[lang]
    assert len(lx) == int(lex_golden[lang + "_size"])
    for wid, s in zip(lex_golden[lang + "_probe_wid"], lex_golden[lang + "_probe_str"]):
        assert lx.wordstr(int(wid)) == str(s) and lx.wordid(str(s)) == int(wid)
    assert lx.wordid("no such word") == -1 and lx.wordstr(len(lx)) is None
    assert lx.is_filler(lx.wordid("<sil>")) and not lx.is_filler(lx.wordid("<s>"))
    assert not lx.is_filler(0)


@pytest.mark.parametrize("lang", ["en-us", "fr-fr"])
def test_populate_matches_reference(lex_golden, lexicons, lang):
    _, lx = lexicons[lang]
    off, wids, ph = lex_golden[lang + "_wid_off"], lex_golden[lang + "_wids"], lex_golden[lang + "_phones"]
    p0 = 0
    for i, n in enumerate(lex_golden[lang + "_n_phones"]):
        c = lx.populate(wids[off[i]:off[i + 1]])
        want = ph[p0:p0 + n]
        p0 += n
        assert len(c["ssid"]) == n
        assert np.array_equal(c["ci"], want[:, 0]) and np.array_equal(c["ssid"], want[:, 1])
        assert np.array_equal(c["tmat"], want[:, 2]) and np.array_equal(c["parent"], want[:, 3])


@pytest.mark.parametrize("lang,text", [("en-us", "go forward ten meters"),
                                       ("fr-fr", "avance de dix mètres")])
def test_text_to_chain_equals_the_golden_alignment_chain(lexicons, golden, lang, text):
    """<sil> words </sil> as decoder_alignment receives them: the chain equals the one the
    reference's CLI aligned (tests/golden/align_*.npz), windows included."""
    _, lx = lexicons[lang]
    g = golden[lang]
    wids = g["words"][:, 0]
    # (pass 1 may have chosen alternate pronunciations: "de(2)", "mètres(4)")
    assert [lx.wordstr(int(w)).split("(")[0] for w in wids] == ["<sil>"] + text.split() + ["<sil>"]
    c = lx.populate(wids, g["words"][:, 1], g["words"][:, 2])
    want = chain_from_golden(g)
    for k in ("ssid", "tmat", "sf", "ef"):
        assert np.array_equal(c[k], want[k]), k


def test_populate_errors(lexicons):
    _, lx = lexicons["en-us"]
    with pytest.raises(ssb.SsbError):
        lx.populate([len(lx) + 5])
    assert len(lx.populate([])["ssid"]) == 0
    m = lexicons["en-us"][0]
    with pytest.raises(ssb.SsbError):
        ssb.Lexicon(m, dictfile="/nonexistent/dict.txt")


def test_reference_agrees_when_present(lexicons):
    from oracle import refshim
    if not refshim.available():
        pytest.skip("oracle/_ref/libssref.so not built here")
    rs = np.random.RandomState(5)
    for lang in ("en-us", "fr-fr"):
        _, lx = lexicons[lang]
        r = refshim.Ref(model_dir(lang))
        for _ in range(150):
            wids = rs.randint(0, len(lx), rs.randint(1, 20)).astype(np.int32)
            a, b = r.populate(wids)["phones"], lx.populate(wids)
            assert np.array_equal(a[:, 1], b["ssid"]) and np.array_equal(a[:, 2], b["tmat"])
        r.close()


# ------------------------------------------------------------------ alignment grammar + lextree
GRAPH_KEYS = ("link", "link_flag", "arc_off", "root", "pnode", "ctxt")
GRAPH_DIMS = ("n_state", "start", "final", "n_ciphone", "sil", "beam", "pbeam", "wbeam", "maxhmmpf")


@pytest.mark.parametrize("lang,text", [("en-us", "go forward ten meters"),
                                       ("fr-fr", "avance de dix mètres")])
def test_align_graph_equals_the_reference_graph(lexicons, lang, text):
    """decoder_set_align_text + fsg_search_init + fsg_lextree_init, flattened: identical --
    link order, node order, context sets -- to the graph the reference searched
    (tests/golden/fsg_*.npz, dumped by oracle/ref_shim.c:ref_fsg_dump)."""
    from test_oracle_fsg import graph_of
    _, lx = lexicons[lang]
    want = graph_of(np.load(os.path.join(GOLDEN, "fsg_%s.npz" % lang)), "align")
    got = lx.align_graph(text)
    for k in GRAPH_DIMS:
        assert int(got[k]) == int(want[k]), k
    for k in GRAPH_KEYS:
        assert np.array_equal(got[k], np.asarray(want[k]).reshape(got[k].shape)), k
    assert got["words"][:len(text.split())] == text.split()
    with pytest.raises(ssb.SsbError, match="Unknown word"):
        lx.align_graph("go forward xyzzyplugh")


def test_align_graphs_agree_with_reference_when_present(lexicons):
    """Random transcripts (1 to 150 words, repeated words, words with alternate
    pronunciations, single-phone words) against the compiled reference."""
    from oracle import refshim
    if not refshim.available():
        pytest.skip("oracle/_ref/libssref.so not built here")
    rs = np.random.RandomState(3)
    for lang in ("en-us", "fr-fr"):
        _, lx = lexicons[lang]
        r = refshim.Ref(model_dir(lang))
        n_main = lx.wordid("<sil>") - 2
        for it in range(40):
            ws = []
            while len(ws) < int(rs.choice([1, 2, 3, 5, 8, 13, 40, 150])):
                w = lx.wordstr(int(rs.randint(0, n_main)))
                if "(" not in w and w[0] not in "<[":
                    ws.append(w)
            if it % 7 == 0:
                ws += ws[:2]
            text = " ".join(ws)
            got, want = lx.align_graph(text), r.fsg_graph(align_text=text)
            assert all(int(got[k]) == int(want[k]) for k in GRAPH_DIMS), text
            assert all(np.array_equal(got[k], want[k]) for k in GRAPH_KEYS), text
        r.close()


# ------------------------------------------------------------------ general grammars (ssb_fsg_build)
def test_fsg_file_graph_equals_reference():
    """decoder_set_fsg on tests/data/goforward.fsg (word + null transitions, null closure,
    silence / filler loops, lextree): ssb_fsg_build's arrays equal the reference's dump
    (tests/golden/fsg_file_en-us.npz, tools/make_golden.py --fsg-file)."""
    g = np.load(os.path.join(GOLDEN, "fsg_file_en-us.npz"))
    m = ssb.AcousticModel(model_dir("en-us"), device=-1)
    lx = ssb.Lexicon(m, hmmdir=model_dir("en-us"))
    n_state, start, final, trans = ssb.read_fsg_file(os.path.join(DATA, "goforward.fsg"))
    assert (n_state, start, final, len(trans)) == (7, 0, 6, 17) and trans[3][3] is None
    got = lx.fsg_graph(n_state, start, final, trans)
    for k in ("n_state", "start", "final", "n_ciphone", "sil", "beam", "pbeam", "wbeam", "maxhmmpf"):
        assert int(got[k]) == int(g["file_" + k]), k
    for k in GRAPH_KEYS:
        assert np.array_equal(got[k], g["file_" + k]), k
    assert (got["link"][:, 3] < 0).sum() == 2          # the two null transitions survive as links
    with pytest.raises(ssb.SsbError, match="Unknown word"):
        lx.fsg_graph(2, 0, 1, [(0, 1, 1.0, "xyzzyq")])
    with pytest.raises(ssb.SsbError, match="probability"):
        lx.fsg_graph(2, 0, 1, [(0, 1, 0.0, "go")])
    # a chain of null transitions: the closure adds 0 -> 2
    c = lx.fsg_graph(4, 0, 3, [(0, 1, 1.0, None), (1, 2, 0.5, None), (2, 3, 1.0, "go")])
    nulls = {(int(l[0]), int(l[1])) for l in c["link"] if l[3] < 0}
    assert nulls == {(0, 1), (1, 2), (0, 2)}
    lx.close()
    m.close()
