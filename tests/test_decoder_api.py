"""The decoder-level mirror (soundswallower_b200.Decoder): the reference's Python surface
(start_utt / process_raw / end_utt / hyp / seg / alignment / dumps / decode_file) and the
JSON of decoder_result_json, against the reference CLI's output and the
golden alignments."""
import json
import os

import numpy as np
import pytest

import soundswallower_b200 as ssb
from soundswallower_b200 import decoder as dec_mod
from conftest import DATA, model_dir

# `soundswallower --align tests/data/goforward.txt --phone-align tests/data/goforward.wav`
CLI_JSON = ('{"b":0.000,"d":2.790,"p":1.000,"t":"go forward ten meters","w":[{"b":0.000,"d":0.460,"p":0.993,"t":"<sil>","w":[{"b":0.000,"d":0.460,"p":0.993,"t":"SIL"}]},'
            '{"b":0.460,"d":0.180,"p":0.990,"t":"go","w":[{"b":0.460,"d":0.080,"p":0.995,"t":"G"},{"b":0.540,"d":0.100,"p":0.994,"t":"OW"}]},'
            '{"b":0.640,"d":0.530,"p":0.967,"t":"forward","w":[{"b":0.640,"d":0.140,"p":0.990,"t":"F"},{"b":0.780,"d":0.060,"p":0.995,"t":"AO"},{"b":0.840,"d":0.100,"p":0.995,"t":"R"},{"b":0.940,"d":0.070,"p":0.996,"t":"W"},{"b":1.010,"d":0.110,"p":0.994,"t":"ER"},{"b":1.120,"d":0.050,"p":0.997,"t":"D"}]},'
            '{"b":1.170,"d":0.360,"p":0.962,"t":"ten","w":[{"b":1.170,"d":0.150,"p":0.980,"t":"T"},{"b":1.320,"d":0.090,"p":0.995,"t":"EH"},{"b":1.410,"d":0.120,"p":0.987,"t":"N"}]},'
            '{"b":1.530,"d":0.580,"p":0.956,"t":"meters","w":[{"b":1.530,"d":0.060,"p":0.996,"t":"M"},{"b":1.590,"d":0.120,"p":0.994,"t":"IY"},{"b":1.710,"d":0.030,"p":0.986,"t":"T"},{"b":1.740,"d":0.160,"p":0.993,"t":"ER"},{"b":1.900,"d":0.210,"p":0.986,"t":"Z"}]},'
            '{"b":2.110,"d":0.670,"p":0.963,"t":"<sil>","w":[{"b":2.110,"d":0.670,"p":0.963,"t":"SIL"}]}]}\n')
TEXT = {"en-us": "go forward ten meters", "fr-fr": "avance de dix mètres"}
RAW = {"en-us": "goforward.raw", "fr-fr": "goforward_fr.raw"}


def _alignment_from_golden(m, lx, g):
    E = m.n_emit
    words = []
    for i, w in enumerate(g["words"]):
        phones = []
        for q in np.nonzero(g["phones"][:, 6] == i)[0]:
            st = [dec_mod.AlignmentEntry(str(int(s[0])), s[1], s[2], s[3]) for s in g["states"][q * E:(q + 1) * E]]
            p = g["phones"][q]
            phones.append(dec_mod.AlignmentEntry(m.ciname(int(p[0])), p[3], p[4], p[5], st))
        words.append(dec_mod.AlignmentEntry(lx.wordstr(int(w[0])), w[1], w[2], w[3], phones))
    return dec_mod.Alignment(words)


def test_result_json_formatting_matches_the_cli(golden):
    """decoder_result_json's layout (ref: src/decoder.c:1339-1593) from the golden alignment:
    no GPU needed, the numbers are the reference's."""
    d = ssb.Decoder(model_dir("en-us"), device=-1)
    g = golden["en-us"]
    fg = np.load(os.path.join(os.path.dirname(__file__), "golden", "fsg_active_en-us.npz"))
    d.set_align_text(TEXT["en-us"])
    graph = d._graph
    # pass-1 segmentation rows -> link ids of the alignment grammar
    segs, state = [], 0
    for wid, sf, ef, ascr, lscr in fg["align_segs"]:
        cand = [k for k in range(len(graph["link"])) if graph["link"][k, 0] == state
                and graph["link"][k, 3] >= 0 and graph["dict_wid"][graph["link"][k, 3]] == wid]
        assert cand
        segs.append([cand[0], sf, ef, ascr, lscr])
        state = int(graph["link"][cand[0], 1])
    p1 = dict(rv=0, exit=1, segs=np.array(segs, np.int32), hyp_score=int(fg["align_hyp_score"]))
    r = dec_mod._Result(d, len(g["feat"]), graph, p1)
    r.alignment = _alignment_from_golden(d.model, d.lexicon, g)
    d._res = r
    assert d.n_frames == 279
    assert d.dumps(align_level=1) == CLI_JSON
    assert d.hyp.text == TEXT["en-us"] and d.hyp.prob == 1.0
    two = json.loads(d.dumps(start_time=10.0, align_level=2))
    assert two["b"] == 10.0 and two["w"][1]["w"][0]["w"][0] == {"b": 10.46, "d": 0.03, "p": 0.998, "t": "2085"}
    assert [len(p["w"]) for w in two["w"] for p in w["w"]] == [3] * 18
    zero = json.loads(d.dumps(align_level=0))
    assert [w["t"] for w in zero["w"]] == ["<sil>", "go", "forward", "ten", "meters", "<sil>"]
    assert zero["w"][1] == {"b": 0.46, "d": 0.18, "p": round(1.0001 ** -133, 3), "t": "go"}
    assert zero["w"][0]["p"] == round(1.0001 ** (-230 - 337), 3)      # App. B: <sil> 0 45 -230 -337
    assert abs(d.hyp.score - 1.0001 ** -2761) < 1e-12
    assert [s.text for s in d.seg] == [w["t"] for w in zero["w"]]
    assert [w.name for w in d.alignment] == [w["t"] for w in zero["w"]]
    assert [p.name for p in list(d.alignment.words())[2]] == ["F", "AO", "R", "W", "ER", "D"]
    assert len(list(d.alignment.phones())) == 18 and len(list(d.alignment.states())) == 54
    d.close()


def test_protocol_errors():
    d = ssb.Decoder(model_dir("en-us"), device=-1)
    with pytest.raises(RuntimeError):
        d.end_utt()
    with pytest.raises(RuntimeError):
        d.process_raw(b"\0\0")
    d.start_utt()
    with pytest.raises(RuntimeError):
        d.start_utt()
    with pytest.raises(RuntimeError, match="grammar"):
        d.end_utt()
    with pytest.raises(RuntimeError, match="Failed to set up alignment"):
        d.set_align_text("go xyzzyq")
    assert d.lookup_word("forward") == "F AO R W ER D" and d.lookup_word("xyzzyq") is None
    assert d.hyp == ssb.Hyp(None, 0., 0.) and list(d.seg) == [] and d.alignment is None
    d.close()


@pytest.mark.gpu
def test_cli_json_from_audio():
    """Audio + transcript in, the reference CLI's JSON line out."""
    d = ssb.Decoder(model_dir("en-us"))
    d.set_align_text(TEXT["en-us"])
    pcm = open(os.path.join(DATA, "goforward.raw"), "rb").read()
    d.start_utt()
    d.process_raw(pcm[:30000], full_utt=False)
    d.process_raw(pcm[30000:], full_utt=False)
    d.end_utt()
    assert d.n_frames == 279
    assert d.dumps(align_level=1) == CLI_JSON
    assert d.hyp.text == TEXT["en-us"]
    assert abs(d.hyp.score - 1.0001 ** -2761) < 1e-12               # decoder_hyp of the default CLI
    word_level = json.loads(d.dumps(align_level=0))                   # App. B word ascr / lscr
    assert [(w["t"], w["p"]) for w in word_level["w"]] == [
        ("<sil>", round(1.0001 ** -567, 3)), ("go", round(1.0001 ** -133, 3)),
        ("forward", round(1.0001 ** -369, 3)), ("ten", round(1.0001 ** -468, 3)),
        ("meters", round(1.0001 ** -563, 3)), ("<sil>", round(1.0001 ** -661, 3))]
    text, seg = d.decode_file(os.path.join(DATA, "goforward.raw"))
    assert text == TEXT["en-us"]
    seg = list(seg)
    assert [s.text for s in seg] == ["<sil>", "go", "forward", "ten", "meters", "<sil>"]
    assert [(round(s.start, 2), round(s.duration, 2)) for s in seg][:3] == [(0.0, 0.46), (0.46, 0.18), (0.64, 0.53)]
    d.close()


@pytest.mark.gpu
@pytest.mark.parametrize("lang", ["en-us", "fr-fr"])
def test_alignment_equals_golden(golden, lang):
    d = ssb.Decoder(model_dir(lang))
    g = golden[lang]
    d.set_align_text(TEXT[lang])
    d.decode_file(os.path.join(DATA, RAW[lang]))
    al = d.alignment
    lx, m = d.lexicon, d.model
    assert [(w.name, w.start, w.duration, w.score) for w in al.words()] == \
        [(lx.wordstr(int(w[0])), w[1], w[2], w[3]) for w in g["words"].tolist()]
    assert [(p.name, p.start, p.duration, p.score) for p in al.phones()] == \
        [(m.ciname(int(p[0])), p[3], p[4], p[5]) for p in g["phones"].tolist()]
    assert [(int(s.name), s.start, s.duration, s.score) for s in al.states()] == \
        [tuple(s[:4]) for s in g["states"].tolist()]
    d.close()


@pytest.mark.gpu
def test_align_batch_equals_one_at_a_time():
    d = ssb.Decoder(model_dir("en-us"))
    pcm = np.frombuffer(open(os.path.join(DATA, "goforward.raw"), "rb").read(), np.int16)
    rs = np.random.RandomState(3)
    pcms = [pcm, pcm[:40000], (pcm + rs.randint(-40, 40, len(pcm))).astype(np.int16), pcm[:9000]]
    texts = [TEXT["en-us"], "go forward ten", TEXT["en-us"], TEXT["en-us"]]
    res = d.align_batch(pcms, texts, align_level=2)
    js = d.dumps_batch(align_level=2)
    assert res[0] is not None and js[0] is not None
    one = ssb.Decoder(model_dir("en-us"))
    n_ok = 0
    for p, t, r, j in zip(pcms, texts, res, js):
        one.set_align_text(t)
        one.start_utt(); one.process_raw(p.tobytes(), full_utt=True); one.end_utt()
        if one.hyp.text is None:
            assert r is None and j is None
            continue
        n_ok += 1
        assert r["text"] == one.hyp.text and r["seg"] == list(one.seg)
        assert j == one.dumps(align_level=2)
    assert n_ok >= 3
    assert js[0].replace(',"w":[{"b":0.000,"d":0.440,"p":1.000,"t":"96"}', "X")  # states present
    d.close(); one.close()


@pytest.mark.gpu
def test_decode_file_8khz_wav():
    """The reference's 8 kHz fixture (ref: tests/test_word_align.c:6): the sampling rate of the
    WAV header re-creates the frontend, as Decoder.decode_file does."""
    d = ssb.Decoder(model_dir("en-us"))
    d.set_align_text("he was not an ill disposed young man")
    text, seg = d.decode_file(os.path.join(DATA, "sense_and_sensibility_01_austen_64kb-0880.wav"))
    assert d.samprate == 8000
    assert text == "he was not an ill disposed young man"
    words = [s.text for s in seg if not s.text.startswith("<")]
    assert [w.split("(")[0] for w in words] == text.split()
    al = d.alignment
    ph = list(al.phones())
    assert ph[0].start == 0 and all(a.start + a.duration == b.start for a, b in zip(ph, ph[1:]))
    assert ph[-1].start + ph[-1].duration == d.n_frames - 1
    d.close()


# ---- ssb_align_texts: the same thing for a batch in one C call
def test_align_texts_unknown_word_and_no_cpu_path():
    m = ssb.AcousticModel(model_dir("en-us"), device=-1)
    lx = ssb.Lexicon(m, hmmdir=model_dir("en-us"))
    with pytest.raises(ssb.SsbError, match="Unknown word"):
        ssb.TextAlignment(m, lx, [np.zeros((5, 39), np.float32)], ["go xyzzyq"])
    with pytest.raises(ssb.SsbError, match="no CPU compute path"):
        ssb.TextAlignment(m, lx, [np.zeros((5, 39), np.float32)], ["go forward"])
    assert ssb.TextAlignment(m, lx, [], []).n == 0


@pytest.mark.gpu
def test_align_texts_c_call_equals_cli_and_decoder(golden):
    m = ssb.AcousticModel(model_dir("en-us"))
    lx = ssb.Lexicon(m, hmmdir=model_dir("en-us"))
    fe = ssb.Frontend(model_dir("en-us"))
    pcm = np.frombuffer(open(os.path.join(DATA, "goforward.raw"), "rb").read(), np.int16)
    rs = np.random.RandomState(3)
    pcms = [pcm, pcm[:40000], (pcm + rs.randint(-40, 40, len(pcm))).astype(np.int16), pcm[:9000], pcm]
    texts = [TEXT["en-us"], "go forward ten", TEXT["en-us"], TEXT["en-us"], "go forward ten meters"]
    ta = ssb.TextAlignment(m, lx, fe.run(pcms), texts, align_level=2)
    assert ta.json(0, align_level=1) == CLI_JSON and ta.json(4, align_level=1) == CLI_JSON
    assert ta.status(0) == (0, -2761, 279) and ta.hyp(0) == TEXT["en-us"]
    g = golden["en-us"]
    assert np.array_equal(ta.entries(0, "words")[:, :4], g["words"])
    assert np.array_equal(ta.entries(0, "states"), g["states"])
    assert np.array_equal(ta.entries(0, "phones")[:, [0, 1, 2, 3, 4]], g["phones"][:, [0, 3, 4, 5, 6]])
    seg = ta.entries(0, "seg")     # App. B: word sf ef ascr lscr of the default CLI
    assert seg[:, 1:].tolist() == [[0, 45, -230, -337], [46, 63, -133, 0], [64, 116, -369, 0],
                                   [117, 152, -468, 0], [153, 210, -563, 0], [211, 277, -324, -337]]
    d = ssb.Decoder(model_dir("en-us"))
    n_ok = 0
    for u, (p, t) in enumerate(zip(pcms, texts)):
        d.set_align_text(t)
        d.start_utt(); d.process_raw(p.tobytes(), full_utt=True); d.end_utt()
        if d.hyp.text is None:
            assert ta.status(u)[0] == -1 and ta.json(u) is None and ta.hyp(u) is None
            continue
        n_ok += 1
        assert ta.hyp(u) == d.hyp.text
        for lvl in (0, 1, 2):
            assert ta.json(u, start=1.5, align_level=lvl) == d.dumps(start_time=1.5, align_level=lvl), (u, lvl)
    assert n_ok >= 4 and ta.kernel_ms()["fsg_search"] > 0
    one_by_one = [ta.json(u, start=2.0, align_level=2) for u in range(len(pcms))]
    ta2 = ssb.TextAlignment(m, lx, fe.run(pcms), texts, align_level=2)
    ta2.render(start=2.0, align_level=2)      # the same lines, rendered by host threads
    assert [ta2.json(u, start=2.0, align_level=2) for u in range(len(pcms))] == one_by_one
    d.close()
