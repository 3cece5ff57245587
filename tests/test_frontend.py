"""Acoustic frontend: PCM -> MFCC -> CMN -> 1s_c_d_dd.

tests/golden/frontend.npz holds what the compiled reference produced
(tools/make_golden.py --frontend) for the parameter sets of model_variants.FE_CASES.

Tolerance: everything but the natural logarithm is evaluated with IEEE operations in the
reference's order, so results normally agree to the bit; log() differs between libm builds
and CUDA by at most an ulp of a float64, which can (rarely) move a float32 cepstrum by one
ulp.  The bound asserted is 2e-4 absolute on cepstra of magnitude 1..100 -- ten times tighter
than the reference's own frontend regression tolerance (test_fe.c compares to 0.002, SURVEY
§4) -- and the number of frames that are not bit-identical is reported and bounded.
"""
import os

import numpy as np
import pytest

import model_variants as mv
from conftest import GOLDEN, model_dir

TOL = 2e-4


@pytest.fixture(scope="module")
def fe_golden():
    return np.load(os.path.join(GOLDEN, "frontend.npz"))


@pytest.fixture(scope="module")
def fe_dirs(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("fe"))
    out = {}
    for tag, params, sr, spec in mv.FE_CASES:
        src = model_dir("fr-fr" if tag == "fr" else "en-us")
        out[tag] = (mv.write_frontend_variant(src, os.path.join(root, tag), **params), sr, spec)
    return out


def close(a, b, what):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.size:
        err = np.abs(a.astype(np.float64) - b).max()
        assert err <= TOL, (what, err)
    return int((a != b).any(axis=-1).sum()) if a.size else 0


# ------------------------------------------------------------------ oracle vs reference goldens
@pytest.mark.parametrize("tag", [c[0] for c in mv.FE_CASES])
def test_oracle_frontend_matches_reference(fe_golden, fe_dirs, tag):
    from oracle.oracle import OracleFrontend, fe_config
    d, sr, spec = fe_dirs[tag]
    fe = OracleFrontend(fe_config(d, samprate=sr))
    pcm = mv.fe_input(spec, sr)
    mfcc, feat = fe.features(pcm)
    bad = close(mfcc, fe_golden["mfcc_" + tag], "mfcc") + close(feat, fe_golden["feat_" + tag], "feat")
    assert bad == 0 or bad <= len(mfcc) // 50
    if "mfcc32_" + tag in fe_golden.files:
        x = (pcm.astype(np.float32) / 32768 * 0.7).astype(np.float32)
        close(fe.mfcc(x), fe_golden["mfcc32_" + tag], "mfcc f32")
    fe.close()


def test_oracle_frontend_ragged_lengths(fe_golden, fe_dirs):
    from oracle.oracle import OracleFrontend, fe_config
    fe = OracleFrontend(fe_config(fe_dirs["base"][0], samprate=16000))
    for n in mv.FE_LENGTHS:
        x = mv.synthetic_pcm(n, n)
        assert fe.n_frames(n) == len(fe_golden["mfcc_len%d" % n])
        mfcc, feat = fe.features(x)
        close(mfcc, fe_golden["mfcc_len%d" % n], "mfcc len %d" % n)
        close(feat, fe_golden["feat_len%d" % n], "feat len %d" % n)
    fe.close()


# ------------------------------------------------------------------ product, host side (no GPU)
@pytest.mark.parametrize("tag", [c[0] for c in mv.FE_CASES])
def test_host_tables_equal_oracle(fe_dirs, tag):
    """feat_params.json parsing + filter/DCT/window tables, bit for bit."""
    import soundswallower_b200 as ssb
    from oracle.oracle import OracleFrontend, fe_config
    d, sr, _ = fe_dirs[tag]
    fe = ssb.Frontend(d, device=-1, samprate=sr)
    o = OracleFrontend(fe_config(d, samprate=sr))
    assert (fe.frame_size, fe.frame_shift, fe.fft_size, fe.n_coeffs) == \
        (o.frame_size, o.frame_shift, o.fft_size, o.n_coeffs)
    a, b = fe.tables(), o.tables()
    for k in b:
        assert np.array_equal(a[k], b[k]), k
    for n in list(mv.FE_LENGTHS) + [44580, 160000, 57600000]:
        assert fe.n_frames(n) == o.n_frames(n)
    fe.close()
    o.close()


def test_unsupported_frontend_parameters_fail_loudly(tmp_path):
    import soundswallower_b200 as ssb
    src = model_dir("en-us")
    for i, params in enumerate([dict(feat="s3_1x39"), dict(cmn="live"), dict(svspec="0-19/20-38"),
                                dict(dither=True), dict(logspec=True), dict(transform="mystery"),
                                dict(warp_type="affine"), dict(agc="max")]):
        d = mv.write_frontend_variant(src, str(tmp_path / ("v%d" % i)), **params)
        with pytest.raises(ssb.SsbError):
            ssb.Frontend(d, device=-1)
    for params in [dict(nfft=500), dict(nfft=256), dict(wlen=0.001), dict(upperf=9000.0),
                   dict(nfilt=100), dict(ncep=40), dict(frate=0)]:
        with pytest.raises(ssb.SsbError):
            ssb.Frontend(src, device=-1, **params)
    fe = ssb.Frontend(src, device=-1)
    with pytest.raises(ssb.SsbError):     # no device: tables only, no CPU path
        fe.run([np.zeros(1000, np.int16)])


# ------------------------------------------------------------------ GPU parity
@pytest.mark.gpu
@pytest.mark.parametrize("tag", [c[0] for c in mv.FE_CASES])
def test_gpu_frontend_matches_reference(fe_golden, fe_dirs, tag):
    import soundswallower_b200 as ssb
    d, sr, spec = fe_dirs[tag]
    fe = ssb.Frontend(d, device=0, samprate=sr)
    pcm = mv.fe_input(spec, sr)
    mfcc, feat = fe.features([pcm])[0]
    bad = close(mfcc, fe_golden["mfcc_" + tag], "mfcc") + close(feat, fe_golden["feat_" + tag], "feat")
    assert bad <= max(1, len(mfcc) // 50), bad
    if "mfcc32_" + tag in fe_golden.files:
        x = (pcm.astype(np.float32) / 32768 * 0.7).astype(np.float32)
        close(fe.features([x])[0][0], fe_golden["mfcc32_" + tag], "mfcc f32")
    fe.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["base", "wav8k", "nfft1024_wlen"])
def test_gpu_generic_melspec_kernel_agrees(fe_golden, fe_dirs, tag, monkeypatch):
    """SSB_FE=generic forces the shared-memory-only mel spectrum kernel (the one used for
    remove_dc and for FFT sizes other than 256/512/1024): same bits as the register kernel."""
    import soundswallower_b200 as ssb
    d, sr, spec = fe_dirs[tag]
    pcm = mv.fe_input(spec, sr)
    fast = ssb.Frontend(d, device=0, samprate=sr)
    a = fast.features([pcm])[0]
    monkeypatch.setenv("SSB_FE", "generic")
    slow = ssb.Frontend(d, device=0, samprate=sr)
    b = slow.features([pcm])[0]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    close(a[0], fe_golden["mfcc_" + tag], "mfcc")
    fast.close()
    slow.close()


@pytest.mark.gpu
def test_gpu_frontend_ragged_batch(fe_golden, fe_dirs):
    """One batch holding every boundary length (incl. the empty utterance) plus real audio."""
    import soundswallower_b200 as ssb
    fe = ssb.Frontend(fe_dirs["base"][0], device=0, samprate=16000)
    pcms = [mv.synthetic_pcm(n, n) for n in mv.FE_LENGTHS] + [mv.fe_input("goforward.raw", 16000)]
    want = [(fe_golden["mfcc_len%d" % n], fe_golden["feat_len%d" % n]) for n in mv.FE_LENGTHS] + \
        [(fe_golden["mfcc_base"], fe_golden["feat_base"])]
    dev = fe.run(pcms)
    assert [int(x) for x in np.diff(dev.frame_off)] == [len(w[0]) for w in want]
    for (mfcc, feat), (wm, wf) in zip(fe.download(), want):
        close(mfcc, wm, "mfcc")
        close(feat, wf, "feat")
    ms = fe.kernel_ms()
    assert ms["total"] > 0
    fe.close()


@pytest.mark.gpu
def test_gpu_frontend_equals_oracle_on_random_batch(fe_dirs):
    from oracle.oracle import OracleFrontend, fe_config
    import soundswallower_b200 as ssb
    rs = np.random.RandomState(5)
    d = fe_dirs["base"][0]
    fe = ssb.Frontend(d, device=0, samprate=16000)
    o = OracleFrontend(fe_config(d, samprate=16000))
    pcms = [mv.synthetic_pcm(int(n), 100 + i) for i, n in enumerate(rs.randint(0, 40000, 37))]
    got = fe.features(pcms)
    n_bad = n_all = 0
    for x, (mfcc, feat) in zip(pcms, got):
        wm, wf = o.features(x)
        n_bad += close(mfcc, wm, "mfcc") + close(feat, wf, "feat")
        n_all += len(wm)
    assert n_bad <= max(1, n_all // 100), (n_bad, n_all)
    fe.close()
    o.close()


@pytest.mark.gpu
@pytest.mark.parametrize("lang", ["en-us", "fr-fr"])
def test_audio_to_alignment_on_gpu_equals_cli(models, golden, lang):
    """BASELINE config #1 from raw audio, everything on the GPU: frontend -> (features stay in
    HBM) -> FSG pass 1 -> word windows -> chain Viterbi pass 2.  Word boundaries and the state
    segmentation must be the reference CLI's (tests/golden/align_*.npz)."""
    import soundswallower_b200 as ssb
    from test_gpu_fsg import graph_of
    m, g = models(lang), golden[lang]
    fg = np.load(os.path.join(GOLDEN, "fsg_%s.npz" % lang))
    raw = "goforward.raw" if lang == "en-us" else "goforward_fr.raw"
    fe = ssb.Frontend(model_dir(lang), device=0, samprate=16000)
    dev = fe.run([mv.fe_input(raw, 16000)])
    assert int(dev.frame_off[-1]) == len(g["feat"])
    p1 = ssb.fsg_batch(m, dev, [graph_of(fg, "align")])[0]
    assert p1["rv"] == 0 and p1["exit"] > 0
    segs = p1["segs"]
    w_start, w_dur = segs[:, 1], segs[:, 2] - segs[:, 1] + 1
    assert np.array_equal(w_start, g["words"][:, 1]) and np.array_equal(w_dur, g["words"][:, 2])
    parent = g["phones"][:, 6]
    sf, ef = ssb.windows(w_start[parent], w_dur[parent])
    chain = dict(ssid=g["phones"][:, 1].astype(np.int32), tmat=g["phones"][:, 2].astype(np.int32),
                 sf=sf, ef=ef)
    p2 = ssb.align_batch(m, dev, [chain])[0]
    st = g["states"]
    assert p2["rv"] == 0
    assert np.array_equal(p2["start"], st[:, 1]) and np.array_equal(p2["dur"], st[:, 2])
    assert np.array_equal(p2["score"], st[:, 3])
    fe.close()


@pytest.mark.gpu
@pytest.mark.parametrize("lang,raw,text", [("en-us", "goforward.raw", "go forward ten meters"),
                                           ("fr-fr", "goforward_fr.raw", "avance de dix mètres")])
def test_cli_alignment_from_audio_and_text(models, golden, lang, raw, text):
    """BASELINE config #1 with nothing but the audio, the transcript and the model directory:
    frontend -> alignment grammar (host) -> grammar search -> word windows -> phone chain
    (host) -> chain Viterbi -> propagate.  Words, phones and states equal the reference CLI's
    (tests/golden/align_*.npz = SURVEY Appendix A/B)."""
    import soundswallower_b200 as ssb
    m, g = models(lang), golden[lang]
    lx = ssb.Lexicon(m, hmmdir=model_dir(lang))
    fe = ssb.Frontend(model_dir(lang), device=0, samprate=16000)
    pcm = mv.fe_input(raw, 16000)
    dev = fe.run([pcm, pcm[:len(pcm) // 3]])          # the truncated copy cannot match
    res = ssb.align_texts(m, lx, dev, [text, text])
    assert res[1] is None
    r = res[0]
    # (pass 1 runs on dense "compallsen" scores: same search decisions and word boundaries as
    # the CLI's default mode, but its path score is normalised over all senones, so hyp_score
    # is the compallsen=yes one -- tests/test_gpu_fsg.py pins that value)
    gw = g["words"]
    assert [w[0] for w in r["words"]] == [lx.wordstr(int(w)) for w in gw[:, 0]]
    assert [list(w[1:]) for w in r["words"]] == gw[:, 1:4].tolist()
    gp = g["phones"]
    assert [m.ciname(int(c)) for c in gp[:, 0]] == [p[0] for p in r["phones"]]
    assert [list(p[1:]) for p in r["phones"]] == gp[:, [3, 4, 5, 6]].tolist()
    assert np.array_equal(r["states"], g["states"])
    fe.close()
