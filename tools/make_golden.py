#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libssref.so).

Run in the build container, where /root/reference exists and `make -C oracle ref`
has produced libssref.so.  The fixtures are small (features, chains, integer
results, sha256 digests of the big matrices) and are committed; the GPU box only
ever reads the fixtures.

    python tools/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.refshim import Ref, available  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "tests", "data")
MODELS = os.path.join(ROOT, "soundswallower_b200", "model")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def phone_windows(words, phones):
    """phones inherit their word's (start, duration) (ref: src/ps_alignment.c:168-305)."""
    parent = phones[:, 6]
    return words[parent, 1].astype(np.int32), words[parent, 2].astype(np.int32)


def utterance(lang, raw, text):
    hmm = os.path.join(MODELS, lang)
    pcm = np.fromfile(os.path.join(DATA, raw), np.int16)
    ref = Ref(hmm)
    refc = Ref(hmm, compallsen=True)
    g = {}
    arrays = ref.model_arrays()
    for k, v in arrays.items():
        g["model_sha_" + k] = sha(v)
    g["lut"] = arrays["lut"]
    g["tp"] = arrays["tp"]
    g["dims"] = np.array([ref.n_mgau, ref.n_feat, ref.n_density, ref.veclen, ref.n_sen, ref.n_sseq,
                          ref.n_emit, ref.n_tmat, ref.n_ciphone, ref.n_phone, ref.sil], np.int32)
    feat = ref.features_from_pcm(pcm)
    g["feat"] = feat
    g["mfcc_head"] = ref.mfcc_from_pcm(pcm)[:8]
    # 2-pass alignment exactly as the CLI does it
    al = ref.align_pcm(pcm, text)
    g["segs"] = al["segs"]
    g["words"] = al["words"]
    g["phones"] = al["phones"]
    g["states"] = al["states"]
    g["hyp_score"] = np.int32(al["hyp_score"])
    g["n_frames"] = np.int32(al["n_frames"])
    wstart, wdur = phone_windows(al["words"], al["phones"])
    g["ph_start"] = wstart
    g["ph_dur"] = wdur
    # dense senone scores, compallsen
    dense = refc.score_all(feat)
    g["senscr_sha"] = sha(dense)
    g["senscr_rows"] = np.array([0, 1, 100, feat.shape[0] - 1], np.int32)
    g["senscr_sample"] = dense[g["senscr_rows"]]
    g["senscr_argmin"] = dense.argmin(1).astype(np.int32)
    # raw top-N of the first frames (fresh history)
    refc.reset_hist()
    tn = []
    for t in range(6):
        _, topn = refc.frame_eval(feat[t], t, compallsen=True, want_topn=True)
        tn.append(topn)
    g["topn_norm_head"] = np.stack(tn)  # post-normalisation (cw, score)
    wids = al["words"][:, 0]
    chain_sen = arrays["sseq"][al["phones"][:, 1]].reshape(-1)
    for name, r, kw in (("win", ref, dict(start=al["words"][:, 1], dur=al["words"][:, 2])),
                        ("nowin", ref, dict()),
                        ("win_call", refc, dict(start=al["words"][:, 1], dur=al["words"][:, 2]))):
        res = r.state_align(feat, wids, clear_active=True, want_tokens=True, want_senscr=True, **kw)
        g[name + "_rv"] = np.int32(res["rv"])
        g[name + "_best"] = np.int32(res["best_score"])
        g[name + "_states"] = res["states"]
        g[name + "_phones"] = res["phones"]
        g[name + "_words"] = res["words"]
        g[name + "_tokens_sha"] = sha(res["tokens"])
        g[name + "_tokens_head"] = res["tokens"][:4]
        g[name + "_senscr_sha"] = sha(res["senscr"])
        g[name + "_chain_scr"] = res["senscr"][:, chain_sen]
    ref.close()
    refc.close()
    np.savez_compressed(os.path.join(OUT, "align_%s.npz" % lang), **g)
    print(lang, "frames", feat.shape[0], "phones", len(al["phones"]), "hyp", al["hyp_score"])
    return g


def synthetic(lang="en-us", seed=20261017):
    """Seeded random utterances through the reference: scoring in both modes plus
    alignment of random phone chains (no dictionary involved)."""
    hmm = os.path.join(MODELS, lang)
    ref = Ref(hmm)
    refc = Ref(hmm, compallsen=True)
    rs = np.random.RandomState(seed)
    arrays = ref.model_arrays()
    mean = arrays["mean"]
    g = {}
    feats = []
    lens = [1, 7, 40, 33]
    for u, T in enumerate(lens):
        # draw frames around randomly chosen Gaussians so the scores are in a realistic range
        cb = rs.randint(0, ref.n_mgau, T)
        dn = rs.randint(0, ref.n_density, T)
        x = np.stack([np.concatenate([mean[cb[t], f, dn[t]] for f in range(ref.n_feat)]) for t in range(T)])
        x = (x + rs.normal(0, 0.7, x.shape)).astype(np.float32)
        feats.append(x)
        g["feat%d" % u] = x
        refc.reset_hist()
        d = refc.score_all(x)
        g["senscr_sha%d" % u] = sha(d)
        g["senscr_head%d" % u] = d[:2, :64]
    # hmm_vit_eval known answers
    n_case = 256
    st_in = np.zeros((n_case, 12), np.int32)
    st_out = np.zeros((n_case, 12), np.int32)
    best = np.zeros(n_case, np.int32)
    senids = np.zeros((n_case, 3), np.uint16)
    tmats = rs.randint(0, ref.n_tmat, n_case).astype(np.int32)
    senscr = rs.randint(0, 400, (n_case, ref.n_sen)).astype(np.int16)
    W = -536870912
    for i in range(n_case):
        st = np.full(12, W, np.int32)
        st[5:10] = -1
        st[11] = -1
        k = rs.randint(0, 4)  # how many states are alive
        for j in range(min(k, 3)):
            st[j] = -int(rs.randint(0, 50000))
            st[5 + j] = int(rs.randint(0, 200))
        if rs.rand() < 0.15:
            st[rs.randint(0, 3)] = W + int(rs.randint(0, 300))  # clamp region
        if rs.rand() < 0.2 and k >= 2:
            st[1] = st[0]  # provoke ties
        if rs.rand() < 0.3:
            st[10] = -int(rs.randint(0, 50000))
            st[11] = int(rs.randint(0, 200))
        senids[i] = rs.randint(0, ref.n_sen, 3)
        st_in[i] = st
        best[i], st_out[i] = ref.hmm_vit_eval(3, int(tmats[i]), senids[i], senscr[i], st)
    g.update(hmm_st_in=st_in, hmm_st_out=st_out, hmm_best=best, hmm_senid=senids, hmm_tmat=tmats,
             hmm_senscr_seed=np.int64(seed + 1))
    # store the senscr rows only for the senones used (3 per case)
    g["hmm_senscr3"] = np.take_along_axis(senscr, senids.astype(np.int64), 1)
    ref.close()
    refc.close()
    np.savez_compressed(os.path.join(OUT, "synthetic_%s.npz" % lang), **g)
    print("synthetic", lang, lens)


def fsg(lang, raw_feat_key, text, gram):
    """FSG search (first pass): the flattened graph the reference searches, and its complete
    history table / segmentation on the test utterance with dense (compallsen) scores."""
    hmm = os.path.join(MODELS, lang)
    ref = Ref(hmm, compallsen=True)
    feat = np.load(os.path.join(OUT, "align_%s.npz" % lang))["feat"]
    jsgf = open(os.path.join(DATA, gram)).read()
    g = {}
    for name, kw in (("align", dict(align_text=text)), ("jsgf", dict(jsgf=jsgf))):
        G = ref.fsg_graph(**kw)
        for k, v in G.items():
            g["%s_%s" % (name, k)] = np.asarray(v)
        d = ref.fsg_decode(feat, **kw)
        H = ref.fsg_history()
        g[name + "_hist"] = H["hist"]
        g[name + "_segs"] = d["segs"]
        g[name + "_n_hmm_eval"] = np.int64(H["n_hmm_eval"])
        g[name + "_hyp_score"] = np.int32(H["hyp_score"])
        print(lang, name, "pnodes", len(G["pnode"]), "links", len(G["link"]), "hist", len(H["hist"]),
              "hyp", H["hyp_score"])
    ref.close()
    np.savez_compressed(os.path.join(OUT, "fsg_%s.npz" % lang), **g)


def fsg_file(lang="en-us", fsgfile="goforward.fsg"):
    """decoder_set_fsg on a text .fsg file (fsg_model_readfile: word + null transitions, null
    closure): the flattened graph, and the history table / segmentation of the search on the test
    utterance with dense scores -> fsg_file_<lang>.npz (checks ssb_fsg_build)."""
    hmm = os.path.join(MODELS, lang)
    ref = Ref(hmm, compallsen=True)
    feat = np.load(os.path.join(OUT, "align_%s.npz" % lang))["feat"]
    path = os.path.join(DATA, fsgfile)
    g = {}
    G = ref.fsg_graph(fsg_file=path)
    for k, v in G.items():
        g["file_%s" % k] = np.asarray(v)
    d = ref.fsg_decode(feat)   # (the grammar selected above)
    H = ref.fsg_history()
    g["file_hist"] = H["hist"]
    g["file_segs"] = d["segs"]
    g["file_n_hmm_eval"] = np.int64(H["n_hmm_eval"])
    g["file_hyp_score"] = np.int32(H["hyp_score"])
    print(lang, fsgfile, "pnodes", len(G["pnode"]), "links", len(G["link"]), "hist", len(H["hist"]),
          "hyp", H["hyp_score"])
    ref.close()
    np.savez_compressed(os.path.join(OUT, "fsg_file_%s.npz" % lang), **g)


def hmm5(n_case=512, seed=55):
    """hmm_vit_eval_5st_lr / _3st_lr known answers on random left-to-right transition matrices
    (self loop, next, skip; 255 = impossible) -- no bundled model has 5-state HMMs -> hmm5.npz."""
    ref = Ref(os.path.join(MODELS, "en-us"))
    rs = np.random.RandomState(seed)
    W = -536870912
    g = {}
    for E in (5, 3):
        tps = np.full((n_case, E, E + 1), 255, np.uint8)
        scr = rs.randint(0, 400, (n_case, E)).astype(np.int16)
        st_in = np.zeros((n_case, 12), np.int32)
        st_out = np.zeros((n_case, 12), np.int32)
        best = np.zeros(n_case, np.int32)
        for i in range(n_case):
            for a in range(E):
                for b in (a, a + 1, a + 2):
                    if b <= E and rs.rand() > (0.25 if b == a + 2 else 0.03):
                        tps[i, a, b] = rs.randint(0, 90)
            st = np.full(12, W, np.int32)
            st[5:10] = -1
            st[11] = -1
            k = rs.randint(0, E + 2)
            for j in range(min(k, E)):
                st[j] = -int(rs.randint(0, 50000))
                st[5 + j] = int(rs.randint(0, 200))
            if rs.rand() < 0.2:
                st[rs.randint(0, E)] = W + int(rs.randint(0, 300))   # clamp region
            if rs.rand() < 0.25 and k >= 2:
                j = rs.randint(1, min(k, E))
                st[j] = st[j - 1]                                      # provoke ties
            if rs.rand() < 0.3:
                st[10] = -int(rs.randint(0, 50000))
                st[11] = int(rs.randint(0, 200))
            st_in[i] = st
            best[i], st_out[i] = ref.hmm_vit_eval_tp(tps[i], scr[i], st)
        g.update({"tp%d" % E: tps, "senscr%d" % E: scr, "st_in%d" % E: st_in, "st_out%d" % E: st_out,
                  "best%d" % E: best})
        print("hmm", E, "states:", n_case, "cases, best range", best.min(), best.max())
    ref.close()
    np.savez_compressed(os.path.join(OUT, "hmm5.npz"), **g)


def five_state(lang="en-us"):
    """Forced alignment through the reference on a model whose HMMs have five emitting states
    (tests/model_variants.py:write_five_state_model): hmm_vit_eval_5st_lr inside
    state_align_search -> five_state_<lang>.npz."""
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import model_variants as mv
    g0 = np.load(os.path.join(OUT, "align_%s.npz" % lang))
    feat, words = g0["feat"], g0["words"]
    rs = np.random.RandomState(5)
    noisy = feat + rs.randn(*feat.shape).astype(np.float32) * np.float32(0.2)
    g = {}
    with tempfile.TemporaryDirectory() as tmp:
        d = mv.write_five_state_model(os.path.join(MODELS, lang), os.path.join(tmp, "five"))
        ref = Ref(d)
        g["tp_sha"] = sha(ref.model_arrays()["tp"])
        g["sseq_sha"] = sha(ref.model_arrays()["sseq"])
        for name, f, win in (("win", feat, True), ("nowin", feat, False), ("noisy", noisy, True),
                             ("short", feat[:120], False)):
            wids = words[:, 0] if name != "short" else words[:3, 0]
            st = words[:len(wids), 1] if win else None
            du = words[:len(wids), 2] if win else None
            a = ref.state_align(f, wids, st, du, clear_active=True, want_tokens=True)
            g[name + "_rv"] = np.int32(a["rv"])
            g[name + "_best"] = np.int32(a["best_score"])
            g[name + "_states"] = a["states"]
            g[name + "_phones"] = a["phones"]
            g[name + "_tokens_sha"] = sha(a["tokens"])
            print("five-state", lang, name, "rv", a["rv"], "best", a["best_score"], "states", len(a["states"]))
        ref.close()
        # both passes in the reference's default mode on one decoder (grammar search with
        # 5-state HMMs, then decoder_alignment from what it left in acmod and the scorer)
        text = "go forward ten meters"
        for name, f in (("fsg", feat), ("fsg_noisy", noisy)):
            ref = Ref(d, compallsen=False)
            dd = ref.fsg_decode(f, align_text=text)
            H = ref.fsg_history()
            bits, n_sen_eval = ref.active_bits()
            g[name + "_hist"] = H["hist"]
            g[name + "_segs"] = dd["segs"]
            g[name + "_hyp_score"] = np.int32(dd["hyp_score"])
            g[name + "_n_hmm_eval"] = np.int64(H["n_hmm_eval"])
            g[name + "_n_sen_eval"] = np.int64(n_sen_eval)
            g[name + "_active"] = bits
            segs = dd["segs"][dd["segs"][:, 0] >= 0]
            a = ref.state_align(f, segs[:, 0], segs[:, 1], segs[:, 2] - segs[:, 1] + 1, clear_active=False)
            g[name + "_p2_rv"] = np.int32(a["rv"])
            g[name + "_p2_states"] = a["states"]
            print("five-state", lang, name, "hist", len(H["hist"]), "hyp", dd["hyp_score"], "p2 rv", a["rv"])
            ref.close()
    np.savez_compressed(os.path.join(OUT, "five_state_%s.npz" % lang), **g)


def fsg_active_cases(lang, text, feat, jsgf):
    """(name, grammar selection, features) of the default-mode fixtures; the noisy case is
    regenerated from its seed by the tests."""
    rs = np.random.RandomState(1)
    noisy = feat + rs.randn(*feat.shape).astype(np.float32) * np.float32(0.3)
    # frames on which every Gaussian score clamps to INT32_MIN, placed where a codebook is first
    # scanned in the second pass: the lists the first pass left in the scorer decide there
    # (they are not reset between passes, ref: src/ptm_mgau.c:426-440); "odd" has an odd frame
    # count, so history slot 1 holds the second-to-last frame's lists
    mid46 = feat.copy()
    mid46[46:49] *= np.float32(3000)
    odd = feat[:-1].copy()
    odd[46:49] *= np.float32(3000)
    wide = feat.copy()
    wide[44:50] *= np.float32(300)
    return [("align", dict(align_text=text), feat), ("jsgf", dict(jsgf=jsgf), feat),
            ("trunc", dict(align_text=text), feat[:150]),
            ("noisy", dict(align_text=text), noisy),
            ("tiled", dict(align_text=" ".join([text] * 2)), np.concatenate([feat, feat])),
            ("mid46", dict(align_text=text), mid46), ("odd", dict(align_text=text), odd),
            ("wide", dict(align_text=text), wide)]


def fsg_active(lang, text, gram):
    """The same searches in the reference's DEFAULT mode (compallsen = no: only the senones of
    the active HMMs are scored, frame by frame): history table, segmentation, hypothesis score,
    senones evaluated, the active-senone flags the search leaves in acmod, and the second pass
    (decoder_alignment) that starts from those flags -> fsg_active_<lang>.npz."""
    hmm = os.path.join(MODELS, lang)
    ref = Ref(hmm, compallsen=False)
    feat = np.load(os.path.join(OUT, "align_%s.npz" % lang))["feat"]
    jsgf = open(os.path.join(DATA, gram)).read()
    g = {}
    for name, kw, f in fsg_active_cases(lang, text, feat, jsgf):
        G = ref.fsg_graph(**kw)
        if name == "tiled":
            for k, v in G.items():
                g["%s_%s" % (name, k)] = np.asarray(v)
        d = ref.fsg_decode(f, **kw)
        H = ref.fsg_history()
        bits, n_sen_eval = ref.active_bits()
        g[name + "_hist"] = H["hist"]
        g[name + "_segs"] = d["segs"]
        g[name + "_n_hmm_eval"] = np.int64(H["n_hmm_eval"])
        g[name + "_n_sen_eval"] = np.int64(n_sen_eval)
        g[name + "_hyp_score"] = np.int32(d["hyp_score"])
        g[name + "_active"] = bits
        # second pass exactly as decoder_alignment runs it: pass 1's words (null transitions
        # dropped) with their windows, acmod's flags left as pass 1 set them
        segs = d["segs"][d["segs"][:, 0] >= 0]
        if len(segs):
            a = ref.state_align(f, segs[:, 0], segs[:, 1], segs[:, 2] - segs[:, 1] + 1, clear_active=False)
            g[name + "_p2_rv"] = np.int32(a["rv"])
            for k in ("words", "phones", "states"):
                g["%s_p2_%s" % (name, k)] = a[k]
        print(lang, name, "hist", len(H["hist"]), "hyp", d["hyp_score"], "n_sen_eval", n_sen_eval,
              "flags left", int(sum(bin(int(x)).count("1") for x in bits)),
              "pass 2 rv", int(g.get(name + "_p2_rv", -9)))
    ref.close()
    np.savez_compressed(os.path.join(OUT, "fsg_active_%s.npz" % lang), **g)


def loaders(lang="en-us"):
    """Weight tables and scores of the reference for the mixture-weight formats the bundled
    models do not use (tests/model_variants.py writes the files) -> loader_variants.json."""
    import json
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import model_variants as mv
    src = os.path.join(MODELS, lang)
    base = Ref(src)
    mixw = base.model_arrays()["mixw"]
    feat = np.load(os.path.join(OUT, "synthetic_%s.npz" % lang))["feat1"]
    base.close()
    out = {"lang": lang, "n_frames": int(len(feat))}
    with tempfile.TemporaryDirectory() as tmp:
        mv.write_clustered_sendump(src, os.path.join(tmp, "clustered"), mixw)
        mv.write_float_mixw(src, os.path.join(tmp, "floatmixw"), mixw)
        for name in ("clustered", "floatmixw"):
            r = Ref(os.path.join(tmp, name), compallsen=True)
            out[name] = {"mixw": sha(r.model_arrays()["mixw"]),
                         "senscr": sha(r.score_all(feat).astype(np.int16))}
            r.close()
    with open(os.path.join(OUT, "loader_variants.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print(out)


def frontend():
    """MFCCs and dynamic features of the reference frontend (fe_process_int16/float32 + fe_end,
    feat_s2mfc2feat_live whole-utterance) for a few parameter sets -> frontend.npz."""
    import tempfile
    import model_variants as mv
    src = os.path.join(MODELS, "en-us")
    g = {}
    with tempfile.TemporaryDirectory() as tmp:
        for tag, params, sr, spec in mv.FE_CASES:
            hmm = os.path.join(MODELS, "fr-fr") if tag == "fr" else src
            d = mv.write_frontend_variant(hmm, os.path.join(tmp, tag), **params)
            r = Ref(d, samprate=sr)
            pcm = mv.fe_input(spec, sr)
            g["mfcc_" + tag] = r.mfcc_from_pcm(pcm)
            g["feat_" + tag] = r.features_from_pcm(pcm)
            if tag in ("base", "plain"):
                g["mfcc32_" + tag] = r.mfcc_from_f32((pcm.astype(np.float32) / 32768 * 0.7).astype(np.float32))
            if tag == "base":  # ragged lengths around the framing boundaries
                for n in mv.FE_LENGTHS:
                    x = mv.synthetic_pcm(n, n)
                    g["mfcc_len%d" % n] = r.mfcc_from_pcm(x)
                    g["feat_len%d" % n] = r.features_from_pcm(x) if n else np.zeros((0, 39), np.float32)
            r.close()
            print("frontend", tag, g["mfcc_" + tag].shape)
    np.savez_compressed(os.path.join(OUT, "frontend.npz"), **g)


def semi(lang="en-us"):
    """Semi-continuous (s2_semi_mgau) scoring and alignment through the reference on the
    synthetic single-codebook models of tests/model_variants.py -> semi_<lang>.npz."""
    import tempfile
    import model_variants as mv
    src = os.path.join(MODELS, lang)
    ga = np.load(os.path.join(OUT, "align_%s.npz" % lang))
    feat, words = ga["feat"], ga["words"]
    g = {}
    with tempfile.TemporaryDirectory() as tmp:
        for tag, kw, _ in mv.SEMI_CASES:
            d = os.path.join(tmp, tag)
            mv.write_semi_model(src, d, int(ga["dims"][4]), **kw)
            ref = Ref(d)
            refc = Ref(d, compallsen=True)
            assert ref.n_mgau == 1
            arrays = ref.model_arrays()
            g[tag + "_mixw_sha"] = sha(arrays["mixw"])
            g[tag + "_det_sha"] = sha(arrays["det"])
            dense = refc.score_all(feat)
            g[tag + "_senscr_sha"] = sha(dense)
            g[tag + "_senscr_rows"] = dense[[0, 1, 100, len(feat) - 1]]
            chain_sen = arrays["sseq"][ga["phones"][:, 1]].reshape(-1)
            for name, r, k2 in (("win", ref, dict(start=words[:, 1], dur=words[:, 2])),
                                ("nowin", ref, dict()),
                                ("win_call", refc, dict(start=words[:, 1], dur=words[:, 2]))):
                res = r.state_align(feat, words[:, 0], clear_active=True, want_tokens=True,
                                    want_senscr=True, **k2)
                key = "%s_%s_" % (tag, name)
                g[key + "rv"] = np.int32(res["rv"])
                g[key + "best"] = np.int32(res["best_score"])
                g[key + "states"] = res["states"]
                g[key + "tokens_sha"] = sha(res["tokens"])
                g[key + "chain_scr"] = res["senscr"][:, chain_sen]
            ref.close()
            refc.close()
            print("semi", tag, dense.shape, int(dense.min()), int(dense.max()))
    np.savez_compressed(os.path.join(OUT, "semi_%s.npz" % lang), **g)


def cont(lang="en-us"):
    """Fully continuous (ms_mgau) scoring and alignment through the reference on the synthetic
    one-codebook-per-senone models of tests/model_variants.py -> cont_<lang>.npz."""
    import tempfile
    import model_variants as mv
    src = os.path.join(MODELS, lang)
    ga = np.load(os.path.join(OUT, "align_%s.npz" % lang))
    feat, words = ga["feat"], ga["words"]
    g = {}
    with tempfile.TemporaryDirectory() as tmp:
        for tag, kw in mv.CONT_CASES:
            d = os.path.join(tmp, tag)
            mv.write_cont_model(src, d, int(ga["dims"][4]), **kw)
            ref = Ref(d)
            refc = Ref(d, compallsen=True)
            assert ref.n_mgau == ref.n_sen
            arrays = ref.model_arrays()
            g[tag + "_mixw_sha"] = sha(arrays["mixw"])
            g[tag + "_det_sha"] = sha(arrays["det"])
            dense = refc.score_all(feat)
            g[tag + "_senscr_sha"] = sha(dense)
            g[tag + "_senscr_rows"] = dense[[0, 1, 100, len(feat) - 1]]
            chain_sen = arrays["sseq"][ga["phones"][:, 1]].reshape(-1)
            for name, r, k2 in (("win", ref, dict(start=words[:, 1], dur=words[:, 2])),
                                ("nowin", ref, dict()),
                                ("win_call", refc, dict(start=words[:, 1], dur=words[:, 2]))):
                res = r.state_align(feat, words[:, 0], clear_active=True, want_tokens=True,
                                    want_senscr=True, **k2)
                key = "%s_%s_" % (tag, name)
                g[key + "rv"] = np.int32(res["rv"])
                g[key + "best"] = np.int32(res["best_score"])
                g[key + "states"] = res["states"]
                g[key + "tokens_sha"] = sha(res["tokens"])
                g[key + "tokens"] = res["tokens"]
                if name == "win_call":   # (with active lists the reference's buffer keeps stale
                    g[key + "chain_scr"] = res["senscr"][:, chain_sen]   # values of other senones)
            ref.close()
            refc.close()
            print("cont", tag, dense.shape, int(dense.min()), int(dense.max()))
    np.savez_compressed(os.path.join(OUT, "cont_%s.npz" % lang), **g)


def lexicon():
    """Word ids and alignment_populate phone chains of the reference for seeded random word
    sequences (both bundled dictionaries) -> lexicon.npz."""
    g = {}
    for lang in ("en-us", "fr-fr"):
        ref = Ref(os.path.join(MODELS, lang))
        rs = np.random.RandomState(77)
        n = ref.wordid("<sil>") + 1
        while ref.wordstr(n) is not None:
            n += 1
        g[lang + "_size"] = np.int32(n)
        probe = np.unique(np.concatenate([rs.randint(0, n, 48), np.arange(n - 8, n)])).astype(np.int32)
        g[lang + "_probe_wid"] = probe
        g[lang + "_probe_str"] = np.array([ref.wordstr(int(w)) for w in probe])
        seqs, off, ph = [], [0], []
        for _ in range(200):
            wids = rs.randint(0, n, rs.randint(1, 14)).astype(np.int32)
            seqs.append(wids)
            off.append(off[-1] + len(wids))
            ph.append(ref.populate(wids)["phones"][:, [0, 1, 2, 6]])
        g[lang + "_wids"] = np.concatenate(seqs)
        g[lang + "_wid_off"] = np.array(off, np.int32)
        g[lang + "_phones"] = np.concatenate(ph).astype(np.int32)   # ci ssid tmat parent
        g[lang + "_n_phones"] = np.array([len(p) for p in ph], np.int32)
        ref.close()
        print("lexicon", lang, n, len(g[lang + "_phones"]))
    np.savez_compressed(os.path.join(OUT, "lexicon.npz"), **g)


def fsg_partial():
    """decoder_hyp / decoder_seg_iter BETWEEN steps of the grammar search (find_exit with final =
    FALSE, ref: src/fsg_search.c:853-960), default mode, both models -> fsg_partial.npz."""
    g = {}
    for lang, text in (("en-us", "go forward ten meters"), ("fr-fr", "avance de dix mètres")):
        ref = Ref(os.path.join(MODELS, lang))
        feat = np.load(os.path.join(OUT, "align_%s.npz" % lang))["feat"]
        T = feat.shape[0]
        stops = sorted(set([1, 5, 20, 50, 60, 100, 150, 200, 250, T - 1, T]) & set(range(1, T + 1)))
        res = ref.fsg_partial(feat, text, stops)
        g["%s_stops" % lang] = np.asarray(stops, np.int32)
        g["%s_hyp" % lang] = np.asarray([x["hyp"] if x["hyp"] is not None else "" for x in res])
        g["%s_score" % lang] = np.asarray([x["hyp_score"] if x["hyp_score"] is not None else 0 for x in res], np.int32)
        segs = np.full((len(stops), 16, 5), -9, np.int32)
        nseg = np.zeros(len(stops), np.int32)
        for k, x in enumerate(res):
            nseg[k] = len(x["segs"])
            segs[k, :nseg[k]] = x["segs"]
        g["%s_segs" % lang] = segs
        g["%s_nseg" % lang] = nseg
        print(lang, [(t, x["hyp"], x["hyp_score"]) for t, x in zip(stops, res)])
        ref.close()
    np.savez_compressed(os.path.join(OUT, "fsg_partial.npz"), **g)


def main():
    if not available():
        raise SystemExit("oracle/_ref/libssref.so missing: run `make -C oracle ref` first")
    os.makedirs(OUT, exist_ok=True)
    if "--loaders" in sys.argv:
        return loaders()
    if "--frontend" in sys.argv:
        return frontend()
    if "--semi" in sys.argv:
        return semi()
    if "--cont" in sys.argv:
        return cont()
    if "--lexicon" in sys.argv:
        return lexicon()
    if "--fsg-file" in sys.argv:
        return fsg_file()
    if "--fsg-partial" in sys.argv:
        return fsg_partial()
    if "--hmm5" in sys.argv:
        return hmm5()
    if "--five-state" in sys.argv:
        return five_state()
    if "--fsg-active" in sys.argv:
        fsg_active("en-us", "go forward ten meters", "goforward.gram")
        return fsg_active("fr-fr", "avance de dix mètres", "goforward_fr.gram")
    utterance("en-us", "goforward.raw", "go forward ten meters")
    utterance("fr-fr", "goforward_fr.raw", "avance de dix mètres")
    synthetic("en-us")
    fsg("en-us", "goforward.raw", "go forward ten meters", "goforward.gram")
    fsg("fr-fr", "goforward_fr.raw", "avance de dix mètres", "goforward_fr.gram")
    fsg_active("en-us", "go forward ten meters", "goforward.gram")
    fsg_active("fr-fr", "avance de dix mètres", "goforward_fr.gram")
    fsg_file()
    fsg_partial()
    hmm5()
    five_state()
    loaders()
    frontend()
    semi()
    cont()
    lexicon()


if __name__ == "__main__":
    main()
