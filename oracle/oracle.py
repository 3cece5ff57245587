"""ctypes wrapper around oracle/liboracle.so (the plain-C restatement).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg -- never from the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
INT_MAX = 2**31 - 1


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("ss_oracle.c", "ss_oracle.h", "ss_oracle_fsg.c", "ss_oracle_fe.c", "ss_oracle_cont.c")]
    srcs = [s for s in srcs if os.path.exists(s)]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", HERE, "liboracle"])
    return LIB


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


class AlignOut(C.Structure):
    _fields_ = [("rv", C.c_int32), ("best_score", C.c_int32), ("n_renorm", C.c_int32)]


class Oracle:
    def __init__(self, model_dir, logbase=1.0001, varfloor=1e-4, tmatfloor=1e-4, topn=4):
        build()
        self.lib = L = C.CDLL(LIB)
        L.orc_model_load.restype = C.c_void_p
        L.orc_model_load.argtypes = [C.c_char_p, C.c_double, C.c_float, C.c_double]
        L.orc_model_free.argtypes = [C.c_void_p]
        L.orc_ptm_new.restype = C.c_void_p
        L.orc_ptm_new.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_ptm_free.argtypes = [C.c_void_p]
        L.orc_ptm_reset.argtypes = [C.c_void_p]
        L.orc_logmath_log.restype = C.c_int32
        L.orc_logmath_log.argtypes = [C.c_double, C.c_int, C.c_double]
        h = L.orc_model_load(model_dir.encode(), logbase, varfloor, tmatfloor)
        if not h:
            raise RuntimeError("oracle model load failed: " + model_dir)
        self.h = C.c_void_p(h)
        self.topn = topn
        dims = np.zeros(16, np.int32)
        L.orc_model_dims(self.h, _p(dims, C.c_int32))
        (self.n_mgau, self.n_feat, self.n_density, self.veclen, self.n_sen, self.n_sseq,
         self.n_emit, self.n_tmat, self.n_ciphone, self.n_phone, self.sil) = [int(x) for x in dims[:11]]
        self.D = self.n_feat * self.veclen
        self._arrays = None

    def close(self):
        if self.h:
            self.lib.orc_model_free(self.h)
            self.h = None

    def set_topn_beam(self, beam):
        """s2_semi "topn_beam" (comma list in the reference's config), per stream."""
        b = np.ascontiguousarray(beam, np.int32)
        self.lib.orc_model_set_topn_beam.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        self.lib.orc_model_set_topn_beam(self.h, b.ctypes.data, len(b))

    def model_arrays(self):
        if self._arrays is None:
            mean = np.zeros((self.n_mgau, self.n_feat, self.n_density, self.veclen), np.float32)
            var = np.zeros_like(mean)
            det = np.zeros((self.n_mgau, self.n_feat, self.n_density), np.float32)
            mixw = np.zeros((self.n_feat, self.n_density, self.n_sen), np.uint8)
            sen2cb = np.zeros(self.n_sen, np.uint8)
            tp = np.zeros((self.n_tmat, self.n_emit, self.n_emit + 1), np.uint8)
            sseq = np.zeros((self.n_sseq, self.n_emit), np.uint16)
            lut = np.zeros(256, np.uint8)
            self.lib.orc_model_copy(self.h, _p(mean, C.c_float), _p(var, C.c_float), _p(det, C.c_float),
                                    _p(mixw, C.c_uint8), _p(sen2cb, C.c_uint8), _p(tp, C.c_uint8),
                                    _p(sseq, C.c_uint16), _p(lut, C.c_uint8))
            self._arrays = dict(mean=mean, var=var, det=det, mixw=mixw, sen2cb=sen2cb, tp=tp,
                                sseq=sseq, lut=lut)
        return self._arrays

    def phone_table(self):
        ssid = np.zeros(self.n_phone, np.int32)
        tmat = np.zeros(self.n_phone, np.int32)
        ci = np.zeros(self.n_phone, np.int32)
        self.lib.orc_phone_table(self.h, _p(ssid, C.c_int32), _p(tmat, C.c_int32), _p(ci, C.c_int32))
        return ssid, tmat, ci

    @staticmethod
    def logadd_table8(base=1.0001, shift=10):
        build()
        L = C.CDLL(LIB)
        out = np.zeros(256, np.uint8)
        L.orc_logadd_table8.argtypes = [C.c_double, C.c_int, C.POINTER(C.c_uint8)]
        rv = L.orc_logadd_table8(base, shift, _p(out, C.c_uint8))
        return rv, out

    # ---- scorer
    def score_all(self, feat):
        feat = np.ascontiguousarray(feat, np.float32)
        T = feat.shape[0]
        out = np.zeros((T, self.n_sen), np.int16)
        self.lib.orc_ptm_score_all(self.h, self.topn, _p(feat, C.c_float), T, _p(out, C.c_int16))
        return out

    def topn_all(self, feat):
        feat = np.ascontiguousarray(feat, np.float32)
        T = feat.shape[0]
        cw = np.zeros((T, self.n_mgau, self.n_feat, self.topn), np.uint8)
        sc = np.zeros((T, self.n_mgau, self.n_feat, self.topn), np.int32)
        self.lib.orc_ptm_topn_all(self.h, self.topn, _p(feat, C.c_float), T, _p(cw, C.c_uint8),
                                  _p(sc, C.c_int32))
        return cw, sc

    def new_ptm(self):
        return C.c_void_p(self.lib.orc_ptm_new(self.h, self.topn, 1))

    def set_frame_idx(self, p, v):
        self.lib.orc_ptm_set_frame_idx.argtypes = [C.c_void_p, C.c_int]
        self.lib.orc_ptm_set_frame_idx(p, int(v))

    def reset_ptm(self, p):
        self.lib.orc_ptm_reset(p)

    def free_ptm(self, p):
        self.lib.orc_ptm_free(p)

    def frame_eval(self, ptm, feat, frame, active=None, compallsen=True, want_topn=False):
        feat = np.ascontiguousarray(feat, np.float32).reshape(-1)
        out = np.zeros(self.n_sen, np.int16)
        if active is None:
            active = np.zeros(1, np.uint8)
            n_active = 0
        else:
            active = np.ascontiguousarray(active, np.uint8)
            n_active = len(active)
        topn = np.zeros((self.n_mgau, self.n_feat, self.topn, 2), np.int32) if want_topn else None
        self.lib.orc_ptm_frame_eval(ptm, _p(out, C.c_int16), _p(active, C.c_uint8), n_active,
                                    _p(feat, C.c_float), int(frame), int(compallsen),
                                    _p(topn, C.c_int32))
        return (out, topn) if want_topn else out

    def flags2list(self, senones):
        bits = np.zeros((self.n_sen + 31) // 32, np.uint32)
        for s in senones:
            bits[s >> 5] |= np.uint32(1 << (s & 31))
        out = np.zeros(self.n_sen + 64, np.uint8)
        n = self.lib.orc_flags2list(_p(bits, C.c_uint32), self.n_sen, _p(out, C.c_uint8))
        return out[:n].copy()

    # ---- HMM / aligner
    def hmm_eval(self, n_emit, tp, senid, senscr, st):
        tp = np.ascontiguousarray(tp, np.uint8)
        senid = np.ascontiguousarray(senid, np.uint16)
        senscr = np.ascontiguousarray(senscr, np.int16)
        st = np.ascontiguousarray(st, np.int32).copy()
        self.lib.orc_hmm_eval.restype = C.c_int32
        best = self.lib.orc_hmm_eval(n_emit, _p(tp, C.c_uint8), _p(senid, C.c_uint16),
                                     _p(senscr, C.c_int16), _p(st, C.c_int32))
        return best, st

    @staticmethod
    def windows(start, dur):
        """state_align_search_init's sf/ef rule (ref: state_align_search.c:464-471)."""
        start = np.asarray(start, np.int32)
        dur = np.asarray(dur, np.int32)
        sf = np.where(start > 0, start, 0).astype(np.int32)
        ef = np.where(dur > 0, start + dur, INT_MAX).astype(np.int32)
        return sf, ef

    def _al_common(self, n_phones, T, want_tokens):
        ns = n_phones * self.n_emit
        return (np.zeros(ns, np.int32), np.zeros(ns, np.int32), np.zeros(ns, np.int32),
                np.zeros((T, ns, 2), np.int32) if want_tokens else None, AlignOut())

    def state_align_dense(self, senscr, ssid, tmat, sf, ef, want_tokens=False):
        senscr = np.ascontiguousarray(senscr, np.int16)
        T = senscr.shape[0]
        ssid, tmat, sf, ef = [np.ascontiguousarray(a, np.int32) for a in (ssid, tmat, sf, ef)]
        n = len(ssid)
        st, du, sc, tok, out = self._al_common(n, T, want_tokens)
        self.lib.orc_state_align_dense(self.h, _p(senscr, C.c_int16), T, n, _p(ssid, C.c_int32),
                                       _p(tmat, C.c_int32), _p(sf, C.c_int32), _p(ef, C.c_int32),
                                       _p(st, C.c_int32), _p(du, C.c_int32), _p(sc, C.c_int32),
                                       _p(tok, C.c_int32), C.byref(out))
        return dict(rv=out.rv, best_score=out.best_score, n_renorm=out.n_renorm, start=st, dur=du,
                    score=sc, tokens=tok)

    def state_align(self, feat, ssid, tmat, sf, ef, init_active=None, compallsen=False,
                    want_tokens=False, want_senscr=False, init_topn=None):
        feat = np.ascontiguousarray(feat, np.float32)
        T = feat.shape[0]
        ssid, tmat, sf, ef = [np.ascontiguousarray(a, np.int32) for a in (ssid, tmat, sf, ef)]
        n = len(ssid)
        st, du, sc, tok, out = self._al_common(n, T, want_tokens)
        bits = None
        if init_active is not None:
            bits = np.zeros((self.n_sen + 31) // 32, np.uint32)
            for s in init_active:
                bits[s >> 5] |= np.uint32(1 << (s & 31))
        senscr = np.zeros((T, self.n_sen), np.int16) if want_senscr else None
        it = np.ascontiguousarray(init_topn, np.uint8) if init_topn is not None else None
        self.lib.orc_state_align2(self.h, self.topn, _p(feat, C.c_float), T, n, _p(ssid, C.c_int32),
                                  _p(tmat, C.c_int32), _p(sf, C.c_int32), _p(ef, C.c_int32),
                                  _p(bits, C.c_uint32), int(compallsen), _p(st, C.c_int32),
                                  _p(du, C.c_int32), _p(sc, C.c_int32), _p(tok, C.c_int32),
                                  _p(senscr, C.c_int16), C.byref(out), _p(it, C.c_uint8))
        return dict(rv=out.rv, best_score=out.best_score, n_renorm=out.n_renorm, start=st, dur=du,
                    score=sc, tokens=tok, senscr=senscr)

    # ---- FSG search
    class _Fsg(C.Structure):
        _fields_ = [(k, C.c_int32) for k in ("n_state", "start", "final", "n_link", "n_pnode",
                                              "n_ciphone", "sil", "beam", "pbeam", "wbeam", "maxhmmpf")] + \
                   [("link4", C.c_void_p), ("link_flag", C.c_void_p), ("arc_off", C.c_void_p),
                    ("root", C.c_void_p), ("pnode8", C.c_void_p), ("ctxt", C.c_void_p)]

    def _fsg_struct(self, G):
        keep = dict(link=np.ascontiguousarray(G["link"], np.int32),
                    link_flag=np.ascontiguousarray(G["link_flag"], np.uint8),
                    arc_off=np.ascontiguousarray(G["arc_off"], np.int32),
                    root=np.ascontiguousarray(G["root"], np.int32),
                    pnode=np.ascontiguousarray(G["pnode"], np.int32),
                    ctxt=np.ascontiguousarray(G["ctxt"], np.uint32))
        f = self._Fsg()
        f.n_state, f.start, f.final = int(G["n_state"]), int(G["start"]), int(G["final"])
        f.n_link, f.n_pnode = len(keep["link"]), len(keep["pnode"])
        f.n_ciphone, f.sil = int(G["n_ciphone"]), int(G["sil"])
        f.beam, f.pbeam, f.wbeam, f.maxhmmpf = (int(G[k]) for k in ("beam", "pbeam", "wbeam", "maxhmmpf"))
        f.link4, f.link_flag, f.arc_off = (keep[k].ctypes.data for k in ("link", "link_flag", "arc_off"))
        f.root, f.pnode8, f.ctxt = (keep[k].ctypes.data for k in ("root", "pnode", "ctxt"))
        return f, keep

    def fsg_search(self, G, senscr, cap=1 << 16):
        """Token passing over the flattened graph G (refshim.Ref.fsg_graph layout) on dense senone
        scores [T][n_sen].  Returns history [n][9], counters, best exit, segs [n][5]
        (link sf ef ascr lscr)."""
        senscr = np.ascontiguousarray(senscr, np.int16)
        T = senscr.shape[0]
        f, keep = self._fsg_struct(G)
        hist = np.zeros((cap, 9), np.int32)
        out = np.zeros(8, np.int64)
        L = self.lib
        L.orc_fsg_search.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                     C.c_void_p]
        rv = L.orc_fsg_search(self.h, C.byref(f), senscr.ctypes.data, T, hist.ctypes.data, cap,
                              out.ctypes.data)
        n = int(out[0])
        hist = hist[:n].copy()
        score = C.c_int32(0)
        L.orc_fsg_find_exit.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        bp = L.orc_fsg_find_exit(C.byref(f), hist.ctypes.data, n, int(out[2]), 1, C.byref(score))
        segs = np.zeros((4096, 5), np.int32)
        ns = 0
        if bp > 0:
            L.orc_fsg_segs.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
            ns = L.orc_fsg_segs(C.byref(f), hist.ctypes.data, bp, segs.ctypes.data, 4096)
        del keep
        return dict(rv=rv, hist=hist, n_hmm_eval=int(out[1]), n_frames=int(out[2]), exit=bp,
                    hyp_score=int(score.value), segs=segs[:max(ns, 0)].copy())

    def fsg_search_active(self, G, feat, topn=4, cap=1 << 16):
        """The same search in the reference's default mode (compallsen = no): scores computed
        frame by frame for the active HMMs' senones.  Adds `active` (acmod's flags after the
        last frame, uint32 words) and n_sen_eval."""
        feat = np.ascontiguousarray(feat, np.float32)
        T = feat.shape[0]
        f, keep = self._fsg_struct(G)
        hist = np.zeros((cap, 9), np.int32)
        out = np.zeros(8, np.int64)
        active = np.zeros((self.n_sen + 31) // 32, np.uint32)
        L = self.lib
        carried = np.zeros((self.n_mgau * self.n_feat, topn), np.uint8)
        L.orc_fsg_search_active2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                             C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        rv = L.orc_fsg_search_active2(self.h, topn, C.byref(f), feat.ctypes.data, T, hist.ctypes.data,
                                      cap, out.ctypes.data, active.ctypes.data, carried.ctypes.data)
        n = int(out[0])
        hist = hist[:n].copy()
        score = C.c_int32(0)
        L.orc_fsg_find_exit.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        bp = L.orc_fsg_find_exit(C.byref(f), hist.ctypes.data, n, int(out[2]), 1, C.byref(score))
        segs = np.zeros((4096, 5), np.int32)
        ns = 0
        if bp > 0:
            L.orc_fsg_segs.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
            ns = L.orc_fsg_segs(C.byref(f), hist.ctypes.data, bp, segs.ctypes.data, 4096)
        del keep
        return dict(rv=rv, hist=hist, n_hmm_eval=int(out[1]), n_frames=int(out[2]), exit=bp,
                    hyp_score=int(score.value), segs=segs[:max(ns, 0)].copy(), active=active,
                    n_sen_eval=int(out[3]), carried=carried)

    def propagate(self, start, dur, score):
        n = len(start)
        E = self.n_emit
        ps, pd, pc = (np.zeros(n // E, np.int32) for _ in range(3))
        self.lib.orc_propagate(n, E, _p(np.ascontiguousarray(start, np.int32), C.c_int32),
                               _p(np.ascontiguousarray(dur, np.int32), C.c_int32),
                               _p(np.ascontiguousarray(score, np.int32), C.c_int32),
                               _p(ps, C.c_int32), _p(pd, C.c_int32), _p(pc, C.c_int32))
        return ps, pd, pc


class FeCfg(C.Structure):
    """orc_fe_cfg_t; defaults = config_defs.h as the non-web build has them."""
    _fields_ = [(k, C.c_int32) for k in ("samprate", "frate", "ncep", "nfft", "nfilt", "lifter",
                                          "remove_dc", "remove_noise", "unit_area", "round_filters",
                                          "doublebw", "transform", "cmn", "varnorm")] + \
               [(k, C.c_float) for k in ("wlen", "alpha", "lowerf", "upperf")]


TRANSFORMS = {"dct": 0, "legacy": 1, "htk": 2}
CMN_TYPES = {"none": 0, "batch": 1, "current": 1}


def fe_config(model_dir=None, **kw):
    """Frontend parameters: reference defaults, overridden by the model's feat_params.json
    (ref: src/acmod.c / config.c expansion of -hmm), overridden by keywords."""
    import json
    d = dict(samprate=16000, frate=100, ncep=13, nfft=0, nfilt=40, lifter=0, remove_dc=0,
             remove_noise=0, unit_area=1, round_filters=1, doublebw=0, transform="legacy",
             cmn="batch", varnorm=0, wlen=0.025625, alpha=0.97, lowerf=133.33334, upperf=6855.4976)
    if model_dir:
        with open(os.path.join(model_dir, "feat_params.json")) as fh:
            for k, v in json.load(fh).items():
                if k in d:
                    d[k] = v
    d.update(kw)
    c = FeCfg()
    for k, _ in FeCfg._fields_:
        v = d[k]
        if k == "transform":
            v = TRANSFORMS[v]
        elif k == "cmn":
            v = CMN_TYPES[v]
        setattr(c, k, int(v) if k not in ("wlen", "alpha", "lowerf", "upperf") else float(v))
    return c


class OracleFrontend:
    """PCM -> MFCC -> features through the C restatement (whole utterances)."""

    def __init__(self, cfg):
        build()
        self.lib = L = C.CDLL(LIB)
        L.orc_fe_new.restype = C.c_void_p
        L.orc_fe_new.argtypes = [C.POINTER(FeCfg)]
        L.orc_fe_free.argtypes = [C.c_void_p]
        L.orc_fe_n_frames.restype = C.c_long
        L.orc_fe_n_frames.argtypes = [C.c_void_p, C.c_long]
        L.orc_fe_mfcc.restype = C.c_long
        L.orc_fe_mfcc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p]
        L.orc_fe_feat.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
        L.orc_fe_dims.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_fe_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        self.cfg = cfg
        h = L.orc_fe_new(C.byref(cfg))
        if not h:
            raise RuntimeError("orc_fe_new: unsupported frontend parameters")
        self.h = C.c_void_p(h)
        d = np.zeros(4, np.int32)
        L.orc_fe_dims(self.h, d.ctypes.data)
        self.frame_size, self.frame_shift, self.fft_size, self.n_coeffs = [int(x) for x in d]

    def close(self):
        if self.h:
            self.lib.orc_fe_free(self.h)
            self.h = None

    def n_frames(self, n_samples):
        return int(self.lib.orc_fe_n_frames(self.h, n_samples))

    def tables(self):
        c = self.cfg
        out = dict(spec_start=np.zeros(c.nfilt, np.int32), filt_width=np.zeros(c.nfilt, np.int32),
                   coeffs=np.zeros(self.n_coeffs, np.float32),
                   mel_cosine=np.zeros((c.ncep, c.nfilt), np.float32),
                   lifter=np.zeros(c.ncep, np.float32), hamming=np.zeros(self.frame_size // 2))
        self.lib.orc_fe_tables(self.h, *[out[k].ctypes.data for k in
                                         ("spec_start", "filt_width", "coeffs", "mel_cosine", "lifter", "hamming")])
        return out

    def mfcc(self, pcm, want_melspec=False):
        pcm = np.ascontiguousarray(pcm)
        assert pcm.dtype in (np.int16, np.float32)
        n = self.n_frames(len(pcm))
        out = np.zeros((n, self.cfg.ncep), np.float32)
        mel = np.zeros((n, self.cfg.nfilt), np.float64) if want_melspec else None
        p16 = pcm.ctypes.data if pcm.dtype == np.int16 else None
        p32 = pcm.ctypes.data if pcm.dtype == np.float32 else None
        got = self.lib.orc_fe_mfcc(self.h, p16, p32, len(pcm), out.ctypes.data,
                                   mel.ctypes.data if want_melspec else None)
        assert got == n
        return (out, mel) if want_melspec else out

    def features(self, pcm):
        """(mfcc before CMN, feat [T][3*ncep])"""
        raw = self.mfcc(pcm)
        work = raw.copy()
        feat = np.zeros((len(raw), 3 * self.cfg.ncep), np.float32)
        self.lib.orc_fe_feat(self.h, work.ctypes.data, len(raw), feat.ctypes.data)
        return raw, feat
