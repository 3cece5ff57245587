"""ctypes wrapper around oracle/_ref/libssref.so (TEST INFRASTRUCTURE ONLY).

libssref.so is the unmodified reference (ReadAlongs/SoundSwallower 0.6.1) built
by oracle/Makefile from /root/reference plus oracle/ref_shim.c.  Only tests/,
tools/make_golden.py, __graft_entry__.smoke() and bench.py's CPU-baseline legs
may import this module; the product package never does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_SSB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libssref_ssb.so")
LIB = os.path.join(HERE, "_ref", "libssref.so")


def available():
    return os.path.exists(LIB)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Ref:
    """One reference decoder_t."""

    def __init__(self, hmmdir, dictfile=None, compallsen=False, samprate=0, loglevel="FATAL", lib=None):
        # lib: another build of the same sources (oracle/_ref/libssref_ssb.so = the reference linked
        # against libssb200.so, its scorer replaced through acmod_load_am)
        self.lib = C.CDLL(lib or LIB)
        self.external_scorer = lib is not None
        L = self.lib
        L.ref_new.restype = C.c_void_p
        L.ref_new.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_char_p]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_wordstr.restype = C.c_char_p
        L.ref_wordstr.argtypes = [C.c_void_p, C.c_int]
        L.ref_wordid.argtypes = [C.c_void_p, C.c_char_p]
        self.h = L.ref_new(hmmdir.encode(), (dictfile or "").encode(), int(compallsen),
                           int(samprate), loglevel.encode())
        if not self.h:
            raise RuntimeError("reference decoder_init failed")
        self.h = C.c_void_p(self.h)
        dims = np.zeros(16, np.int32)
        if self.external_scorer:
            return  # (ref_model_dims reads ptm_mgau_t's private fields)
        if L.ref_model_dims(self.h, _p(dims, C.c_int32)) != 0:
            raise RuntimeError("not a PTM model")
        (self.n_mgau, self.n_feat, self.n_density, self.veclen, self.n_sen, self.n_sseq,
         self.n_emit, self.n_tmat, self.n_ciphone, self.n_phone, self.sil) = [int(x) for x in dims[:11]]
        self.D = self.n_feat * self.veclen

    def close(self):
        if self.h:
            self.lib.ref_free(self.h)
            self.h = None

    def model_arrays(self):
        G = self.n_mgau * self.n_feat * self.n_density
        mean = np.zeros((self.n_mgau, self.n_feat, self.n_density, self.veclen), np.float32)
        var = np.zeros_like(mean)
        det = np.zeros((self.n_mgau, self.n_feat, self.n_density), np.float32)
        mixw = np.zeros((self.n_feat, self.n_density, self.n_sen), np.uint8)
        sen2cb = np.zeros(self.n_sen, np.uint8)
        tp = np.zeros((self.n_tmat, self.n_emit, self.n_emit + 1), np.uint8)
        sseq = np.zeros((self.n_sseq, self.n_emit), np.uint16)
        lut = np.zeros(256, np.uint8)
        rv = self.lib.ref_model_copy(self.h, _p(mean, C.c_float), _p(var, C.c_float),
                                     _p(det, C.c_float), _p(mixw, C.c_uint8),
                                     _p(sen2cb, C.c_uint8), _p(tp, C.c_uint8),
                                     _p(sseq, C.c_uint16), _p(lut, C.c_uint8))
        assert rv == 0 and G > 0
        return dict(mean=mean, var=var, det=det, mixw=mixw, sen2cb=sen2cb, tp=tp, sseq=sseq, lut=lut)

    def phone_table(self):
        ssid = np.zeros(self.n_phone, np.int32)
        tmat = np.zeros(self.n_phone, np.int32)
        ci = np.zeros(self.n_phone, np.int32)
        self.lib.ref_phone_table(self.h, _p(ssid, C.c_int32), _p(tmat, C.c_int32), _p(ci, C.c_int32))
        return ssid, tmat, ci

    def features_from_pcm(self, pcm):
        pcm = np.ascontiguousarray(pcm, np.int16)
        maxf = len(pcm) // 80 + 16
        out = np.zeros((maxf, self.D), np.float32)
        n = self.lib.ref_features_from_pcm(self.h, _p(pcm, C.c_int16), C.c_long(len(pcm)),
                                           _p(out, C.c_float), maxf)
        assert n >= 0, n
        return out[:n].copy()

    def mfcc_from_pcm(self, pcm, ncep=13):
        pcm = np.ascontiguousarray(pcm, np.int16)
        maxf = len(pcm) // 80 + 16
        out = np.zeros((maxf, ncep), np.float32)
        n = self.lib.ref_mfcc_from_pcm(self.h, _p(pcm, C.c_int16), C.c_long(len(pcm)),
                                       _p(out, C.c_float), maxf)
        assert n >= 0, n
        return out[:n].copy()

    def mfcc_from_f32(self, pcm, ncep=13):
        pcm = np.ascontiguousarray(pcm, np.float32)
        maxf = len(pcm) // 80 + 16
        out = np.zeros((maxf, ncep), np.float32)
        n = self.lib.ref_mfcc_from_f32(self.h, _p(pcm, C.c_float), C.c_long(len(pcm)),
                                       _p(out, C.c_float), maxf)
        assert n >= 0, n
        return out[:n].copy()

    def reset_hist(self):
        self.lib.ref_reset_hist(self.h)

    def score_all(self, feat, reset_hist=True):
        feat = np.ascontiguousarray(feat, np.float32)
        T = feat.shape[0]
        out = np.zeros((T, self.n_sen), np.int16)
        rv = self.lib.ref_score_all(self.h, _p(feat, C.c_float), T, _p(out, C.c_int16), int(reset_hist))
        assert rv == T, rv
        return out

    def frame_eval(self, feat, frame, active=None, compallsen=True, want_topn=False):
        feat = np.ascontiguousarray(feat, np.float32).reshape(-1)
        out = np.zeros(self.n_sen, np.int16)
        if active is None:
            active = np.zeros(1, np.uint8)
            n_active = 0
        else:
            active = np.ascontiguousarray(active, np.uint8)
            n_active = len(active)
        topn = np.zeros((self.n_mgau, self.n_feat, 4, 2), np.int32) if want_topn else None
        self.lib.ref_frame_eval(self.h, _p(feat, C.c_float), int(frame), _p(active, C.c_uint8),
                                n_active, int(compallsen), _p(out, C.c_int16),
                                _p(topn, C.c_int32) if want_topn else None)
        return (out, topn) if want_topn else out

    def wordid(self, w):
        return self.lib.ref_wordid(self.h, w.encode())

    def wordstr(self, wid):
        s = self.lib.ref_wordstr(self.h, int(wid))
        return s.decode() if s else None

    @staticmethod
    def _albufs(maxw=4096, maxp=16384, maxs=49152):
        return (np.zeros(16, np.int32), np.zeros((maxw, 4), np.int32),
                np.zeros((maxp, 7), np.int32), np.zeros((maxs, 5), np.int32))

    @staticmethod
    def _alout(n_out, words, phones, states):
        return dict(words=words[:n_out[0]].copy(), phones=phones[:n_out[1]].copy(),
                    states=states[:n_out[2]].copy())

    def align_pcm(self, pcm, text):
        """2-pass alignment like the CLI.  Returns dict with segs (pass 1) and
        words[wid,start,dur,score] phones[ci,ssid,tmat,start,dur,score,parent]
        states[senid,start,dur,score,parent]."""
        pcm = np.ascontiguousarray(pcm, np.int16)
        n_out, words, phones, states = self._albufs()
        segs = np.zeros((4096, 5), np.int32)
        rv = self.lib.ref_align_pcm(self.h, _p(pcm, C.c_int16), C.c_long(len(pcm)), text.encode(),
                                    _p(n_out, C.c_int32), _p(segs, C.c_int32), 4096,
                                    _p(words, C.c_int32), _p(phones, C.c_int32), _p(states, C.c_int32),
                                    words.shape[0], phones.shape[0], states.shape[0])
        assert rv == 0, rv
        r = self._alout(n_out, words, phones, states)
        r.update(segs=segs[:n_out[3]].copy(), hyp_score=int(n_out[4]), n_frames=int(n_out[5]))
        return r

    def populate(self, wids, start=None, dur=None):
        wids = np.ascontiguousarray(wids, np.int32)
        nw = len(wids)
        start = np.zeros(nw, np.int32) if start is None else np.ascontiguousarray(start, np.int32)
        dur = np.zeros(nw, np.int32) if dur is None else np.ascontiguousarray(dur, np.int32)
        n_out, words, phones, states = self._albufs()
        rv = self.lib.ref_populate(self.h, _p(wids, C.c_int32), _p(start, C.c_int32), _p(dur, C.c_int32),
                                   nw, _p(n_out, C.c_int32), _p(words, C.c_int32), _p(phones, C.c_int32),
                                   _p(states, C.c_int32), words.shape[0], phones.shape[0], states.shape[0])
        assert rv == 0, rv
        return self._alout(n_out, words, phones, states)

    def state_align(self, feat, wids, start=None, dur=None, clear_active=True, want_tokens=False,
                    want_senscr=False):
        feat = np.ascontiguousarray(feat, np.float32)
        T = feat.shape[0]
        wids = np.ascontiguousarray(wids, np.int32)
        nw = len(wids)
        start = np.zeros(nw, np.int32) if start is None else np.ascontiguousarray(start, np.int32)
        dur = np.zeros(nw, np.int32) if dur is None else np.ascontiguousarray(dur, np.int32)
        n_out, words, phones, states = self._albufs()
        chain = self.populate(wids, start, dur)
        ns = len(chain["states"])
        tokens = np.zeros((T, ns, 2), np.int32) if want_tokens else None
        senscr = np.zeros((T, self.n_sen), np.int16) if want_senscr else None
        rv = self.lib.ref_state_align(self.h, _p(feat, C.c_float), T, _p(wids, C.c_int32),
                                      _p(start, C.c_int32), _p(dur, C.c_int32), nw, int(clear_active),
                                      _p(n_out, C.c_int32), _p(words, C.c_int32), _p(phones, C.c_int32),
                                      _p(states, C.c_int32), words.shape[0], phones.shape[0],
                                      states.shape[0],
                                      _p(tokens, C.c_int32) if want_tokens else None,
                                      _p(senscr, C.c_int16) if want_senscr else None)
        r = self._alout(n_out, words, phones, states)
        r.update(rv=rv, best_score=int(n_out[6]), tokens=tokens, senscr=senscr)
        return r

    def fsg_graph(self, align_text=None, jsgf=None, fsg_file=None):
        """Select a grammar and return the flattened FSG + lextree the search runs on
        (see ref_fsg_dump in ref_shim.c for the layouts)."""
        L = self.lib
        if fsg_file:
            rv = L.ref_fsg_prepare_file(self.h, fsg_file.encode())
        else:
            rv = L.ref_fsg_prepare(self.h, align_text.encode() if align_text else None,
                                   jsgf.encode() if jsgf else None)
        assert rv == 0, rv
        d = np.zeros(16, np.int32)
        assert L.ref_fsg_dims(self.h, _p(d, C.c_int32)) == 0
        (n_state, start, final, n_link, n_pnode, n_ci, sil, beam, pbeam, wbeam, maxhmmpf, wip,
         pip) = [int(x) for x in d[:13]]
        link4 = np.zeros((n_link, 4), np.int32)
        flag = np.zeros(n_link, np.uint8)
        arc_off = np.zeros(n_state + 1, np.int32)
        root = np.zeros(n_state, np.int32)
        pnode8 = np.zeros((n_pnode, 8), np.int32)
        ctxt = np.zeros((n_pnode, 4), np.uint32)
        rv = L.ref_fsg_dump(self.h, _p(link4, C.c_int32), _p(flag, C.c_uint8), _p(arc_off, C.c_int32),
                            _p(root, C.c_int32), _p(pnode8, C.c_int32), _p(ctxt, C.c_uint32))
        assert rv == 0, rv
        return dict(n_state=n_state, start=start, final=final, n_ciphone=n_ci, sil=sil, beam=beam,
                    pbeam=pbeam, wbeam=wbeam, maxhmmpf=maxhmmpf, wip=wip, pip=pip, link=link4,
                    link_flag=flag, arc_off=arc_off, root=root, pnode=pnode8, ctxt=ctxt)

    def fsg_history(self, max_ent=1 << 16):
        """History table of the last fsg_decode: [n][9] = link score pred frame lc rc[4]."""
        ent = np.zeros((max_ent, 9), np.int32)
        n_out = np.zeros(8, np.int32)
        n = self.lib.ref_fsg_history(self.h, _p(ent, C.c_int32), max_ent, _p(n_out, C.c_int32))
        assert n >= 0, n
        return dict(hist=ent[:n].copy(), n_hmm_eval=int(n_out[1]), n_frames=int(n_out[2]),
                    hyp_score=int(n_out[3]))

    def fsg_decode(self, feat, align_text=None, jsgf=None):
        feat = np.ascontiguousarray(feat, np.float32)
        n_out = np.zeros(16, np.int32)
        segs = np.zeros((4096, 5), np.int32)
        rv = self.lib.ref_fsg_decode(self.h, _p(feat, C.c_float), feat.shape[0],
                                     align_text.encode() if align_text else None,
                                     jsgf.encode() if jsgf else None,
                                     _p(n_out, C.c_int32), _p(segs, C.c_int32), 4096)
        assert rv >= 0, rv
        return dict(segs=segs[:rv].copy(), hyp_score=int(n_out[4]), n_frames=int(n_out[5]))

    def fsg_partial(self, feat, align_text, stops, maxseg=64):
        """decoder_hyp / decoder_seg_iter between steps: after each of `stops` frames.  Returns a
        list of dicts hyp (str or None), hyp_score, segs [n][5] (wid sf ef ascr lscr)."""
        feat = np.ascontiguousarray(feat, np.float32)
        stops = np.ascontiguousarray(stops, np.int32)
        out = np.zeros((len(stops), 2 + 5 * maxseg), np.int32)
        hyps = C.create_string_buffer(len(stops) * 512)
        self.lib.ref_fsg_partial.restype = C.c_int
        self.lib.ref_fsg_partial.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_char_p,
                                             C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32), C.c_int,
                                             C.c_char_p, C.c_int]
        rv = self.lib.ref_fsg_partial(self.h, _p(feat, C.c_float), feat.shape[0], align_text.encode(),
                                      _p(stops, C.c_int32), len(stops), _p(out, C.c_int32), maxseg, hyps, 512)
        assert rv == len(stops), rv
        res = []
        for k in range(len(stops)):
            sc, n = int(out[k, 0]), int(out[k, 1])
            txt = hyps.raw[k * 512:(k + 1) * 512].split(b"\0")[0].decode()
            res.append(dict(hyp=None if sc == -2**31 else txt, hyp_score=None if sc == -2**31 else sc,
                            segs=out[k, 2:2 + 5 * n].reshape(n, 5).copy()))
        return res

    def hmm_vit_eval_tp(self, tp, senscr, st):
        """hmm_vit_eval on a caller-provided transition matrix [n_emit][n_emit+1] and the n_emit
        scores of the HMM's own states (3- or 5-state)."""
        tp = np.ascontiguousarray(tp, np.uint8)
        n_emit = tp.shape[0]
        senscr = np.ascontiguousarray(senscr, np.int16)
        st = np.ascontiguousarray(st, np.int32).copy()
        best = self.lib.ref_hmm_vit_eval_tp(self.h, n_emit, _p(tp, C.c_uint8), _p(senscr, C.c_int16),
                                            _p(st, C.c_int32))
        return best, st

    def active_bits(self):
        """(acmod's active-senone flags as uint32 words, senones evaluated by the last fsg search)"""
        out = np.zeros((self.n_sen + 31) // 32, np.uint32)
        n = C.c_int32(0)
        assert self.lib.ref_active_bits(self.h, _p(out, C.c_uint32), C.byref(n)) == 0
        return out, int(n.value)

    def hmm_vit_eval(self, n_emit, tmatid, senid, senscr, st):
        senid = np.ascontiguousarray(senid, np.uint16)
        senscr = np.ascontiguousarray(senscr, np.int16)
        st = np.ascontiguousarray(st, np.int32).copy()
        best = self.lib.ref_hmm_vit_eval(self.h, n_emit, tmatid, _p(senid, C.c_uint16),
                                         _p(senscr, C.c_int16), _p(st, C.c_int32))
        return best, st
