"""soundswallower_b200 -- B200 (sm_100a) acoustic scoring + Viterbi alignment.

Python mirror of the reference interfaces for this one path, over the C ABI of
libssb200.so (include/ssb200.h):

  AcousticModel    the model tables acmod_load_am builds (ref: src/acmod.c:62-129)
  PtmMgau          mgau_t / mgaufuncs_t, one frame per call (ref: include/soundswallower/acmod.h:93-119)
  StateAlignBatch  state_align_search over a batch of utterances
                   (ref: src/state_align_search.c:177-268, 429-474)
  score_batch / topn_batch / hmm_vit_eval   the pieces, for tests and benchmarks

numpy arrays in, numpy arrays out; all compute happens in the CUDA library.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import Config, SsbError

INT_MAX = 2**31 - 1
WORST_SCORE = -536870912  # ref: include/soundswallower/hmm.h:80
MODEL_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "model")

__all__ = ["AcousticModel", "PtmMgau", "StateAlignBatch", "align_batch", "score_batch",
           "topn_batch", "tc_probe", "fsg_batch", "hmm_vit_eval", "windows", "plan_chain", "propagate", "flags2list", "device_count",
           "align_batch_multi", "Frontend", "DeviceFeatures", "Lexicon", "read_fsg_file", "align_texts", "SsbError", "Config", "MODEL_DIR", "INT_MAX", "WORST_SCORE"]


def _ptr(a, t=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(t)) if t is not None else a.ctypes.data_as(C.c_void_p)


def device_count():
    return _lib.load().ssb_device_count()


def model_path(name):
    """Bundled model directory ("en-us", "fr-fr")."""
    return os.path.join(MODEL_DIR, name)


class AcousticModel:
    """Parsed model directory + its packed image in HBM.

    device=-1 parses only (loader tests on machines without a GPU); every compute
    call on such a model fails -- there is no CPU path."""

    def __init__(self, hmmdir, device=0, topn=4, ds=1, logbase=1.0001, varfloor=1e-4,
                 mixwfloor=1e-7, tmatfloor=1e-4, topn_beam=None):
        self.lib = L = _lib.load()
        cfg = Config()
        L.ssb_config_defaults(C.byref(cfg))
        cfg.logbase, cfg.varfloor, cfg.mixwfloor, cfg.tmatfloor = logbase, varfloor, mixwfloor, tmatfloor
        cfg.topn, cfg.ds, cfg.device = topn, ds, device
        if topn_beam is not None:  # "topn_beam" of semi-continuous models; missing entries
            beam = [int(x) for x in (topn_beam.split(",") if isinstance(topn_beam, str) else topn_beam)]
            for f in range(4):     # repeat the largest one (ref: src/s2_semi_mgau.c:877-907)
                cfg.topn_beam[f] = beam[f] if f < len(beam) else max(beam)
        h = L.ssb_model_load(os.fsencode(hmmdir), C.byref(cfg))
        if not h:
            raise SsbError("ssb_model_load(%s): %s" % (hmmdir, _lib.last_error()))
        self.h = C.c_void_p(h)
        self.device = device
        self.topn = topn
        d = np.zeros(16, np.int32)
        L.ssb_model_dims(self.h, _ptr(d, C.c_int32))
        (self.n_mgau, self.n_feat, self.n_density, self.veclen, self.n_sen, self.n_sseq,
         self.n_emit, self.n_tmat, self.n_ciphone, self.n_phone, self.sil) = [int(x) for x in d[:11]]
        self.featlen = [int(x) for x in d[11:11 + self.n_feat]]
        self.blk = int(d[15])
        self.kind = int(L.ssb_model_kind(self.h))   # 0 PTM, 1 semi-continuous
        self._arrays = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.ssb_model_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def fsg_active_ok(self):
        """The grammar search can run in the reference's default (active-list) mode."""
        return bool(self.lib.ssb_model_fsg_active_ok(self.h))

    def arrays(self):
        """Host copies of the parsed tables, in the reference's in-memory shapes."""
        if self._arrays is None:
            L = self.n_mgau, self.n_feat, self.n_density
            if len(set(self.featlen)) == 1:
                mean = np.zeros(L + (self.veclen,), np.float32)
            else:
                mean = np.zeros(self.n_mgau * self.n_density * self.blk, np.float32)
            var = np.zeros_like(mean)
            det = np.zeros(L, np.float32)
            mixw = np.zeros((self.n_feat, self.n_density, self.n_sen), np.uint8)
            sen2cb = np.zeros(self.n_sen, np.uint8)
            tp = np.zeros((self.n_tmat, self.n_emit, self.n_emit + 1), np.uint8)
            sseq = np.zeros((self.n_sseq, self.n_emit), np.uint16)
            lut = np.zeros(256, np.uint8)
            self.lib.ssb_model_copy(self.h, _ptr(mean), _ptr(var), _ptr(det), _ptr(mixw),
                                    _ptr(sen2cb), _ptr(tp), _ptr(sseq), _ptr(lut))
            self._arrays = dict(mean=mean, var=var, det=det, mixw=mixw, sen2cb=sen2cb, tp=tp,
                                sseq=sseq, lut=lut)
        return self._arrays

    def ciname(self, ci):
        s = self.lib.ssb_model_ciphone_str(self.h, int(ci))
        return s.decode("utf-8") if s is not None else None

    def phone_table(self):
        ssid = np.zeros(self.n_phone, np.int32)
        tmat = np.zeros(self.n_phone, np.int32)
        ci = np.zeros(self.n_phone, np.int32)
        self.lib.ssb_model_phones(self.h, _ptr(ssid), _ptr(tmat), _ptr(ci))
        return ssid, tmat, ci


def flags2list(senones, n_sen=None):
    """acmod_flags2list (ref: src/acmod.c:947-999): sorted senone ids -> uint8 delta list,
    gaps above 255 bridged with extra entries."""
    out, last = [], 0
    for s in sorted(set(int(x) for x in senones)):
        delta = s - last
        while delta > 255:
            out.append(255)
            delta -= 255
        out.append(delta)
        last = s
    return np.asarray(out, np.uint8)


class PtmMgau:
    """The scorer object behind acmod_score: same layout ({vt, frame_idx} first) and the same
    frame_eval contract as mgau_t / mgaufuncs_t (ref: acmod.h:93-119, src/ptm_mgau.c:408-454).
    Calls go through the object's own vtable pointer, like ps_mgau_frame_eval() does."""

    def __init__(self, model):
        self.model = model
        self.lib = model.lib
        self.p = self.lib.ssb_mgau_init(model.h)
        if not self.p:
            raise SsbError("ssb_mgau_init: " + _lib.last_error())

    @property
    def name(self):
        return self.p.contents.vt.contents.name.decode()

    @property
    def frame_idx(self):
        return self.p.contents.frame_idx

    @frame_idx.setter
    def frame_idx(self, v):
        self.p.contents.frame_idx = int(v)  # what acmod_advance / acmod_rewind poke

    def reset(self):
        self.lib.ssb_mgau_reset(self.p)

    def frame_eval(self, feat, frame, senone_active=None, compallsen=False):
        """feat: [blk] floats of one frame (streams concatenated).  senone_active: uint8
        delta list as produced by flags2list.  Returns int16[n_sen]."""
        m = self.model
        feat = np.ascontiguousarray(feat, np.float32).reshape(-1)
        assert feat.size == m.blk
        ptrs = (C.POINTER(C.c_float) * m.n_feat)()
        off = 0
        for f in range(m.n_feat):
            ptrs[f] = C.cast(feat.ctypes.data + 4 * off, C.POINTER(C.c_float))
            off += m.featlen[f]
        if senone_active is None:
            act = np.zeros(1, np.uint8)
            n_act = 0
        else:
            act = np.ascontiguousarray(senone_active, np.uint8)
            n_act = len(act)
        out = np.zeros(m.n_sen, np.int16)
        fn = self.p.contents.vt.contents.frame_eval
        rv = fn(C.cast(self.p, C.c_void_p), _ptr(out, C.c_int16), _ptr(act, C.c_uint8), n_act, ptrs,
                int(frame), int(bool(compallsen)))
        _lib.check(rv, "mgau frame_eval")
        return out

    def close(self):
        if getattr(self, "p", None):
            self.lib.ssb_mgau_free(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def windows(start, dur):
    """state_align_search_init's window rule (ref: src/state_align_search.c:464-471)."""
    start = np.asarray(start, np.int32)
    dur = np.asarray(dur, np.int32)
    sf = np.where(start > 0, start, 0).astype(np.int32)
    ef = np.where(dur > 0, start.astype(np.int64) + dur, INT_MAX).astype(np.int32)
    return sf, ef


def plan_chain(n_frames, sf, ef):
    """Host-only: the frame each phone of a chain is first evaluated on (-1 never)."""
    sf = np.ascontiguousarray(sf, np.int32)
    ef = np.ascontiguousarray(ef, np.int32)
    enter = np.full(len(sf), -1, np.int32)
    rv = _lib.load().ssb_plan_chain(len(sf), int(n_frames), _ptr(sf), _ptr(ef), _ptr(enter))
    _lib.check(rv, "ssb_plan_chain")
    return enter


def propagate(start, dur, score, n_emit):
    """alignment_propagate, states -> phones (ref: src/ps_alignment.c:317-341)."""
    start = np.asarray(start).reshape(-1, n_emit)
    dur = np.asarray(dur).reshape(-1, n_emit)
    score = np.asarray(score).reshape(-1, n_emit)
    return start[:, 0].copy(), dur.sum(1).astype(np.int32), score.sum(1).astype(np.int32)


def _concat_i32(seq):
    seq = [np.ascontiguousarray(a, np.int32).reshape(-1) for a in seq]
    return np.concatenate(seq) if seq else np.zeros(0, np.int32)


class _DevPtr:
    """A raw (host or device) address with whatever keeps it alive."""

    def __init__(self, addr, keep):
        self.addr, self.keep = addr, keep

    def cast(self):
        return C.cast(C.c_void_p(self.addr), C.POINTER(C.c_float))


class StateAlignBatch:
    """A batch of state_align_search problems resident on one GPU.

    upload(feats, chains) -> run() -> download().  `chains` is a list of dicts with the
    per-phone arrays state_align_search_init derives: ssid, tmat, sf, ef."""

    def __init__(self, model, stream=None):
        self.model = model
        self.lib = model.lib
        self.b = self.lib.ssb_batch_create(model.h, C.c_void_p(stream or 0))
        if not self.b:
            raise SsbError("ssb_batch_create: " + _lib.last_error())
        self.b = C.c_void_p(self.b)
        self._keep = None

    def close(self):
        if getattr(self, "b", None):
            self.lib.ssb_batch_free(self.b)
            self.b = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _align_in(self, feat, frame_off, phone_off, ssid, tmat, sf, ef, init_active, compallsen,
                  init_topn=None):
        n_utts = len(frame_off) - 1
        a = _lib.AlignIn()
        a.n_utts = n_utts
        a.feat = feat.cast() if isinstance(feat, _DevPtr) else _ptr(feat, C.c_float)
        a.frame_off = _ptr(frame_off, C.c_int64)
        a.phone_off = _ptr(phone_off, C.c_int64)
        a.ssid = _ptr(ssid, C.c_int32)
        a.tmat = _ptr(tmat, C.c_int32)
        a.sf = _ptr(sf, C.c_int32)
        a.ef = _ptr(ef, C.c_int32)
        a.init_active = _ptr(init_active, C.c_uint32) if init_active is not None else None
        a.compallsen = int(bool(compallsen))
        if init_topn is not None:
            init_topn = np.ascontiguousarray(init_topn, np.uint8)
            assert init_topn.size == n_utts * self.model.n_mgau * self.model.n_feat * 4
        a.init_topn = _ptr(init_topn, C.c_uint8) if init_topn is not None else None
        self._keep = (feat, frame_off, phone_off, ssid, tmat, sf, ef, init_active, init_topn)
        self.n_utts = n_utts
        self.frame_off = np.asarray(frame_off)
        self.phone_off = np.asarray(phone_off)
        return a

    def upload_raw(self, feat, frame_off, phone_off, ssid, tmat, sf, ef, init_active=None,
                   compallsen=False, init_topn=None):
        """Flat arrays exactly as ssb_align_in_t takes them (feat may be pinned memory)."""
        a = self._align_in(feat, frame_off, phone_off, ssid, tmat, sf, ef, init_active, compallsen,
                           init_topn)
        _lib.check(self.lib.ssb_batch_upload(self.b, C.byref(a)), "ssb_batch_upload")

    def upload(self, feats, chains, init_active=None, compallsen=False, init_topn=None):
        m = self.model
        assert len(feats) == len(chains)
        fptr, frame_off, keep = _flat_feats(m, feats)
        feat = _DevPtr(fptr, keep)
        phone_off = np.zeros(len(chains) + 1, np.int64)
        for i, c in enumerate(chains):
            phone_off[i + 1] = phone_off[i] + len(c["ssid"])
        ia = None
        if init_active is not None:
            nw = (m.n_sen + 31) // 32
            ia = np.zeros((len(chains), nw), np.uint32)
            for u, sens in enumerate(init_active):
                for s in (sens or ()):
                    ia[u, s >> 5] |= np.uint32(1 << (s & 31))
        self.upload_raw(feat, frame_off, phone_off, _concat_i32([c["ssid"] for c in chains]),
                        _concat_i32([c["tmat"] for c in chains]),
                        _concat_i32([c["sf"] for c in chains]),
                        _concat_i32([c["ef"] for c in chains]), ia, compallsen,
                        None if init_topn is None else np.stack([np.asarray(t, np.uint8).reshape(-1, 4)
                                                                 for t in init_topn]))

    def debug_tokens(self, on=True):
        self.lib.ssb_batch_debug_tokens(self.b, int(on))
        self._tokens = bool(on)

    def run(self):
        _lib.check(self.lib.ssb_batch_run(self.b), "ssb_batch_run")

    def _align_out(self, want_chain_scr, want_tokens, init):
        E = self.model.n_emit
        ns = int(self.phone_off[-1]) * E
        U = self.n_utts
        if init is None:
            st = [np.zeros(ns, np.int32) for _ in range(3)]
        else:
            st = [np.ascontiguousarray(a, np.int32).copy() for a in init]
        rv, best, ren = (np.zeros(U, np.int32) for _ in range(3))
        T = np.diff(self.frame_off)
        npz = np.diff(self.phone_off)
        nsf = int((T * npz * E).sum())
        cs = np.zeros(nsf, np.int16) if want_chain_scr else None
        tk = np.zeros((nsf, 2), np.int32) if want_tokens else None
        o = _lib.AlignOut()
        o.st_start, o.st_dur, o.st_score = (_ptr(a, C.c_int32) for a in st)
        o.utt_rv, o.utt_best, o.utt_renorm = (_ptr(a, C.c_int32) for a in (rv, best, ren))
        o.chain_scr = _ptr(cs, C.c_int16) if cs is not None else None
        o.tokens = _ptr(tk, C.c_int32) if tk is not None else None
        return o, dict(start=st[0], dur=st[1], score=st[2], rv=rv, best_score=best, n_renorm=ren,
                       chain_scr=cs, tokens=tk)

    def download(self, want_chain_scr=False, want_tokens=False, init=None):
        """Returns flat arrays; `per_utt()` splits them.  `init` = (start, dur, score) the
        caller's pre-filled state entries (states off the best path keep them)."""
        o, res = self._align_out(want_chain_scr, want_tokens, init)
        _lib.check(self.lib.ssb_batch_download(self.b, C.byref(o)), "ssb_batch_download")
        return res

    def per_utt(self, res):
        E = self.model.n_emit
        out = []
        sf_off = 0
        for u in range(self.n_utts):
            p0, p1 = int(self.phone_off[u]) * E, int(self.phone_off[u + 1]) * E
            T = int(self.frame_off[u + 1] - self.frame_off[u])
            n = T * (p1 - p0)
            d = dict(start=res["start"][p0:p1], dur=res["dur"][p0:p1], score=res["score"][p0:p1],
                     rv=int(res["rv"][u]), best_score=int(res["best_score"][u]),
                     n_renorm=int(res["n_renorm"][u]))
            if res.get("chain_scr") is not None:
                d["chain_scr"] = res["chain_scr"][sf_off:sf_off + n].reshape(T, p1 - p0)
            if res.get("tokens") is not None:
                d["tokens"] = res["tokens"][sf_off:sf_off + n].reshape(T, p1 - p0, 2)
            sf_off += n
            out.append(d)
        return out

    def kernel_ms(self):
        ms = np.zeros(8, np.float32)
        _lib.check(self.lib.ssb_batch_kernel_ms(self.b, _ptr(ms, C.c_float)), "ssb_batch_kernel_ms")
        return dict(gmm_topn=float(ms[0]), senone_mix=float(ms[1]), chain_viterbi=float(ms[2]),
                    backtrace=float(ms[3]), total=float(ms[4]))

    def n_launches(self):
        return int(self.lib.ssb_batch_n_launches(self.b))

    def stats(self):
        s = np.zeros(8, np.int64)
        self.lib.ssb_batch_stats(self.b, _ptr(s, C.c_int64))
        return dict(frames=int(s[0]), state_frames=int(s[1]), active_senone_frames=int(s[2]),
                    scanned_cb_frames=int(s[3]), device_bytes=int(s[4]), max_union=int(s[5]),
                    max_phones=int(s[6]), plan_us=int(s[7]),
                    band_state_frames=int(self.lib.ssb_batch_band_state_frames(self.b)),
                    segments=int(self.lib.ssb_batch_n_segments(self.b)))


class _AlignCall(StateAlignBatch):
    """Argument marshalling only: upload()/upload_raw() keep the ssb_align_in_t for a later call."""

    def __init__(self, model):
        self.model = model
        self.lib = model.lib
        self.b = None
        self._keep = None
        self._in = None

    def upload_raw(self, feat, frame_off, phone_off, ssid, tmat, sf, ef, init_active=None,
                   compallsen=False, init_topn=None):
        self._in = self._align_in(feat, frame_off, phone_off, ssid, tmat, sf, ef, init_active,
                                  compallsen, init_topn)

    def run(self):
        raise SsbError("not a resident batch: use align()")

    def align(self, want_chain_scr=False, want_tokens=False, init=None):
        """ssb_align_batch: upload + run + download in one C call."""
        o, res = self._align_out(want_chain_scr, want_tokens, init)
        _lib.check(self.lib.ssb_align_batch(self.model.h, C.byref(self._in), C.byref(o)),
                   "ssb_align_batch")
        return res


class AlignPipeline(_AlignCall):
    """ssb_pipeline_*: the batch cut into chunks of whole utterances that travel through a few
    lanes (stream + host thread each), so that planning and copies overlap the kernels.  Same
    upload()/upload_raw() arguments as StateAlignBatch; align() = upload + run + download."""

    def __init__(self, model, n_lanes=0, chunk_frames=0, overlap_kernels=True):
        _AlignCall.__init__(self, model)
        self.p = self.lib.ssb_pipeline_create(model.h, int(n_lanes), int(chunk_frames))
        if not self.p:
            raise SsbError("ssb_pipeline_create: " + _lib.last_error())
        self.p = C.c_void_p(self.p)
        self.lib.ssb_pipeline_set_overlap(self.p, int(bool(overlap_kernels)))
        self._inflight = {}

    def submit(self, want_chain_scr=False, want_tokens=False, init=None):
        """ssb_pipeline_submit of the uploaded arguments: returns a ticket at once; the arrays
        are kept alive here until collect(ticket) returns the result dict."""
        o, res = self._align_out(want_chain_scr, want_tokens, init)
        t = int(self.lib.ssb_pipeline_submit(self.p, C.byref(self._in), C.byref(o)))
        _lib.check(t, "ssb_pipeline_submit")
        self._inflight[t] = (self._in, self._keep, o, res)
        return t

    def collect(self, ticket):
        _in, _keep, _o, res = self._inflight.pop(ticket)
        _lib.check(self.lib.ssb_pipeline_collect(self.p, int(ticket)), "ssb_pipeline_collect")
        return res

    def close(self):
        if getattr(self, "p", None):
            self.lib.ssb_pipeline_free(self.p)
            self.p = None

    def align(self, want_chain_scr=False, want_tokens=False, init=None):
        """Runs the uploaded arguments through the pipeline; returns download()'s dict."""
        o, res = self._align_out(want_chain_scr, want_tokens, init)
        _lib.check(self.lib.ssb_pipeline_align(self.p, C.byref(self._in), C.byref(o)),
                   "ssb_pipeline_align")
        return res

    def n_launches(self):
        return int(self.lib.ssb_pipeline_n_launches(self.p))

    def n_chunks(self):
        return int(self.lib.ssb_pipeline_n_chunks(self.p))

    def trace(self):
        """[chunk][lane, first utt, upload start, upload end, download end, top-N ms, kernels ms, 0]"""
        t = np.zeros((max(self.n_chunks(), 1), 8), np.float64)
        n = self.lib.ssb_pipeline_trace(self.p, _ptr(t, C.c_double), t.shape[0])
        return t[:max(int(n), 0)]


def align_batch(model, feats, chains, init_active=None, compallsen=False, want_chain_scr=False,
                want_tokens=False, init_topn=None):
    """One-shot: upload + run + download; returns a list of per-utterance dicts
    (start/dur/score per state, rv, best_score, n_renorm[, chain_scr, tokens]).
    init_active / init_topn: per utterance what a first pass left in the shared acmod (the
    active-senone flags and the scorer's carried top-N codewords: fsg_batch's `active`, `carried`)."""
    b = _AlignCall(model)
    b.upload(feats, chains, init_active=init_active, compallsen=compallsen, init_topn=init_topn)
    return b.per_utt(b.align(want_chain_scr=want_chain_scr, want_tokens=want_tokens))


def align_batch_multi(models, feats, chains, init_active=None, compallsen=False, init_topn=None):
    """ssb_align_batch_multi: the batch over several GPUs of one box, models[d] loaded on device d
    (one host thread per GPU inside the C call, contiguous utterance ranges, no collective)."""
    b = _AlignCall(models[0])
    b.upload(feats, chains, init_active=init_active, compallsen=compallsen, init_topn=init_topn)
    o, res = b._align_out(False, False, None)
    hs = (C.c_void_p * len(models))(*[m.h for m in models])
    _lib.check(models[0].lib.ssb_align_batch_multi(hs, len(models), C.byref(b._in), C.byref(o)),
               "ssb_align_batch_multi")
    return b.per_utt(res)


def score_batch(model, feats, want=True):
    """acmod_score over whole utterances with compallsen semantics: list of int16 [T][n_sen]."""
    fptr, off, _keep_feat = _flat_feats(model, feats)
    out = np.zeros((int(off[-1]), model.n_sen), np.int16) if want else None
    n = model.lib.ssb_score_batch(model.h, C.c_void_p(fptr), _ptr(off), len(off) - 1, _ptr(out))
    _lib.check(int(n), "ssb_score_batch")
    if not want:
        return int(n)
    return [out[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def topn_batch(model, feats):
    """Raw top-N (codeword ids, int32 scores) of every frame/codebook/stream."""
    feats = [np.ascontiguousarray(f, np.float32).reshape(-1, model.blk) for f in feats]
    off = np.zeros(len(feats) + 1, np.int64)
    for i, f in enumerate(feats):
        off[i + 1] = off[i] + f.shape[0]
    feat = np.concatenate(feats) if feats else np.zeros((0, model.blk), np.float32)
    G = int(off[-1])
    cw = np.zeros((G, model.n_mgau, model.n_feat, model.topn), np.uint8)
    sc = np.zeros((G, model.n_mgau, model.n_feat, model.topn), np.int32)
    n = model.lib.ssb_topn_batch(model.h, _ptr(feat), _ptr(off), len(feats), _ptr(cw), _ptr(sc))
    _lib.check(int(n), "ssb_topn_batch")
    return ([cw[off[i]:off[i + 1]] for i in range(len(feats))],
            [sc[off[i]:off[i + 1]] for i in range(len(feats))])


def tc_probe(model, feats):
    """Tensor-core scorer verification: returns (cw, score, approx, eps, counters) with
    approx [frames][mgau][feat][n_density] the TF32 screening scores and eps (same shape)
    their guaranteed error bound (regular or "hot" class bound, per density)."""
    feats = [np.ascontiguousarray(f, np.float32).reshape(-1, model.blk) for f in feats]
    off = np.zeros(len(feats) + 1, np.int64)
    for i, f in enumerate(feats):
        off[i + 1] = off[i] + f.shape[0]
    feat = np.concatenate(feats) if feats else np.zeros((0, model.blk), np.float32)
    G = int(off[-1])
    cw = np.zeros((G, model.n_mgau, model.n_feat, model.topn), np.uint8)
    sc = np.zeros((G, model.n_mgau, model.n_feat, model.topn), np.int32)
    approx = np.zeros((G, model.n_mgau, model.n_feat, model.n_density), np.float32)
    eps2 = np.zeros((G, model.n_mgau, model.n_feat, 2), np.float32)
    cnt = np.zeros(4, np.int64)
    n = model.lib.ssb_tc_probe(model.h, _ptr(feat), _ptr(off), len(feats), _ptr(cw), _ptr(sc),
                               _ptr(approx), _ptr(eps2), _ptr(cnt))
    _lib.check(int(n), "ssb_tc_probe")
    # per-density bound: the regular one, or the hot one for densities flagged hot
    words = np.zeros((model.n_mgau * model.n_feat, 4), np.uint32)
    _lib.check(model.lib.ssb_tc_hot_mask(model.h, _ptr(words)), "ssb_tc_hot_mask")
    hot = ((words[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool)
    hot = hot.reshape(model.n_mgau, model.n_feat, 128)[:, :, :model.n_density]
    eps = np.where(hot[None], eps2[..., 1:2], eps2[..., 0:1])
    return cw, sc, approx, eps, dict(hot=hot, eps_regular=eps2[..., 0], exact_evals=int(cnt[0]), scan_steps=int(cnt[1]), slow_steps=int(cnt[2]))


def fsg_batch(model, feats, graphs, utt_graph=None, hist_cap=None, max_seg=256, want_hist=False,
              compallsen=True, partial=False):
    """fsg_search over a batch: see _fsg_batch_once.  The reference's history table and segment
    iterator are unbounded (ref: src/fsg_history.c:129-232); here they have capacities, so the
    call sizes them from the longest utterance (as the C entry points do) and repeats with
    larger ones when an utterance overflows (rv == -2 / n_seg < 0) instead of reporting a
    wrong "no hypothesis"."""
    n_frames = getattr(feats, "frame_off", None)
    if n_frames is not None:
        max_T = int(np.max(np.diff(np.asarray(n_frames)))) if len(n_frames) > 1 else 0
    else:
        max_T = max([int(np.asarray(f).reshape(-1, model.blk).shape[0]) for f in feats] or [0])
    cap = int(hist_cap) if hist_cap else max(4096, 8 * max_T)
    for _attempt in range(6):
        out = _fsg_batch_once(model, feats, graphs, utt_graph, cap, max_seg, want_hist, compallsen,
                              partial)
        over_hist = any(r["rv"] == -2 for r in out)
        over_seg = [-r["n_seg"] for r in out if r["n_seg"] < 0]
        if not over_hist and not over_seg:
            return out
        if over_hist:
            cap *= 2
        if over_seg:
            max_seg = max(max(over_seg), 2 * max_seg)
    raise SsbError("fsg_batch: history / segment capacity exceeded after 6 attempts "
                   "(hist_cap %d, max_seg %d)" % (cap, max_seg))


def _fsg_batch_once(model, feats, graphs, utt_graph=None, hist_cap=4096, max_seg=256, want_hist=False,
                    compallsen=True, partial=False):
    """fsg_search over a batch (first pass / grammar decoding) on dense senone scores.

    graphs: list of dicts with the flattened FSG + lextree (keys n_state start final n_ciphone
    sil beam pbeam wbeam maxhmmpf link link_flag arc_off root pnode ctxt -- the reference's
    fsg_model_t / fsg_lextree_t, see ssb_fsg_graph_t).  utt_graph[u] = graph of utterance u
    (default: all use graph 0).  Returns a list of per-utterance dicts: segs [n][5] (link sf ef
    ascr lscr), hyp_score, exit, rv, n_hist, n_hmm_eval[, hist [n_hist][9]] and, on the first
    one, kernel_ms / n_launches of the call.

    compallsen=False is the reference's default mode: only the senones of the active HMMs are
    scored, frame by frame inside the search; each result then also carries `active` (acmod's
    active-senone flags after the last frame, uint32 words: the second pass' init_active) and
    n_sen_eval."""
    fptr, off, _keep_feat = _flat_feats(model, feats)
    U = len(off) - 1
    ug = np.zeros(U, np.int32) if utt_graph is None else np.ascontiguousarray(utt_graph, np.int32)
    keep = []
    garr = (_lib.FsgGraph * max(len(graphs), 1))()
    for i, G in enumerate(graphs):
        a = dict(link=np.ascontiguousarray(G["link"], np.int32).reshape(-1, 4),
                 link_flag=np.ascontiguousarray(G["link_flag"], np.uint8),
                 arc_off=np.ascontiguousarray(G["arc_off"], np.int32),
                 root=np.ascontiguousarray(G["root"], np.int32),
                 pnode=np.ascontiguousarray(G["pnode"], np.int32).reshape(-1, 8),
                 ctxt=np.ascontiguousarray(G["ctxt"], np.uint32).reshape(-1, 4))
        keep.append(a)
        g = garr[i]
        g.n_state, g.start, g.final = int(G["n_state"]), int(G["start"]), int(G["final"])
        g.n_link, g.n_pnode = len(a["link"]), len(a["pnode"])
        g.n_ciphone, g.sil = int(G["n_ciphone"]), int(G["sil"])
        g.beam, g.pbeam, g.wbeam, g.maxhmmpf = (int(G[k]) for k in ("beam", "pbeam", "wbeam", "maxhmmpf"))
        g.link4, g.link_flag, g.arc_off = (a[k].ctypes.data for k in ("link", "link_flag", "arc_off"))
        g.root, g.pnode8, g.ctxt = (a[k].ctypes.data for k in ("root", "pnode", "ctxt"))
    fin = _lib.FsgIn()
    fin.n_utts, fin.feat, fin.frame_off = U, fptr, off.ctypes.data
    fin.n_graphs, fin.graphs, fin.utt_graph = len(graphs), garr, ug.ctypes.data
    fin.hist_cap, fin.max_seg = int(hist_cap), int(max_seg)
    fin.active_lists = 0 if compallsen else 1
    fin.partial = 1 if partial else 0   # hypothesis of a running utterance (find_exit, final = FALSE)
    segs = np.zeros((U, max_seg, 5), np.int32)
    n_seg, score, exit_bp, rv, n_hist = (np.zeros(U, np.int32) for _ in range(5))
    n_eval = np.zeros(U, np.int64)
    hist = np.zeros((U, hist_cap, 9), np.int32) if want_hist else None
    ms = np.zeros(4, np.float32)
    fo = _lib.FsgOut()
    fo.segs, fo.n_seg, fo.hyp_score, fo.exit_bp = (a.ctypes.data for a in (segs, n_seg, score, exit_bp))
    fo.utt_rv, fo.n_hist, fo.n_hmm_eval = rv.ctypes.data, n_hist.ctypes.data, n_eval.ctypes.data
    fo.hist9 = hist.ctypes.data if hist is not None else None
    fo.kernel_ms = ms.ctypes.data
    nw = (model.n_sen + 31) // 32
    fact = np.zeros((U, nw), np.uint32) if not compallsen else None
    nsen = np.zeros(U, np.int64) if not compallsen else None
    fo.final_active = fact.ctypes.data if fact is not None else None
    fo.n_sen_eval = nsen.ctypes.data if nsen is not None else None
    CS = model.n_mgau * model.n_feat
    ftopn = np.zeros((U, CS, 4), np.uint8) if model.kind != 2 else None
    fo.final_topn = ftopn.ctypes.data if ftopn is not None else None
    _lib.check(model.lib.ssb_fsg_batch(model.h, C.byref(fin), C.byref(fo)), "ssb_fsg_batch")
    out = []
    for u in range(U):
        d = dict(segs=segs[u, :max(int(n_seg[u]), 0)].copy(), n_seg=int(n_seg[u]), hyp_score=int(score[u]),
                 exit=int(exit_bp[u]), rv=int(rv[u]), n_hist=int(n_hist[u]), n_hmm_eval=int(n_eval[u]))
        if hist is not None:
            d["hist"] = hist[u, :int(n_hist[u])].copy()
        if fact is not None:
            d["active"], d["n_sen_eval"] = fact[u].copy(), int(nsen[u])
        if ftopn is not None:
            d["carried"] = ftopn[u].copy()   # the scorer's top-N codewords after the search
        out.append(d)
    if out:
        out[0]["kernel_ms"] = dict(gmm_topn=float(ms[0]), senone_mix=float(ms[1]), fsg_search=float(ms[2]),
                                   backtrace=float(ms[3]))
        out[0]["n_launches"] = int(fo.n_launches)
    del keep
    return out


def hmm_vit_eval(model, tmatid, senid, senscr, st):
    """hmm_vit_eval on one HMM (ref: src/hmm.c:741-759).  st = score[5] hist[5] out_score
    out_hist; returns (best, new st)."""
    senid = np.ascontiguousarray(senid, np.uint16)
    senscr = np.ascontiguousarray(senscr, np.int16)
    st = np.ascontiguousarray(st, np.int32).copy()
    best = C.c_int32(0)
    rv = model.lib.ssb_hmm_vit_eval(model.h, model.n_emit, int(tmatid), _ptr(senid), _ptr(senscr),
                                    _ptr(st), C.byref(best))
    _lib.check(rv, "ssb_hmm_vit_eval")
    return best.value, st


# ---------------------------------------------------------------------------- frontend
TRANSFORMS = {"dct": 0, "legacy": 1, "htk": 2}
CMN_TYPES = {"none": 0, "batch": 1, "current": 1}


def hmm_vit_eval_tp(model, tp, senscr, st):
    """Many hmm_vit_eval steps on given transition matrices: tp [n][E][E+1] uint8, senscr [n][E],
    st [n][12]; returns (best [n], new st [n][12]).  E = 3 or 5."""
    tp = np.ascontiguousarray(tp, np.uint8)
    n, E = tp.shape[0], tp.shape[1]
    senscr = np.ascontiguousarray(senscr, np.int16).reshape(n, E)
    st = np.ascontiguousarray(st, np.int32).reshape(n, 12).copy()
    best = np.zeros(n, np.int32)
    _lib.check(model.lib.ssb_hmm_vit_eval_tp(model.h, E, n, _ptr(tp), _ptr(senscr), _ptr(st), _ptr(best)),
               "ssb_hmm_vit_eval_tp")
    return best, st


class DeviceFeatures:
    """Features of a batch that stayed in HBM (Frontend.run): accepted wherever a list of
    per-utterance feature arrays is (align_batch, fsg_batch, score_batch, StateAlignBatch)."""

    def __init__(self, frontend, ptr, frame_off, dim):
        self.frontend, self.ptr, self.frame_off, self.dim = frontend, ptr, frame_off, dim

    def __len__(self):
        return len(self.frame_off) - 1


def _flat_feats(model, feats):
    """(pointer value, frame_off, keep-alive) of host arrays or DeviceFeatures."""
    if isinstance(feats, DeviceFeatures):
        if feats.dim != model.blk:
            raise SsbError("feature dimension %d does not match the model (%d)" % (feats.dim, model.blk))
        return feats.ptr, np.ascontiguousarray(feats.frame_off, np.int64), feats
    feats = [np.ascontiguousarray(f, np.float32).reshape(-1, model.blk) for f in feats]
    off = np.zeros(len(feats) + 1, np.int64)
    for i, f in enumerate(feats):
        off[i + 1] = off[i] + f.shape[0]
    feat = np.concatenate(feats) if feats else np.zeros((0, model.blk), np.float32)
    return feat.ctypes.data, off, feat


class Frontend:
    """Batched fe_t + feat_t for whole utterances (ref: src/fe_interface.c, src/fe_sigproc.c,
    src/fe_noise.c, src/cmn.c, src/feat.c): PCM -> MFCC -> CMN -> 1s_c_d_dd features.

    Frontend(hmmdir) reads <hmmdir>/feat_params.json over the reference defaults, keywords
    override single parameters (names = the reference's config keys)."""

    def __init__(self, hmmdir=None, device=0, stream=None, **params):
        self.lib = L = _lib.load()
        cfg = _lib.FeConfig()
        if hmmdir is not None:
            _lib.check(L.ssb_fe_config_from_model(os.fsencode(hmmdir), C.byref(cfg)),
                       "ssb_fe_config_from_model")
        else:
            L.ssb_fe_config_defaults(C.byref(cfg))
        for k, v in params.items():
            if k == "transform":
                v = TRANSFORMS[v] if isinstance(v, str) else v
            elif k == "cmn":
                v = CMN_TYPES[v] if isinstance(v, str) else v
            if not hasattr(cfg, k):
                raise SsbError("unknown frontend parameter " + k)
            setattr(cfg, k, v)
        self.cfg = cfg
        h = L.ssb_frontend_create(C.byref(cfg), device, C.c_void_p(stream or 0))
        if not h:
            raise SsbError("ssb_frontend_create: " + _lib.last_error())
        self.h = C.c_void_p(h)
        d = np.zeros(8, np.int32)
        L.ssb_frontend_dims(self.h, _ptr(d))
        (self.frame_size, self.frame_shift, self.fft_size, self.nfilt, self.ncep, self.feat_dim,
         self.n_coeffs) = [int(x) for x in d[:7]]
        self._keep = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.ssb_frontend_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def n_frames(self, n_samples):
        return int(self.lib.ssb_frontend_n_frames(self.h, int(n_samples)))

    def tables(self):
        out = dict(spec_start=np.zeros(self.nfilt, np.int32), filt_width=np.zeros(self.nfilt, np.int32),
                   coeffs=np.zeros(self.n_coeffs, np.float32),
                   mel_cosine=np.zeros((self.ncep, self.nfilt), np.float32),
                   lifter=np.zeros(self.ncep, np.float32), hamming=np.zeros(self.frame_size // 2))
        self.lib.ssb_frontend_tables(self.h, *[_ptr(out[k]) for k in
                                               ("spec_start", "filt_width", "coeffs", "mel_cosine",
                                                "lifter", "hamming")])
        return out

    def run(self, pcms):
        """pcms: list of int16 (or float32 in [-1,1)) sample arrays, one per utterance.
        Computes on the device and returns DeviceFeatures (nothing is copied back)."""
        if len(pcms) and any(np.asarray(p).dtype != np.asarray(pcms[0]).dtype for p in pcms):
            raise SsbError("all utterances of a batch must share one sample type")
        dt = np.asarray(pcms[0]).dtype if len(pcms) else np.dtype(np.int16)
        if dt not in (np.dtype(np.int16), np.dtype(np.float32)):
            raise SsbError("samples must be int16 or float32")
        pcms = [np.ascontiguousarray(p, dt).reshape(-1) for p in pcms]
        off = np.zeros(len(pcms) + 1, np.int64)
        for i, p in enumerate(pcms):
            off[i + 1] = off[i] + len(p)
        flat = np.concatenate(pcms) if pcms else np.zeros(0, dt)
        return self.run_raw(flat, off)

    def run_raw(self, pcm, samp_off):
        """Flat samples + offsets exactly as ssb_frontend_run takes them (pcm may be pinned)."""
        enc = 0 if pcm.dtype == np.int16 else 1
        samp_off = np.ascontiguousarray(samp_off, np.int64)
        n = self.lib.ssb_frontend_run(self.h, _ptr(pcm), enc, _ptr(samp_off), len(samp_off) - 1)
        _lib.check(int(n), "ssb_frontend_run")
        self._keep = (pcm, samp_off)
        frame_off = np.zeros(len(samp_off), np.int64)
        _lib.check(self.lib.ssb_frontend_download(self.h, _ptr(frame_off), None, None),
                   "ssb_frontend_download")
        ptr = self.lib.ssb_frontend_feat_device(self.h)
        return DeviceFeatures(self, ptr, frame_off, self.feat_dim)

    def download(self, want_mfcc=True):
        """Per-utterance (mfcc [T][ncep] before CMN, feat [T][3*ncep]) of the last run."""
        frame_off = np.zeros(len(self._keep[1]), np.int64)
        self.lib.ssb_frontend_download(self.h, _ptr(frame_off), None, None)
        G = int(frame_off[-1])
        mfcc = np.zeros((G, self.ncep), np.float32) if want_mfcc else None
        feat = np.zeros((G, self.feat_dim), np.float32)
        _lib.check(self.lib.ssb_frontend_download(self.h, None, _ptr(mfcc), _ptr(feat)),
                   "ssb_frontend_download")
        sl = [slice(int(frame_off[u]), int(frame_off[u + 1])) for u in range(len(frame_off) - 1)]
        return [(mfcc[s] if want_mfcc else None, feat[s]) for s in sl]

    def features(self, pcms):
        self.run(pcms)
        return self.download()

    def kernel_ms(self):
        ms = np.zeros(8, np.float32)
        _lib.check(self.lib.ssb_frontend_kernel_ms(self.h, _ptr(ms)), "ssb_frontend_kernel_ms")
        return dict(melspec=float(ms[0]), noise=float(ms[1]), cepstrum=float(ms[2]),
                    cmn=float(ms[3]), feat=float(ms[4]), total=float(ms[5]))


# ---------------------------------------------------------------------------- lexicon
class Lexicon:
    """dict_t + dict2pid_t of the reference (ref: src/dict.c, src/dict2pid.c): word ids,
    pronunciations and the word -> phone-chain expansion of alignment_populate
    (ref: src/ps_alignment.c:133-248).  Host only.  Defaults to the model directory's
    dict.txt / noisedict.txt like `decoder_init` expands them (ref: src/decoder.c:122-123)."""

    def __init__(self, model, dictfile=None, fdictfile=None, hmmdir=None):
        self.model = model
        self.lib = model.lib
        if hmmdir is not None:
            dictfile = dictfile or os.path.join(hmmdir, "dict.txt")
            fd = os.path.join(hmmdir, "noisedict.txt")
            fdictfile = fdictfile or (fd if os.path.exists(fd) else None)
        h = self.lib.ssb_lexicon_load(model.h, os.fsencode(dictfile) if dictfile else None,
                                      os.fsencode(fdictfile) if fdictfile else None)
        if not h:
            raise SsbError("ssb_lexicon_load: " + _lib.last_error())
        self.h = C.c_void_p(h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.ssb_lexicon_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return int(self.lib.ssb_lexicon_size(self.h))

    def wordid(self, word):
        return int(self.lib.ssb_lexicon_wordid(self.h, word.encode("utf-8")))

    def wordstr(self, wid):
        s = self.lib.ssb_lexicon_wordstr(self.h, int(wid))
        return s.decode("utf-8") if s is not None else None

    def pron(self, wid):
        out = np.zeros(64, np.int32)
        n = int(self.lib.ssb_lexicon_pron(self.h, int(wid), _ptr(out), 64))
        return out[:n]

    def is_filler(self, wid):
        return bool(self.lib.ssb_lexicon_is_filler(self.h, int(wid)))

    def basewid(self, wid):
        """dict_basewid: the id of "word" for "word(2)"."""
        return int(self.lib.ssb_lexicon_basewid(self.h, int(wid)))

    def populate(self, wids, start=None, dur=None):
        """Phone chain of a word sequence: dict(ssid, tmat, ci, parent[, sf, ef]) -- with
        word windows (start, dur) the sf/ef arrays state_align_search_init derives."""
        wids = np.ascontiguousarray(wids, np.int32)
        n = int(self.lib.ssb_chain_populate(self.h, _ptr(wids), len(wids), None, None, None, None, 0))
        _lib.check(n, "ssb_chain_populate")
        ssid, tmat, ci, parent = (np.zeros(n, np.int32) for _ in range(4))
        _lib.check(int(self.lib.ssb_chain_populate(self.h, _ptr(wids), len(wids), _ptr(ssid), _ptr(tmat),
                                                   _ptr(ci), _ptr(parent), n)), "ssb_chain_populate")
        out = dict(ssid=ssid, tmat=tmat, ci=ci, parent=parent)
        if start is not None:
            start, dur = np.asarray(start, np.int32), np.asarray(dur, np.int32)
            out["sf"], out["ef"] = windows(start[parent], dur[parent])
        else:
            out["sf"], out["ef"] = windows(np.zeros(n, np.int32), np.zeros(n, np.int32))
        return out


def _lexicon_align_graph(self, text, **cfg):
    """decoder_set_align_text + fsg_search_init + fsg_lextree_init (ref: src/decoder.c:685-735,
    src/fsg_search.c:171-260, src/fsg_lextree.c:226-716): the flattened alignment grammar of
    `text` in the form fsg_batch takes, plus `words` (the grammar's vocabulary: link[:, 3]
    indexes it) and `dict_wid`."""
    c = _lib.FsgConfig()
    self.lib.ssb_fsg_config_defaults(C.byref(c))
    for k, v in cfg.items():
        if not hasattr(c, k):
            raise SsbError("unknown search parameter " + k)
        setattr(c, k, v)
    b = self.lib.ssb_fsg_build_align(self.h, text.encode("utf-8"), C.byref(c))
    if not b:
        raise SsbError("ssb_fsg_build_align: " + _lib.last_error())
    return _built_graph(self, C.c_void_p(b))


def _built_graph(self, b):
    try:
        g = self.lib.ssb_fsg_built_graph(b).contents

        def arr(ptr, shape, dt):
            n = int(np.prod(shape))
            if n == 0:
                return np.zeros(shape, dt)
            buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dt).reshape(shape).copy()
        out = dict(n_state=g.n_state, start=g.start, final=g.final, n_ciphone=g.n_ciphone, sil=g.sil,
                   beam=g.beam, pbeam=g.pbeam, wbeam=g.wbeam, maxhmmpf=g.maxhmmpf,
                   link=arr(g.link4, (g.n_link, 4), np.int32),
                   link_flag=arr(g.link_flag, (g.n_link,), np.uint8),
                   arc_off=arr(g.arc_off, (g.n_state + 1,), np.int32),
                   root=arr(g.root, (g.n_state,), np.int32),
                   pnode=arr(g.pnode8, (g.n_pnode, 8), np.int32),
                   ctxt=arr(g.ctxt, (g.n_pnode, 4), np.uint32))
        words, dwids = [], []
        for i in range(int(self.lib.ssb_fsg_built_n_words(b))):
            dw = C.c_int32(-1)
            words.append(self.lib.ssb_fsg_built_word(b, i, C.byref(dw)).decode("utf-8"))
            dwids.append(dw.value)
        out["words"], out["dict_wid"] = words, np.array(dwids, np.int32)
        return out
    finally:
        self.lib.ssb_fsg_built_free(b)


Lexicon.align_graph = _lexicon_align_graph


def _lexicon_fsg_graph(self, n_state, start, final, transitions, null_closure=True, **cfg):
    """Any grammar from its transition list [(from, to, prob, word or None), ...] in the order the
    reference's reader / JSGF compiler adds them (ref: src/fsg_model.c:506-690): decoder_set_fsg's
    graph, flattened like align_graph's (ssb_fsg_build)."""
    c = _lib.FsgConfig()
    self.lib.ssb_fsg_config_defaults(C.byref(c))
    for k, v in cfg.items():
        if not hasattr(c, k):
            raise SsbError("unknown search parameter " + k)
        setattr(c, k, v)
    n = len(transitions)
    fr = np.array([t[0] for t in transitions], np.int32)
    to = np.array([t[1] for t in transitions], np.int32)
    pr = np.array([t[2] for t in transitions], np.float32)
    words = (C.c_char_p * max(n, 1))()
    for i, t in enumerate(transitions):
        words[i] = t[3].encode("utf-8") if t[3] else None
    b = self.lib.ssb_fsg_build(self.h, int(n_state), int(start), int(final), n, _ptr(fr), _ptr(to), _ptr(pr),
                               words, 1 if null_closure else 0, C.byref(c))
    if not b:
        raise SsbError("ssb_fsg_build: " + _lib.last_error())
    return _built_graph(self, C.c_void_p(b))


def read_fsg_file(path):
    """The reference's text FSG format (ref: src/fsg_model.c:506-690): returns
    (n_state, start, final, [(from, to, prob, word or None), ...])."""
    n_state = start = final = None
    trans = []
    for ln in open(path, encoding="utf-8"):
        ln = ln.split("#", 1)[0].split()
        if not ln:
            continue
        k = ln[0]
        if k in ("NUM_STATES", "N"):
            n_state = int(ln[1])
        elif k in ("START_STATE", "S"):
            start = int(ln[1])
        elif k in ("FINAL_STATE", "F"):
            final = int(ln[1])
        elif k in ("TRANSITION", "T"):
            trans.append((int(ln[1]), int(ln[2]), float(ln[3]), ln[4] if len(ln) > 4 else None))
    if None in (n_state, start, final):
        raise SsbError("%s: not an FSG file" % path)
    return n_state, start, final, trans


Lexicon.fsg_graph = _lexicon_fsg_graph


# ---------------------------------------------------------------------------- search drop-in
class Search:
    """A search object behind the reference's search_module_t interface
    (ref: include/soundswallower/search_module.h:72-113): every method below goes through the
    object's own vtable, the way the search_module_* macros do.  Features come either from a
    `source(frame_idx) -> 39 floats or None` callable (the part acmod plays in the reference)
    or from feed()."""

    def __init__(self, model, lexicon, ptr, source=None, cb=None):
        if not ptr:
            raise SsbError("search init: " + _lib.last_error())
        self.model, self.lexicon = model, lexicon
        self.ptr = C.c_void_p(ptr)
        self.base = C.cast(self.ptr, C.POINTER(_lib.SearchBase)).contents
        self.vt = self.base.vt.contents
        self._cb, self._source = cb, source

    @staticmethod
    def _make_source(model, source):
        if source is None:
            return None, None
        hold = {}

        def get(_ctx, frame_idx):
            x = source(int(frame_idx))
            if x is None:
                return None
            hold["x"] = np.ascontiguousarray(x, np.float32).reshape(model.blk)
            return hold["x"].ctypes.data
        return _lib.FEAT_SOURCE_FN(get), hold

    type = property(lambda self: self.base.type.decode())
    name = property(lambda self: self.base.name.decode())

    def feed(self, feat):
        feat = np.ascontiguousarray(feat, np.float32).reshape(-1, self.model.blk)
        return _lib.check(int(self.model.lib.ssb_search_feed(self.ptr, _ptr(feat), len(feat))),
                          "ssb_search_feed")

    def start(self):
        return int(self.vt.start(self.ptr))

    def step(self, frame_idx):
        return int(self.vt.step(self.ptr, int(frame_idx)))

    def finish(self):
        return int(self.vt.finish(self.ptr))

    def forward(self, feat):
        """search_module_forward over a whole utterance (ref: src/decoder.c:935-957)."""
        self.feed(feat)
        n = 0
        for t in range(len(feat)):
            if self.step(t) < 0:
                raise SsbError("search step: " + _lib.last_error())
            n += 1
        return n

    def hyp(self):
        """(hypothesis string or None, score)"""
        sc = C.c_int32(0)
        h = self.vt.hyp(self.ptr, C.byref(sc))
        return (h.decode("utf-8") if h is not None else None), int(sc.value)

    def seg(self):
        """[(word, sf, ef, ascr, lscr)] through seg_iter / seg_next."""
        out = []
        it = self.vt.seg_iter(self.ptr)
        while it:
            s = C.cast(C.c_void_p(it), C.POINTER(_lib.SegIter)).contents
            out.append((s.word.decode("utf-8"), int(s.sf), int(s.ef), int(s.ascr), int(s.lscr)))
            it = s.vt.contents.seg_next(it)
        return out

    def final_active(self):
        """acmod's active-senone flags as the grammar search left them (uint32 words)."""
        out = np.zeros((self.model.n_sen + 31) // 32, np.uint32)
        _lib.check(self.model.lib.ssb_search_final_active(self.ptr, _ptr(out)), "ssb_search_final_active")
        return out

    def set_init_active(self, bits):
        """The flags the aligner finds in acmod when it starts (it never clears them)."""
        bits = np.ascontiguousarray(bits, np.uint32) if bits is not None else None
        _lib.check(self.model.lib.ssb_search_set_init_active(self.ptr, _ptr(bits) if bits is not None else None),
                   "ssb_search_set_init_active")

    def final_topn(self):
        """The scorer's carried top-N codewords as the grammar search left them, uint8 [CS][4]."""
        out = np.zeros((self.model.n_mgau * self.model.n_feat, 4), np.uint8)
        _lib.check(self.model.lib.ssb_search_final_topn(self.ptr, _ptr(out)), "ssb_search_final_topn")
        return out

    def set_init_topn(self, cw):
        cw = np.ascontiguousarray(cw, np.uint8) if cw is not None else None
        _lib.check(self.model.lib.ssb_search_set_init_topn(self.ptr, _ptr(cw) if cw is not None else None),
                   "ssb_search_set_init_topn")

    def alignment(self, level):
        """alignment_words/phones/states: int32 [n][5] id start duration score parent."""
        lvl = {"words": 0, "phones": 1, "states": 2}[level]
        n = _lib.check(int(self.model.lib.ssb_search_alignment(self.ptr, lvl, None, 0)), "ssb_search_alignment")
        out = np.zeros((n, 5), np.int32)
        self.model.lib.ssb_search_alignment(self.ptr, lvl, _ptr(out), n)
        return out

    def close(self):
        if getattr(self, "ptr", None):
            self.vt.free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def state_align_search(model, lexicon, wids, wstart=None, wdur=None, name="_state_align", source=None):
    """state_align_search_init (ref: src/state_align_search.c:429-474)."""
    wids = np.ascontiguousarray(wids, np.int32)
    ws = np.ascontiguousarray(wstart if wstart is not None else np.zeros(len(wids)), np.int32)
    wd = np.ascontiguousarray(wdur if wdur is not None else np.zeros(len(wids)), np.int32)
    cb, hold = Search._make_source(model, source)
    p = model.lib.ssb_state_align_search_init(name.encode(), model.h, lexicon.h, _ptr(wids), _ptr(ws),
                                              _ptr(wd), len(wids), C.cast(cb, C.c_void_p) if cb else None,
                                              None)
    return Search(model, lexicon, p, hold, cb)


def fsg_search(model, lexicon, text, name="_default", source=None, **cfg):
    """decoder_set_align_text + fsg_search_init (ref: src/decoder.c:685-735,
    src/fsg_search.c:171-260) for the alignment grammar of `text`."""
    c = _lib.FsgConfig()
    model.lib.ssb_fsg_config_defaults(C.byref(c))
    for k, v in cfg.items():
        if not hasattr(c, k):
            raise SsbError("unknown search parameter " + k)
        setattr(c, k, v)
    b = model.lib.ssb_fsg_build_align(lexicon.h, text.encode("utf-8"), C.byref(c))
    if not b:
        raise SsbError("ssb_fsg_build_align: " + _lib.last_error())
    cb, hold = Search._make_source(model, source)
    p = model.lib.ssb_fsg_search_init(name.encode(), model.h, lexicon.h, C.c_void_p(b),
                                      C.cast(cb, C.c_void_p) if cb else None, None)
    return Search(model, lexicon, p, hold, cb)


# ---------------------------------------------------------------------------- two-pass alignment
def _left_active(p1):
    """Per utterance the senones whose acmod flags the first pass left set."""
    out = []
    for r in p1:
        bits = r.get("active")
        out.append([] if bits is None else
                   [int(w * 32 + b) for w in np.nonzero(bits)[0] for b in range(32) if (int(bits[w]) >> b) & 1])
    return out


def align_texts(model, lexicon, feats, texts, **search_cfg):
    """The reference's forced alignment of `soundswallower --align` for a batch of
    (utterance, transcript) pairs, both passes on the GPU:

      pass 1  decoder_set_align_text + search_module_forward (ref: src/decoder.c:685-735,
              935-957): grammar search of the transcript -> word segmentation;
      pass 2  decoder_alignment (ref: src/decoder.c:737-798): the words of pass 1 with their
              frame windows -> alignment_populate -> state_align_search -> backtrace ->
              alignment_propagate.

    feats: list of feature arrays, or DeviceFeatures from Frontend.run (audio never leaves the
    GPU).  Returns per utterance None when the transcript does not match the audio (pass 1
    has no final exit), else dict(words=[(word, start, dur, score)], phones=[(phone, start,
    dur, score, word index)], states=int32 [n][5] (senone, start, dur, score, phone index),
    hyp_score)."""
    graphs = [lexicon.align_graph(t, **search_cfg) for t in texts]
    # the reference's default mode (compallsen = no) wherever the model allows it: pass-1 path
    # scores are then the CLI's, and pass 2 starts from the flags pass 1 left in acmod
    active = model.fsg_active_ok
    p1 = fsg_batch(model, feats, graphs, utt_graph=np.arange(len(texts), dtype=np.int32),
                   compallsen=not active)
    chains, metas = [], []
    for g, r in zip(graphs, p1):
        if r["rv"] != 0 or r["exit"] <= 0:
            chains.append(dict(ssid=np.zeros(0, np.int32), tmat=np.zeros(0, np.int32),
                               sf=np.zeros(0, np.int32), ef=np.zeros(0, np.int32)))
            metas.append(None)
            continue
        segs = r["segs"]
        fw = g["link"][segs[:, 0], 3]
        keep = fw >= 0                      # null transitions carry no word
        wids = g["dict_wid"][fw[keep]]
        start, dur = segs[keep, 1], segs[keep, 2] - segs[keep, 1] + 1
        c = lexicon.populate(wids, start, dur)
        chains.append(c)
        metas.append((wids, start, dur, c, int(r["hyp_score"])))
    carried = [r.get("carried") for r in p1]
    p2 = align_batch(model, feats, chains, init_active=_left_active(p1) if active else None,
                     init_topn=carried if all(c is not None for c in carried) else None)
    arrays = model.arrays()
    E = model.n_emit
    out = []
    for meta, r in zip(metas, p2):
        if meta is None or r["rv"] != 0:
            out.append(None)
            continue
        wids, wstart, wdur, c, hyp = meta
        ps, pd, pc = propagate(r["start"], r["dur"], r["score"], E)
        parent = c["parent"]
        words = []
        for i, w in enumerate(wids):   # alignment_propagate, phones -> words
            sel = parent == i
            words.append((lexicon.wordstr(int(w)), int(ps[sel][0]), int(pd[sel].sum()), int(pc[sel].sum())))
        phones = [(model.ciname(int(c["ci"][i])), int(ps[i]), int(pd[i]), int(pc[i]), int(parent[i]))
                  for i in range(len(ps))]
        sen = arrays["sseq"][c["ssid"]].reshape(-1).astype(np.int32)
        states = np.stack([sen, r["start"], r["dur"], r["score"],
                           np.repeat(np.arange(len(ps), dtype=np.int32), E)], 1).astype(np.int32)
        out.append(dict(words=words, phones=phones, states=states, hyp_score=hyp))
    return out


class TextAlignment:
    """ssb_align_texts: `soundswallower --align` for a batch in one C call -- both passes on the
    GPU, chains and JSON on the host in C++.  status(u) / hyp(u) / entries(u, level) / json(u)."""

    def __init__(self, model, lexicon, feats, texts, align_level=1, frate=0, **search_cfg):
        self.model, self.lexicon = model, lexicon
        self.lib = model.lib
        c = None
        if search_cfg:
            c = _lib.FsgConfig()
            self.lib.ssb_fsg_config_defaults(C.byref(c))
            for k, v in search_cfg.items():
                if not hasattr(c, k):
                    raise SsbError("unknown search parameter " + k)
                setattr(c, k, v)
        fptr, off, keep = _flat_feats(model, feats)
        arr = (C.c_char_p * max(len(texts), 1))(*[t.encode("utf-8") for t in texts])
        self.n = len(texts)
        r = self.lib.ssb_align_texts(model.h, lexicon.h, C.c_void_p(fptr), _ptr(off), arr, self.n,
                                     C.byref(c) if c is not None else None, int(align_level), int(frate))
        if not r:
            raise SsbError("ssb_align_texts: " + _lib.last_error())
        self.r = C.c_void_p(r)
        del keep

    def close(self):
        if getattr(self, "r", None):
            self.lib.ssb_text_align_free(self.r)
            self.r = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def status(self, u):
        """(rv, hyp_score, n_frames): rv 0 ok, -1 no hypothesis, -2 second pass failed."""
        sc, nf = C.c_int32(0), C.c_int32(0)
        rv = int(self.lib.ssb_text_align_status(self.r, int(u), C.byref(sc), C.byref(nf)))
        return rv, int(sc.value), int(nf.value)

    def hyp(self, u):
        h = self.lib.ssb_text_align_hyp(self.r, int(u))
        return h.decode("utf-8") if h is not None else None

    def entries(self, u, level):
        lvl = {"words": 0, "phones": 1, "states": 2, "seg": 3}[level]
        n = _lib.check(int(self.lib.ssb_text_align_entries(self.r, int(u), lvl, None, 0)), "ssb_text_align_entries")
        out = np.zeros((n, 5), np.int32)
        self.lib.ssb_text_align_entries(self.r, int(u), lvl, _ptr(out), n)
        return out

    def render(self, start=0., align_level=1):
        """All JSON lines at once (host threads); json(u, same arguments) then just returns them."""
        _lib.check(self.lib.ssb_text_align_render(self.r, float(start), int(align_level)), "ssb_text_align_render")

    def json(self, u, start=0., align_level=1):
        j = self.lib.ssb_text_align_json(self.r, int(u), float(start), int(align_level))
        return j.decode("utf-8") if j is not None else None

    def kernel_ms(self):
        ms = np.zeros(8, np.float32)
        self.lib.ssb_text_align_kernel_ms(self.r, _ptr(ms))
        return dict(gmm_topn=float(ms[0]), senone_mix=float(ms[1]), fsg_search=float(ms[2]), backtrace=float(ms[3]),
                    wall_pass1=float(ms[4]), wall_chains=float(ms[5]), wall_pass2=float(ms[6]), wall_total=float(ms[7]))


from .decoder import (Alignment, AlignmentEntry, Decoder, Hyp, Seg, get_audio_data)  # noqa: E402
