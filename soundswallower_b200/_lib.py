"""ctypes binding of libssb200.so (the C ABI declared in include/ssb200.h).

There is deliberately no fallback: if the shared library is missing this module
raises, and if it loads but finds no B200 every compute call returns -1 and the
wrappers raise SsbError with the library's message.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libssb200.so")


class SsbError(RuntimeError):
    pass


class Config(C.Structure):
    """ssb_config_t -- the reference config keys that reach the hot path
    (ref: include/soundswallower/config_defs.h:78-257)."""
    _fields_ = [("logbase", C.c_double), ("varfloor", C.c_float), ("mixwfloor", C.c_double),
                ("tmatfloor", C.c_double), ("topn", C.c_int32), ("ds", C.c_int32),
                ("device", C.c_int32), ("topn_beam", C.c_int32 * 4)]


class FeConfig(C.Structure):
    """ssb_fe_config_t -- the frontend keys of the reference config
    (ref: include/soundswallower/config_defs.h:296-449)."""
    _fields_ = [(k, C.c_int32) for k in ("samprate", "frate", "ncep", "nfft", "nfilt", "lifter",
                                          "remove_dc", "remove_noise", "unit_area", "round_filters",
                                          "doublebw", "transform", "cmn", "varnorm")] + \
               [(k, C.c_float) for k in ("wlen", "alpha", "lowerf", "upperf")]


class FsgConfig(C.Structure):
    """ssb_fsg_config_t -- the search keys of the reference config
    (ref: include/soundswallower/config_defs.h:79-159)."""
    _fields_ = [("beam", C.c_double), ("pbeam", C.c_double), ("wbeam", C.c_double),
                ("lw", C.c_float), ("wip", C.c_float), ("pip", C.c_float),
                ("silprob", C.c_float), ("fillprob", C.c_float), ("maxhmmpf", C.c_int32),
                ("fsgusefiller", C.c_int32), ("fsgusealtpron", C.c_int32)]


class AlignIn(C.Structure):
    _fields_ = [("n_utts", C.c_int32), ("feat", C.POINTER(C.c_float)),
                ("frame_off", C.POINTER(C.c_int64)), ("phone_off", C.POINTER(C.c_int64)),
                ("ssid", C.POINTER(C.c_int32)), ("tmat", C.POINTER(C.c_int32)),
                ("sf", C.POINTER(C.c_int32)), ("ef", C.POINTER(C.c_int32)),
                ("init_active", C.POINTER(C.c_uint32)), ("compallsen", C.c_int32),
                ("init_topn", C.POINTER(C.c_uint8))]


class AlignOut(C.Structure):
    _fields_ = [("st_start", C.POINTER(C.c_int32)), ("st_dur", C.POINTER(C.c_int32)),
                ("st_score", C.POINTER(C.c_int32)), ("utt_rv", C.POINTER(C.c_int32)),
                ("utt_best", C.POINTER(C.c_int32)), ("utt_renorm", C.POINTER(C.c_int32)),
                ("chain_scr", C.POINTER(C.c_int16)), ("tokens", C.POINTER(C.c_int32))]


class FsgGraph(C.Structure):
    """ssb_fsg_graph_t: the reference's fsg_model_t + fsg_lextree_t flattened."""
    _fields_ = [(k, C.c_int32) for k in ("n_state", "start", "final", "n_link", "n_pnode", "n_ciphone",
                                          "sil", "beam", "pbeam", "wbeam", "maxhmmpf")] + \
               [("link4", C.c_void_p), ("link_flag", C.c_void_p), ("arc_off", C.c_void_p),
                ("root", C.c_void_p), ("pnode8", C.c_void_p), ("ctxt", C.c_void_p)]


class FsgIn(C.Structure):
    _fields_ = [("n_utts", C.c_int32), ("feat", C.c_void_p), ("frame_off", C.c_void_p),
                ("n_graphs", C.c_int32), ("graphs", C.POINTER(FsgGraph)), ("utt_graph", C.c_void_p),
                ("hist_cap", C.c_int32), ("max_seg", C.c_int32), ("active_lists", C.c_int32),
                ("partial", C.c_int32)]


class FsgOut(C.Structure):
    _fields_ = [("segs", C.c_void_p), ("n_seg", C.c_void_p), ("hyp_score", C.c_void_p),
                ("exit_bp", C.c_void_p), ("utt_rv", C.c_void_p), ("n_hist", C.c_void_p),
                ("n_hmm_eval", C.c_void_p), ("hist9", C.c_void_p), ("kernel_ms", C.c_void_p),
                ("n_launches", C.c_int32), ("final_active", C.c_void_p), ("n_sen_eval", C.c_void_p),
                ("final_topn", C.c_void_p)]


MGAU_FRAME_EVAL = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_int16), C.POINTER(C.c_uint8),
                              C.c_int32, C.POINTER(C.POINTER(C.c_float)), C.c_int32, C.c_int32)


class MgauFuncs(C.Structure):
    """ssb_mgaufuncs_t == mgaufuncs_t (ref: include/soundswallower/acmod.h:93-106)."""
    _fields_ = [("name", C.c_char_p), ("frame_eval", MGAU_FRAME_EVAL),
                ("transform", C.c_void_p), ("free", C.c_void_p)]


class MgauBase(C.Structure):
    """ssb_mgau_t == mgau_t (ref: acmod.h:108-111): {vt, frame_idx}."""
    _fields_ = [("vt", C.POINTER(MgauFuncs)), ("frame_idx", C.c_int)]


SEARCH_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)
SEARCH_STEP_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int)
SEARCH_HYP_FN = C.CFUNCTYPE(C.c_char_p, C.c_void_p, C.POINTER(C.c_int32))
SEARCH_SEG_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p)
SEARCH_FREE_FN = C.CFUNCTYPE(None, C.c_void_p)
FEAT_SOURCE_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_int)


class SearchFuncs(C.Structure):
    """ssb_searchfuncs_t == searchfuncs_t (ref: include/soundswallower/search_module.h:72-84)."""
    _fields_ = [("start", SEARCH_FN), ("step", SEARCH_STEP_FN), ("finish", SEARCH_FN),
                ("reinit", C.c_void_p), ("free", SEARCH_FREE_FN), ("lattice", C.c_void_p),
                ("hyp", SEARCH_HYP_FN), ("prob", C.c_void_p), ("seg_iter", SEARCH_SEG_FN)]


class SearchBase(C.Structure):
    """ssb_search_t == search_module_t (ref: search_module.h:89-113)."""
    _fields_ = [("vt", C.POINTER(SearchFuncs)), ("type", C.c_char_p), ("name", C.c_char_p),
                ("config", C.c_void_p), ("acmod", C.c_void_p), ("dict", C.c_void_p),
                ("d2p", C.c_void_p), ("hyp_str", C.c_char_p), ("dag", C.c_void_p),
                ("last_link", C.c_void_p), ("post", C.c_int32), ("n_words", C.c_int32),
                ("start_wid", C.c_int32), ("silence_wid", C.c_int32), ("finish_wid", C.c_int32)]


class SegFuncs(C.Structure):
    """ssb_segfuncs_t == ps_segfuncs_t (ref: search_module.h:157-160)."""
    _fields_ = [("seg_next", SEARCH_SEG_FN), ("seg_free", SEARCH_FREE_FN)]


class SegIter(C.Structure):
    """ssb_seg_iter_t == seg_iter_t (ref: search_module.h:165-174)."""
    _fields_ = [("vt", C.POINTER(SegFuncs)), ("search", C.c_void_p), ("word", C.c_char_p),
                ("sf", C.c_int32), ("ef", C.c_int32), ("ascr", C.c_int32), ("lscr", C.c_int32),
                ("prob", C.c_int32)]


# every symbol include/ssb200.h declares
SYMBOLS = [
    "ssb_version", "ssb_last_error", "ssb_device_count", "ssb_config_defaults",
    "ssb_model_load", "ssb_model_kind", "ssb_model_ciphone_str", "ssb_model_free", "ssb_model_dims", "ssb_model_copy", "ssb_model_phones",
    "ssb_mgau_init", "ssb_mgau_frame_eval", "ssb_mgau_reset", "ssb_mgau_free", "ssb_mgau_own_model", "ssb_mgau_model",
    "ssb_plan_chain", "ssb_batch_create", "ssb_batch_free", "ssb_batch_upload", "ssb_batch_run",
    "ssb_batch_download", "ssb_batch_debug_tokens", "ssb_batch_kernel_ms",
    "ssb_batch_n_launches", "ssb_batch_stats", "ssb_batch_band_state_frames", "ssb_batch_n_segments", "ssb_align_batch", "ssb_align_batch_multi", "ssb_pipeline_create",
    "ssb_pipeline_align", "ssb_pipeline_submit", "ssb_pipeline_collect", "ssb_pipeline_set_overlap", "ssb_pipeline_n_launches", "ssb_pipeline_n_chunks", "ssb_pipeline_trace", "ssb_pipeline_free",
    "ssb_score_batch", "ssb_lexicon_basewid", "ssb_fsg_built_is_filler",
    "ssb_state_align_search_init", "ssb_fsg_search_init", "ssb_search_feed", "ssb_search_alignment", "ssb_search_final_active",
    "ssb_search_set_init_active", "ssb_search_final_topn", "ssb_search_set_init_topn",
    "ssb_model_fsg_active_ok", "ssb_align_texts", "ssb_text_align_free",
    "ssb_text_align_status", "ssb_text_align_hyp", "ssb_text_align_entries", "ssb_text_align_json",
    "ssb_text_align_kernel_ms", "ssb_text_align_render", "ssb_device_cache_trim", "ssb_hmm_vit_eval_tp",
    "ssb_topn_batch", "ssb_tc_probe", "ssb_tc_hot_mask", "ssb_fsg_batch", "ssb_hmm_vit_eval",
    "ssb_fe_config_defaults", "ssb_fe_config_from_model", "ssb_frontend_create",
    "ssb_frontend_free", "ssb_frontend_dims", "ssb_frontend_n_frames", "ssb_frontend_tables",
    "ssb_frontend_run", "ssb_frontend_download", "ssb_frontend_feat_device",
    "ssb_frontend_kernel_ms",
    "ssb_lexicon_load", "ssb_lexicon_free", "ssb_lexicon_size", "ssb_lexicon_wordid",
    "ssb_lexicon_wordstr", "ssb_lexicon_pron", "ssb_lexicon_is_filler", "ssb_chain_populate",
    "ssb_fsg_config_defaults", "ssb_fsg_build_align", "ssb_fsg_build", "ssb_fsg_build_logp", "ssb_fsg_built_graph",
    "ssb_fsg_built_n_words", "ssb_fsg_built_word", "ssb_fsg_built_free",
]

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SsbError("libssb200.so is not built (run `python -m soundswallower_b200._build`); "
                       "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    P = C.POINTER
    L.ssb_version.restype = C.c_int
    L.ssb_last_error.restype = C.c_char_p
    L.ssb_device_count.restype = C.c_int
    L.ssb_config_defaults.argtypes = [P(Config)]
    L.ssb_config_defaults.restype = None
    L.ssb_model_load.restype = vp
    L.ssb_model_load.argtypes = [C.c_char_p, P(Config)]
    L.ssb_model_free.argtypes = [vp]
    L.ssb_model_free.restype = None
    L.ssb_model_dims.argtypes = [vp, P(i32)]
    L.ssb_model_copy.argtypes = [vp] + [vp] * 8
    L.ssb_model_phones.argtypes = [vp, vp, vp, vp]
    L.ssb_mgau_init.restype = P(MgauBase)
    L.ssb_mgau_init.argtypes = [vp]
    L.ssb_mgau_frame_eval.argtypes = [P(MgauBase), P(C.c_int16), P(C.c_uint8), i32,
                                      P(P(C.c_float)), i32, i32]
    L.ssb_mgau_reset.argtypes = [P(MgauBase)]
    L.ssb_mgau_reset.restype = None
    L.ssb_mgau_free.argtypes = [P(MgauBase)]
    L.ssb_mgau_free.restype = None
    L.ssb_mgau_own_model.argtypes = [P(MgauBase), C.c_int]
    L.ssb_mgau_own_model.restype = None
    L.ssb_plan_chain.argtypes = [i32, i32, vp, vp, vp]
    L.ssb_batch_create.restype = vp
    L.ssb_batch_create.argtypes = [vp, vp]
    L.ssb_batch_free.argtypes = [vp]
    L.ssb_batch_free.restype = None
    L.ssb_batch_upload.argtypes = [vp, P(AlignIn)]
    L.ssb_batch_run.argtypes = [vp]
    L.ssb_batch_download.argtypes = [vp, P(AlignOut)]
    L.ssb_batch_debug_tokens.argtypes = [vp, C.c_int]
    L.ssb_batch_kernel_ms.argtypes = [vp, P(C.c_float)]
    L.ssb_batch_n_launches.argtypes = [vp]
    L.ssb_batch_stats.argtypes = [vp, P(i64)]
    L.ssb_batch_band_state_frames.argtypes = [vp]
    L.ssb_batch_band_state_frames.restype = i64
    L.ssb_batch_n_segments.argtypes = [vp]
    L.ssb_batch_n_segments.restype = C.c_int32
    L.ssb_align_batch.argtypes = [vp, P(AlignIn), P(AlignOut)]
    L.ssb_align_batch_multi.argtypes = [P(vp), i32, P(AlignIn), P(AlignOut)]
    L.ssb_pipeline_create.restype = vp
    L.ssb_pipeline_create.argtypes = [vp, i32, i64]
    L.ssb_pipeline_align.argtypes = [vp, P(AlignIn), P(AlignOut)]
    L.ssb_pipeline_submit.restype = i64
    L.ssb_pipeline_submit.argtypes = [vp, P(AlignIn), P(AlignOut)]
    L.ssb_pipeline_collect.argtypes = [vp, i64]
    L.ssb_pipeline_set_overlap.argtypes = [vp, i32]
    L.ssb_pipeline_n_launches.argtypes = [vp]
    L.ssb_pipeline_n_chunks.argtypes = [vp]
    L.ssb_pipeline_trace.argtypes = [vp, vp, i32]
    L.ssb_pipeline_free.restype = None
    L.ssb_pipeline_free.argtypes = [vp]
    L.ssb_score_batch.restype = i64
    L.ssb_score_batch.argtypes = [vp, vp, vp, i32, vp]
    L.ssb_topn_batch.restype = i64
    L.ssb_topn_batch.argtypes = [vp, vp, vp, i32, vp, vp]
    L.ssb_tc_probe.restype = i64
    L.ssb_tc_probe.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, vp]
    L.ssb_tc_hot_mask.argtypes = [vp, vp]
    L.ssb_fsg_batch.argtypes = [vp, P(FsgIn), P(FsgOut)]
    L.ssb_hmm_vit_eval.argtypes = [vp, i32, i32, vp, vp, vp, P(i32)]
    L.ssb_hmm_vit_eval_tp.argtypes = [vp, i32, i32, vp, vp, vp, vp]
    L.ssb_model_kind.argtypes = [vp]
    L.ssb_model_ciphone_str.restype = C.c_char_p
    L.ssb_model_ciphone_str.argtypes = [vp, i32]
    L.ssb_lexicon_load.restype = vp
    L.ssb_lexicon_load.argtypes = [vp, C.c_char_p, C.c_char_p]
    L.ssb_lexicon_free.restype = None
    L.ssb_lexicon_free.argtypes = [vp]
    L.ssb_lexicon_size.argtypes = [vp]
    L.ssb_lexicon_wordid.argtypes = [vp, C.c_char_p]
    L.ssb_lexicon_wordstr.restype = C.c_char_p
    L.ssb_lexicon_wordstr.argtypes = [vp, i32]
    L.ssb_lexicon_pron.argtypes = [vp, i32, vp, i32]
    L.ssb_lexicon_is_filler.argtypes = [vp, i32]
    L.ssb_lexicon_basewid.argtypes = [vp, i32]
    L.ssb_fsg_built_is_filler.argtypes = [vp, i32]
    L.ssb_state_align_search_init.restype = vp
    L.ssb_state_align_search_init.argtypes = [C.c_char_p, vp, vp, vp, vp, vp, i32, vp, vp]
    L.ssb_fsg_search_init.restype = vp
    L.ssb_fsg_search_init.argtypes = [C.c_char_p, vp, vp, vp, vp, vp]
    L.ssb_search_feed.argtypes = [vp, vp, i32]
    L.ssb_search_alignment.argtypes = [vp, i32, vp, i32]
    L.ssb_search_final_active.argtypes = [vp, vp]
    L.ssb_search_set_init_active.argtypes = [vp, vp]
    L.ssb_model_fsg_active_ok.argtypes = [vp]
    L.ssb_search_final_topn.argtypes = [vp, vp]
    L.ssb_search_set_init_topn.argtypes = [vp, vp]
    L.ssb_align_texts.restype = vp
    L.ssb_align_texts.argtypes = [vp, vp, vp, vp, P(C.c_char_p), i32, vp, i32, i32]
    L.ssb_text_align_free.restype = None
    L.ssb_text_align_free.argtypes = [vp]
    L.ssb_text_align_status.argtypes = [vp, i32, P(i32), P(i32)]
    L.ssb_text_align_hyp.restype = C.c_char_p
    L.ssb_text_align_hyp.argtypes = [vp, i32]
    L.ssb_text_align_entries.argtypes = [vp, i32, i32, vp, i32]
    L.ssb_text_align_json.restype = C.c_char_p
    L.ssb_text_align_json.argtypes = [vp, i32, C.c_double, i32]
    L.ssb_text_align_kernel_ms.argtypes = [vp, vp]
    L.ssb_text_align_render.argtypes = [vp, C.c_double, i32]
    L.ssb_chain_populate.argtypes = [vp, vp, i32, vp, vp, vp, vp, i32]
    L.ssb_fsg_config_defaults.restype = None
    L.ssb_fsg_config_defaults.argtypes = [P(FsgConfig)]
    L.ssb_fsg_build_align.restype = vp
    L.ssb_fsg_build_align.argtypes = [vp, C.c_char_p, P(FsgConfig)]
    L.ssb_fsg_build.restype = vp
    L.ssb_fsg_build.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, P(C.c_char_p), i32, P(FsgConfig)]
    L.ssb_fsg_built_graph.restype = P(FsgGraph)
    L.ssb_fsg_built_graph.argtypes = [vp]
    L.ssb_fsg_built_n_words.argtypes = [vp]
    L.ssb_fsg_built_word.restype = C.c_char_p
    L.ssb_fsg_built_word.argtypes = [vp, i32, P(i32)]
    L.ssb_fsg_built_free.restype = None
    L.ssb_fsg_built_free.argtypes = [vp]
    L.ssb_fe_config_defaults.restype = None
    L.ssb_fe_config_defaults.argtypes = [P(FeConfig)]
    L.ssb_fe_config_from_model.argtypes = [C.c_char_p, P(FeConfig)]
    L.ssb_frontend_create.restype = vp
    L.ssb_frontend_create.argtypes = [P(FeConfig), i32, vp]
    L.ssb_frontend_free.restype = None
    L.ssb_frontend_free.argtypes = [vp]
    L.ssb_frontend_dims.argtypes = [vp, vp]
    L.ssb_frontend_n_frames.restype = i64
    L.ssb_frontend_n_frames.argtypes = [vp, i64]
    L.ssb_frontend_tables.argtypes = [vp] + [vp] * 6
    L.ssb_frontend_run.restype = i64
    L.ssb_frontend_run.argtypes = [vp, vp, i32, vp, i32]
    L.ssb_frontend_download.argtypes = [vp, vp, vp, vp]
    L.ssb_frontend_feat_device.restype = vp
    L.ssb_frontend_feat_device.argtypes = [vp]
    L.ssb_frontend_kernel_ms.argtypes = [vp, vp]
    _lib = L
    return L


def last_error():
    return load().ssb_last_error().decode("utf-8", "replace")


def check(rv, what):
    if rv is None or (isinstance(rv, int) and rv < 0):
        raise SsbError("%s failed: %s" % (what, last_error()))
    return rv
