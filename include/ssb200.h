/*
 * ssb200.h -- C ABI of libssb200.so: the B200 (sm_100a) acoustic-scoring and
 * Viterbi-alignment hot path of SoundSwallower.
 *
 * Plain pointers and sizes only; no torch / C++ types cross this boundary.
 * Every entry point returns 0 (or a valid handle) on success and -1 / NULL on
 * failure, like the reference's C API; ssb_last_error() holds the message the
 * reference would have logged with E_ERROR.  There is NO CPU fallback: without a
 * CUDA device (or without the sm_100a kernels) every compute call fails.
 *
 * `ref:` citations are relative to ReadAlongs/SoundSwallower 0.6.1.
 *
 * Two layers:
 *   1. drop-in objects that keep the reference's static "plugin" layout
 *      (ssb_mgau_t <-> mgau_t / mgaufuncs_t, ref: include/soundswallower/acmod.h:93-111);
 *   2. additive *batched* entry points (many utterances per launch).  The
 *      reference API is one frame at a time (ref: src/decoder.c:935-957) and would
 *      serialise the GPU; the batched calls return exactly what a loop over the
 *      per-utterance reference calls returns.
 */
#ifndef SSB200_H
#define SSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSB_WORST_SCORE ((int32_t)0xE0000000) /* ref: hmm.h:80 */
#define SSB_MAX_TOPN 4                        /* ref: config_defs.h "topn" default */
#define SSB_MAX_FEAT 4
#define SSB_MAX_CB 256 /* ref: ptm_mgau.c:754 */

/* ------------------------------------------------------------------ */
/* library                                                             */
/* ------------------------------------------------------------------ */
int ssb_version(void);
const char *ssb_last_error(void);
/* number of visible CUDA devices with compute capability 10.x; 0 if none */
int ssb_device_count(void);
/* Device blocks released by one call are kept for the next one (per device, at most
 * $SSB_DEV_CACHE_GB, default 64; 0 = off); this returns them to the driver. */
int ssb_device_cache_trim(void);

/* ------------------------------------------------------------------ */
/* model: loaders + packed device image                                */
/* replaces: gauden_init_s3file + gauden_dist_precompute (ref: src/ms_gauden.c:105-258),
 *           read_sendump / read_mixw (ref: src/ptm_mgau.c:456-692),
 *           bin_mdef_read (ref: src/bin_mdef.c:333-520),
 *           tmat_init_s3file (ref: src/tmat.c:125-225),
 *           logmath_init (ref: src/logmath.c:61-163)                  */
/* ------------------------------------------------------------------ */
typedef struct ssb_config_s {
    double logbase;  /* "logbase"   1.0001 */
    float varfloor;  /* "varfloor"  1e-4   */
    double mixwfloor; /* "mixwfloor" 1e-7   */
    double tmatfloor; /* "tmatfloor" 1e-4   */
    int32_t topn;    /* "topn"      4      */
    int32_t ds;      /* "ds"        1      */
    int32_t device;  /* CUDA ordinal; -1 = host arrays only (loader tests) */
    int32_t topn_beam[SSB_MAX_FEAT]; /* "topn_beam" per stream, 0 = off; semi-continuous
                                      * models only (ref: src/s2_semi_mgau.c:184-202, 877-907) */
} ssb_config_t;
void ssb_config_defaults(ssb_config_t *cfg);

/* Which of the reference's scorers the model directory selects, in acmod_load_am's order
 * (ref: src/acmod.c:101-119): ptm_mgau when there is one codebook per CI phone
 * (ref: src/ptm_mgau.c:722-816), s2_semi_mgau when there is a single codebook
 * (ref: src/s2_semi_mgau.c:829-1058), else ms_mgau (ref: src/ms_mgau.c:279-368) -- offered for
 * fully continuous models with one codebook per senone (the ".cont." map of
 * src/ms_senone.c:262-275); other shapes and explicit "senmgau" maps are declined (NULL) so
 * that a caller falls through to the reference's own scorer. */
enum { SSB_SCORER_PTM = 0, SSB_SCORER_SEMI = 1, SSB_SCORER_CONT = 2 };
typedef struct ssb_model_s ssb_model_t;
ssb_model_t *ssb_model_load(const char *hmmdir, const ssb_config_t *cfg);
int ssb_model_kind(const ssb_model_t *m);
/* bin_mdef_ciphone_str (ref: src/bin_mdef.c:590-595); NULL when out of range */
const char *ssb_model_ciphone_str(const ssb_model_t *m, int32_t ci);
void ssb_model_free(ssb_model_t *m);
/* out[0..10] = n_mgau n_feat n_density veclen(stream 0) n_sen n_sseq n_emit n_tmat
 *              n_ciphone n_phone sil ; out[11..14] = featlen[0..3] ; out[15] = sum featlen */
int ssb_model_dims(const ssb_model_t *m, int32_t *out16);
/* host copies of the parsed tables (any pointer may be NULL):
 * mean/var [mgau][feat][density][len]  det [mgau][feat][density]
 * mixw [feat][density][n_sen]  sen2cb [n_sen]  tp [n_tmat][n_emit][n_emit+1]
 * sseq [n_sseq][n_emit]  lut8 [256] */
int ssb_model_copy(const ssb_model_t *m, float *mean, float *var, float *det, uint8_t *mixw,
                   uint8_t *sen2cb, uint8_t *tp, uint16_t *sseq, uint8_t *lut8);
/* per phone id: sequence id, transition matrix id, CI phone (ref: bin_mdef.h:160-175) */
int ssb_model_phones(const ssb_model_t *m, int32_t *ssid, int32_t *tmat, int32_t *ci);

/* ------------------------------------------------------------------ */
/* scorer drop-in: same first members as mgau_t so that                 */
/* acmod->mgau->frame_idx pokes (ref: src/acmod.c:367,748,760) and       */
/* ps_mgau_frame_eval() dispatch (ref: acmod.h:113-114) work unchanged.  */
/* ------------------------------------------------------------------ */
typedef struct ssb_mgau_s ssb_mgau_t;
typedef struct ssb_mgaufuncs_s {
    const char *name;
    int (*frame_eval)(ssb_mgau_t *mgau, int16_t *senscr, uint8_t *senone_active,
                      int32_t n_senone_active, float **feat, int32_t frame, int32_t compallsen);
    int (*transform)(ssb_mgau_t *mgau, void *mllr);
    void (*free)(ssb_mgau_t *mgau);
} ssb_mgaufuncs_t;
struct ssb_mgau_s {
    ssb_mgaufuncs_t *vt;
    int frame_idx;
    /* private state follows */
};
/* replaces ptm_mgau_init (ref: src/ptm_mgau.c:722-816); the model is borrowed */
ssb_mgau_t *ssb_mgau_init(ssb_model_t *m);
/* replaces ptm_mgau_frame_eval (ref: src/ptm_mgau.c:408-454): host buffers in, host
 * senscr[n_sen] out, one frame, history kept on the device */
int ssb_mgau_frame_eval(ssb_mgau_t *mgau, int16_t *senscr, uint8_t *senone_active,
                        int32_t n_senone_active, float **feat, int32_t frame, int32_t compallsen);
/* replaces ptm_mgau_reset_fast_hist (ref: src/ptm_mgau.c:694-720) */
void ssb_mgau_reset(ssb_mgau_t *mgau);
void ssb_mgau_free(ssb_mgau_t *mgau);
/* own != 0: the scorer owns its model, ssb_mgau_free (= vt->free, what acmod_free calls, ref:
 * src/acmod.c:278-279) releases both -- the lifetime ptm_mgau_init's object has in the reference
 * (see integration/ssb_glue.c, the reference-side binding). */
void ssb_mgau_own_model(ssb_mgau_t *mgau, int own);
/* the model a scorer object was made for (the searches of the same decoder need it) */
ssb_model_t *ssb_mgau_model(const ssb_mgau_t *mgau);

/* ------------------------------------------------------------------ */
/* batched path                                                        */
/* ------------------------------------------------------------------ */
/* One batch of independent utterances.  All arrays are host memory.
 * Utterance u owns frames [frame_off[u], frame_off[u+1]) of `feat`
 * ([frame][sum featlen] fp32, the layout of acmod's feat_buf, ref: src/feat.c:369-398)
 * and phones [phone_off[u], phone_off[u+1]) of the chain arrays, which carry what
 * state_align_search_init derives from the alignment (ref: src/state_align_search.c:429-474):
 * ssid/tmat per phone, sf = window start (0 if none), ef = window end (INT32_MAX if none). */
typedef struct ssb_align_in_s {
    int32_t n_utts;
    const float *feat;        /* [frames][sum featlen]; host memory, or device memory of the
                               * model's GPU (e.g. ssb_frontend_feat_device) that is complete
                               * when the call is made / ordered before the batch's stream */
    const int64_t *frame_off;
    const int64_t *phone_off;
    const int32_t *ssid;
    const int32_t *tmat;
    const int32_t *sf;
    const int32_t *ef;
    /* optional [n_utts][(n_sen+31)/32]: senones already flagged active when pass 2
     * starts (the reference never clears them, ref: src/state_align_search.c:186-188) */
    const uint32_t *init_active;
    int32_t compallsen; /* config key "compallsen" */
    /* optional [n_utts][n_mgau*n_feat][4]: the top-N codewords the scorer carries when pass 2
     * starts -- ptm_mgau's lists are not reset between passes, frame 0 copies history slot 1 as
     * the first pass left it (ref: src/ptm_mgau.c:426-440).  They only matter on frames whose
     * integer Gaussian scores tie.  ssb_fsg_out_t.final_topn provides them; NULL = initial lists. */
    const uint8_t *init_topn;
} ssb_align_in_t;

typedef struct ssb_align_out_s {
    /* [total_phones * n_emit]: state start frame, duration, acoustic score
     * (ref: src/state_align_search.c:215-268) */
    int32_t *st_start;
    int32_t *st_dur;
    int32_t *st_score;
    int32_t *utt_rv;     /* [n_utts] 0 ok, -1 "failed to reach final state" */
    int32_t *utt_best;   /* [n_utts] best score of the last frame */
    int32_t *utt_renorm; /* [n_utts] number of renormalisations */
    /* optional debug outputs, NULL to skip */
    int16_t *chain_scr; /* [sum_u T_u * n_states_u] senone score seen by each chain state */
    int32_t *tokens;    /* [sum_u T_u * n_states_u * 2] {history id, score} */
} ssb_align_out_t;

/* Host-only (no device): the data-independent schedule the planner derives for one chain.
 * enter[i] = first frame on which phone i is evaluated (== T: entered at the very end, never
 * evaluated), -1 = never entered.  Phone i is then evaluated on [enter[i], max(enter[i], ef[i])].
 * Restates the hmm_frame() bookkeeping of prune_hmms / phone_transition
 * (ref: src/state_align_search.c:88-133) for window ends that do not decrease along the chain;
 * returns -1 if they do. */
int ssb_plan_chain(int32_t n_phones, int32_t n_frames, const int32_t *sf, const int32_t *ef,
                   int32_t *enter);

typedef struct ssb_batch_s ssb_batch_t;
/* `stream` = a cudaStream_t (0 = the legacy default stream).  All kernels of the
 * batch run on it, so a caller can bracket ssb_batch_run with its own events. */
ssb_batch_t *ssb_batch_create(ssb_model_t *m, void *stream);
void ssb_batch_free(ssb_batch_t *b);
/* plan (active-senone epochs, offsets) + host->device copies */
int ssb_batch_upload(ssb_batch_t *b, const ssb_align_in_t *in);
/* score + align on the device; asynchronous on the batch's stream */
int ssb_batch_run(ssb_batch_t *b);
/* device->host copies of the results; synchronises */
int ssb_batch_download(ssb_batch_t *b, ssb_align_out_t *out);
/* Debug outputs.  By default chain scores and tokens are kept BANDED: every phone keeps only the
 * frames it can be evaluated on (the reference's dense [T][n_states] token stack is its memory
 * hazard -- 86 GB for one hour, SURVEY section 5 -- and 98 % of it is never touched inside word
 * windows), and ssb_batch_download cannot return them.  on & 1: keep the whole dense token stack
 * ({-1,-1} where the reference records nothing); on & 2: keep the dense chain scores.  Call
 * before ssb_batch_upload.  (ssb_align_batch / the pipeline do this themselves when the
 * output struct asks for chain_scr / tokens.) */
int ssb_batch_debug_tokens(ssb_batch_t *b, int on);
/* CUDA-event durations (ms) of the kernels of the last ssb_batch_run:
 * [0] gmm_topn [1] senone_mix [2] chain_viterbi [3] backtrace [4] whole run; synchronises */
int ssb_batch_kernel_ms(ssb_batch_t *b, float *ms8);
/* kernels launched by the last ssb_batch_run */
int ssb_batch_n_launches(const ssb_batch_t *b);
/* counters of the uploaded batch: [0] frames [1] state-frames [2] active senone-frames
 * [3] scanned codebook-frames [4] device bytes held [5] largest active-senone union
 * [6] longest chain (phones) [7] host microseconds the last upload spent planning */
int ssb_batch_stats(const ssb_batch_t *b, int64_t *out8);
/* state-frames the chain Viterbi really evaluates for the uploaded batch: phone i on frames
 * [enter[i], max(enter[i], ef[i])] (the word-window band; ref: src/state_align_search.c:88-133).
 * out8[1] of ssb_batch_stats is the dense count T x states the reference allocates tokens for. */
int64_t ssb_batch_band_state_frames(const ssb_batch_t *b);
/* chain segments the Viterbi pass of the last upload runs on: chains are cut where one word window
 * ends exactly where the next begins (the only frame the chain can be crossed on) and the
 * segments are searched in parallel, with identical results; = n_utts when nothing is cut */
int32_t ssb_batch_n_segments(const ssb_batch_t *b);

/* upload + run + download in one call (the call a host program makes).  A batch of at least
 * two chunks (see ssb_pipeline_create) is routed through a temporary pipeline. */
int ssb_align_batch(ssb_model_t *m, const ssb_align_in_t *in, ssb_align_out_t *out);
/* The same batch over several GPUs of one box (the path shards by utterance, no
 * exchange step): models[d] = the model loaded on device d (ssb_config_t.device); one host
 * thread per GPU takes a contiguous range of utterances with about 1/n of the frames -- slices
 * of the caller's arrays, nothing is re-packed -- through its own pipeline and writes its part
 * of `out`.  No collective; results are those of ssb_align_batch on one device.  The debug
 * outputs (chain_scr, tokens) must be NULL. */
int ssb_align_batch_multi(ssb_model_t *const *models, int32_t n_models, const ssb_align_in_t *in,
                          ssb_align_out_t *out);

/* The same call for callers that keep coming back (a server, a long file list): device
 * buffers are kept between calls and a large batch is cut into chunks of whole utterances
 * (about `chunk_frames` frames each, 0 = 1 024 000 or $SSB_PIPE_CHUNK_FRAMES) that travel through
 * `n_lanes` (0 = 4 or $SSB_PIPE_LANES) batches, each on its own stream and host thread, so that
 * planning, host->device and device->host copies of one chunk overlap the kernels of another.
 * Utterances are independent (ref: one decoder_t per utterance, src/decoder.c:737-798), so the
 * results are those of ssb_batch_upload/run/download on the whole batch, utterance by
 * utterance.  `in->feat` should be pinned host memory for the copies to overlap. */
typedef struct ssb_pipeline_s ssb_pipeline_t;
ssb_pipeline_t *ssb_pipeline_create(ssb_model_t *m, int32_t n_lanes, int64_t chunk_frames);
int ssb_pipeline_align(ssb_pipeline_t *p, const ssb_align_in_t *in, ssb_align_out_t *out);
/* The two halves of ssb_pipeline_align, for a stream of batches: submit returns a ticket (< 0 on
 * error) at once, collect waits for that batch.  The arrays behind `in` and `out` must stay
 * valid until the batch is collected; batches are processed in submission order and several
 * may be in flight.  With ssb_pipeline_set_overlap(p, 0) the kernels of different chunks never
 * share the GPU (only planning and copies overlap them) -- the setting for whole batches as
 * chunks (chunk_frames >= the batch), where batch i+1 is planned and uploaded while batch i
 * computes; the default, 1, lets kernels of different chunks overlap (chunks of one batch). */
int64_t ssb_pipeline_submit(ssb_pipeline_t *p, const ssb_align_in_t *in, ssb_align_out_t *out);
int ssb_pipeline_collect(ssb_pipeline_t *p, int64_t ticket);
int ssb_pipeline_set_overlap(ssb_pipeline_t *p, int32_t overlap_kernels);
int ssb_pipeline_n_launches(const ssb_pipeline_t *p); /* kernels launched for the last collected batch */
int ssb_pipeline_n_chunks(const ssb_pipeline_t *p);   /* chunks of the last collected batch */
/* timeline of the last call, 8 doubles per chunk: lane, first utterance, ms since the call
 * started at which the chunk's upload began / its upload returned / its download returned,
 * CUDA-event ms of its top-N kernel and of all its kernels, 0; returns chunks written */
int ssb_pipeline_trace(const ssb_pipeline_t *p, double *out, int32_t max_chunks);
void ssb_pipeline_free(ssb_pipeline_t *p);

/* Dense senone scoring of whole utterances with "compallsen" semantics
 * (what acmod_score returns frame by frame, ref: src/acmod.c:822-860):
 * senscr [total_frames][n_sen] int16, 0 = best.  senscr may be NULL to keep the
 * result on the device only (throughput measurement); returns frames scored. */
int64_t ssb_score_batch(ssb_model_t *m, const float *feat, const int64_t *frame_off,
                        int32_t n_utts, int16_t *senscr);
/* raw top-N of every (frame, codebook, stream), before normalisation:
 * cw [frames][mgau][feat][topn] u8, score [..] int32 (ref: src/ptm_mgau.c:86-225) */
int64_t ssb_topn_batch(ssb_model_t *m, const float *feat, const int64_t *frame_off,
                       int32_t n_utts, uint8_t *cw, int32_t *score);

/* Verification hook of the tensor-core scorer (gmm_topn_tc.cu): ssb_topn_batch plus the TF32
 * screening scores approx [frames][mgau][feat][n_density], their per-row error bound
 * eps [frames][mgau][feat] (|approx - exact fp32 distance| <= eps is what makes the screening
 * safe) and counters[4] {exact evaluations of scan survivors, scanned (utterance, frame,
 * codebook-stream) steps, steps that took the literal slow path, 0}.  Any of approx/eps/counters may be NULL, not all. */
int64_t ssb_tc_probe(ssb_model_t *m, const float *feat, const int64_t *frame_off, int32_t n_utts,
                     uint8_t *cw, int32_t *score, float *approx, float *eps, int64_t *counters);
/* eps above is [frames][mgau][feat][2]: the bound of the regular densities and the bound of
 * the "hot" ones (outlier precisions, e.g. floored variances).  out[mgau*feat][4]: bit n of
 * the 128-bit mask says density n of that codebook-stream is hot. */
int ssb_tc_hot_mask(const ssb_model_t *m, uint32_t *out);

/* ------------------------------------------------------------------ */
/* FSG token-passing search (first pass / grammar decoding)            */
/* replaces fsg_search_start/step/finish, fsg_history_*, and the backtrace behind
 * fsg_search_hyp / fsg_search_seg_iter (ref: src/fsg_search.c:309-1142, src/fsg_history.c:129-232).
 * The graph is prepared on the host by the caller: it is the reference's fsg_model_t +
 * fsg_lextree_t flattened (INTEGRATION.md shows the loop over alloc_head / fsg_model_arcs).   */
/* ------------------------------------------------------------------ */
typedef struct ssb_fsg_graph_s {
    int32_t n_state, start, final, n_link, n_pnode, n_ciphone, sil;
    int32_t beam, pbeam, wbeam, maxhmmpf; /* log-domain beams of fsg_search_init; -1 = no maxhmmpf */
    const int32_t *link4;     /* [n_link][4] from_state to_state logs2prob wid (-1: null arc); the
                                 links of state s are arc_off[s]..arc_off[s+1], in fsg_model_arcs order */
    const uint8_t *link_flag; /* [n_link] bit0: word gives no right context (filler or single phone) */
    const int32_t *arc_off;   /* [n_state+1] */
    const int32_t *root;      /* [n_state] first root pnode of the state's lextree, -1 if none */
    const int32_t *pnode8;    /* [n_pnode][8] ssid tmatid logs2prob ci_ext leaf (succ | link) sibling ppos */
    const uint32_t *ctxt;     /* [n_pnode][4] fsg_pnode_ctxt_t */
} ssb_fsg_graph_t;

typedef struct ssb_fsg_in_s {
    int32_t n_utts;
    const float *feat;        /* [frames][sum featlen], utterance u = frame_off[u]..frame_off[u+1] */
    const int64_t *frame_off;
    int32_t n_graphs;
    const ssb_fsg_graph_t *graphs;
    const int32_t *utt_graph; /* [n_utts] graph searched for each utterance */
    int32_t hist_cap;         /* history entries kept per utterance (overflow: utt_rv = -2) */
    int32_t max_seg;          /* segmentation entries returned per utterance */
    /* config key "compallsen" negated: 0 = every senone is scored on every frame (compallsen =
     * yes), 1 = the reference's default: only the senones of the active HMMs are scored, frame
     * by frame inside the search (fsg_search_sen_active + acmod_score with an active list, ref:
     * src/fsg_search.c:309-328, 686-690) -- path scores then equal the default CLI's.  PTM
     * models with 128 densities only. */
    int32_t active_lists;
    /* 1 = the utterances are still running: the hypothesis is the best word exit of the last
     * frame that has one, whatever state it leads to (fsg_search_hyp before
     * search_module_finish: find_exit with final = FALSE, ref: src/fsg_search.c:853-924) */
    int32_t partial;
} ssb_fsg_in_t;

typedef struct ssb_fsg_out_s {
    int32_t *segs;       /* [n_utts][max_seg][5] link sf ef ascr lscr (fsg_seg_bp2itor) */
    int32_t *n_seg;      /* [n_utts] entries used; < 0: -needed when max_seg is too small */
    int32_t *hyp_score;  /* [n_utts] score of the best final exit (fsg_search_hyp) */
    int32_t *exit_bp;    /* [n_utts] its history index; -1: "does not match the grammar", 0: no hypothesis */
    int32_t *utt_rv;     /* [n_utts] 0, or -2 on history overflow */
    int32_t *n_hist;     /* [n_utts] history entries created (optional) */
    int64_t *n_hmm_eval; /* [n_utts] HMM evaluations (optional) */
    int32_t *hist9;      /* optional [n_utts][hist_cap][9] link score pred frame lc rc[4] */
    float *kernel_ms;    /* optional [4] gmm_topn, senone_mix, fsg_search, backtrace */
    int32_t n_launches;  /* kernels launched (written by the call) */
    /* active_lists = 1 only, optional: acmod's active-senone flags as the last frame left them,
     * [n_utts][(n_sen+31)/32] -- hand them to ssb_align_in_t.init_active for the second pass,
     * which never clears them (ref: src/state_align_search.c:186-188); and the number of
     * senones evaluated per utterance (fsgs->n_sen_eval) */
    uint32_t *final_active;
    int64_t *n_sen_eval;
    /* optional (PTM / semi-continuous models): [n_utts][n_mgau*n_feat][4], the top-N codewords
     * the scorer is left carrying (history slot 1 = after the last odd frame): the second pass'
     * ssb_align_in_t.init_topn */
    uint8_t *final_topn;
} ssb_fsg_out_t;
/* Senone scoring + search + backtrace of a batch of utterances, with dense ("compallsen")
 * scores or, with in->active_lists, the reference's default active-list scoring. */
int ssb_fsg_batch(ssb_model_t *m, const ssb_fsg_in_t *in, ssb_fsg_out_t *out);
/* 1 when in->active_lists = 1 is available for this model (PTM, 128 densities, on a device) */
int ssb_model_fsg_active_ok(const ssb_model_t *m);

/* single HMM step on the device (ref: src/hmm.c:482-567); st = score[5] hist[5]
 * out_score out_hist, updated in place; returns best score via *best */
int ssb_hmm_vit_eval(ssb_model_t *m, int32_t n_emit, int32_t tmatid, const uint16_t *senid,
                     const int16_t *senscr, int32_t *st12, int32_t *best);

/* n_cases independent steps on caller-provided transition matrices: tp [n][n_emit][n_emit+1]
 * (255 = impossible), senscr [n][n_emit] (the scores of the HMM's own states), st12 [n][12]
 * updated in place, best [n].  n_emit 3 or 5 (hmm_vit_eval_3st_lr / _5st_lr, ref:
 * src/hmm.c:166-304, 482-567) whatever the model has. */
int ssb_hmm_vit_eval_tp(ssb_model_t *m, int32_t n_emit, int32_t n_cases, const uint8_t *tp,
                        const int16_t *senscr, int32_t *st12, int32_t *best);

/* ------------------------------------------------------------------ lexicon (host only)
 * Graph preparation for the chain aligner: pronunciation dictionary + context-dependent phone
 * lookup + the word -> phone-chain expansion.  Word ids are the reference's (main dictionary
 * in file order, then the filler dictionary, then <s> </s> <sil> when missing). */
typedef struct ssb_lexicon_s ssb_lexicon_t;
/* replaces dict_init (ref: src/dict.c:134-366) + dict2pid_build (ref: src/dict2pid.c:372-480);
 * either path may be NULL.  The model is borrowed. */
ssb_lexicon_t *ssb_lexicon_load(const ssb_model_t *m, const char *dictfile, const char *fdictfile);
void ssb_lexicon_free(ssb_lexicon_t *lx);
int32_t ssb_lexicon_size(const ssb_lexicon_t *lx);
int32_t ssb_lexicon_wordid(const ssb_lexicon_t *lx, const char *word);   /* dict_wordid; -1 */
const char *ssb_lexicon_wordstr(const ssb_lexicon_t *lx, int32_t wid);
/* CI phones of a word; returns the pronunciation length */
int32_t ssb_lexicon_pron(const ssb_lexicon_t *lx, int32_t wid, int32_t *ciphones, int32_t max);
int32_t ssb_lexicon_is_filler(const ssb_lexicon_t *lx, int32_t wid);     /* dict_filler_word */
int32_t ssb_lexicon_basewid(const ssb_lexicon_t *lx, int32_t wid);       /* dict_basewid */
/* replaces alignment_populate (ref: src/ps_alignment.c:133-248): per phone of the word
 * sequence its senone-sequence id, transition matrix, CI phone and parent word index (any
 * output may be NULL).  Returns the number of phones, -1 on error. */
int32_t ssb_chain_populate(const ssb_lexicon_t *lx, const int32_t *wids, int32_t n_words,
                           int32_t *ssid, int32_t *tmat, int32_t *cipid, int32_t *parent,
                           int32_t max_phones);

/* Alignment grammar + lextree on the host: what decoder_set_align_text builds
 * (ref: src/decoder.c:685-735: one state per word boundary, one link per word), augmented as
 * fsg_search_init does (silence / filler self-loops on every state, alternate pronunciations;
 * ref: src/fsg_search.c:83-168, 171-260) and expanded by fsg_lextree_init
 * (ref: src/fsg_lextree.c:226-716), flattened to ssb_fsg_graph_t with the reference's link
 * and node order. */
typedef struct ssb_fsg_config_s {
    double beam, pbeam, wbeam;        /* "beam" 1e-48, "pbeam" 1e-48, "wbeam" 7e-29 */
    float lw, wip, pip;               /* "lw" 6.5, "wip" 0.65, "pip" 1.0 */
    float silprob, fillprob;          /* "silprob" 0.005, "fillprob" 1e-8 */
    int32_t maxhmmpf;                 /* "maxhmmpf" 30000 */
    int32_t fsgusefiller, fsgusealtpron; /* yes, yes */
} ssb_fsg_config_t;
void ssb_fsg_config_defaults(ssb_fsg_config_t *c);
typedef struct ssb_fsg_built_s ssb_fsg_built_t;
/* NULL with "Unknown word ..." when a word of `text` is not in the dictionary */
ssb_fsg_built_t *ssb_fsg_build_align(const ssb_lexicon_t *lx, const char *text,
                                     const ssb_fsg_config_t *cfg);
/* Any grammar from its transition list (a compiled JSGF, a .fsg file's TRANSITION lines): what
 * fsg_model_read_s3file / jsgf_build_fsg hand to decoder_set_fsg (ref: src/fsg_model.c:506-690,
 * :62-140, :146-213; src/decoder.c:600-645), then the same augmentation + lextree as above.
 * Transitions in the order the reference would add them (link order decides ties in the search);
 * word[i] NULL or "" = null transition; prob[i] in (0, 1] is converted as the reference does,
 * (int32)(logmath_log(p) * lw); null_closure != 0 computes the transitive closure of the null
 * transitions first (fsg_model_null_trans_closure), as both of the reference's producers do. */
ssb_fsg_built_t *ssb_fsg_build(const ssb_lexicon_t *lx, int32_t n_state, int32_t start, int32_t final,
                               int32_t n_trans, const int32_t *from, const int32_t *to,
                               const float *prob, const char *const *word, int32_t null_closure,
                               const ssb_fsg_config_t *cfg);
/* the same with fsg_link_t.logs2prob values (integers already scaled by lw) instead of linear
 * probabilities: what an fsg_model_t in memory holds (integration/ssb_glue.c) */
ssb_fsg_built_t *ssb_fsg_build_logp(const ssb_lexicon_t *lx, int32_t n_state, int32_t start, int32_t final,
                                    int32_t n_trans, const int32_t *from, const int32_t *to,
                                    const int32_t *logs2prob, const char *const *word,
                                    int32_t null_closure, const ssb_fsg_config_t *cfg);
/* the graph (arrays owned by the object) and its vocabulary: link4[.][3] indexes these words */
const ssb_fsg_graph_t *ssb_fsg_built_graph(const ssb_fsg_built_t *b);
int32_t ssb_fsg_built_n_words(const ssb_fsg_built_t *b);
const char *ssb_fsg_built_word(const ssb_fsg_built_t *b, int32_t fsg_wid, int32_t *dict_wid);
int32_t ssb_fsg_built_is_filler(const ssb_fsg_built_t *b, int32_t fsg_wid); /* fsg_model_is_filler */
void ssb_fsg_built_free(ssb_fsg_built_t *b);

/* ------------------------------------------------------------------ search-module drop-in
 * Objects with the layout and the vtable of the reference's search_module_t / seg_iter_t
 * (ref: include/soundswallower/search_module.h:72-113, 157-174) for the two searches on the
 * path: state_align_search (ref: src/state_align_search.c:46-474) and fsg_search
 * (ref: src/fsg_search.c:171-260, 664-851, 945-1142).  search_module_forward's loop
 * (ref: src/decoder.c:935-957) and decoder_end_utt / decoder_hyp / decoder_seg_iter work on
 * them unchanged through the macros of search_module.h.  step() collects the frame's feature
 * vector; finish() runs the utterance through the batched kernels; hyp / seg_iter / the
 * alignment entries are then served from the results.  hyp / seg_iter of the grammar search
 * BETWEEN two steps (the reference answers from its history table as it stands, find_exit with
 * final = FALSE, ref: src/fsg_search.c:853-960) search the frames collected so far and give what
 * the reference gives at that frame; the aligner's hyp between steps reads the alignment as
 * populated, like the reference's (ref: src/state_align_search.c:365-411). */
typedef struct ssb_search_s ssb_search_t;
typedef struct ssb_seg_iter_s ssb_seg_iter_t;
typedef struct ssb_searchfuncs_s { /* searchfuncs_t, ref: search_module.h:72-84 */
    int (*start)(ssb_search_t *search);
    int (*step)(ssb_search_t *search, int frame_idx); /* aligner: 0, grammar search: 1, error < 0 */
    int (*finish)(ssb_search_t *search);              /* aligner: -1 "Failed to reach final state" */
    int (*reinit)(ssb_search_t *search, void *dict, void *d2p);
    void (*free)(ssb_search_t *search);
    void *(*lattice)(ssb_search_t *search);           /* aligner: NULL pointer; grammar search: returns NULL */
    const char *(*hyp)(ssb_search_t *search, int32_t *out_score);
    int32_t (*prob)(ssb_search_t *search);            /* aligner: NULL pointer; grammar search: 0 (no bestpath) */
    ssb_seg_iter_t *(*seg_iter)(ssb_search_t *search);
} ssb_searchfuncs_t;
struct ssb_search_s { /* search_module_t field for field, ref: search_module.h:89-113 */
    ssb_searchfuncs_t *vt;
    char *type; /* "state_align" | "fsg" */
    char *name;
    void *config;
    void *acmod; /* handed back to the feature source */
    void *dict;  /* the ssb_lexicon_t */
    void *d2p;
    char *hyp_str;
    void *dag;
    void *last_link;
    int32_t post;
    int32_t n_words;
    int32_t start_wid, silence_wid, finish_wid;
};
typedef struct ssb_segfuncs_s { /* ps_segfuncs_t, ref: search_module.h:157-160 */
    ssb_seg_iter_t *(*seg_next)(ssb_seg_iter_t *seg); /* frees the iterator and returns NULL at the end */
    void (*seg_free)(ssb_seg_iter_t *seg);
} ssb_segfuncs_t;
struct ssb_seg_iter_s { /* seg_iter_t, ref: search_module.h:165-174 */
    ssb_segfuncs_t *vt;
    ssb_search_t *search;
    const char *word;
    int32_t sf, ef;
    int32_t ascr, lscr, prob;
};
/* The frame's feature vector ([sum featlen] fp32, what acmod_score scores for frame_idx,
 * ref: src/acmod.c:822-860); NULL when the frame is not available.  With a NULL source the
 * search reads the frames given to ssb_search_feed. */
typedef const float *(*ssb_feat_source_fn)(void *acmod, int frame_idx);
/* replaces state_align_search_init (ref: src/state_align_search.c:429-474) for the alignment
 * whose word level is (wids, wstart, wdur) -- pass 1's words with their frame windows, 0/0 =
 * no window; the phone and state levels are populated as alignment_populate does. */
ssb_search_t *ssb_state_align_search_init(const char *name, ssb_model_t *m, const ssb_lexicon_t *lx,
                                          const int32_t *wids, const int32_t *wstart,
                                          const int32_t *wdur, int32_t n_words,
                                          ssb_feat_source_fn src, void *acmod);
/* replaces fsg_search_init (ref: src/fsg_search.c:171-260); consumes `fsg` like the reference */
ssb_search_t *ssb_fsg_search_init(const char *name, ssb_model_t *m, const ssb_lexicon_t *lx,
                                  ssb_fsg_built_t *fsg, ssb_feat_source_fn src, void *acmod);
/* appends frames to the search's own feature buffer (the NULL-source case); returns the
 * number of frames held */
int ssb_search_feed(ssb_search_t *search, const float *feat, int32_t n_frames);
/* acmod's active-senone flags ((n_sen+31)/32 words): after finish() of a grammar search in the
 * default mode, what its last frame left (ssb_search_final_active; all zero in compallsen mode);
 * before start() of the aligner, what it starts from (ssb_search_set_init_active) -- in the
 * reference both live in the one acmod the two searches share, and the aligner never clears
 * them (ref: src/state_align_search.c:186-188). */
int ssb_search_final_active(const ssb_search_t *search, uint32_t *bits);
int ssb_search_set_init_active(ssb_search_t *search, const uint32_t *bits);
/* likewise the scorer's carried top-N codewords ([n_mgau*n_feat][4] bytes; returns that row
 * count): ptm_mgau's lists are not reset between the passes either */
int ssb_search_final_topn(const ssb_search_t *search, uint8_t *cw);
int ssb_search_set_init_topn(ssb_search_t *search, const uint8_t *cw);
/* alignment_words / alignment_phones / alignment_states of the aligner's alignment
 * (level 0 / 1 / 2, ref: src/ps_alignment.c:357-420): [n][5] = id (word id | CI phone |
 * senone), start, duration, score, parent; returns the number of entries at that level */
int32_t ssb_search_alignment(const ssb_search_t *search, int32_t level, int32_t *out5,
                             int32_t max_entries);

/* ------------------------------------------------------------------ batched two-pass alignment
 * `soundswallower --align` for a batch: decoder_set_align_text + search_module_forward (pass 1,
 * the reference's default active-list mode where the model allows it), decoder_alignment
 * (pass 2, starting from the acmod flags pass 1 left), alignment_propagate and
 * decoder_result_json (ref: src/decoder.c:685-798, 935-957, 1339-1593; src/ps_alignment.c:
 * 317-355).  `feat` is host or device memory (ssb_frontend_feat_device), `texts[u]` the
 * whitespace-separated transcript of utterance u, `cfg` NULL for the reference's defaults,
 * align_level 0 = first pass only, >= 1 = both passes, frate 0 = 100 frames/s.
 * NULL with "Unknown word ..." when a transcript has a word the dictionary lacks. */
typedef struct ssb_text_align_s ssb_text_align_t;
ssb_text_align_t *ssb_align_texts(ssb_model_t *m, const ssb_lexicon_t *lx, const float *feat,
                                  const int64_t *frame_off, const char *const *texts, int32_t n_utts,
                                  const ssb_fsg_config_t *cfg, int32_t align_level, int32_t frate);
void ssb_text_align_free(ssb_text_align_t *r);
/* 0 ok, -1 no hypothesis ("Final result does not match the grammar"), -2 "Failed to reach final
 * state in alignment"; hyp_score = decoder_hyp's score, n_frames = decoder_n_frames */
int32_t ssb_text_align_status(const ssb_text_align_t *r, int32_t u, int32_t *hyp_score, int32_t *n_frames);
const char *ssb_text_align_hyp(const ssb_text_align_t *r, int32_t u);   /* decoder_hyp; NULL if none */
/* level 0 / 1 / 2 = alignment_words / phones / states: [n][5] id start duration score parent;
 * level 3 = the first pass' seg_iter: [n][5] fsg word id (-1 null) sf ef ascr lscr.  Returns n. */
int32_t ssb_text_align_entries(const ssb_text_align_t *r, int32_t u, int32_t level, int32_t *out5,
                               int32_t max_entries);
/* decoder_result_json(d, start, align_level) of utterance u: the line the reference CLI prints;
 * owned by the result object, valid until the next call for the same utterance; NULL where the
 * reference returns NULL */
const char *ssb_text_align_json(ssb_text_align_t *r, int32_t u, double start, int32_t align_level);
/* renders the JSON line of every utterance with up to 16 host threads; ssb_text_align_json
 * with the same (start, align_level) then returns the stored lines */
int ssb_text_align_render(ssb_text_align_t *r, double start, int32_t align_level);
/* ms8: [0..3] kernels of pass 1 (top-N, mix, search, backtrace); wall clock of [4] grammars +
 * first pass, [5] chains on the host, [6] second pass + propagate, [7] the whole call */
int ssb_text_align_kernel_ms(const ssb_text_align_t *r, float *ms8);

/* ------------------------------------------------------------------ frontend
 * Batched PCM -> MFCC -> CMN -> dynamic features for whole utterances: what
 * acmod_process_raw(full_utt=TRUE) obtains from fe_start / fe_process_int16 |
 * fe_process_float32 / fe_end (ref: src/fe_interface.c:352-360, 578-713,
 * src/fe_sigproc.c:238-738, src/fe_noise.c:266-327) followed by
 * feat_s2mfc2feat_block_utt (ref: src/feat.c:978-1007, 589-632; src/cmn.c:159-229).
 * Arithmetic types are the reference's: float64 up to the log mel spectrum, float32
 * cepstra.  Everything except the natural logarithm is evaluated in the reference's
 * operation order with IEEE operations. */
enum { SSB_FE_DCT = 0, SSB_FE_LEGACY = 1, SSB_FE_HTK = 2 };   /* "transform" */
enum { SSB_FE_CMN_NONE = 0, SSB_FE_CMN_BATCH = 1 };          /* "cmn": none | batch/current */
enum { SSB_PCM_INT16 = 0, SSB_PCM_FLOAT32 = 1 };
typedef struct ssb_fe_config_s {
    int32_t samprate;      /* "samprate"      16000 */
    int32_t frate;         /* "frate"         100 */
    int32_t ncep;          /* "ncep"          13 */
    int32_t nfft;          /* "nfft"          0 = smallest power of two >= window */
    int32_t nfilt;         /* "nfilt"         40 */
    int32_t lifter;        /* "lifter"        0 */
    int32_t remove_dc;     /* "remove_dc"     no */
    int32_t remove_noise;  /* "remove_noise"  no */
    int32_t unit_area;     /* "unit_area"     yes */
    int32_t round_filters; /* "round_filters" yes */
    int32_t doublebw;      /* "doublebw"      no */
    int32_t transform;     /* "transform"     legacy */
    int32_t cmn;           /* "cmn"           batch (live CMN is a streaming mode: not offered) */
    int32_t varnorm;       /* "varnorm"       no */
    float wlen;            /* "wlen"          0.025625 */
    float alpha;           /* "alpha"         0.97 */
    float lowerf;          /* "lowerf"        133.33334 */
    float upperf;          /* "upperf"        6855.4976 */
} ssb_fe_config_t;
/* defaults of include/soundswallower/config_defs.h:296-449 (non-web build) */
void ssb_fe_config_defaults(ssb_fe_config_t *c);
/* defaults overridden by <hmmdir>/feat_params.json, as decoder_init expands "hmm"
 * (ref: src/decoder.c:100-160).  -1 when the file asks for something this frontend does
 * not implement (feat other than 1s_c_d_dd, an svspec that is not three equal contiguous
 * streams, lda, agc, dither, frequency warping, logspec/smoothspec, live CMN). */
int ssb_fe_config_from_model(const char *hmmdir, ssb_fe_config_t *c);

typedef struct ssb_frontend_s ssb_frontend_t;
/* `stream` as in ssb_batch_create.  NULL (see ssb_last_error) for parameters fe_init
 * refuses (ref: src/fe_interface.c:270-300). */
ssb_frontend_t *ssb_frontend_create(const ssb_fe_config_t *c, int device, void *stream);
void ssb_frontend_free(ssb_frontend_t *fe);
/* out8: frame_size frame_shift fft_size nfilt ncep feat_dim(=3*ncep) n_filter_coeffs 0 */
int ssb_frontend_dims(const ssb_frontend_t *fe, int32_t *out8);
/* frames a whole-utterance call yields for n_samples samples (fe_process_int16 with a NULL
 * output buffer, ref: src/fe_interface.c:379-391, minus the frame fe_end cannot fill when
 * there is no sample at all) */
int64_t ssb_frontend_n_frames(const ssb_frontend_t *fe, int64_t n_samples);
/* host tables (parity tests): mel filters, DCT basis, lifter, half Hamming window */
int ssb_frontend_tables(const ssb_frontend_t *fe, int32_t *spec_start, int32_t *filt_width,
                        float *coeffs, float *mel_cosine, float *lifter, double *hamming);
/* Upload the samples of n_utts utterances (utterance u = samp_off[u]..samp_off[u+1], int16
 * or float32 in [-1,1) as for fe_process_float32) and compute their features on the device;
 * asynchronous on the frontend's stream.  Returns the total number of frames. */
int64_t ssb_frontend_run(ssb_frontend_t *fe, const void *pcm, int32_t encoding,
                         const int64_t *samp_off, int32_t n_utts);
/* frame_off [n_utts+1]; mfcc [frames][ncep] (before CMN) and feat [frames][3*ncep], either
 * may be NULL; synchronises */
int ssb_frontend_download(ssb_frontend_t *fe, int64_t *frame_off, float *mfcc, float *feat);
/* device pointer to feat [frames][3*ncep] of the last run, valid until the next run: may be
 * passed as `feat` of ssb_align_in_t / ssb_fsg_in_t / ssb_score_batch (those accept host or
 * device memory) so that features never visit the host.  A frontend created on the default
 * stream leaves its work queued there (the consumers wait for that stream); one created on a
 * stream of its own is synchronised by this call. */
const float *ssb_frontend_feat_device(const ssb_frontend_t *fe);
/* CUDA-event durations (ms) of the last run: [0] mel spectrum [1] noise tracker
 * [2] cepstrum [3] CMN sums [4] dynamic features [5] whole run incl. the upload */
int ssb_frontend_kernel_ms(ssb_frontend_t *fe, float *ms8);

#ifdef __cplusplus
}
#endif
#endif /* SSB200_H */
