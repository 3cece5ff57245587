/*
 * ss_oracle.h -- CPU oracle: a plain-C restatement of the reference's
 * acoustic-scoring + Viterbi-alignment hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load liboracle.so; the product
 * (soundswallower_b200/) never links, imports or executes anything in oracle/.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle_*.py)
 * against the unmodified reference compiled as oracle/_ref/libssref.so and
 * against golden vectors generated from it (tests/golden/, tools/make_golden.py),
 * including the SURVEY.md Appendix-B goldens (hyp score -2761, the sha256 of the
 * 278x5126 senone-score matrix, word/phone/state segmentations).
 *
 * All `ref:` citations are relative to ReadAlongs/SoundSwallower 0.6.1.
 */
#ifndef SS_ORACLE_H
#define SS_ORACLE_H
#include <stdint.h>

#define ORC_WORST_SCORE ((int32_t)0xE0000000) /* ref: include/soundswallower/hmm.h:80 */
#define ORC_TMAT_WORST (-255)                 /* ref: hmm.h:86 */
#define ORC_SENSCR_SHIFT 10                   /* ref: hmm.h:69 */
#define ORC_MAX_NEG_ASCR 96                   /* ref: tied_mgau_common.h:81 */
#define ORC_MAX_NEG_MIXW 159                  /* ref: tied_mgau_common.h:80 */
#define ORC_WORST_DIST INT32_MIN              /* ref: tied_mgau_common.h:60 */
#define ORC_MAX_FEAT 8
#define ORC_MAX_TOPN 8

typedef struct orc_model_s {
    /* gauden (ref: ms_gauden.h:83-92), flattened */
    int32_t n_mgau, n_feat, n_density;
    int32_t featlen[ORC_MAX_FEAT];
    int32_t featoff[ORC_MAX_FEAT + 1]; /* offsets of each stream in a frame */
    int32_t blk;                       /* sum of featlen */
    float *mean;                       /* [mgau][feat][density][featlen[f]] (file order) */
    float *var;                        /* same, after precompute */
    float *det;                        /* [mgau][feat][density] */
    int64_t *gau_off;                  /* [mgau][feat] offset into mean/var */
    /* senones */
    int32_t n_sen;
    uint8_t *mixw;   /* [feat][density][n_sen], 4-bit clusters expanded */
    uint8_t *sen2cb; /* [n_sen] */
    /* mdef */
    int32_t n_ciphone, n_phone, n_emit, n_ci_sen, n_tmat_mdef, n_sseq, n_ctx, n_cd_tree, sil;
    uint16_t *sseq;   /* [n_sseq][n_emit] */
    int32_t *ph_ssid; /* [n_phone] */
    int32_t *ph_tmat;
    int32_t *ph_ci;
    char **ciname;
    /* tmat */
    int32_t n_tmat, n_state;
    uint8_t *tp; /* [n_tmat][n_state][n_state+1] */
    /* log math */
    double logbase;
    uint8_t lut8[256];
    int32_t lmath_zero; /* shift-0 zero */
    /* which scorer acmod_load_am ends up with (ref: acmod.c:101-119): ptm_mgau when
     * n_mgau == n_ciphone, s2_semi_mgau when there is a single codebook */
    int32_t kind;
    int32_t topn_beam[ORC_MAX_FEAT]; /* s2_semi "topn_beam" per stream, 0 = off */
} orc_model_t;
enum { ORC_KIND_PTM = 0, ORC_KIND_SEMI = 1, ORC_KIND_CONT = 2 };

/* ---- load-time (ref: logmath.c, ms_gauden.c, ptm_mgau.c read_sendump, bin_mdef.c, tmat.c) */
int orc_logadd_table8(double base, int shift, uint8_t *out256);
int32_t orc_logmath_log(double base, int shift, double p);
orc_model_t *orc_model_load(const char *dir, double logbase, float varfloor, double tmatfloor);
void orc_model_free(orc_model_t *m);
int orc_model_dims(const orc_model_t *m, int32_t *out);
int orc_model_copy(const orc_model_t *m, float *mean, float *var, float *det, uint8_t *mixw,
                   uint8_t *sen2cb, uint8_t *tp, uint16_t *sseq, uint8_t *lut8);
int orc_phone_table(const orc_model_t *m, int32_t *ssid, int32_t *tmat, int32_t *ci);

/* ---- PTM scorer (ref: ptm_mgau.c) */
typedef struct orc_ptm_s orc_ptm_t;
orc_ptm_t *orc_ptm_new(const orc_model_t *m, int topn, int ds_ratio);
void orc_ptm_free(orc_ptm_t *p);
void orc_ptm_reset(orc_ptm_t *p);
void orc_ptm_set_frame_idx(orc_ptm_t *p, int frame_idx);
/* s2_semi only: per-stream "topn_beam" (ref: s2_semi_mgau.c:184-202, 877-907; missing
 * entries repeat the largest given one like split_topn does); applies to scorers created
 * afterwards */
void orc_model_set_topn_beam(orc_model_t *m, const int32_t *beam, int n);
/* One frame_eval; `feat` = blk floats. active = delta list (ref acmod.c:947). If
 * topn_out != NULL receives post-norm [mgau][feat][topn][2] = (cw, score). */
int orc_ptm_frame_eval(orc_ptm_t *p, int16_t *senscr, const uint8_t *active, int32_t n_active,
                       const float *feat, int32_t frame, int32_t compallsen, int32_t *topn_out);
/* Whole-utterance compallsen scoring: out[T][n_sen]. Fresh history. */
int orc_ptm_score_all(const orc_model_t *m, int topn, const float *feat, int T, int16_t *out);
/* Raw (pre-norm) top-N for every frame, fresh history per call:
 * cw[T][mgau][feat][topn] (u8), score[T][mgau][feat][topn] (i32) */
int orc_ptm_topn_all(const orc_model_t *m, int topn, const float *feat, int T, uint8_t *cw,
                     int32_t *score);

/* ---- continuous scorer (ss_oracle_cont.c; ref: ms_mgau.c:279-368, ms_gauden.c:342-457,
 * ms_senone.c:315-362).  mixw is [sen][feat][density] for such models.  In active-list mode
 * only the listed entries of senscr are written, like the reference. */
int orc_cont_frame_eval(const orc_model_t *m, int topn, int16_t *senscr, const uint8_t *active,
                        int32_t n_active, const float *feat, int32_t compallsen);

/* ---- active list (ref: acmod.c:947-999) */
int orc_flags2list(const uint32_t *bits, int n_sen, uint8_t *out);

/* ---- HMM (ref: hmm.c:166-304, 482-567). st = score[5] hist[5] out_score out_hist */
int32_t orc_hmm_eval(int n_emit, const uint8_t *tp, const uint16_t *senid, const int16_t *senscr,
                     int32_t *st);

/* ---- chain aligner (ref: state_align_search.c) */
typedef struct orc_align_out_s {
    int32_t rv;         /* 0 ok, -1 failed */
    int32_t best_score; /* last frame's best */
    int32_t n_renorm;
} orc_align_out_t;
/* Dense-score variant: senscr[T][n_sen] supplied ("given identical senone scores").
 * phones: ssid[n], tmat[n], sf[n], ef[n] (ref state_align_search_init semantics already
 * applied: sf=0 / ef=INT_MAX when unconstrained).
 * Outputs: st_start/st_dur/st_score [n*n_emit]; tokens (optional) [T][n*n_emit][2]. */
int orc_state_align_dense(const orc_model_t *m, const int16_t *senscr, int T, int n_phones,
                          const int32_t *ssid, const int32_t *tmat, const int32_t *sf,
                          const int32_t *ef, int32_t *st_start, int32_t *st_dur, int32_t *st_score,
                          int32_t *tokens, orc_align_out_t *out);
/* Full variant: scores computed frame by frame with the PTM scorer exactly as
 * state_align_search_step does (activate -> acmod_score -> ...).
 * init_active: optional bit vector (n_sen bits) the acmod starts with (pass 1
 * leftovers), or NULL for empty.  compallsen selects the reference config key. */
int orc_state_align(const orc_model_t *m, int topn, const float *feat, int T, int n_phones,
                    const int32_t *ssid, const int32_t *tmat, const int32_t *sf, const int32_t *ef,
                    const uint32_t *init_active, int compallsen, int32_t *st_start, int32_t *st_dur,
                    int32_t *st_score, int32_t *tokens, int16_t *senscr_out, orc_align_out_t *out);
int orc_state_align2(const orc_model_t *m, int topn, const float *feat, int T, int n_phones,
                     const int32_t *ssid, const int32_t *tmat, const int32_t *sf, const int32_t *ef,
                     const uint32_t *init_active, int compallsen, int32_t *st_start, int32_t *st_dur,
                     int32_t *st_score, int32_t *tokens, int16_t *senscr_out, orc_align_out_t *out,
                     const uint8_t *init_topn);
void orc_ptm_get_carried(const orc_ptm_t *p, uint8_t *cw);
void orc_ptm_set_carried(orc_ptm_t *p, const uint8_t *cw);
/* ---- FSG token-passing search on a flattened lextree (ss_oracle_fsg.c; ref: fsg_search.c,
 * fsg_history.c).  Array layouts = oracle/ref_shim.c:ref_fsg_dump. */
typedef struct orc_fsg_s {
    int32_t n_state, start, final, n_link, n_pnode, n_ciphone, sil;
    int32_t beam, pbeam, wbeam, maxhmmpf;
    const int32_t *link4;     /* [n_link][4] from to logs2prob wid; state s owns arc_off[s]..arc_off[s+1] */
    const uint8_t *link_flag; /* bit0: word models no right context (filler / single phone) */
    const int32_t *arc_off;   /* [n_state+1] */
    const int32_t *root;      /* [n_state] first root pnode or -1 */
    const int32_t *pnode8;    /* [n_pnode][8] ssid tmat logs2prob ci_ext leaf succ|link sibling ppos */
    const uint32_t *ctxt;     /* [n_pnode][4] */
} orc_fsg_t;
int orc_fsg_search(const orc_model_t *m, const orc_fsg_t *g, const int16_t *senscr, int T,
                   int32_t *hist9, int cap, int64_t *out);
int orc_fsg_search_active(const orc_model_t *m, int topn, const orc_fsg_t *g, const float *feat, int T,
                          int32_t *hist9, int cap, int64_t *out, uint32_t *active_out);
int orc_fsg_search_active2(const orc_model_t *m, int topn, const orc_fsg_t *g, const float *feat, int T,
                           int32_t *hist9, int cap, int64_t *out, uint32_t *active_out,
                           uint8_t *carried_out);
int orc_fsg_find_exit(const orc_fsg_t *g, const int32_t *hist9, int n_hist, int frame_idx, int final,
                      int32_t *out_score);
int orc_fsg_segs(const orc_fsg_t *g, const int32_t *hist9, int bpidx, int32_t *segs, int max_seg);

/* ref: ps_alignment.c:317-355 (durations/scores summed upward) */
int orc_propagate(int n_states, int n_emit, const int32_t *st_start, const int32_t *st_dur,
                  const int32_t *st_score, int32_t *ph_start, int32_t *ph_dur, int32_t *ph_score);

/* ---- acoustic frontend (ss_oracle_fe.c), whole utterances ---- */
enum { ORC_FE_DCT = 0, ORC_FE_LEGACY = 1, ORC_FE_HTK = 2 };
enum { ORC_FE_CMN_NONE = 0, ORC_FE_CMN_BATCH = 1 };
typedef struct {
    int32_t samprate, frate, ncep, nfft, nfilt, lifter;
    int32_t remove_dc, remove_noise, unit_area, round_filters, doublebw;
    int32_t transform, cmn, varnorm;
    float wlen, alpha, lowerf, upperf;
} orc_fe_cfg_t;
typedef struct orc_fe_s orc_fe_t;
orc_fe_t *orc_fe_new(const orc_fe_cfg_t *cfg);
void orc_fe_free(orc_fe_t *fe);
int orc_fe_dims(const orc_fe_t *fe, int32_t *out4); /* frame_size shift fft_size n_coeffs */
int orc_fe_tables(const orc_fe_t *fe, int32_t *spec_start, int32_t *filt_width, float *coeffs,
                  float *mel_cosine, float *lifter, double *hamming);
long orc_fe_n_frames(const orc_fe_t *fe, long n_samples);
long orc_fe_mfcc(const orc_fe_t *fe, const int16_t *pcm16, const float *pcm32, long n_samples,
                 float *mfcc, double *melspec);
int orc_fe_feat(const orc_fe_t *fe, float *mfcc, long nfr, float *feat);

#endif
