// device.cuh -- device-side model image and kernel launch prototypes.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "model.h"

namespace ssb {

#define SSB_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            ssb::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,             \
                           cudaGetErrorString(e_));                                         \
            return -1;                                                                      \
        }                                                                                   \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is state of the (device, kernel) pair, shared by
// every host thread: several batches launch the same kernel with different sizes (the lanes of
// ssb_pipeline_*), so the limit is only ever raised, under a lock (api.cu).
cudaError_t raise_dyn_smem_limit(const void *kernel, size_t bytes);
#define SSB_DYN_SMEM(kernel, bytes) \
    SSB_CUDA(ssb::raise_dyn_smem_limit(reinterpret_cast<const void *>(&kernel), (bytes)))

constexpr int32_t WORST_SCORE = (int32_t)0xE0000000;
constexpr int MAX_NEG_ASCR = 96;
constexpr int SENSCR_SHIFT = 10;

// Packed model in HBM.  Gaussians: one record per density,
//   rec = [det, mean[0..L), prec[0..L), 0-pad]  (rec_len floats, multiple of 4)
// so that a whole record is a run of aligned float4 loads.
struct DevModel {
    int32_t n_mgau, n_feat, n_density, n_sen, n_emit, n_tmat, n_sseq, blk, topn, ds;
    int32_t featlen[SSB_MAX_FEAT];
    int32_t featoff[SSB_MAX_FEAT];
    int32_t rec_len[SSB_MAX_FEAT];   // floats per density record
    int64_t gau_base[SSB_MAX_FEAT];  // float offset of stream f of codebook 0
    int64_t gau_cb_stride;           // floats between consecutive codebooks
    const float *gau;                // packed records
    const uint8_t *mixw;             // [feat][density][n_sen]
    const uint8_t *sen2cb;           // [n_sen]
    const uint16_t *sseq;            // [n_sseq][n_emit]
    const uint8_t *tp;               // [n_tmat][n_emit][n_emit+1]
    const uint8_t *lut8;             // [256]
    // senones grouped by codebook (for the dense scorer)
    const int32_t *cb_sen_off;       // [n_mgau+1]
    const uint16_t *cb_sen;          // [n_sen] senone ids sorted by (codebook, id)
    int32_t max_cb_sen;
    // tensor-core screening operands (gmm_topn_tc.cu): per codebook-stream
    //   gB   [cs][128][32] TF32-rounded rows [2 mu' v (13) | -v (13) | c_hi | c_lo | 0 x4]
    //   gAux [cs][SSB_TC_AUX] centre[13] | max|2 mu' v|[13] | max v[13] | max|c| (all densities)
    //                      | [40..65] the same maxima over regular densities | [67..92] over hot ones
    //   gBlo [cs][128][32] TF32 residuals B - gB (3xTF32 split, v2 kernel)
    //   gHot [cs][4]       bit n: density n is "hot" (outlier precision, own error bound)
    const float *gB;
    const float *gBlo;
    const float *gAux;
    const uint32_t *gHot;
    // frame-tiled kernel (gmm_scan_ft.cu): ONE centre per stream, so that the A tile of a frame
    // serves every codebook
    //   gBft   [cs][2][128][32]  B_hi then B_lo, already in the SWIZZLE_128B shared-memory order
    //                            (one 32 KB cp.async.bulk per codebook-stream)
    //   gAuxFt [cs][64]  [0..12] max|2 mu' v|, [16..28] max v over the regular densities;
    //                    [32..44], [48..60] the same over the hot ones; [13]/[29] max|c| regular,
    //                    [14]/[30] max|c| hot, [15]/[31] 1.0 if the codebook-stream has hot densities
    //   ft_centre [feat][16]
    const float *gBft;
    const float *gAuxFt;
    float ft_centre[SSB_MAX_FEAT * 16];
    // scorer family (ssb200.h SSB_SCORER_*): PTM, or the single-codebook semi-continuous one
    // (ref: src/s2_semi_mgau.c) with its per-stream top-N beam
    int32_t kind;
    int32_t topn_beam[SSB_MAX_FEAT];
};
constexpr int SSB_TC_AUX = 96;

__host__ __device__ inline int64_t gau_offset(const DevModel &m, int cb, int f)
{
    return m.gau_base[f] + (int64_t)cb * m.gau_cb_stride;
}

// ---- active-set plan of one batch (mode "compallsen = no") ----
// Utterance u has epochs [ep_off[u], ep_off[u+1]); epoch e starts at frame
// ep_start[e] and activates union slots ep_slot[ep_slot_off[e] .. ep_slot_off[e+1])
// of the utterance's senone union usen[us_off[u] .. us_off[u+1]).
struct DevPlan {
    int32_t n_utts;
    const int64_t *frame_off;   // [U+1]
    const int64_t *phone_off;   // [U+1]
    const int64_t *scr_off;     // [U+1] offsets into chain_scr / tokens (state-frames)
    const int32_t *ssid, *tmat, *sf, *ef;  // [total phones]
    const int32_t *ep_off;      // [U+1]
    const int32_t *ep_start;    // [n_epochs]
    const uint32_t *ep_cbmask;  // [n_epochs][8]   256-bit active-codebook mask
    const int32_t *ep_slot_off; // [n_epochs+1]
    const uint16_t *ep_slot;    // union slots active in the epoch, ascending senone order
    const int32_t *us_off;      // [U+1]
    const uint16_t *usen;       // union senone ids (sorted) per utterance
    const uint16_t *st_slot;    // [total states] union slot of each chain state
    const int32_t *enter_plan;  // [total phones] planned first-entry frame (-1 never)
    int32_t all_active;         // compallsen: every codebook scanned on every frame
    // optional (K4 with active lists): bit (global frame) of word row cs is set by the top-N
    // kernel when the step's integer scores tie, i.e. when the reference's list depends on the
    // list it carried in; [n_mgau*n_feat][tie_w] words, zeroed by the caller
    uint32_t *tie_bits;
    int64_t tie_w;
    // optional: the top-N codewords the scorer carries when the pass starts, [n_utts][CS] (what
    // a first pass on the same decoder left: the lists are not reset between passes, ref:
    // src/ptm_mgau.c:426-440); null = the initial lists (codewords 0..N-1)
    const uchar4 *init_topn;
    // banded layout of chain_scr / tokens (the reference's dense [T][n_states] stack is its memory
    // hazard: 86 GB for an hour): phone i keeps only the frames it can be
    // evaluated on, as one contiguous block per phone
    //   scores: frames [enter[i], last[i]]           at chain_scr + scr_boff[i] + t * E + j
    //   tokens: frames [max(enter[i]-1, 0), last[i]] at tokens + tok_boff[i] + t * E + j
    // with last[i] = min(max(enter[i], ef[i]), T-1); the offsets are absolute (whole batch), per
    // phone, and VIRTUAL: block start minus (first frame) * E, so that no kernel needs the first
    // frame to address a block
    int32_t banded;
    const int64_t *scr_boff;
    const int64_t *tok_boff;
    // chain cutting (api.cu: cut_chains): the "utterances" of this plan are SEGMENTS of the real
    // ones -- runs of phones between two points where the chain can only be crossed on one
    // known frame (a word window ending where the next begins) -- and seg_t0[u] is the frame of
    // the real utterance the segment starts on; null = the plan's utterances are the real ones
    const int32_t *seg_t0;
};

// ---- kernel launchers (each returns 0 or -1 with the error set) ----
// K1: stateful top-N of every (utterance, codebook, stream) chain.
// featp: scratch of tc2_featp_bytes() for the tensor-core path (NULL: CUDA-core / v1 kernels)
int launch_gmm_topn(const DevModel &m, const DevPlan &p, const float *feat, int64_t n_frames,
                    int4 *tn_score, uchar4 *tn_cw, float *featp, cudaStream_t st);
size_t tc2_featp_bytes(const DevModel &m, int64_t n_frames);
// K1 on the tensor cores (tcgen05 TF32 screening + exact FP32 re-scoring); same results.
// dbg_* may be NULL: approx [cs][frame][128], eps [cs][frame], counters [2].
bool tc_supported(const DevModel &m);
int launch_gmm_topn_tc(const DevModel &m, const DevPlan &p, const float *feat, int64_t n_frames,
                       int4 *tn_score, uchar4 *tn_cw, float *featp, float *dbg_approx,
                       float *dbg_eps, unsigned long long *dbg_counters, cudaStream_t st);
// K1, frame-tiled (gmm_scan_ft.cu): tiles of 128 consecutive frames (tile_utt, tile_t0), tie steps
// flagged in p.tie_bits; scores are stored as (score >> 10) << 10 unless `exact`
struct TcDebug {
    float *approx;     // [cs][frame][128] or null
    float *eps;        // [cs][frame][2]
    unsigned long long *counters;  // [0] exact evaluations of scan survivors [1] scanned (lane, frame) steps [2] slow-path steps
};
bool ft_supported(const DevModel &m);
int launch_gmm_scan_ft(const DevModel &m, const DevPlan &p, const float *feat, int64_t n_frames,
                       int4 *tn_score, uchar4 *tn_cw, const int32_t *tile_utt, const int32_t *tile_t0,
                       int n_tiles, int *tile_counter, int exact, TcDebug dbg, cudaStream_t st);
// K2 (active lists): normalise, mix, subtract best, gather to chain states.
// max_union counts the always-zero slot that inactive chain states read.
int launch_senone_mix_active(const DevModel &m, const DevPlan &p, const int4 *tn_score,
                             const uchar4 *tn_cw, int64_t n_frames, int max_union,
                             int max_frames_per_utt, int16_t *chain_scr, cudaStream_t st);
// K2 (dense, compallsen): every senone of frames [g0, g0+n) -> dense [n][n_sen], best subtracted
int launch_senone_mix_all(const DevModel &m, const int4 *tn_score, const uchar4 *tn_cw,
                          int64_t n_frames_total, int64_t g0, int64_t n, int16_t *dense,
                          cudaStream_t st);
// dense scores of frames [g0, ...) -> chain states of utterances [u0, u1)
int launch_gather_chain(const DevModel &m, const DevPlan &p, const int16_t *dense, int u0, int u1,
                        int64_t g0, int16_t *chain_scr, cudaStream_t st);
// K3: chain Viterbi + token stack; K3b: backtrace
int launch_chain_viterbi(const DevModel &m, const DevPlan &p, const int16_t *chain_scr,
                         int2 *tokens, int32_t *spill, int64_t spill_stride, int32_t *utt_best,
                         int32_t *utt_renorm, int32_t *fin_hist, int32_t *fin_score,
                         int max_phones, int max_band, cudaStream_t st);
int launch_backtrace(const DevModel &m, const DevPlan &p, const int2 *tokens,
                     const int32_t *fin_hist, const int32_t *fin_score, int32_t *st_start,
                     int32_t *st_dur, int32_t *st_score, int32_t *utt_rv, cudaStream_t st);
// ---- K4: FSG token passing (fsg_search.cu) ----
struct DevFsg {  // one grammar: sizes, beams, offsets into the concatenated arrays of DevFsgSet
    int32_t n_state, start, final, n_link, n_pnode, n_ciphone, sil, beam, pbeam, wbeam, maxhmmpf;
    int32_t link_off, arcoff_off, root_off, pnode_off;
};
struct DevFsgSet {
    const DevFsg *graph;       // [n_graphs]
    const int32_t *link4;      // concatenated [.][4]: from to logs2prob wid
    const uint8_t *link_flag;  // concatenated
    const int32_t *arc_off;    // concatenated, n_state+1 per graph (values local to the graph)
    const int32_t *root;       // concatenated
    const int32_t *pnode8;     // concatenated [.][8]: ssid tmat logs2prob ci_ext leaf succ|link sibling ppos
    const uint32_t *ctxt;      // concatenated [.][4]
};
constexpr int FSG_PH = 14;  // ints of HMM state per pnode
constexpr int FSG_TE = 10;  // ints per tentative history entry
int launch_fsg_search(const DevModel &m, const DevFsgSet &gs, const int64_t *frame_off,
                      const int32_t *utt_graph, const int64_t *ws_off, int32_t *ws,
                      const int16_t *dense, int64_t g0, int u0, int n_utts, int32_t *hist,
                      int hist_cap, int tent_cap, int32_t *n_hist, int64_t *n_eval, int32_t *frames,
                      int32_t *rv, cudaStream_t st);
// the same search in the reference's default mode: senone scores computed inside the kernel for
// the active HMMs' senones, from K1's dense top-N lists (+ tie flags) and the mixture weights
size_t fsg_active_ws_ints(const DevModel &m, int n_pnode);
int launch_fsg_search_active(const DevModel &m, const DevFsgSet &gs, const int64_t *frame_off,
                             const int32_t *utt_graph, const int64_t *ws_off, int32_t *ws,
                             const float *feat, const int4 *tn_s, const uchar4 *tn_c,
                             const uint32_t *tie, int64_t G, int64_t tie_w, const int64_t *aws_off,
                             int32_t *aws, uint32_t *final_active, int64_t *n_sen_eval,
                             uchar4 *final_topn, int n_utts,
                             int32_t *hist, int hist_cap, int tent_cap, int32_t *n_hist,
                             int64_t *n_eval, int32_t *frames, int32_t *rv, cudaStream_t st);
// dense mode: the carried lists after the search are K1's own lists of the last odd frame
int launch_fsg_final_topn_dense(const DevModel &m, const int64_t *frame_off, int n_utts,
                                const uchar4 *tn_c, int64_t G, uchar4 *final_topn, cudaStream_t st);
int launch_fsg_backtrace(const DevFsgSet &gs, const int32_t *utt_graph, int u0, int n_utts,
                         const int32_t *hist, int hist_cap, const int32_t *n_hist,
                         const int32_t *frames, int32_t *exit_bp, int32_t *hyp_score, int32_t *segs,
                         int max_seg, int32_t *n_seg, int final, cudaStream_t st);
// single-frame scorer behind the mgau vtable
struct FrameHist {
    int4 *score[2];    // [CS] per history slot
    uchar4 *cw[2];
    uint8_t *act[2];   // [n_mgau]
};
int launch_frame_topn(const DevModel &m, const FrameHist &h, int slot, int prev, const float *x,
                      int do_scan, cudaStream_t st);
int launch_frame_senones(const DevModel &m, const FrameHist &h, int slot, int do_norm,
                         const uint16_t *act_sen, int n_act, int compallsen, int16_t *senscr,
                         cudaStream_t st);
int launch_hmm_eval_tp(int n_emit, int n_cases, const uint8_t *tp, const int16_t *senscr, int32_t *st12,
                       int32_t *best, cudaStream_t st);
int launch_hmm_eval(const DevModel &m, int n_emit, int tmatid, const uint16_t *senid,
                    const int16_t *senscr, int32_t *st12, int32_t *best, cudaStream_t st);

}  // namespace ssb
