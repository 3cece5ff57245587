// topn_fixup.cu -- K1 over time: long utterances of small batches are cut into segments so that
// the Gaussian top-N kernel (thread = utterance, walking its frames in order) has rows to fill
// its CTAs with -- one hour of audio is 360 000 dependent steps on a single thread otherwise.
//
// That is possible because the reference's list after a scan (ref: src/ptm_mgau.c:86-225) is
// "the N best, sorted" whatever list it carried in, UNLESS integer scores tie; the tensor-core
// kernel flags those steps (DevPlan.tie_bits).  Segments are scored independently; this kernel
// then visits the flagged steps of each (utterance, codebook-stream) in time order -- one warp
// each -- and replays them literally from the list the reference carried:
//   * a codebook is scanned on every frame from the one it becomes active on (the aligner's
//     active set only grows, ref: src/state_align_search.c:186-188; compallsen scans
//     everything), so at a flagged frame t the carried list is the final list of frame t-1;
//   * on the very first scan the carried list is the initial one (codewords 0..N-1, or what a
//     previous pass left: DevPlan.init_topn) as
//     eval_topn has re-scored and stably re-sorted it on every frame before (ref :234-237) --
//     re-played from the most recent frame on which those N scores are pairwise distinct
//     (their order is forced there), or from the utterance's first frame.
#include "tc_common.cuh"

namespace ssb {

__device__ __forceinline__ float fix_gau_dist(const float *__restrict__ rec, const float *__restrict__ x, int L)
{
    float d = rec[0];  // the reference's operation order, no contraction (ref :63-68)
    for (int j = 0; j < L; ++j) {
        const float diff = __fsub_rn(x[j], rec[1 + j]);
        const float sq = __fmul_rn(diff, diff);
        const float c = __fmul_rn(sq, rec[1 + L + j]);
        d = __fsub_rn(d, c);
    }
    return d;
}

template <int N>
__device__ __forceinline__ void fix_eval_topn(TcTopN<N> &tn, const float *rec, int RL, const float *x, int L)
{
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        int32_t ci = tn.c[0];
#pragma unroll
        for (int k = 1; k < N; ++k)
            if (k == i)
                ci = tn.c[k];
        const int32_t sc = __float2int_rz(fix_gau_dist(rec + (int64_t)ci * RL, x, L));
#pragma unroll
        for (int k = 0; k < N; ++k)
            if (k == i)
                tn.s[k] = sc;
        tn.settle(i);
    }
}

template <int N>
__global__ void __launch_bounds__(32)
topn_fixup_kernel(DevModel m, DevPlan p, const int32_t *__restrict__ seg_utts,
                  const float *__restrict__ feat, int64_t G, int4 *__restrict__ tn_s,
                  uchar4 *__restrict__ tn_c, const uint32_t *__restrict__ tie, int64_t tie_w)
{
    __shared__ float sd[256];
    const int cs = blockIdx.x, u = seg_utts[blockIdx.y], lane = threadIdx.x;
    const int cb = cs / m.n_feat, f = cs - cb * m.n_feat;
    const int RL = m.rec_len[f], L = m.featlen[f], ND = m.n_density;
    const float *rec = m.gau + gau_offset(m, cb, f);
    const int64_t g0 = p.frame_off[u];
    const int T = (int)(p.frame_off[u + 1] - g0);
    const int e_lo = p.all_active ? 0 : p.ep_off[u], e_hi = p.all_active ? 0 : p.ep_off[u + 1];
    // the utterance's flag words, 32 at a time (one coalesced read; a long recording has tens of
    // thousands of them per codebook-stream and almost all are zero); the flagged steps are then
    // replayed in time order -- a replay's list is what the next one starts from
    const int64_t w_lo = g0 >> 5, w_hi = T > 0 ? (g0 + T - 1) >> 5 : w_lo - 1;
    for (int64_t wb = w_lo; wb <= w_hi; wb += 32) {
      const uint32_t mine = wb + lane <= w_hi ? tie[(int64_t)cs * tie_w + wb + lane] : 0u;
      uint32_t nz = __ballot_sync(0xffffffffu, mine != 0u);
      while (nz) {
        const int src = __ffs((int)nz) - 1;
        nz &= nz - 1u;
        const int64_t w = wb + src;
        uint32_t bits = __shfl_sync(0xffffffffu, mine, src);
        while (bits) {
            const int64_t g = w * 32 + (__ffs((int)bits) - 1);
            bits &= bits - 1;
            if (g < g0 || g >= g0 + T)
                continue;
            const int t = (int)(g - g0);
            const float *x = feat + g * m.blk + m.featoff[f];
            for (int n = lane; n < ND; n += 32)
                sd[n] = fix_gau_dist(rec + (int64_t)n * RL, x, L);
            __syncwarp();
            if (lane == 0) {
                TcTopN<N> tn;
                tn.reset();
                if (p.init_topn) {  // the lists a previous pass left
                    const uchar4 c = p.init_topn[(int64_t)u * gridDim.x + cs];
                    const int cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                    for (int k = 0; k < N; ++k)
                        tn.c[k] = cc[k];
                }
                // last frame before t on which this codebook was scanned: the frames of the epochs
                // whose mask holds it (the evaluated list is not monotone -- the >255-gap bridging
                // senones come and go, ref: src/acmod.c:968-973)
                int t_prev = p.all_active ? t - 1 : -1;
                {
                    // (the last epoch that starts before t and holds the codebook: binary search for
                    // the first epoch at or after t, then backwards)
                    int a = e_lo, b = e_hi;
                    while (a < b) {
                        const int mid = (a + b) >> 1;
                        if (p.ep_start[mid] >= t)
                            b = mid;
                        else
                            a = mid + 1;
                    }
                    for (int e = a - 1; e >= e_lo; --e)
                        if ((p.ep_cbmask[(int64_t)e * 8 + (cb >> 5)] >> (cb & 31)) & 1u) {
                            const int s1 = e + 1 < e_hi ? p.ep_start[e + 1] : T;
                            t_prev = min(s1, t) - 1;
                            break;
                        }
                }
                if (t_prev >= 0) {
                    // the final list of that scan (already replayed if it was flagged itself)
                    const uchar4 c = tn_c[(int64_t)cs * G + g0 + t_prev];
                    const int cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                    for (int k = 0; k < N; ++k)
                        tn.c[k] = cc[k];
                }
                if (t_prev < t - 1) {
                    // frames t_prev+1 .. t-1 did not scan it: eval_topn re-scored and stably
                    // re-sorted the carried codewords on each of them (ref :234-237); replayed
                    // from the most recent frame on which their scores are pairwise distinct
                    const int lo = t_prev + 1;
                    int tt0 = lo;
                    for (int tt = t - 1; tt > lo; --tt) {
                        const float *xx = feat + (g0 + tt) * m.blk + m.featoff[f];
                        int32_t s[N];
#pragma unroll
                        for (int k = 0; k < N; ++k)
                            s[k] = __float2int_rz(fix_gau_dist(rec + (int64_t)tn.c[k] * RL, xx, L));
                        bool distinct = true;
#pragma unroll
                        for (int i = 0; i < N; ++i)
#pragma unroll
                            for (int j = i + 1; j < N; ++j)
                                distinct = distinct && s[i] != s[j];
                        if (distinct) {
                            tt0 = tt;
                            break;
                        }
                    }
#pragma unroll 1
                    for (int tt = tt0; tt < t; ++tt)
                        fix_eval_topn<N>(tn, rec, RL, feat + (g0 + tt) * m.blk + m.featoff[f], L);
                }
                // eval_topn + eval_cb of frame t on the distances held in shared memory
#pragma unroll 1
                for (int i = 0; i < N; ++i) {
                    int32_t ci = tn.c[0];
#pragma unroll
                    for (int k = 1; k < N; ++k)
                        if (k == i)
                            ci = tn.c[k];
                    const int32_t sc = __float2int_rz(sd[ci]);
#pragma unroll
                    for (int k = 0; k < N; ++k)
                        if (k == i)
                            tn.s[k] = sc;
                    tn.settle(i);
                }
#pragma unroll 1
                for (int cw = 0; cw < ND; ++cw) {
                    const float d = sd[cw];
                    if (d < __int2float_rn(tn.s[N - 1]))
                        continue;
                    if (tn.has(cw))
                        continue;
                    tn.insert(__float2int_rz(d), cw);
                }
                int4 sv = make_int4(INT32_MIN, INT32_MIN, INT32_MIN, INT32_MIN);
                uchar4 cv = make_uchar4(0, 0, 0, 0);
                sv.x = tn.s[0];
                cv.x = (unsigned char)tn.c[0];
                if (N > 1) {
                    sv.y = tn.s[N > 1 ? 1 : 0];
                    cv.y = (unsigned char)tn.c[N > 1 ? 1 : 0];
                }
                if (N > 2) {
                    sv.z = tn.s[N > 2 ? 2 : 0];
                    cv.z = (unsigned char)tn.c[N > 2 ? 2 : 0];
                }
                if (N > 3) {
                    sv.w = tn.s[N > 3 ? 3 : 0];
                    cv.w = (unsigned char)tn.c[N > 3 ? 3 : 0];
                }
                tn_s[(int64_t)cs * G + g] = sv;
                tn_c[(int64_t)cs * G + g] = cv;
            }
            __syncwarp();
        }
      }
    }
}

int launch_topn_fixup(const DevModel &m, const DevPlan &p, const int32_t *seg_utts, int n_seg_utts,
                      const float *feat, int64_t n_frames, int4 *tn_s, uchar4 *tn_c,
                      const uint32_t *tie, int64_t tie_w, cudaStream_t st)
{
    if (n_seg_utts <= 0)
        return 0;
    if (m.n_density > 256) {
        set_error("topn_fixup: more than 256 densities");
        return -1;
    }
    dim3 grid(m.n_mgau * m.n_feat, n_seg_utts);
    switch (m.topn) {
    case 1: topn_fixup_kernel<1><<<grid, 32, 0, st>>>(m, p, seg_utts, feat, n_frames, tn_s, tn_c, tie, tie_w); break;
    case 2: topn_fixup_kernel<2><<<grid, 32, 0, st>>>(m, p, seg_utts, feat, n_frames, tn_s, tn_c, tie, tie_w); break;
    case 3: topn_fixup_kernel<3><<<grid, 32, 0, st>>>(m, p, seg_utts, feat, n_frames, tn_s, tn_c, tie, tie_w); break;
    case 4: topn_fixup_kernel<4><<<grid, 32, 0, st>>>(m, p, seg_utts, feat, n_frames, tn_s, tn_c, tie, tie_w); break;
    default:
        set_error("topn %d not supported (1..4)", m.topn);
        return -1;
    }
    SSB_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

}  // namespace ssb
