// cont_score.cu -- senone scoring for fully continuous models (one codebook per senone).
//
// Replaces ms_cont_mgau_frame_eval and its callees (ref: src/ms_mgau.c:279-368,
// src/ms_gauden.c:342-457 gauden_dist, src/ms_senone.c:315-362 senone_eval,
// src/logmath.c:229-275 logmath_add).  Unlike the tied scorers this one is stateless: per
// frame and senone the top-N densities are selected from scratch on FLOAT distances
// (newcomers go in front of equal scores; with topn >= n_density the list is simply all
// densities in index order), each is converted with (int + 1023) >> 10, mixed with the 8-bit
// weights through the table log-add, negated, clamped to int16; then the best score over the
// evaluated senones is subtracted (clamped again).  Bit-exact: distances use the reference's
// operation order with explicit IEEE intrinsics.
//
// Mapping: thread = senone, tile of CT_F frames per thread -- a density record (det, means,
// precisions) is read once per tile and applied to all its frames, whose feature vectors sit in
// shared memory (broadcast reads).  Raw scores go to HBM as int16; small finishing kernels do the
// per-frame minimum (dense rows, or the epoch's active list of an utterance for the aligner).
#include "device.cuh"

namespace ssb {

constexpr int CT_THREADS = 128;
constexpr int CT_F = 8;
constexpr int CT_ZERO = (int)0x80000000 >> (SENSCR_SHIFT + 2);  // logmath zero at shift 10

__device__ __forceinline__ int cont_table_add(int x, int y, const uint8_t *__restrict__ lut)
{
    if (x <= CT_ZERO)
        return y;
    if (y <= CT_ZERO)
        return x;
    int d, r;
    if (x > y) {
        d = x - y;
        r = x;
    } else {
        d = y - x;
        r = y;
    }
    if (d < 0 || d >= 256)  // the reference's table ends (with zeros) well before 256
        return r;
    return r + __ldg(lut + d);
}

__device__ __forceinline__ int cont_density_score(float dist)
{
    // ref: src/ms_senone.c:333-337
    if (dist < -2147483648.0f)
        return (int)0x80000000 >> SENSCR_SHIFT;
    return (__float2int_rz(dist) + ((1 << SENSCR_SHIFT) - 1)) >> SENSCR_SHIFT;
}

// Scores of senone `sen` for the nf <= CT_F frames whose features are xs[k][blk].
template <int N>
__device__ __forceinline__ void cont_senone_tile(const DevModel &m, int sen, const float *xs,
                                                 int nf, int (&scr)[CT_F])
{
    const int ND = m.n_density;
    const bool all = N >= ND;  // compute_dist_all: every density, index order
#pragma unroll
    for (int k = 0; k < CT_F; ++k)
        scr[k] = 0;
    for (int f = 0; f < m.n_feat; ++f) {
        const int L = m.featlen[f], RL = m.rec_len[f], xo = m.featoff[f];
        const float *rec = m.gau + gau_offset(m, sen, f);
        float dist[CT_F][N];
        uint32_t ids[CT_F];
#pragma unroll
        for (int k = 0; k < CT_F; ++k) {
            ids[k] = 0u;
#pragma unroll
            for (int t = 0; t < N; ++t)
                dist[k][t] = -2147483648.0f;  // WORST_DIST as a float
        }
        for (int d = 0; d < ND; ++d) {
            const float *r = rec + (size_t)d * RL;
            float dv[CT_F];
            const float det = __ldg(r);
#pragma unroll
            for (int k = 0; k < CT_F; ++k)
                dv[k] = det;
            for (int i = 0; i < L; ++i) {
                const float mu = __ldg(r + 1 + i), pv = __ldg(r + 1 + L + i);
#pragma unroll
                for (int k = 0; k < CT_F; ++k) {
                    const float diff = __fsub_rn(xs[k * m.blk + xo + i], mu);
                    dv[k] = __fsub_rn(dv[k], __fmul_rn(__fmul_rn(diff, diff), pv));
                }
            }
#pragma unroll
            for (int k = 0; k < CT_F; ++k) {
                if (all) {
#pragma unroll
                    for (int t = 0; t < N; ++t)
                        if (t == d) {
                            dist[k][t] = dv[k];
                            ids[k] |= (uint32_t)d << (8 * t);
                        }
                    continue;
                }
                // ref: src/ms_gauden.c:398-418 -- dropped when below the worst; otherwise
                // placed in front of the first entry that is not strictly better
                if (dv[k] < dist[k][N - 1])
                    continue;
                float cd = dv[k];
                uint32_t ci = (uint32_t)d;
                bool placed = false;
#pragma unroll
                for (int t = 0; t < N; ++t) {
                    const uint32_t it = (ids[k] >> (8 * t)) & 0xffu;
                    if (placed || !(dv[k] < dist[k][t])) {
                        // this slot takes the carried entry; the old one is carried on
                        const float od = dist[k][t];
                        dist[k][t] = cd;
                        ids[k] = (ids[k] & ~(0xffu << (8 * t))) | (ci << (8 * t));
                        cd = od;
                        ci = it;
                        placed = true;
                    }
                }
            }
        }
        const uint8_t *pdf = m.mixw + ((size_t)sen * m.n_feat + f) * ND;
        const int n = all ? ND : N;
#pragma unroll
        for (int k = 0; k < CT_F; ++k) {
            if (k >= nf)
                continue;
            int fscr = cont_density_score(dist[k][0]) - (int)__ldg(pdf + (ids[k] & 0xffu));
#pragma unroll
            for (int t = 1; t < N; ++t)
                if (t < n)
                    fscr = cont_table_add(fscr, cont_density_score(dist[k][t])
                                                    - (int)__ldg(pdf + ((ids[k] >> (8 * t)) & 0xffu)),
                                          m.lut8);
            scr[k] -= fscr;
        }
    }
#pragma unroll
    for (int k = 0; k < CT_F; ++k)  // "aw" is 1; avoid overflowing int16 (ref :355-361)
        scr[k] = max(-32768, min(32767, scr[k]));
}

// dense: raw [n][n_sen] of frames [g0, g0+n)
template <int N>
__global__ void __launch_bounds__(CT_THREADS)
cont_raw_dense_kernel(DevModel m, const float *__restrict__ feat, int64_t g0, int64_t n,
                      int16_t *__restrict__ raw)
{
    extern __shared__ float ct_x[];
    const int64_t t0 = (int64_t)blockIdx.y * CT_F;
    const int nf = (int)min((int64_t)CT_F, n - t0);
    for (int i = threadIdx.x; i < CT_F * m.blk; i += blockDim.x)
        ct_x[i] = i < nf * m.blk ? feat[(g0 + t0) * m.blk + i] : 0.f;
    __syncthreads();
    const int sen = blockIdx.x * CT_THREADS + threadIdx.x;
    if (sen >= m.n_sen)
        return;
    int scr[CT_F];
    cont_senone_tile<N>(m, sen, ct_x, nf, scr);
#pragma unroll
    for (int k = 0; k < CT_F; ++k)
        if (k < nf)
            raw[(t0 + k) * m.n_sen + sen] = (int16_t)scr[k];
}

// aligner: raw [frame][W] for the senone union of every utterance
template <int N>
__global__ void __launch_bounds__(CT_THREADS)
cont_raw_union_kernel(DevModel m, DevPlan p, const float *__restrict__ feat, int W,
                      int16_t *__restrict__ raw)
{
    extern __shared__ float ct_x[];
    const int u = blockIdx.y;
    const int64_t f0 = p.frame_off[u];
    const int T = (int)(p.frame_off[u + 1] - f0);
    const int t0 = blockIdx.x * CT_F;
    if (t0 >= T)
        return;
    const int nf = min(CT_F, T - t0);
    for (int i = threadIdx.x; i < CT_F * m.blk; i += blockDim.x)
        ct_x[i] = i < nf * m.blk ? feat[(f0 + t0) * m.blk + i] : 0.f;
    __syncthreads();
    const int us0 = p.us_off[u], n_us = p.us_off[u + 1] - us0;
    for (int j = threadIdx.x; j < n_us; j += blockDim.x) {
        int scr[CT_F];
        cont_senone_tile<N>(m, p.usen[us0 + j], ct_x, nf, scr);
#pragma unroll
        for (int k = 0; k < CT_F; ++k)
            if (k < nf)
                raw[(f0 + t0 + k) * W + j] = (int16_t)scr[k];
    }
}

__device__ __forceinline__ int16_t cont_clamp16(int v)
{
    return (int16_t)max(-32768, min(32767, v));
}

// one CTA per frame: subtract the best of the row (ref: src/ms_mgau.c:303-321)
__global__ void __launch_bounds__(256)
cont_dense_finish_kernel(int n_sen, int16_t *__restrict__ dense)
{
    __shared__ int s_best;
    int16_t *row = dense + (int64_t)blockIdx.x * n_sen;
    if (threadIdx.x == 0)
        s_best = INT32_MAX;
    __syncthreads();
    int best = INT32_MAX;
    for (int s = threadIdx.x; s < n_sen; s += blockDim.x)
        best = min(best, (int)row[s]);
    for (int o = 16; o > 0; o >>= 1)
        best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0)
        atomicMin(&s_best, best);
    __syncthreads();
    best = s_best;
    for (int s = threadIdx.x; s < n_sen; s += blockDim.x)
        row[s] = cont_clamp16((int)row[s] - best);
}

// one warp per (utterance, frame): best over the epoch's active senones, gather to the chain
// states (ref: src/ms_mgau.c:323-364); states whose senone is not active read 0
__global__ void __launch_bounds__(256)
cont_chain_finish_kernel(DevModel m, DevPlan p, const int16_t *__restrict__ raw, int W,
                         int16_t *__restrict__ chain_scr)
{
    const int u = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t f0 = p.frame_off[u];
    const int T = (int)(p.frame_off[u + 1] - f0);
    const int64_t ph0 = p.phone_off[u];
    const int ns = (int)(p.phone_off[u + 1] - ph0) * m.n_emit;
    const int e0 = p.ep_off[u], e1 = p.ep_off[u + 1];
    const int n_us = p.us_off[u + 1] - p.us_off[u];
    if (ns == 0 || e0 == e1)
        return;
    const uint16_t *st_slot = p.st_slot + ph0 * m.n_emit;
    for (int t = blockIdx.x * 8 + warp; t < T; t += gridDim.x * 8) {
        int lo = e0, hi = e1;  // last epoch that started at or before t
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (p.ep_start[mid] <= t)
                lo = mid;
            else
                hi = mid;
        }
        const int sl0 = p.ep_slot_off[lo], na = p.ep_slot_off[lo + 1] - sl0;
        const int16_t *row = raw + (f0 + t) * W;
        int best = INT32_MAX;
        for (int i = lane; i < na; i += 32)
            best = min(best, (int)row[p.ep_slot[sl0 + i]]);
        for (int o = 16; o > 0; o >>= 1)
            best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        int16_t *dst = chain_scr + p.scr_off[u] + (int64_t)t * ns;
        for (int si = lane; si < ns; si += 32) {
            const int slot = st_slot[si];
            dst[si] = slot < n_us ? cont_clamp16((int)row[slot] - best) : (int16_t)0;
        }
    }
}

// vtable path, one frame: raw [n_sen] -> senscr; with an active list only the listed entries
// are touched, the others keep what earlier frames left there (like the reference's buffer)
__global__ void __launch_bounds__(512)
cont_frame_finish_kernel(int n_sen, const int16_t *__restrict__ raw,
                         const uint16_t *__restrict__ act_sen, int n_act, int compallsen,
                         int16_t *__restrict__ senscr)
{
    __shared__ int s_best;
    if (threadIdx.x == 0)
        s_best = INT32_MAX;
    __syncthreads();
    const int n = compallsen ? n_sen : n_act;
    int best = INT32_MAX;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        best = min(best, (int)raw[compallsen ? i : act_sen[i]]);
    if (best != INT32_MAX)
        atomicMin(&s_best, best);
    __syncthreads();
    best = s_best;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int s = compallsen ? i : act_sen[i];
        senscr[s] = cont_clamp16((int)raw[s] - best);
    }
}

#define SSB_CT_DISPATCH(KERNEL, GRID, SMEM, ...)                                     \
    switch (m.topn) {                                                                \
    case 1: KERNEL<1><<<GRID, CT_THREADS, SMEM, st>>>(__VA_ARGS__); break;           \
    case 2: KERNEL<2><<<GRID, CT_THREADS, SMEM, st>>>(__VA_ARGS__); break;           \
    case 3: KERNEL<3><<<GRID, CT_THREADS, SMEM, st>>>(__VA_ARGS__); break;           \
    case 4: KERNEL<4><<<GRID, CT_THREADS, SMEM, st>>>(__VA_ARGS__); break;           \
    default:                                                                         \
        set_error("topn %d not supported (1..4)", m.topn);                           \
        return -1;                                                                   \
    }

static int cont_check(const DevModel &m)
{
    if (m.n_density > 256 || (size_t)CT_F * m.blk * sizeof(float) > 48 * 1024) {
        set_error("continuous scorer: %d densities / %d feature dimensions not supported",
                  m.n_density, m.blk);
        return -1;
    }
    return 0;
}

// dense [n][n_sen] = final senone scores of frames [g0, g0 + n)  (compallsen semantics)
int launch_cont_dense(const DevModel &m, const float *feat, int64_t g0, int64_t n, int16_t *dense,
                      cudaStream_t st)
{
    if (n == 0)
        return 0;
    if (cont_check(m) != 0)
        return -1;
    const size_t smem = (size_t)CT_F * m.blk * sizeof(float);
    dim3 grid((m.n_sen + CT_THREADS - 1) / CT_THREADS, (unsigned)((n + CT_F - 1) / CT_F));
    SSB_CT_DISPATCH(cont_raw_dense_kernel, grid, smem, m, feat, g0, n, dense)
    SSB_CUDA(cudaGetLastError());
    note_launch();
    cont_dense_finish_kernel<<<(unsigned)n, 256, 0, st>>>(m.n_sen, dense);
    SSB_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// aligner default mode: raw scores of every utterance's senone union (scratch [frames][W]),
// then per frame the best over the epoch's active list and the gather to chain states
int launch_cont_active(const DevModel &m, const DevPlan &p, const float *feat, int W,
                       int max_frames_per_utt, int16_t *scratch, int16_t *chain_scr,
                       cudaStream_t st)
{
    if (p.n_utts == 0 || max_frames_per_utt == 0)
        return 0;
    if (cont_check(m) != 0)
        return -1;
    const size_t smem = (size_t)CT_F * m.blk * sizeof(float);
    dim3 grid((max_frames_per_utt + CT_F - 1) / CT_F, p.n_utts);
    SSB_CT_DISPATCH(cont_raw_union_kernel, grid, smem, m, p, feat, W, scratch)
    SSB_CUDA(cudaGetLastError());
    note_launch();
    dim3 g2(min(64, (max_frames_per_utt + 7) / 8), p.n_utts);
    cont_chain_finish_kernel<<<g2, 256, 0, st>>>(m, p, scratch, W, chain_scr);
    SSB_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// vtable path: x = one frame [blk] on the device, raw = scratch [n_sen]
int launch_cont_frame(const DevModel &m, const float *x, const uint16_t *act_sen, int n_act,
                      int compallsen, int16_t *raw, int16_t *senscr, cudaStream_t st)
{
    if (cont_check(m) != 0)
        return -1;
    const size_t smem = (size_t)CT_F * m.blk * sizeof(float);
    dim3 grid((m.n_sen + CT_THREADS - 1) / CT_THREADS, 1);
    SSB_CT_DISPATCH(cont_raw_dense_kernel, grid, smem, m, x, (int64_t)0, (int64_t)1, raw)
    SSB_CUDA(cudaGetLastError());
    note_launch();
    cont_frame_finish_kernel<<<1, 512, 0, st>>>(m.n_sen, raw, act_sen, n_act, compallsen, senscr);
    SSB_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

}  // namespace ssb
