// gmm_topn.cu -- K1: diagonal-Gaussian evaluation + stateful top-N selection.
//
// Replaces eval_topn / eval_cb / ptm_mgau_codebook_eval of the reference PTM
// scorer (ref: src/ptm_mgau.c:63-253).  Semantics kept bit for bit:
//   * distance d = det - sum_j ((x_j - mu_j)^2 * prec_j) in fp32, one rounding per
//     operation, dimensions in order, no fused multiply-add (ref :63-68, :106-127);
//   * the top-N list of frame t starts from frame t-1's codewords, re-scored and
//     insertion-sorted with strict '>' (ref :70-135);
//   * active codebooks are then scanned density by density; a density enters the
//     list when d >= (float)worst and it is not already listed, inserted in front
//     of equal scores (ref :139-225).  The reference's early-out is result-neutral.
//
// B200 mapping: one (codebook, stream) per CTA -- its 128 density records
// (13.8 KB for the bundled models) sit in shared memory and every lane reads the
// same record (broadcast LDS.128) -- and one *utterance* per thread, walking its
// frames in order with the top-N list in registers.  So the sequential
// dependence of the reference (frame t needs frame t-1's list) costs nothing,
// there are no cross-lane operations, and the FP32 pipe is the binding unit:
// 4 dependent-free FP32 ops per (density, dimension).
#include <cstdlib>
#include <cstring>

#include "device.cuh"

namespace ssb {

constexpr int K1_THREADS = 128;

__device__ __forceinline__ int32_t dist_to_int(float d)
{
    // (int32)d truncation toward zero, clamped below INT32_MIN (ref: src/ptm_mgau.c:128-131)
    return __float2int_rz(d);
}

// rec = [det, mean[L], prec[L], pad]; x = L feature values
template <int L>
__device__ __forceinline__ float gau_dist(const float *__restrict__ rec, const float (&x)[L])
{
    constexpr int RL = (1 + 2 * L + 3) & ~3;
    float v[RL];
    const float4 *r4 = reinterpret_cast<const float4 *>(rec);
#pragma unroll
    for (int i = 0; i < RL / 4; ++i) {
        float4 q = r4[i];
        v[4 * i] = q.x;
        v[4 * i + 1] = q.y;
        v[4 * i + 2] = q.z;
        v[4 * i + 3] = q.w;
    }
    float d = v[0];
#pragma unroll
    for (int j = 0; j < L; ++j) {
        float diff = __fsub_rn(x[j], v[1 + j]);
        float sq = __fmul_rn(diff, diff);
        float c = __fmul_rn(sq, v[1 + L + j]);
        d = __fsub_rn(d, c);
    }
    return d;
}

// Same, also returning the partial sum before the last dimension: the semi-continuous
// scorer's early-out tests the partial sums against the worst score (ref:
// src/s2_semi_mgau.c:131-143); they decrease monotonically, so the last one decides.
template <int L>
__device__ __forceinline__ float gau_dist2(const float *__restrict__ rec, const float (&x)[L],
                                           float &before_last)
{
    constexpr int RL = (1 + 2 * L + 3) & ~3;
    float v[RL];
    const float4 *r4 = reinterpret_cast<const float4 *>(rec);
#pragma unroll
    for (int i = 0; i < RL / 4; ++i) {
        float4 q = r4[i];
        v[4 * i] = q.x;
        v[4 * i + 1] = q.y;
        v[4 * i + 2] = q.z;
        v[4 * i + 3] = q.w;
    }
    float d = v[0];
    before_last = d;
#pragma unroll
    for (int j = 0; j < L; ++j) {
        if (j == L - 1)
            before_last = d;
        float diff = __fsub_rn(x[j], v[1 + j]);
        float sq = __fmul_rn(diff, diff);
        float c = __fmul_rn(sq, v[1 + L + j]);
        d = __fsub_rn(d, c);
    }
    return d;
}

__device__ __forceinline__ float gau_dist_rt2(const float *__restrict__ rec,
                                              const float *__restrict__ x, int L,
                                              float &before_last)
{
    float d = rec[0];
    before_last = d;
    for (int j = 0; j < L; ++j) {
        if (j == L - 1)
            before_last = d;
        float diff = __fsub_rn(__ldg(x + j), rec[1 + j]);
        float sq = __fmul_rn(diff, diff);
        float c = __fmul_rn(sq, rec[1 + L + j]);
        d = __fsub_rn(d, c);
    }
    return d;
}

// run-time length variant (models whose streams are not 13 wide)
__device__ __forceinline__ float gau_dist_rt(const float *__restrict__ rec,
                                             const float *__restrict__ x, int L)
{
    float d = rec[0];
    for (int j = 0; j < L; ++j) {
        float diff = __fsub_rn(__ldg(x + j), rec[1 + j]);
        float sq = __fmul_rn(diff, diff);
        float c = __fmul_rn(sq, rec[1 + L + j]);
        d = __fsub_rn(d, c);
    }
    return d;
}

template <int N>
struct TopN {
    int32_t s[N];
    int32_t c[N];
    __device__ __forceinline__ void reset()
    {
        // ref: src/ptm_mgau.c:694-720
#pragma unroll
        for (int k = 0; k < N; ++k) {
            s[k] = INT32_MIN;
            c[k] = k;
        }
    }
    // entry i just received score s[i]; move it up past strictly smaller scores (ref :70-84)
    __device__ __forceinline__ void settle(int i)
    {
        bool moving = true;
#pragma unroll
        for (int j = N - 2; j >= 0; --j) {
            if (j < i) {
                bool sw = moving && (s[j + 1] > s[j]);
                if (sw) {
                    int32_t ts = s[j], tc = c[j];
                    s[j] = s[j + 1];
                    c[j] = c[j + 1];
                    s[j + 1] = ts;
                    c[j + 1] = tc;
                }
                moving = sw;
            }
        }
    }
    // newcomer replaces the last entry and moves up past scores <= its own (ref :139-148)
    __device__ __forceinline__ void insert(int32_t id, int32_t cw)
    {
        s[N - 1] = id;
        c[N - 1] = cw;
        bool moving = true;
#pragma unroll
        for (int j = N - 2; j >= 0; --j) {
            bool sw = moving && (s[j + 1] >= s[j]);
            if (sw) {
                int32_t ts = s[j], tc = c[j];
                s[j] = s[j + 1];
                c[j] = c[j + 1];
                s[j + 1] = ts;
                c[j + 1] = tc;
            }
            moving = sw;
        }
    }
    __device__ __forceinline__ bool has(int32_t cw) const
    {
        bool h = false;
#pragma unroll
        for (int k = 0; k < N; ++k)
            h |= (c[k] == cw);
        return h;
    }
};

template <int N>
__device__ __forceinline__ void store_topn(const TopN<N> &tn, int4 *so, uchar4 *co)
{
    int4 sv;
    uchar4 cv;
    sv.x = tn.s[0];
    cv.x = (unsigned char)tn.c[0];
    sv.y = N > 1 ? tn.s[N > 1 ? 1 : 0] : INT32_MIN;
    cv.y = N > 1 ? (unsigned char)tn.c[N > 1 ? 1 : 0] : 0;
    sv.z = N > 2 ? tn.s[N > 2 ? 2 : 0] : INT32_MIN;
    cv.z = N > 2 ? (unsigned char)tn.c[N > 2 ? 2 : 0] : 0;
    sv.w = N > 3 ? tn.s[N > 3 ? 3 : 0] : INT32_MIN;
    cv.w = N > 3 ? (unsigned char)tn.c[N > 3 ? 3 : 0] : 0;
    *so = sv;
    *co = cv;
}

// grid: x = codebook*n_feat + stream, y = utterance group.  L = 0: run-time length.
template <int L, int N>
__global__ void __launch_bounds__(K1_THREADS)
gmm_topn_kernel(DevModel m, DevPlan p, const float *__restrict__ feat, int64_t G,
                int4 *__restrict__ out_s, uchar4 *__restrict__ out_c)
{
    extern __shared__ float4 smem4[];
    float *rec = reinterpret_cast<float *>(smem4);
    const int cs = blockIdx.x;
    const int cb = cs / m.n_feat, f = cs - cb * m.n_feat;
    const int RL = m.rec_len[f];
    const int ND = m.n_density;
    const int Lrt = m.featlen[f];
    {
        const float4 *src = reinterpret_cast<const float4 *>(m.gau + gau_offset(m, cb, f));
        for (int i = threadIdx.x; i < ND * RL / 4; i += blockDim.x)
            smem4[i] = src[i];
    }
    __syncthreads();
    const int u = blockIdx.y * blockDim.x + threadIdx.x;
    if (u >= p.n_utts)
        return;
    const int64_t g0 = p.frame_off[u];
    const int T = (int)(p.frame_off[u + 1] - g0);
    const float *xp = feat + g0 * m.blk + m.featoff[f];
    int4 *so = out_s + (int64_t)cs * G + g0;
    uchar4 *co = out_c + (int64_t)cs * G + g0;

    // active-codebook epochs of this utterance (the semi-continuous scorer has no notion of
    // an inactive codebook: it scans its only one on every frame)
    const bool semi = m.kind == SSB_SCORER_SEMI;
    int e = 0, e_end = 0, t_next = INT32_MAX;
    bool active = true;
    if (!p.all_active && !semi) {
        e = p.ep_off[u];
        e_end = p.ep_off[u + 1];
        active = false;
        t_next = e < e_end ? p.ep_start[e] : INT32_MAX;
    }

    TopN<N> tn;
    tn.reset();
    constexpr int LX = L > 0 ? L : 1;
    float x[LX], xn[LX];
    if (L > 0 && T > 0) {
#pragma unroll
        for (int j = 0; j < LX; ++j)
            xn[j] = __ldg(xp + j);
    }
    for (int t = 0; t < T; ++t) {
        const float *xt = xp + (int64_t)t * m.blk;
        if (L > 0) {
#pragma unroll
            for (int j = 0; j < LX; ++j)
                x[j] = xn[j];
            if (t + 1 < T) {
#pragma unroll
                for (int j = 0; j < LX; ++j)
                    xn[j] = __ldg(xt + m.blk + j);
            }
        }
        while (t >= t_next) {
            active = (p.ep_cbmask[(int64_t)e * 8 + (cb >> 5)] >> (cb & 31)) & 1u;
            ++e;
            t_next = e < e_end ? p.ep_start[e] : INT32_MAX;
        }
        // re-score last frame's codewords (ref: eval_topn)
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const float *r = rec + tn.c[i] * RL;
            float d;
            if (L > 0)
                d = gau_dist<LX>(r, x);
            else
                d = gau_dist_rt(r, xt, Lrt);
            tn.s[i] = dist_to_int(d);
            tn.settle(i);
        }
        // scan the codebook (ref: eval_cb)
        if (semi && (m.ds <= 1 || t % m.ds == 0)) {
            // ref: src/s2_semi_mgau.c:110-169 -- dropped when a partial sum falls below the
            // worst score (as floats), else the TRUNCATED score is compared as an integer
#pragma unroll 2
            for (int cw = 0; cw < ND; ++cw) {
                const float *r = rec + cw * RL;
                float d, dp;
                if (L > 0)
                    d = gau_dist2<LX>(r, x, dp);
                else
                    d = gau_dist_rt2(r, xt, Lrt, dp);
                if (dp < __int2float_rn(tn.s[N - 1]))
                    continue;
                const int32_t id = dist_to_int(d);
                if (id < tn.s[N - 1])
                    continue;
                if (tn.has(cw))
                    continue;
                tn.insert(id, cw);
            }
        } else if (active && (m.ds <= 1 || t % m.ds == 0)) {
#pragma unroll 2
            for (int cw = 0; cw < ND; ++cw) {
                const float *r = rec + cw * RL;
                float d;
                if (L > 0)
                    d = gau_dist<LX>(r, x);
                else
                    d = gau_dist_rt(r, xt, Lrt);
                if (d < __int2float_rn(tn.s[N - 1]))
                    continue;
                if (tn.has(cw))
                    continue;
                tn.insert(dist_to_int(d), cw);
            }
        }
        if (active || semi)
            store_topn<N>(tn, so + t, co + t);
    }
}

template <int L>
static int launch_k1_n(const DevModel &m, const DevPlan &p, const float *feat, int64_t G,
                       int4 *s, uchar4 *c, cudaStream_t st)
{
    int max_rl = 0;
    for (int f = 0; f < m.n_feat; ++f)
        max_rl = m.rec_len[f] > max_rl ? m.rec_len[f] : max_rl;
    size_t smem = (size_t)m.n_density * max_rl * sizeof(float);
    dim3 grid(m.n_mgau * m.n_feat, (p.n_utts + K1_THREADS - 1) / K1_THREADS);
#define SSB_K1(NN)                                                                          \
    case NN:                                                                                \
        SSB_DYN_SMEM((gmm_topn_kernel<L, NN>), smem); \
        gmm_topn_kernel<L, NN><<<grid, K1_THREADS, smem, st>>>(m, p, feat, G, s, c);        \
        break;
    switch (m.topn) {
        SSB_K1(1)
        SSB_K1(2)
        SSB_K1(3)
        SSB_K1(4)
    default:
        set_error("topn %d not supported (1..4)", m.topn);
        return -1;
    }
#undef SSB_K1
    SSB_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

int launch_gmm_topn(const DevModel &m, const DevPlan &p, const float *feat, int64_t n_frames,
                    int4 *tn_score, uchar4 *tn_cw, float *featp, cudaStream_t st)
{
    if (p.n_utts == 0 || n_frames == 0 || m.kind == SSB_SCORER_CONT)
        return 0;  // (the continuous scorer has no top-N stage: cont_score.cu)
    // default: tensor-core screening kernel (gmm_topn_tc.cu); SSB_K1=fp32 keeps the plain
    // CUDA-core scan below (same results, used for A/B timing and as the generic-shape path)
    const char *force = getenv("SSB_K1");
    if (tc_supported(m) && m.ds <= 1 && featp != nullptr && !(force && strcmp(force, "fp32") == 0))
        return launch_gmm_topn_tc(m, p, feat, n_frames, tn_score, tn_cw, featp, nullptr, nullptr,
                                  nullptr, st);
    bool all13 = true;
    for (int f = 0; f < m.n_feat; ++f)
        all13 = all13 && m.featlen[f] == 13;
    if (all13)
        return launch_k1_n<13>(m, p, feat, n_frames, tn_score, tn_cw, st);
    return launch_k1_n<0>(m, p, feat, n_frames, tn_score, tn_cw, st);
}

// ------------------------------------------------------------------------
// Single-frame variant behind the mgau vtable (one frame_eval call = one frame).
// grid = codebook*stream, one thread per density; thread 0 replays the
// reference's sequential insertion over the distances held in shared memory.
// History lives in HBM as two slots, exactly like ptm_mgau_t.hist (ref :803-811).
// ------------------------------------------------------------------------
__global__ void frame_topn_kernel(DevModel m, FrameHist h, int slot, int prev,
                                  const float *__restrict__ x, int do_scan)
{
    extern __shared__ float dist[];
    const int cs = blockIdx.x;
    const int cb = cs / m.n_feat, f = cs - cb * m.n_feat;
    const int RL = m.rec_len[f], L = m.featlen[f], ND = m.n_density, N = m.topn;
    const float *rec = m.gau + gau_offset(m, cb, f);
    const float *xf = x + m.featoff[f];
    const bool semi = m.kind == SSB_SCORER_SEMI;
    float *part = dist + ND;  // semi only: partial sums before the last dimension
    for (int cw = threadIdx.x; cw < ND; cw += blockDim.x) {
        if (semi) {
            float dp;
            dist[cw] = gau_dist_rt2(rec + (int64_t)cw * RL, xf, L, dp);
            part[cw] = dp;
        } else
            dist[cw] = gau_dist_rt(rec + (int64_t)cw * RL, xf, L);
    }
    __syncthreads();
    if (threadIdx.x != 0)
        return;
    int4 ps = h.score[prev][cs];
    uchar4 pc = h.cw[prev][cs];
    int32_t s[4] = {ps.x, ps.y, ps.z, ps.w};
    int32_t c[4] = {pc.x, pc.y, pc.z, pc.w};
    for (int i = 0; i < N; ++i) {
        int32_t cw = c[i], v = dist_to_int(dist[cw]);
        int j;
        for (j = i - 1; j >= 0 && v > s[j]; --j) {
            s[j + 1] = s[j];
            c[j + 1] = c[j];
        }
        s[j + 1] = v;
        c[j + 1] = cw;
    }
    if (do_scan && (semi || h.act[slot][cb])) {
        for (int cw = 0; cw < ND; ++cw) {
            float d = dist[cw];
            if (semi) {  // ref: src/s2_semi_mgau.c:131-152
                if (part[cw] < __int2float_rn(s[N - 1]) || dist_to_int(d) < s[N - 1])
                    continue;
            } else if (d < __int2float_rn(s[N - 1]))
                continue;
            int i;
            for (i = 0; i < N; ++i)
                if (c[i] == cw)
                    break;
            if (i < N)
                continue;
            int32_t id = dist_to_int(d);
            for (i = N - 2; i >= 0 && id >= s[i]; --i) {
                s[i + 1] = s[i];
                c[i + 1] = c[i];
            }
            s[i + 1] = id;
            c[i + 1] = cw;
        }
    }
    h.score[slot][cs] = make_int4(s[0], s[1], s[2], s[3]);
    h.cw[slot][cs] = make_uchar4((unsigned char)c[0], (unsigned char)c[1], (unsigned char)c[2],
                                 (unsigned char)c[3]);
}

int launch_frame_topn(const DevModel &m, const FrameHist &h, int slot, int prev, const float *x,
                      int do_scan, cudaStream_t st)
{
    int threads = m.n_density < 128 ? 128 : 256;
    frame_topn_kernel<<<m.n_mgau * m.n_feat, threads, 2 * m.n_density * sizeof(float), st>>>(
        m, h, slot, prev, x, do_scan);
    SSB_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

}  // namespace ssb
