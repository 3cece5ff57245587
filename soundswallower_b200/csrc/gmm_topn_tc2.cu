// gmm_topn_tc2.cu -- K1, tensor-core version 2: stateless 3xTF32 screening + exact survivors.
// Same results, bit for bit, as gmm_topn.cu / the reference's eval_topn + eval_cb
// (ref: src/ptm_mgau.c:63-253).
//
// Measured on the reference's test audio (tools/probe_stats.py): the top-N SET of a
// codebook-stream changes on 93 % of the frames, so last frame's list is a poor threshold
// (36 survivors on average) and re-scoring it every frame is wasted work.  This kernel takes
// the threshold from the current frame's own screening scores:
//   * 3xTF32 split GEMM (A_hi B_hi + A_lo B_hi + A_hi B_lo, accumulators in TMEM):
//     |approx - exact| <= eps ~ 10 raw units (1/100 of an output unit) -- measured worst case
//     2^-20.4 of the term magnitudes, bound used 2^-18; densities with outlier precisions
//     ("hot", e.g. floored variances) get their own, larger bound so they do not inflate
//     everybody else's;
//   * L = N-th largest of the 8 group maxima (16 columns each, regular densities only): at
//     least N densities have approx >= L, hence the N-th best EXACT score is >= L - eps;
//   * survivors = {approx >= L - 2 eps - 2}: every other density is more than 2 raw units
//     below the N-th best exact score, i.e. strictly below it after the (int) truncation;
//   * survivors are evaluated with the reference's own fp32 operation order.  If the N+1
//     best integer scores are pairwise distinct, the reference's final list is "the N best,
//     sorted" whatever list it carried in (its insertion rules only matter under ties):
//     accept.  Otherwise (a tie, ~2e-4 of the steps) replay the reference literally: catch
//     the carried list up through the frames since it was last exact, then eval_topn +
//     eval_cb over all 128 densities (slow path).
// Frames on which no utterance of the CTA scans this codebook are skipped wholesale (the CTA
// jumps to the next frame where the active set of one of its utterances grows).
//
// CTA = one (codebook, stream) x 256 utterances: two 128-row MMA groups sharing B and the
// exact records; thread = utterance = MMA row = TMEM lane.  Per frame step 2 x 12 tcgen05.mma
// (M=128, N=128, K=8, kind::tf32).  2 CTAs per SM use the 512 TMEM columns.  Features are
// re-packed once per batch to [frame][stream][16] so that a thread's 13 values are four
// aligned 16-byte loads.
#include "tc_common.cuh"

namespace ssb {

constexpr int TC2_THREADS = 256;
constexpr int TC2_GROUPS = TC2_THREADS / 128;
constexpr int TC2_XP = 16;  // packed floats per (frame, stream)

struct Tc2Smem {
    float Bhi[TC_ND * TC_K];                 // 16 KB, swizzled
    float Blo[TC_ND * TC_K];                 // 16 KB, swizzled
    float A[TC2_GROUPS][2][128 * TC_K];      // per row group: hi, lo tiles (16 KB each)
    float rec[TC_ND * TC_RL];                // 14 KB exact records
    float aux[SSB_TC_AUX];
    uint32_t hot[4];
    uint64_t mbar[TC2_GROUPS];  // the two row groups run out of phase: own barriers, own MMAs
    uint32_t tmem_base;
    int tmax[TC2_GROUPS];
    int jump[TC2_GROUPS];
};

// named barrier of one 128-thread row group (ids 1, 2; 0 is __syncthreads)
__device__ __forceinline__ void group_sync(int grp)
{
    asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
}
__device__ __forceinline__ int group_sync_or(int grp, int pred)
{
    uint32_t r;
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "setp.ne.b32 q, %1, 0;\n\t"
        "bar.red.or.pred p, %2, 128, q;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(r)
        : "r"(pred), "r"(grp + 1)
        : "memory");
    return (int)r;
}

__global__ void pack_features_kernel(DevModel m, const float *__restrict__ feat, int64_t G,
                                     float *__restrict__ featp)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = G * m.n_feat * TC2_XP;
    if (i >= n)
        return;
    const int k = (int)(i % TC2_XP);
    const int64_t gf = i / TC2_XP;
    const int f = (int)(gf % m.n_feat);
    const int64_t g = gf / m.n_feat;
    featp[i] = k < m.featlen[f] ? feat[g * m.blk + m.featoff[f] + k] : 0.f;
}

__device__ __forceinline__ float max3(float a, float b, float c)
{
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// max of v[o .. o+16), skipping the columns whose bit is set in `exclude`
__device__ __forceinline__ float max16(const float (&v)[32], int o, uint32_t exclude)
{
    float w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i)
        w[i] = ((exclude >> i) & 1u) ? -3.4028235e38f : v[o + i];
    float m0 = max3(w[0], w[1], w[2]), m1 = max3(w[3], w[4], w[5]), m2 = max3(w[6], w[7], w[8]);
    float m3 = max3(w[9], w[10], w[11]), m4 = max3(w[12], w[13], w[14]);
    return fmaxf(max3(m0, m1, m2), max3(m3, m4, w[15]));
}

__device__ __forceinline__ float max16_plain(const float (&v)[32], int o)
{
    float m0 = max3(v[o], v[o + 1], v[o + 2]), m1 = max3(v[o + 3], v[o + 4], v[o + 5]);
    float m2 = max3(v[o + 6], v[o + 7], v[o + 8]), m3 = max3(v[o + 9], v[o + 10], v[o + 11]);
    float m4 = max3(v[o + 12], v[o + 13], v[o + 14]);
    return fmaxf(max3(m0, m1, m2), max3(m3, m4, v[o + 15]));
}

#define TC_CE(a, b)              \
    {                            \
        float hi_ = fmaxf(a, b); \
        b = fminf(a, b);         \
        a = hi_;                 \
    }

// k-th largest (k = 1..4) of 8 values: 19-comparator sorting network, descending
__device__ __forceinline__ float kth_largest8(float (&g)[8], int k)
{
    TC_CE(g[0], g[1]) TC_CE(g[2], g[3]) TC_CE(g[4], g[5]) TC_CE(g[6], g[7])
    TC_CE(g[0], g[2]) TC_CE(g[1], g[3]) TC_CE(g[4], g[6]) TC_CE(g[5], g[7])
    TC_CE(g[1], g[2]) TC_CE(g[5], g[6]) TC_CE(g[0], g[4]) TC_CE(g[3], g[7])
    TC_CE(g[1], g[5]) TC_CE(g[2], g[6])
    TC_CE(g[1], g[4]) TC_CE(g[3], g[6])
    TC_CE(g[2], g[4]) TC_CE(g[3], g[5])
    TC_CE(g[3], g[4])
    return k == 1 ? g[0] : (k == 2 ? g[1] : (k == 3 ? g[2] : g[3]));
}

__device__ __forceinline__ float4 lds128(uint32_t saddr)
{
    float4 q;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(saddr));
    return q;
}

// exact distance with the record read by explicit shared-memory loads (ref: src/ptm_mgau.c:63-68)
__device__ __forceinline__ float exact_dist_s(uint32_t rec_saddr, const float (&x)[TC_L])
{
    float v[TC_RL];
#pragma unroll
    for (int i = 0; i < TC_RL / 4; ++i) {
        const float4 q = lds128(rec_saddr + 16u * i);
        v[4 * i] = q.x;
        v[4 * i + 1] = q.y;
        v[4 * i + 2] = q.z;
        v[4 * i + 3] = q.w;
    }
    float d = v[0];
#pragma unroll
    for (int j = 0; j < TC_L; ++j) {
        const float diff = __fsub_rn(x[j], v[1 + j]);
        const float sq = __fmul_rn(diff, diff);
        const float c = __fmul_rn(sq, v[1 + TC_L + j]);
        d = __fsub_rn(d, c);
    }
    return d;
}

__device__ __forceinline__ void load_x(const float *__restrict__ p, float4 (&q)[4])
{
    const float4 *p4 = reinterpret_cast<const float4 *>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        q[i] = __ldg(p4 + i);
}

// the same 13 values straight from the caller's [frame][sum featlen] rows (4-byte aligned only)
__device__ __forceinline__ void load_x_raw(const float *__restrict__ p, float4 (&q)[4])
{
    q[0] = make_float4(__ldg(p + 0), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
    q[1] = make_float4(__ldg(p + 4), __ldg(p + 5), __ldg(p + 6), __ldg(p + 7));
    q[2] = make_float4(__ldg(p + 8), __ldg(p + 9), __ldg(p + 10), __ldg(p + 11));
    q[3] = make_float4(__ldg(p + 12), 0.f, 0.f, 0.f);
}

__device__ __forceinline__ void unpack_x(const float4 (&q)[4], float (&x)[TC_L])
{
    x[0] = q[0].x; x[1] = q[0].y; x[2] = q[0].z; x[3] = q[0].w;
    x[4] = q[1].x; x[5] = q[1].y; x[6] = q[1].z; x[7] = q[1].w;
    x[8] = q[2].x; x[9] = q[2].y; x[10] = q[2].z; x[11] = q[2].w;
    x[12] = q[3].x;
}

// INIT: start from the lists a previous pass left (DevPlan.init_topn) -- a separate
// instantiation, because the kernel sits exactly at its 128-register budget
template <int N, bool DEBUG, bool RAW, bool INIT>
__global__ void __launch_bounds__(TC2_THREADS, 2)
gmm_topn_tc2_kernel(DevModel m, DevPlan p, const float *__restrict__ featp, int64_t G,
                    int4 *__restrict__ out_s, uchar4 *__restrict__ out_c, TcDebug dbg)
{
    extern __shared__ uint8_t smem_raw[];
    Tc2Smem &S = *reinterpret_cast<Tc2Smem *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, grp = tid >> 7, row = tid & 127;
    const int cs = blockIdx.x;
    const int cb = cs / m.n_feat, f = cs - cb * m.n_feat;

    {
        const float4 *shi = reinterpret_cast<const float4 *>(m.gB + (size_t)cs * TC_ND * TC_K);
        const float4 *slo = reinterpret_cast<const float4 *>(m.gBlo + (size_t)cs * TC_ND * TC_K);
        float4 *dhi = reinterpret_cast<float4 *>(S.Bhi), *dlo = reinterpret_cast<float4 *>(S.Blo);
        for (int i = tid; i < TC_ND * TC_K / 4; i += TC2_THREADS) {
            const int n = i >> 3, j = i & 7;
            const int o = (n >> 3) * 64 + (n & 7) * 8 + (j ^ (n & 7));
            dhi[o] = shi[i];
            dlo[o] = slo[i];
        }
        const float4 *rsrc = reinterpret_cast<const float4 *>(m.gau + gau_offset(m, cb, f));
        float4 *rdst = reinterpret_cast<float4 *>(S.rec);
        for (int i = tid; i < TC_ND * TC_RL / 4; i += TC2_THREADS)
            rdst[i] = rsrc[i];
        if (tid < SSB_TC_AUX)
            S.aux[tid] = m.gAux[(size_t)cs * SSB_TC_AUX + tid];
        if (tid < 4)
            S.hot[tid] = m.gHot[(size_t)cs * 4 + tid];
        float4 *dA = reinterpret_cast<float4 *>(&S.A[0][0][0]);
        for (int i = tid; i < TC2_GROUPS * 2 * 128 * TC_K / 4; i += TC2_THREADS)
            dA[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid == 0) {
            for (int g = 0; g < TC2_GROUPS; ++g) {
                mbar_init(&S.mbar[g], 1);
                S.tmax[g] = 0;
                S.jump[g] = INT32_MAX;
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)),
                     "r"(128 * TC2_GROUPS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    const int u = blockIdx.y * TC2_THREADS + tid;
    const bool has_utt = u < p.n_utts;
    const int64_t g0 = has_utt ? p.frame_off[u] : 0;
    const int T = has_utt ? (int)(p.frame_off[u + 1] - g0) : 0;
    atomicMax(&S.tmax[grp], T);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;
    const int Tmax = S.tmax[grp];
    const bool has_hot = (S.hot[0] | S.hot[1] | S.hot[2] | S.hot[3]) != 0u;
    // TMEM: lanes = rows of the group (a warp may only touch lanes 32*(warp%4)..+31), columns
    // [128*grp, 128*grp+128) = the group's accumulator
    const uint32_t tmem_row = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(grp * 128);
    const uint32_t rec_s = smem_u32(S.rec);

    // RAW: featp is the caller's feature array itself (no re-packing pass)
    const int64_t xstride = RAW ? (int64_t)m.blk : (int64_t)m.n_feat * TC2_XP;
    const float *xp = RAW ? featp + g0 * m.blk + m.featoff[f] : featp + (g0 * m.n_feat + f) * TC2_XP;
    int4 *so = out_s + (int64_t)cs * G + g0;
    uchar4 *co = out_c + (int64_t)cs * G + g0;
    float4 *myAhi = reinterpret_cast<float4 *>(&S.A[grp][0][0]) + (row >> 3) * 64 + (row & 7) * 8;
    float4 *myAlo = reinterpret_cast<float4 *>(&S.A[grp][1][0]) + (row >> 3) * 64 + (row & 7) * 8;
    const int swz = row & 7;

    int e = 0, e_end = 0, t_next = INT32_MAX;
    bool active = true;
    if (!p.all_active && has_utt) {
        e = p.ep_off[u];
        e_end = p.ep_off[u + 1];
        active = false;
        t_next = e < e_end ? p.ep_start[e] : INT32_MAX;
    }

    TcTopN<N> tn;
    tn.reset();
    int t_last = -1;  // last frame on which tn was the reference's exact list
    float4 xn[4];
    int xn_t = -1;    // frame held in xn
    unsigned long long n_exact = 0, n_steps = 0, n_slow = 0;
    uint32_t n_mma = 0;

    int t = 0;
    while (t < Tmax) {
        const bool live = t < T;
        if (live) {
            while (t >= t_next) {
                active = (p.ep_cbmask[(int64_t)e * 8 + (cb >> 5)] >> (cb & 31)) & 1u;
                ++e;
                t_next = e < e_end ? p.ep_start[e] : INT32_MAX;
            }
        }
        const bool scan = live && active;
        float x[TC_L];
        float eps = 0.f, eps_hot = 0.f;
        if (scan) {
            if (xn_t != t) {
                if (RAW)
                    load_x_raw(xp + (int64_t)t * xstride, xn);
                else
                    load_x(xp + (int64_t)t * xstride, xn);
            }
            unpack_x(xn, x);
            // A rows (hi, lo): [x' (13), x'^2 (13), 1, 1, 0 x4] / [x'_lo, x'^2_lo, 0 ...]
            float hi[TC_K], lo[TC_K];
            float acc = S.aux[39], acc_hot = S.aux[39];
#pragma unroll
            for (int j = 0; j < TC_L; ++j) {
                const float xc = __fsub_rn(x[j], S.aux[j]);
                const float sq = __fmul_rn(xc, xc);
                hi[j] = to_tf32(xc);
                lo[j] = to_tf32(__fsub_rn(xc, hi[j]));
                hi[TC_L + j] = to_tf32(sq);
                lo[TC_L + j] = to_tf32(__fsub_rn(sq, hi[TC_L + j]));
                acc = fmaf(fabsf(xc), S.aux[40 + j], acc);
                acc = fmaf(sq, S.aux[53 + j], acc);
                if (has_hot) {
                    acc_hot = fmaf(fabsf(xc), S.aux[67 + j], acc_hot);
                    acc_hot = fmaf(sq, S.aux[80 + j], acc_hot);
                }
            }
            hi[26] = 1.f;
            hi[27] = 1.f;
            lo[26] = 0.f;
            lo[27] = 0.f;
#pragma unroll
            for (int j = 28; j < TC_K; ++j) {
                hi[j] = 0.f;
                lo[j] = 0.f;
            }
            // split-TF32 products are good to ~2^-20; the accumulation in the tensor core and the
            // reference's own fp32 chain add a few 2^-24 each.  Measured worst case on the test
            // audio: 2^-20.4 of the term magnitudes; the bound used is 2^-18 (+4 raw units).
            eps = fmaf(acc, 1.f / 262144.f, 4.f);
            eps_hot = fmaf(acc_hot, 1.f / 262144.f, 4.f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                myAhi[j ^ swz] = make_float4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                myAlo[j ^ swz] = make_float4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
            }
        }
        fence_async_proxy();
        tc_fence_before();
        const int any = group_sync_or(grp, scan ? 1 : 0);
        if (!any) {
            // nobody in this row group scans this codebook on frame t: jump to the next frame
            // on which the active set of one of its utterances grows
            const int cand = (live && t_next < T) ? t_next : INT32_MAX;
            if (cand != INT32_MAX)
                atomicMin(&S.jump[grp], cand);
            group_sync(grp);
            const int tj = S.jump[grp];
            group_sync(grp);
            if (row == 0)
                S.jump[grp] = INT32_MAX;
            if (tj == INT32_MAX)
                break;
            t = tj;
            continue;
        }
        if (row == 0) {
            tc_fence_after();
            const uint64_t bhi = umma_desc_sw128(smem_u32(S.Bhi)), blo = umma_desc_sw128(smem_u32(S.Blo));
            const uint64_t ahi = umma_desc_sw128(smem_u32(&S.A[grp][0][0]));
            const uint64_t alo = umma_desc_sw128(smem_u32(&S.A[grp][1][0]));
            const uint32_t d = tmem + (uint32_t)(grp * 128);
#pragma unroll
            for (int k = 0; k < TC_K / 8; ++k)
                umma_tf32(d, ahi + (uint64_t)(2 * k), bhi + (uint64_t)(2 * k), k > 0 ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < TC_K / 8; ++k)
                umma_tf32(d, alo + (uint64_t)(2 * k), bhi + (uint64_t)(2 * k), 1u);
#pragma unroll
            for (int k = 0; k < TC_K / 8; ++k)
                umma_tf32(d, ahi + (uint64_t)(2 * k), blo + (uint64_t)(2 * k), 1u);
            umma_commit(&S.mbar[grp]);
        }
        // next frame's features travel while the tensor core works
        if (scan && t + 1 < T) {
            if (RAW)
                load_x_raw(xp + (int64_t)(t + 1) * xstride, xn);
            else
                load_x(xp + (int64_t)(t + 1) * xstride, xn);
            xn_t = t + 1;
        }
        mbar_wait(&S.mbar[grp], n_mma & 1u);
        ++n_mma;
        tc_fence_after();
        // pass 1: N-th largest of the 8 group maxima (regular densities only, so that the N
        // witnesses carry the regular bound)
        float gm[8];
#pragma unroll 1
        for (int ch = 0; ch < TC_ND / 32; ++ch) {
            float v[32];
            tmem_ld32(tmem_row + (uint32_t)(ch * 32), v);
            float a, b;
            if (has_hot) {  // block-uniform: only the few codebook-streams with hot densities
                const uint32_t hot = S.hot[ch];
                a = max16(v, 0, hot & 0xffffu);
                b = max16(v, 16, hot >> 16);
            } else {
                a = max16_plain(v, 0);
                b = max16_plain(v, 16);
            }
            // gm[2*ch], gm[2*ch+1] without dynamic register indexing
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (q == ch) {
                    gm[2 * q] = a;
                    gm[2 * q + 1] = b;
                }
        }
        const float L = kth_largest8(gm, N);
        // regular density n survives iff approx_n >= L - 2 eps - 2; a hot one iff
        // approx_n >= L - eps - eps_hot - 2 (its own error bound on its side of the inequality)
        const float nthr = -(L - 2.f * eps - 2.f) * TC_BIG;
        const float nthr_hot = -(L - eps - eps_hot - 2.f) * TC_BIG;
        // pass 2: survivors, evaluated exactly; keep the N+1 best integer scores
        TcTopN<N + 1> best;
#pragma unroll
        for (int k = 0; k <= N; ++k) {
            best.s[k] = INT32_MIN;
            best.c[k] = -1;
        }
        int cnt = 0;
        // survivor bitmasks of the four 32-column chunks first, then ONE loop over all
        // survivors: the warp iterates max-over-lanes(total survivors) times instead of the
        // sum over chunks of the per-chunk maxima
        uint32_t mk0 = 0u, mk1 = 0u, mk2 = 0u, mk3 = 0u;
#pragma unroll 1
        for (int ch = 0; ch < TC_ND / 32; ++ch) {
            float v[32];
            tmem_ld32(tmem_row + (uint32_t)(ch * 32), v);
            if (DEBUG && scan && dbg.approx) {
                float *o = dbg.approx + (((int64_t)cs * G + g0 + t) * TC_ND + ch * 32);
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    o[i] = v[i];
                if (ch == 0) {
                    dbg.eps[((int64_t)cs * G + g0 + t) * 2] = eps;
                    dbg.eps[((int64_t)cs * G + g0 + t) * 2 + 1] = eps_hot;
                }
            }
            if (scan) {
                // s = sat((approx - thr) * 2^96) is exactly 0 or 1; bit i <-> column i
                float m0 = 0.f, m1 = 0.f;
#pragma unroll
                for (int i = 15; i >= 0; --i) {
                    m0 = fmaf(m0, 2.f, __saturatef(fmaf(v[i], TC_BIG, nthr)));
                    m1 = fmaf(m1, 2.f, __saturatef(fmaf(v[16 + i], TC_BIG, nthr)));
                }
                uint32_t mask = (uint32_t)m0 | ((uint32_t)m1 << 16);
                const uint32_t hot = S.hot[ch];
                if (hot) {  // block-uniform
                    float h0 = 0.f, h1 = 0.f;
#pragma unroll
                    for (int i = 15; i >= 0; --i) {
                        h0 = fmaf(h0, 2.f, __saturatef(fmaf(v[i], TC_BIG, nthr_hot)));
                        h1 = fmaf(h1, 2.f, __saturatef(fmaf(v[16 + i], TC_BIG, nthr_hot)));
                    }
                    mask = (mask & ~hot) | (((uint32_t)h0 | ((uint32_t)h1 << 16)) & hot);
                }
                mk0 = ch == 0 ? mask : mk0;
                mk1 = ch == 1 ? mask : mk1;
                mk2 = ch == 2 ? mask : mk2;
                mk3 = ch == 3 ? mask : mk3;
            }
        }
        {
            unsigned long long lo = (unsigned long long)mk0 | ((unsigned long long)mk1 << 32);
            unsigned long long hi = (unsigned long long)mk2 | ((unsigned long long)mk3 << 32);
            while (lo | hi) {  // ascending density index, as the chunked loop visited them
                int cw;
                if (lo) {
                    cw = __ffsll((long long)lo) - 1;
                    lo &= lo - 1;
                } else {
                    cw = 63 + __ffsll((long long)hi);
                    hi &= hi - 1;
                }
                const int32_t sc = __float2int_rz(exact_dist_s(rec_s + (uint32_t)(cw * TC_RL * 4), x));
                ++cnt;
                if (sc >= best.s[N])
                    best.insert(sc, cw);
            }
        }
        tc_fence_before();
        if (scan) {
            if (DEBUG) {
                ++n_steps;
                n_exact += cnt;
            }
            bool distinct = cnt >= N;
#pragma unroll
            for (int k = 0; k < N; ++k)
                distinct = distinct && (best.s[k] > best.s[k + 1]);
            if (distinct) {
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    tn.s[k] = best.s[k];
                    tn.c[k] = best.c[k];
                }
            } else {
                // The reference's own procedure, literally.  First catch the carried list up
                // through the frames since it was last exact (they did not scan this codebook:
                // re-score + stable sort only), then eval_topn + eval_cb on frame t.
                if (DEBUG)
                    ++n_slow;
                if (p.tie_bits)
                    atomicOr(&p.tie_bits[(int64_t)cs * p.tie_w + ((g0 + t) >> 5)], 1u << ((g0 + t) & 31));
                if (INIT && t_last < 0) {
                    // no exact step yet: the carried list is still the one a previous pass left
                    // (loaded here, on the cold path: the kernel sits at its register budget)
                    const uchar4 c = p.init_topn[(int64_t)u * gridDim.x + cs];
                    const int cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                    for (int k = 0; k < N; ++k)
                        tn.c[k] = cc[k];
                }
#pragma unroll 1
                for (int tt = t_last + 1; tt <= t; ++tt) {
                    float xx[TC_L];
                    if (tt < t) {
                        float4 q[4];
                        if (RAW)
                            load_x_raw(xp + (int64_t)tt * xstride, q);
                        else
                            load_x(xp + (int64_t)tt * xstride, q);
                        unpack_x(q, xx);
                    } else {
#pragma unroll
                        for (int j = 0; j < TC_L; ++j)
                            xx[j] = x[j];
                    }
#pragma unroll 1
                    for (int i = 0; i < N; ++i) {
                        int32_t ci = tn.c[0];
#pragma unroll
                        for (int k = 1; k < N; ++k)
                            if (k == i)
                                ci = tn.c[k];
                        const int32_t sc = __float2int_rz(exact_dist_s(rec_s + (uint32_t)(ci * TC_RL * 4), xx));
#pragma unroll
                        for (int k = 0; k < N; ++k)
                            if (k == i)
                                tn.s[k] = sc;
                        tn.settle(i);
                    }
                }
#pragma unroll 1
                for (int cw = 0; cw < TC_ND; ++cw) {
                    const float d = exact_dist_s(rec_s + (uint32_t)(cw * TC_RL * 4), x);
                    if (d < __int2float_rn(tn.s[N - 1]))
                        continue;
                    if (tn.has(cw))
                        continue;
                    tn.insert(__float2int_rz(d), cw);
                }
            }
            t_last = t;
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
            so[t] = sv;
            co[t] = cv;
        }
        ++t;
    }
    if (DEBUG && dbg.counters) {
        atomicAdd(&dbg.counters[0], n_exact);
        atomicAdd(&dbg.counters[1], n_steps);
        atomicAdd(&dbg.counters[2], n_slow);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128 * TC2_GROUPS));
}

int launch_gmm_topn_tc2(const DevModel &m, const DevPlan &p, const float *feat, int64_t n_frames,
                        int4 *tn_score, uchar4 *tn_cw, float *featp, TcDebug dbg, cudaStream_t st)
{
    const bool debug = dbg.approx != nullptr || dbg.counters != nullptr;
    // SSB_K1_PACK=1 keeps the re-packing pass ([frame][stream][16], four aligned 16-byte loads
    // per step); the default reads the caller's rows directly (13 scalar loads per step, one
    // launch and 786 MB of scratch traffic less)
    const char *pk = getenv("SSB_K1_PACK");
    const bool raw = !(pk && *pk == '1');
    if (!raw) {
        const int64_t n = n_frames * m.n_feat * TC2_XP;
        pack_features_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m, feat, n_frames, featp);
        SSB_CUDA(cudaGetLastError());
        note_launch();
    }
    const float *src = raw ? feat : featp;
    const size_t smem = sizeof(Tc2Smem) + 1024;
    dim3 grid(m.n_mgau * m.n_feat, (p.n_utts + TC2_THREADS - 1) / TC2_THREADS);
#define SSB_TC2_(NN, DBG, RW, IN)                                                                 \
    do {                                                                                          \
        SSB_DYN_SMEM((gmm_topn_tc2_kernel<NN, DBG, RW, IN>), smem);                               \
        gmm_topn_tc2_kernel<NN, DBG, RW, IN><<<grid, TC2_THREADS, smem, st>>>(m, p, src, n_frames, \
                                                                              tn_score, tn_cw, dbg); \
    } while (0)
#define SSB_TC2(NN, DBG)                                                                          \
    do {                                                                                          \
        if (raw && !p.init_topn) SSB_TC2_(NN, DBG, true, false);                                  \
        else if (raw) SSB_TC2_(NN, DBG, true, true);                                              \
        else if (!p.init_topn) SSB_TC2_(NN, DBG, false, false);                                   \
        else SSB_TC2_(NN, DBG, false, true);                                                      \
    } while (0)
    switch (m.topn) {
    case 1:
        if (debug) SSB_TC2(1, true); else SSB_TC2(1, false);
        break;
    case 2:
        if (debug) SSB_TC2(2, true); else SSB_TC2(2, false);
        break;
    case 3:
        if (debug) SSB_TC2(3, true); else SSB_TC2(3, false);
        break;
    case 4:
        if (debug) SSB_TC2(4, true); else SSB_TC2(4, false);
        break;
    default:
        set_error("topn %d not supported (1..4)", m.topn);
        return -1;
    }
#undef SSB_TC2
#undef SSB_TC2_
    SSB_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

bool tc_supported(const DevModel &m)
{
    if (m.n_density != TC_ND || m.gB == nullptr || m.kind != SSB_SCORER_PTM)
        return false;
    for (int f = 0; f < m.n_feat; ++f)
        if (m.featlen[f] != TC_L)
            return false;
    return true;
}

// dispatcher kept from the time when two tensor-core kernels existed (the v1 kernel, single
// TF32 + carried-list threshold, is gone: 61 ms against 27 ms on config #2)
int launch_gmm_topn_tc(const DevModel &m, const DevPlan &p, const float *feat, int64_t n_frames,
                       int4 *tn_score, uchar4 *tn_cw, float *featp, float *dbg_approx,
                       float *dbg_eps, unsigned long long *dbg_counters, cudaStream_t st)
{
    if (p.n_utts == 0 || n_frames == 0)
        return 0;
    if (!tc_supported(m) || m.ds > 1 || featp == nullptr) {
        set_error("tensor-core top-N needs 128 densities, 13-wide streams and ds = 1");
        return -1;
    }
    TcDebug dbg{dbg_approx, dbg_eps, dbg_counters};
    return launch_gmm_topn_tc2(m, p, feat, n_frames, tn_score, tn_cw, featp, dbg, st);
}

size_t tc2_featp_bytes(const DevModel &m, int64_t n_frames)
{
    const char *pk = getenv("SSB_K1_PACK");
    if (!(pk && *pk == '1'))
        return 16;  // the kernel reads the caller's rows directly; the pointer only has to exist
    return (size_t)n_frames * m.n_feat * TC2_XP * sizeof(float);
}

}  // namespace ssb
