// gmm_topn_tc.cu -- K1 on the tensor cores: tcgen05 TF32 screening GEMM + exact FP32
// re-scoring of the survivors.  Same results, bit for bit, as gmm_topn.cu (and therefore as
// eval_topn / eval_cb of the reference, ref: src/ptm_mgau.c:63-253).
//
// Idea.  The reference scans all 128 densities of a codebook-stream every frame to find the
// top-N.  Almost all of that work only proves that a density is NOT in the top-N.  That proof
// does not need the reference's arithmetic: any approximation with a rigorous error bound
// will do.  So per (utterance, frame, codebook-stream):
//   1. tensor cores compute approx[n] = sum_k A[k] * B[n][k] for the 128 densities, with
//        A = [x', x'^2, 1, 1, 0...]           (x' = x - centre of the codebook, TF32)
//        B[n] = [2 mu' v, -v, c_hi, c_lo, 0...]  (c = det - sum mu'^2 v; TF32, packed at load)
//      = det - sum (x - mu)^2 v  up to  eps = O(2^-10 * sum |terms|), bounded per row;
//   2. the carried top-N codewords are re-scored exactly (the reference's eval_topn);
//   3. a density can only enter the list if its exact score d >= (float)worst, hence only
//      if approx >= worst - eps: those few survivors are evaluated exactly, in codeword
//      order, with the reference's own insertion rules.  Everything else is provably a
//      "continue" of the reference's loop.
//
// Mapping: CTA = one (codebook, stream) x 128 utterances; thread = utterance = MMA row =
// TMEM lane, walking the frames in order (top-N carry in registers).  Per frame step one
// 128x128x32 TF32 MMA (4 tcgen05.mma, K = 8 each): A rows written by their own threads into
// the canonical K-major SWIZZLE_128B layout, B resident in shared memory, D in TMEM (128
// columns), read back with tcgen05.ld 32x32b (each thread gets its own row).  4 CTAs per SM
// share the 512 TMEM columns; the MMA of one CTA overlaps the exact phase of the others.
#include "device.cuh"

namespace ssb {

constexpr int TC_THREADS = 128;
constexpr int TC_ND = 128;    // densities per codebook-stream (MMA N)
constexpr int TC_L = 13;      // stream width
constexpr int TC_K = 32;      // padded K (floats) = one 128-byte swizzle row
constexpr int TC_RL = 28;     // exact record: det, mean[13], prec[13], pad
constexpr float TC_BIG = 7.9228163e28f;  // 2^96

// ---- PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ float to_tf32(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    // try_wait suspends the thread for a hardware time slice; the iteration cap turns a lost
    // completion (a malformed descriptor, say) into a trap instead of a hung GPU
    const uint32_t addr = smem_u32(bar);
    for (uint32_t it = 0; it < (1u << 22); ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok)
            return;
    }
    __trap();
}
__device__ __forceinline__ void fence_async_proxy()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row groups
// 1024 bytes apart (SBO), LBO unused (=1), descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address      [0,14)
    d |= (uint64_t)1 << 16;                           // leading byte off   [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte off    [32,46)
    d |= (uint64_t)1 << 46;                           // version            [46,48)
    d |= (uint64_t)2 << 61;                           // SWIZZLE_128B       [61,64)
    return d;
}
// instruction descriptor, kind::tf32: D=F32, A=B=TF32, both K-major, N=128, M=128
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((TC_ND >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(TC_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
          "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
          "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
          "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i)
        v[i] = __uint_as_float(r[i]);
}

// exact distance, the reference's operation order (ref: src/ptm_mgau.c:63-68, 106-127)
__device__ __forceinline__ float tc_exact_dist(const float *__restrict__ rec, const float (&x)[TC_L])
{
    float v[TC_RL];
    const float4 *r4 = reinterpret_cast<const float4 *>(rec);
#pragma unroll
    for (int i = 0; i < TC_RL / 4; ++i) {
        float4 q = r4[i];
        v[4 * i] = q.x;
        v[4 * i + 1] = q.y;
        v[4 * i + 2] = q.z;
        v[4 * i + 3] = q.w;
    }
    float d = v[0];
#pragma unroll
    for (int j = 0; j < TC_L; ++j) {
        float diff = __fsub_rn(x[j], v[1 + j]);
        float sq = __fmul_rn(diff, diff);
        float c = __fmul_rn(sq, v[1 + TC_L + j]);
        d = __fsub_rn(d, c);
    }
    return d;
}

template <int N>
struct TcTopN {
    int32_t s[N];
    int32_t c[N];
    __device__ __forceinline__ void reset()
    {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            s[k] = INT32_MIN;  // WORST_DIST (ref: src/ptm_mgau.c:694-720)
            c[k] = k;
        }
    }
    // eval_topn's insertion sort step: entry i moves up past strictly smaller scores (ref :70-84)
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
    // eval_cb's insertion: replaces the last entry, moves up past scores <= its own (ref :139-148)
    __device__ __forceinline__ void insert(int32_t sc, int32_t cw)
    {
        s[N - 1] = sc;
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

struct TcDebug {
    float *approx;     // [cs][frame][128] or null
    float *eps;        // [cs][frame]
    unsigned long long *counters;  // [0] exact evaluations of scan survivors [1] scanned (lane, frame) steps
};

// shared memory carve-up (dynamic, 1024-byte aligned for the swizzle atoms)
struct TcSmem {
    float B[TC_ND * TC_K];    // 16 KB, swizzled
    float A[TC_THREADS * TC_K];  // 16 KB, swizzled
    float rec[TC_ND * TC_RL];   // 14 KB exact records
    float aux[48];            // centre[13] | bm1[13] | bm2[13] | cmax
    uint64_t mbar;
    uint32_t tmem_base;
    int tmax;
};

template <int N, bool DEBUG>
__global__ void __launch_bounds__(TC_THREADS, 4)
gmm_topn_tc_kernel(DevModel m, DevPlan p, const float *__restrict__ feat, int64_t G,
                   const float *__restrict__ gB, const float *__restrict__ gAux,
                   int4 *__restrict__ out_s, uchar4 *__restrict__ out_c, TcDebug dbg)
{
    extern __shared__ uint8_t smem_raw[];
    TcSmem &S = *reinterpret_cast<TcSmem *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int cs = blockIdx.x;
    const int cb = cs / m.n_feat, f = cs - cb * m.n_feat;

    // ---- one-time setup: B (swizzled), exact records, aux, barrier, TMEM
    {
        const float4 *src = reinterpret_cast<const float4 *>(gB + (size_t)cs * TC_ND * TC_K);
        float4 *dstB = reinterpret_cast<float4 *>(S.B);
        for (int i = tid; i < TC_ND * TC_K / 4; i += TC_THREADS) {
            int n = i >> 3, j = i & 7;  // row, 16-byte chunk
            dstB[(n >> 3) * 64 + (n & 7) * 8 + (j ^ (n & 7))] = src[i];
        }
        const float4 *rsrc = reinterpret_cast<const float4 *>(m.gau + gau_offset(m, cb, f));
        float4 *rdst = reinterpret_cast<float4 *>(S.rec);
        for (int i = tid; i < TC_ND * TC_RL / 4; i += TC_THREADS)
            rdst[i] = rsrc[i];
        if (tid < 48)
            S.aux[tid] = gAux[(size_t)cs * 48 + tid];
        float4 *dstA = reinterpret_cast<float4 *>(S.A);
        for (int i = tid; i < TC_THREADS * TC_K / 4; i += TC_THREADS)
            dstA[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid == 0) {
            mbar_init(&S.mbar, 1);
            S.tmax = 0;
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)),
                     "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    const int u = blockIdx.y * TC_THREADS + tid;
    const bool has_utt = u < p.n_utts;
    const int64_t g0 = has_utt ? p.frame_off[u] : 0;
    const int T = has_utt ? (int)(p.frame_off[u + 1] - g0) : 0;
    atomicMax(&S.tmax, T);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;
    const int Tmax = S.tmax;
    const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);
    const uint64_t adesc = umma_desc_sw128(smem_u32(S.A));
    const uint64_t bdesc = umma_desc_sw128(smem_u32(S.B));

    float centre[TC_L];
#pragma unroll
    for (int j = 0; j < TC_L; ++j)
        centre[j] = S.aux[j];

    const float *xp = feat + g0 * m.blk + m.featoff[f];
    int4 *so = out_s + (int64_t)cs * G + g0;
    uchar4 *co = out_c + (int64_t)cs * G + g0;
    float4 *myA = reinterpret_cast<float4 *>(S.A) + (tid >> 3) * 64 + (tid & 7) * 8;
    const int swz = tid & 7;

    // active-codebook epochs of this utterance
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
    float x[TC_L], xn[TC_L];
#pragma unroll
    for (int j = 0; j < TC_L; ++j)
        xn[j] = (T > 0) ? __ldg(xp + j) : 0.f;
    unsigned long long n_exact = 0, n_steps = 0;

    for (int t = 0; t < Tmax; ++t) {
        const bool live = t < T;
        float eps = 0.f;
        if (live) {
            const float *xt = xp + (int64_t)t * m.blk;
#pragma unroll
            for (int j = 0; j < TC_L; ++j)
                x[j] = xn[j];
            if (t + 1 < T) {
#pragma unroll
                for (int j = 0; j < TC_L; ++j)
                    xn[j] = __ldg(xt + m.blk + j);
            }
            while (t >= t_next) {
                active = (p.ep_cbmask[(int64_t)e * 8 + (cb >> 5)] >> (cb & 31)) & 1u;
                ++e;
                t_next = e < e_end ? p.ep_start[e] : INT32_MAX;
            }
            // A row: [x' (13), x'^2 (13), 1, 1, 0 x4], TF32-rounded, 128-byte swizzled
            float a[TC_K];
            float acc = 0.f;  // the constant is split hi+lo (error 2^-22 |c|): covered by the margin
#pragma unroll
            for (int j = 0; j < TC_L; ++j) {
                float xc = __fsub_rn(x[j], centre[j]);
                float sq = __fmul_rn(xc, xc);
                a[j] = to_tf32(xc);
                a[TC_L + j] = to_tf32(sq);
                acc = fmaf(fabsf(xc), S.aux[13 + j], acc);
                acc = fmaf(sq, S.aux[26 + j], acc);
            }
            a[26] = 1.f;
            a[27] = 1.f;
#pragma unroll
            for (int j = 28; j < TC_K; ++j)
                a[j] = 0.f;
            // |approx - d| <= 2^-10 * sum(|x'||2 mu' v| + x'^2 v) (+ accumulation, centring) ; 1.5x margin
            eps = fmaf(acc, 1.5f / 1024.f, 16.f);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                myA[j ^ swz] = make_float4(a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
        }
        fence_async_proxy();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < TC_K / 8; ++k)
                umma_tf32(tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), k > 0 ? 1u : 0u);
            umma_commit(&S.mbar);
        }
        // eval_topn: re-score last frame's codewords exactly while the MMA runs
        if (live) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                float d = tc_exact_dist(S.rec + tn.c[i] * TC_RL, x);
                tn.s[i] = __float2int_rz(d);
                tn.settle(i);
            }
        }
        mbar_wait(&S.mbar, (uint32_t)(t & 1));
        tc_fence_after();
        const bool scan = live && active && (m.ds <= 1 || t % m.ds == 0);
        if (scan)
            ++n_steps;
#pragma unroll 1
        for (int ch = 0; ch < TC_ND / 32; ++ch) {
            float v[32];
            tmem_ld32(tmem_row + (uint32_t)(ch * 32), v);
            if (DEBUG && live && dbg.approx) {
                float *o = dbg.approx + (((int64_t)cs * G + g0 + t) * TC_ND + ch * 32);
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    o[i] = v[i];
                if (ch == 0)
                    dbg.eps[(int64_t)cs * G + g0 + t] = eps;
            }
            if (scan) {
                // survivors: approx > (float)worst - eps.  s = sat((approx - thr) * 2^96) is exactly 0 or 1.
                const float thr = __int2float_rn(tn.s[N - 1]) - eps;
                const float nthr = -thr * TC_BIG;
                // bit i of the mask <-> column i: accumulate from the top column down
                float m0 = 0.f, m1 = 0.f;
#pragma unroll
                for (int i = 15; i >= 0; --i) {
                    m0 = fmaf(m0, 2.f, __saturatef(fmaf(v[i], TC_BIG, nthr)));
                    m1 = fmaf(m1, 2.f, __saturatef(fmaf(v[16 + i], TC_BIG, nthr)));
                }
                uint32_t mask = (uint32_t)m0 | ((uint32_t)m1 << 16);
                while (mask) {
                    const int i = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int cw = ch * 32 + i;
                    const float d = tc_exact_dist(S.rec + cw * TC_RL, x);
                    ++n_exact;
                    if (d < __int2float_rn(tn.s[N - 1]))
                        continue;  // ref: src/ptm_mgau.c:174,209
                    if (tn.has(cw))
                        continue;  // ref :211-217
                    tn.insert(__float2int_rz(d), cw);
                }
            }
        }
        tc_fence_before();
        if (live && active) {
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
    }
    if (DEBUG && dbg.counters) {
        atomicAdd(&dbg.counters[0], n_exact);
        atomicAdd(&dbg.counters[1], n_steps);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

bool tc_supported(const DevModel &m)
{
    if (m.n_density != TC_ND || m.gB == nullptr)
        return false;
    for (int f = 0; f < m.n_feat; ++f)
        if (m.featlen[f] != TC_L)
            return false;
    return true;
}

int launch_gmm_topn_tc(const DevModel &m, const DevPlan &p, const float *feat, int64_t n_frames,
                       int4 *tn_score, uchar4 *tn_cw, float *dbg_approx, float *dbg_eps,
                       unsigned long long *dbg_counters, cudaStream_t st)
{
    if (p.n_utts == 0 || n_frames == 0)
        return 0;
    if (!tc_supported(m)) {
        set_error("tensor-core top-N needs 128 densities and 13-wide streams");
        return -1;
    }
    const size_t smem = sizeof(TcSmem) + 1024;
    dim3 grid(m.n_mgau * m.n_feat, (p.n_utts + TC_THREADS - 1) / TC_THREADS);
    TcDebug dbg{dbg_approx, dbg_eps, dbg_counters};
    const bool debug = dbg_approx != nullptr || dbg_counters != nullptr;
#define SSB_TC(NN, DBG)                                                                         \
    do {                                                                                        \
        SSB_CUDA(cudaFuncSetAttribute(gmm_topn_tc_kernel<NN, DBG>,                              \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        gmm_topn_tc_kernel<NN, DBG><<<grid, TC_THREADS, smem, st>>>(m, p, feat, n_frames, m.gB, \
                                                                    m.gAux, tn_score, tn_cw, dbg); \
    } while (0)
    switch (m.topn) {
    case 1:
        if (debug) SSB_TC(1, true); else SSB_TC(1, false);
        break;
    case 2:
        if (debug) SSB_TC(2, true); else SSB_TC(2, false);
        break;
    case 3:
        if (debug) SSB_TC(3, true); else SSB_TC(3, false);
        break;
    case 4:
        if (debug) SSB_TC(4, true); else SSB_TC(4, false);
        break;
    default:
        set_error("topn %d not supported (1..4)", m.topn);
        return -1;
    }
#undef SSB_TC
    SSB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ssb
