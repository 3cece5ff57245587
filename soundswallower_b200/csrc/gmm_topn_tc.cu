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
#include "tc_common.cuh"

namespace ssb {

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
            S.aux[tid] = gAux[(size_t)cs * SSB_TC_AUX + tid];
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
                if (ch == 0) {
                    dbg.eps[((int64_t)cs * G + g0 + t) * 2] = eps;
                    dbg.eps[((int64_t)cs * G + g0 + t) * 2 + 1] = eps;
                }
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
    if (m.n_density != TC_ND || m.gB == nullptr || m.kind != SSB_SCORER_PTM)
        return false;
    for (int f = 0; f < m.n_feat; ++f)
        if (m.featlen[f] != TC_L)
            return false;
    return true;
}

int launch_gmm_topn_tc(const DevModel &m, const DevPlan &p, const float *feat, int64_t n_frames,
                       int4 *tn_score, uchar4 *tn_cw, float *featp, float *dbg_approx,
                       float *dbg_eps, unsigned long long *dbg_counters, cudaStream_t st)
{
    if (p.n_utts == 0 || n_frames == 0)
        return 0;
    if (!tc_supported(m)) {
        set_error("tensor-core top-N needs 128 densities and 13-wide streams");
        return -1;
    }
    TcDebug dbg{dbg_approx, dbg_eps, dbg_counters};
    const bool debug = dbg_approx != nullptr || dbg_counters != nullptr;
    const char *force = getenv("SSB_K1");
    const bool v1 = m.ds > 1 || featp == nullptr || (force && strcmp(force, "tc1") == 0);
    if (!v1)
        return launch_gmm_topn_tc2(m, p, feat, n_frames, tn_score, tn_cw, featp, dbg, st);
    const size_t smem = sizeof(TcSmem) + 1024;
    dim3 grid(m.n_mgau * m.n_feat, (p.n_utts + TC_THREADS - 1) / TC_THREADS);
#define SSB_TC(NN, DBG)                                                                         \
    do {                                                                                        \
        SSB_DYN_SMEM((gmm_topn_tc_kernel<NN, DBG>), smem); \
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
    note_launch();
    return 0;
}

}  // namespace ssb
