// gmm_scan_ft.cu -- K1 "frame-tiled": Gaussian top-N of a tile of 128 consecutive frames
// against every active codebook, one CTA per tile (persistent, tiles handed out by an atomic
// counter).  Same results as the reference's eval_topn + eval_cb (ref: src/ptm_mgau.c:63-253)
// wherever integer scores do not tie; tie steps are flagged (DevPlan.tie_bits) for the
// literal replay (topn_fixup.cu / the grammar search's own replay).
//
// Against gmm_topn_tc2.cu (CTA = codebook-stream x 256 utterances, thread = utterance):
//   * the MMA's M rows are FRAMES, so the A tile [x', x'^2, 1] (3xTF32 hi / lo split, centred on
//     one global centre per stream) is built ONCE per tile and shared by all codebooks; it
//     lives in TENSOR MEMORY (tcgen05.st, A-from-TMEM MMA), not in shared memory;
//   * B_hi / B_lo of a codebook-stream (32 KB, pre-swizzled in HBM, L2 resident) stream
//     through a two-slot shared-memory ring filled by cp.async.bulk + mbarrier (TMA warp),
//     consumed by a single MMA-issuing thread; two accumulators in TMEM let the tensor core
//     run one codebook ahead of the epilogue;
//   * the epilogue is warp-specialised: 8 warps, two threads per frame row (64 of the 128
//     densities each).  One TMEM read-out; group maxima -> N-th largest -> survivor mask
//     (one FADD + one funnel shift per density); the survivors' screening scores are parked
//     in a shared-memory stash and picked up by index;
//   * NO exact evaluation on the common path: if the N+1 best screening scores are separated
//     by more than the rigorous error bound and none lies within the bound of a 1024-raw-unit
//     boundary, the list (codewords in order + scores >> 10, which is all that the mixing stage
//     consumes: ref src/ptm_mgau.c:276-285) is decided.  Otherwise (~13 % of the steps) the row
//     files the step in a shared-memory queue; the queue is drained by all threads together
//     (no divergence) with exact evaluations in the reference's operation order.
#include "tc_common.cuh"

namespace ssb {

constexpr int FT_ROWS = 128;
constexpr int FT_EPI_THREADS = 256;
constexpr int FT_THREADS = 320;  // 8 epilogue warps + TMA warp + MMA warp
constexpr int FT_NSLOT = 2;
constexpr int FT_BTILE = TC_ND * TC_K;  // floats of one operand tile (16 KB)
constexpr int FT_STASH_LD = 132;        // floats per stash row (bank-conflict-free float4 stores)
constexpr int FT_QCAP = 512;
constexpr int FT_QWORDS = 5;            // header + 128-bit survivor mask
constexpr int FT_QDRAIN = 160;          // drain the queue when it holds this many items
constexpr int FT_MAXCB = 64;            // codebooks a tile can list
constexpr uint32_t FT_A_COL = 256;      // TMEM columns: accumulators [0,256), A tiles from 256

struct FtSmem {
    float B[FT_NSLOT][2 * FT_BTILE];          // B_hi | B_lo per slot (SWIZZLE_128B, as stored in HBM)
    float stash[FT_ROWS * FT_STASH_LD];
    float4 xg[FT_ROWS][2];                    // pair exchange: sorted group maxima
    float2 xe[FT_ROWS][2];                    //                error-bound partial sums
    unsigned long long xm[FT_ROWS][2];        //                survivor masks
    uint32_t q[FT_QCAP][FT_QWORDS];
    uint64_t b_full[FT_NSLOT], b_empty[FT_NSLOT], acc_full[2], acc_empty[2], a_ready;
    uint32_t tmem_base;
    int q_count;
    int next_tile;
    int n_cb;
    uint8_t cb_list[FT_MAXCB];
    int16_t cb_r0[FT_MAXCB];
};

// ---- PTX -----------------------------------------------------------------------------------
__device__ __forceinline__ void ft_bar_pair(int wq)
{
    asm volatile("bar.sync %0, 64;" ::"r"(wq + 1) : "memory");
}
__device__ __forceinline__ void ft_bar_epi()
{
    asm volatile("bar.sync 5, 256;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(TC_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
        "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
        "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])),
        "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])),
        "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])),
        "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])),
        "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float ft_max3(float a, float b, float c)
{
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float ft_max16(const float (&v)[64], int o)
{
    const float m0 = ft_max3(v[o], v[o + 1], v[o + 2]), m1 = ft_max3(v[o + 3], v[o + 4], v[o + 5]);
    const float m2 = ft_max3(v[o + 6], v[o + 7], v[o + 8]), m3 = ft_max3(v[o + 9], v[o + 10], v[o + 11]);
    const float m4 = ft_max3(v[o + 12], v[o + 13], v[o + 14]);
    return fmaxf(ft_max3(m0, m1, m2), ft_max3(m3, m4, v[o + 15]));
}
__device__ __forceinline__ float ft_max16_ex(const float (&v)[64], int o, uint32_t exclude)
{
    float w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i)
        w[i] = ((exclude >> i) & 1u) ? -3.4028235e38f : v[o + i];
    const float m0 = ft_max3(w[0], w[1], w[2]), m1 = ft_max3(w[3], w[4], w[5]), m2 = ft_max3(w[6], w[7], w[8]);
    const float m3 = ft_max3(w[9], w[10], w[11]), m4 = ft_max3(w[12], w[13], w[14]);
    return fmaxf(ft_max3(m0, m1, m2), ft_max3(m3, m4, w[15]));
}
#define FT_CE(a, b)              \
    {                            \
        float hi_ = fmaxf(a, b); \
        b = fminf(a, b);         \
        a = hi_;                 \
    }
// bit i of the result: v[o + i] >= thr  (sign bit of v - thr, collected by funnel shifts)
__device__ __forceinline__ uint32_t ft_mask32(const float (&v)[64], int o, float thr)
{
    uint32_t below = 0u;
#pragma unroll
    for (int i = 31; i >= 0; --i)
        below = __funnelshift_l(__float_as_uint(__fsub_rn(v[o + i], thr)), below, 1);
    return ~below;
}

struct FtArgs {
    const float *feat;
    int64_t G;
    int4 *out_s;
    uchar4 *out_c;
    const int32_t *tile_utt, *tile_t0;
    int n_tiles;
    int *tile_counter;  // zeroed before the launch; tile = gridDim.x + ticket
    int exact;          // every step through the exact path: raw scores out (ssb_topn_batch)
    TcDebug dbg;
};

template <int N>
__device__ __forceinline__ void ft_store(const FtArgs &a, int cs, int64_t g, const int32_t (&s)[N + 1],
                                         const int32_t (&c)[N + 1])
{
    int4 sv = make_int4(INT32_MIN, INT32_MIN, INT32_MIN, INT32_MIN);
    uchar4 cv = make_uchar4(0, 0, 0, 0);
    sv.x = s[0];
    cv.x = (unsigned char)c[0];
    if (N > 1) {
        sv.y = s[N > 1 ? 1 : 0];
        cv.y = (unsigned char)c[N > 1 ? 1 : 0];
    }
    if (N > 2) {
        sv.z = s[N > 2 ? 2 : 0];
        cv.z = (unsigned char)c[N > 2 ? 2 : 0];
    }
    if (N > 3) {
        sv.w = s[N > 3 ? 3 : 0];
        cv.w = (unsigned char)c[N > 3 ? 3 : 0];
    }
    a.out_s[(int64_t)cs * a.G + g] = sv;
    a.out_c[(int64_t)cs * a.G + g] = cv;
}

// One queued step: exact evaluation (the reference's fp32 operation order, no contraction) of
// the survivors named by the mask; the N+1 best decide.  Ties -> flagged, list as found.
template <int N, bool DEBUG>
__device__ __noinline__ void ft_exact_item(const DevModel &m, const DevPlan &p, const FtArgs &a, int cs,
                                           int64_t g, uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3)
{
    const int cb = cs / m.n_feat, f = cs - cb * m.n_feat;
    const float *rec = m.gau + gau_offset(m, cb, f);
    const float *xp = a.feat + g * m.blk + m.featoff[f];
    float x[TC_L];
#pragma unroll
    for (int j = 0; j < TC_L; ++j)
        x[j] = __ldg(xp + j);
    TcTopN<N + 1> best;
#pragma unroll
    for (int k = 0; k <= N; ++k) {
        best.s[k] = INT32_MIN;
        best.c[k] = -1;
    }
    int cnt = 0;
    uint32_t w[4] = {k0, k1, k2, k3};
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
        uint32_t bits = w[q];
        while (bits) {
            const int cw = q * 32 + __ffs((int)bits) - 1;
            bits &= bits - 1;
            const int32_t sc = __float2int_rz(tc_exact_dist(rec + (size_t)cw * TC_RL, x));
            ++cnt;
            if (sc >= best.s[N])
                best.insert(sc, cw);
        }
    }
    bool distinct = cnt >= N;
#pragma unroll
    for (int k = 0; k < N; ++k)
        distinct = distinct && (best.s[k] > best.s[k + 1]);
    if (!distinct && p.tie_bits)
        atomicOr(&p.tie_bits[(int64_t)cs * p.tie_w + (g >> 5)], 1u << (g & 31));
    if (DEBUG && a.dbg.counters) {
        atomicAdd(&a.dbg.counters[0], (unsigned long long)cnt);
        if (!distinct)
            atomicAdd(&a.dbg.counters[2], 1ull);
    }
    ft_store<N>(a, cs, g, best.s, best.c);
}

template <int N, bool DEBUG>
__device__ __forceinline__ void ft_drain(FtSmem &S, const DevModel &m, const DevPlan &p, const FtArgs &a,
                                         int64_t gtile, int tid, int n)
{
    for (int i = tid; i < n; i += FT_EPI_THREADS) {
        const uint32_t hd = S.q[i][0];
        ft_exact_item<N, DEBUG>(m, p, a, (int)(hd >> 8), gtile + (int)(hd & 0xffu), S.q[i][1], S.q[i][2],
                                S.q[i][3], S.q[i][4]);
    }
}

template <int N, bool DEBUG>
__global__ void __launch_bounds__(FT_THREADS, 1)
gmm_scan_ft_kernel(DevModel m, DevPlan p, FtArgs a)
{
    extern __shared__ uint8_t smem_raw[];
    FtSmem &S = *reinterpret_cast<FtSmem *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NF = m.n_feat;

    if (tid == 0) {
        for (int s = 0; s < FT_NSLOT; ++s) {
            mbar_init(&S.b_full[s], 1);
            mbar_init(&S.b_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&S.acc_full[s], 1);
            mbar_init(&S.acc_empty[s], FT_EPI_THREADS / 32);
        }
        mbar_init(&S.a_ready, FT_EPI_THREADS / 32);
        S.q_count = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)),
                     "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;

    uint32_t step = 0;      // codebook-stream steps since the kernel began (same in every role)
    uint32_t tile_it = 0;   // tiles with work this CTA has started (phase of a_ready)
    int tile = blockIdx.x;
    while (tile < a.n_tiles) {
        const int u = a.tile_utt[tile], t0 = a.tile_t0[tile];
        const int64_t gu = p.frame_off[u];
        const int T = (int)(p.frame_off[u + 1] - gu);
        const int nr = min(FT_ROWS, T - t0);
        const int64_t gtile = gu + t0;
        // ---- codebooks of the tile: those active on its last frame (the aligner's active set
        // only grows, ref: src/state_align_search.c:186-188), each from the row it enters on
        if (warp == 0) {
            int e_lo = 0, e_hi = 0, e_last = -1;
            if (!p.all_active) {
                e_lo = p.ep_off[u];
                e_hi = p.ep_off[u + 1];
                for (int e = e_lo; e < e_hi && p.ep_start[e] <= t0 + nr - 1; ++e)
                    e_last = e;
            }
            int n_before = 0;
            for (int c0 = 0; c0 < m.n_mgau; c0 += 32) {
                const int cb = c0 + lane;
                bool on = false;
                int r0 = 0;
                if (cb < m.n_mgau) {
                    if (p.all_active)
                        on = true;
                    else if (e_last >= 0
                             && ((p.ep_cbmask[(int64_t)e_last * 8 + (cb >> 5)] >> (cb & 31)) & 1u)) {
                        on = true;
                        for (int e = e_lo; e <= e_last; ++e)
                            if ((p.ep_cbmask[(int64_t)e * 8 + (cb >> 5)] >> (cb & 31)) & 1u) {
                                r0 = max(p.ep_start[e] - t0, 0);
                                break;
                            }
                    }
                }
                const uint32_t bal = __ballot_sync(0xffffffffu, on);
                if (on) {
                    const int pos = n_before + __popc(bal & ((1u << lane) - 1u));
                    if (pos < FT_MAXCB) {
                        S.cb_list[pos] = (uint8_t)cb;
                        S.cb_r0[pos] = (int16_t)r0;
                    }
                }
                n_before += __popc(bal);
            }
            if (lane == 0)
                S.n_cb = min(n_before, FT_MAXCB);
        }
        __syncthreads();
        const int n_cb = S.n_cb;
        const uint32_t n_steps = (uint32_t)(n_cb * NF);

        if (warp == FT_EPI_THREADS / 32) {
            // ================= TMA warp: B_hi | B_lo of each listed codebook-stream =================
            if (lane == 0) {
                uint32_t s = step;
                for (int f = 0; f < NF; ++f)
                    for (int i = 0; i < n_cb; ++i, ++s) {
                        const int slot = (int)(s % FT_NSLOT);
                        const int cs = (int)S.cb_list[i] * NF + f;
                        mbar_wait(&S.b_empty[slot], ((s / FT_NSLOT) & 1u) ^ 1u);
                        mbar_expect_tx(&S.b_full[slot], 2 * FT_BTILE * 4);
                        bulk_g2s(&S.B[slot][0], m.gBft + (size_t)cs * 2 * FT_BTILE, 2 * FT_BTILE * 4,
                                 &S.b_full[slot]);
                    }
            }
            __syncwarp();
        } else if (warp == FT_EPI_THREADS / 32 + 1) {
            // ================= MMA warp: 12 x tcgen05.mma (M128 N128 K8, tf32) per step =================
            if (lane == 0 && n_steps > 0) {
                mbar_wait(&S.a_ready, tile_it & 1u);
                tc_fence_after();
                uint32_t s = step;
                for (int f = 0; f < NF; ++f) {
                    const uint32_t a_hi = tmem + FT_A_COL + (uint32_t)(64 * f), a_lo = a_hi + 32u;
                    for (int i = 0; i < n_cb; ++i, ++s) {
                        const int slot = (int)(s % FT_NSLOT);
                        const uint32_t accb = s & 1u;
                        mbar_wait(&S.b_full[slot], (s / FT_NSLOT) & 1u);
                        mbar_wait(&S.acc_empty[accb], ((s >> 1) & 1u) ^ 1u);
                        tc_fence_after();
                        const uint64_t bhi = umma_desc_sw128(smem_u32(&S.B[slot][0]));
                        const uint64_t blo = umma_desc_sw128(smem_u32(&S.B[slot][FT_BTILE]));
                        const uint32_t d = tmem + accb * 128u;
#pragma unroll
                        for (int k = 0; k < TC_K / 8; ++k)
                            umma_tf32_ts(d, a_hi + (uint32_t)(8 * k), bhi + (uint64_t)(2 * k), k > 0 ? 1u : 0u);
#pragma unroll
                        for (int k = 0; k < TC_K / 8; ++k)
                            umma_tf32_ts(d, a_lo + (uint32_t)(8 * k), bhi + (uint64_t)(2 * k), 1u);
#pragma unroll
                        for (int k = 0; k < TC_K / 8; ++k)
                            umma_tf32_ts(d, a_hi + (uint32_t)(8 * k), blo + (uint64_t)(2 * k), 1u);
                        umma_commit(&S.b_empty[slot]);
                        umma_commit(&S.acc_full[accb]);
                    }
                }
            }
            __syncwarp();
        } else if (n_steps > 0) {
            // ================= epilogue warps =================
            const int wq = warp & 3, h = warp >> 2, row = wq * 32 + lane;
            const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
            const bool row_valid = row < nr;
            const float *xrow = a.feat + (gtile + (row_valid ? row : 0)) * m.blk;
            // ---- A tiles of the three streams -> TMEM (thread h = 0: hi parts, h = 1: lo parts)
            for (int f = 0; f < NF; ++f) {
                float av[32];
#pragma unroll
                for (int j = 0; j < TC_L; ++j) {
                    const float xv = row_valid ? __ldg(xrow + m.featoff[f] + j) : 0.f;
                    const float xc = __fsub_rn(xv, m.ft_centre[f * 16 + j]);
                    const float sq = __fmul_rn(xc, xc);
                    const float xh = to_tf32(xc), sh = to_tf32(sq);
                    av[j] = h == 0 ? xh : to_tf32(__fsub_rn(xc, xh));
                    av[TC_L + j] = h == 0 ? sh : to_tf32(__fsub_rn(sq, sh));
                }
                av[26] = h == 0 ? 1.f : 0.f;
                av[27] = h == 0 ? 1.f : 0.f;
#pragma unroll
                for (int j = 28; j < 32; ++j)
                    av[j] = 0.f;
                tmem_st32(tmem + lane_base + FT_A_COL + (uint32_t)(64 * f + 32 * h), av);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&S.a_ready);

            uint32_t s = step;
            for (int f = 0; f < NF; ++f) {
                // this thread's half of the error-bound sum: |x'| (h = 0) or x'^2 (h = 1)
                float ev[TC_L];
#pragma unroll
                for (int j = 0; j < TC_L; ++j) {
                    const float xv = row_valid ? __ldg(xrow + m.featoff[f] + j) : 0.f;
                    const float xc = __fsub_rn(xv, m.ft_centre[f * 16 + j]);
                    ev[j] = h == 0 ? fabsf(xc) : __fmul_rn(xc, xc);
                }
                for (int i = 0; i < n_cb; ++i, ++s) {
                    const int cb = S.cb_list[i];
                    const int cs = cb * NF + f;
                    const bool active = row_valid && row >= (int)S.cb_r0[i];
                    const uint32_t accb = s & 1u;
                    // bound operands of this codebook-stream (L1-resident, 256 B per cs)
                    const float4 *ax = reinterpret_cast<const float4 *>(m.gAuxFt + (size_t)cs * 64 + 16 * h);
                    const float4 a0 = __ldg(ax), a1 = __ldg(ax + 1), a2 = __ldg(ax + 2), a3 = __ldg(ax + 3);
                    const bool has_hot = a3.w != 0.f;
                    mbar_wait(&S.acc_full[accb], (s >> 1) & 1u);
                    tc_fence_after();
                    const bool warp_active = __any_sync(0xffffffffu, active);
                    float v[64];
                    if (warp_active) {
                        float t[32];
                        tmem_ld32(tmem + lane_base + accb * 128u + (uint32_t)(64 * h), t);
#pragma unroll
                        for (int k = 0; k < 32; ++k)
                            v[k] = t[k];
                        tmem_ld32(tmem + lane_base + accb * 128u + (uint32_t)(64 * h + 32), t);
#pragma unroll
                        for (int k = 0; k < 32; ++k)
                            v[32 + k] = t[k];
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0)
                        mbar_arrive(&S.acc_empty[accb]);
                    if (warp_active) {
                        // ---- front: group maxima (regular densities), bound, exchange with the pair
                        uint32_t hot_lo = 0u, hot_hi = 0u;
                        float part_hot = 0.f;
                        float g0, g1, g2, g3;
                        if (has_hot) {  // uniform per step
                            hot_lo = __ldg(m.gHot + (size_t)cs * 4 + 2 * h);
                            hot_hi = __ldg(m.gHot + (size_t)cs * 4 + 2 * h + 1);
                            g0 = ft_max16_ex(v, 0, hot_lo & 0xffffu);
                            g1 = ft_max16_ex(v, 16, hot_lo >> 16);
                            g2 = ft_max16_ex(v, 32, hot_hi & 0xffffu);
                            g3 = ft_max16_ex(v, 48, hot_hi >> 16);
                            const float4 *hx = ax + 8;  // + 32 floats: the hot densities' maxima
                            const float4 b0 = __ldg(hx), b1 = __ldg(hx + 1), b2 = __ldg(hx + 2), b3 = __ldg(hx + 3);
                            const float hm[13] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w,
                                                  b2.x, b2.y, b2.z, b2.w, b3.x};
#pragma unroll
                            for (int j = 0; j < TC_L; ++j)
                                part_hot = fmaf(ev[j], hm[j], part_hot);
                            if (h == 1)
                                part_hot += a3.z;  // max |c| over the hot densities
                        } else {
                            g0 = ft_max16(v, 0);
                            g1 = ft_max16(v, 16);
                            g2 = ft_max16(v, 32);
                            g3 = ft_max16(v, 48);
                        }
                        FT_CE(g0, g1) FT_CE(g2, g3) FT_CE(g0, g2) FT_CE(g1, g3) FT_CE(g1, g2)
                        float part = 0.f;
                        {
                            const float rm[13] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w,
                                                  a2.x, a2.y, a2.z, a2.w, a3.x};
#pragma unroll
                            for (int j = 0; j < TC_L; ++j)
                                part = fmaf(ev[j], rm[j], part);
                            if (h == 1)
                                part += a3.y;  // max |c| over the regular densities
                        }
                        S.xg[row][h] = make_float4(g0, g1, g2, g3);
                        S.xe[row][h] = make_float2(part, part_hot);
                        ft_bar_pair(wq);
                        const float4 og = S.xg[row][1 - h];
                        const float2 oe = S.xe[row][1 - h];
                        // N-th largest of the eight group maxima (two sorted quadruples)
                        float L;
                        if (N == 4)
                            L = fmaxf(ft_max3(og.w, fminf(g0, og.z), fminf(g1, og.y)), fmaxf(fminf(g2, og.x), g3));
                        else if (N == 3)
                            L = fmaxf(ft_max3(og.z, fminf(g0, og.y), fminf(g1, og.x)), g2);
                        else if (N == 2)
                            L = ft_max3(og.y, fminf(g0, og.x), g1);
                        else
                            L = fmaxf(g0, og.x);
                        // split-TF32 products are good to ~2^-20 of the term magnitudes, measured
                        // worst case 2^-20.4; the bound used is 2^-18 (+4 raw units), as in v2
                        const float eps = fmaf(part + oe.x, 1.f / 262144.f, 4.f);
                        const float eps_hot = fmaf(part_hot + oe.y, 1.f / 262144.f, 4.f);
                        const float thr = L - 2.f * eps - 2.f;
                        uint32_t mk_lo = ft_mask32(v, 0, thr), mk_hi = ft_mask32(v, 32, thr);
                        if (has_hot) {
                            const float thr_hot = L - eps - eps_hot - 2.f;
                            if (hot_lo)
                                mk_lo = (mk_lo & ~hot_lo) | (ft_mask32(v, 0, thr_hot) & hot_lo);
                            if (hot_hi)
                                mk_hi = (mk_hi & ~hot_hi) | (ft_mask32(v, 32, thr_hot) & hot_hi);
                        }
                        if (!active)
                            mk_lo = mk_hi = 0u;
                        {
                            float4 *sp = reinterpret_cast<float4 *>(&S.stash[row * FT_STASH_LD + 64 * h]);
#pragma unroll
                            for (int k = 0; k < 16; ++k)
                                sp[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                        }
                        S.xm[row][h] = (unsigned long long)mk_lo | ((unsigned long long)mk_hi << 32);
                        if (DEBUG && active && a.dbg.approx) {
                            float *o = a.dbg.approx + (((int64_t)cs * a.G + gtile + row) * TC_ND + 64 * h);
#pragma unroll
                            for (int k = 0; k < 64; ++k)
                                o[k] = v[k];
                            if (h == 0) {
                                a.dbg.eps[((int64_t)cs * a.G + gtile + row) * 2] = eps;
                                a.dbg.eps[((int64_t)cs * a.G + gtile + row) * 2 + 1] = eps_hot;
                            }
                        }
                        ft_bar_pair(wq);
                        // ---- back: one thread of the pair (alternating) decides the step
                        if (active && h == (int)(s & 1u)) {
                            const unsigned long long om = S.xm[row][1 - h];
                            const unsigned long long mine = (unsigned long long)mk_lo | ((unsigned long long)mk_hi << 32);
                            unsigned long long mlo = h == 0 ? mine : om, mhi = h == 0 ? om : mine;
                            const uint32_t k0 = (uint32_t)mlo, k1 = (uint32_t)(mlo >> 32);
                            const uint32_t k2 = (uint32_t)mhi, k3 = (uint32_t)(mhi >> 32);
                            bool decided = false;
                            const int nsurv = __popcll(mlo) + __popcll(mhi);
                            if (DEBUG && a.dbg.counters)
                                atomicAdd(&a.dbg.counters[1], 1ull);
                            if (!a.exact && nsurv <= 12) {
                                float tv[N + 1];
                                int ti[N + 1];
#pragma unroll
                                for (int k = 0; k <= N; ++k) {
                                    tv[k] = -3.4028235e38f;
                                    ti[k] = 0;
                                }
                                const float *srow = &S.stash[row * FT_STASH_LD];
                                while (mlo | mhi) {
                                    int cw;
                                    if (mlo) {
                                        cw = __ffsll((long long)mlo) - 1;
                                        mlo &= mlo - 1;
                                    } else {
                                        cw = 63 + __ffsll((long long)mhi);
                                        mhi &= mhi - 1;
                                    }
                                    const float val = srow[cw];
                                    if (val > tv[N]) {
                                        tv[N] = val;
                                        ti[N] = cw;
#pragma unroll
                                        for (int j = N - 1; j >= 0; --j)
                                            if (tv[j + 1] > tv[j]) {
                                                const float fv = tv[j];
                                                tv[j] = tv[j + 1];
                                                tv[j + 1] = fv;
                                                const int iv = ti[j];
                                                ti[j] = ti[j + 1];
                                                ti[j + 1] = iv;
                                            }
                                    }
                                }
                                // decided iff the N+1 best are ordered beyond doubt (2 eps + 2 apart)
                                // and the N best sit clear of every 1024-unit boundary (eps + 2)
                                bool ok = true;
                                const float sep = 2.f * eps + 2.f, clr = eps + 2.f;
#pragma unroll
                                for (int k = 0; k < N; ++k) {
                                    ok = ok && (tv[k] - tv[k + 1] > sep);
                                    const float r = tv[k] * (1.f / 1024.f);
                                    ok = ok && (fabsf(r - rintf(r)) * 1024.f > clr);
                                }
                                if (has_hot) {
                                    // a hot density among the candidates carries the larger bound
#pragma unroll
                                    for (int k = 0; k <= N; ++k) {
                                        const uint32_t hw = __ldg(m.gHot + (size_t)cs * 4 + (ti[k] >> 5));
                                        ok = ok && !((hw >> (ti[k] & 31)) & 1u);
                                    }
                                }
                                if (ok) {
                                    int32_t qs[N + 1], qc[N + 1];
#pragma unroll
                                    for (int k = 0; k <= N; ++k) {
                                        // (int)d >> 10 of the exact score; stored << 10 (consumers shift)
                                        qs[k] = (__float2int_rz(tv[k]) >> SENSCR_SHIFT) * 1024;
                                        qc[k] = ti[k];
                                    }
                                    ft_store<N>(a, cs, gtile + row, qs, qc);
                                    decided = true;
                                }
                            }
                            if (!decided) {
                                const int qi = atomicAdd(&S.q_count, 1);
                                if (qi < FT_QCAP) {
                                    S.q[qi][0] = (uint32_t)row | ((uint32_t)cs << 8);
                                    S.q[qi][1] = k0;
                                    S.q[qi][2] = k1;
                                    S.q[qi][3] = k2;
                                    S.q[qi][4] = k3;
                                } else {
                                    ft_exact_item<N, DEBUG>(m, p, a, cs, gtile + row, k0, k1, k2, k3);
                                }
                            }
                        }
                    }
                    // ---- drain the queue every 4 steps (all epilogue threads, evenly)
                    if ((i & 3) == 3 || i == n_cb - 1) {
                        ft_bar_epi();
                        const int nq = min(S.q_count, FT_QCAP);
                        const bool last = (i == n_cb - 1) && (f == NF - 1);
                        if (nq >= FT_QDRAIN || (last && nq > 0)) {
                            ft_drain<N, DEBUG>(S, m, p, a, gtile, tid, nq);
                            ft_bar_epi();
                            if (tid == 0)
                                S.q_count = 0;
                            ft_bar_epi();
                        }
                    }
                }
            }
        }
        step += n_steps;
        if (n_steps > 0)
            ++tile_it;
        // ---- next tile
        __syncthreads();
        if (tid == 0)
            S.next_tile = (int)gridDim.x + atomicAdd(a.tile_counter, 1);
        __syncthreads();
        tile = S.next_tile;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

bool ft_supported(const DevModel &m)
{
    return tc_supported(m) && m.gBft != nullptr && m.n_mgau <= FT_MAXCB && m.ds <= 1 && m.topn >= 1 && m.topn <= 4;
}

int launch_gmm_scan_ft(const DevModel &m, const DevPlan &p, const float *feat, int64_t n_frames,
                       int4 *tn_score, uchar4 *tn_cw, const int32_t *tile_utt, const int32_t *tile_t0,
                       int n_tiles, int *tile_counter, int exact, TcDebug dbg, cudaStream_t st)
{
    if (p.n_utts == 0 || n_frames == 0 || n_tiles == 0)
        return 0;
    if (!ft_supported(m)) {
        set_error("frame-tiled top-N: model shape not supported");
        return -1;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = n_tiles < sms ? n_tiles : sms;
    SSB_CUDA(cudaMemsetAsync(tile_counter, 0, sizeof(int), st));
    FtArgs a;
    a.feat = feat;
    a.G = n_frames;
    a.out_s = tn_score;
    a.out_c = tn_cw;
    a.tile_utt = tile_utt;
    a.tile_t0 = tile_t0;
    a.n_tiles = n_tiles;
    a.tile_counter = tile_counter;
    a.exact = exact;
    a.dbg = dbg;
    const bool debug = dbg.approx != nullptr || dbg.counters != nullptr;
    const size_t smem = sizeof(FtSmem) + 1024;
#define SSB_FT(NN, DBG)                                                        \
    do {                                                                       \
        SSB_DYN_SMEM((gmm_scan_ft_kernel<NN, DBG>), smem);                     \
        gmm_scan_ft_kernel<NN, DBG><<<grid, FT_THREADS, smem, st>>>(m, p, a);  \
    } while (0)
    switch (m.topn) {
    case 1:
        if (debug) SSB_FT(1, true); else SSB_FT(1, false);
        break;
    case 2:
        if (debug) SSB_FT(2, true); else SSB_FT(2, false);
        break;
    case 3:
        if (debug) SSB_FT(3, true); else SSB_FT(3, false);
        break;
    default:
        if (debug) SSB_FT(4, true); else SSB_FT(4, false);
        break;
    }
#undef SSB_FT
    SSB_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

}  // namespace ssb
