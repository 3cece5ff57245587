// gmm_scan_ft.cu -- K1 "frame-tiled": Gaussian top-N of tiles of 128 consecutive frames against
// every active codebook (persistent CTAs, tile pairs handed out by an atomic counter).  Same
// results as the reference's eval_topn + eval_cb (ref: src/ptm_mgau.c:63-253) wherever integer
// scores do not tie; tie steps are flagged (DevPlan.tie_bits) for the literal replay
// (topn_fixup.cu / the grammar search's own replay).
//
// Against gmm_topn_tc2.cu (CTA = codebook-stream x 256 utterances, thread = utterance):
//   * the MMA's M rows are FRAMES, so the A tile [x', x'^2, 1] (3xTF32 hi / lo split, centred on
//     one global centre per stream) is built once per (tile, stream) and shared by all
//     codebooks; it lives in TENSOR MEMORY (tcgen05.st, A-from-TMEM MMA), not in shared memory;
//   * a CTA works on TWO tiles (row groups) at a time: B_hi / B_lo of a codebook-stream (32 KB,
//     pre-swizzled in HBM, L2 resident) travel once per 256 frames through a two-slot
//     shared-memory ring filled by cp.async.bulk + mbarrier (TMA warp) and are consumed by a
//     single MMA-issuing thread, one accumulator per row group in TMEM;
//   * the epilogue is warp-specialised and free of block barriers: 8 warps, thread = frame row,
//     all 128 densities in registers.  One TMEM read-out; group maxima -> N-th largest ->
//     survivor mask (one FADD + one funnel shift per density); the row's screening scores are
//     parked in a private shared-memory stash and the survivors picked up by index, packed with
//     their index into one sortable 32-bit key, branch-free insertion;
//   * NO exact evaluation on the common path: if the N+1 best screening scores are separated
//     by more than the rigorous error bound and none lies within the bound of a 1024-raw-unit
//     boundary, the list (codewords in order + scores >> 10, which is all that the mixing stage
//     consumes: ref src/ptm_mgau.c:276-285) is decided.  Otherwise (~13 % of the steps) the row
//     files the step in its own small queue; a warp works its lanes' queues off together
//     (exact evaluations in the reference's operation order) when enough lanes have one.
#include "tc_common.cuh"

namespace ssb {

constexpr int FT_ROWS = 128;
constexpr int FT_RG = 2;                // row groups (tiles) per CTA
constexpr int FT_EPI_THREADS = FT_RG * FT_ROWS;
constexpr int FT_THREADS = FT_EPI_THREADS + 64;  // + TMA warp + MMA warp
constexpr int FT_NSLOT = 2;
constexpr int FT_BTILE = TC_ND * TC_K;  // floats of one operand tile (16 KB)
constexpr int FT_STASH_LD = 132;        // floats per stash row (bank-conflict-free float4 stores)
constexpr int FT_PQ = 3;                // undecided steps a row can hold
constexpr int FT_MAXCB = 64;            // codebooks a model can have here
constexpr uint32_t FT_A_COL = 256;      // TMEM columns: accumulators [0,256), A tiles from 256

struct FtSmem {
    float B[FT_NSLOT][2 * FT_BTILE];          // B_hi | B_lo per slot (SWIZZLE_128B, as stored in HBM)
    float stash[FT_EPI_THREADS * FT_STASH_LD];
    uint32_t pq[FT_PQ][6][FT_EPI_THREADS];    // per-row queue of undecided steps: cs | need << 16, 5 keys
    uint64_t b_full[FT_NSLOT], b_empty[FT_NSLOT], acc_full[FT_RG], acc_empty[FT_RG], a_ready[FT_RG];
    uint32_t tmem_base;
    int next_pair;
    int n_cb;
    uint32_t um[FT_RG][2];                    // codebooks each row group needs (n_mgau <= 64)
    uint8_t cb_list[FT_MAXCB];
};

// ---- PTX -----------------------------------------------------------------------------------
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
// 32 accumulator columns of this thread's row, no wait: the values may only be used after
// FT_TMEM_WAIT32 on the same array (which names the registers, so that nothing is scheduled
// across it); a second load can be in flight while the first one's values are processed
#define FT_TMEM_LD32_NOWAIT(taddr, v)                                                                        \
    asm volatile(                                                                                            \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                            \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                            \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"            \
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]),    \
          "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]),           \
          "=f"(v[15]), "=f"(v[16]), "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]),         \
          "=f"(v[22]), "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]),         \
          "=f"(v[29]), "=f"(v[30]), "=f"(v[31])                                                              \
        : "r"(taddr)                                                                                         \
        : "memory")
#define FT_TMEM_WAIT32(v)                                                                                    \
    asm volatile("tcgen05.wait::ld.sync.aligned;"                                                            \
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]),       \
                   "+f"(v[7]), "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]),   \
                   "+f"(v[14]), "+f"(v[15]), "+f"(v[16]), "+f"(v[17]), "+f"(v[18]), "+f"(v[19]),             \
                   "+f"(v[20]), "+f"(v[21]), "+f"(v[22]), "+f"(v[23]), "+f"(v[24]), "+f"(v[25]),             \
                   "+f"(v[26]), "+f"(v[27]), "+f"(v[28]), "+f"(v[29]), "+f"(v[30]), "+f"(v[31])              \
                 :                                                                                           \
                 : "memory")

__device__ __forceinline__ float ft_max3(float a, float b, float c)
{
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float ft_max16(const float (&v)[32], int o)
{
    const float m0 = ft_max3(v[o], v[o + 1], v[o + 2]), m1 = ft_max3(v[o + 3], v[o + 4], v[o + 5]);
    const float m2 = ft_max3(v[o + 6], v[o + 7], v[o + 8]), m3 = ft_max3(v[o + 9], v[o + 10], v[o + 11]);
    const float m4 = ft_max3(v[o + 12], v[o + 13], v[o + 14]);
    return fmaxf(ft_max3(m0, m1, m2), ft_max3(m3, m4, v[o + 15]));
}
__device__ __forceinline__ float ft_max16_ex(const float (&v)[32], int o, uint32_t exclude)
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
// k-th largest (k = 1..4) of 8 values: 19-comparator sorting network, descending
__device__ __forceinline__ float ft_kth_largest8(float (&g)[8], int k)
{
    FT_CE(g[0], g[1]) FT_CE(g[2], g[3]) FT_CE(g[4], g[5]) FT_CE(g[6], g[7])
    FT_CE(g[0], g[2]) FT_CE(g[1], g[3]) FT_CE(g[4], g[6]) FT_CE(g[5], g[7])
    FT_CE(g[1], g[2]) FT_CE(g[5], g[6]) FT_CE(g[0], g[4]) FT_CE(g[3], g[7])
    FT_CE(g[1], g[5]) FT_CE(g[2], g[6])
    FT_CE(g[1], g[4]) FT_CE(g[3], g[6])
    FT_CE(g[2], g[4]) FT_CE(g[3], g[5])
    FT_CE(g[3], g[4])
    return k == 1 ? g[0] : (k == 2 ? g[1] : (k == 3 ? g[2] : g[3]));
}
// bit i of the result: v[i] >= thr  (sign bit of v - thr, collected by funnel shifts; two
// independent chains)
__device__ __forceinline__ uint32_t ft_mask32(const float (&v)[32], float thr)
{
    uint32_t lo = 0u, hi = 0u;
#pragma unroll
    for (int i = 15; i >= 0; --i) {
        lo = __funnelshift_l(__float_as_uint(__fsub_rn(v[i], thr)), lo, 1);
        hi = __funnelshift_l(__float_as_uint(__fsub_rn(v[16 + i], thr)), hi, 1);
    }
    return ~(lo | (hi << 16));
}

struct FtArgs {
    const float *feat;
    int64_t G;
    int4 *out_s;
    uchar4 *out_c;
    const int32_t *tile_utt, *tile_t0;
    int n_tiles;
    int *tile_counter;  // zeroed before the launch; pair = gridDim.x + ticket
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

// One deferred step: exact evaluation (the reference's fp32 operation order, no contraction) of
// the survivors named by the mask; the N+1 best decide.  Ties -> flagged, list as found.
// Called with the warp converged; lanes without work pass have = false.
template <int N, bool DEBUG>
__device__ __forceinline__ void ft_exact_step(const DevModel &m, const DevPlan &p, const FtArgs &a, bool have,
                                              int cs, int64_t g, const float (&x)[TC_L], uint32_t k0,
                                              uint32_t k1, uint32_t k2, uint32_t k3)
{
    const int cb = cs / m.n_feat, f = cs - cb * m.n_feat;
    const float *rec = m.gau + gau_offset(m, cb, f);
    TcTopN<N + 1> best;
#pragma unroll
    for (int k = 0; k <= N; ++k) {
        best.s[k] = INT32_MIN;
        best.c[k] = -1;
    }
    int cnt = 0;
    unsigned long long lo = have ? ((unsigned long long)k0 | ((unsigned long long)k1 << 32)) : 0ull;
    unsigned long long hi = have ? ((unsigned long long)k2 | ((unsigned long long)k3 << 32)) : 0ull;
    while (lo | hi) {  // ascending density index
        int cw;
        if (lo) {
            cw = __ffsll((long long)lo) - 1;
            lo &= lo - 1;
        } else {
            cw = 63 + __ffsll((long long)hi);
            hi &= hi - 1;
        }
        const int32_t sc = __float2int_rz(tc_exact_dist(rec + (size_t)cw * TC_RL, x));
        ++cnt;
        if (sc >= best.s[N])
            best.insert(sc, cw);
    }
    if (!have)
        return;
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

// A step whose N+1 best screening scores (keys, descending) do not decide it: the candidates in
// `need` get their exact score (the reference's fp32 operation order), all N+1 take their
// places, ties are flagged, the list is stored.  Whole warp; lanes without work pass have = false.
template <int N, bool DEBUG>
__device__ __forceinline__ void ft_resolve(const DevModel &m, const DevPlan &p, const FtArgs &a, bool have,
                                           int cs, int64_t g, const float (&x)[TC_L], uint32_t need,
                                           const float (&tk)[N + 1])
{
    const int cb = cs / m.n_feat, f = cs - cb * m.n_feat;
    const float *rec = m.gau + gau_offset(m, cb, f);
    int32_t fs[N + 1], fc[N + 1];
#pragma unroll
    for (int k = 0; k <= N; ++k) {
        fs[k] = tk[k] > -1.0e30f ? __float2int_rz(tk[k]) : INT32_MIN;
        fc[k] = (int)(__float_as_uint(tk[k]) & 127u);
    }
    uint32_t nd = have ? need : 0u;
    while (nd) {
        const int k = __ffs((int)nd) - 1;
        nd &= nd - 1u;
        int cw = fc[0];
#pragma unroll
        for (int q = 1; q <= N; ++q)
            cw = q == k ? fc[q] : cw;
        const int32_t sc = __float2int_rz(tc_exact_dist(rec + (size_t)cw * TC_RL, x));
#pragma unroll
        for (int q = 0; q <= N; ++q)
            fs[q] = q == k ? sc : fs[q];
        if (DEBUG && a.dbg.counters)
            atomicAdd(&a.dbg.counters[0], 1ull);
    }
    if (!have)
        return;
    // candidates that were in doubt take their exact places (the others are separated from
    // everybody beyond doubt): sort the N+1, descending
#define FT_CES(i, j)                                                   \
    if (fs[j] > fs[i]) {                                               \
        const int32_t ts_ = fs[i], tc_ = fc[i];                        \
        fs[i] = fs[j];                                                 \
        fc[i] = fc[j];                                                 \
        fs[j] = ts_;                                                   \
        fc[j] = tc_;                                                   \
    }
    if (N == 4) {
        FT_CES(0, 1) FT_CES(3, 4) FT_CES(2, 4) FT_CES(2, 3) FT_CES(1, 4)
        FT_CES(0, 3) FT_CES(0, 2) FT_CES(1, 3) FT_CES(1, 2)
    } else if (N == 3) {
        FT_CES(0, 1) FT_CES(2, 3) FT_CES(0, 2) FT_CES(1, 3) FT_CES(1, 2)
    } else if (N == 2) {
        FT_CES(0, 1) FT_CES(1, 2) FT_CES(0, 1)
    } else {
        FT_CES(0, 1)
    }
#undef FT_CES
    // which entries hold exact scores after the sort?  (exact ones keep their raw value, the
    // others are stored as (score >> 10) << 10; telling them apart is not needed: both shift right)
    bool distinct = true;
#pragma unroll
    for (int k = 0; k < N; ++k)
        distinct = distinct && (fs[k] > fs[k + 1]);
    if (!distinct) {  // exact scores tie: the list depends on the one carried in
        if (p.tie_bits)
            atomicOr(&p.tie_bits[(int64_t)cs * p.tie_w + (g >> 5)], 1u << (g & 31));
        if (DEBUG && a.dbg.counters)
            atomicAdd(&a.dbg.counters[2], 1ull);
    }
    ft_store<N>(a, cs, g, fs, fc);
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
        for (int r = 0; r < FT_RG; ++r) {
            mbar_init(&S.acc_full[r], 1);
            mbar_init(&S.acc_empty[r], FT_ROWS / 32);
            mbar_init(&S.a_ready[r], FT_ROWS / 32);
        }
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
    // phase counters, advanced identically by every role
    uint32_t step = 0;                 // B loads since the kernel began
    uint32_t acc_it[FT_RG] = {0, 0};   // accumulator hand-overs per row group
    uint32_t a_it[FT_RG] = {0, 0};     // A tiles built per row group
    const int n_pairs = (a.n_tiles + FT_RG - 1) / FT_RG;
    int pair = blockIdx.x;
    while (pair < n_pairs) {
        // ---- the pair's tiles; codebooks = union over the epochs that overlap them (the evaluated
        // list is NOT monotone: the >255-gap bridging senones of acmod_flags2list come and go,
        // ref: src/acmod.c:968-973); every row then checks its own epoch's mask
        if (warp < FT_RG) {
            const int tile = pair * FT_RG + warp;
            uint32_t um0 = 0u, um1 = 0u;
            if (tile < a.n_tiles) {
                const int u = a.tile_utt[tile], t0 = a.tile_t0[tile];
                const int T = (int)(p.frame_off[u + 1] - p.frame_off[u]);
                const int nr = min(FT_ROWS, T - t0);
                if (p.all_active) {
                    um0 = um1 = 0xffffffffu;
                } else {
                    const int e_lo = p.ep_off[u], e_hi = p.ep_off[u + 1];
                    for (int e = e_lo + lane; e < e_hi; e += 32)
                        if (p.ep_start[e] <= t0 + nr - 1 && (e + 1 >= e_hi || p.ep_start[e + 1] > t0)) {
                            um0 |= p.ep_cbmask[(int64_t)e * 8];
                            um1 |= p.ep_cbmask[(int64_t)e * 8 + 1];
                        }
                    um0 = __reduce_or_sync(0xffffffffu, um0);
                    um1 = __reduce_or_sync(0xffffffffu, um1);
                }
                if (m.n_mgau < 32)
                    um0 &= (1u << m.n_mgau) - 1u;
                um1 = m.n_mgau > 32 ? (m.n_mgau < 64 ? um1 & ((1u << (m.n_mgau - 32)) - 1u) : um1) : 0u;
            }
            if (lane == 0) {
                S.um[warp][0] = um0;
                S.um[warp][1] = um1;
            }
        }
        __syncthreads();
        if (warp == 0) {
            uint32_t w0 = 0u, w1 = 0u;
            for (int r = 0; r < FT_RG; ++r) {
                w0 |= S.um[r][0];
                w1 |= S.um[r][1];
            }
            int n_before = 0;
            for (int c0 = 0; c0 < 64; c0 += 32) {
                const bool on = ((c0 ? w1 : w0) >> lane) & 1u;
                const uint32_t bal = __ballot_sync(0xffffffffu, on);
                if (on)
                    S.cb_list[n_before + __popc(bal & ((1u << lane) - 1u))] = (uint8_t)(c0 + lane);
                n_before += __popc(bal);
            }
            if (lane == 0)
                S.n_cb = n_before;
        }
        __syncthreads();
        const int n_cb = S.n_cb;
        uint32_t rgm[FT_RG][2];
        for (int r = 0; r < FT_RG; ++r) {
            rgm[r][0] = S.um[r][0];
            rgm[r][1] = S.um[r][1];
        }

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
            // ================= MMA warp: 12 x tcgen05.mma (M128 N128 K8, tf32) per row group and step =================
            if (lane == 0) {
                uint32_t s = step, ai[FT_RG] = {acc_it[0], acc_it[1]};
                for (int f = 0; f < NF; ++f) {
                    for (int r = 0; r < FT_RG; ++r)
                        if (rgm[r][0] | rgm[r][1])
                            mbar_wait(&S.a_ready[r], (a_it[r] + (uint32_t)f) & 1u);
                    tc_fence_after();
                    for (int i = 0; i < n_cb; ++i, ++s) {
                        const int slot = (int)(s % FT_NSLOT);
                        const int cb = S.cb_list[i];
                        mbar_wait(&S.b_full[slot], (s / FT_NSLOT) & 1u);
                        const uint64_t bhi = umma_desc_sw128(smem_u32(&S.B[slot][0]));
                        const uint64_t blo = umma_desc_sw128(smem_u32(&S.B[slot][FT_BTILE]));
                        for (int r = 0; r < FT_RG; ++r) {
                            if (!((rgm[r][cb >> 5] >> (cb & 31)) & 1u))
                                continue;
                            mbar_wait(&S.acc_empty[r], (ai[r] & 1u) ^ 1u);
                            tc_fence_after();
                            const uint32_t a_hi = tmem + FT_A_COL + (uint32_t)(64 * r), a_lo = a_hi + 32u;
                            const uint32_t d = tmem + (uint32_t)(128 * r);
#pragma unroll
                            for (int k = 0; k < TC_K / 8; ++k)
                                umma_tf32_ts(d, a_hi + (uint32_t)(8 * k), bhi + (uint64_t)(2 * k), k > 0 ? 1u : 0u);
#pragma unroll
                            for (int k = 0; k < TC_K / 8; ++k)
                                umma_tf32_ts(d, a_lo + (uint32_t)(8 * k), bhi + (uint64_t)(2 * k), 1u);
#pragma unroll
                            for (int k = 0; k < TC_K / 8; ++k)
                                umma_tf32_ts(d, a_hi + (uint32_t)(8 * k), blo + (uint64_t)(2 * k), 1u);
                            umma_commit(&S.acc_full[r]);
                            ++ai[r];
                        }
                        umma_commit(&S.b_empty[slot]);
                    }
                }
            }
            __syncwarp();
        } else {
            // ================= epilogue warps: thread = frame row =================
            const int rg = warp >> 2, wq = warp & 3, row = wq * 32 + lane;
            const int tile = pair * FT_RG + rg;
            if ((rgm[rg][0] | rgm[rg][1]) != 0u) {  // (implies tile < n_tiles)
                const int u = a.tile_utt[tile], t0 = a.tile_t0[tile];
                const int64_t gu = p.frame_off[u];
                const int T = (int)(p.frame_off[u + 1] - gu);
                const int nr = min(FT_ROWS, T - t0);
                const int64_t g = gu + t0 + row;
                const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
                const bool row_valid = row < nr;
                // active codebooks of this row's epoch
                uint32_t rm0 = 0u, rm1 = 0u;
                if (row_valid) {
                    if (p.all_active) {
                        rm0 = rm1 = 0xffffffffu;
                    } else {
                        int er = -1;
                        for (int e = p.ep_off[u]; e < p.ep_off[u + 1] && p.ep_start[e] <= t0 + row; ++e)
                            er = e;
                        if (er >= 0) {
                            rm0 = p.ep_cbmask[(int64_t)er * 8];
                            rm1 = p.ep_cbmask[(int64_t)er * 8 + 1];
                        }
                    }
                }
                const float *xrow = a.feat + (row_valid ? g : gu + t0) * m.blk;
                float *srow = &S.stash[(rg * FT_ROWS + row) * FT_STASH_LD];
                uint32_t ai = acc_it[rg];
                const int qrow = rg * FT_ROWS + row;
                int pq_n = 0;
                for (int f = 0; f < NF; ++f) {
                    // ---- this stream's features; A tile (hi | lo) -> TMEM
                    float xs[TC_L];
                    {
                        float ah[32], al[32], xc[TC_L];
#pragma unroll
                        for (int j = 0; j < TC_L; ++j) {
                            xs[j] = row_valid ? __ldg(xrow + m.featoff[f] + j) : 0.f;
                            xc[j] = __fsub_rn(xs[j], m.ft_centre[f * 16 + j]);
                            const float sq = __fmul_rn(xc[j], xc[j]);
                            ah[j] = to_tf32(xc[j]);
                            ah[TC_L + j] = to_tf32(sq);
                            al[j] = to_tf32(__fsub_rn(xc[j], ah[j]));
                            al[TC_L + j] = to_tf32(__fsub_rn(sq, ah[TC_L + j]));
                        }
                        ah[26] = ah[27] = 1.f;
                        al[26] = al[27] = 0.f;
#pragma unroll
                        for (int j = 28; j < 32; ++j)
                            ah[j] = al[j] = 0.f;
                        // (the row group's MMAs of the previous stream are complete: their last
                        // accumulator has been read)
                        tmem_st32(tmem + lane_base + FT_A_COL + (uint32_t)(64 * rg), ah);
                        tmem_st32(tmem + lane_base + FT_A_COL + (uint32_t)(64 * rg + 32), al);
                        tmem_st_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0)
                            mbar_arrive(&S.a_ready[rg]);
                    }
                    for (int i = 0; i < n_cb; ++i) {
                        const int cb = S.cb_list[i];
                        if (!((rgm[rg][cb >> 5] >> (cb & 31)) & 1u))
                            continue;  // (uniform for the row group; the MMA warp skips it too)
                        const int cs = cb * NF + f;
                        const bool active = ((cb < 32 ? rm0 >> cb : rm1 >> (cb - 32)) & 1u) != 0u;
                        const float4 *ax = reinterpret_cast<const float4 *>(m.gAuxFt + (size_t)cs * 64);
                        // error bound of the screening scores (2^-18 of the term magnitudes + 4)
                        float xa[TC_L], xq[TC_L];
#pragma unroll
                        for (int j = 0; j < TC_L; ++j) {
                            const float xcj = __fsub_rn(xs[j], m.ft_centre[f * 16 + j]);
                            xa[j] = fabsf(xcj);
                            xq[j] = __fmul_rn(xcj, xcj);
                        }
                        float eps, eps_hot = 0.f;
                        bool has_hot;
                        {
                            const float4 q0 = __ldg(ax), q1 = __ldg(ax + 1), q2 = __ldg(ax + 2), q3 = __ldg(ax + 3);
                            const float4 r0 = __ldg(ax + 4), r1 = __ldg(ax + 5), r2 = __ldg(ax + 6), r3 = __ldg(ax + 7);
                            const float m1[13] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x};
                            const float m2[13] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w, r3.x};
                            float e0 = q3.y, e1 = 0.f;  // q3.y = max |c| over the regular densities
#pragma unroll
                            for (int j = 0; j < TC_L; ++j) {
                                e0 = fmaf(xa[j], m1[j], e0);
                                e1 = fmaf(xq[j], m2[j], e1);
                            }
                            eps = fmaf(e0 + e1, 1.f / 262144.f, 4.f);
                            has_hot = q3.w != 0.f;
                            eps_hot = q3.z;  // max |c| over the hot ones (completed below)
                        }
                        uint32_t hot[4] = {0u, 0u, 0u, 0u};
                        if (has_hot) {  // uniform
                            const float4 q0 = __ldg(ax + 8), q1 = __ldg(ax + 9), q2 = __ldg(ax + 10), q3 = __ldg(ax + 11);
                            const float4 r0 = __ldg(ax + 12), r1 = __ldg(ax + 13), r2 = __ldg(ax + 14), r3 = __ldg(ax + 15);
                            const float m1[13] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x};
                            const float m2[13] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w, r3.x};
                            float h0 = eps_hot, h1 = 0.f;
#pragma unroll
                            for (int j = 0; j < TC_L; ++j) {
                                h0 = fmaf(xa[j], m1[j], h0);
                                h1 = fmaf(xq[j], m2[j], h1);
                            }
                            eps_hot = fmaf(h0 + h1, 1.f / 262144.f, 4.f);
                            const uint4 hw = __ldg(reinterpret_cast<const uint4 *>(m.gHot + (size_t)cs * 4));
                            hot[0] = hw.x;
                            hot[1] = hw.y;
                            hot[2] = hw.z;
                            hot[3] = hw.w;
                        }
                        mbar_wait(&S.acc_full[rg], ai & 1u);
                        ++ai;
                        tc_fence_after();
                        const bool warp_active = __any_sync(0xffffffffu, active);
                        if (!warp_active) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0)
                                mbar_arrive(&S.acc_empty[rg]);
                            continue;
                        }
                        const uint32_t ta = tmem + lane_base + (uint32_t)(128 * rg);
                        float va[32], vb[32];
                        // ---- pass 1 over the accumulator row: group maxima over the regular densities
                        float gm[8];
#define FT_GM(v, c)                                                                     \
    do {                                                                                \
        if (has_hot) {                                                                  \
            gm[2 * (c)] = ft_max16_ex(v, 0, hot[c] & 0xffffu);                          \
            gm[2 * (c) + 1] = ft_max16_ex(v, 16, hot[c] >> 16);                         \
        } else {                                                                        \
            gm[2 * (c)] = ft_max16(v, 0);                                               \
            gm[2 * (c) + 1] = ft_max16(v, 16);                                          \
        }                                                                               \
    } while (0)
                        FT_TMEM_LD32_NOWAIT(ta, va);
                        FT_TMEM_WAIT32(va);
                        FT_TMEM_LD32_NOWAIT(ta + 32u, vb);
                        FT_GM(va, 0);
                        FT_TMEM_WAIT32(vb);
                        FT_TMEM_LD32_NOWAIT(ta + 64u, va);
                        FT_GM(vb, 1);
                        FT_TMEM_WAIT32(va);
                        FT_TMEM_LD32_NOWAIT(ta + 96u, vb);
                        FT_GM(va, 2);
                        FT_TMEM_WAIT32(vb);
                        FT_TMEM_LD32_NOWAIT(ta, va);  // (pass 2 starts travelling)
                        FT_GM(vb, 3);
#undef FT_GM
                        const float L = ft_kth_largest8(gm, N);
                        const float thr = L - 2.f * eps - 2.f;
                        const float thr_hot = L - eps - eps_hot - 2.f;
                        // ---- pass 2: survivor masks; chunks with survivors are parked in the stash
                        uint32_t mk[4];
#define FT_MK(v, c)                                                                                  \
    do {                                                                                             \
        uint32_t mm = ft_mask32(v, thr);                                                             \
        if (has_hot && hot[c])                                                                       \
            mm = (mm & ~hot[c]) | (ft_mask32(v, thr_hot) & hot[c]);                                  \
        mk[c] = active ? mm : 0u;                                                                    \
        if (__any_sync(0xffffffffu, mk[c] != 0u)) {                                                  \
            float4 *sp = reinterpret_cast<float4 *>(srow + 32 * (c));                                \
            _Pragma("unroll") for (int k = 0; k < 8; ++k)                                            \
                sp[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);             \
        }                                                                                            \
        if (DEBUG && active && a.dbg.approx) {                                                       \
            float *o = a.dbg.approx + ((int64_t)cs * a.G + g) * TC_ND + 32 * (c);                    \
            _Pragma("unroll") for (int k = 0; k < 32; ++k) o[k] = v[k];                              \
        }                                                                                            \
    } while (0)
                        FT_TMEM_WAIT32(va);
                        FT_TMEM_LD32_NOWAIT(ta + 32u, vb);
                        FT_MK(va, 0);
                        FT_TMEM_WAIT32(vb);
                        FT_TMEM_LD32_NOWAIT(ta + 64u, va);
                        FT_MK(vb, 1);
                        FT_TMEM_WAIT32(va);
                        FT_TMEM_LD32_NOWAIT(ta + 96u, vb);
                        FT_MK(va, 2);
                        FT_TMEM_WAIT32(vb);
                        // the accumulator has been read: the next codebook's MMAs may overwrite it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0)
                            mbar_arrive(&S.acc_empty[rg]);
                        FT_MK(vb, 3);
#undef FT_MK
                        if (DEBUG && active && a.dbg.approx) {
                            a.dbg.eps[((int64_t)cs * a.G + g) * 2] = eps;
                            a.dbg.eps[((int64_t)cs * a.G + g) * 2 + 1] = eps_hot;
                        }
                        // ---- the N+2 best survivors: value and index packed in one sortable key
                        // (the 7 low mantissa bits carry the density index: 2^-16 of the value lost).
                        // Fixed trip count: the first 8 survivors are picked up with all their
                        // shared-memory loads in flight and sorted by a network; a row with more
                        // inserts the rest one by one.
                        float tk[N + 2];
                        {   // (all lanes: the votes below need the whole warp; inactive rows have no survivors)
                            if (DEBUG && active && a.dbg.counters)
                                atomicAdd(&a.dbg.counters[1], 1ull);
                            uint32_t m0 = mk[0], m1 = mk[1], m2 = mk[2], m3 = mk[3];
                            float k8[8];
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                const bool z0 = m0 == 0u, z1 = m1 == 0u, z2 = m2 == 0u;
                                const uint32_t w = z0 ? (z1 ? (z2 ? m3 : m2) : m1) : m0;
                                const int base = z0 ? (z1 ? (z2 ? 96 : 64) : 32) : 0;
                                const uint32_t cl = w & (w - 1u);
                                const int cw = base + (w ? __ffs((int)w) - 1 : 0);
                                m0 = z0 ? m0 : cl;
                                m1 = (z0 && !z1) ? cl : m1;
                                m2 = (z0 && z1 && !z2) ? cl : m2;
                                m3 = (z0 && z1 && z2) ? cl : m3;
                                const float key = __uint_as_float((__float_as_uint(srow[cw]) & 0xffffff80u) | (uint32_t)cw);
                                k8[it] = w ? key : -3.4028235e38f;
                            }
                            FT_CE(k8[0], k8[1]) FT_CE(k8[2], k8[3]) FT_CE(k8[4], k8[5]) FT_CE(k8[6], k8[7])
                            FT_CE(k8[0], k8[2]) FT_CE(k8[1], k8[3]) FT_CE(k8[4], k8[6]) FT_CE(k8[5], k8[7])
                            FT_CE(k8[1], k8[2]) FT_CE(k8[5], k8[6]) FT_CE(k8[0], k8[4]) FT_CE(k8[3], k8[7])
                            FT_CE(k8[1], k8[5]) FT_CE(k8[2], k8[6])
                            FT_CE(k8[1], k8[4]) FT_CE(k8[3], k8[6])
                            FT_CE(k8[2], k8[4]) FT_CE(k8[3], k8[5])
                            FT_CE(k8[3], k8[4])
#pragma unroll
                            for (int k = 0; k < N + 2; ++k)
                                tk[k] = k8[k];
                            uint32_t rest[4] = {m0, m1, m2, m3};
                            if (__any_sync(0xffffffffu, (m0 | m1 | m2 | m3) != 0u)) {
#pragma unroll
                                for (int c = 0; c < 4; ++c) {
                                    uint32_t w = rest[c];
                                    while (w) {
                                        const int cw = 32 * c + __ffs((int)w) - 1;
                                        w &= w - 1u;
                                        const float key = __uint_as_float((__float_as_uint(srow[cw]) & 0xffffff80u) | (uint32_t)cw);
#pragma unroll
                                        for (int k = N + 1; k >= 1; --k)
                                            tk[k] = fmaxf(tk[k], fminf(tk[k - 1], key));
                                        tk[0] = fmaxf(tk[0], key);
                                    }
                                }
                            }
                        }
                        // Which of the N+1 best need their exact score?  Those within the error
                        // bound of a 1024-unit boundary (their score >> 10 is in doubt) and those
                        // not separated from a neighbour beyond doubt (their order is).
                        // |key - score| <= 2^-16 |score|; hot densities carry their own bound.
                        uint32_t need = 0u;
                        bool full;
                        {
                            const float lastv = tk[N + 1] > -1.0e30f ? tk[N + 1] : (tk[N] > -1.0e30f ? tk[N] : tk[N - 1]);
                            const float qe = fmaxf(fabsf(tk[0]), fabsf(lastv)) * (1.f / 32768.f);
                            float ek[N + 2];
#pragma unroll
                            for (int k = 0; k < N + 2; ++k) {
                                ek[k] = eps;
                                if (has_hot) {
                                    const int ci = (int)(__float_as_uint(tk[k]) & 127u);
                                    const int hq = ci >> 5;  // (select chain: no dynamic register index)
                                    const uint32_t hw = hq == 0 ? hot[0] : (hq == 1 ? hot[1] : (hq == 2 ? hot[2] : hot[3]));
                                    if ((hw >> (ci & 31)) & 1u)
                                        ek[k] = eps_hot;
                                }
                            }
                            full = a.exact || !(fabsf(tk[0]) < 1.0e9f) || !(fabsf(lastv) < 1.0e9f);
#pragma unroll
                            for (int k = 0; k < N; ++k) {
                                // distance of the score to the nearest multiple of 1024
                                const int lowb = __float2int_rz(tk[k]) & 1023;
                                if (!(__int2float_rn(min(lowb, 1024 - lowb)) > ek[k] + 3.f + qe))
                                    need |= 1u << k;
                                full = full || !(fabsf(tk[k]) >= 1.f);  // (keys near 0 could be flushed)
                                if (!(tk[k] - tk[k + 1] > ek[k] + ek[k + 1] + 2.f + qe))
                                    need |= 3u << k;
                            }
                            // the (N+2)-th best within reach of the N-th: more than N+1 candidates
                            // for the list -> every survivor is evaluated (rare)
                            full = full || !(tk[N - 1] - tk[N + 1] > ek[N - 1] + ek[N + 1] + 2.f + qe);
                            full = full && active;
                            if (full || !active)
                                need = 0u;
                        }
                        if (active && !full && need == 0u) {
                            // decided by the screening scores alone: (int)d >> 10, stored << 10
                            int32_t fs[N + 1], fc[N + 1];
#pragma unroll
                            for (int k = 0; k <= N; ++k) {
                                fs[k] = tk[k] > -1.0e30f ? (__float2int_rz(tk[k]) >> SENSCR_SHIFT) * 1024 : INT32_MIN;
                                fc[k] = (int)(__float_as_uint(tk[k]) & 127u);
                            }
                            ft_store<N>(a, cs, g, fs, fc);
                        }
                        if (__any_sync(0xffffffffu, full))
                            ft_exact_step<N, DEBUG>(m, p, a, full, cs, g, xs, mk[0], mk[1], mk[2], mk[3]);
                        // ---- undecided steps wait in the row's queue until the warp has enough of them
                        {
                            const bool want_push = need != 0u;
                            if (__any_sync(0xffffffffu, want_push && pq_n >= FT_PQ)) {
                                // a full queue: one round first
                                const bool have = pq_n > 0;
                                float qk[N + 1];
                                uint32_t hd = 0u;
                                if (have) {
                                    --pq_n;
                                    hd = S.pq[pq_n][0][qrow];
#pragma unroll
                                    for (int k = 0; k <= N; ++k)
                                        qk[k] = __uint_as_float(S.pq[pq_n][1 + k][qrow]);
                                } else {
#pragma unroll
                                    for (int k = 0; k <= N; ++k)
                                        qk[k] = 0.f;
                                }
                                ft_resolve<N, DEBUG>(m, p, a, have, (int)(hd & 0xffffu), g, xs, hd >> 16, qk);
                            }
                            if (want_push) {
                                S.pq[pq_n][0][qrow] = (uint32_t)cs | (need << 16);
#pragma unroll
                                for (int k = 0; k <= N; ++k)
                                    S.pq[pq_n][1 + k][qrow] = __float_as_uint(tk[k]);
                                ++pq_n;
                            }
                            const uint32_t hv = __ballot_sync(0xffffffffu, pq_n > 0);
                            if (__popc(hv) >= 26) {
                                const bool have = pq_n > 0;
                                float qk[N + 1];
                                uint32_t hd = 0u;
                                if (have) {
                                    --pq_n;
                                    hd = S.pq[pq_n][0][qrow];
#pragma unroll
                                    for (int k = 0; k <= N; ++k)
                                        qk[k] = __uint_as_float(S.pq[pq_n][1 + k][qrow]);
                                } else {
#pragma unroll
                                    for (int k = 0; k <= N; ++k)
                                        qk[k] = 0.f;
                                }
                                ft_resolve<N, DEBUG>(m, p, a, have, (int)(hd & 0xffffu), g, xs, hd >> 16, qk);
                            }
                        }
                    }
                    // ---- end of the stream: the queues are emptied (xs changes)
                    while (__any_sync(0xffffffffu, pq_n > 0)) {
                        const bool have = pq_n > 0;
                        float qk[N + 1];
                        uint32_t hd = 0u;
                        if (have) {
                            --pq_n;
                            hd = S.pq[pq_n][0][qrow];
#pragma unroll
                            for (int k = 0; k <= N; ++k)
                                qk[k] = __uint_as_float(S.pq[pq_n][1 + k][qrow]);
                        } else {
#pragma unroll
                            for (int k = 0; k <= N; ++k)
                                qk[k] = 0.f;
                        }
                        ft_resolve<N, DEBUG>(m, p, a, have, (int)(hd & 0xffffu), g, xs, hd >> 16, qk);
                    }
                }
            }
        }
        // ---- phase counters, then the next pair
        for (int r = 0; r < FT_RG; ++r) {
            const uint32_t w0 = rgm[r][0], w1 = rgm[r][1];
            if (w0 | w1) {
                acc_it[r] += (uint32_t)((__popc(w0) + __popc(w1)) * NF);
                a_it[r] += (uint32_t)NF;
            }
        }
        step += (uint32_t)(n_cb * NF);
        __syncthreads();
        if (tid == 0)
            S.next_pair = (int)gridDim.x + atomicAdd(a.tile_counter, 1);
        __syncthreads();
        pair = S.next_pair;
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
    const int n_pairs = (n_tiles + FT_RG - 1) / FT_RG;
    const int grid = n_pairs < sms ? n_pairs : sms;
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
