// tc_common.cuh -- tcgen05 / TMEM / mbarrier PTX wrappers and the exact-arithmetic pieces
// shared by the tensor-core top-N kernels (gmm_topn_tc.cu, gmm_topn_tc2.cu).
#pragma once
#include <cstdlib>
#include <cstring>

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
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 22); ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\t"
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

// v2 kernel (gmm_topn_tc2.cu); featp = scratch for the re-packed features
int launch_gmm_topn_tc2(const DevModel &m, const DevPlan &p, const float *feat, int64_t n_frames,
                        int4 *tn_score, uchar4 *tn_cw, float *featp, TcDebug dbg, cudaStream_t st);

}  // namespace ssb
