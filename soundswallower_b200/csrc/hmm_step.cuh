// hmm_step.cuh -- one Viterbi step of a left-to-right HMM (non-multiplex), shared by the chain
// aligner (chain_viterbi.cu) and the FSG token-passing search (fsg_search.cu).
// Replaces hmm_vit_eval_3st_lr / hmm_vit_eval_5st_lr (ref: src/hmm.c:166-304, 482-567); every
// tie-break, clamp and stale value of the reference is kept.
#pragma once
#include "device.cuh"

namespace ssb {

constexpr int32_t TMAT_WORST = -255;  // ref: include/soundswallower/hmm.h:86

__device__ __forceinline__ int32_t clampw(int32_t x) { return x < WORST_SCORE ? WORST_SCORE : x; }

// ref: src/hmm.c:482-567.  sc/hi: emitting-state scores/histories; ss: senone scores
// (non-negative costs); tp: [3][4] uint8 costs.  Returns the HMM's best score.
__device__ __forceinline__ int32_t hmm_step3(const uint8_t *__restrict__ tp, const int (&ss)[3],
                                             int32_t (&sc)[3], int32_t (&hi)[3], int32_t &osc,
                                             int32_t &ohi)
{
#define TP(i, j) (-(int32_t)tp[(i) * 4 + (j)])
    int32_t s2 = sc[2] - ss[2], s1 = sc[1] - ss[1], s0 = sc[0] - ss[0];
    int32_t t0, t1, t2 = INT32_MIN, best = WORST_SCORE;
    if (s1 > WORST_SCORE) {
        int32_t s3;
        t1 = s2 + TP(2, 3);
        if (TP(1, 3) > TMAT_WORST)
            t2 = s1 + TP(1, 3);
        if (t1 > t2) {
            s3 = t1;
            ohi = hi[2];
        } else {
            s3 = t2;
            ohi = hi[1];
        }
        s3 = clampw(s3);
        osc = s3;
        best = s3;
    }
    t0 = s2 + TP(2, 2);
    t1 = s1 + TP(1, 2);
    if (TP(0, 2) > TMAT_WORST)
        t2 = s0 + TP(0, 2);  // otherwise t2 keeps what the exit block left (ref :496,519)
    if (t0 > t1) {
        if (t2 > t0) {
            s2 = t2;
            hi[2] = hi[0];
        } else
            s2 = t0;
    } else {
        if (t2 > t1) {
            s2 = t2;
            hi[2] = hi[0];
        } else {
            s2 = t1;
            hi[2] = hi[1];
        }
    }
    s2 = clampw(s2);
    best = max(best, s2);
    sc[2] = s2;
    t0 = s1 + TP(1, 1);
    t1 = s0 + TP(0, 1);
    if (t0 > t1)
        s1 = t0;
    else {
        s1 = t1;
        hi[1] = hi[0];
    }
    s1 = clampw(s1);
    best = max(best, s1);
    sc[1] = s1;
    s0 = clampw(s0 + TP(0, 0));
    best = max(best, s0);
    sc[0] = s0;
    return best;
#undef TP
}

// ref: src/hmm.c:166-304
__device__ __forceinline__ int32_t hmm_step5(const uint8_t *__restrict__ tp, const int (&ss)[5],
                                             int32_t (&sc)[5], int32_t (&hi)[5], int32_t &osc,
                                             int32_t &ohi)
{
#define TP(i, j) (-(int32_t)tp[(i) * 6 + (j)])
    int32_t sv[5], t0, t1, t2, best = WORST_SCORE;
#pragma unroll
    for (int j = 0; j < 5; ++j)
        sv[j] = sc[j] - ss[j];
    if (sv[3] > WORST_SCORE) {
        int32_t s5;
        t1 = sv[4] + TP(4, 5);
        t2 = sv[3] + TP(3, 5);
        if (t1 > t2) {
            s5 = t1;
            ohi = hi[4];
        } else {
            s5 = t2;
            ohi = hi[3];
        }
        s5 = clampw(s5);
        osc = s5;
        best = s5;
    }
    // states 4 and 3 only move when their skip source is alive (ref :191,:218); 2 always
#pragma unroll
    for (int j = 4; j >= 2; --j) {
        if (j > 2 && !(sv[j - 2] > WORST_SCORE))
            continue;
        int32_t nv;
        t0 = sv[j] + TP(j, j);
        t1 = sv[j - 1] + TP(j - 1, j);
        t2 = sv[j - 2] + TP(j - 2, j);
        if (t0 > t1) {
            if (t2 > t0) {
                nv = t2;
                hi[j] = hi[j - 2];
            } else
                nv = t0;
        } else {
            if (t2 > t1) {
                nv = t2;
                hi[j] = hi[j - 2];
            } else {
                nv = t1;
                hi[j] = hi[j - 1];
            }
        }
        nv = clampw(nv);
        best = max(best, nv);
        sc[j] = nv;
    }
    t0 = sv[1] + TP(1, 1);
    t1 = sv[0] + TP(0, 1);
    int32_t s1;
    if (t0 > t1)
        s1 = t0;
    else {
        s1 = t1;
        hi[1] = hi[0];
    }
    s1 = clampw(s1);
    best = max(best, s1);
    sc[1] = s1;
    int32_t s0 = clampw(sv[0] + TP(0, 0));
    best = max(best, s0);
    sc[0] = s0;
    return best;
#undef TP
}

template <int E>
__device__ __forceinline__ int32_t hmm_step(const uint8_t *tp, const int (&ss)[E],
                                            int32_t (&sc)[E], int32_t (&hi)[E], int32_t &osc,
                                            int32_t &ohi)
{
    if constexpr (E == 3)
        return hmm_step3(tp, ss, sc, hi, osc, ohi);
    else
        return hmm_step5(tp, ss, sc, hi, osc, ohi);
}

}  // namespace ssb
