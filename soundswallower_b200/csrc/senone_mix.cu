// senone_mix.cu -- K2: codebook normalisation + mixture-weight log-sum.
//
// Replaces ptm_mgau_codebook_norm and ptm_mgau_senone_eval of the reference
// (ref: src/ptm_mgau.c:264-295, 326-403) and fast_logmath_add
// (ref: include/soundswallower/tied_mgau_common.h:100-117).  All integer work,
// bit-exact:
//   norm_f   = max over ACTIVE codebooks of (top-1 score >> 10)
//   score_k  = min(96, norm_f - (raw_k >> 10))
//   fden_f   = fold_k  a (+) b = min(a,b) - LUT[|a-b|]   over  mixw[f][cw_k][sen] + score_k
//   senscr   = sum_f fden_f  - min over ACTIVE senones
//
// Two shapes:
//  * senone_mix_active (the aligner's default mode): per utterance only the
//    senones of its phone chain are active (plus the >255-gap bridging entries of
//    acmod_flags2list, ref: src/acmod.c:947-999).  One CTA owns one utterance for a
//    run of frames, stages the mixture-weight COLUMNS of that utterance's senones
//    in shared memory once (3*128*W bytes), and emits the int16 score of every
//    chain state -- 2 B per state-frame is all that reaches HBM.
//  * senone_dense (compallsen): CTAs own tiles of frames, keep the tile's raw scores in
//    shared memory until the frame minimum is known, and read the mixture weights as
//    coalesced 32-byte runs straight from L2 (neighbouring senones share their codebook).
#include "device.cuh"

namespace ssb {

// fast_logmath_add on the 8-bit table (ref: src/ptm_mgau.c:40-56 via logmath's shifted table): the
// reference keeps the smaller of the two and subtracts table[difference] -- it takes its
// (mlx > mly) branch only when strictly greater, which for equal values picks the same number --
// i.e. min(a, b) - lut[|a - b|]: one min, one absolute difference (VABSDIFF), one table read
__device__ __forceinline__ int logadd8(int a, int b, const uint8_t *lut)
{
    return min(a, b) - (int)lut[__sad(a, b, 0u)];
}

// Normalised, clamped scores of one codebook-stream.  For the semi-continuous scorer the
// entries from the first one beyond the stream's top-N beam onwards are not mixed in
// (ref: src/s2_semi_mgau.c:184-202); they are parked as K2_UNUSED.
constexpr int K2_UNUSED = 255;
__device__ __forceinline__ uchar4 norm_scores(const DevModel &m, int f, int nm, int4 rs)
{
    int s[4] = {min(MAX_NEG_ASCR, nm - (rs.x >> SENSCR_SHIFT)), min(MAX_NEG_ASCR, nm - (rs.y >> SENSCR_SHIFT)),
                min(MAX_NEG_ASCR, nm - (rs.z >> SENSCR_SHIFT)), min(MAX_NEG_ASCR, nm - (rs.w >> SENSCR_SHIFT))};
    if (m.kind == SSB_SCORER_SEMI) {
        const int beam = m.topn_beam[f];
        bool cut = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            cut = cut || (beam != 0 && s[k] > beam);
            if (cut)
                s[k] = K2_UNUSED;
        }
    }
    return make_uchar4((unsigned char)s[0], (unsigned char)s[1], (unsigned char)s[2],
                       (unsigned char)s[3]);
}

constexpr int K2_THREADS = 256;
constexpr int K2_WARPS = K2_THREADS / 32;
constexpr int K2_CS_PER_LANE = 4;
// resident CTAs per SM the active-list kernel is compiled for: left alone the compiler takes 120
// registers (2 CTAs per SM, 6.07 ms on config #2); 3 -> 80 registers, no spills, 5.41 ms; 4 -> 64
// registers with spills, 5.70 ms; 5 -> 6.09 ms
#ifndef K2_MIN_BLOCKS
#define K2_MIN_BLOCKS 3
#endif  // codebook-streams per lane: CS <= 128

// ---------------------------------------------------------------- active lists
// CTA = (utterance, run of frames); the mixture-weight columns of the utterance's senone
// union are staged once.  Then every WARP takes frames on its own (t = t_begin + warp,
// +8, ...): lanes over codebook-streams for the normaliser, lanes over active senones for
// the mixing, lanes over chain states for the gather -- only __syncwarp inside the loop.
// PTM4 = the bundled models' case fixed at compile time (PTM scorer, top-4, 3 streams x 128): no
// unused-entry test and no trip-count test in the innermost loop (they cost 2.8 of 11 ms on
// config #2), no runtime divisions by the stream count.
template <bool STAGED, bool PTM4>
__global__ void __launch_bounds__(K2_THREADS, K2_MIN_BLOCKS)
senone_mix_active_kernel(DevModel m, DevPlan p, const int4 *__restrict__ tn_s,
                         const uchar4 *__restrict__ tn_c, int64_t G, int W, int chunk,
                         int16_t *__restrict__ chain_scr)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int u = blockIdx.y;
    const int64_t g0 = p.frame_off[u];
    const int T = (int)(p.frame_off[u + 1] - g0);
    const int t_begin = blockIdx.x * chunk;
    if (t_begin >= T)
        return;
    const int t_end = min(T, t_begin + chunk);
    // PTM4 also fixes the stream count at 3 (the launcher checks): `% NF`, `/ NF` and the loops
    // over streams are compile-time then
    const int NF = PTM4 ? 3 : m.n_feat;
    const int CS = m.n_mgau * NF;
    const int ND = PTM4 ? 128 : m.n_density, N = m.topn;
    const int us0 = p.us_off[u], n_us = p.us_off[u + 1] - us0;
    const int64_t ph0 = p.phone_off[u];
    const int ns = (int)(p.phone_off[u + 1] - ph0) * m.n_emit;
    const uint16_t *st_slot = p.st_slot + ph0 * m.n_emit;
    const int e0 = p.ep_off[u], e1 = p.ep_off[u + 1];
    if (ns == 0 || e0 == e1)
        return;  // nothing to score for this utterance (uniform for the whole CTA)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // shared layout
    const int CSP = (CS + 3) & ~3;
    uint8_t *lut = smem;                                                 // 256
    uchar4 *wt_s = reinterpret_cast<uchar4 *>(smem + 256) + warp * CSP;  // [warps][CSP]
    uchar4 *wt_c = reinterpret_cast<uchar4 *>(smem + 256) + (K2_WARPS + warp) * CSP;
    int16_t *wscr = reinterpret_cast<int16_t *>(smem + 256 + (size_t)2 * K2_WARPS * CSP * 4) + warp * W;
    uint8_t *ucb = smem + 256 + (size_t)2 * K2_WARPS * CSP * 4 + (size_t)K2_WARPS * W * 2;  // [W]
    uint8_t *mw = ucb + ((W + 15) & ~15);                                // [NF*ND][W] staged columns

    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        lut[i] = m.lut8[i];
    for (int i = threadIdx.x; i < n_us; i += blockDim.x)
        ucb[i] = m.sen2cb[p.usen[us0 + i]];
    if (STAGED) {
        const int rows = NF * ND;
        for (int idx = threadIdx.x; idx < rows * n_us; idx += blockDim.x) {
            int r = idx / n_us, j = idx - r * n_us;
            mw[r * W + j] = m.mixw[(int64_t)r * m.n_sen + p.usen[us0 + j]];
        }
    }
    __syncthreads();

    int e = e0;
    uint32_t okmask = 0;  // bit it: codebook of this lane's it-th codebook-stream is active
    int sl0 = 0, na = 0;
    bool fresh = true;
    int b_lo = 0, b_hi = -1;  // banded layout: phones [b_lo, b_hi] are evaluated on the warp's frame
    const int per_round = 32 / m.n_emit;
    const int lane_ph = (int)((lane + 0.5f) / (float)m.n_emit), lane_j = lane - lane_ph * m.n_emit;
    if (p.banded) {
        // first frame of this warp: binary search for the last phone entered by then
        const int np = ns / m.n_emit, t = t_begin + warp;
        const int32_t *enter = p.enter_plan + ph0, *ef = p.ef + ph0;
        int a = 0, b = np;
        while (a < b) {
            const int mid = (a + b) >> 1;
            const int en = enter[mid];
            if (en >= 0 && en <= t)
                a = mid + 1;
            else
                b = mid;
        }
        b_hi = a - 1;
        a = 0;
        b = b_hi + 1;
        while (a < b) {
            const int mid = (a + b) >> 1;
            if (max(enter[mid], ef[mid]) >= t)
                b = mid;
            else
                a = mid + 1;
        }
        b_lo = a;
    }
    for (int t = t_begin + warp; t < t_end; t += K2_WARPS) {
        while (e + 1 < e1 && p.ep_start[e + 1] <= t) {
            ++e;
            fresh = true;
        }
        if (fresh) {
            okmask = 0;
#pragma unroll
            for (int it = 0; it < K2_CS_PER_LANE; ++it) {
                const int cs = lane + 32 * it;
                if (cs < CS) {
                    const int cb = cs / NF;
                    okmask |= ((p.ep_cbmask[(int64_t)e * 8 + (cb >> 5)] >> (cb & 31)) & 1u) << it;
                }
            }
            sl0 = p.ep_slot_off[e];
            na = p.ep_slot_off[e + 1] - sl0;
            fresh = false;
        }
        // A: raw top-N of the active codebooks, per-stream normaliser over them
        int4 rs[K2_CS_PER_LANE];
        uchar4 rc[K2_CS_PER_LANE];
        bool ok[K2_CS_PER_LANE];
        int nm[SSB_MAX_FEAT];
#pragma unroll
        for (int f = 0; f < SSB_MAX_FEAT; ++f)
            nm[f] = WORST_SCORE;
#pragma unroll
        for (int it = 0; it < K2_CS_PER_LANE; ++it) {
            const int cs = lane + 32 * it;
            ok[it] = (okmask >> it) & 1u;
            if (ok[it]) {
                const int64_t g = (int64_t)cs * G + g0 + t;
                rs[it] = tn_s[g];
                rc[it] = tn_c[g];
            }
        }
#pragma unroll
        for (int it = 0; it < K2_CS_PER_LANE; ++it)
            if (ok[it]) {
                const int f = (lane + 32 * it) % NF;
                const int v = rs[it].x >> SENSCR_SHIFT;
#pragma unroll
                for (int ff = 0; ff < SSB_MAX_FEAT; ++ff)
                    if (ff == f)
                        nm[ff] = max(nm[ff], v);
            }
#pragma unroll
        for (int f = 0; f < SSB_MAX_FEAT; ++f)
            for (int o = 16; o > 0; o >>= 1)
                nm[f] = max(nm[f], __shfl_xor_sync(0xffffffffu, nm[f], o));
        // B: normalise, clamp, park in this warp's tile; clear the warp's score row
#pragma unroll
        for (int it = 0; it < K2_CS_PER_LANE; ++it)
            if (ok[it]) {
                const int cs = lane + 32 * it, f = cs % NF;
                int n0 = nm[0];
#pragma unroll
                for (int ff = 1; ff < SSB_MAX_FEAT; ++ff)
                    if (ff == f)
                        n0 = nm[ff];
                if (PTM4)
                    wt_s[cs] = make_uchar4((unsigned char)min(MAX_NEG_ASCR, n0 - (rs[it].x >> SENSCR_SHIFT)),
                                           (unsigned char)min(MAX_NEG_ASCR, n0 - (rs[it].y >> SENSCR_SHIFT)),
                                           (unsigned char)min(MAX_NEG_ASCR, n0 - (rs[it].z >> SENSCR_SHIFT)),
                                           (unsigned char)min(MAX_NEG_ASCR, n0 - (rs[it].w >> SENSCR_SHIFT)));
                else
                    wt_s[cs] = norm_scores(m, f, n0, rs[it]);
                wt_c[cs] = rc[it];
            }
        for (int i = lane; i < W / 2; i += 32)
            reinterpret_cast<uint32_t *>(wscr)[i] = 0u;
        __syncwarp();
        // C: active senones of this frame's epoch
        int local_best = INT32_MAX;
        for (int i = lane; i < na; i += 32) {
            const int slot = p.ep_slot[sl0 + i];
            const int cb = ucb[slot];
            int ascore = 0;
            for (int f = 0; f < NF; ++f) {
                const uchar4 sv = wt_s[cb * NF + f];
                const uchar4 cv = wt_c[cb * NF + f];
                const int sc[4] = {sv.x, sv.y, sv.z, sv.w};
                const int cw[4] = {cv.x, cv.y, cv.z, cv.w};
                int fden = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (PTM4 || (k < N && sc[k] != K2_UNUSED)) {
                        int w;
                        if (STAGED)
                            w = mw[(f * ND + cw[k]) * W + slot];
                        else
                            w = m.mixw[(int64_t)(f * ND + cw[k]) * m.n_sen + p.usen[us0 + slot]];
                        int v = w + sc[k];
                        fden = k == 0 ? v : logadd8(fden, v, lut);
                    }
                }
                ascore += fden;
            }
            wscr[slot] = (int16_t)ascore;
            local_best = min(local_best, ascore);
        }
        for (int o = 16; o > 0; o >>= 1)
            local_best = min(local_best, __shfl_xor_sync(0xffffffffu, local_best, o));
        __syncwarp();
        // D: gather to chain states, subtract the frame's best (ref :398-400; the
        // semi-continuous scorer does not normalise over senones)
        {
            const int16_t b16 = (!PTM4 && m.kind == SSB_SCORER_SEMI) ? (int16_t)0 : (int16_t)local_best;
            if (!p.banded) {
                int16_t *dst = chain_scr + p.scr_off[u] + (int64_t)t * ns;
                for (int si = lane; si < ns; si += 32)
                    dst[si] = (int16_t)(wscr[st_slot[si]] - b16);
            } else {
                // only the phones the chain Viterbi evaluates on frame t: enter[i] <= t <= last[i];
                // both bounds are non-decreasing along the entered prefix of the chain, and the
                // warp's frames only go forward: two running pointers
                const int E = m.n_emit, np = ns / E;
                const int32_t *enter = p.enter_plan + ph0, *ef = p.ef + ph0;
                while (b_hi + 1 < np) {
                    const int en = enter[b_hi + 1];
                    if (en < 0 || en > t)
                        break;
                    ++b_hi;
                }
                while (b_lo < b_hi && max(enter[b_lo], ef[b_lo]) < t)
                    ++b_lo;
                const int a = b_lo, hiq = b_hi;
                // lanes = (phone, state) pairs, 32 / E phones per round; virtual block bases
                // (scr_boff - enter * E) make the address base + t * E + j
                for (int ph = a + lane_ph; ph <= hiq; ph += per_round)
                    if (lane < per_round * E) {
                        const int si = ph * E + lane_j;
                        chain_scr[p.scr_boff[ph0 + ph] + (int64_t)t * E + lane_j] =
                            (int16_t)(wscr[st_slot[si]] - b16);
                    }
            }
        }
        __syncwarp();
    }
}

static size_t k2_active_smem(const DevModel &m, int W, bool staged)
{
    int CSP = (m.n_mgau * m.n_feat + 3) & ~3;
    size_t b = 256 + (size_t)2 * K2_WARPS * CSP * 4 + (size_t)K2_WARPS * W * 2 + ((W + 15) & ~15);
    if (staged)
        b += (size_t)m.n_feat * m.n_density * W;
    return b + 16;
}

int launch_senone_mix_active(const DevModel &m, const DevPlan &p, const int4 *tn_score,
                             const uchar4 *tn_cw, int64_t n_frames, int max_union,
                             int max_frames_per_utt, int16_t *chain_scr, cudaStream_t st)
{
    if (p.n_utts == 0 || n_frames == 0)
        return 0;
    if (m.n_mgau * m.n_feat > 32 * K2_CS_PER_LANE) {
        set_error("senone_mix: %d codebook-streams exceed the per-warp tile (max %d)",
                  m.n_mgau * m.n_feat, 32 * K2_CS_PER_LANE);
        return -1;
    }
    int W = (max_union + 3) & ~3;
    if (W < 4)
        W = 4;
    bool staged = k2_active_smem(m, W, true) <= 200 * 1024;
    size_t smem = k2_active_smem(m, W, staged);
    // Frames per CTA: staging the mixture-weight columns costs ~NF*ND*W byte gathers per CTA, so
    // a CTA keeps an utterance as long as the grid still fills the machine a few times over.
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int64_t want_ctas = (int64_t)sms * 12;
    int chunk = (int)((n_frames + want_ctas - 1) / want_ctas);
    chunk = (max(chunk, 64) + K2_WARPS - 1) / K2_WARPS * K2_WARPS;
    dim3 grid((max_frames_per_utt + chunk - 1) / chunk, p.n_utts);
    const bool ptm4 = m.kind == SSB_SCORER_PTM && m.topn == 4 && m.n_feat == 3 && m.n_density == 128;
#define SSB_K2(ST, P4)                                                                              \
    do {                                                                                            \
        SSB_DYN_SMEM((senone_mix_active_kernel<ST, P4>), smem);     \
        senone_mix_active_kernel<ST, P4>                                                            \
            <<<grid, K2_THREADS, smem, st>>>(m, p, tn_score, tn_cw, n_frames, W, chunk, chain_scr); \
    } while (0)
    if (staged) {
        if (ptm4) SSB_K2(true, true); else SSB_K2(true, false);
    } else {
        if (ptm4) SSB_K2(false, true); else SSB_K2(false, false);
    }
#undef SSB_K2
    SSB_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// ---------------------------------------------------------------- dense (compallsen)
// Every senone of every frame: what acmod_score returns with compallsen (ref: src/acmod.c:
// 822-860), best score already subtracted.  CTA = tile of F frames.  Phase A parks the
// normalised top-N of all codebook-streams of the tile in shared memory (normaliser = max over
// ALL codebooks); phase B gives every thread a run of senone ids: neighbouring senones share
// their codebook almost always (ids are grouped by phone), so a warp reads the same 12
// mixture-weight rows at 32 consecutive bytes -- coalesced straight out of L2, no staging;
// the raw scores of the tile stay in shared memory until the frame's minimum is known, then
// one coalesced int16 row per frame goes to HBM (2 B per score is all the HBM traffic).
constexpr int KD_THREADS = 256;

__global__ void __launch_bounds__(KD_THREADS)
senone_dense_kernel(DevModel m, const int4 *__restrict__ tn_s, const uchar4 *__restrict__ tn_c,
                    int64_t G, int64_t g0, int64_t n, int F, int16_t *__restrict__ dense)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int CS = m.n_mgau * m.n_feat, NF = m.n_feat, ND = m.n_density, N = m.topn;
    const int n_sen = m.n_sen;
    const int CSP = (CS + 3) & ~3;
    uint8_t *lut = smem;                                          // 256
    int *norm = reinterpret_cast<int *>(smem + 256);              // [F][SSB_MAX_FEAT]
    int *best = norm + F * SSB_MAX_FEAT;                          // [F]
    uchar4 *wt_s = reinterpret_cast<uchar4 *>(best + F);          // [F][CSP]
    uchar4 *wt_c = wt_s + (size_t)F * CSP;                        // [F][CSP]
    uint8_t *s2c = reinterpret_cast<uint8_t *>(wt_c + (size_t)F * CSP);  // [n_sen]
    int16_t *scr = reinterpret_cast<int16_t *>(s2c + ((n_sen + 15) & ~15));  // [F][n_sen]
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        lut[i] = m.lut8[i];
    for (int i = threadIdx.x; i < n_sen; i += blockDim.x)
        s2c[i] = m.sen2cb[i];
    for (int64_t base = (int64_t)blockIdx.x * F; base < n; base += (int64_t)gridDim.x * F) {
        const int nf = (int)min((int64_t)F, n - base);
        __syncthreads();
        for (int i = threadIdx.x; i < F * SSB_MAX_FEAT; i += blockDim.x)
            norm[i] = WORST_SCORE;
        if (threadIdx.x < F)
            best[threadIdx.x] = INT32_MAX;
        __syncthreads();
        // A: raw top-N -> per-stream normaliser over all codebooks
        for (int idx = threadIdx.x; idx < nf * CS; idx += blockDim.x) {
            const int fr = idx % nf, cs = idx / nf;
            const int v = tn_s[(int64_t)cs * G + g0 + base + fr].x >> SENSCR_SHIFT;
            atomicMax(&norm[fr * SSB_MAX_FEAT + cs % NF], v);
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < nf * CS; idx += blockDim.x) {
            const int fr = idx % nf, cs = idx / nf;
            const int64_t g = (int64_t)cs * G + g0 + base + fr;
            const int4 rs = tn_s[g];
            const int nm = norm[fr * SSB_MAX_FEAT + cs % NF];
            wt_s[fr * CSP + cs] = norm_scores(m, cs % NF, nm, rs);
            wt_c[fr * CSP + cs] = tn_c[g];
        }
        __syncthreads();
        // B: every senone of every frame of the tile
        for (int fr = 0; fr < nf; ++fr) {
            int local_best = INT32_MAX;
            for (int sen = threadIdx.x; sen < n_sen; sen += blockDim.x) {
                const int cb = s2c[sen];
                int ascore = 0;
                if (NF == 3 && N == 4 && m.kind == SSB_SCORER_PTM) {
                    // the bundled shape: all 12 weight loads in flight before any arithmetic
                    uchar4 sv[3], cv[3];
                    int w[12];
#pragma unroll
                    for (int f = 0; f < 3; ++f) {
                        sv[f] = wt_s[fr * CSP + cb * 3 + f];
                        cv[f] = wt_c[fr * CSP + cb * 3 + f];
                    }
                    const uint32_t base_s = (uint32_t)sen;
#pragma unroll
                    for (int f = 0; f < 3; ++f) {
                        const uint8_t *row = m.mixw + (uint32_t)(f * ND) * (uint32_t)n_sen + base_s;
                        w[4 * f + 0] = __ldg(row + (uint32_t)cv[f].x * (uint32_t)n_sen);
                        w[4 * f + 1] = __ldg(row + (uint32_t)cv[f].y * (uint32_t)n_sen);
                        w[4 * f + 2] = __ldg(row + (uint32_t)cv[f].z * (uint32_t)n_sen);
                        w[4 * f + 3] = __ldg(row + (uint32_t)cv[f].w * (uint32_t)n_sen);
                    }
#pragma unroll
                    for (int f = 0; f < 3; ++f) {
                        int fden = w[4 * f] + sv[f].x;
                        fden = logadd8(fden, w[4 * f + 1] + sv[f].y, lut);
                        fden = logadd8(fden, w[4 * f + 2] + sv[f].z, lut);
                        fden = logadd8(fden, w[4 * f + 3] + sv[f].w, lut);
                        ascore += fden;
                    }
                } else {
                    for (int f = 0; f < NF; ++f) {
                        const uchar4 sv = wt_s[fr * CSP + cb * NF + f];
                        const uchar4 cv = wt_c[fr * CSP + cb * NF + f];
                        const int sc[4] = {sv.x, sv.y, sv.z, sv.w};
                        const int cw[4] = {cv.x, cv.y, cv.z, cv.w};
                        const uint8_t *row = m.mixw + (int64_t)f * ND * n_sen + sen;
                        int fden = 0;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (k < N && sc[k] != K2_UNUSED) {
                                const int v = (int)__ldg(row + (int64_t)cw[k] * n_sen) + sc[k];
                                fden = k == 0 ? v : logadd8(fden, v, lut);
                            }
                        }
                        ascore += fden;
                    }
                }
                scr[(size_t)fr * n_sen + sen] = (int16_t)ascore;
                local_best = min(local_best, ascore);
            }
            for (int o = 16; o > 0; o >>= 1)
                local_best = min(local_best, __shfl_xor_sync(0xffffffffu, local_best, o));
            if ((threadIdx.x & 31) == 0)
                atomicMin(&best[fr], local_best);
        }
        __syncthreads();
        // C: subtract the frame's best (ref: src/ptm_mgau.c:398-400), one coalesced row per frame
        for (int fr = 0; fr < nf; ++fr) {
            const int16_t b16 = m.kind == SSB_SCORER_SEMI ? (int16_t)0 : (int16_t)best[fr];
            int16_t *dst = dense + (base + fr) * n_sen;
            for (int sen = threadIdx.x; sen < n_sen; sen += blockDim.x)
                dst[sen] = (int16_t)(scr[(size_t)fr * n_sen + sen] - b16);
        }
    }
}

// dense [n][n_sen] = final senone scores of frames [g0, g0 + n)
int launch_senone_mix_all(const DevModel &m, const int4 *tn_score, const uchar4 *tn_cw,
                          int64_t n_frames_total, int64_t g0, int64_t n, int16_t *dense,
                          cudaStream_t st)
{
    if (n == 0)
        return 0;
    const int CSP = (m.n_mgau * m.n_feat + 3) & ~3;
    const size_t fixed = 256 + ((m.n_sen + 15) & ~15) + 16;
    auto need = [&](int F) {
        return fixed + (size_t)F * (SSB_MAX_FEAT + 1) * 4 + (size_t)2 * F * CSP * 4 + (size_t)F * m.n_sen * 2;
    };
    int F = 8;
    while (F > 1 && need(F) > 56 * 1024)
        F >>= 1;
    if (need(F) > 200 * 1024) {
        set_error("dense senone scoring: %d senones do not fit the shared-memory tile", m.n_sen);
        return -1;
    }
    const size_t smem = need(F);
    SSB_DYN_SMEM((senone_dense_kernel), smem);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int64_t tiles = (n + F - 1) / F;
    const int grid = (int)(tiles < (int64_t)sms * 8 ? tiles : (int64_t)sms * 8);
    senone_dense_kernel<<<grid, KD_THREADS, smem, st>>>(m, tn_score, tn_cw, n_frames_total, g0, n, F,
                                                        dense);
    SSB_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// chain_scr[u][t][si] = dense[g][senone(si)] for utterances [u0,u1) whose frames start at g0
__global__ void gather_chain_kernel(DevModel m, DevPlan p, const int16_t *__restrict__ dense,
                                    int u0, int64_t g0, int16_t *__restrict__ chain_scr)
{
    const int u = u0 + blockIdx.y;
    const int64_t f0 = p.frame_off[u];
    const int T = (int)(p.frame_off[u + 1] - f0);
    const int64_t ph0 = p.phone_off[u];
    const int ns = (int)(p.phone_off[u + 1] - ph0) * m.n_emit;
    const int E = m.n_emit;
    for (int t = blockIdx.x; t < T; t += gridDim.x) {
        const int64_t row = f0 + t - g0;
        int16_t *dst = chain_scr + p.scr_off[u] + (int64_t)t * ns;
        for (int si = threadIdx.x; si < ns; si += blockDim.x) {
            int ph = si / E, j = si - ph * E;
            int sen = m.sseq[(int64_t)p.ssid[ph0 + ph] * E + j];
            dst[si] = dense[row * m.n_sen + sen];
        }
    }
}

int launch_gather_chain(const DevModel &m, const DevPlan &p, const int16_t *dense, int u0, int u1,
                        int64_t g0, int16_t *chain_scr, cudaStream_t st)
{
    if (u1 <= u0)
        return 0;
    dim3 grid(64, u1 - u0);
    gather_chain_kernel<<<grid, 128, 0, st>>>(m, p, dense, u0, g0, chain_scr);
    SSB_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// ---------------------------------------------------------------- single frame (vtable)
// One CTA.  Works on history slot `slot` in place like the reference does: active
// codebooks are normalised in place when do_norm, and a senone whose codebook is not
// active forces that codebook's entries to 96 (ref :353-364).
__global__ void __launch_bounds__(512)
frame_senones_kernel(DevModel m, FrameHist h, int slot, int do_norm,
                     const uint16_t *__restrict__ act_sen, int n_act, int compallsen,
                     int16_t *__restrict__ senscr)
{
    __shared__ int s_norm[SSB_MAX_FEAT];
    __shared__ int s_best;
    __shared__ uint8_t lut[256];
    const int NF = m.n_feat, ND = m.n_density, N = m.topn;
    int4 *hs = h.score[slot];
    const uchar4 *hc = h.cw[slot];
    const uint8_t *act = h.act[slot];
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        lut[i] = m.lut8[i];
    if (threadIdx.x < NF)
        s_norm[threadIdx.x] = WORST_SCORE;
    if (threadIdx.x == 0)
        s_best = INT32_MAX;
    for (int i = threadIdx.x; i < m.n_sen; i += blockDim.x)
        senscr[i] = 0;
    __syncthreads();
    if (do_norm) {
        for (int cs = threadIdx.x; cs < m.n_mgau * NF; cs += blockDim.x)
            if (act[cs / NF])
                atomicMax(&s_norm[cs % NF], hs[cs].x >> SENSCR_SHIFT);
        __syncthreads();
        for (int cs = threadIdx.x; cs < m.n_mgau * NF; cs += blockDim.x)
            if (act[cs / NF]) {
                const uchar4 q = norm_scores(m, cs % NF, s_norm[cs % NF], hs[cs]);
                hs[cs] = make_int4(q.x, q.y, q.z, q.w);
            }
        __syncthreads();
    }
    const int n = compallsen ? m.n_sen : n_act;
    // pass 1: knock out codebooks that are referenced but not active
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int sen = compallsen ? i : act_sen[i];
        int cb = m.sen2cb[sen];
        if (!act[cb])
            for (int f = 0; f < NF; ++f)
                hs[cb * NF + f] = make_int4(MAX_NEG_ASCR, MAX_NEG_ASCR, MAX_NEG_ASCR, MAX_NEG_ASCR);
    }
    __syncthreads();
    int local_best = INT32_MAX;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int sen = compallsen ? i : act_sen[i];
        int cb = m.sen2cb[sen];
        int ascore = 0;
        for (int f = 0; f < NF; ++f) {
            int4 sv = hs[cb * NF + f];
            uchar4 cv = hc[cb * NF + f];
            const int sc[4] = {sv.x, sv.y, sv.z, sv.w};
            const int cw[4] = {cv.x, cv.y, cv.z, cv.w};
            int fden = 0;
            for (int k = 0; k < N && sc[k] != K2_UNUSED; ++k) {
                int v = m.mixw[(int64_t)(f * ND + cw[k]) * m.n_sen + sen] + sc[k];
                fden = k == 0 ? v : logadd8(fden, v, lut);
            }
            ascore += fden;
        }
        senscr[sen] = (int16_t)ascore;
        local_best = min(local_best, ascore);
    }
    if (local_best != INT32_MAX)
        atomicMin(&s_best, local_best);
    __syncthreads();
    // ref: src/ptm_mgau.c:398-400; s2_semi leaves the sums as they are (and unlisted senones 0)
    const int16_t b = m.kind == SSB_SCORER_SEMI ? (int16_t)0 : (int16_t)s_best;
    for (int i = threadIdx.x; i < m.n_sen; i += blockDim.x)
        senscr[i] = (int16_t)(senscr[i] - b);
}

int launch_frame_senones(const DevModel &m, const FrameHist &h, int slot, int do_norm,
                         const uint16_t *act_sen, int n_act, int compallsen, int16_t *senscr,
                         cudaStream_t st)
{
    frame_senones_kernel<<<1, 512, 0, st>>>(m, h, slot, do_norm, act_sen, n_act, compallsen, senscr);
    SSB_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

}  // namespace ssb
