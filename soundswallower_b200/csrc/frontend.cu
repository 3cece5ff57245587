// frontend.cu -- batched acoustic frontend on the device: PCM -> mel spectrum -> noise
// tracker -> cepstra -> CMN -> dynamic features, whole utterances at a time.
//
// ref: src/fe_interface.c:83-178, 270-350 (parameters), :352-360, 379-391, 578-713 (framing),
//      src/fe_sigproc.c:70-236 (filters, DCT basis, lifter), :238-321 (pre-emphasis, window),
//      :447-594 (real FFT, power and mel spectrum), :596-715 (log, DCT, lifter),
//      src/fe_noise.c:110-186, 266-327 (noise tracker), src/cmn.c:159-229 (batch CMN),
//      src/feat.c:589-632, 978-1007 (1s_c_d_dd with edge replication).
//
// The reference streams an utterance through small buffers; here a frame is addressed
// directly: frame t of an utterance covers samples [t*shift, t*shift+frame_size), the last
// frame is the partial remainder (zero padded), and the pre-emphasis carry is simply the
// sample before the frame.  All float64/float32 operations are issued as explicit IEEE
// intrinsics (__dadd_rn, __dmul_rn, ...) in the reference's order so that nothing is fused;
// the one operation that is not reproducible to the bit is the natural logarithm.
//
// Kernels (data stays in HBM between them):
//   fe_melspec_kernel  warp per frame: window, 2^m-point real FFT in shared memory, power
//                      spectrum, mel filters                       -> mel [frames][nfilt] f64
//   fe_noise_kernel    warp per utterance, sequential over frames (the tracker is a non-linear
//                      recurrence per band), in place on mel
//   fe_cepstrum_kernel 64 frames per CTA: log, DCT, lifter         -> mfcc [frames][ncep] f32
//   fe_cmn_kernel      warp per utterance: the reference's sequential float32 sums
//   fe_feat_kernel     thread per (frame, coefficient): CMN + deltas -> feat [frames][3*ncep]
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "host_util.cuh"

using namespace ssb;

namespace {

constexpr int kMaxFft = 2048;
constexpr int kMaxFilt = 64;
constexpr int kMaxCep = 32;
constexpr int kCepFrames = 64;
constexpr int kMelSmem = 64 * 1024;

struct FeDev {
    int frame_size, frame_shift, fft_size, fft_order, nfilt, ncep;
    int remove_dc, transform, has_lifter, cmn, varnorm;
    float alpha, sqrt_inv_n, sqrt_inv_2n;
    const double *hamming, *ccc, *sss;
    const int *spec_start, *filt_start, *filt_width;
    const float *coeffs, *mel_cosine, *lifter;
};

__device__ __forceinline__ float pcm_sample(const void *pcm, int enc, int64_t i)
{
    if (enc == SSB_PCM_INT16)
        return (float)static_cast<const int16_t *>(pcm)[i];
    return __fmul_rn(static_cast<const float *>(pcm)[i], 32768.0f);
}

// largest u with off[u] <= x (off non-decreasing, off[0] = 0)
__device__ __forceinline__ int find_utt(const int64_t *off, int n, int64_t x)
{
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= x)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

// ------------------------------------------------------------------ mel spectrum
__global__ void __launch_bounds__(256)
fe_melspec_kernel(FeDev fe, const void *__restrict__ pcm, int enc,
                  const int64_t *__restrict__ samp_off, const int64_t *__restrict__ frame_off,
                  int n_utts, int64_t fr_base, int64_t n_frames, double *__restrict__ mel)
{
    extern __shared__ double fe_sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t fr = fr_base + (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;  // frames [fr_base, n_frames)
    if (fr >= n_frames)
        return;
    const int n = fe.fft_size, m = fe.fft_order, fs = fe.frame_size;
    double *x = fe_sm + (size_t)warp * 2 * n, *y = x + n;
    const int u = find_utt(frame_off, n_utts, fr);
    const int64_t t = fr - frame_off[u];
    const int64_t s0 = samp_off[u] + t * fe.frame_shift;
    const int64_t left = samp_off[u + 1] - s0;
    const int len = left < fs ? (int)left : fs;

    // pre-emphasis (ref: fe_sigproc.c:238-247), zero padding
    const double alpha = (double)fe.alpha;
    for (int i = lane; i < n; i += 32) {
        double v = 0.0;
        if (i < len) {
            const float s = pcm_sample(pcm, enc, s0 + i);
            if (fe.alpha != 0.0f) {
                const float p = (i > 0 || t > 0) ? pcm_sample(pcm, enc, s0 + i - 1) : 0.0f;
                v = __dsub_rn((double)s, __dmul_rn((double)p, alpha));
            } else
                v = (double)s;
        }
        y[i] = v;
    }
    __syncwarp();
    if (fe.remove_dc) {  // ref: fe_sigproc.c:277-285 -- one running sum, in order
        double mean = 0.0;
        if (lane == 0) {
            for (int i = 0; i < fs; ++i)
                mean = __dadd_rn(mean, y[i]);
            mean = __ddiv_rn(mean, (double)fs);
        }
        mean = __shfl_sync(0xffffffffu, mean, 0);
        for (int i = lane; i < fs; i += 32)
            y[i] = __dsub_rn(y[i], mean);
        __syncwarp();
    }
    // Hamming window (ref: fe_sigproc.c:287-290), stored in bit-reversed order
    const int half = fs >> 1;
    for (int i = lane; i < n; i += 32) {
        double v = y[i];
        if (i < fs) {
            const int k = i < half ? i : fs - 1 - i;
            if (k < half)
                v = __dmul_rn(v, fe.hamming[k]);
        }
        x[__brev((unsigned)i) >> (32 - m)] = v;
    }
    __syncwarp();
    // real FFT, the reference's butterflies (fe_sigproc.c:460-550) spread over the lanes
    for (int p = lane; p < (n >> 1); p += 32) {
        const double a = x[2 * p], b = x[2 * p + 1];
        x[2 * p] = __dadd_rn(a, b);
        x[2 * p + 1] = __dsub_rn(a, b);
    }
    __syncwarp();
    for (int k = 1; k < m; ++k) {
        const int h = 1 << k, q = h >> 1, tw = m - k - 1;
        for (int w = lane; w < (n >> 2); w += 32) {
            const int j = w & (q - 1), i = (w >> (k - 1)) << (k + 1);
            if (j == 0) {
                const double a = x[i], b = x[i + h];
                x[i] = __dadd_rn(a, b);
                x[i + h] = __dsub_rn(a, b);
                x[i + h + q] = -x[i + h + q];
            } else {
                const int i1 = i + j, i2 = i + h - j, i3 = i + h + j, i4 = i + 2 * h - j;
                const double cc = __ldg(fe.ccc + (j << tw)), ss = __ldg(fe.sss + (j << tw));
                const double x1 = x[i1], x2 = x[i2], x3 = x[i3], x4 = x[i4];
                const double t1 = __dadd_rn(__dmul_rn(x3, cc), __dmul_rn(x4, ss));
                const double t2 = __dsub_rn(__dmul_rn(x3, ss), __dmul_rn(x4, cc));
                x[i4] = __dsub_rn(x2, t2);
                x[i3] = __dsub_rn(-x2, t2);
                x[i2] = __dsub_rn(x1, t1);
                x[i1] = __dadd_rn(x1, t1);
            }
        }
        __syncwarp();
    }
    // power spectrum (ref: fe_sigproc.c:552-577; bin n/2 counts its real part twice)
    for (int j = lane; j <= (n >> 1); j += 32) {
        const double re = x[j];
        double pw = __dmul_rn(re, re);
        if (j > 0) {
            const double im = x[n - j];
            pw = __dadd_rn(pw, __dmul_rn(im, im));
        }
        y[j] = pw;
    }
    __syncwarp();
    // mel filters (ref: fe_sigproc.c:579-594)
    for (int f = lane; f < fe.nfilt; f += 32) {
        const int ss = fe.spec_start[f], cs = fe.filt_start[f], wd = fe.filt_width[f];
        double acc = 0.0;
        for (int i = 0; i < wd; ++i)
            acc = __dadd_rn(acc, __dmul_rn(y[ss + i], (double)__ldg(fe.coeffs + cs + i)));
        mel[fr * fe.nfilt + f] = acc;
    }
}

// Same arithmetic, register-resident start: for fft_size = 32 * R (R = 8, 16, 32) lane l owns the
// R consecutive elements [l*R, l*R + R) of the bit-reversed frame.  Their source samples are
// rev_LR(e) * 32 + rev5(l): each load instruction reads 32 consecutive samples across the warp,
// and the first LR butterfly stages (spans up to R) never leave the registers.  Only the last
// five stages go through shared memory, in a layout padded by one double per 16 so that both
// the lane-contiguous hand-over and the strided butterflies are (nearly) conflict-free.
__device__ __forceinline__ int fe_pad(int i) { return i + (i >> 4); }

template <int LR>
__global__ void __launch_bounds__(256)
fe_melspec_reg_kernel(FeDev fe, const void *__restrict__ pcm, int enc,
                      const int64_t *__restrict__ samp_off,
                      const int64_t *__restrict__ frame_off, int n_utts, int64_t fr_base,
                      int64_t n_frames, double *__restrict__ mel)
{
    constexpr int R = 1 << LR, NN = 32 * R, M = 5 + LR, LDW = NN + NN / 16;
    extern __shared__ double fe_sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t fr = fr_base + (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;  // frames [fr_base, n_frames)
    if (fr >= n_frames)
        return;
    double *x = fe_sm + (size_t)warp * LDW;
    const int fs = fe.frame_size, half = fs >> 1;
    const int u = find_utt(frame_off, n_utts, fr);
    const int64_t t = fr - frame_off[u];
    const int64_t s0 = samp_off[u] + t * fe.frame_shift;
    const int64_t left = samp_off[u + 1] - s0;
    const int len = left < fs ? (int)left : fs;
    const double alpha = (double)fe.alpha;
    const int rl = (int)(__brev((unsigned)lane) >> 27);

    double v[R];
#pragma unroll
    for (int e = 0; e < R; ++e) {
        const int pos = (int)((__brev((unsigned)e) >> (32 - LR)) << 5) | rl;
        double a = 0.0;
        if (pos < len) {
            const float s = pcm_sample(pcm, enc, s0 + pos);
            if (fe.alpha != 0.0f) {
                const float p = (pos > 0 || t > 0) ? pcm_sample(pcm, enc, s0 + pos - 1) : 0.0f;
                a = __dsub_rn((double)s, __dmul_rn((double)p, alpha));
            } else
                a = (double)s;
        }
        if (pos < fs) {
            const int k = pos < half ? pos : fs - 1 - pos;
            if (k < half)
                a = __dmul_rn(a, __ldg(fe.hamming + k));
        }
        v[e] = a;
    }
    // stages 0 .. LR-1 in registers (ref: fe_sigproc.c:482-550)
#pragma unroll
    for (int e = 0; e < R; e += 2) {
        const double a = v[e], b = v[e + 1];
        v[e] = __dadd_rn(a, b);
        v[e + 1] = __dsub_rn(a, b);
    }
#pragma unroll
    for (int k = 1; k < LR; ++k) {
        const int h = 1 << k, q = h >> 1, tw = M - k - 1;
#pragma unroll
        for (int i = 0; i < R; i += 2 * h) {
            const double a = v[i], b = v[i + h];
            v[i] = __dadd_rn(a, b);
            v[i + h] = __dsub_rn(a, b);
            v[i + h + q] = -v[i + h + q];
#pragma unroll
            for (int j = 1; j < q; ++j) {
                const int i1 = i + j, i2 = i + h - j, i3 = i + h + j, i4 = i + 2 * h - j;
                const double cc = __ldg(fe.ccc + (j << tw)), ss = __ldg(fe.sss + (j << tw));
                const double x1 = v[i1], x2 = v[i2], x3 = v[i3], x4 = v[i4];
                const double t1 = __dadd_rn(__dmul_rn(x3, cc), __dmul_rn(x4, ss));
                const double t2 = __dsub_rn(__dmul_rn(x3, ss), __dmul_rn(x4, cc));
                v[i4] = __dsub_rn(x2, t2);
                v[i3] = __dsub_rn(-x2, t2);
                v[i2] = __dsub_rn(x1, t1);
                v[i1] = __dadd_rn(x1, t1);
            }
        }
    }
#pragma unroll
    for (int e = 0; e < R; ++e)
        x[fe_pad(lane * R + e)] = v[e];
    __syncwarp();
    // stages LR .. M-1 through shared memory
#pragma unroll
    for (int k = LR; k < M; ++k) {
        const int h = 1 << k, q = h >> 1, tw = M - k - 1;
#pragma unroll
        for (int r = 0; r < R / 4; ++r) {
            const int w = lane + 32 * r;
            const int j = w & (q - 1), i = (w >> (k - 1)) << (k + 1);
            if (j == 0) {
                const int pa = fe_pad(i), pb = fe_pad(i + h), pc = fe_pad(i + h + q);
                const double a = x[pa], b = x[pb];
                x[pa] = __dadd_rn(a, b);
                x[pb] = __dsub_rn(a, b);
                x[pc] = -x[pc];
            } else {
                const int p1 = fe_pad(i + j), p2 = fe_pad(i + h - j), p3 = fe_pad(i + h + j),
                          p4 = fe_pad(i + 2 * h - j);
                const double cc = __ldg(fe.ccc + (j << tw)), ss = __ldg(fe.sss + (j << tw));
                const double x1 = x[p1], x2 = x[p2], x3 = x[p3], x4 = x[p4];
                const double t1 = __dadd_rn(__dmul_rn(x3, cc), __dmul_rn(x4, ss));
                const double t2 = __dsub_rn(__dmul_rn(x3, ss), __dmul_rn(x4, cc));
                x[p4] = __dsub_rn(x2, t2);
                x[p3] = __dsub_rn(-x2, t2);
                x[p2] = __dsub_rn(x1, t1);
                x[p1] = __dadd_rn(x1, t1);
            }
        }
        __syncwarp();
    }
    // power spectrum in place: bin j <= NN/2 overwrites x[j], its imaginary part x[NN-j] lives
    // in the upper half which is only read
    for (int j = lane; j <= NN / 2; j += 32) {
        const double re = x[fe_pad(j)];
        double pw = __dmul_rn(re, re);
        if (j > 0) {
            const double im = x[fe_pad(NN - j)];
            pw = __dadd_rn(pw, __dmul_rn(im, im));
        }
        x[fe_pad(j)] = pw;
    }
    __syncwarp();
    for (int f = lane; f < fe.nfilt; f += 32) {
        const int ss = fe.spec_start[f], cs = fe.filt_start[f], wd = fe.filt_width[f];
        double acc = 0.0;
        for (int i = 0; i < wd; ++i)
            acc = __dadd_rn(acc, __dmul_rn(x[fe_pad(ss + i)], (double)__ldg(fe.coeffs + cs + i)));
        mel[fr * fe.nfilt + f] = acc;
    }
}

// ------------------------------------------------------------------ noise tracker
__device__ __forceinline__ double lower_envelope(double buf, double fl)
{
    // ref: fe_noise.c:110-126 (LAMBDA_A 0.995, LAMBDA_B 0.5)
    if (buf >= fl)
        return __dadd_rn(__dmul_rn(0.995, fl), __dmul_rn(1 - 0.995, buf));
    return __dadd_rn(__dmul_rn(0.5, fl), __dmul_rn(1 - 0.5, buf));
}

__global__ void __launch_bounds__(32)
fe_noise_kernel(FeDev fe, const int64_t *__restrict__ frame_off, double *__restrict__ mel)
{
    __shared__ double gain[2][kMaxFilt];
    const int lane = threadIdx.x, nf = fe.nfilt;
    const int64_t f0 = frame_off[blockIdx.x], T = frame_off[blockIdx.x + 1] - f0;
    double *base = mel + f0 * nf;
    const double max_gain = 20.0, inv_max_gain = 1.0 / 20;
    double power[2] = {0, 0}, noise[2] = {0, 0}, flr[2] = {0, 0}, peak[2] = {0, 0};
    double cur[2] = {0, 0}, nxt[2] = {0, 0};
#pragma unroll
    for (int b = 0; b < 2; ++b)
        if (lane + 32 * b < nf && T > 0)
            nxt[b] = base[lane + 32 * b];
    for (int64_t t = 0; t < T; ++t) {
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int band = lane + 32 * b;
            if (band >= nf)
                continue;
            const double mf = nxt[b];
            cur[b] = mf;
            if (t + 1 < T)
                nxt[b] = base[(t + 1) * nf + band];
            if (t == 0) {  // ref: fe_noise.c:281-290
                power[b] = mf;
                noise[b] = __ddiv_rn(mf, max_gain);
                flr[b] = __ddiv_rn(mf, max_gain);
                peak[b] = 0.0;
            }
            power[b] = __dadd_rn(__dmul_rn(0.7, power[b]), __dmul_rn(1 - 0.7, mf));
            noise[b] = lower_envelope(power[b], noise[b]);
            double sig = __dsub_rn(power[b], noise[b]);
            if (sig < 1.0)
                sig = 1.0;
            flr[b] = lower_envelope(sig, flr[b]);
            {  // temporal masking, ref: fe_noise.c:129-154 (LAMBDA_T 0.85, MU_T 0.2)
                const double in = sig;
                peak[b] = __dmul_rn(peak[b], 0.85);
                if (sig < __dmul_rn(0.85, peak[b]))
                    sig = __dmul_rn(peak[b], 0.2);
                if (in > peak[b])
                    peak[b] = in;
            }
            if (sig < flr[b])
                sig = flr[b];
            double g = sig < __dmul_rn(max_gain, power[b]) ? __ddiv_rn(sig, power[b]) : max_gain;
            if (g < inv_max_gain)
                g = inv_max_gain;
            gain[t & 1][band] = g;
        }
        __syncwarp();
#pragma unroll
        for (int b = 0; b < 2; ++b) {  // ref: fe_noise.c:156-186, SMOOTH_WINDOW 4
            const int band = lane + 32 * b;
            if (band >= nf)
                continue;
            const int l1 = band - 4 > 0 ? band - 4 : 0, l2 = band + 4 < nf - 1 ? band + 4 : nf - 1;
            double coef = 0.0;
            for (int j = l1; j <= l2; ++j)
                coef = __dadd_rn(coef, gain[t & 1][j]);
            base[t * nf + band] = __dmul_rn(cur[b], __ddiv_rn(coef, (double)(l2 - l1 + 1)));
        }
    }
}

// ------------------------------------------------------------------ cepstrum
__global__ void __launch_bounds__(128)
fe_cepstrum_kernel(FeDev fe, int64_t n_frames, const double *__restrict__ mel,
                   float *__restrict__ mfcc)
{
    extern __shared__ double fe_sm[];
    const int nf = fe.nfilt, nc = fe.ncep, ld = nf | 1;
    const int64_t f0 = (int64_t)blockIdx.x * kCepFrames;
    const int F = (int)(n_frames - f0 < kCepFrames ? n_frames - f0 : kCepFrames);
    for (int idx = threadIdx.x; idx < F * nf; idx += blockDim.x) {
        const int fr = idx / nf, j = idx - fr * nf;
        // ref: fe_sigproc.c:596-609, LOG_FLOOR 1e-4
        fe_sm[fr * ld + j] = log(__dadd_rn(mel[f0 * nf + idx], 1e-4));
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < F * nc; idx += blockDim.x) {
        const int fr = idx / nc, i = idx - fr * nc;
        const double *lg = fe_sm + fr * ld;
        const float *cs = fe.mel_cosine + i * nf;
        float c;
        if (fe.transform == SSB_FE_LEGACY) {  // ref: fe_sigproc.c:642-673
            if (i == 0) {
                c = __double2float_rn(__ddiv_rn(lg[0], 2.0));
                for (int j = 1; j < nf; ++j)
                    c = __double2float_rn(__dadd_rn((double)c, lg[j]));
                c = __double2float_rn(__ddiv_rn((double)c, (double)nf));
            } else {
                c = 0.0f;
                for (int j = 0; j < nf; ++j) {
                    const double pr = __dmul_rn(__dmul_rn(lg[j], (double)__ldg(cs + j)),
                                                j == 0 ? 1.0 : 2.0);
                    c = __double2float_rn(__dadd_rn((double)c, pr));
                }
                c = __double2float_rn(__ddiv_rn((double)c, __dmul_rn((double)nf, 2.0)));
            }
        } else {  // ref: fe_sigproc.c:675-700
            if (i == 0) {
                c = __double2float_rn(lg[0]);
                for (int j = 1; j < nf; ++j)
                    c = __double2float_rn(__dadd_rn((double)c, lg[j]));
                c = __fmul_rn(c, fe.transform == SSB_FE_HTK ? fe.sqrt_inv_2n : fe.sqrt_inv_n);
            } else {
                c = 0.0f;
                for (int j = 0; j < nf; ++j)
                    c = __double2float_rn(
                        __dadd_rn((double)c, __dmul_rn(lg[j], (double)__ldg(cs + j))));
                c = __fmul_rn(c, fe.sqrt_inv_2n);
            }
        }
        if (fe.has_lifter)  // ref: fe_sigproc.c:702-715
            c = __fmul_rn(c, __ldg(fe.lifter + i));
        mfcc[f0 * nc + idx] = c;
    }
}

// ------------------------------------------------------------------ CMN sums
__global__ void __launch_bounds__(32)
fe_cmn_kernel(FeDev fe, const int64_t *__restrict__ frame_off, const float *__restrict__ mfcc,
              float *__restrict__ mean, float *__restrict__ scale)
{
    const int i = threadIdx.x, nc = fe.ncep;
    const int64_t f0 = frame_off[blockIdx.x], T = frame_off[blockIdx.x + 1] - f0;
    if (i >= nc || T <= 0)
        return;
    const float *m = mfcc + f0 * nc;
    // ref: cmn.c:173-190 -- frames with c0 < 0 are left out of the mean
    float sum = 0.0f;
    int cnt = 0;
#pragma unroll 8
    for (int64_t t = 0; t < T; ++t) {
        const float c0 = m[t * nc], v = m[t * nc + i];
        if (c0 < 0.0f)
            continue;
        sum = __fadd_rn(sum, v);
        ++cnt;
    }
    const float mu = __fdiv_rn(sum, (float)cnt);
    mean[(int64_t)blockIdx.x * nc + i] = mu;
    float sc = 1.0f;
    if (fe.varnorm) {  // ref: cmn.c:200-217
        float var = 0.0f;
#pragma unroll 8
        for (int64_t t = 0; t < T; ++t) {
            const float d = __fsub_rn(m[t * nc + i], mu);
            var = __fadd_rn(var, __fmul_rn(d, d));
        }
        sc = __double2float_rn(sqrt(__ddiv_rn((double)(int)T, (double)var)));
    }
    scale[(int64_t)blockIdx.x * nc + i] = sc;
}

// ------------------------------------------------------------------ dynamic features
__global__ void __launch_bounds__(256)
fe_feat_kernel(FeDev fe, const int64_t *__restrict__ frame_off, int n_utts, int64_t n_frames,
               const float *__restrict__ mfcc, const float *__restrict__ mean,
               const float *__restrict__ scale, float *__restrict__ feat)
{
    const int nc = fe.ncep;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_frames * nc)
        return;
    const int64_t fr = idx / nc;
    const int i = (int)(idx - fr * nc);
    const int u = find_utt(frame_off, n_utts, fr);
    const int64_t f0 = frame_off[u], T = frame_off[u + 1] - f0, t = fr - f0;
    const float mu = fe.cmn ? mean[(int64_t)u * nc + i] : 0.0f;
    const float sc = fe.varnorm ? scale[(int64_t)u * nc + i] : 1.0f;
    auto c = [&](int64_t tt) {
        tt = tt < 0 ? 0 : (tt >= T ? T - 1 : tt);  // ref: feat.c:993-1001 edge replication
        float v = mfcc[(f0 + tt) * nc + i];
        if (fe.cmn) {
            v = __fsub_rn(v, mu);
            if (fe.varnorm)
                v = __fmul_rn(v, sc);
        }
        return v;
    };
    // ref: feat.c:589-632
    float *o = feat + fr * 3 * nc;
    o[i] = c(t);
    o[nc + i] = __fsub_rn(c(t + 2), c(t - 2));
    const float d1 = __fsub_rn(c(t + 3), c(t - 1)), d2 = __fsub_rn(c(t + 1), c(t - 3));
    o[2 * nc + i] = __fsub_rn(d1, d2);
}

// ------------------------------------------------------------------ host tables
struct FeHost {
    ssb_fe_config_t c;
    int frame_size = 0, frame_shift = 0, fft_size = 0, fft_order = 0;
    float samprate = 0;
    std::vector<double> hamming, ccc, sss;
    std::vector<int> spec_start, filt_start, filt_width;
    std::vector<float> coeffs, mel_cosine, lifter;
    float sqrt_inv_n = 0, sqrt_inv_2n = 0;
};

// ref: fe_sigproc.c:70-84 with the neutral (default) warp
float hz_to_mel(float hz) { return (float)(2595.0 * log10(1.0 + hz / 700.0)); }
float mel_to_hz(float mel) { return (float)(700.0 * (pow(10.0, mel / 2595.0) - 1.0)); }

struct MelLayout {
    float melmin, melbw, fftfreq;
    bool doublewide, round;
    void edges(int i, float fr[3]) const
    {
        for (int j = 0; j < 3; ++j) {
            fr[j] = mel_to_hz((i + (doublewide ? j * 2 : j)) * melbw + melmin);
            if (round)
                fr[j] = ((int)(fr[j] / fftfreq + 0.5)) * fftfreq;
        }
    }
};

// ref: fe_sigproc.c:86-199.  All of this is float32 arithmetic in the reference.
bool build_mel_filters(FeHost &h)
{
    const int nf = h.c.nfilt, half = h.fft_size / 2;
    float melmin = hz_to_mel(h.c.lowerf), melmax = hz_to_mel(h.c.upperf);
    const float melbw = (melmax - melmin) / (nf + 1);
    if (h.c.doublebw) {
        melmin -= melbw;
        melmax += melbw;
        if (mel_to_hz(melmin) < 0 || mel_to_hz(melmax) > h.samprate / 2) {
            set_error("doublebw: filter edges %g..%g Hz out of range", mel_to_hz(melmin),
                      mel_to_hz(melmax));
            return false;
        }
    }
    const MelLayout lay{melmin, melbw, h.samprate / (float)h.fft_size, h.c.doublebw != 0,
                        h.c.round_filters != 0};
    h.spec_start.assign(nf, -1);
    h.filt_start.assign(nf, 0);
    h.filt_width.assign(nf, 0);
    int n = 0;
    for (int i = 0; i < nf; ++i) {
        float fr[3];
        lay.edges(i, fr);
        for (int j = 0; j < half + 1; ++j) {
            const float hz = j * lay.fftfreq;
            if (hz < fr[0])
                continue;
            if (hz > fr[2] || j == half) {
                h.filt_width[i] = j - h.spec_start[i];
                h.filt_start[i] = n;
                n += h.filt_width[i];
                break;
            }
            if (h.spec_start[i] == -1)
                h.spec_start[i] = j;
        }
        if (h.spec_start[i] < 0 || h.filt_width[i] < 0) {
            set_error("mel filter %d does not cover any DFT point", i);
            return false;
        }
    }
    h.coeffs.assign(n, 0.0f);
    n = 0;
    for (int i = 0; i < nf; ++i) {
        float fr[3];
        lay.edges(i, fr);
        for (int j = 0; j < h.filt_width[i]; ++j) {
            const float hz = (h.spec_start[i] + j) * lay.fftfreq;
            if (hz < fr[0] || hz > fr[2]) {
                set_error("failed to create filterbank: %g Hz outside %g..%g", hz, fr[0], fr[2]);
                return false;
            }
            float lo = (hz - fr[0]) / (fr[1] - fr[0]);
            float hi = (fr[2] - hz) / (fr[2] - fr[1]);
            if (h.c.unit_area) {
                lo *= 2 / (fr[2] - fr[0]);
                hi *= 2 / (fr[2] - fr[0]);
            }
            h.coeffs[n++] = lo < hi ? lo : hi;
        }
    }
    return true;
}

bool build_host(FeHost &h)
{
    const ssb_fe_config_t &c = h.c;
    h.samprate = (float)c.samprate;
    if (c.samprate < 1 || c.frate < 1 || c.frate > 32767 || c.frate > c.samprate) {
        set_error("frame rate %d can not be bigger than sample rate %d", c.frate, c.samprate);
        return false;
    }
    if (c.ncep < 1 || c.ncep > kMaxCep || c.nfilt < 1 || c.nfilt > kMaxFilt) {
        set_error("ncep %d / nfilt %d outside 1..%d / 1..%d", c.ncep, c.nfilt, kMaxCep, kMaxFilt);
        return false;
    }
    if (c.transform < SSB_FE_DCT || c.transform > SSB_FE_HTK || c.cmn < 0 || c.cmn > 1) {
        set_error("invalid transform / cmn type");
        return false;
    }
    const int window_samples = (int)(c.wlen * h.samprate);
    if (c.nfft == 0) {
        h.fft_order = 0;
        h.fft_size = 1;
        while (h.fft_size < window_samples) {
            ++h.fft_order;
            h.fft_size <<= 1;
        }
    } else {
        h.fft_size = c.nfft;
        h.fft_order = 0;
        for (int j = c.nfft; j > 1; j >>= 1, ++h.fft_order)
            if (j % 2 != 0) {
                set_error("fft: number of points must be a power of 2 (is %d)", c.nfft);
                return false;
            }
        if (h.fft_size < window_samples) {
            set_error("FFT: number of points must be greater or equal to frame size");
            return false;
        }
    }
    h.frame_shift = (int)(h.samprate / (short)c.frate + 0.5);
    h.frame_size = (int)(c.wlen * h.samprate + 0.5);
    if (h.frame_shift <= 1 || h.frame_size < h.frame_shift) {
        set_error("frame size %d (wlen) must be greater than frame shift %d (frate)",
                  h.frame_size, h.frame_shift);
        return false;
    }
    if (h.frame_size > h.fft_size || h.fft_size < 8 || h.fft_size > kMaxFft) {
        set_error("FFT size %d must be a power of two in [max(8, frame size %d), %d]", h.fft_size,
                  h.frame_size, kMaxFft);
        return false;
    }
    if (c.upperf > h.samprate / 2 + 1.0) {
        set_error("upper frequency %.1f is higher than samprate/2 (%.1f)", c.upperf,
                  h.samprate / 2);
        return false;
    }
    h.hamming.resize(h.frame_size / 2);
    for (int i = 0; i < h.frame_size / 2; ++i)
        h.hamming[i] = 0.54 - 0.46 * cos(2 * M_PI * i / ((double)h.frame_size - 1.0));
    h.ccc.resize(h.fft_size / 4);
    h.sss.resize(h.fft_size / 4);
    for (int i = 0; i < h.fft_size / 4; ++i) {
        const double a = 2 * M_PI * i / h.fft_size;
        h.ccc[i] = cos(a);
        h.sss[i] = sin(a);
    }
    if (!build_mel_filters(h))
        return false;
    h.mel_cosine.resize((size_t)c.ncep * c.nfilt);
    const double step = M_PI / c.nfilt;
    for (int i = 0; i < c.ncep; ++i)
        for (int j = 0; j < c.nfilt; ++j)
            h.mel_cosine[(size_t)i * c.nfilt + j] = (float)cos(step * i * (j + 0.5));
    h.sqrt_inv_n = (float)sqrt(1.0 / c.nfilt);
    h.sqrt_inv_2n = (float)sqrt(2.0 / c.nfilt);
    h.lifter.assign(c.ncep, 1.0f);
    if (c.lifter)
        for (int i = 0; i < c.ncep; ++i)
            h.lifter[i] = (float)(1 + c.lifter / 2 * sin(i * M_PI / c.lifter));
    return true;
}

int64_t frames_for(const FeHost &h, int64_t n_samples)
{
    if (n_samples <= 0)
        return 0;
    if (n_samples < h.frame_size)
        return 1;
    return 1 + (n_samples - h.frame_size) / h.frame_shift + 1;
}

// ---- feat_params.json: a flat object of strings, numbers and booleans
struct JsonKV {
    std::string key, val;
    bool is_string;
};

bool parse_flat_json(const std::string &s, std::vector<JsonKV> &out)
{
    size_t i = 0;
    auto ws = [&] {
        while (i < s.size() && isspace((unsigned char)s[i]))
            ++i;
    };
    auto str = [&](std::string &d) {
        if (i >= s.size() || s[i] != '"')
            return false;
        for (++i; i < s.size() && s[i] != '"'; ++i) {
            if (s[i] == '\\' && i + 1 < s.size())
                ++i;
            d.push_back(s[i]);
        }
        if (i >= s.size())
            return false;
        ++i;
        return true;
    };
    ws();
    if (i >= s.size() || s[i] != '{')
        return false;
    ++i;
    for (;;) {
        ws();
        if (i < s.size() && s[i] == '}')
            return true;
        JsonKV kv;
        if (!str(kv.key))
            return false;
        ws();
        if (i >= s.size() || s[i] != ':')
            return false;
        ++i;
        ws();
        kv.is_string = i < s.size() && s[i] == '"';
        if (kv.is_string) {
            if (!str(kv.val))
                return false;
        } else {
            while (i < s.size() && s[i] != ',' && s[i] != '}' && !isspace((unsigned char)s[i]))
                kv.val.push_back(s[i++]);
            if (kv.val.empty())
                return false;
        }
        out.push_back(kv);
        ws();
        if (i < s.size() && s[i] == ',')
            ++i;
    }
}

bool truthy(const std::string &v)
{
    // ref: src/configuration.c boolean spellings
    return !v.empty() && (v[0] == 'y' || v[0] == 't' || v[0] == 'Y' || v[0] == 'T' || v[0] == '1');
}

}  // namespace

// ------------------------------------------------------------------ C ABI
struct ssb_frontend_s {
    FeHost h;
    FeDev d;
    int device = -1;
    cudaStream_t st = nullptr;
    DBuf tables, pcm, samp_off, frame_off, mel, mfcc, feat, mean, scale;
    std::vector<int64_t> h_frame_off;
    int32_t n_utts = 0;
    int64_t n_frames = 0;
    cudaEvent_t ev[7] = {};
    // the audio travels in pieces on a stream of its own; the mel spectrum of a piece starts when
    // the piece has landed, under the copy of the next one
    static constexpr int kPieces = 4;
    cudaStream_t copy_st = nullptr;
    cudaEvent_t ev_piece[kPieces] = {}, ev_free = nullptr;
    bool have_ev = false, ran = false;
    bool force_generic = false;  // SSB_FE=generic: the shared-memory-only mel spectrum kernel
};

extern "C" void ssb_fe_config_defaults(ssb_fe_config_t *c)
{
    std::memset(c, 0, sizeof(*c));
    c->samprate = 16000;
    c->frate = 100;
    c->ncep = 13;
    c->nfft = 0;
    c->nfilt = 40;
    c->lifter = 0;
    c->unit_area = 1;
    c->round_filters = 1;
    c->transform = SSB_FE_LEGACY;
    c->cmn = SSB_FE_CMN_BATCH;
    c->wlen = 0.025625f;
    c->alpha = 0.97f;
    c->lowerf = 133.33334f;
    c->upperf = 6855.4976f;
}

extern "C" int ssb_fe_config_from_model(const char *hmmdir, ssb_fe_config_t *c)
{
    if (!hmmdir || !c) {
        set_error("ssb_fe_config_from_model: bad arguments");
        return -1;
    }
    ssb_fe_config_defaults(c);
    const std::string path = std::string(hmmdir) + "/feat_params.json";
    FILE *fh = fopen(path.c_str(), "rb");
    if (!fh)
        return 0;  // the reference silently keeps its defaults (src/decoder.c:137-138)
    std::string txt;
    char buf[4096];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), fh)) > 0)
        txt.append(buf, n);
    fclose(fh);
    std::vector<JsonKV> kv;
    if (!parse_flat_json(txt, kv)) {
        set_error("%s: not a flat JSON object", path.c_str());
        return -1;
    }
    int ceplen = -1;
    for (const JsonKV &e : kv) {
        const std::string &k = e.key, &v = e.val;
        const double num = atof(v.c_str());
        if (k == "samprate")
            c->samprate = (int)num;
        else if (k == "frate")
            c->frate = (int)num;
        else if (k == "ncep")
            c->ncep = (int)num;
        else if (k == "ceplen")
            ceplen = (int)num;
        else if (k == "nfft")
            c->nfft = (int)num;
        else if (k == "nfilt")
            c->nfilt = (int)num;
        else if (k == "lifter")
            c->lifter = (int)num;
        else if (k == "wlen")
            c->wlen = (float)num;
        else if (k == "alpha")
            c->alpha = (float)num;
        else if (k == "lowerf")
            c->lowerf = (float)num;
        else if (k == "upperf")
            c->upperf = (float)num;
        else if (k == "remove_dc")
            c->remove_dc = truthy(v);
        else if (k == "remove_noise")
            c->remove_noise = truthy(v);
        else if (k == "unit_area")
            c->unit_area = truthy(v);
        else if (k == "round_filters")
            c->round_filters = truthy(v);
        else if (k == "doublebw")
            c->doublebw = truthy(v);
        else if (k == "varnorm")
            c->varnorm = truthy(v);
        else if (k == "transform") {
            if (v == "dct")
                c->transform = SSB_FE_DCT;
            else if (v == "legacy")
                c->transform = SSB_FE_LEGACY;
            else if (v == "htk")
                c->transform = SSB_FE_HTK;
            else {
                set_error("invalid transform type '%s' (values are 'dct', 'legacy', 'htk')",
                          v.c_str());
                return -1;
            }
        } else if (k == "cmn") {
            if (v == "none")
                c->cmn = SSB_FE_CMN_NONE;
            else if (v == "batch" || v == "current")
                c->cmn = SSB_FE_CMN_BATCH;
            else {
                set_error("cmn '%s' is a streaming mode; whole-utterance batches use batch CMN",
                          v.c_str());
                return -1;
            }
        } else if (k == "feat") {
            if (v != "1s_c_d_dd") {
                set_error("feature type '%s' not implemented (only 1s_c_d_dd)", v.c_str());
                return -1;
            }
        } else if (k == "svspec") {
            // only the split of c/d/dd into three equal contiguous streams is the identity
            int a[6];
            if (sscanf(v.c_str(), "%d-%d/%d-%d/%d-%d", &a[0], &a[1], &a[2], &a[3], &a[4], &a[5]) != 6
                || a[0] != 0 || a[2] != a[1] + 1 || a[4] != a[3] + 1 || a[1] - a[0] != a[3] - a[2]
                || a[3] - a[2] != a[5] - a[4]) {
                set_error("svspec '%s' not implemented (only three equal contiguous streams)",
                          v.c_str());
                return -1;
            }
        } else if ((k == "dither" || k == "logspec" || k == "smoothspec") && truthy(v)) {
            set_error("%s is not implemented by this frontend", k.c_str());
            return -1;
        } else if ((k == "agc" && v != "none") || k == "warp_params" || k == "lda"
                   || (k == "warp_type" && v != "inverse_linear")) {
            set_error("%s is not implemented by this frontend", k.c_str());
            return -1;
        }
    }
    if (ceplen >= 0 && ceplen != c->ncep) {
        set_error("ceplen %d does not match ncep %d", ceplen, c->ncep);
        return -1;
    }
    return 0;
}

extern "C" ssb_frontend_t *ssb_frontend_create(const ssb_fe_config_t *c, int device, void *stream)
{
    if (!c) {
        set_error("ssb_frontend_create: config is NULL");
        return nullptr;
    }
    auto *fe = new ssb_frontend_s();
    fe->h.c = *c;
    if (!build_host(fe->h)) {
        delete fe;
        return nullptr;
    }
    fe->device = device;
    fe->st = (cudaStream_t)stream;
    {
        const char *env = getenv("SSB_FE");
        fe->force_generic = env && std::strcmp(env, "generic") == 0;
    }
    if (device < 0)  // tables only (host tests); every compute call fails
        return fe;
    const FeHost &h = fe->h;
    auto fail = [&](const char *what, cudaError_t e) {
        set_error("ssb_frontend_create: %s: %s", what, cudaGetErrorString(e));
        ssb_frontend_free(fe);
        return (ssb_frontend_t *)nullptr;
    };
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess)
        return fail("cudaSetDevice", e);
    // one allocation, 16-byte aligned sub-tables
    size_t off = 0;
    auto place = [&](size_t bytes) {
        const size_t at = off;
        off += (bytes + 15) & ~(size_t)15;
        return at;
    };
    const size_t o_ham = place(h.hamming.size() * 8), o_ccc = place(h.ccc.size() * 8),
                 o_sss = place(h.sss.size() * 8), o_ss = place(h.spec_start.size() * 4),
                 o_fs = place(h.filt_start.size() * 4), o_fw = place(h.filt_width.size() * 4),
                 o_co = place(std::max<size_t>(h.coeffs.size(), 1) * 4),
                 o_mc = place(h.mel_cosine.size() * 4), o_li = place(h.lifter.size() * 4);
    if (fe->tables.ensure(off) != 0) {
        ssb_frontend_free(fe);
        return nullptr;
    }
    std::vector<char> img(off, 0);
    std::memcpy(&img[o_ham], h.hamming.data(), h.hamming.size() * 8);
    std::memcpy(&img[o_ccc], h.ccc.data(), h.ccc.size() * 8);
    std::memcpy(&img[o_sss], h.sss.data(), h.sss.size() * 8);
    std::memcpy(&img[o_ss], h.spec_start.data(), h.spec_start.size() * 4);
    std::memcpy(&img[o_fs], h.filt_start.data(), h.filt_start.size() * 4);
    std::memcpy(&img[o_fw], h.filt_width.data(), h.filt_width.size() * 4);
    std::memcpy(&img[o_co], h.coeffs.data(), h.coeffs.size() * 4);
    std::memcpy(&img[o_mc], h.mel_cosine.data(), h.mel_cosine.size() * 4);
    std::memcpy(&img[o_li], h.lifter.data(), h.lifter.size() * 4);
    e = cudaMemcpy(fe->tables.p, img.data(), off, cudaMemcpyHostToDevice);
    if (e != cudaSuccess)
        return fail("table upload", e);
    char *base = fe->tables.as<char>();
    FeDev &d = fe->d;
    d.frame_size = h.frame_size;
    d.frame_shift = h.frame_shift;
    d.fft_size = h.fft_size;
    d.fft_order = h.fft_order;
    d.nfilt = c->nfilt;
    d.ncep = c->ncep;
    d.remove_dc = c->remove_dc;
    d.transform = c->transform;
    d.has_lifter = c->lifter != 0;
    d.cmn = c->cmn;
    d.varnorm = c->varnorm && c->cmn;
    d.alpha = c->alpha;
    d.sqrt_inv_n = h.sqrt_inv_n;
    d.sqrt_inv_2n = h.sqrt_inv_2n;
    d.hamming = (const double *)(base + o_ham);
    d.ccc = (const double *)(base + o_ccc);
    d.sss = (const double *)(base + o_sss);
    d.spec_start = (const int *)(base + o_ss);
    d.filt_start = (const int *)(base + o_fs);
    d.filt_width = (const int *)(base + o_fw);
    d.coeffs = (const float *)(base + o_co);
    d.mel_cosine = (const float *)(base + o_mc);
    d.lifter = (const float *)(base + o_li);
    for (auto &ev : fe->ev)
        if ((e = cudaEventCreate(&ev)) != cudaSuccess)
            return fail("cudaEventCreate", e);
    for (auto &ev : fe->ev_piece)
        if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess)
            return fail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&fe->ev_free, cudaEventDisableTiming)) != cudaSuccess)
        return fail("cudaEventCreate", e);
    if ((e = cudaStreamCreateWithFlags(&fe->copy_st, cudaStreamNonBlocking)) != cudaSuccess)
        return fail("cudaStreamCreate", e);
    fe->have_ev = true;
    // two fft_size arrays of doubles per warp, 64 KB per CTA (8 warps up to 512 points)
    e = cudaFuncSetAttribute(fe_melspec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kMelSmem);
    if (e != cudaSuccess)
        return fail("cudaFuncSetAttribute", e);
    return fe;
}

extern "C" void ssb_frontend_free(ssb_frontend_t *fe)
{
    if (!fe)
        return;
    if (fe->device >= 0) {
        cudaSetDevice(fe->device);
        if (fe->ran)
            cudaStreamSynchronize(fe->st);
        if (fe->copy_st) {
            cudaStreamSynchronize(fe->copy_st);
            cudaStreamDestroy(fe->copy_st);
        }
        for (auto &ev : fe->ev)
            if (ev)
                cudaEventDestroy(ev);
        for (auto &ev : fe->ev_piece)
            if (ev)
                cudaEventDestroy(ev);
        if (fe->ev_free)
            cudaEventDestroy(fe->ev_free);
        for (DBuf *b : {&fe->tables, &fe->pcm, &fe->samp_off, &fe->frame_off, &fe->mel, &fe->mfcc,
                        &fe->feat, &fe->mean, &fe->scale})
            b->release();
    }
    delete fe;
}

extern "C" int ssb_frontend_dims(const ssb_frontend_t *fe, int32_t *out8)
{
    if (!fe || !out8) {
        set_error("ssb_frontend_dims: bad arguments");
        return -1;
    }
    const int32_t v[8] = {fe->h.frame_size, fe->h.frame_shift, fe->h.fft_size, fe->h.c.nfilt,
                          fe->h.c.ncep, 3 * fe->h.c.ncep, (int32_t)fe->h.coeffs.size(), 0};
    std::memcpy(out8, v, sizeof(v));
    return 0;
}

extern "C" int64_t ssb_frontend_n_frames(const ssb_frontend_t *fe, int64_t n_samples)
{
    if (!fe) {
        set_error("ssb_frontend_n_frames: frontend is NULL");
        return -1;
    }
    return frames_for(fe->h, n_samples);
}

extern "C" int ssb_frontend_tables(const ssb_frontend_t *fe, int32_t *spec_start,
                                   int32_t *filt_width, float *coeffs, float *mel_cosine,
                                   float *lifter, double *hamming)
{
    if (!fe) {
        set_error("ssb_frontend_tables: frontend is NULL");
        return -1;
    }
    const FeHost &h = fe->h;
    if (spec_start)
        std::copy(h.spec_start.begin(), h.spec_start.end(), spec_start);
    if (filt_width)
        std::copy(h.filt_width.begin(), h.filt_width.end(), filt_width);
    if (coeffs)
        std::copy(h.coeffs.begin(), h.coeffs.end(), coeffs);
    if (mel_cosine)
        std::copy(h.mel_cosine.begin(), h.mel_cosine.end(), mel_cosine);
    if (lifter)
        std::copy(h.lifter.begin(), h.lifter.end(), lifter);
    if (hamming)
        std::copy(h.hamming.begin(), h.hamming.end(), hamming);
    return 0;
}

extern "C" int64_t ssb_frontend_run(ssb_frontend_t *fe, const void *pcm, int32_t encoding,
                                    const int64_t *samp_off, int32_t n_utts)
{
    if (!fe || n_utts < 0 || (n_utts > 0 && !samp_off)
        || (encoding != SSB_PCM_INT16 && encoding != SSB_PCM_FLOAT32)) {
        set_error("ssb_frontend_run: bad arguments");
        return -1;
    }
    if (fe->device < 0) {
        set_error("this frontend was created without a device (device=-1): tables only, there "
                  "is no CPU path");
        return -1;
    }
    API_CUDA(cudaSetDevice(fe->device), -1);
    const FeHost &h = fe->h;
    const int U = n_utts, nf = h.c.nfilt, nc = h.c.ncep;
    fe->ran = false;
    fe->h_frame_off.assign(U + 1, 0);
    if (U > 0 && samp_off[0] != 0) {
        set_error("samp_off[0] must be 0");
        return -1;
    }
    for (int u = 0; u < U; ++u) {
        const int64_t ns = samp_off[u + 1] - samp_off[u];
        if (ns < 0) {
            set_error("utterance %d: sample offsets must be non-decreasing", u);
            return -1;
        }
        fe->h_frame_off[u + 1] = fe->h_frame_off[u] + frames_for(h, ns);
    }
    const int64_t G = fe->h_frame_off[U], S = U > 0 ? samp_off[U] : 0;
    if (S > 0 && !pcm) {
        set_error("ssb_frontend_run: pcm is NULL");
        return -1;
    }
    fe->n_utts = U;
    fe->n_frames = G;
    const size_t ssz = encoding == SSB_PCM_INT16 ? 2 : 4;
    if (fe->pcm.ensure(std::max<size_t>((size_t)S * ssz, 16)) != 0
        || fe->samp_off.ensure((size_t)(U + 1) * 8) != 0
        || fe->frame_off.ensure((size_t)(U + 1) * 8) != 0
        || fe->mel.ensure(std::max<size_t>((size_t)G * nf * 8, 16)) != 0
        || fe->mfcc.ensure(std::max<size_t>((size_t)G * nc * 4, 16)) != 0
        || fe->feat.ensure(std::max<size_t>((size_t)G * 3 * nc * 4, 16)) != 0
        || fe->mean.ensure(std::max<size_t>((size_t)U * nc * 4, 16)) != 0
        || fe->scale.ensure(std::max<size_t>((size_t)U * nc * 4, 16)) != 0)
        return -1;
    cudaStream_t st = fe->st;
    launch_count(true);
    API_CUDA(cudaEventRecord(fe->ev[0], st), -1);
    const int64_t zero = 0;
    API_CUDA(cudaMemcpyAsync(fe->samp_off.p, U > 0 ? samp_off : &zero, (size_t)(U + 1) * 8,
                             cudaMemcpyHostToDevice, st), -1);
    API_CUDA(cudaMemcpyAsync(fe->frame_off.p, fe->h_frame_off.data(), (size_t)(U + 1) * 8,
                             cudaMemcpyHostToDevice, st), -1);
    // pieces of whole utterances with about equal sample counts (one piece for small inputs)
    int piece_u[ssb_frontend_s::kPieces + 1];
    int n_pieces = (S * (int64_t)ssz >= (64 << 20) && U >= 2 * ssb_frontend_s::kPieces) ? ssb_frontend_s::kPieces : 1;
    piece_u[0] = 0;
    for (int k = 1; k < n_pieces; ++k) {
        const int64_t want = S * k / n_pieces;
        int u = piece_u[k - 1];
        while (u < U && samp_off[u] < want)
            ++u;
        piece_u[k] = std::max(u, piece_u[k - 1]);
    }
    piece_u[n_pieces] = U;
    if (n_pieces > 1) {
        // the copy stream may not overwrite the buffer while earlier work on `st` still reads it
        API_CUDA(cudaEventRecord(fe->ev_free, st), -1);
        API_CUDA(cudaStreamWaitEvent(fe->copy_st, fe->ev_free, 0), -1);
    }
    API_CUDA(cudaEventRecord(fe->ev[1], st), -1);
    const int64_t *d_so = fe->samp_off.as<int64_t>(), *d_fo = fe->frame_off.as<int64_t>();
    const bool reg = !h.c.remove_dc && !fe->force_generic
                     && (h.fft_size == 256 || h.fft_size == 512 || h.fft_size == 1024);
    for (int k = 0; k < n_pieces; ++k) {
        const int u0 = piece_u[k], u1 = piece_u[k + 1];
        if (u1 <= u0)
            continue;
        const int64_t s0 = samp_off[u0], s1 = samp_off[u1];
        cudaStream_t cs = n_pieces > 1 ? fe->copy_st : st;
        if (s1 > s0)
            API_CUDA(cudaMemcpyAsync(fe->pcm.as<char>() + (size_t)s0 * ssz, (const char *)pcm + (size_t)s0 * ssz,
                                     (size_t)(s1 - s0) * ssz, cudaMemcpyDefault, cs), -1);
        if (n_pieces > 1) {
            API_CUDA(cudaEventRecord(fe->ev_piece[k], cs), -1);
            API_CUDA(cudaStreamWaitEvent(st, fe->ev_piece[k], 0), -1);
        }
        const int64_t f0 = fe->h_frame_off[u0], f1 = fe->h_frame_off[u1];
        if (f1 <= f0)
            continue;
        if (reg) {  // register-resident first stages
            const int warps = h.fft_size == 1024 ? 4 : 8;
            const int64_t blocks = (f1 - f0 + warps - 1) / warps;
            const size_t smem = (size_t)warps * (h.fft_size + h.fft_size / 16) * 8;
            double *mel = fe->mel.as<double>();
            if (h.fft_size == 256)
                fe_melspec_reg_kernel<3><<<(unsigned)blocks, warps * 32, smem, st>>>(
                    fe->d, fe->pcm.p, encoding, d_so, d_fo, U, f0, f1, mel);
            else if (h.fft_size == 512)
                fe_melspec_reg_kernel<4><<<(unsigned)blocks, warps * 32, smem, st>>>(
                    fe->d, fe->pcm.p, encoding, d_so, d_fo, U, f0, f1, mel);
            else
                fe_melspec_reg_kernel<5><<<(unsigned)blocks, warps * 32, smem, st>>>(
                    fe->d, fe->pcm.p, encoding, d_so, d_fo, U, f0, f1, mel);
        } else {  // any power of two up to 2048, per-frame DC removal
            const int warps = std::max(1, std::min(8, kMelSmem / (2 * h.fft_size * 8)));
            const int64_t blocks = (f1 - f0 + warps - 1) / warps;
            fe_melspec_kernel<<<(unsigned)blocks, warps * 32, (size_t)warps * 2 * h.fft_size * 8,
                                st>>>(fe->d, fe->pcm.p, encoding, d_so, d_fo, U, f0, f1,
                                      fe->mel.as<double>());
        }
        note_launch();
    }
    API_CUDA(cudaEventRecord(fe->ev[2], st), -1);
    if (G > 0 && h.c.remove_noise) {
        fe_noise_kernel<<<U, 32, 0, st>>>(fe->d, d_fo, fe->mel.as<double>());
        note_launch();
    }
    API_CUDA(cudaEventRecord(fe->ev[3], st), -1);
    if (G > 0) {
        const int64_t blocks = (G + kCepFrames - 1) / kCepFrames;
        fe_cepstrum_kernel<<<(unsigned)blocks, 128, (size_t)kCepFrames * (nf | 1) * 8, st>>>(
            fe->d, G, fe->mel.as<double>(), fe->mfcc.as<float>());
        note_launch();
    }
    API_CUDA(cudaEventRecord(fe->ev[4], st), -1);
    if (G > 0 && h.c.cmn) {
        fe_cmn_kernel<<<U, 32, 0, st>>>(fe->d, d_fo, fe->mfcc.as<float>(), fe->mean.as<float>(),
                                        fe->scale.as<float>());
        note_launch();
    }
    API_CUDA(cudaEventRecord(fe->ev[5], st), -1);
    if (G > 0) {
        const int64_t items = G * nc, blocks = (items + 255) / 256;
        fe_feat_kernel<<<(unsigned)blocks, 256, 0, st>>>(fe->d, d_fo, U, G, fe->mfcc.as<float>(),
                                                         fe->mean.as<float>(),
                                                         fe->scale.as<float>(),
                                                         fe->feat.as<float>());
        note_launch();
    }
    API_CUDA(cudaEventRecord(fe->ev[6], st), -1);
    API_CUDA(cudaGetLastError(), -1);
    fe->ran = true;
    return G;
}

extern "C" int ssb_frontend_download(ssb_frontend_t *fe, int64_t *frame_off, float *mfcc,
                                     float *feat)
{
    if (!fe || !fe->ran) {
        set_error("ssb_frontend_download: nothing has been run");
        return -1;
    }
    API_CUDA(cudaSetDevice(fe->device), -1);
    const int nc = fe->h.c.ncep;
    if (frame_off)
        std::copy(fe->h_frame_off.begin(), fe->h_frame_off.end(), frame_off);
    if (mfcc && fe->n_frames > 0)
        API_CUDA(cudaMemcpyAsync(mfcc, fe->mfcc.p, (size_t)fe->n_frames * nc * 4,
                                 cudaMemcpyDeviceToHost, fe->st), -1);
    if (feat && fe->n_frames > 0)
        API_CUDA(cudaMemcpyAsync(feat, fe->feat.p, (size_t)fe->n_frames * 3 * nc * 4,
                                 cudaMemcpyDeviceToHost, fe->st), -1);
    API_CUDA(cudaStreamSynchronize(fe->st), -1);
    return 0;
}

extern "C" const float *ssb_frontend_feat_device(const ssb_frontend_t *fe)
{
    if (!fe || !fe->ran) {
        set_error("ssb_frontend_feat_device: nothing has been run");
        return nullptr;
    }
    // consumers run on other streams (the one-shot entry points on a stream of their own, ordered
    // after the legacy default stream only): a frontend on a stream of its own is waited for here
    if (fe->st != nullptr) {
        cudaSetDevice(fe->device);
        cudaStreamSynchronize(fe->st);
    }
    return fe->feat.as<float>();
}

extern "C" int ssb_frontend_kernel_ms(ssb_frontend_t *fe, float *ms8)
{
    if (!fe || !fe->ran || !ms8) {
        set_error("ssb_frontend_kernel_ms: nothing has been run");
        return -1;
    }
    API_CUDA(cudaSetDevice(fe->device), -1);
    API_CUDA(cudaEventSynchronize(fe->ev[6]), -1);
    std::memset(ms8, 0, 8 * sizeof(float));
    for (int i = 0; i < 5; ++i)
        API_CUDA(cudaEventElapsedTime(&ms8[i], fe->ev[i + 1], fe->ev[i + 2]), -1);
    API_CUDA(cudaEventElapsedTime(&ms8[5], fe->ev[0], fe->ev[6]), -1);
    return 0;
}
