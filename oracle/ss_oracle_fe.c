/*
 * ss_oracle_fe.c -- CPU oracle of the acoustic frontend (TEST INFRASTRUCTURE ONLY).
 *
 * Whole-utterance restatement of PCM -> MFCC -> CMN -> dynamic features, i.e. what
 * acmod_process_full_raw (ref: src/acmod.c:424-455) obtains from fe_start /
 * fe_process_int16 / fe_end and feat_s2mfc2feat_block_utt.  Array-based: the
 * reference's streaming buffers (spch / overflow_samps) are replaced by their
 * closed form -- frame t covers samples [t*shift, t*shift + frame_size), the last
 * frame is the partial remainder, the pre-emphasis carry is the sample before the
 * frame (ref: src/fe_interface.c:352-360, 379-391, 578-669, 694-713).
 *
 * Arithmetic types follow the reference: frame_t / powspec_t / window_t are float64
 * (ref: fe_type.h:42-44), filter coefficients, DCT basis, lifter and MFCCs float32.
 * Built with -ffp-contract=off like the rest of the oracle.
 */
#include "ss_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

struct orc_fe_s {
    orc_fe_cfg_t c;
    int frame_size, frame_shift, fft_size, fft_order;
    float samprate;
    double *hamming;      /* first half */
    double *ccc, *sss;    /* fft_size / 4 twiddles */
    int *spec_start, *filt_start, *filt_width;
    float *filt_coeffs;
    int n_coeffs;
    float *mel_cosine;    /* [ncep][nfilt] */
    float sqrt_inv_n, sqrt_inv_2n;
    float *lifter;
};

/* ref: fe_sigproc.c:70-84 (no warping: the only warp the bundled models use is the
 * neutral default) */
static float
mel_of(float hz)
{
    return (float)(2595.0 * log10(1.0 + hz / 700.0));
}

static float
mel_inv(float mel)
{
    return (float)(700.0 * (pow(10.0, mel / 2595.0) - 1.0));
}

static void
filter_edges(const orc_fe_t *fe, int i, float melmin, float melbw, float fftfreq, float *freqs)
{
    int j;
    for (j = 0; j < 3; ++j) {
        if (fe->c.doublebw)
            freqs[j] = mel_inv((i + j * 2) * melbw + melmin);
        else
            freqs[j] = mel_inv((i + j) * melbw + melmin);
        if (fe->c.round_filters)
            freqs[j] = ((int)(freqs[j] / fftfreq + 0.5)) * fftfreq;
    }
}

/* ref: fe_sigproc.c:86-199 */
static int
build_filters(orc_fe_t *fe)
{
    const int nf = fe->c.nfilt, half = fe->fft_size / 2;
    float melmin = mel_of(fe->c.lowerf), melmax = mel_of(fe->c.upperf);
    float melbw = (melmax - melmin) / (nf + 1);
    float fftfreq = fe->samprate / (float)fe->fft_size;
    int i, j, n = 0;
    fe->spec_start = calloc(nf, sizeof(int));
    fe->filt_start = calloc(nf, sizeof(int));
    fe->filt_width = calloc(nf, sizeof(int));
    if (fe->c.doublebw) {
        melmin -= melbw;
        melmax += melbw;
        if (mel_inv(melmin) < 0 || mel_inv(melmax) > fe->samprate / 2)
            return -1;
    }
    for (i = 0; i < nf; ++i) {
        float fr[3];
        filter_edges(fe, i, melmin, melbw, fftfreq, fr);
        fe->spec_start[i] = -1;
        for (j = 0; j < half + 1; ++j) {
            float hz = j * fftfreq;
            if (hz < fr[0])
                continue;
            if (hz > fr[2] || j == half) {
                fe->filt_width[i] = j - fe->spec_start[i];
                fe->filt_start[i] = n;
                n += fe->filt_width[i];
                break;
            }
            if (fe->spec_start[i] == -1)
                fe->spec_start[i] = j;
        }
    }
    fe->n_coeffs = n;
    fe->filt_coeffs = calloc(n > 0 ? n : 1, sizeof(float));
    n = 0;
    for (i = 0; i < nf; ++i) {
        float fr[3];
        filter_edges(fe, i, melmin, melbw, fftfreq, fr);
        for (j = 0; j < fe->filt_width[i]; ++j) {
            float hz = (fe->spec_start[i] + j) * fftfreq, lo, hi;
            if (hz < fr[0] || hz > fr[2])
                return -1;
            lo = (hz - fr[0]) / (fr[1] - fr[0]);
            hi = (fr[2] - hz) / (fr[2] - fr[1]);
            if (fe->c.unit_area) {
                lo *= 2 / (fr[2] - fr[0]);
                hi *= 2 / (fr[2] - fr[0]);
            }
            fe->filt_coeffs[n++] = lo < hi ? lo : hi;
        }
    }
    return 0;
}

void
orc_fe_free(orc_fe_t *fe)
{
    if (!fe)
        return;
    free(fe->hamming);
    free(fe->ccc);
    free(fe->sss);
    free(fe->spec_start);
    free(fe->filt_start);
    free(fe->filt_width);
    free(fe->filt_coeffs);
    free(fe->mel_cosine);
    free(fe->lifter);
    free(fe);
}

/* ref: fe_interface.c:83-178 (general parameters), :270-350 (fe_init),
 * fe_sigproc.c:201-236 (DCT basis, lifter), :258-270 (Hamming), :447-458 (twiddles) */
orc_fe_t *
orc_fe_new(const orc_fe_cfg_t *cfg)
{
    orc_fe_t *fe = calloc(1, sizeof(*fe));
    int i, j, window_samples;
    float wlen = cfg->wlen;
    fe->c = *cfg;
    fe->samprate = (float)cfg->samprate;
    if (cfg->frate < 1 || cfg->frate > cfg->samprate || cfg->ncep < 1 || cfg->nfilt < 1)
        goto fail;
    window_samples = (int)(wlen * fe->samprate);
    if (cfg->nfft == 0) {
        fe->fft_order = 0;
        fe->fft_size = 1;
        while (fe->fft_size < window_samples) {
            fe->fft_order++;
            fe->fft_size <<= 1;
        }
    } else {
        fe->fft_size = cfg->nfft;
        for (j = cfg->nfft, fe->fft_order = 0; j > 1; j >>= 1, fe->fft_order++)
            if (j % 2 != 0)
                goto fail;
        if (fe->fft_size < window_samples)
            goto fail;
    }
    fe->frame_shift = (int)(fe->samprate / (short)cfg->frate + 0.5);
    fe->frame_size = (int)(wlen * fe->samprate + 0.5);
    if (fe->frame_shift <= 1 || fe->frame_size < fe->frame_shift || fe->frame_size > fe->fft_size
        || fe->fft_size < 4)
        goto fail;
    if (cfg->upperf > fe->samprate / 2 + 1.0)
        goto fail;
    fe->hamming = calloc(fe->frame_size / 2 + 1, sizeof(double));
    for (i = 0; i < fe->frame_size / 2; ++i)
        fe->hamming[i] = 0.54 - 0.46 * cos(2 * M_PI * i / ((double)fe->frame_size - 1.0));
    fe->ccc = calloc(fe->fft_size / 4, sizeof(double));
    fe->sss = calloc(fe->fft_size / 4, sizeof(double));
    for (i = 0; i < fe->fft_size / 4; ++i) {
        double a = 2 * M_PI * i / fe->fft_size;
        fe->ccc[i] = cos(a);
        fe->sss[i] = sin(a);
    }
    if (build_filters(fe) < 0)
        goto fail;
    fe->mel_cosine = calloc((size_t)cfg->ncep * cfg->nfilt, sizeof(float));
    {
        double step = M_PI / cfg->nfilt;
        for (i = 0; i < cfg->ncep; ++i)
            for (j = 0; j < cfg->nfilt; ++j)
                fe->mel_cosine[i * cfg->nfilt + j] = (float)cos(step * i * (j + 0.5));
    }
    fe->sqrt_inv_n = (float)sqrt(1.0 / cfg->nfilt);
    fe->sqrt_inv_2n = (float)sqrt(2.0 / cfg->nfilt);
    if (cfg->lifter) {
        fe->lifter = calloc(cfg->ncep, sizeof(float));
        for (i = 0; i < cfg->ncep; ++i)
            fe->lifter[i] = (float)(1 + cfg->lifter / 2 * sin(i * M_PI / cfg->lifter));
    }
    return fe;
fail:
    orc_fe_free(fe);
    return NULL;
}

int
orc_fe_dims(const orc_fe_t *fe, int32_t *out)
{
    out[0] = fe->frame_size;
    out[1] = fe->frame_shift;
    out[2] = fe->fft_size;
    out[3] = fe->n_coeffs;
    return 0;
}

int
orc_fe_tables(const orc_fe_t *fe, int32_t *spec_start, int32_t *filt_width, float *coeffs,
              float *mel_cosine, float *lifter, double *hamming)
{
    int i;
    for (i = 0; i < fe->c.nfilt; ++i) {
        spec_start[i] = fe->spec_start[i];
        filt_width[i] = fe->filt_width[i];
    }
    memcpy(coeffs, fe->filt_coeffs, sizeof(float) * fe->n_coeffs);
    memcpy(mel_cosine, fe->mel_cosine, sizeof(float) * fe->c.ncep * fe->c.nfilt);
    for (i = 0; i < fe->c.ncep; ++i)
        lifter[i] = fe->lifter ? fe->lifter[i] : 1.0f;
    memcpy(hamming, fe->hamming, sizeof(double) * (fe->frame_size / 2));
    return 0;
}

/* ref: fe_interface.c:379-391 (every whole-utterance call ends with one partial frame) */
long
orc_fe_n_frames(const orc_fe_t *fe, long n_samples)
{
    if (n_samples <= 0)
        return 0;
    if (n_samples < fe->frame_size)
        return 1;
    return 1 + (n_samples - fe->frame_size) / fe->frame_shift + 1;
}

/* ref: fe_sigproc.c:460-550.  In-place real FFT, output x[j] = Re, x[n-j] = Im. */
static void
fft_real(const orc_fe_t *fe, double *x)
{
    const int n = fe->fft_size, m = fe->fft_order;
    int i, j, k;
    for (i = 0, j = 0; i < n - 1; ++i) {
        if (i < j) {
            double t = x[j];
            x[j] = x[i];
            x[i] = t;
        }
        k = n / 2;
        while (k <= j) {
            j -= k;
            k /= 2;
        }
        j += k;
    }
    for (i = 0; i < n; i += 2) {
        double a = x[i], b = x[i + 1];
        x[i] = a + b;
        x[i + 1] = a - b;
    }
    for (k = 1; k < m; ++k) {
        const int h = 1 << k, q = h >> 1, tw = m - k - 1;
        for (i = 0; i < n; i += 2 * h) {
            double a = x[i], b = x[i + h];
            x[i] = a + b;
            x[i + h] = a - b;
            x[i + h + q] = -x[i + h + q];
            for (j = 1; j < q; ++j) {
                const int i1 = i + j, i2 = i + h - j, i3 = i + h + j, i4 = i + 2 * h - j;
                const double cc = fe->ccc[j << tw], ss = fe->sss[j << tw];
                const double t1 = x[i3] * cc + x[i4] * ss;
                const double t2 = x[i3] * ss - x[i4] * cc;
                x[i4] = x[i2] - t2;
                x[i3] = -x[i2] - t2;
                x[i2] = x[i1] - t1;
                x[i1] = x[i1] + t1;
            }
        }
    }
}

typedef struct {
    double *power, *noise, *floor, *peak, *signal, *gain;
    int undefined;
} noise_t;

/* ref: fe_noise.c:110-126 */
static void
lower_envelope(const double *buf, double *fl, int n)
{
    int i;
    for (i = 0; i < n; ++i) {
        if (buf[i] >= fl[i])
            fl[i] = 0.995 * fl[i] + (1 - 0.995) * buf[i];
        else
            fl[i] = 0.5 * fl[i] + (1 - 0.5) * buf[i];
    }
}

/* ref: fe_noise.c:266-327 (+ :129-186 temporal masking, weight smoothing) */
static void
remove_noise(noise_t *ns, double *mfspec, int n)
{
    const double max_gain = 20, inv_max_gain = 1.0 / 20, lambda_power = 0.7,
                 comp_lambda_power = 1 - 0.7, lambda_t = 0.85, mu_t = 0.2;
    int i, j;
    if (ns->undefined) {
        for (i = 0; i < n; ++i) {
            ns->power[i] = mfspec[i];
            ns->noise[i] = mfspec[i] / max_gain;
            ns->floor[i] = mfspec[i] / max_gain;
            ns->peak[i] = 0.0;
        }
        ns->undefined = 0;
    }
    for (i = 0; i < n; ++i)
        ns->power[i] = lambda_power * ns->power[i] + comp_lambda_power * mfspec[i];
    lower_envelope(ns->power, ns->noise, n);
    for (i = 0; i < n; ++i) {
        ns->signal[i] = ns->power[i] - ns->noise[i];
        if (ns->signal[i] < 1.0)
            ns->signal[i] = 1.0;
    }
    lower_envelope(ns->signal, ns->floor, n);
    for (i = 0; i < n; ++i) {
        double cur = ns->signal[i];
        ns->peak[i] *= lambda_t;
        if (ns->signal[i] < lambda_t * ns->peak[i])
            ns->signal[i] = ns->peak[i] * mu_t;
        if (cur > ns->peak[i])
            ns->peak[i] = cur;
    }
    for (i = 0; i < n; ++i)
        if (ns->signal[i] < ns->floor[i])
            ns->signal[i] = ns->floor[i];
    for (i = 0; i < n; ++i) {
        if (ns->signal[i] < max_gain * ns->power[i])
            ns->gain[i] = ns->signal[i] / ns->power[i];
        else
            ns->gain[i] = max_gain;
        if (ns->gain[i] < inv_max_gain)
            ns->gain[i] = inv_max_gain;
    }
    for (i = 0; i < n; ++i) {
        int l1 = i - 4 > 0 ? i - 4 : 0, l2 = i + 4 < n - 1 ? i + 4 : n - 1;
        double coef = 0;
        for (j = l1; j <= l2; ++j)
            coef += ns->gain[j];
        mfspec[i] = mfspec[i] * (coef / (l2 - l1 + 1));
    }
}

/* ref: fe_sigproc.c:596-640 (log, DCT), :642-715 */
static void
mel_cep(const orc_fe_t *fe, double *mfspec, float *cep)
{
    const int nf = fe->c.nfilt, nc = fe->c.ncep;
    int i, j;
    for (i = 0; i < nf; ++i)
        mfspec[i] = log(mfspec[i] + 1e-4);
    if (fe->c.transform == ORC_FE_LEGACY) {
        cep[0] = mfspec[0] / 2;
        for (j = 1; j < nf; ++j)
            cep[0] += mfspec[j];
        cep[0] /= (double)nf;
        for (i = 1; i < nc; ++i) {
            cep[i] = 0;
            for (j = 0; j < nf; ++j)
                cep[i] += mfspec[j] * fe->mel_cosine[i * nf + j] * (j == 0 ? 1 : 2);
            cep[i] /= (double)nf * 2;
        }
    } else {
        cep[0] = mfspec[0];
        for (j = 1; j < nf; ++j)
            cep[0] += mfspec[j];
        cep[0] = cep[0] * (fe->c.transform == ORC_FE_HTK ? fe->sqrt_inv_2n : fe->sqrt_inv_n);
        for (i = 1; i < nc; ++i) {
            cep[i] = 0;
            for (j = 0; j < nf; ++j)
                cep[i] += mfspec[j] * fe->mel_cosine[i * nf + j];
            cep[i] = cep[i] * fe->sqrt_inv_2n;
        }
    }
    if (fe->lifter)
        for (i = 0; i < nc; ++i)
            cep[i] = cep[i] * fe->lifter[i];
}

/* One utterance, PCM (int16 when pcm16 != NULL, else float32 in [-1,1)) -> MFCC.
 * Optionally also returns the mel spectrum after noise removal ([frames][nfilt] f64). */
long
orc_fe_mfcc(const orc_fe_t *fe, const int16_t *pcm16, const float *pcm32, long n_samples,
            float *mfcc, double *melspec)
{
    const int fs = fe->frame_size, sh = fe->frame_shift, n = fe->fft_size, nf = fe->c.nfilt;
    const long nfr = orc_fe_n_frames(fe, n_samples);
    double *x = calloc(n, sizeof(double)), *spec = calloc(n, sizeof(double));
    double *buf = calloc((size_t)nf * 7, sizeof(double)), *mf = buf + 6 * nf;
    float *spch = calloc(fs, sizeof(float));
    noise_t ns = { buf, buf + nf, buf + 2 * nf, buf + 3 * nf, buf + 4 * nf, buf + 5 * nf, 1 };
    const float alpha = fe->c.alpha;
    long t;
    for (t = 0; t < nfr; ++t) {
        const long s0 = t * sh;
        const int len = n_samples - s0 < fs ? (int)(n_samples - s0) : fs;
        float prior = 0;
        int i, f;
        for (i = 0; i < len; ++i)
            spch[i] = pcm16 ? (float)pcm16[s0 + i] : pcm32[s0 + i] * 32768.0F;
        if (s0 > 0)
            prior = pcm16 ? (float)pcm16[s0 - 1] : pcm32[s0 - 1] * 32768.0F;
        /* ref: fe_sigproc.c:238-247, 292-321 */
        if (alpha != 0.0) {
            x[0] = (double)spch[0] - (double)prior * alpha;
            for (i = 1; i < len; ++i)
                x[i] = (double)spch[i] - (double)spch[i - 1] * alpha;
        } else
            for (i = 0; i < len; ++i)
                x[i] = spch[i];
        memset(x + len, 0, (n - len) * sizeof(double));
        if (fe->c.remove_dc) {
            double mean = 0;
            for (i = 0; i < fs; ++i)
                mean += x[i];
            mean /= fs;
            for (i = 0; i < fs; ++i)
                x[i] -= mean;
        }
        for (i = 0; i < fs / 2; ++i) {
            x[i] = x[i] * fe->hamming[i];
            x[fs - 1 - i] = x[fs - 1 - i] * fe->hamming[i];
        }
        fft_real(fe, x);
        /* ref: fe_sigproc.c:552-577 */
        spec[0] = x[0] * x[0];
        for (i = 1; i <= n / 2; ++i)
            spec[i] = x[i] * x[i] + x[n - i] * x[n - i];
        /* ref: fe_sigproc.c:579-594 */
        for (f = 0; f < nf; ++f) {
            mf[f] = 0;
            for (i = 0; i < fe->filt_width[f]; ++i)
                mf[f] += spec[fe->spec_start[f] + i] * fe->filt_coeffs[fe->filt_start[f] + i];
        }
        if (fe->c.remove_noise)
            remove_noise(&ns, mf, nf);
        if (melspec)
            memcpy(melspec + (size_t)t * nf, mf, nf * sizeof(double));
        mel_cep(fe, mf, mfcc + (size_t)t * fe->c.ncep);
    }
    free(x);
    free(spec);
    free(buf);
    free(spch);
    return nfr;
}

/* ref: cmn.c:159-229 (batch CMN, frames with c0 < 0 left out of the mean),
 * feat.c:978-1007 (edge replication AFTER normalisation), :589-632 (1s_c_d_dd).
 * mfcc is normalised in place, like the reference does to its buffer. */
int
orc_fe_feat(const orc_fe_t *fe, float *mfcc, long nfr, float *feat)
{
    const int nc = fe->c.ncep;
    long t;
    int i;
    if (nfr <= 0)
        return 0;
    if (fe->c.cmn == ORC_FE_CMN_BATCH) {
        float *sum = calloc(nc, sizeof(float)), *mean = calloc(nc, sizeof(float));
        float *var = calloc(nc, sizeof(float));
        int cnt = 0;
        for (t = 0; t < nfr; ++t) {
            if (mfcc[t * nc] < 0)
                continue;
            for (i = 0; i < nc; ++i)
                sum[i] += mfcc[t * nc + i];
            ++cnt;
        }
        for (i = 0; i < nc; ++i)
            mean[i] = sum[i] / cnt;
        if (!fe->c.varnorm) {
            for (t = 0; t < nfr; ++t)
                for (i = 0; i < nc; ++i)
                    mfcc[t * nc + i] -= mean[i];
        } else {
            for (t = 0; t < nfr; ++t)
                for (i = 0; i < nc; ++i) {
                    float d = mfcc[t * nc + i] - mean[i];
                    var[i] += d * d;
                }
            for (i = 0; i < nc; ++i)
                var[i] = (float)sqrt((double)nfr / var[i]);
            for (t = 0; t < nfr; ++t)
                for (i = 0; i < nc; ++i)
                    mfcc[t * nc + i] = (mfcc[t * nc + i] - mean[i]) * var[i];
        }
        free(sum);
        free(mean);
        free(var);
    }
#define C(tt) (mfcc + (size_t)((tt) < 0 ? 0 : (tt) >= nfr ? nfr - 1 : (tt)) * nc)
    for (t = 0; t < nfr; ++t) {
        float *f = feat + (size_t)t * 3 * nc;
        for (i = 0; i < nc; ++i) {
            float d1, d2;
            f[i] = C(t)[i];
            f[nc + i] = C(t + 2)[i] - C(t - 2)[i];
            d1 = C(t + 3)[i] - C(t - 1)[i];
            d2 = C(t + 1)[i] - C(t - 3)[i];
            f[2 * nc + i] = d1 - d2;
        }
    }
#undef C
    return 0;
}
