/*
 * ss_oracle_cont.c -- CPU oracle of the fully continuous scorer (TEST INFRASTRUCTURE ONLY).
 *
 * ref: src/ms_mgau.c:279-368 (ms_cont_mgau_frame_eval), src/ms_gauden.c:342-457 (gauden_dist:
 * stateless float top-N per codebook and stream), src/ms_senone.c:315-362 (senone_eval:
 * (int + 1023) >> 10, table logmath_add, / aw, int16 clamp), src/logmath.c:229-275.
 * One codebook per senone (the ".cont." mapping, ms_senone.c:262-275).
 */
#include "ss_oracle.h"

#include <stdlib.h>
#include <string.h>

typedef struct {
    float dist;
    int32_t id;
} gd_t;

/* ref: ms_gauden.c:342-426.  The early-out of the reference is result-neutral here: the
 * partial sums only decrease and nothing is truncated before the comparison. */
static int
top_densities(const orc_model_t *m, int mgau, int f, const float *x, int n_top, gd_t *out)
{
    const int L = m->featlen[f], nd = m->n_density;
    const float *mean = m->mean + m->gau_off[mgau * m->n_feat + f];
    const float *var = m->var + m->gau_off[mgau * m->n_feat + f];
    const float *det = m->det + (size_t)(mgau * m->n_feat + f) * nd;
    int d, i, j;
    if (n_top >= nd) {
        for (d = 0; d < nd; ++d) {
            float dval = det[d];
            for (i = 0; i < L; ++i) {
                float diff = x[i] - mean[d * L + i];
                dval -= diff * diff * var[d * L + i];
            }
            out[d].dist = dval;
            out[d].id = d;
        }
        return nd;
    }
    for (i = 0; i < n_top; ++i) {
        out[i].dist = (float)ORC_WORST_DIST;
        out[i].id = 0; /* ckd_calloc'ed in the reference */
    }
    for (d = 0; d < nd; ++d) {
        float dval = det[d];
        for (i = 0; i < L; ++i) {
            float diff = x[i] - mean[d * L + i];
            dval -= diff * diff * var[d * L + i];
        }
        if (dval < out[n_top - 1].dist)
            continue;
        for (i = 0; i < n_top && dval < out[i].dist; ++i)
            ;
        for (j = n_top - 1; j > i; --j)
            out[j] = out[j - 1];
        out[i].dist = dval;
        out[i].id = d;
    }
    return n_top;
}

/* ref: logmath.c:229-275 with the 8-bit shift-10 table (zero = INT32_MIN >> 12) */
static int32_t
table_add(const orc_model_t *m, int32_t x, int32_t y)
{
    const int32_t zero = INT32_MIN >> (ORC_SENSCR_SHIFT + 2);
    int32_t d, r;
    if (x <= zero)
        return y;
    if (y <= zero)
        return x;
    if (x > y) {
        d = x - y;
        r = x;
    } else {
        d = y - x;
        r = y;
    }
    if (d < 0 || d >= 256)
        return r;
    return r + m->lut8[d]; /* padded with zeros past the reference's table_size */
}

static int32_t
density_score(float dist)
{
    if (dist < (float)INT32_MIN)
        return INT32_MIN >> ORC_SENSCR_SHIFT;
    return ((int32_t)dist + ((1 << ORC_SENSCR_SHIFT) - 1)) >> ORC_SENSCR_SHIFT;
}

/* ref: ms_senone.c:315-362 (aw = 1) */
static int32_t
senone_score(const orc_model_t *m, int s, int topn, const float *feat)
{
    gd_t top[ORC_MAX_TOPN > 64 ? ORC_MAX_TOPN : 64];
    int32_t scr = 0;
    int f, t, n;
    for (f = 0; f < m->n_feat; ++f) {
        const uint8_t *pdf = m->mixw + ((size_t)s * m->n_feat + f) * m->n_density;
        int32_t fscr;
        n = top_densities(m, s, f, feat + m->featoff[f], topn, top);
        fscr = density_score(top[0].dist) - pdf[top[0].id];
        for (t = 1; t < n; ++t)
            fscr = table_add(m, fscr, density_score(top[t].dist) - pdf[top[t].id]);
        scr -= fscr;
    }
    if (scr > 32767)
        scr = 32767;
    if (scr < -32768)
        scr = -32768;
    return scr;
}

static int16_t
clamp16(int32_t v)
{
    return (int16_t)(v > 32767 ? 32767 : v < -32768 ? -32768 : v);
}

/* ref: ms_mgau.c:279-368 */
int
orc_cont_frame_eval(const orc_model_t *m, int topn, int16_t *senscr, const uint8_t *active,
                    int32_t n_active, const float *feat, int32_t compallsen)
{
    int32_t best = INT32_MAX, i, n, s;
    if (topn == 0 || topn > m->n_density)
        topn = m->n_density;
    if (m->n_density > 64)
        return -1;
    if (compallsen) {
        for (s = 0; s < m->n_sen; ++s) {
            senscr[s] = (int16_t)senone_score(m, s, topn, feat);
            if (best > senscr[s])
                best = senscr[s];
        }
        for (s = 0; s < m->n_sen; ++s)
            senscr[s] = clamp16(senscr[s] - best);
        return 0;
    }
    for (n = i = 0; i < n_active; ++i) {
        s = active[i] + n;
        senscr[s] = (int16_t)senone_score(m, s, topn, feat);
        if (best > senscr[s])
            best = senscr[s];
        n = s;
    }
    for (n = i = 0; i < n_active; ++i) {
        s = active[i] + n;
        senscr[s] = clamp16(senscr[s] - best);
        n = s;
    }
    return 0;
}
