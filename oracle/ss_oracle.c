/*
 * ss_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see ss_oracle.h).
 *
 * A from-scratch, array-based restatement of the reference algorithms; each
 * function cites the reference file:line it follows.  Compile with
 * -ffp-contract=off (oracle/Makefile) so the fp32 distance chain is unfused.
 */
#include "ss_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* small file helpers                                                  */
/* ------------------------------------------------------------------ */
typedef struct {
    uint8_t *buf;
    size_t len, pos;
    int swap, do_chk;
    uint32_t chk;
} rd_t;

static int
rd_open(rd_t *r, const char *dir, const char *name)
{
    char path[4096];
    FILE *fh;
    long n;
    snprintf(path, sizeof(path), "%s/%s", dir, name);
    memset(r, 0, sizeof(*r));
    if ((fh = fopen(path, "rb")) == NULL)
        return -1;
    fseek(fh, 0, SEEK_END);
    n = ftell(fh);
    fseek(fh, 0, SEEK_SET);
    r->buf = malloc(n + 1);
    if (fread(r->buf, 1, n, fh) != (size_t)n) {
        fclose(fh);
        free(r->buf);
        return -1;
    }
    fclose(fh);
    r->len = n;
    return 0;
}

static uint32_t
bswap32(uint32_t v)
{
    return (v >> 24) | ((v >> 8) & 0xff00) | ((v << 8) & 0xff0000) | (v << 24);
}

/* ref: s3file.c:415-440 (get) + :365-397 (checksum: rotl 20 for 4-byte items) */
static int
rd_u32(rd_t *r, void *out, size_t n)
{
    uint32_t *o = out;
    size_t i;
    if (r->pos + 4 * n > r->len)
        return -1;
    memcpy(o, r->buf + r->pos, 4 * n);
    r->pos += 4 * n;
    for (i = 0; i < n; ++i) {
        if (r->swap)
            o[i] = bswap32(o[i]);
        if (r->do_chk)
            r->chk = ((r->chk << 20) | (r->chk >> 12)) + o[i];
    }
    return 0;
}

/* ref: s3file.c:210-326 -- "s3\n", name/value lines, "endhdr", byte-order magic */
static int
rd_s3_header(rd_t *r)
{
    int chk = 0;
    uint32_t magic;
    if (r->len < 3 || memcmp(r->buf, "s3\n", 3) != 0)
        return -1; /* pre-1996 format not needed by any bundled model */
    r->pos = 3;
    for (;;) {
        size_t s = r->pos, e;
        while (r->pos < r->len && r->buf[r->pos] != '\n')
            ++r->pos;
        if (r->pos >= r->len)
            return -1;
        e = r->pos++;
        while (s < e && (r->buf[s] == ' ' || r->buf[s] == '\t'))
            ++s;
        if (s == e)
            return -1;
        if (r->buf[s] == '#')
            continue;
        if (e - s >= 6 && memcmp(r->buf + s, "endhdr", 6) == 0)
            break;
        if (e - s >= 7 && memcmp(r->buf + s, "chksum0", 7) == 0)
            chk = 1;
    }
    if (rd_u32(r, &magic, 1) < 0)
        return -1;
    if (magic != 0x11223344u) {
        if (bswap32(magic) != 0x11223344u)
            return -1;
        r->swap = 1;
    }
    r->do_chk = chk;
    r->chk = 0;
    return 0;
}

/* ref: s3file.c:551-570 */
static int
rd_verify(rd_t *r)
{
    uint32_t want, have = r->chk;
    if (!r->do_chk)
        return 0;
    r->do_chk = 0;
    if (rd_u32(r, &want, 1) < 0)
        return -1;
    return want == have ? 0 : -1;
}

/* ------------------------------------------------------------------ */
/* log math (ref: logmath.c:61-163, 283-302)                           */
/* ------------------------------------------------------------------ */
int32_t
orc_logmath_log(double base, int shift, double p)
{
    double inv = 1.0 / log(base);
    if (p <= 0)
        return INT32_MIN >> (shift + 2);
    return (int32_t)(log(p) * inv) >> shift;
}

static int32_t
ln_to_log(double base, int shift, double lnp)
{
    double inv = 1.0 / log(base);
    return (int32_t)(lnp * inv) >> shift;
}

/* The 8-bit add table: entry i>>shift keeps the FIRST (largest) rounded value of
 * log_b(1 + b^-i); table padded to >=256 entries. */
int
orc_logadd_table8(double base, int shift, uint8_t *out256)
{
    double inv = 1.0 / log(base), byx = 1.0;
    uint32_t maxyx = (uint32_t)(log(2.0) / log(base) + 0.5) >> shift;
    uint32_t i;
    if (maxyx >= 256)
        return -1; /* would need a wider table: reference refuses too (ptm_mgau.c:741) */
    memset(out256, 0, 256);
    for (i = 0;; ++i) {
        double lobyx = log(1.0 + byx) * inv;
        int32_t k = (int32_t)(lobyx + 0.5 * (1 << shift)) >> shift;
        if ((i >> shift) < 256 && out256[i >> shift] == 0)
            out256[i >> shift] = (uint8_t)k;
        if (k <= 0)
            break;
        byx /= base;
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* model loading                                                       */
/* ------------------------------------------------------------------ */
/* ref: ms_gauden.c:105-202 */
static float *
load_gauden_file(const char *dir, const char *name, int32_t *n_mgau, int32_t *n_feat,
                 int32_t *n_density, int32_t *featlen)
{
    rd_t r;
    int32_t n, i, blk = 0;
    float *buf;
    if (rd_open(&r, dir, name) < 0)
        return NULL;
    if (rd_s3_header(&r) < 0 || rd_u32(&r, n_mgau, 1) < 0 || rd_u32(&r, n_feat, 1) < 0
        || rd_u32(&r, n_density, 1) < 0 || *n_feat > ORC_MAX_FEAT || rd_u32(&r, featlen, *n_feat) < 0
        || rd_u32(&r, &n, 1) < 0)
        goto fail;
    for (i = 0; i < *n_feat; ++i)
        blk += featlen[i];
    if (n != *n_mgau * *n_density * blk)
        goto fail;
    buf = malloc((size_t)n * sizeof(float));
    if (rd_u32(&r, buf, n) < 0 || rd_verify(&r) < 0) {
        free(buf);
        goto fail;
    }
    free(r.buf);
    return buf;
fail:
    free(r.buf);
    return NULL;
}

/* ref: ms_gauden.c:217-258 (precompute) -- double math, int truncation, float cast */
static void
gauden_precompute(orc_model_t *m, float varfloor)
{
    int c, f, d, i;
    double base = m->logbase;
    for (c = 0; c < m->n_mgau; ++c)
        for (f = 0; f < m->n_feat; ++f) {
            int L = m->featlen[f];
            float *v = m->var + m->gau_off[c * m->n_feat + f];
            float *det = m->det + (size_t)(c * m->n_feat + f) * m->n_density;
            for (d = 0; d < m->n_density; ++d) {
                float acc = 0;
                for (i = 0; i < L; ++i) {
                    float *fv = v + d * L + i;
                    if (*fv < varfloor)
                        *fv = varfloor;
                    acc += (float)orc_logmath_log(base, 0, 1.0 / sqrt(*fv * 2.0 * M_PI));
                    *fv = (float)ln_to_log(base, 0, (1.0 / (*fv * 2.0)));
                }
                det[d] = acc;
            }
        }
}

/* ref: ptm_mgau.c:456-609 (sendump) */
static int
load_sendump(orc_model_t *m, const char *dir)
{
    rd_t r;
    int32_t n, rows, cols, n_clust = 0, n_bits = 8, n_feat, n_density, n_sen;
    const uint8_t *cb = NULL;
    int f, k, s;
    if (rd_open(&r, dir, "sendump") < 0)
        return -2; /* no sendump: the caller tries mixture_weights */
    n_feat = m->n_feat;
    n_density = m->n_density;
    n_sen = m->n_sen;
    if (rd_u32(&r, &n, 1) < 0)
        goto fail;
    if (n < 1 || n > 999) {
        n = bswap32(n);
        if (n < 1 || n > 999)
            goto fail;
        r.swap = 1;
    }
    if (r.pos + n > r.len || r.buf[r.pos + n - 1] != 0)
        goto fail;
    r.pos += n; /* title */
    if (rd_u32(&r, &n, 1) < 0 || r.pos + n > r.len || r.buf[r.pos + n - 1] != 0)
        goto fail;
    r.pos += n; /* header */
    for (;;) {
        const char *p;
        if (rd_u32(&r, &n, 1) < 0)
            goto fail;
        if (n == 0)
            break;
        if (r.pos + n > r.len)
            goto fail;
        p = (const char *)r.buf + r.pos;
        if (!strncmp(p, "feature_count ", 14))
            n_feat = atoi(p + 14);
        if (!strncmp(p, "mixture_count ", 14))
            n_density = atoi(p + 14);
        if (!strncmp(p, "model_count ", 12))
            n_sen = atoi(p + 12);
        if (!strncmp(p, "cluster_count ", 14))
            n_clust = atoi(p + 14);
        if (!strncmp(p, "cluster_bits ", 13))
            n_bits = atoi(p + 13);
        r.pos += n;
    }
    cols = n_sen;
    rows = n_density;
    if (n_clust == 0) {
        if (rd_u32(&r, &rows, 1) < 0 || rd_u32(&r, &cols, 1) < 0)
            goto fail;
    }
    if (n_feat != m->n_feat || n_density != m->n_density || n_sen != m->n_sen)
        goto fail;
    if (!(n_clust == 0 || n_clust == 15 || n_clust == 16) || !(n_bits == 8 || n_bits == 4))
        goto fail;
    if (n_clust == 15)
        n_clust = 16;
    if (n_clust) {
        cb = r.buf + r.pos;
        r.pos += n_clust;
    }
    m->mixw = calloc((size_t)n_feat * n_density * n_sen, 1);
    for (f = 0; f < n_feat; ++f) {
        int step = (n_bits == 4) ? (cols + 1) / 2 : cols;
        for (k = 0; k < rows; ++k) {
            const uint8_t *row = r.buf + r.pos;
            uint8_t *dst = m->mixw + ((size_t)f * n_density + k) * n_sen;
            r.pos += step;
            if (r.pos > r.len)
                goto fail;
            if (k >= n_density)
                continue;
            for (s = 0; s < n_sen; ++s) {
                if (cb) { /* ref: ptm_mgau.c:375-378: the nibble is selected by the low
                           * bit of the packed byte, not by the senone's parity; s2_semi
                           * selects by senone parity (s2_semi_mgau.c:733-757, 795-824) */
                    int dcw = row[s / 2];
                    if (m->kind == ORC_KIND_SEMI)
                        dcw = (s & 1) ? dcw >> 4 : dcw & 0x0f;
                    else
                        dcw = (dcw & 1) ? dcw >> 4 : dcw & 0x0f;
                    dst[s] = cb[dcw];
                } else
                    dst[s] = row[s];
            }
        }
    }
    free(r.buf);
    return 0;
fail:
    free(r.buf);
    return -1;
}

/* ref: ptm_mgau.c:611-692 (uncompressed mixture_weights, used when there is no sendump),
 * vector.c:87-113.  The floor is the "mixwfloor" default of config_defs.h:218-221, held
 * in a float32 by the caller (ptm_mgau.c:787). */
static void
sum_norm_f32(float *v, int n)
{
    double sum = 0.0, f;
    int i;
    for (i = 0; i < n; ++i)
        sum += v[i];
    if (sum != 0.0) {
        f = 1.0 / sum;
        for (i = 0; i < n; ++i)
            v[i] = (float)(v[i] * f);
    }
}

static int
load_mixw_float(orc_model_t *m, const char *dir)
{
    rd_t r;
    int32_t n_sen, n_feat, n_comp, n, i, f, c;
    const float flr_f = 1e-7f;
    const double flr = flr_f;
    float *pdf = NULL;
    if (rd_open(&r, dir, "mixture_weights") < 0)
        return -1;
    if (rd_s3_header(&r) < 0)
        goto fail;
    if (rd_u32(&r, &n_sen, 1) < 0 || rd_u32(&r, &n_feat, 1) < 0 || rd_u32(&r, &n_comp, 1) < 0
        || rd_u32(&r, &n, 1) < 0)
        goto fail;
    if (n_feat != m->n_feat || n_comp != m->n_density || n_sen != m->n_sen
        || n != n_sen * n_feat * n_comp)
        goto fail;
    m->mixw = calloc((size_t)n_feat * n_comp * n_sen, 1);
    pdf = malloc(sizeof(float) * n_comp);
    for (i = 0; i < n_sen; ++i)
        for (f = 0; f < n_feat; ++f) {
            if (rd_u32(&r, pdf, n_comp) < 0)
                goto fail;
            sum_norm_f32(pdf, n_comp);
            for (c = 0; c < n_comp; ++c)
                if (pdf[c] < flr)
                    pdf[c] = (float)flr;
            sum_norm_f32(pdf, n_comp);
            for (c = 0; c < n_comp; ++c) {
                int32_t q = -orc_logmath_log(m->logbase, ORC_SENSCR_SHIFT, pdf[c]);
                if (q > 159 || q < 0) /* MAX_NEG_MIXW */
                    q = 159;
                m->mixw[((size_t)f * n_comp + c) * n_sen + i] = (uint8_t)q;
            }
        }
    if (rd_verify(&r) < 0)
        goto fail;
    free(pdf);
    free(r.buf);
    return 0;
fail:
    free(pdf);
    free(r.buf);
    return -1;
}

/* ref: ms_senone.c:103-190 (senone_mixw_read): the continuous scorer's own quantisation --
 * shift-0 log, rounded before the >> 10, clamped at 255 -- stored [sen][feat][density]. */
static int
load_mixw_cont(orc_model_t *m, const char *dir)
{
    rd_t r;
    int32_t n_sen, n_feat, n_cw, n, i, f, c;
    const float flr_f = 1e-7f;
    const double flr = flr_f;
    float *pdf = NULL;
    if (rd_open(&r, dir, "mixture_weights") < 0)
        return -1;
    if (rd_s3_header(&r) < 0)
        goto fail;
    if (rd_u32(&r, &n_sen, 1) < 0 || rd_u32(&r, &n_feat, 1) < 0 || rd_u32(&r, &n_cw, 1) < 0
        || rd_u32(&r, &n, 1) < 0)
        goto fail;
    if (n_feat != m->n_feat || n_cw != m->n_density || n_sen != m->n_sen
        || n != n_sen * n_feat * n_cw)
        goto fail;
    m->mixw = calloc((size_t)n_feat * n_cw * n_sen, 1);
    pdf = malloc(sizeof(float) * n_cw);
    for (i = 0; i < n_sen; ++i)
        for (f = 0; f < n_feat; ++f) {
            if (rd_u32(&r, pdf, n_cw) < 0)
                goto fail;
            sum_norm_f32(pdf, n_cw);
            for (c = 0; c < n_cw; ++c)
                if (pdf[c] < flr)
                    pdf[c] = (float)flr;
            sum_norm_f32(pdf, n_cw);
            for (c = 0; c < n_cw; ++c) {
                int32_t p = -orc_logmath_log(m->logbase, 0, pdf[c]);
                p += (1 << (ORC_SENSCR_SHIFT - 1)) - 1;
                m->mixw[((size_t)i * n_feat + f) * n_cw + c]
                    = (uint8_t)(p < (255 << ORC_SENSCR_SHIFT) ? p >> ORC_SENSCR_SHIFT : 255);
            }
        }
    if (rd_verify(&r) < 0)
        goto fail;
    free(pdf);
    free(r.buf);
    return 0;
fail:
    free(pdf);
    free(r.buf);
    return -1;
}

/* ref: bin_mdef.c:333-520 */
static int
load_mdef(orc_model_t *m, const char *dir)
{
    rd_t r;
    int32_t val, hdr[10], i, j, sseq_size;
    size_t p;
    if (rd_open(&r, dir, "mdef") < 0)
        return -1;
    if (r.len < 4)
        goto fail;
    if (memcmp(r.buf, "BMDF", 4) == 0)
        r.swap = 0;
    else if (memcmp(r.buf, "FDMB", 4) == 0)
        r.swap = 1;
    else
        goto fail;
    r.pos = 4;
    if (rd_u32(&r, &val, 1) < 0 || val > 1)
        goto fail;
    if (rd_u32(&r, &val, 1) < 0)
        goto fail;
    r.pos += val;
    if (rd_u32(&r, hdr, 10) < 0)
        goto fail;
    m->n_ciphone = hdr[0];
    m->n_phone = hdr[1];
    m->n_emit = hdr[2];
    m->n_ci_sen = hdr[3];
    m->n_sen = hdr[4];
    m->n_tmat_mdef = hdr[5];
    m->n_sseq = hdr[6];
    m->n_ctx = hdr[7];
    m->n_cd_tree = hdr[8];
    m->sil = hdr[9];
    if (m->n_emit <= 0)
        goto fail; /* heterogeneous topologies: not needed by bundled models */
    m->ciname = calloc(m->n_ciphone, sizeof(char *));
    p = r.pos;
    for (i = 0; i < m->n_ciphone; ++i) {
        m->ciname[i] = strdup((char *)r.buf + p);
        p += strlen((char *)r.buf + p) + 1;
    }
    p = r.pos + (((p - r.pos) + 3) & ~(size_t)3);
    p += (size_t)m->n_cd_tree * 8; /* cd_tree_t: not needed on the hot path */
    m->ph_ssid = malloc(sizeof(int32_t) * m->n_phone);
    m->ph_tmat = malloc(sizeof(int32_t) * m->n_phone);
    m->ph_ci = malloc(sizeof(int32_t) * m->n_phone);
    for (i = 0; i < m->n_phone; ++i) {
        uint32_t a, b;
        memcpy(&a, r.buf + p, 4);
        memcpy(&b, r.buf + p + 4, 4);
        if (r.swap) {
            a = bswap32(a);
            b = bswap32(b);
        }
        m->ph_ssid[i] = (int32_t)a;
        m->ph_tmat[i] = (int32_t)b;
        /* ref: bin_mdef.h:167 pid2ci */
        m->ph_ci[i] = (i < m->n_ciphone) ? i : r.buf[p + 9];
        p += 12;
    }
    r.pos = p;
    if (rd_u32(&r, &sseq_size, 1) < 0 || sseq_size != m->n_sseq * m->n_emit)
        goto fail;
    m->sseq = malloc(sizeof(uint16_t) * sseq_size);
    memcpy(m->sseq, r.buf + r.pos, sizeof(uint16_t) * sseq_size);
    if (r.swap)
        for (i = 0; i < sseq_size; ++i)
            m->sseq[i] = (uint16_t)((m->sseq[i] >> 8) | (m->sseq[i] << 8));
    /* sen2cimap: first phone (in phone order) whose sseq contains s (bin_mdef.c:470-516) */
    m->sen2cb = malloc(m->n_sen);
    {
        int16_t *map = malloc(sizeof(int16_t) * m->n_sen);
        for (i = 0; i < m->n_sen; ++i)
            map[i] = -1;
        for (i = 0; i < m->n_phone; ++i)
            for (j = 0; j < m->n_emit; ++j) {
                int s = m->sseq[m->ph_ssid[i] * m->n_emit + j];
                if (map[s] == -1)
                    map[s] = (int16_t)m->ph_ci[i];
            }
        for (i = 0; i < m->n_sen; ++i)
            m->sen2cb[i] = (uint8_t)map[i];
        free(map);
    }
    /* sil = id of "SIL" (bin_mdef.c:519) */
    for (i = 0; i < m->n_ciphone; ++i)
        if (strcmp(m->ciname[i], "SIL") == 0)
            m->sil = i;
    free(r.buf);
    return 0;
fail:
    free(r.buf);
    return -1;
}

/* ref: tmat.c:125-225, vector.c:87-123 */
static int
load_tmat(orc_model_t *m, const char *dir, double tpfloor)
{
    rd_t r;
    int32_t n_tmat, n_src, n_dst, n, i, j, k;
    float *tp;
    if (rd_open(&r, dir, "transition_matrices") < 0)
        return -1;
    if (rd_s3_header(&r) < 0 || rd_u32(&r, &n_tmat, 1) < 0 || rd_u32(&r, &n_src, 1) < 0
        || rd_u32(&r, &n_dst, 1) < 0 || rd_u32(&r, &n, 1) < 0 || n_dst != n_src + 1
        || n != n_tmat * n_src * n_dst)
        goto fail;
    m->n_tmat = n_tmat;
    m->n_state = n_src;
    m->tp = malloc((size_t)n);
    tp = malloc(sizeof(float) * n_src * n_dst);
    for (i = 0; i < n_tmat; ++i) {
        if (rd_u32(&r, tp, n_src * n_dst) < 0)
            goto fail;
        for (j = 0; j < n_src; ++j) {
            float *row = tp + j * n_dst;
            int pass;
            for (pass = 0; pass < 2; ++pass) {
                double sum = 0.0;
                for (k = 0; k < n_dst; ++k)
                    sum += row[k];
                if (sum != 0.0) {
                    double f = 1.0 / sum;
                    for (k = 0; k < n_dst; ++k)
                        row[k] = (float)(row[k] * f);
                }
                if (pass == 0)
                    for (k = 0; k < n_dst; ++k)
                        if (row[k] != 0.0 && row[k] < tpfloor)
                            row[k] = (float)tpfloor;
            }
            for (k = 0; k < n_dst; ++k) {
                int ltp = (-orc_logmath_log(m->logbase, 0, row[k])) >> ORC_SENSCR_SHIFT;
                if (ltp > 255)
                    ltp = 255;
                m->tp[(i * n_src + j) * n_dst + k] = (uint8_t)ltp;
            }
        }
    }
    free(tp);
    if (rd_verify(&r) < 0)
        goto fail;
    free(r.buf);
    return 0;
fail:
    free(r.buf);
    return -1;
}

orc_model_t *
orc_model_load(const char *dir, double logbase, float varfloor, double tmatfloor)
{
    orc_model_t *m = calloc(1, sizeof(*m));
    int32_t a[3], fl[ORC_MAX_FEAT], c, f;
    int64_t off = 0;
    m->logbase = logbase;
    m->lmath_zero = INT32_MIN >> 2;
    if ((m->mean = load_gauden_file(dir, "means", &m->n_mgau, &m->n_feat, &m->n_density, m->featlen))
        == NULL)
        goto fail;
    if ((m->var = load_gauden_file(dir, "variances", &a[0], &a[1], &a[2], fl)) == NULL)
        goto fail;
    if (a[0] != m->n_mgau || a[1] != m->n_feat || a[2] != m->n_density
        || memcmp(fl, m->featlen, sizeof(int32_t) * m->n_feat))
        goto fail;
    m->gau_off = malloc(sizeof(int64_t) * m->n_mgau * m->n_feat);
    for (f = 0; f < m->n_feat; ++f) {
        m->featoff[f] = m->blk;
        m->blk += m->featlen[f];
    }
    m->featoff[m->n_feat] = m->blk;
    for (c = 0; c < m->n_mgau; ++c)
        for (f = 0; f < m->n_feat; ++f) {
            m->gau_off[c * m->n_feat + f] = off;
            off += (int64_t)m->n_density * m->featlen[f];
        }
    m->det = calloc((size_t)m->n_mgau * m->n_feat * m->n_density, sizeof(float));
    gauden_precompute(m, varfloor);
    if (load_mdef(m, dir) < 0)
        goto fail;
    /* acmod_load_am tries ptm_mgau (needs n_mgau == n_ciphone, ptm_mgau.c:760), then
     * s2_semi_mgau (needs a single codebook, s2_semi_mgau.c:947-949) */
    if (m->n_mgau == m->n_ciphone)
        m->kind = ORC_KIND_PTM;
    else if (m->n_mgau == 1) {
        m->kind = ORC_KIND_SEMI;
        memset(m->sen2cb, 0, m->n_sen);
    } else if (m->n_mgau == m->n_sen)
        m->kind = ORC_KIND_CONT; /* ms_mgau with the 1-to-1 senone-codebook map (ms_senone.c:262-275) */
    else
        goto fail;
    if (m->kind == ORC_KIND_CONT)
        c = load_mixw_cont(m, dir);
    else if ((c = load_sendump(m, dir)) == -2)
        c = load_mixw_float(m, dir);
    if (c < 0)
        goto fail;
    if (load_tmat(m, dir, tmatfloor) < 0)
        goto fail;
    if (orc_logadd_table8(logbase, ORC_SENSCR_SHIFT, m->lut8) < 0)
        goto fail;
    return m;
fail:
    orc_model_free(m);
    return NULL;
}

void
orc_model_free(orc_model_t *m)
{
    int i;
    if (!m)
        return;
    free(m->mean);
    free(m->var);
    free(m->det);
    free(m->gau_off);
    free(m->mixw);
    free(m->sen2cb);
    free(m->sseq);
    free(m->ph_ssid);
    free(m->ph_tmat);
    free(m->ph_ci);
    if (m->ciname)
        for (i = 0; i < m->n_ciphone; ++i)
            free(m->ciname[i]);
    free(m->ciname);
    free(m->tp);
    free(m);
}

int
orc_model_dims(const orc_model_t *m, int32_t *out)
{
    out[0] = m->n_mgau;
    out[1] = m->n_feat;
    out[2] = m->n_density;
    out[3] = m->featlen[0];
    out[4] = m->n_sen;
    out[5] = m->n_sseq;
    out[6] = m->n_emit;
    out[7] = m->n_tmat;
    out[8] = m->n_ciphone;
    out[9] = m->n_phone;
    out[10] = m->sil;
    return 0;
}

int
orc_model_copy(const orc_model_t *m, float *mean, float *var, float *det, uint8_t *mixw,
               uint8_t *sen2cb, uint8_t *tp, uint16_t *sseq, uint8_t *lut8)
{
    size_t ng = (size_t)m->n_mgau * m->n_density * m->blk;
    memcpy(mean, m->mean, ng * sizeof(float));
    memcpy(var, m->var, ng * sizeof(float));
    memcpy(det, m->det, (size_t)m->n_mgau * m->n_feat * m->n_density * sizeof(float));
    memcpy(mixw, m->mixw, (size_t)m->n_feat * m->n_density * m->n_sen);
    memcpy(sen2cb, m->sen2cb, m->n_sen);
    memcpy(tp, m->tp, (size_t)m->n_tmat * m->n_state * (m->n_state + 1));
    memcpy(sseq, m->sseq, (size_t)m->n_sseq * m->n_emit * sizeof(uint16_t));
    memcpy(lut8, m->lut8, 256);
    return 0;
}

int
orc_phone_table(const orc_model_t *m, int32_t *ssid, int32_t *tmat, int32_t *ci)
{
    memcpy(ssid, m->ph_ssid, sizeof(int32_t) * m->n_phone);
    memcpy(tmat, m->ph_tmat, sizeof(int32_t) * m->n_phone);
    memcpy(ci, m->ph_ci, sizeof(int32_t) * m->n_phone);
    return m->n_phone;
}

/* ------------------------------------------------------------------ */
/* PTM scorer                                                          */
/* ------------------------------------------------------------------ */
typedef struct {
    int32_t cw, score;
} topn_t;

struct orc_ptm_s {
    const orc_model_t *m;
    int topn, ds;
    topn_t *hist[2]; /* [mgau][feat][topn], ring of 2 (ptm_mgau.c:803) */
    uint8_t *cb_active[2];
    int cur;
    int32_t frame_idx; /* mgau_t.frame_idx (acmod.h:110) */
    int32_t topn_beam[ORC_MAX_FEAT]; /* s2_semi only */
    uint8_t *topn_n[2];              /* s2_semi: entries inside the beam, [feat] per slot */
};

orc_ptm_t *
orc_ptm_new(const orc_model_t *m, int topn, int ds_ratio)
{
    orc_ptm_t *p = calloc(1, sizeof(*p));
    int i;
    if (topn > ORC_MAX_TOPN)
        topn = ORC_MAX_TOPN;
    p->m = m;
    p->topn = topn;
    p->ds = ds_ratio < 1 ? 1 : ds_ratio;
    for (i = 0; i < 2; ++i) {
        p->hist[i] = malloc(sizeof(topn_t) * m->n_mgau * m->n_feat * topn);
        p->cb_active[i] = malloc(m->n_mgau);
        p->topn_n[i] = calloc(ORC_MAX_FEAT, 1);
    }
    memcpy(p->topn_beam, m->topn_beam, sizeof(p->topn_beam));
    orc_ptm_reset(p);
    return p;
}

void
orc_ptm_free(orc_ptm_t *p)
{
    if (!p)
        return;
    free(p->hist[0]);
    free(p->hist[1]);
    free(p->cb_active[0]);
    free(p->cb_active[1]);
    free(p->topn_n[0]);
    free(p->topn_n[1]);
    free(p);
}

/* ref: ptm_mgau.c:694-720 */
void
orc_ptm_reset(orc_ptm_t *p)
{
    int i, j, k;
    for (i = 0; i < 2; ++i) {
        for (j = 0; j < p->m->n_mgau * p->m->n_feat; ++j)
            for (k = 0; k < p->topn; ++k) {
                p->hist[i][j * p->topn + k].cw = k;
                p->hist[i][j * p->topn + k].score = ORC_WORST_DIST;
            }
        memset(p->cb_active[i], 1, p->m->n_mgau);
    }
    p->cur = 0;
    p->frame_idx = 0;
}

/* The codewords of history slot 1: what frame 0 of a following utterance copies in (ptm_mgau.c:
 * 426-440: lastf = hist[n_fast_hist - 1] when fast_eval_idx == 0; the lists are not reset between
 * utterances or passes).  Scores are irrelevant: eval_topn re-scores every entry first. */
void
orc_ptm_get_carried(const orc_ptm_t *p, uint8_t *cw)
{
    int j, k, CS = p->m->n_mgau * p->m->n_feat;
    for (j = 0; j < CS; ++j)
        for (k = 0; k < p->topn; ++k)
            cw[j * p->topn + k] = (uint8_t)p->hist[1][j * p->topn + k].cw;
}

void
orc_ptm_set_carried(orc_ptm_t *p, const uint8_t *cw)
{
    int j, k, CS = p->m->n_mgau * p->m->n_feat;
    for (j = 0; j < CS; ++j)
        for (k = 0; k < p->topn; ++k) {
            p->hist[1][j * p->topn + k].cw = cw[j * p->topn + k];
            p->hist[1][j * p->topn + k].score = ORC_WORST_DIST;
        }
}

/* The fp32 distance, in the reference's op order (ptm_mgau.c:63-68,106-127):
 * per dimension: diff = x - mu; sq = diff*diff; c = sq*var; d = d - c. */
static float
gau_dist(const float *x, const float *mean, const float *var, float det, int L)
{
    float d = det;
    int j;
    for (j = 0; j < L; ++j) {
        float diff = x[j] - mean[j];
        float sq = diff * diff;
        float c = sq * var[j];
        d = d - c;
    }
    return d;
}

static int32_t
dist_to_int(float d)
{
    if (d < (float)INT32_MIN)
        return INT32_MIN;
    return (int32_t)d;
}

/* ref: ptm_mgau.c:86-135 (eval_topn) + :70-84 (insertion_sort_topn, strict >) */
static void
rescore_topn(orc_ptm_t *p, topn_t *tn, int c, int f, const float *x)
{
    const orc_model_t *m = p->m;
    int L = m->featlen[f], i, j;
    const float *mean = m->mean + m->gau_off[c * m->n_feat + f];
    const float *var = m->var + m->gau_off[c * m->n_feat + f];
    const float *det = m->det + (size_t)(c * m->n_feat + f) * m->n_density;
    for (i = 0; i < p->topn; ++i) {
        int cw = tn[i].cw;
        topn_t v;
        v.cw = cw;
        v.score = dist_to_int(gau_dist(x, mean + cw * L, var + cw * L, det[cw], L));
        for (j = i - 1; j >= 0 && v.score > tn[j].score; --j)
            tn[j + 1] = tn[j];
        tn[j + 1] = v;
    }
}

/* ref: ptm_mgau.c:150-225 (eval_cb) + :139-148 (insertion_sort_cb, >=).
 * The early-out of the reference is result-neutral (terms are >= 0). */
static void
scan_codebook(orc_ptm_t *p, topn_t *tn, int c, int f, const float *x)
{
    const orc_model_t *m = p->m;
    int L = m->featlen[f], cw, i, N = p->topn;
    const float *mean = m->mean + m->gau_off[c * m->n_feat + f];
    const float *var = m->var + m->gau_off[c * m->n_feat + f];
    const float *det = m->det + (size_t)(c * m->n_feat + f) * m->n_density;
    for (cw = 0; cw < m->n_density; ++cw) {
        float thresh = (float)tn[N - 1].score;
        float d = gau_dist(x, mean + cw * L, var + cw * L, det[cw], L);
        int32_t id;
        if (d < thresh)
            continue;
        for (i = 0; i < N; ++i)
            if (tn[i].cw == cw)
                break;
        if (i < N)
            continue;
        id = dist_to_int(d);
        for (i = N - 2; i >= 0 && id >= tn[i].score; --i)
            tn[i + 1] = tn[i];
        tn[i + 1].cw = cw;
        tn[i + 1].score = id;
    }
}

/* ref: s2_semi_mgau.c:110-169 (eval_cb of the semi-continuous scorer).  Unlike the PTM
 * one its early-out is NOT result-neutral: a density is dropped when a partial sum (before the
 * last dimension at the latest) is below the worst score as a float, otherwise its TRUNCATED
 * score is compared with the worst as integers. */
static void
scan_codebook_semi(orc_ptm_t *p, topn_t *tn, int f, const float *x)
{
    const orc_model_t *m = p->m;
    int L = m->featlen[f], cw, i, j, N = p->topn;
    const float *mean = m->mean + m->gau_off[f];
    const float *var = m->var + m->gau_off[f];
    const float *det = m->det + (size_t)f * m->n_density;
    for (cw = 0; cw < m->n_density; ++cw) {
        float d = det[cw];
        int32_t id;
        for (j = 0; j < L && d >= tn[N - 1].score; ++j) {
            float diff = x[j] - mean[cw * L + j];
            float sq = diff * diff;
            float c = sq * var[cw * L + j];
            d = d - c;
        }
        if (j < L)
            continue;
        id = dist_to_int(d);
        if (id < tn[N - 1].score)
            continue;
        for (i = 0; i < N; ++i)
            if (tn[i].cw == cw)
                break;
        if (i < N)
            continue;
        for (i = N - 2; i >= 0 && id >= tn[i].score; --i)
            tn[i + 1] = tn[i];
        tn[i + 1].cw = cw;
        tn[i + 1].score = id;
    }
}

void
orc_model_set_topn_beam(orc_model_t *m, const int32_t *beam, int n)
{
    int i, maxn = 0;
    for (i = 0; i < n && i < ORC_MAX_FEAT; ++i)
        if (beam[i] > maxn)
            maxn = beam[i];
    for (i = 0; i < ORC_MAX_FEAT; ++i)
        m->topn_beam[i] = (uint8_t)(i < n ? beam[i] : maxn);
}

/* ref: s2_semi_mgau.c:829-875 and callees */
static int
semi_frame_eval(orc_ptm_t *p, int16_t *senscr, const uint8_t *active, int32_t n_active,
                const float *feat, int32_t frame, int32_t compallsen, int32_t *topn_out)
{
    const orc_model_t *m = p->m;
    int N = p->topn, f, k, i, lastsen;
    int slot = frame % 2;
    topn_t *cur = p->hist[slot];
    uint8_t *nn = p->topn_n[slot];
    memset(senscr, 0, sizeof(int16_t) * m->n_sen);
    for (f = 0; f < m->n_feat; ++f) {
        topn_t *tn = cur + f * N;
        if (frame >= p->frame_idx) {
            int32_t norm;
            memcpy(tn, p->hist[slot ? slot - 1 : 1] + f * N, sizeof(topn_t) * N);
            rescore_topn(p, tn, 0, f, feat + m->featoff[f]); /* :68-108, same as the PTM one */
            if (frame % p->ds == 0)
                scan_codebook_semi(p, tn, f, feat + m->featoff[f]);
            /* mgau_norm (:184-202) */
            norm = tn[0].score >> ORC_SENSCR_SHIFT;
            for (k = 0; k < N; ++k) {
                tn[k].score = -((tn[k].score >> ORC_SENSCR_SHIFT) - norm);
                if (tn[k].score > ORC_MAX_NEG_ASCR)
                    tn[k].score = ORC_MAX_NEG_ASCR;
                if (p->topn_beam[f] && tn[k].score > p->topn_beam[f])
                    break;
            }
            nn[f] = (uint8_t)k;
        }
        /* get_scores_{8b,4b}_feat[_all] (:204-827): no normalisation over senones, senones
         * that are not listed stay 0 */
        for (lastsen = i = 0; i < (compallsen ? m->n_sen : n_active); ++i) {
            int sen = compallsen ? i : active[i] + lastsen;
            int fden = 0;
            lastsen = sen;
            for (k = 0; k < nn[f]; ++k) {
                int v = m->mixw[((size_t)f * m->n_density + tn[k].cw) * m->n_sen + sen] + tn[k].score;
                if (k == 0)
                    fden = v;
                else {
                    int d, r;
                    if (fden > v) {
                        d = fden - v;
                        r = v;
                    } else {
                        d = v - fden;
                        r = fden;
                    }
                    fden = r - m->lut8[d];
                }
            }
            senscr[sen] = (int16_t)(senscr[sen] + fden);
        }
    }
    p->cur = slot;
    if (topn_out)
        for (i = 0; i < m->n_feat * N; ++i) {
            *topn_out++ = cur[i].cw;
            *topn_out++ = cur[i].score;
        }
    return 0;
}

/* ref: ptm_mgau.c:408-454 and callees */
int
orc_ptm_frame_eval(orc_ptm_t *p, int16_t *senscr, const uint8_t *active, int32_t n_active,
                   const float *feat, int32_t frame, int32_t compallsen, int32_t *topn_out)
{
    const orc_model_t *m = p->m;
    int N = p->topn, c, f, k, i, lastsen;
    int slot = frame % 2;
    topn_t *cur = p->hist[slot];
    uint8_t *act = p->cb_active[slot];
    int32_t best;

    if (m->kind == ORC_KIND_SEMI)
        return semi_frame_eval(p, senscr, active, n_active, feat, frame, compallsen, topn_out);
    if (m->kind == ORC_KIND_CONT)
        return orc_cont_frame_eval(m, p->topn, senscr, active, n_active, feat, compallsen);
    if (frame >= p->frame_idx) {
        topn_t *prev = p->hist[slot ? slot - 1 : 1];
        memcpy(cur, prev, sizeof(topn_t) * m->n_mgau * m->n_feat * N);
        /* calc_cb_active (:297-321) */
        if (compallsen)
            memset(act, 1, m->n_mgau);
        else {
            memset(act, 0, m->n_mgau);
            for (lastsen = i = 0; i < n_active; ++i) {
                int sen = active[i] + lastsen;
                act[m->sen2cb[sen]] = 1;
                lastsen = sen;
            }
        }
        /* codebook_eval (:230-253) */
        for (c = 0; c < m->n_mgau; ++c)
            for (f = 0; f < m->n_feat; ++f)
                rescore_topn(p, cur + (c * m->n_feat + f) * N, c, f, feat + m->featoff[f]);
        if (frame % p->ds == 0)
            for (c = 0; c < m->n_mgau; ++c) {
                if (!act[c])
                    continue;
                for (f = 0; f < m->n_feat; ++f)
                    scan_codebook(p, cur + (c * m->n_feat + f) * N, c, f, feat + m->featoff[f]);
            }
        /* codebook_norm (:264-295) */
        for (f = 0; f < m->n_feat; ++f) {
            int32_t norm = ORC_WORST_SCORE;
            for (c = 0; c < m->n_mgau; ++c) {
                int32_t s;
                if (!act[c])
                    continue;
                s = cur[(c * m->n_feat + f) * N].score >> ORC_SENSCR_SHIFT;
                if (norm < s)
                    norm = s;
            }
            for (c = 0; c < m->n_mgau; ++c) {
                topn_t *tn = cur + (c * m->n_feat + f) * N;
                if (!act[c])
                    continue;
                for (k = 0; k < N; ++k) {
                    int32_t s = -((tn[k].score >> ORC_SENSCR_SHIFT) - norm);
                    tn[k].score = s > ORC_MAX_NEG_ASCR ? ORC_MAX_NEG_ASCR : s;
                }
            }
        }
    }
    p->cur = slot;
    /* senone_eval (:326-403) */
    memset(senscr, 0, sizeof(int16_t) * m->n_sen);
    if (compallsen)
        n_active = m->n_sen;
    best = INT32_MAX;
    for (lastsen = i = 0; i < n_active; ++i) {
        int sen = compallsen ? i : active[i] + lastsen;
        int cb, ascore = 0;
        lastsen = sen;
        cb = m->sen2cb[sen];
        if (!act[cb])
            for (f = 0; f < m->n_feat; ++f)
                for (k = 0; k < N; ++k)
                    cur[(cb * m->n_feat + f) * N + k].score = ORC_MAX_NEG_ASCR;
        for (f = 0; f < m->n_feat; ++f) {
            const topn_t *tn = cur + (cb * m->n_feat + f) * N;
            int fden = 0;
            for (k = 0; k < N; ++k) {
                int mw = m->mixw[((size_t)f * m->n_density + tn[k].cw) * m->n_sen + sen];
                int v = mw + tn[k].score;
                if (k == 0)
                    fden = v;
                else { /* fast_logmath_add, tied_mgau_common.h:100-117 */
                    int d, r;
                    if (fden > v) {
                        d = fden - v;
                        r = v;
                    } else {
                        d = v - fden;
                        r = fden;
                    }
                    fden = r - m->lut8[d];
                }
            }
            ascore += fden;
        }
        if (ascore < best)
            best = ascore;
        senscr[sen] = (int16_t)ascore;
    }
    for (i = 0; i < m->n_sen; ++i)
        senscr[i] = (int16_t)(senscr[i] - best);
    if (topn_out)
        for (i = 0; i < m->n_mgau * m->n_feat * N; ++i) {
            *topn_out++ = cur[i].cw;
            *topn_out++ = cur[i].score;
        }
    return 0;
}

/* what acmod_advance / acmod_rewind do to mgau_t.frame_idx (ref: acmod.c:367,748,760) */
void
orc_ptm_set_frame_idx(orc_ptm_t *p, int frame_idx)
{
    p->frame_idx = frame_idx;
}

int
orc_ptm_score_all(const orc_model_t *m, int topn, const float *feat, int T, int16_t *out)
{
    orc_ptm_t *p = orc_ptm_new(m, topn, 1);
    int t;
    for (t = 0; t < T; ++t) {
        orc_ptm_frame_eval(p, out + (size_t)t * m->n_sen, NULL, 0, feat + (size_t)t * m->blk, t, 1,
                           NULL);
        p->frame_idx = t + 1;
    }
    orc_ptm_free(p);
    return T;
}

int
orc_ptm_topn_all(const orc_model_t *m, int topn, const float *feat, int T, uint8_t *cw,
                 int32_t *score)
{
    orc_ptm_t *p = orc_ptm_new(m, topn, 1);
    int t, c, f, k, N = p->topn;
    size_t n = (size_t)m->n_mgau * m->n_feat * N;
    for (t = 0; t < T; ++t) {
        topn_t *cur = p->hist[t % 2], *prev = p->hist[(t + 1) % 2];
        memcpy(cur, prev, sizeof(topn_t) * n);
        for (c = 0; c < m->n_mgau; ++c)
            for (f = 0; f < m->n_feat; ++f) {
                const float *x = feat + (size_t)t * m->blk + m->featoff[f];
                rescore_topn(p, cur + (c * m->n_feat + f) * N, c, f, x);
                scan_codebook(p, cur + (c * m->n_feat + f) * N, c, f, x);
            }
        for (k = 0; k < (int)n; ++k) {
            cw[t * n + k] = (uint8_t)cur[k].cw;
            score[t * n + k] = cur[k].score;
        }
    }
    orc_ptm_free(p);
    return T;
}

/* ------------------------------------------------------------------ */
/* active list: bit vector -> uint8 deltas with lossy bridging         */
/* ref: acmod.c:947-999                                                */
/* ------------------------------------------------------------------ */
int
orc_flags2list(const uint32_t *bits, int n_sen, uint8_t *out)
{
    int s, n = 0, last = 0;
    for (s = 0; s < n_sen; ++s) {
        int delta;
        if (!(bits[s >> 5] & (1u << (s & 31))))
            continue;
        delta = s - last;
        while (delta > 255) {
            out[n++] = 255;
            delta -= 255;
        }
        out[n++] = (uint8_t)delta;
        last = s;
    }
    return n;
}

/* ------------------------------------------------------------------ */
/* HMM evaluation                                                      */
/* ------------------------------------------------------------------ */
#define CLAMPW(x) ((x) < ORC_WORST_SCORE ? ORC_WORST_SCORE : (x))

/* ref: hmm.c:482-567 */
static int32_t
hmm_eval_3(const uint8_t *tp, const uint16_t *sid, const int16_t *ss, int32_t *sc, int32_t *hi,
           int32_t *out_score, int32_t *out_hist)
{
#define TP3(i, j) (-(int32_t)tp[(i) * 4 + (j)])
    int32_t s2 = sc[2] - ss[sid[2]], s1 = sc[1] - ss[sid[1]], s0 = sc[0] - ss[sid[0]];
    int32_t s3, t0, t1, t2 = INT_MIN, best = ORC_WORST_SCORE;
    if (s1 > ORC_WORST_SCORE) {
        t1 = s2 + TP3(2, 3);
        if (TP3(1, 3) > ORC_TMAT_WORST)
            t2 = s1 + TP3(1, 3);
        if (t1 > t2) {
            s3 = t1;
            *out_hist = hi[2];
        } else {
            s3 = t2;
            *out_hist = hi[1];
        }
        s3 = CLAMPW(s3);
        *out_score = s3;
        best = s3;
    }
    t0 = s2 + TP3(2, 2);
    t1 = s1 + TP3(1, 2);
    if (TP3(0, 2) > ORC_TMAT_WORST)
        t2 = s0 + TP3(0, 2); /* else t2 keeps whatever the exit block left in it */
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
    s2 = CLAMPW(s2);
    if (s2 > best)
        best = s2;
    sc[2] = s2;
    t0 = s1 + TP3(1, 1);
    t1 = s0 + TP3(0, 1);
    if (t0 > t1)
        s1 = t0;
    else {
        s1 = t1;
        hi[1] = hi[0];
    }
    s1 = CLAMPW(s1);
    if (s1 > best)
        best = s1;
    sc[1] = s1;
    s0 = s0 + TP3(0, 0);
    s0 = CLAMPW(s0);
    if (s0 > best)
        best = s0;
    sc[0] = s0;
    return best;
#undef TP3
}

/* ref: hmm.c:166-304 */
static int32_t
hmm_eval_5(const uint8_t *tp, const uint16_t *sid, const int16_t *ss, int32_t *sc, int32_t *hi,
           int32_t *out_score, int32_t *out_hist)
{
#define TP5(i, j) (-(int32_t)tp[(i) * 6 + (j)])
    int32_t s5, s4, s3, s2, s1, s0, t0, t1, t2, best = ORC_WORST_SCORE;
    int j;
    s4 = sc[4] - ss[sid[4]];
    s3 = sc[3] - ss[sid[3]];
    if (s3 > ORC_WORST_SCORE) {
        t1 = s4 + TP5(4, 5);
        t2 = s3 + TP5(3, 5);
        if (t1 > t2) {
            s5 = t1;
            *out_hist = hi[4];
        } else {
            s5 = t2;
            *out_hist = hi[3];
        }
        s5 = CLAMPW(s5);
        *out_score = s5;
        best = s5;
    }
    s2 = sc[2] - ss[sid[2]];
    s1 = sc[1] - ss[sid[1]];
    s0 = sc[0] - ss[sid[0]];
    {
        int32_t sv[5];
        sv[0] = s0;
        sv[1] = s1;
        sv[2] = s2;
        sv[3] = s3;
        sv[4] = s4;
        /* states 4 and 3 are only updated when their skip source is alive (:191,:218);
         * state 2 always (:245).  All read the pre-update sv[]. */
        for (j = 4; j >= 2; --j) {
            int32_t nv;
            if (j > 2 && !(sv[j - 2] > ORC_WORST_SCORE))
                continue;
            t0 = sv[j] + TP5(j, j);
            t1 = sv[j - 1] + TP5(j - 1, j);
            t2 = sv[j - 2] + TP5(j - 2, j);
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
            nv = CLAMPW(nv);
            if (nv > best)
                best = nv;
            sc[j] = nv;
        }
    }
    t0 = s1 + TP5(1, 1);
    t1 = s0 + TP5(0, 1);
    if (t0 > t1)
        s1 = t0;
    else {
        s1 = t1;
        hi[1] = hi[0];
    }
    s1 = CLAMPW(s1);
    if (s1 > best)
        best = s1;
    sc[1] = s1;
    s0 = s0 + TP5(0, 0);
    s0 = CLAMPW(s0);
    if (s0 > best)
        best = s0;
    sc[0] = s0;
    return best;
#undef TP5
}

int32_t
orc_hmm_eval(int n_emit, const uint8_t *tp, const uint16_t *senid, const int16_t *senscr,
             int32_t *st)
{
    if (n_emit == 3)
        return hmm_eval_3(tp, senid, senscr, st, st + 5, st + 10, st + 11);
    if (n_emit == 5)
        return hmm_eval_5(tp, senid, senscr, st, st + 5, st + 10, st + 11);
    return ORC_WORST_SCORE;
}

/* ------------------------------------------------------------------ */
/* chain aligner (ref: state_align_search.c:46-268)                    */
/* ------------------------------------------------------------------ */
typedef struct {
    int32_t sc[5], hi[5], out_score, out_hist, frame;
} chmm_t;

typedef struct {
    const orc_model_t *m;
    int n_phones, E;
    chmm_t *h;
    const int32_t *ssid, *tmat, *sf, *ef;
    int32_t best_score;
    int32_t *tokens; /* [T][n_states][2] */
    int n_renorm;
} chain_t;

static void
chain_init(chain_t *c, const orc_model_t *m, int n_phones, const int32_t *ssid,
           const int32_t *tmat, const int32_t *sf, const int32_t *ef, int T)
{
    int i, j;
    memset(c, 0, sizeof(*c));
    c->m = m;
    c->n_phones = n_phones;
    c->E = m->n_emit;
    c->ssid = ssid;
    c->tmat = tmat;
    c->sf = sf;
    c->ef = ef;
    c->h = calloc(n_phones, sizeof(chmm_t));
    for (i = 0; i < n_phones; ++i) { /* hmm_clear (hmm.c:121-135) */
        for (j = 0; j < 5; ++j) {
            c->h[i].sc[j] = ORC_WORST_SCORE;
            c->h[i].hi[j] = -1;
        }
        c->h[i].out_score = ORC_WORST_SCORE;
        c->h[i].out_hist = -1;
        c->h[i].frame = -1;
    }
    c->tokens = malloc(sizeof(int32_t) * 2 * (size_t)T * n_phones * c->E);
    /* start: hmm_enter(hmms, 0, 0, 0) (:52) */
    c->h[0].sc[0] = 0;
    c->h[0].hi[0] = 0;
    c->h[0].frame = 0;
    c->best_score = 0;
}

/* everything in state_align_search_step after acmod_score (:191-212) */
static void
chain_step(chain_t *c, const int16_t *senscr, int t)
{
    const orc_model_t *m = c->m;
    int E = c->E, i, j, nf = t + 1, ns = c->n_phones * E;
    int32_t bs = ORC_WORST_SCORE;
    int32_t *tok = c->tokens + (size_t)t * ns * 2;
    /* renormalize (:193-197, hmm.c:150-161) */
    if (c->best_score - 0x300000 < ORC_WORST_SCORE) {
        for (i = 0; i < c->n_phones; ++i) {
            for (j = 0; j < E; ++j)
                if (c->h[i].sc[j] > ORC_WORST_SCORE)
                    c->h[i].sc[j] -= c->best_score;
            if (c->h[i].out_score > ORC_WORST_SCORE)
                c->h[i].out_score -= c->best_score;
        }
        c->n_renorm++;
    }
    /* evaluate_hmms (:66-86) */
    for (i = 0; i < c->n_phones; ++i) {
        chmm_t *h = c->h + i;
        int32_t s;
        if (h->frame < t)
            continue;
        s = (E == 3 ? hmm_eval_3 : hmm_eval_5)(m->tp + (size_t)c->tmat[i] * E * (E + 1),
                                                 m->sseq + (size_t)c->ssid[i] * E, senscr, h->sc,
                                                 h->hi, &h->out_score, &h->out_hist);
        if (s > bs)
            bs = s;
    }
    c->best_score = bs;
    /* prune_hmms (:88-106) */
    for (i = 0; i < c->n_phones; ++i) {
        if (c->h[i].frame < t)
            continue;
        if (nf > c->ef[i])
            continue;
        c->h[i].frame = nf;
    }
    /* phone_transition (:108-133) */
    for (i = 0; i < c->n_phones - 1; ++i) {
        chmm_t *h = c->h + i, *nh = h + 1;
        if (h->frame != nf)
            continue;
        if (nf < c->sf[i + 1])
            continue;
        if (nh->frame < t || h->out_score > nh->sc[0]) {
            nh->sc[0] = h->out_score;
            nh->hi[0] = h->out_hist;
            nh->frame = nf;
        }
    }
    /* record_transitions (:149-175) */
    memset(tok, 0xff, sizeof(int32_t) * 2 * ns);
    for (i = 0; i < c->n_phones; ++i) {
        chmm_t *h = c->h + i;
        if (h->frame < t)
            continue;
        for (j = 0; j < E; ++j) {
            int si = i * E + j;
            tok[si * 2] = h->hi[j];
            tok[si * 2 + 1] = h->sc[j];
            h->hi[j] = si;
        }
    }
}

/* ref: state_align_search.c:215-268 */
static int
chain_finish(chain_t *c, int T, int32_t *st_start, int32_t *st_dur, int32_t *st_score)
{
    int ns = c->n_phones * c->E, cur_frame, last_frame = T;
    chmm_t *fin = c->h + c->n_phones - 1;
    int32_t last_id = fin->out_hist, last_score = fin->out_score, cur_id = last_id, cur_score;
    if (last_id == -1)
        return -1;
    for (cur_frame = T - 2; cur_frame >= 0; --cur_frame) {
        const int32_t *tok = c->tokens + ((size_t)cur_frame * ns + cur_id) * 2;
        cur_id = tok[0];
        cur_score = tok[1];
        if (cur_id == -1)
            return -1;
        if (cur_id != last_id) {
            st_start[last_id] = cur_frame + 1;
            st_dur[last_id] = last_frame - st_start[last_id];
            st_score[last_id] = last_score - cur_score;
            last_id = cur_id;
            last_score = cur_score;
            last_frame = cur_frame + 1;
        }
    }
    st_start[0] = 0;
    st_dur[0] = last_frame;
    return 0;
}

int
orc_state_align_dense(const orc_model_t *m, const int16_t *senscr, int T, int n_phones,
                      const int32_t *ssid, const int32_t *tmat, const int32_t *sf, const int32_t *ef,
                      int32_t *st_start, int32_t *st_dur, int32_t *st_score, int32_t *tokens,
                      orc_align_out_t *out)
{
    chain_t c;
    int t, ns = n_phones * m->n_emit;
    chain_init(&c, m, n_phones, ssid, tmat, sf, ef, T);
    for (t = 0; t < T; ++t)
        chain_step(&c, senscr + (size_t)t * m->n_sen, t);
    out->rv = chain_finish(&c, T, st_start, st_dur, st_score);
    out->best_score = c.best_score;
    out->n_renorm = c.n_renorm;
    if (tokens)
        memcpy(tokens, c.tokens, sizeof(int32_t) * 2 * (size_t)T * ns);
    free(c.h);
    free(c.tokens);
    return out->rv;
}

int
orc_state_align(const orc_model_t *m, int topn, const float *feat, int T, int n_phones,
                const int32_t *ssid, const int32_t *tmat, const int32_t *sf, const int32_t *ef,
                const uint32_t *init_active, int compallsen, int32_t *st_start, int32_t *st_dur,
                int32_t *st_score, int32_t *tokens, int16_t *senscr_out, orc_align_out_t *out)
{
    return orc_state_align2(m, topn, feat, T, n_phones, ssid, tmat, sf, ef, init_active, compallsen,
                            st_start, st_dur, st_score, tokens, senscr_out, out, NULL);
}

/* init_topn: the scorer's carried top-N codewords ([mgau*feat][topn]) when the pass starts --
 * after a first pass on the same decoder they are what that pass left (orc_ptm_get_carried). */
int
orc_state_align2(const orc_model_t *m, int topn, const float *feat, int T, int n_phones,
                 const int32_t *ssid, const int32_t *tmat, const int32_t *sf, const int32_t *ef,
                 const uint32_t *init_active, int compallsen, int32_t *st_start, int32_t *st_dur,
                 int32_t *st_score, int32_t *tokens, int16_t *senscr_out, orc_align_out_t *out,
                 const uint8_t *init_topn)
{
    chain_t c;
    orc_ptm_t *p = orc_ptm_new(m, topn, 1);
    int t, i, j, E = m->n_emit, ns = n_phones * E, nw = (m->n_sen + 31) / 32;
    uint32_t *bits = calloc(nw, sizeof(uint32_t));
    uint8_t *list = malloc(m->n_sen + 64);
    int16_t *senscr = malloc(sizeof(int16_t) * m->n_sen);
    if (init_active)
        memcpy(bits, init_active, nw * sizeof(uint32_t));
    if (init_topn)
        orc_ptm_set_carried(p, init_topn);
    chain_init(&c, m, n_phones, ssid, tmat, sf, ef, T);
    for (t = 0; t < T; ++t) {
        int n_active = 0;
        /* activate HMMs entering this frame; the vector is never cleared (:186-188) */
        if (!compallsen) {
            for (i = 0; i < n_phones; ++i)
                if (c.h[i].frame == t)
                    for (j = 0; j < E; ++j) {
                        int s = m->sseq[(size_t)ssid[i] * E + j];
                        bits[s >> 5] |= 1u << (s & 31);
                    }
            n_active = orc_flags2list(bits, m->n_sen, list);
        }
        orc_ptm_frame_eval(p, senscr, list, n_active, feat + (size_t)t * m->blk, t, compallsen, NULL);
        p->frame_idx = t + 1;
        if (senscr_out)
            memcpy(senscr_out + (size_t)t * m->n_sen, senscr, sizeof(int16_t) * m->n_sen);
        chain_step(&c, senscr, t);
    }
    out->rv = chain_finish(&c, T, st_start, st_dur, st_score);
    out->best_score = c.best_score;
    out->n_renorm = c.n_renorm;
    if (tokens)
        memcpy(tokens, c.tokens, sizeof(int32_t) * 2 * (size_t)T * ns);
    free(c.h);
    free(c.tokens);
    free(bits);
    free(list);
    free(senscr);
    orc_ptm_free(p);
    return out->rv;
}

/* ref: ps_alignment.c:317-341 (states -> phones; phones -> words is the same fold) */
int
orc_propagate(int n_states, int n_emit, const int32_t *st_start, const int32_t *st_dur,
              const int32_t *st_score, int32_t *ph_start, int32_t *ph_dur, int32_t *ph_score)
{
    int i;
    for (i = 0; i < n_states; ++i) {
        int p = i / n_emit;
        if (i % n_emit == 0) {
            ph_start[p] = st_start[i];
            ph_dur[p] = 0;
            ph_score[p] = 0;
        }
        ph_dur[p] += st_dur[i];
        ph_score[p] += st_score[i];
    }
    return n_states / n_emit;
}
