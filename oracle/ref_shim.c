/*
 * ref_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A thin ctypes-friendly driver compiled INTO oracle/_ref/libssref.so together
 * with the unmodified reference sources (see oracle/Makefile).  It calls the
 * reference's own functions (decoder_*, acmod_*, search_module_*, the mgau
 * vtable) and copies results/internal arrays out through flat buffers so that
 * Python tests and tools/make_golden.py can
 *   (1) validate our restatement in oracle/ss_oracle.c,
 *   (2) generate the golden fixtures committed under tests/golden/,
 *   (3) serve as the "reference" CPU baseline in bench.py.
 *
 * Nothing here restates an algorithm: every number comes out of reference code.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <soundswallower.h>
#include <soundswallower/acmod.h>
#include <soundswallower/alignment.h>
#include <soundswallower/bin_mdef.h>
#include <soundswallower/ckd_alloc.h>
#include <soundswallower/dict.h>
#include <soundswallower/dict2pid.h>
#include <soundswallower/feat.h>
#include <soundswallower/fsg_search.h>
#include <soundswallower/hmm.h>
#include <soundswallower/ms_mgau.h>
#include <soundswallower/ptm_mgau.h>
#include <soundswallower/s2_semi_mgau.h>
#include <soundswallower/search_module.h>
#include <soundswallower/state_align_search.h>
#include <soundswallower/tied_mgau_common.h>
#include <soundswallower/tmat.h>

typedef struct ref_s {
    decoder_t *d;
} ref_t;

void *
ref_new(const char *hmmdir, const char *dictfile, int compallsen, int samprate,
        const char *loglevel)
{
    config_t *config = config_init(NULL);
    ref_t *r;
    config_set_str(config, "loglevel", loglevel ? loglevel : "FATAL");
    config_set_str(config, "hmm", hmmdir);
    if (dictfile && dictfile[0])
        config_set_str(config, "dict", dictfile);
    if (samprate > 0)
        config_set_int(config, "samprate", samprate);
    config_set_bool(config, "compallsen", compallsen);
    r = calloc(1, sizeof(*r));
    r->d = decoder_init(config);
    if (r->d == NULL) {
        free(r);
        return NULL;
    }
    return r;
}

void
ref_free(void *h)
{
    ref_t *r = h;
    if (r == NULL)
        return;
    decoder_free(r->d);
    free(r);
}

/* dims: n_mgau n_feat n_density veclen0 n_sen n_sseq n_emit n_tmat n_ciphone n_phone sil */
int
ref_model_dims(void *h, int32 *out)
{
    ref_t *r = h;
    ptm_mgau_t *s = (ptm_mgau_t *)r->d->acmod->mgau;
    bin_mdef_t *m = r->d->acmod->mdef;
    gauden_t *g;
    int n_sen;
    if (strcmp(r->d->acmod->mgau->vt->name, "ptm") == 0) {
        g = s->g;
        n_sen = s->n_sen;
    } else if (strcmp(r->d->acmod->mgau->vt->name, "s2_semi") == 0) {
        g = ((s2_semi_mgau_t *)r->d->acmod->mgau)->g;
        n_sen = ((s2_semi_mgau_t *)r->d->acmod->mgau)->n_sen;
    } else if (strcmp(r->d->acmod->mgau->vt->name, "ms") == 0) {
        g = ((ms_mgau_model_t *)r->d->acmod->mgau)->g;
        n_sen = ((ms_mgau_model_t *)r->d->acmod->mgau)->s->n_sen;
    } else
        return -1;
    out[0] = g->n_mgau;
    out[1] = g->n_feat;
    out[2] = g->n_density;
    out[3] = g->featlen[0];
    out[4] = n_sen;
    out[5] = m->n_sseq;
    out[6] = m->n_emit_state;
    out[7] = r->d->acmod->tmat->n_tmat;
    out[8] = m->n_ciphone;
    out[9] = m->n_phone;
    out[10] = m->sil;
    return 0;
}

/* Copy out the in-memory model arrays the reference actually scores with. */
int
ref_model_copy(void *h, float *mean, float *var, float *det, uint8 *mixw,
               uint8 *sen2cb, uint8 *tp, uint16 *sseq, uint8 *lut8)
{
    ref_t *r = h;
    bin_mdef_t *m = r->d->acmod->mdef;
    tmat_t *t = r->d->acmod->tmat;
    int i, j, k, c, f, n = 0, nd = 0;
    int semi = strcmp(r->d->acmod->mgau->vt->name, "s2_semi") == 0;
    int cont = strcmp(r->d->acmod->mgau->vt->name, "ms") == 0;
    gauden_t *g;
    uint8 ***mw, *mw_cb;
    logmath_t *lm8;
    int n_sen;
    if (cont) {
        /* ms_mgau: weights are senone->pdf[sen][feat][codeword] (n_gauden > 1) */
        ms_mgau_model_t *s = (ms_mgau_model_t *)r->d->acmod->mgau;
        if (s->s->n_gauden <= 1)
            return -3;
        g = s->g, mw = (uint8 ***)s->s->pdf, mw_cb = NULL, lm8 = s->s->lmath, n_sen = s->s->n_sen;
    } else if (semi) {
        s2_semi_mgau_t *s = (s2_semi_mgau_t *)r->d->acmod->mgau;
        g = s->g, mw = s->mixw, mw_cb = s->mixw_cb, lm8 = s->lmath_8b, n_sen = s->n_sen;
    } else {
        ptm_mgau_t *s = (ptm_mgau_t *)r->d->acmod->mgau;
        g = s->g, mw = s->mixw, mw_cb = s->mixw_cb, lm8 = s->lmath_8b, n_sen = s->n_sen;
    }
    for (c = 0; c < g->n_mgau; ++c)
        for (f = 0; f < g->n_feat; ++f)
            for (k = 0; k < g->n_density; ++k) {
                int L = g->featlen[f];
                memcpy(mean + n, g->mean[c][f][k], L * sizeof(float));
                memcpy(var + n, g->var[c][f][k], L * sizeof(float));
                n += L;
                det[nd++] = g->det[c][f][k];
            }
    for (f = 0; cont && f < n_sen; ++f) /* [sen][feat][codeword] as stored */
        for (k = 0; k < g->n_feat; ++k)
            memcpy(mixw + ((size_t)f * g->n_feat + k) * g->n_density, mw[f][k], g->n_density);
    for (f = 0; !cont && f < g->n_feat; ++f)
        for (k = 0; k < g->n_density; ++k) {
            uint8 *dst = mixw + ((size_t)f * g->n_density + k) * n_sen;
            if (!mw_cb) {
                memcpy(dst, mw[f][k], n_sen);
                continue;
            }
            /* 4-bit clustered sendump: expand with the expression the scoring loops use.
             * ptm_mgau.c:375-378 chooses the nibble by the low bit of the packed byte
             * itself; s2_semi_mgau.c:733-757 by the parity of the senone. */
            for (i = 0; i < n_sen; ++i) {
                int dcw = mw[f][k][i / 2];
                if (semi)
                    dcw = (i & 1) ? dcw >> 4 : dcw & 0x0f;
                else
                    dcw = (dcw & 1) ? dcw >> 4 : dcw & 0x0f;
                dst[i] = mw_cb[dcw];
            }
        }
    if (semi || cont)
        memset(sen2cb, 0, n_sen);
    else
        memcpy(sen2cb, ((ptm_mgau_t *)r->d->acmod->mgau)->sen2cb, n_sen);
    for (i = 0; i < t->n_tmat; ++i)
        for (j = 0; j < t->n_state; ++j)
            for (k = 0; k < t->n_state + 1; ++k)
                tp[(i * t->n_state + j) * (t->n_state + 1) + k] = t->tp[i][j][k];
    for (i = 0; i < m->n_sseq; ++i)
        for (j = 0; j < m->n_emit_state; ++j)
            sseq[i * m->n_emit_state + j] = m->sseq[i][j];
    {
        /* 8-bit log-add table of the scorer (logmath_init(base, 10, TRUE)) */
        int d;
        for (d = 0; d < 256; ++d)
            lut8[d] = (uint8)(0 - fast_logmath_add(lm8, 0, d));
    }
    return 0;
}

/* phone table: ssid, tmat, ci for each of n_phone entries */
int
ref_phone_table(void *h, int32 *ssid, int32 *tmat, int32 *ci)
{
    ref_t *r = h;
    bin_mdef_t *m = r->d->acmod->mdef;
    int i;
    for (i = 0; i < m->n_phone; ++i) {
        ssid[i] = m->phone[i].ssid;
        tmat[i] = m->phone[i].tmat;
        ci[i] = bin_mdef_pid2ci(m, i);
    }
    return m->n_phone;
}

/* Run the reference frontend on a whole utterance (full_utt=1) and copy the
 * dynamic features out: out[T][n_feat*veclen]. Returns T. */
int
ref_features_from_pcm(void *h, const int16 *pcm, long nsamp, float *out,
                      int max_frames)
{
    ref_t *r = h;
    acmod_t *a = r->d->acmod;
    int nfr, t, f, D = 0, nf;
    /* Feed acmod directly so that no search consumes the frames (and no
     * grammar needs to be set). */
    acmod_start_utt(a);
    {
        const int16 *p = pcm;
        size_t n = nsamp;
        nfr = acmod_process_raw(a, &p, &n, TRUE);
    }
    acmod_end_utt(a);
    nf = feat_dimension1(a->fcb);
    for (f = 0; f < nf; ++f)
        D += feat_dimension2(a->fcb, f);
    nfr = a->n_feat_frame;
    if (out) {
        if (nfr > max_frames)
            return -2;
        for (t = 0; t < nfr; ++t)
            memcpy(out + (size_t)t * D, a->feat_buf[t][0], D * sizeof(float));
    }
    return nfr;
}

/* Also expose 13-dim MFCCs (before CMN / deltas) for frontend parity. */
int
ref_mfcc_from_pcm(void *h, const int16 *pcm, long nsamp, float *out, int max_frames)
{
    ref_t *r = h;
    fe_t *fe = r->d->acmod->fe;
    const int16 *p = pcm;
    size_t n = nsamp;
    int nfr, nvec, ncep = fe_get_output_size(fe), t;
    mfcc_t **buf;
    nfr = fe_process_int16(fe, NULL, &n, NULL, 0);
    if (nfr > max_frames)
        return -2;
    buf = (mfcc_t **)ckd_calloc_2d(nfr, ncep, sizeof(mfcc_t));
    fe_start(fe);
    p = pcm;
    n = nsamp;
    nvec = fe_process_int16(fe, &p, &n, buf, nfr);
    nvec += fe_end(fe, buf + nvec, nfr - nvec);
    for (t = 0; t < nvec; ++t)
        memcpy(out + (size_t)t * ncep, buf[t], ncep * sizeof(float));
    ckd_free_2d(buf);
    return nvec;
}

/* Same through the float32 entry point (samples in [-1, 1)). */
int
ref_mfcc_from_f32(void *h, const float *pcm, long nsamp, float *out, int max_frames)
{
    ref_t *r = h;
    fe_t *fe = r->d->acmod->fe;
    const float *p = pcm;
    size_t n = nsamp;
    int nfr, nvec, ncep = fe_get_output_size(fe), t;
    mfcc_t **buf;
    nfr = fe_process_float32(fe, NULL, &n, NULL, 0);
    if (nfr > max_frames)
        return -2;
    buf = (mfcc_t **)ckd_calloc_2d(nfr > 0 ? nfr : 1, ncep, sizeof(mfcc_t));
    fe_start(fe);
    p = pcm;
    n = nsamp;
    nvec = fe_process_float32(fe, (float32 **)&p, &n, buf, nfr);
    nvec += fe_end(fe, buf + nvec, nfr - nvec);
    for (t = 0; t < nvec; ++t)
        memcpy(out + (size_t)t * ncep, buf[t], ncep * sizeof(float));
    ckd_free_2d(buf);
    return nvec;
}

/* Load T frames of precomputed features into the acmod as a full utterance. */
static int
load_features(ref_t *r, const float *feat, int T)
{
    acmod_t *a = r->d->acmod;
    int t, f, D = 0, nf = feat_dimension1(a->fcb);
    for (f = 0; f < nf; ++f)
        D += feat_dimension2(a->fcb, f);
    acmod_start_utt(a);
    if (a->n_feat_alloc < T) {
        feat_array_free(a->feat_buf);
        a->feat_buf = feat_array_alloc(a->fcb, T);
        a->framepos = ckd_realloc(a->framepos, T * sizeof(*a->framepos));
        a->n_feat_alloc = T;
    }
    for (t = 0; t < T; ++t)
        memcpy(a->feat_buf[t][0], feat + (size_t)t * D, D * sizeof(float));
    a->n_feat_frame = T;
    a->feat_outidx = 0;
    a->output_frame = 0;
    a->state = ACMOD_ENDED;
    return 0;
}

/* Score every frame the way acmod_score does (in the decoder's compallsen
 * mode, with whatever is in the active vector when compallsen=0). out is
 * [T][n_sen] int16. Resets the PTM top-N history first if reset_hist. */
int
ref_score_all(void *h, const float *feat, int T, int16 *out, int reset_hist)
{
    ref_t *r = h;
    acmod_t *a = r->d->acmod;
    int t, nsen = bin_mdef_n_sen(a->mdef);
    if (reset_hist) {
        ptm_mgau_t *s = (ptm_mgau_t *)a->mgau;
        int i, j, k, m;
        for (i = 0; i < s->n_fast_hist; ++i)
            for (j = 0; j < s->g->n_mgau; ++j)
                for (k = 0; k < s->g->n_feat; ++k)
                    for (m = 0; m < s->max_topn; ++m) {
                        s->hist[i].topn[j][k][m].cw = m;
                        s->hist[i].topn[j][k][m].score = WORST_DIST;
                    }
    }
    load_features(r, feat, T);
    for (t = 0; t < T; ++t) {
        int fi = t;
        int16 const *sc = acmod_score(a, &fi);
        if (sc == NULL)
            return -1;
        memcpy(out + (size_t)t * nsen, sc, nsen * sizeof(int16));
        acmod_advance(a);
    }
    return T;
}

/* Reset PTM history to the freshly-initialised state. */
void
ref_reset_hist(void *h)
{
    ref_t *r = h;
    ptm_mgau_t *s = (ptm_mgau_t *)r->d->acmod->mgau;
    int i, j, k, m;
    for (i = 0; i < s->n_fast_hist; ++i)
        for (j = 0; j < s->g->n_mgau; ++j)
            for (k = 0; k < s->g->n_feat; ++k)
                for (m = 0; m < s->max_topn; ++m) {
                    s->hist[i].topn[j][k][m].cw = m;
                    s->hist[i].topn[j][k][m].score = WORST_DIST;
                }
}

/* Direct vtable call for one frame: frame_eval(mgau, senscr, active list, ...)
 * feat = n_feat*veclen floats.  Also returns the post-norm top-N (cw, score)
 * for every (cb, feat) in topn_out[n_mgau][n_feat][topn][2] if non-NULL. */
int
ref_frame_eval(void *h, const float *feat, int frame, const uint8 *active,
               int n_active, int compallsen, int16 *out, int32 *topn_out)
{
    ref_t *r = h;
    acmod_t *a = r->d->acmod;
    ptm_mgau_t *s = (ptm_mgau_t *)a->mgau;
    mfcc_t *ptrs[8];
    float buf[256];
    int f, off = 0, nf = s->g->n_feat, i, j, k;
    for (f = 0; f < nf; ++f) {
        ptrs[f] = buf + off;
        off += s->g->featlen[f];
    }
    memcpy(buf, feat, off * sizeof(float));
    a->mgau->frame_idx = frame; /* force recompute of this frame */
    a->mgau->vt->frame_eval(a->mgau, out, (uint8 *)active, n_active, ptrs, frame,
                            compallsen);
    if (topn_out) {
        for (i = 0; i < s->g->n_mgau; ++i)
            for (j = 0; j < nf; ++j)
                for (k = 0; k < s->max_topn; ++k) {
                    *topn_out++ = s->f->topn[i][j][k].cw;
                    *topn_out++ = s->f->topn[i][j][k].score;
                }
    }
    return 0;
}

static int
copy_alignment(alignment_t *al, dict_t *dict, int32 *n_out, int32 *words,
               int32 *phones, int32 *states, int maxw, int maxp, int maxs)
{
    int i;
    n_out[0] = al->word.n_ent;
    n_out[1] = al->sseq.n_ent;
    n_out[2] = al->state.n_ent;
    if (al->word.n_ent > maxw || al->sseq.n_ent > maxp || al->state.n_ent > maxs)
        return -2;
    (void)dict;
    for (i = 0; i < al->word.n_ent; ++i) {
        alignment_entry_t *e = al->word.seq + i;
        words[i * 4 + 0] = e->id.wid;
        words[i * 4 + 1] = e->start;
        words[i * 4 + 2] = e->duration;
        words[i * 4 + 3] = e->score;
    }
    for (i = 0; i < al->sseq.n_ent; ++i) {
        alignment_entry_t *e = al->sseq.seq + i;
        phones[i * 7 + 0] = e->id.pid.cipid;
        phones[i * 7 + 1] = e->id.pid.ssid;
        phones[i * 7 + 2] = e->id.pid.tmatid;
        phones[i * 7 + 3] = e->start;
        phones[i * 7 + 4] = e->duration;
        phones[i * 7 + 5] = e->score;
        phones[i * 7 + 6] = e->parent;
    }
    for (i = 0; i < al->state.n_ent; ++i) {
        alignment_entry_t *e = al->state.seq + i;
        states[i * 5 + 0] = e->id.senid;
        states[i * 5 + 1] = e->start;
        states[i * 5 + 2] = e->duration;
        states[i * 5 + 3] = e->score;
        states[i * 5 + 4] = e->parent;
    }
    return 0;
}

/* Two-pass alignment of raw PCM against text, exactly as the CLI does:
 * set_align_text, start_utt, process_int16(full_utt), end_utt, alignment.
 * segs: pass-1 word segments [n][5] = wid sf ef ascr lscr ; n_out[3] = n_segs,
 * n_out[4] = hyp score, n_out[5] = n_frames. */
int
ref_align_pcm(void *h, const int16 *pcm, long nsamp, const char *text,
              int32 *n_out, int32 *segs, int maxseg, int32 *words, int32 *phones,
              int32 *states, int maxw, int maxp, int maxs)
{
    ref_t *r = h;
    decoder_t *d = r->d;
    seg_iter_t *seg;
    alignment_t *al;
    int32 score = 0;
    int n = 0;
    if (decoder_set_align_text(d, text) < 0)
        return -1;
    if (decoder_start_utt(d) < 0)
        return -1;
    if (decoder_process_int16(d, (int16 *)pcm, nsamp, FALSE, TRUE) < 0)
        return -1;
    if (decoder_end_utt(d) < 0)
        return -1;
    decoder_hyp(d, &score);
    n_out[4] = score;
    n_out[5] = decoder_n_frames(d);
    for (seg = decoder_seg_iter(d); seg; seg = seg_iter_next(seg)) {
        int sf, ef;
        int32 ascr, lscr;
        if (n >= maxseg)
            return -2;
        seg_iter_frames(seg, &sf, &ef);
        seg_iter_prob(seg, &ascr, &lscr);
        segs[n * 5 + 0] = dict_wordid(d->dict, seg_iter_word(seg));
        segs[n * 5 + 1] = sf;
        segs[n * 5 + 2] = ef;
        segs[n * 5 + 3] = ascr;
        segs[n * 5 + 4] = lscr;
        ++n;
    }
    n_out[3] = n;
    al = decoder_alignment(d);
    if (al == NULL)
        return -3;
    return copy_alignment(al, d->dict, n_out, words, phones, states, maxw, maxp, maxs);
}

/* Word id lookup / string. */
int
ref_wordid(void *h, const char *w)
{
    ref_t *r = h;
    return dict_wordid(r->d->dict, w);
}
const char *
ref_wordstr(void *h, int wid)
{
    ref_t *r = h;
    return dict_wordstr(r->d->dict, wid);
}

/* Expand a word sequence (with windows) into the phone/state chain the
 * reference would align (alignment_add_word + alignment_populate). */
int
ref_populate(void *h, const int32 *wids, const int32 *start, const int32 *dur,
             int nw, int32 *n_out, int32 *words, int32 *phones, int32 *states,
             int maxw, int maxp, int maxs)
{
    ref_t *r = h;
    alignment_t *al = alignment_init(r->d->d2p);
    int i, rv;
    for (i = 0; i < nw; ++i)
        alignment_add_word(al, wids[i], start[i], dur[i]);
    if (alignment_populate(al) < 0) {
        alignment_free(al);
        return -1;
    }
    rv = copy_alignment(al, r->d->dict, n_out, words, phones, states, maxw, maxp, maxs);
    alignment_free(al);
    return rv;
}

/* Pass-2 only: state_align_search over given features and a word sequence with
 * windows (start/dur; 0/0 = unconstrained).  The acmod active vector is cleared
 * first (clear_active=1) or left as is.  If tokens_out != NULL it receives the
 * reference token stack [T][n_states][2] (id, score). */
int
ref_state_align(void *h, const float *feat, int T, const int32 *wids,
                const int32 *start, const int32 *dur, int nw, int clear_active,
                int32 *n_out, int32 *words, int32 *phones, int32 *states, int maxw,
                int maxp, int maxs, int32 *tokens_out, int16 *senscr_out)
{
    ref_t *r = h;
    decoder_t *d = r->d;
    acmod_t *a = d->acmod;
    alignment_t *al = alignment_init(d->d2p);
    search_module_t *sm;
    state_align_search_t *sas;
    int i, rv, nsen = bin_mdef_n_sen(a->mdef);
    for (i = 0; i < nw; ++i)
        alignment_add_word(al, wids[i], start[i], dur[i]);
    if (alignment_populate(al) < 0) {
        alignment_free(al);
        return -1;
    }
    load_features(r, feat, T);
    if (clear_active)
        acmod_clear_active(a);
    alignment_retain(al);
    sm = state_align_search_init("_sa", d->config, a, al);
    if (sm == NULL)
        return -1;
    sas = (state_align_search_t *)sm;
    if (search_module_start(sm) < 0)
        return -1;
    while (a->output_frame < T) {
        if (search_module_step(sm, a->output_frame) < 0)
            return -1;
        if (senscr_out)
            memcpy(senscr_out + (size_t)a->output_frame * nsen, a->senone_scores,
                   nsen * sizeof(int16));
        acmod_advance(a);
    }
    n_out[6] = sas->best_score;
    rv = search_module_finish(sm);
    n_out[7] = rv;
    if (tokens_out)
        memcpy(tokens_out, sas->tokens,
               (size_t)T * sas->n_emit_state * sizeof(*sas->tokens));
    copy_alignment(al, d->dict, n_out, words, phones, states, maxw, maxp, maxs);
    search_module_free(sm);
    alignment_free(al);
    return rv;
}

/* FSG decode of features with a JSGF grammar string or align text.  Returns
 * number of segs; segs [n][5] = wid sf ef ascr lscr; n_out[4]=score. */
int
ref_fsg_decode(void *h, const float *feat, int T, const char *align_text,
               const char *jsgf_string, int32 *n_out, int32 *segs, int maxseg)
{
    ref_t *r = h;
    decoder_t *d = r->d;
    acmod_t *a = d->acmod;
    seg_iter_t *seg;
    int32 score = 0;
    int n = 0;
    if (align_text) {
        if (decoder_set_align_text(d, align_text) < 0)
            return -1;
    } else if (jsgf_string) {
        if (decoder_set_jsgf_string(d, jsgf_string) < 0)
            return -1;
    }
    if (decoder_start_utt(d) < 0)
        return -1;
    load_features(r, feat, T);
    a->state = ACMOD_ENDED;
    while (a->n_feat_frame > 0) {
        if (search_module_step(d->search, a->output_frame) < 0)
            return -1;
        acmod_advance(a);
    }
    search_module_finish(d->search);
    decoder_hyp(d, &score);
    n_out[4] = score;
    n_out[5] = a->output_frame;
    for (seg = decoder_seg_iter(d); seg; seg = seg_iter_next(seg)) {
        int sf, ef;
        int32 ascr, lscr;
        if (n >= maxseg)
            return -2;
        seg_iter_frames(seg, &sf, &ef);
        seg_iter_prob(seg, &ascr, &lscr);
        segs[n * 5 + 0] = seg_iter_word(seg) ? dict_wordid(d->dict, seg_iter_word(seg)) : -1;
        segs[n * 5 + 1] = sf;
        segs[n * 5 + 2] = ef;
        segs[n * 5 + 3] = ascr;
        segs[n * 5 + 4] = lscr;
        ++n;
    }
    n_out[3] = n;
    return n;
}

/* The same decode, asked for its hypothesis BETWEEN steps: after each of stops[0..n_stops)
 * frames (ascending) decoder_hyp / decoder_seg_iter are called before the next step.  out
 * [n_stops][2 + 5*maxseg] = hyp score (INT32_MIN when decoder_hyp returns NULL), n segs, then
 * segs as in ref_fsg_decode; hyps [n_stops][hyp_len] the hypothesis strings.  The utterance is
 * then run to its end and finished. */
int
ref_fsg_partial(void *h, const float *feat, int T, const char *align_text, const int32 *stops,
                int n_stops, int32 *out, int maxseg, char *hyps, int hyp_len)
{
    ref_t *r = h;
    decoder_t *d = r->d;
    acmod_t *a = d->acmod;
    int k = 0, stride = 2 + 5 * maxseg;
    if (align_text && decoder_set_align_text(d, align_text) < 0)
        return -1;
    if (decoder_start_utt(d) < 0)
        return -1;
    load_features(r, feat, T);
    a->state = ACMOD_ENDED;
    while (a->n_feat_frame > 0) {
        if (search_module_step(d->search, a->output_frame) < 0)
            return -1;
        acmod_advance(a);
        while (k < n_stops && stops[k] == a->output_frame) {
            int32 score = 0, *o = out + (size_t)k * stride;
            const char *hyp = decoder_hyp(d, &score);
            seg_iter_t *seg;
            int n = 0;
            o[0] = hyp ? score : INT32_MIN;
            hyps[(size_t)k * hyp_len] = 0;
            if (hyp)
                strncpy(hyps + (size_t)k * hyp_len, hyp, hyp_len - 1), hyps[(size_t)k * hyp_len + hyp_len - 1] = 0;
            for (seg = decoder_seg_iter(d); seg; seg = seg_iter_next(seg)) {
                int sf, ef;
                int32 ascr, lscr;
                if (n >= maxseg) {
                    seg_iter_free(seg);
                    break;
                }
                seg_iter_frames(seg, &sf, &ef);
                seg_iter_prob(seg, &ascr, &lscr);
                o[2 + n * 5 + 0] = seg_iter_word(seg) ? dict_wordid(d->dict, seg_iter_word(seg)) : -1;
                o[2 + n * 5 + 1] = sf;
                o[2 + n * 5 + 2] = ef;
                o[2 + n * 5 + 3] = ascr;
                o[2 + n * 5 + 4] = lscr;
                ++n;
            }
            o[1] = n;
            ++k;
        }
    }
    search_module_finish(d->search);
    return k;
}

/* acmod's active-senone flags as they stand (after an fsg decode in the default mode: the
 * senones of the last frame's active HMMs); out = (n_sen+31)/32 words.  Also the number of
 * senones the last grammar search evaluated (fsgs->n_sen_eval) when it is an fsg search. */
int
ref_active_bits(void *h, uint32 *out, int32 *n_sen_eval)
{
    ref_t *r = h;
    acmod_t *a = r->d->acmod;
    int n_sen = bin_mdef_n_sen(a->mdef), i;
    for (i = 0; i < (n_sen + 31) / 32; ++i)
        out[i] = a->senone_active_vec[i];
    if (n_sen_eval && r->d->search
        && 0 == strcmp(search_module_type(r->d->search), PS_SEARCH_TYPE_FSG))
        *n_sen_eval = ((fsg_search_t *)r->d->search)->n_sen_eval;
    return 0;
}

/* One hmm_vit_eval on caller-provided state (3- or 5-state, non-mpx).
 * st: score[5] history[5] out_score out_history (12 int32), in/out. */
int
ref_hmm_vit_eval(void *h, int n_emit, int tmatid, const uint16 *senid,
                 const int16 *senscr, int32 *st)
{
    ref_t *r = h;
    acmod_t *a = r->d->acmod;
    hmm_context_t *ctx;
    hmm_t hmm;
    uint16 *sseq_row = (uint16 *)senid;
    uint16 *const sseq[1] = { sseq_row };
    int32 best;
    int i;
    ctx = hmm_context_init(n_emit, a->tmat->tp, senscr, sseq);
    hmm_init(ctx, &hmm, FALSE, 0, tmatid);
    for (i = 0; i < 5; ++i) {
        hmm.score[i] = st[i];
        hmm.history[i] = st[5 + i];
    }
    hmm.out_score = st[10];
    hmm.out_history = st[11];
    best = hmm_vit_eval(&hmm);
    for (i = 0; i < 5; ++i) {
        st[i] = hmm.score[i];
        st[5 + i] = hmm.history[i];
    }
    st[10] = hmm.out_score;
    st[11] = hmm.out_history;
    hmm_context_free(ctx);
    return best;
}

/* The same on a caller-provided transition matrix tp[n_emit][n_emit+1] (255 = impossible) and
 * the n_emit senone scores of the HMM's states: hmm_vit_eval_5st_lr has no bundled model. */
int
ref_hmm_vit_eval_tp(void *h, int n_emit, const uint8 *tp, const int16 *senscr, int32 *st)
{
    uint8 *rows[8];
    uint8 **mat[1];
    uint16 sseq_row[8];
    uint16 *const sseq[1] = { sseq_row };
    hmm_context_t *ctx;
    hmm_t hmm;
    int32 best;
    int i;
    (void)h;
    if (n_emit < 1 || n_emit > 5)
        return 0x7fffffff;
    for (i = 0; i < n_emit; ++i) {
        rows[i] = (uint8 *)tp + i * (n_emit + 1);
        sseq_row[i] = (uint16)i;
    }
    mat[0] = rows;
    ctx = hmm_context_init(n_emit, mat, senscr, sseq);
    hmm_init(ctx, &hmm, FALSE, 0, 0);
    for (i = 0; i < 5; ++i) {
        hmm.score[i] = st[i];
        hmm.history[i] = st[5 + i];
    }
    hmm.out_score = st[10];
    hmm.out_history = st[11];
    best = hmm_vit_eval(&hmm);
    for (i = 0; i < 5; ++i) {
        st[i] = hmm.score[i];
        st[5 + i] = hmm.history[i];
    }
    st[10] = hmm.out_score;
    st[11] = hmm.out_history;
    hmm_context_free(ctx);
    return best;
}

/* ------------------------------------------------------------------ */
/* FSG search: flattened graph + history table (for the K4 oracle/kernel) */
/* ------------------------------------------------------------------ */
#include <soundswallower/fsg_history.h>
#include <soundswallower/fsg_lextree.h>

typedef struct {
    fsg_link_t **link;
    int n_link;
    fsg_pnode_t **pnode;
    int n_pnode;
} fsg_index_t;

static fsg_search_t *
fsg_of(ref_t *r)
{
    search_module_t *s = r->d->search;
    if (s == NULL || strcmp(s->type, PS_SEARCH_TYPE_FSG) != 0)
        return NULL;
    return (fsg_search_t *)s;
}

static void
fsg_index_build(fsg_search_t *fs, fsg_index_t *ix)
{
    fsg_model_t *fsg = fs->fsg;
    int s, n;
    memset(ix, 0, sizeof(*ix));
    for (n = 0, s = 0; s < fsg_model_n_state(fsg); ++s) {
        fsg_arciter_t *it;
        for (it = fsg_model_arcs(fsg, s); it; it = fsg_arciter_next(it))
            ++n;
    }
    ix->link = ckd_calloc(n + 1, sizeof(*ix->link));
    for (s = 0; s < fsg_model_n_state(fsg); ++s) {
        fsg_arciter_t *it;
        for (it = fsg_model_arcs(fsg, s); it; it = fsg_arciter_next(it))
            ix->link[ix->n_link++] = fsg_arciter_get(it);
    }
    ix->pnode = ckd_calloc(fs->lextree->n_pnode + 1, sizeof(*ix->pnode));
    for (s = 0; s < fsg_model_n_state(fsg); ++s) {
        fsg_pnode_t *p;
        for (p = fs->lextree->alloc_head[s]; p; p = p->alloc_next)
            ix->pnode[ix->n_pnode++] = p;
    }
}

static void
fsg_index_free(fsg_index_t *ix)
{
    ckd_free(ix->link);
    ckd_free(ix->pnode);
}

static int
link_id(fsg_index_t *ix, fsg_link_t *l)
{
    int i;
    if (l == NULL)
        return -1;
    for (i = 0; i < ix->n_link; ++i)
        if (ix->link[i] == l)
            return i;
    return -2;
}

static int
pnode_id(fsg_index_t *ix, fsg_pnode_t *p)
{
    int i;
    if (p == NULL)
        return -1;
    for (i = 0; i < ix->n_pnode; ++i)
        if (ix->pnode[i] == p)
            return i;
    return -2;
}

/* Select the grammar (align text or JSGF string) without decoding. */
int
ref_fsg_prepare(void *h, const char *align_text, const char *jsgf_string)
{
    ref_t *r = h;
    if (align_text)
        return decoder_set_align_text(r->d, align_text);
    if (jsgf_string)
        return decoder_set_jsgf_string(r->d, jsgf_string);
    return -1;
}

/* Select the grammar of a text .fsg file (fsg_model_readfile + decoder_set_fsg). */
int
ref_fsg_prepare_file(void *h, const char *path)
{
    ref_t *r = h;
    fsg_model_t *fsg = fsg_model_readfile(path, decoder_logmath(r->d),
                                          config_float(decoder_config(r->d), "lw"));
    if (fsg == NULL)
        return -1;
    return decoder_set_fsg(r->d, fsg);   /* consumes fsg */
}

/* out: n_state start final n_link n_pnode n_ciphone silcipid beam pbeam wbeam maxhmmpf wip pip */
int
ref_fsg_dims(void *h, int32 *out)
{
    ref_t *r = h;
    fsg_search_t *fs = fsg_of(r);
    fsg_index_t ix;
    if (fs == NULL)
        return -1;
    fsg_index_build(fs, &ix);
    out[0] = fsg_model_n_state(fs->fsg);
    out[1] = fsg_model_start_state(fs->fsg);
    out[2] = fsg_model_final_state(fs->fsg);
    out[3] = ix.n_link;
    out[4] = ix.n_pnode;
    out[5] = bin_mdef_n_ciphone(r->d->acmod->mdef);
    out[6] = bin_mdef_ciphone_id(r->d->acmod->mdef, "SIL");
    out[7] = fs->beam_orig;
    out[8] = fs->pbeam_orig;
    out[9] = fs->wbeam_orig;
    out[10] = config_int(search_module_config(fs), "maxhmmpf");
    out[11] = fs->wip;
    out[12] = fs->pip;
    fsg_index_free(&ix);
    return 0;
}

/* link4 [n_link][4] = from to logs2prob wid ; link_flag [n_link] bit0 = the word models no
 * right context (filler or single-phone word, the test of fsg_search_pnode_exit), bit1 = filler;
 * arc_off [n_state+1] into the link list (links are numbered in fsg_model_arcs order, state by
 * state, so state s owns links arc_off[s]..arc_off[s+1]);
 * root [n_state]; pnode8 [n_pnode][8] = ssid tmat logs2prob ci_ext leaf succ_or_link sibling ppos;
 * ctxt [n_pnode][4]. */
int
ref_fsg_dump(void *h, int32 *link4, uint8 *link_flag, int32 *arc_off, int32 *root, int32 *pnode8,
             uint32 *ctxt)
{
    ref_t *r = h;
    fsg_search_t *fs = fsg_of(r);
    fsg_model_t *fsg;
    dict_t *dict = r->d->dict;
    fsg_index_t ix;
    int s, i, n;
    if (fs == NULL)
        return -1;
    fsg = fs->fsg;
    fsg_index_build(fs, &ix);
    for (n = 0, s = 0; s < fsg_model_n_state(fsg); ++s) {
        fsg_arciter_t *it;
        arc_off[s] = n;
        for (it = fsg_model_arcs(fsg, s); it; it = fsg_arciter_next(it))
            ++n;
        root[s] = pnode_id(&ix, fsg_lextree_root(fs->lextree, s));
    }
    arc_off[fsg_model_n_state(fsg)] = n;
    for (i = 0; i < ix.n_link; ++i) {
        fsg_link_t *l = ix.link[i];
        int wid = fsg_link_wid(l);
        link4[i * 4 + 0] = fsg_link_from_state(l);
        link4[i * 4 + 1] = fsg_link_to_state(l);
        link4[i * 4 + 2] = fsg_link_logs2prob(l);
        link4[i * 4 + 3] = wid;
        link_flag[i] = 0;
        if (wid >= 0) {
            int filler = fsg_model_is_filler(fsg, wid);
            int single = dict_is_single_phone(dict, dict_wordid(dict, fsg_model_word_str(fsg, wid)));
            link_flag[i] = (uint8)((filler || single) ? 1 : 0) | (uint8)(filler ? 2 : 0);
        }
    }
    for (i = 0; i < ix.n_pnode; ++i) {
        fsg_pnode_t *p = ix.pnode[i];
        pnode8[i * 8 + 0] = p->hmm.ssid;
        pnode8[i * 8 + 1] = p->hmm.tmatid;
        pnode8[i * 8 + 2] = p->logs2prob;
        pnode8[i * 8 + 3] = p->ci_ext;
        pnode8[i * 8 + 4] = p->leaf;
        pnode8[i * 8 + 5] = p->leaf ? link_id(&ix, p->next.fsglink) : pnode_id(&ix, p->next.succ);
        pnode8[i * 8 + 6] = pnode_id(&ix, p->sibling);
        pnode8[i * 8 + 7] = p->ppos;
        memcpy(ctxt + i * 4, p->ctxt.bv, 4 * sizeof(uint32));
    }
    fsg_index_free(&ix);
    return 0;
}

/* History table of the last decode: ent [n][9] = link score pred frame lc rc[4]. */
int
ref_fsg_history(void *h, int32 *ent, int max_ent, int32 *n_out)
{
    ref_t *r = h;
    fsg_search_t *fs = fsg_of(r);
    fsg_index_t ix;
    int i, n;
    int32 score = 0;
    if (fs == NULL)
        return -1;
    n = fsg_history_n_entries(fs->history);
    if (n > max_ent)
        return -2;
    fsg_index_build(fs, &ix);
    for (i = 0; i < n; ++i) {
        fsg_hist_entry_t *e = fsg_history_entry_get(fs->history, i);
        ent[i * 9 + 0] = link_id(&ix, e->fsglink);
        ent[i * 9 + 1] = e->score;
        ent[i * 9 + 2] = e->pred;
        ent[i * 9 + 3] = e->frame;
        ent[i * 9 + 4] = e->lc;
        memcpy(ent + i * 9 + 5, e->rc.bv, 4 * sizeof(int32));
    }
    fsg_index_free(&ix);
    n_out[0] = n;
    n_out[1] = fs->n_hmm_eval;
    n_out[2] = fs->frame;
    search_module_hyp(r->d->search, &score);
    n_out[3] = score;
    return n;
}
