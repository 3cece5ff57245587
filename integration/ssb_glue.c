/* ssb_glue.c -- the reference-side binding of libssb200.so: the few functions a SoundSwallower
 * maintainer adds to the reference tree so that its own acmod.c / decoder.c run on the B200
 * library.  Compiled against the REFERENCE's headers (it is the only file of
 * this repository that includes them) and linked with -lssb200; it contains no algorithm.
 *
 *   ssb_ptm_mgau_init(acmod_t *)      same contract as ptm_mgau_init (ref: src/ptm_mgau.c:722-
 *                                     ; src/acmod.c:101-119): a scorer object whose first member
 *                                     is mgau_t, or NULL to decline (the caller falls through to
 *                                     the next scorer).
 *
 * oracle/Makefile target ref_ssb builds the unmodified reference sources with
 * -Dptm_mgau_init=ssb_ptm_mgau_init on acmod.c only, i.e. with the one-line change of
 * INTEGRATION.md section 1 applied at compile time, into oracle/_ref/libssref_ssb.so;
 * tests/test_link_dropin.py runs the reference's decoder through it on the GPU. */
#include <string.h>

#include <soundswallower/acmod.h>
#include <soundswallower/configuration.h>
#include <soundswallower/err.h>

#include "ssb200.h"

static int g_n_init = 0;
/* the model the most recent scorers were made from, with the serial number of their creation: a
 * freed model's address can come back for another model, so the searches' lexicon cache is keyed
 * by (address, serial) */
#define GLUE_N_MODELS 64
static struct { const void *model; int serial; } g_models[GLUE_N_MODELS];

static void
glue_note_model(const void *sm, int serial)
{
    int i, at = serial % GLUE_N_MODELS;
    for (i = 0; i < GLUE_N_MODELS; ++i)
        if (g_models[i].model == sm)
            at = i;
    g_models[at].model = sm;
    g_models[at].serial = serial;
}

static int
glue_model_serial(const void *sm)
{
    int i;
    for (i = 0; i < GLUE_N_MODELS; ++i)
        if (g_models[i].model == sm)
            return g_models[i].serial;
    return -1;
}

/* how many scorers this glue has handed to acmod (the test checks the B200 path really ran) */
int
ssb_glue_n_mgau_init(void)
{
    return g_n_init;
}

mgau_t *
ssb_ptm_mgau_init(acmod_t *acmod)
{
    ssb_config_t sc;
    ssb_model_t *sm;
    ssb_mgau_t *g;
    const char *hmm = config_str(acmod->config, "hmm");

    if (hmm == NULL)
        return NULL;
    ssb_config_defaults(&sc);
    sc.logbase = config_float(acmod->config, "logbase");
    sc.varfloor = config_float(acmod->config, "varfloor");
    sc.mixwfloor = config_float(acmod->config, "mixwfloor");
    sc.tmatfloor = config_float(acmod->config, "tmatfloor");
    sc.topn = config_int(acmod->config, "topn");
    sc.ds = config_int(acmod->config, "ds");
    sm = ssb_model_load(hmm, &sc);
    if (sm == NULL) {
        E_INFO("B200 scorer declines: %s\n", ssb_last_error());
        return NULL;
    }
    g = ssb_mgau_init(sm);
    if (g == NULL) {
        E_INFO("B200 scorer declines: %s\n", ssb_last_error());
        ssb_model_free(sm);
        return NULL;
    }
    ++g_n_init;
    glue_note_model(sm, g_n_init);
    /* (the model lives as long as the scorer; ssb_mgau_free releases both when the scorer
     * owns it: see ssb_mgau_own_model) */
    ssb_mgau_own_model(g, 1);
    return (mgau_t *)g;
}

/* ==================================================================================================
 * Searches.  decoder.c is compiled with
 *     -Dfsg_search_init=ssb_glue_fsg_search_init -Dstate_align_search_init=ssb_glue_state_align_search_init
 * (oracle/Makefile target ref_ssb), i.e. decoder_set_fsg / decoder_alignment (ref: src/decoder.c:
 * 600, 777) create the B200 library's search objects.  What they get back is a thin adapter with
 * the reference's own struct in front (decoder_alignment reads state_align_search_t.al / .frame
 * of the object it made, ref :747-751) whose vtable forwards to the library's ssb_search_t:
 * start / step (the frame's features are fetched from acmod->feat_buf through the library's
 * feature-source callback) / finish (the utterance runs through the batched kernels) / hyp /
 * seg_iter.  The flags and top-N lists that travel from the first pass to the second inside the
 * reference's single acmod (ref: src/state_align_search.c:186-188, src/ptm_mgau.c:426-440) are
 * handed over explicitly. */
#include <stdlib.h>

#include <soundswallower/alignment.h>
#include <soundswallower/ckd_alloc.h>
#include <soundswallower/dict.h>
#include <soundswallower/fsg_model.h>
#include <soundswallower/fsg_search.h>
#include <soundswallower/search_module.h>
#include <soundswallower/state_align_search.h>

static int g_n_search_init = 0;
int
ssb_glue_n_search_init(void)
{
    return g_n_search_init;
}

/* one lexicon per scorer model (dictionary + filler dictionary of the decoder's configuration);
 * the searches borrow it, so a few are kept: decoders that are alive at the same time (up to
 * GLUE_N_LX of them) never see theirs freed */
#define GLUE_N_LX 4
static struct { ssb_model_t *model; int serial; ssb_lexicon_t *lx; long used; } g_lxs[GLUE_N_LX];
static long g_lx_clock = 0;
/* acmod state after the first pass of the decoder's current utterance */
static uint32_t *g_flags = NULL;
static uint8_t *g_topn = NULL;
static int g_have_pass1 = 0;

static ssb_model_t *
glue_model(acmod_t *acmod)
{
    if (acmod->mgau == NULL || g_n_init == 0)
        return NULL;
    return ssb_mgau_model((const ssb_mgau_t *)acmod->mgau);
}

static ssb_lexicon_t *
glue_lexicon(ssb_model_t *m, config_t *config)
{
    const int serial = glue_model_serial(m);
    int i, at = 0;
    for (i = 0; i < GLUE_N_LX; ++i)
        if (g_lxs[i].lx && g_lxs[i].model == m && g_lxs[i].serial == serial) {
            g_lxs[i].used = ++g_lx_clock;
            return g_lxs[i].lx;
        }
    for (i = 1; i < GLUE_N_LX; ++i) /* a free slot, else the least recently used */
        if (g_lxs[i].lx == NULL ? g_lxs[at].lx != NULL : (g_lxs[at].lx != NULL && g_lxs[i].used < g_lxs[at].used))
            at = i;
    if (g_lxs[at].lx)
        ssb_lexicon_free(g_lxs[at].lx);
    g_lxs[at].lx = ssb_lexicon_load(m, config_str(config, "dict"), config_str(config, "fdict"));
    g_lxs[at].model = m;
    g_lxs[at].serial = serial;
    g_lxs[at].used = ++g_lx_clock;
    return g_lxs[at].lx;
}

/* what acmod_score scores for frame_idx (ref: src/acmod.c:765-802 calc_feat_idx) */
static const float *
glue_feat(void *ctx, int frame_idx)
{
    acmod_t *a = (acmod_t *)ctx;
    int n_backfr = a->n_feat_alloc - a->n_feat_frame, idx;
    if (frame_idx < 0 || a->output_frame - frame_idx > n_backfr)
        return NULL;
    idx = (a->feat_outidx + frame_idx - a->output_frame) % a->n_feat_alloc;
    if (idx < 0)
        idx += a->n_feat_alloc;
    return a->feat_buf[idx][0];
}

/* ---- seg_iter adapter: the library's iterator copied into a reference-shaped one ---- */
typedef struct glue_seg_s {
    seg_iter_t base;
    ssb_seg_iter_t *it;
} glue_seg_t;

static void
glue_seg_fill(glue_seg_t *g)
{
    g->base.word = g->it->word;
    g->base.sf = g->it->sf;
    g->base.ef = g->it->ef;
    g->base.ascr = g->it->ascr;
    g->base.lscr = g->it->lscr;
    g->base.prob = g->it->prob;
}
static void
glue_seg_free(seg_iter_t *seg)
{
    glue_seg_t *g = (glue_seg_t *)seg;
    if (g->it)
        g->it->vt->seg_free(g->it);
    ckd_free(g);
}
static seg_iter_t *
glue_seg_next(seg_iter_t *seg)
{
    glue_seg_t *g = (glue_seg_t *)seg;
    g->it = g->it->vt->seg_next(g->it); /* frees itself and returns NULL at the end */
    if (g->it == NULL) {
        ckd_free(g);
        return NULL;
    }
    glue_seg_fill(g);
    return seg;
}
static ps_segfuncs_t glue_segfuncs = { glue_seg_next, glue_seg_free };

/* ---- grammar search adapter ---- */
typedef struct glue_fsg_s {
    search_module_t base;
    ssb_search_t *impl;
} glue_fsg_t;

static int
glue_fsg_start(search_module_t *s)
{
    g_have_pass1 = 0;
    return ((glue_fsg_t *)s)->impl->vt->start(((glue_fsg_t *)s)->impl);
}
static int
glue_fsg_step(search_module_t *s, int frame_idx)
{
    return ((glue_fsg_t *)s)->impl->vt->step(((glue_fsg_t *)s)->impl, frame_idx);
}
static int
glue_fsg_finish(search_module_t *s)
{
    glue_fsg_t *g = (glue_fsg_t *)s;
    int rv = g->impl->vt->finish(g->impl);
    if (rv >= 0 && g_flags && g_topn) {
        /* what the reference's acmod holds when decoder_alignment starts */
        ssb_search_final_active(g->impl, g_flags);
        ssb_search_final_topn(g->impl, g_topn);
        g_have_pass1 = 1;
    }
    return rv;
}
static int
glue_reinit(search_module_t *s, dict_t *dict, dict2pid_t *d2p)
{
    (void)s; (void)dict; (void)d2p;
    return 0;
}
static void
glue_fsg_free(search_module_t *s)
{
    glue_fsg_t *g = (glue_fsg_t *)s;
    g->impl->vt->free(g->impl);
    search_module_base_free(s);
    ckd_free(g);
}
static lattice_t *
glue_lattice(search_module_t *s)
{
    (void)s;
    return NULL;
}
static const char *
glue_fsg_hyp(search_module_t *s, int32 *out_score)
{
    glue_fsg_t *g = (glue_fsg_t *)s;
    int32_t score = 0;
    const char *h = g->impl->vt->hyp(g->impl, &score);
    if (out_score)
        *out_score = score;
    ckd_free(s->hyp_str);
    s->hyp_str = h ? ckd_salloc(h) : NULL;
    return s->hyp_str;
}
static int32
glue_prob(search_module_t *s)
{
    (void)s;
    return 0;
}
static seg_iter_t *
glue_fsg_seg_iter(search_module_t *s)
{
    glue_fsg_t *g = (glue_fsg_t *)s;
    ssb_seg_iter_t *it = g->impl->vt->seg_iter(g->impl);
    glue_seg_t *seg;
    if (it == NULL)
        return NULL;
    seg = ckd_calloc(1, sizeof(*seg));
    seg->base.vt = &glue_segfuncs;
    seg->base.search = s;
    seg->it = it;
    glue_seg_fill(seg);
    return &seg->base;
}
static searchfuncs_t glue_fsg_funcs = {
    glue_fsg_start, glue_fsg_step, glue_fsg_finish, glue_reinit, glue_fsg_free,
    glue_lattice, glue_fsg_hyp, glue_prob, glue_fsg_seg_iter
};

/* Same contract as fsg_search_init (ref: src/fsg_search.c:171-260): consumes `fsg`. */
search_module_t *
ssb_glue_fsg_search_init(const char *name, fsg_model_t *fsg, config_t *config, acmod_t *acmod,
                         dict_t *dict, dict2pid_t *d2p)
{
    ssb_model_t *m = glue_model(acmod);
    ssb_lexicon_t *lx;
    ssb_fsg_config_t fc;
    ssb_fsg_built_t *built;
    glue_fsg_t *g;
    int32_t *from, *to, *logp;
    const char **word;
    int n = 0, cap = 0, s, i;

    if (m == NULL || (lx = glue_lexicon(m, config)) == NULL) {
        E_ERROR("B200 grammar search unavailable: %s\n", ssb_last_error());
        return NULL;
    }
    /* the grammar's transitions in an insertion order that reproduces its arc order: per state
     * and destination, fsg_model_arcs lists the most recently added link first */
    for (s = 0; s < fsg_model_n_state(fsg); ++s) {
        fsg_arciter_t *it;
        for (it = fsg_model_arcs(fsg, s); it; it = fsg_arciter_next(it))
            ++cap;
    }
    from = ckd_calloc(cap + 1, sizeof(*from));
    to = ckd_calloc(cap + 1, sizeof(*to));
    logp = ckd_calloc(cap + 1, sizeof(*logp));
    word = ckd_calloc(cap + 1, sizeof(*word));
    for (s = 0; s < fsg_model_n_state(fsg); ++s) {
        fsg_arciter_t *it;
        int first = n;
        for (it = fsg_model_arcs(fsg, s); it; it = fsg_arciter_next(it)) {
            fsg_link_t *l = fsg_arciter_get(it);
            from[n] = fsg_link_from_state(l);
            to[n] = fsg_link_to_state(l);
            logp[n] = fsg_link_logs2prob(l);
            word[n] = fsg_link_wid(l) >= 0 ? fsg_model_word_str(fsg, fsg_link_wid(l)) : NULL;
            ++n;
        }
        /* runs with the same destination: reverse (oldest first) */
        for (i = first; i < n;) {
            int j = i, a, b;
            while (j + 1 < n && to[j + 1] == to[i] && (word[j + 1] != NULL) == (word[i] != NULL))
                ++j;
            for (a = i, b = j; a < b; ++a, --b) {
                int32_t t32;
                const char *tw;
                t32 = logp[a]; logp[a] = logp[b]; logp[b] = t32;
                tw = word[a]; word[a] = word[b]; word[b] = tw;
            }
            i = j + 1;
        }
    }
    ssb_fsg_config_defaults(&fc);
    fc.beam = config_float(config, "beam");
    fc.pbeam = config_float(config, "pbeam");
    fc.wbeam = config_float(config, "wbeam");
    fc.lw = config_float(config, "lw");
    fc.wip = config_float(config, "wip");
    fc.pip = config_float(config, "pip");
    fc.silprob = config_float(config, "silprob");
    fc.fillprob = config_float(config, "fillprob");
    fc.maxhmmpf = config_int(config, "maxhmmpf");
    fc.fsgusefiller = config_bool(config, "fsgusefiller");
    fc.fsgusealtpron = config_bool(config, "fsgusealtpron");
    built = ssb_fsg_build_logp(lx, fsg_model_n_state(fsg), fsg_model_start_state(fsg),
                               fsg_model_final_state(fsg), n, from, to, logp, word, 0, &fc);
    ckd_free(from);
    ckd_free(to);
    ckd_free(logp);
    ckd_free((void *)word);
    if (built == NULL) {
        E_ERROR("B200 grammar search: %s\n", ssb_last_error());
        return NULL;
    }
    g = ckd_calloc(1, sizeof(*g));
    search_module_init(&g->base, &glue_fsg_funcs, PS_SEARCH_TYPE_FSG, name, config, acmod, dict, d2p);
    g->impl = ssb_fsg_search_init(name, m, lx, built, glue_feat, acmod); /* consumes built */
    if (g->impl == NULL) {
        E_ERROR("B200 grammar search: %s\n", ssb_last_error());
        search_module_base_free(&g->base);
        ckd_free(g);
        return NULL;
    }
    if (g_flags == NULL) {
        int32_t dims[16];
        ssb_model_dims(m, dims);
        g_flags = ckd_calloc((dims[4] + 31) / 32 + 1, sizeof(*g_flags));
        g_topn = ckd_calloc((size_t)dims[0] * dims[1] * 4 + 4, 1);
    }
    fsg_model_free(fsg); /* (the reference's search keeps it; ours has its own flattened copy) */
    ++g_n_search_init;
    return &g->base;
}

/* ---- aligner adapter: the reference's struct in front ---- */
typedef struct glue_align_s {
    state_align_search_t ref;
    ssb_search_t *impl;
} glue_align_t;

static int
glue_align_start(search_module_t *s)
{
    glue_align_t *g = (glue_align_t *)s;
    g->ref.frame = 0;
    return g->impl->vt->start(g->impl);
}
static int
glue_align_step(search_module_t *s, int frame_idx)
{
    glue_align_t *g = (glue_align_t *)s;
    int rv = g->impl->vt->step(g->impl, frame_idx);
    g->ref.frame = frame_idx + 1;
    return rv;
}
static int
glue_align_finish(search_module_t *s)
{
    glue_align_t *g = (glue_align_t *)s;
    alignment_t *al = g->ref.al;
    int rv = g->impl->vt->finish(g->impl);
    int n = alignment_n_states(al), i;
    int32_t *e;
    if (rv < 0)
        return rv;
    /* state entries as state_align_search_finish leaves them (ref: src/state_align_search.c:
     * 215-268), then the reference's own alignment_propagate */
    e = ckd_calloc((size_t)n * 5 + 5, sizeof(*e));
    if (ssb_search_alignment(g->impl, 2, e, n) != n) {
        ckd_free(e);
        E_ERROR("B200 aligner: state level does not match alignment_populate's\n");
        return -1;
    }
    for (i = 0; i < n; ++i) {
        al->state.seq[i].start = e[i * 5 + 1];
        al->state.seq[i].duration = e[i * 5 + 2];
        al->state.seq[i].score = e[i * 5 + 3];
    }
    ckd_free(e);
    alignment_propagate(al);
    return rv;
}
static void
glue_align_free(search_module_t *s)
{
    glue_align_t *g = (glue_align_t *)s;
    g->impl->vt->free(g->impl);
    alignment_free(g->ref.al); /* consumed at init, like the reference */
    search_module_base_free(s);
    ckd_free(g);
}
static const char *
glue_align_hyp(search_module_t *s, int32 *out_score)
{
    (void)s;
    if (out_score)
        *out_score = 0;
    return NULL;
}
static seg_iter_t *
glue_align_seg_iter(search_module_t *s)
{
    (void)s;
    return NULL;
}
static searchfuncs_t glue_align_funcs = {
    glue_align_start, glue_align_step, glue_align_finish, glue_reinit, glue_align_free,
    NULL, glue_align_hyp, NULL, glue_align_seg_iter
};

/* Same contract as state_align_search_init (ref: src/state_align_search.c:429-474): consumes
 * `al`, whose word level (ids and frame windows) defines the chain. */
search_module_t *
ssb_glue_state_align_search_init(const char *name, config_t *config, acmod_t *acmod, alignment_t *al)
{
    ssb_model_t *m = glue_model(acmod);
    ssb_lexicon_t *lx;
    glue_align_t *g;
    int nw = alignment_n_words(al), i;
    int32_t *wid, *ws, *wd;

    if (m == NULL || (lx = glue_lexicon(m, config)) == NULL) {
        E_ERROR("B200 aligner unavailable: %s\n", ssb_last_error());
        return NULL;
    }
    wid = ckd_calloc(nw + 1, sizeof(*wid));
    ws = ckd_calloc(nw + 1, sizeof(*ws));
    wd = ckd_calloc(nw + 1, sizeof(*wd));
    for (i = 0; i < nw; ++i) {
        /* dictionary ids are the same on both sides (same files, same order) */
        wid[i] = al->word.seq[i].id.wid;
        ws[i] = al->word.seq[i].start;
        wd[i] = al->word.seq[i].duration;
    }
    g = ckd_calloc(1, sizeof(*g));
    search_module_init(&g->ref.base, &glue_align_funcs, PS_SEARCH_TYPE_STATE_ALIGN, name, config, acmod,
                       al->d2p->dict, al->d2p);
    g->ref.al = al;
    g->impl = ssb_state_align_search_init(name, m, lx, wid, ws, wd, nw, glue_feat, acmod);
    ckd_free(wid);
    ckd_free(ws);
    ckd_free(wd);
    if (g->impl == NULL) {
        E_ERROR("B200 aligner: %s\n", ssb_last_error());
        search_module_base_free(&g->ref.base);
        ckd_free(g);
        return NULL;
    }
    if (g_have_pass1) {
        ssb_search_set_init_active(g->impl, g_flags);
        ssb_search_set_init_topn(g->impl, g_topn);
    }
    ++g_n_search_init;
    return &g->ref.base;
}
