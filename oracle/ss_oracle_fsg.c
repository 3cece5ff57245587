/*
 * ss_oracle_fsg.c -- CPU oracle, part 2: FSG token-passing search on a flattened lextree.
 *
 * TEST INFRASTRUCTURE ONLY (see ss_oracle.h).  Plain-C restatement of
 *   fsg_search_start / step / finish          ref: src/fsg_search.c:746-851, 664-739
 *   hmm_eval / prune_prop / pnode_trans / pnode_exit / null_prop / word_trans
 *                                             ref: src/fsg_search.c:330-662
 *   fsg_history_entry_add / end_frame         ref: src/fsg_history.c:129-232
 *   fsg_search_find_exit, seg iterator        ref: src/fsg_search.c:853-924, 1030-1142
 * operating on the arrays oracle/ref_shim.c:ref_fsg_dump exports from the reference's own
 * fsg_model_t / fsg_lextree_t (graph construction is host-side preparation and is NOT
 * restated), and on dense senone scores (`compallsen` semantics).
 *
 * Parity status: PINNED -- tests/test_oracle_fsg.py compares the full history table, the
 * number of HMM evaluations, the hypothesis score and the word segmentation with the
 * reference's on the bundled test audio (alignment grammar and goforward.gram, en-us/fr-fr).
 */
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include "ss_oracle.h"

typedef struct {
    int32_t st[12]; /* score[5] hist[5] out_score out_hist */
    int32_t frame, bestscore;
} phmm_t;

typedef struct {
    int32_t link, score, pred, frame, lc;
    uint32_t rc[4];
} hent_t;

typedef struct {
    hent_t e;
    int next;
} tent_t;

typedef struct {
    const orc_model_t *m;
    const orc_fsg_t *g;
    phmm_t *h;
    int *act, n_act, *nxt, n_nxt; /* "prepend" lists: appended here, walked backwards */
    hent_t *hist;
    int n_hist, cap;
    tent_t *tent;
    int n_tent, cap_tent;
    int *heads; /* [n_state][n_ciphone] */
    int32_t frame, bestscore, bpidx_start;
    int32_t beam, pbeam, wbeam;
    float beam_factor;
    int64_t n_hmm_eval;
    int overflow;
    /* active-list mode (the reference's default, compallsen = no): scores are computed frame
     * by frame for the senones of the active HMMs only (ref: src/fsg_search.c:309-328, 686-690) */
    orc_ptm_t *ptm;
    const float *feat;
    uint32_t *bits;   /* acmod's senone_active_vec; left as the last frame set it */
    uint8_t *list;
    int16_t *scr;
    int64_t n_sen_eval;
} fs_t;

#define PN(g, i, k) ((g)->pnode8[(i) * 8 + (k)])

static void
hclear(phmm_t *h)
{
    int i;
    for (i = 0; i < 5; ++i) {
        h->st[i] = ORC_WORST_SCORE;
        h->st[5 + i] = -1;
    }
    h->st[10] = ORC_WORST_SCORE;
    h->st[11] = -1;
    h->bestscore = ORC_WORST_SCORE;
    h->frame = -1;
}

static void
henter(phmm_t *h, int32_t score, int32_t hist, int frame)
{
    h->st[0] = score;
    h->st[5] = hist;
    h->frame = frame;
}

static void
hist_append(fs_t *s, const hent_t *e)
{
    if (s->n_hist >= s->cap) {
        s->overflow = 1;
        return;
    }
    s->hist[s->n_hist++] = *e;
}

/* ref: src/fsg_history.c:129-202 */
static void
entry_add(fs_t *s, int link, int32_t frame, int32_t score, int32_t pred, int32_t lc,
          const uint32_t *rc_in)
{
    hent_t ne;
    uint32_t rc[4];
    int *head, gn, prev, k;
    memcpy(rc, rc_in, sizeof(rc));
    ne.link = link;
    ne.frame = frame;
    ne.score = score;
    ne.pred = pred;
    ne.lc = lc;
    if (frame < 0) {
        memcpy(ne.rc, rc, sizeof(rc));
        hist_append(s, &ne);
        return;
    }
    head = &s->heads[s->g->link4[link * 4 + 1] * s->g->n_ciphone + lc];
    prev = -1;
    for (gn = *head; gn >= 0; gn = s->tent[gn].next) {
        uint32_t left = 0;
        hent_t *e = &s->tent[gn].e;
        if (score > e->score)
            break;
        for (k = 0; k < 4; ++k)
            left |= (rc[k] = ~e->rc[k] & rc[k]);
        if (left == 0)
            return;
        prev = gn;
    }
    if (s->n_tent >= s->cap_tent) {
        s->cap_tent *= 2;
        s->tent = realloc(s->tent, sizeof(tent_t) * s->cap_tent);
    }
    memcpy(ne.rc, rc, sizeof(rc));
    k = s->n_tent++;
    s->tent[k].e = ne;
    s->tent[k].next = gn;
    if (prev < 0)
        *head = k;
    else
        s->tent[prev].next = k;
    prev = k;
    while (gn >= 0) {
        uint32_t left = 0;
        hent_t *e = &s->tent[gn].e;
        int j;
        for (j = 0; j < 4; ++j)
            left |= (e->rc[j] = ~rc[j] & e->rc[j]);
        if (left == 0) {
            s->tent[prev].next = s->tent[gn].next; /* pruned */
            gn = s->tent[gn].next;
        } else {
            prev = gn;
            gn = s->tent[gn].next;
        }
    }
}

/* ref: src/fsg_history.c:208-232 */
static void
end_frame(fs_t *s)
{
    int i, n = s->g->n_state * s->g->n_ciphone, gn;
    for (i = 0; i < n; ++i) {
        for (gn = s->heads[i]; gn >= 0; gn = s->tent[gn].next)
            hist_append(s, &s->tent[gn].e);
        s->heads[i] = -1;
    }
    s->n_tent = 0;
}

/* ref: src/fsg_search.c:543-591 */
static void
null_prop(fs_t *s)
{
    const orc_fsg_t *g = s->g;
    int32_t thresh = s->bestscore + s->wbeam;
    int bp, n = s->n_hist, a;
    for (bp = s->bpidx_start; bp < n; ++bp) {
        hent_t he = s->hist[bp]; /* copy: the table may grow (frame < 0 entries) */
        int st = he.link >= 0 ? g->link4[he.link * 4 + 1] : g->start;
        for (a = g->arc_off[st]; a < g->arc_off[st + 1]; ++a) {
            int32_t newscore;
            if (g->link4[a * 4 + 3] != -1)
                continue;
            newscore = he.score + (g->link4[a * 4 + 2] >> ORC_SENSCR_SHIFT);
            if (newscore >= thresh)
                entry_add(s, a, he.frame, newscore, bp, he.lc, he.rc);
        }
    }
}

/* ref: src/fsg_search.c:597-662 */
static void
word_trans(fs_t *s)
{
    const orc_fsg_t *g = s->g;
    int32_t thresh = s->bestscore + s->beam, nf = s->frame + 1;
    int bp, n = s->n_hist, root;
    for (bp = s->bpidx_start; bp < n; ++bp) {
        const hent_t *he = &s->hist[bp];
        int d = he->link >= 0 ? g->link4[he->link * 4 + 1] : g->start;
        int lc = he->lc;
        for (root = g->root[d]; root >= 0; root = PN(g, root, 6)) {
            int rc = PN(g, root, 3);
            if ((g->ctxt[root * 4 + (lc >> 5)] & (1u << (lc & 31)))
                && (he->rc[rc >> 5] & (1u << (rc & 31)))) {
                int32_t newscore = he->score + PN(g, root, 2);
                phmm_t *h = &s->h[root];
                if (newscore > thresh && newscore > h->st[0]) {
                    if (h->frame < nf)
                        s->nxt[s->n_nxt++] = root;
                    henter(h, newscore, bp, nf);
                }
            }
        }
    }
}

/* ref: src/fsg_search.c:400-428 */
static void
pnode_trans(fs_t *s, int pn)
{
    const orc_fsg_t *g = s->g;
    int32_t nf = s->frame + 1, thresh = s->bestscore + s->beam;
    phmm_t *h = &s->h[pn];
    int child;
    for (child = PN(g, pn, 5); child >= 0; child = PN(g, child, 6)) {
        int32_t newscore = h->st[10] + PN(g, child, 2);
        phmm_t *c = &s->h[child];
        if (newscore > thresh && newscore > c->st[0]) {
            if (c->frame < nf)
                s->nxt[s->n_nxt++] = child;
            henter(c, newscore, h->st[11], nf);
        }
    }
}

/* ref: src/fsg_search.c:430-489 */
static void
pnode_exit(fs_t *s, int pn)
{
    const orc_fsg_t *g = s->g;
    phmm_t *h = &s->h[pn];
    int link = PN(g, pn, 5);
    static const uint32_t all[4] = { 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu };
    const uint32_t *rc = (g->link_flag[link] & 1) ? all : &g->ctxt[pn * 4];
    entry_add(s, link, s->frame, h->st[10], h->st[11], PN(g, pn, 3), rc);
}

static int
fsg_run(const orc_model_t *m, const orc_fsg_t *g, const int16_t *senscr, int T, fs_t *s)
{
    static const uint32_t all[4] = { 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu };
    int i, t, E = m->n_emit;
    if (E != 3 && E != 5)
        return -1;
    s->m = m;
    s->g = g;
    s->h = malloc(sizeof(phmm_t) * (g->n_pnode + 1));
    s->act = malloc(sizeof(int) * (g->n_pnode + 1));
    s->nxt = malloc(sizeof(int) * (g->n_pnode + 1));
    s->heads = malloc(sizeof(int) * g->n_state * g->n_ciphone);
    s->cap_tent = 64;
    s->tent = malloc(sizeof(tent_t) * s->cap_tent);
    for (i = 0; i < g->n_pnode; ++i)
        hclear(&s->h[i]);
    for (i = 0; i < g->n_state * g->n_ciphone; ++i)
        s->heads[i] = -1;
    /* fsg_search_start (ref :746-798) */
    s->beam_factor = 1.0f;
    s->beam = g->beam;
    s->pbeam = g->pbeam;
    s->wbeam = g->wbeam;
    s->frame = -1;
    s->bestscore = 0;
    entry_add(s, -1, -1, 0, -1, g->sil, all);
    s->bpidx_start = 0;
    null_prop(s);
    word_trans(s);
    memcpy(s->act, s->nxt, sizeof(int) * s->n_nxt);
    s->n_act = s->n_nxt;
    s->n_nxt = 0;
    s->frame = 0;
    for (t = 0; t < T; ++t) {
        const int16_t *ss = senscr ? senscr + (size_t)t * m->n_sen : s->scr;
        int32_t thresh, phone_thresh, word_thresh;
        if (!senscr) {
            /* fsg_search_sen_active (ref :309-328): acmod_clear_active, then acmod_activate_hmm
             * for every HMM of the active list; acmod_score -> acmod_flags2list -> frame_eval
             * (ref: src/acmod.c:822-860, 889-999) */
            int n_list, k;
            memset(s->bits, 0, sizeof(uint32_t) * ((m->n_sen + 31) / 32));
            for (i = 0; i < s->n_act; ++i) {
                const uint16_t *sid = m->sseq + (size_t)PN(g, s->act[i], 0) * E;
                for (k = 0; k < E; ++k)
                    s->bits[sid[k] >> 5] |= 1u << (sid[k] & 31);
            }
            n_list = orc_flags2list(s->bits, m->n_sen, s->list);
            orc_ptm_frame_eval(s->ptm, s->scr, s->list, n_list, s->feat + (size_t)t * m->blk, t, 0, NULL);
            s->n_sen_eval += n_list;
        }
        s->bpidx_start = s->n_hist;
        /* fsg_search_hmm_eval (ref :330-398) */
        if (s->n_act > 0) {
            int32_t best = ORC_WORST_SCORE;
            for (i = s->n_act - 1; i >= 0; --i) {
                int pn = s->act[i];
                phmm_t *h = &s->h[pn];
                int32_t sc = orc_hmm_eval(E, m->tp + (size_t)PN(g, pn, 1) * E * (E + 1),
                                          m->sseq + (size_t)PN(g, pn, 0) * E, ss, h->st);
                h->bestscore = sc;
                if (sc > best)
                    best = sc;
            }
            s->n_hmm_eval += s->n_act;
            if (g->maxhmmpf != -1 && s->n_act > g->maxhmmpf) {
                if (s->beam_factor > 0.1) {
                    s->beam_factor *= 0.9f;
                    s->beam = (int32_t)(g->beam * s->beam_factor);
                    s->pbeam = (int32_t)(g->pbeam * s->beam_factor);
                    s->wbeam = (int32_t)(g->wbeam * s->beam_factor);
                }
            } else {
                s->beam_factor = 1.0f;
                s->beam = g->beam;
                s->pbeam = g->pbeam;
                s->wbeam = g->wbeam;
            }
            s->bestscore = best;
        }
        /* fsg_search_hmm_prune_prop (ref :497-538) */
        thresh = s->bestscore + s->beam;
        phone_thresh = s->bestscore + s->pbeam;
        word_thresh = s->bestscore + s->wbeam;
        for (i = s->n_act - 1; i >= 0; --i) {
            int pn = s->act[i];
            phmm_t *h = &s->h[pn];
            if (h->bestscore >= thresh) {
                if (h->frame == s->frame) {
                    h->frame = s->frame + 1;
                    s->nxt[s->n_nxt++] = pn;
                }
                if (!PN(g, pn, 4)) {
                    if (h->st[10] >= phone_thresh)
                        pnode_trans(s, pn);
                } else {
                    if (h->st[10] >= word_thresh)
                        pnode_exit(s, pn);
                }
            }
        }
        end_frame(s);
        null_prop(s);
        end_frame(s);
        word_trans(s);
        for (i = s->n_act - 1; i >= 0; --i) {
            phmm_t *h = &s->h[s->act[i]];
            if (h->frame == s->frame)
                hclear(h);
        }
        {
            int *tmp = s->act;
            s->act = s->nxt;
            s->nxt = tmp;
            s->n_act = s->n_nxt;
            s->n_nxt = 0;
        }
        ++s->frame;
    }
    free(s->h);
    free(s->act);
    free(s->nxt);
    free(s->heads);
    free(s->tent);
    return s->overflow ? -2 : 0;
}

/* ref: src/fsg_search.c:853-924 */
int
orc_fsg_find_exit(const orc_fsg_t *g, const int32_t *hist9, int n_hist, int frame_idx, int final,
                  int32_t *out_score)
{
    int bpidx = n_hist - 1, frm, last_frm, besthist = -1;
    int32_t bestscore = INT_MIN;
    const int32_t *e = NULL;
    last_frm = frm = frame_idx;
    while (bpidx > 0) {
        e = hist9 + (size_t)bpidx * 9;
        if (e[3] <= frame_idx) {
            frm = last_frm = e[3];
            break;
        }
        bpidx--;
    }
    if (bpidx <= 0)
        return bpidx;
    while (frm == last_frm) {
        int link = e[0];
        int32_t score = e[1];
        if (link < 0)
            break;
        if (score == bestscore && g->link4[link * 4 + 1] == g->final) {
            besthist = bpidx;
        } else if (score > bestscore) {
            if (!final || g->link4[link * 4 + 1] == g->final) {
                bestscore = score;
                besthist = bpidx;
            }
        }
        --bpidx;
        if (bpidx < 0)
            break;
        e = hist9 + (size_t)bpidx * 9;
        frm = e[3];
    }
    if (besthist == -1)
        return -1;
    if (out_score)
        *out_score = bestscore;
    return besthist;
}

/* ref: src/fsg_search.c:1030-1054, 1110-1142.  segs [n][5] = link sf ef ascr lscr, first word first */
int
orc_fsg_segs(const orc_fsg_t *g, const int32_t *hist9, int bpidx, int32_t *segs, int max_seg)
{
    int n = 0, bp, cur;
    for (bp = bpidx; bp > 0; bp = hist9[(size_t)bp * 9 + 2])
        ++n;
    if (n > max_seg)
        return -2;
    cur = n - 1;
    for (bp = bpidx; bp > 0; bp = hist9[(size_t)bp * 9 + 2], --cur) {
        const int32_t *e = hist9 + (size_t)bp * 9;
        const int32_t *ph = e[2] >= 0 ? hist9 + (size_t)e[2] * 9 : NULL;
        int32_t sf = ph ? ph[3] + 1 : 0, ef = e[3];
        int32_t lscr = g->link4[e[0] * 4 + 2] >> ORC_SENSCR_SHIFT;
        if (sf > ef)
            sf = ef;
        segs[cur * 5 + 0] = e[0];
        segs[cur * 5 + 1] = sf;
        segs[cur * 5 + 2] = ef;
        segs[cur * 5 + 3] = ph ? e[1] - ph[1] - lscr : e[1] - lscr;
        segs[cur * 5 + 4] = lscr;
    }
    return n;
}

/* hist9 [cap][9] = link score pred frame lc rc[4]; out[0] = #entries, out[1] = #HMM evaluations,
 * out[2] = frames searched.  Returns 0, -1 (unsupported model), -2 (history overflow). */
int
orc_fsg_search(const orc_model_t *m, const orc_fsg_t *g, const int16_t *senscr, int T,
               int32_t *hist9, int cap, int64_t *out)
{
    fs_t s;
    int i, rv;
    memset(&s, 0, sizeof(s));
    s.hist = malloc(sizeof(hent_t) * (cap > 0 ? cap : 1));
    s.cap = cap;
    rv = fsg_run(m, g, senscr, T, &s);
    for (i = 0; i < s.n_hist; ++i) {
        int32_t *o = hist9 + (size_t)i * 9;
        o[0] = s.hist[i].link;
        o[1] = s.hist[i].score;
        o[2] = s.hist[i].pred;
        o[3] = s.hist[i].frame;
        o[4] = s.hist[i].lc;
        memcpy(o + 5, s.hist[i].rc, 16);
    }
    out[0] = s.n_hist;
    out[1] = s.n_hmm_eval;
    out[2] = s.frame;
    free(s.hist);
    return rv;
}

/* The same search in the reference's default mode (compallsen = no): senone scores are
 * computed frame by frame by the PTM scorer for the active HMMs' senones only.  `active_out`
 * (optional, (n_sen+31)/32 words) receives acmod's active-senone flags as the last frame left
 * them -- what a following state_align_search starts with (ref: src/state_align_search.c:
 * 186-188 never clears them).  out[3] = senones evaluated (fsgs->n_sen_eval). */
int
orc_fsg_search_active(const orc_model_t *m, int topn, const orc_fsg_t *g, const float *feat, int T,
                      int32_t *hist9, int cap, int64_t *out, uint32_t *active_out)
{
    return orc_fsg_search_active2(m, topn, g, feat, T, hist9, cap, out, active_out, NULL);
}

/* carried_out (optional, [mgau*feat][topn]): the scorer's carried top-N codewords after the
 * search -- what a second pass on the same decoder starts from (orc_ptm_get_carried). */
int
orc_fsg_search_active2(const orc_model_t *m, int topn, const orc_fsg_t *g, const float *feat, int T,
                       int32_t *hist9, int cap, int64_t *out, uint32_t *active_out,
                       uint8_t *carried_out)
{
    fs_t s;
    int i, rv, nw = (m->n_sen + 31) / 32;
    if (m->kind != ORC_KIND_PTM)
        return -1;
    memset(&s, 0, sizeof(s));
    s.hist = malloc(sizeof(hent_t) * (cap > 0 ? cap : 1));
    s.cap = cap;
    s.ptm = orc_ptm_new(m, topn, 1);
    s.feat = feat;
    s.bits = calloc(nw, sizeof(uint32_t));
    s.list = malloc(m->n_sen + 16);
    s.scr = calloc(m->n_sen, sizeof(int16_t));
    rv = fsg_run(m, g, NULL, T, &s);
    for (i = 0; i < s.n_hist; ++i) {
        int32_t *o = hist9 + (size_t)i * 9;
        o[0] = s.hist[i].link;
        o[1] = s.hist[i].score;
        o[2] = s.hist[i].pred;
        o[3] = s.hist[i].frame;
        o[4] = s.hist[i].lc;
        memcpy(o + 5, s.hist[i].rc, 16);
    }
    out[0] = s.n_hist;
    out[1] = s.n_hmm_eval;
    out[2] = s.frame;
    out[3] = s.n_sen_eval;
    if (active_out)
        memcpy(active_out, s.bits, sizeof(uint32_t) * nw);
    if (carried_out)
        orc_ptm_get_carried(s.ptm, carried_out);
    orc_ptm_free(s.ptm);
    free(s.bits);
    free(s.list);
    free(s.scr);
    free(s.hist);
    return rv;
}
