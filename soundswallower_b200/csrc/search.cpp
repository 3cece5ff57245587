// search.cpp -- the search-module drop-in: objects with the layout and vtable of the
// reference's search_module_t (ref: include/soundswallower/search_module.h:72-113, 157-174) for
// the two searches on the hot path,
//   state_align_search (ref: src/state_align_search.c:46-474) and
//   fsg_search         (ref: src/fsg_search.c:171-260, 664-851, 945-1142).
// The reference pulls one frame of senone scores per step() out of its acmod; a GPU cannot be
// driven one frame of one utterance at a time, so step() only collects the frame's feature
// vector (from the acmod through a feature-source callback, or from frames fed with
// ssb_search_feed) and finish() runs the utterance through the batched kernels
// (ssb_align_batch / ssb_fsg_batch).  Everything the reference exposes after finish() -- hyp,
// seg_iter, the alignment's word / phone / state entries -- is then served from the results,
// with the reference's own quirks (hyp score of the aligner = score of the last word;
// "(NULL)" segments for null transitions; sf clipped to ef).
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "model.h"

namespace ssb {
const HostModel *model_host(const ssb_model_t *m);  // api.cu
}
using namespace ssb;

namespace {

struct Entry {
    int32_t id, start, dur, score, parent;
};

struct SearchImpl {
    ssb_search_t base;  // must stay first: callers hold ssb_search_t* / search_module_t*
    int kind = 0;       // 0 state_align, 1 fsg
    ssb_model_t *m = nullptr;
    const ssb_lexicon_t *lx = nullptr;
    ssb_feat_source_fn src = nullptr;
    void *src_ctx = nullptr;
    int blk = 0, n_emit = 0;
    std::vector<float> feat;  // frames collected since start()
    int n_frames = 0;
    bool finished = false;
    std::string type_s, name_s;
    // state_align
    std::vector<Entry> word, phone, state;
    std::vector<int32_t> ssid, tmat, sf, ef;
    int32_t best_score = 0;
    // fsg
    ssb_fsg_built_t *fsg = nullptr;
    std::vector<int32_t> segs;  // [n][5] link sf ef ascr lscr
    int32_t n_seg = 0, hyp_score = 0, exit_bp = 0, n_hist = 0;
    int result_frames = -1;  // frames the results above were computed on (hyp between two steps)
    std::vector<std::string> seg_word_store;
    // acmod's active-senone flags: left by the grammar search / found by the aligner
    std::vector<uint32_t> flags;
    std::vector<uint8_t> carried;  // the scorer's carried top-N codewords ([CS][4]), likewise
    int n_sen = 0, n_cs = 0;
};

struct SegImpl {
    ssb_seg_iter_t base;
    int cur = 0;
};

SearchImpl *impl(ssb_search_t *s) { return reinterpret_cast<SearchImpl *>(s); }

int search_start(ssb_search_t *s)
{
    SearchImpl *S = impl(s);
    S->feat.clear();
    S->n_frames = 0;
    S->finished = false;
    S->n_seg = 0;
    S->exit_bp = 0;
    S->hyp_score = 0;
    S->result_frames = -1;
    return 0;
}

int search_step(ssb_search_t *s, int frame_idx)
{
    SearchImpl *S = impl(s);
    if (frame_idx != S->n_frames) {
        set_error("search step: frame %d out of order (expected %d)", frame_idx, S->n_frames);
        return -1;
    }
    if (S->src) {
        const float *x = S->src(S->src_ctx, frame_idx);
        if (!x) {
            set_error("search step: the feature source has no frame %d", frame_idx);
            return -1;
        }
        S->feat.insert(S->feat.end(), x, x + S->blk);
    } else if ((size_t)(frame_idx + 1) * S->blk > S->feat.size()) {
        set_error("search step: frame %d has not been fed (ssb_search_feed)", frame_idx);
        return -1;
    }
    ++S->n_frames;
    // fsg_search_step returns 1, state_align_search_step 0 (ref: src/fsg_search.c:738,
    // src/state_align_search.c:212)
    return S->kind == 1 ? 1 : 0;
}

// ---- state_align
int align_finish(ssb_search_t *s)
{
    SearchImpl *S = impl(s);
    const int np = (int)S->phone.size(), ns = (int)S->state.size();
    const int64_t frame_off[2] = {0, S->n_frames}, phone_off[2] = {0, np};
    ssb_align_in_t in;
    memset(&in, 0, sizeof in);
    in.n_utts = 1;
    in.feat = S->feat.data();
    in.frame_off = frame_off;
    in.phone_off = phone_off;
    in.ssid = S->ssid.data();
    in.tmat = S->tmat.data();
    in.sf = S->sf.data();
    in.ef = S->ef.data();
    in.init_active = S->flags.empty() ? nullptr : S->flags.data();
    in.init_topn = S->carried.empty() ? nullptr : S->carried.data();
    std::vector<int32_t> st(ns), du(ns), sc(ns);
    for (int i = 0; i < ns; ++i) {
        st[i] = S->state[i].start;
        du[i] = S->state[i].dur;
        sc[i] = S->state[i].score;
    }
    int32_t rv = 0, best = 0, ren = 0;
    ssb_align_out_t out;
    memset(&out, 0, sizeof out);
    out.st_start = st.data();
    out.st_dur = du.data();
    out.st_score = sc.data();
    out.utt_rv = &rv;
    out.utt_best = &best;
    out.utt_renorm = &ren;
    if (ssb_align_batch(S->m, &in, &out) != 0)
        return -1;
    if (rv != 0) {
        set_error("Failed to reach final state in alignment");
        return -1;
    }
    S->best_score = best;
    for (int i = 0; i < ns; ++i) {
        S->state[i].start = st[i];
        S->state[i].dur = du[i];
        S->state[i].score = sc[i];
    }
    // alignment_propagate (ref: src/ps_alignment.c:317-355)
    int last = -1;
    for (const Entry &e : S->state) {
        Entry &p = S->phone[e.parent];
        if (e.parent != last) {
            p.start = e.start;
            p.dur = 0;
            p.score = 0;
        }
        p.dur += e.dur;
        p.score += e.score;
        last = e.parent;
    }
    last = -1;
    for (const Entry &p : S->phone) {
        Entry &w = S->word[p.parent];
        if (p.parent != last) {
            w.start = p.start;
            w.dur = 0;
            w.score = 0;
        }
        w.dur += p.dur;
        w.score += p.score;
        last = p.parent;
    }
    S->finished = true;
    return 0;
}

// dict_basestr (ref: src/dict.c: dict_basewid / dict_wordstr)
const char *basestr(const SearchImpl *S, int32_t wid)
{
    const int32_t b = ssb_lexicon_basewid(S->lx, wid);
    return ssb_lexicon_wordstr(S->lx, b >= 0 ? b : wid);
}

// dict_real_word (ref: src/dict.c:386-398)
bool real_word(const SearchImpl *S, int32_t wid)
{
    if (wid < 0)
        return false;
    const int32_t b = ssb_lexicon_basewid(S->lx, wid);
    if (b == S->base.start_wid || b == S->base.finish_wid)
        return false;
    return ssb_lexicon_is_filler(S->lx, wid) == 0;
}

void set_hyp(SearchImpl *S, const std::string &h)
{
    free(S->base.hyp_str);
    S->base.hyp_str = (char *)malloc(h.size() + 1);
    memcpy(S->base.hyp_str, h.c_str(), h.size() + 1);
}

const char *align_hyp(ssb_search_t *s, int32_t *out_score)
{
    // ref: src/state_align_search.c:365-411 -- the real words of the alignment joined by
    // spaces; *out_score ends up as the score of the LAST word entry
    SearchImpl *S = impl(s);
    free(S->base.hyp_str);
    S->base.hyp_str = nullptr;
    if (S->word.empty())
        return nullptr;
    std::string h;
    for (const Entry &w : S->word) {
        if (real_word(S, w.id)) {
            h += basestr(S, w.id);
            h += ' ';
        }
        if (out_score)
            *out_score = w.score;
    }
    if (!h.empty())
        h.pop_back();
    set_hyp(S, h);
    return S->base.hyp_str;
}

ssb_seg_iter_t *seg_fill(SegImpl *it);
bool fsg_have_result(SearchImpl *S);

void seg_free(ssb_seg_iter_t *seg) { delete reinterpret_cast<SegImpl *>(seg); }

ssb_seg_iter_t *seg_next(ssb_seg_iter_t *seg)
{
    SegImpl *it = reinterpret_cast<SegImpl *>(seg);
    ++it->cur;
    ssb_seg_iter_t *r = seg_fill(it);
    if (!r)
        seg_free(seg);  // like the reference: the iterator is freed when it runs off the end
    return r;
}

ssb_segfuncs_t g_segfuncs = {seg_next, seg_free};

ssb_seg_iter_t *seg_fill(SegImpl *it)
{
    SearchImpl *S = impl(it->base.search);
    if (S->kind == 0) {
        // ref: src/state_align_search.c:312-324
        if (it->cur >= (int)S->word.size())
            return nullptr;
        const Entry &w = S->word[it->cur];
        it->base.sf = w.start;
        it->base.ef = w.start + w.dur - 1;
        it->base.ascr = w.score;
        it->base.lscr = 0;
        it->base.word = ssb_lexicon_wordstr(S->lx, w.id);
        return &it->base;
    }
    // ref: src/fsg_search.c:1029-1054 (fsg_seg_bp2itor; the arithmetic is done in the backtrace
    // kernel, the sf > ef clip for null transitions included)
    if (it->cur >= S->n_seg)
        return nullptr;
    const int32_t *g = &S->segs[(size_t)it->cur * 5];
    it->base.sf = g[1];
    it->base.ef = g[2];
    it->base.ascr = g[3];
    it->base.lscr = g[4];
    it->base.prob = g[3] + g[4];
    it->base.word = S->seg_word_store[it->cur].c_str();
    return &it->base;
}

ssb_seg_iter_t *search_seg_iter(ssb_search_t *s)
{
    SearchImpl *S = impl(s);
    if (S->kind == 0 ? S->word.empty() : (!fsg_have_result(S) || S->exit_bp <= 0 || S->n_seg <= 0))
        return nullptr;
    SegImpl *it = new SegImpl;
    memset(&it->base, 0, sizeof it->base);
    it->base.vt = &g_segfuncs;
    it->base.search = s;
    it->cur = 0;
    if (!seg_fill(it)) {
        delete it;
        return nullptr;
    }
    return &it->base;
}

// ---- fsg
// The search over the frames collected so far.  partial: the utterance is still running -- the
// hypothesis is what fsg_search_hyp / fsg_search_seg_iter give between two steps (best word exit
// of the last frame that has one, final state or not; ref: src/fsg_search.c:853-924, 945-960)
int fsg_run(SearchImpl *S, bool partial)
{
    const int64_t frame_off[2] = {0, S->n_frames};
    const int32_t utt_graph = 0;
    int max_seg = 256;
    // the reference's history table and segment iterator are unbounded (ref: src/fsg_history.c:
    // 129-232); here they have capacities, grown and tried again when an utterance exceeds them
    int hist_cap = std::max(4096, 8 * S->n_frames);
    for (int attempt = 0; attempt < 8; ++attempt) {
        ssb_fsg_in_t in;
        memset(&in, 0, sizeof in);
        in.n_utts = 1;
        in.feat = S->feat.data();
        in.frame_off = frame_off;
        in.n_graphs = 1;
        in.graphs = ssb_fsg_built_graph(S->fsg);
        in.utt_graph = &utt_graph;
        in.hist_cap = hist_cap;
        in.max_seg = max_seg;
        // the reference's default (compallsen = no) wherever the model allows it
        in.active_lists = ssb_model_fsg_active_ok(S->m);
        in.partial = partial ? 1 : 0;
        S->flags.assign((size_t)(S->n_sen + 31) / 32, 0u);
        S->segs.assign((size_t)max_seg * 5, 0);
        int32_t n_seg = 0, score = 0, exit_bp = 0, rv = 0, n_hist = 0;
        ssb_fsg_out_t out;
        memset(&out, 0, sizeof out);
        out.segs = S->segs.data();
        out.n_seg = &n_seg;
        out.hyp_score = &score;
        out.exit_bp = &exit_bp;
        out.utt_rv = &rv;
        out.n_hist = &n_hist;
        out.final_active = S->flags.data();
        S->carried.assign((size_t)S->n_cs * 4, 0);
        out.final_topn = S->n_cs ? S->carried.data() : nullptr;
        if (ssb_fsg_batch(S->m, &in, &out) != 0)
            return -1;
        if (rv != 0) {
            if (attempt + 1 < 8 && hist_cap < (1 << 28)) {
                hist_cap *= 2;
                continue;
            }
            set_error("fsg search: history overflow (%d entries)", in.hist_cap);
            return -1;
        }
        if (n_seg < 0 && attempt + 1 < 8) {  // segmentation longer than max_seg: ask again
            max_seg = -n_seg;
            continue;
        }
        S->n_seg = std::max(n_seg, 0);
        S->hyp_score = score;
        S->exit_bp = exit_bp;
        S->n_hist = n_hist;
        break;
    }
    const ssb_fsg_graph_t *g = ssb_fsg_built_graph(S->fsg);
    S->seg_word_store.clear();
    for (int i = 0; i < S->n_seg; ++i) {
        const int32_t link = S->segs[(size_t)i * 5];
        const int32_t wid = g->link4[(size_t)link * 4 + 3];
        // fsg_model_word_str (ref: include/soundswallower/fsg_model.h:131)
        S->seg_word_store.push_back(wid < 0 ? "(NULL)" : ssb_fsg_built_word(S->fsg, wid, nullptr));
    }
    S->result_frames = S->n_frames;
    return 0;
}

int fsg_finish(ssb_search_t *s)
{
    SearchImpl *S = impl(s);
    if (fsg_run(S, false) != 0)
        return -1;
    S->finished = true;
    return 0;
}

// hyp / seg_iter between two steps: the reference answers from its history table as it stands;
// here the frames collected so far are searched (once per frame count)
bool fsg_have_result(SearchImpl *S)
{
    if (S->finished)
        return true;
    if (S->n_frames <= 0)
        return false;
    if (S->result_frames != S->n_frames && fsg_run(S, true) != 0)
        return false;
    return true;
}

const char *fsg_hyp(ssb_search_t *s, int32_t *out_score)
{
    // ref: src/fsg_search.c:945-1026 -- words of the best final exit's backtrace, null
    // transitions and fillers left out, base strings of alternate pronunciations
    SearchImpl *S = impl(s);
    if (!fsg_have_result(S))
        return nullptr;
    if (out_score)
        *out_score = S->hyp_score;
    if (S->exit_bp <= 0)
        return nullptr;
    const ssb_fsg_graph_t *g = ssb_fsg_built_graph(S->fsg);
    std::string h;
    for (int i = 0; i < S->n_seg; ++i) {
        const int32_t wid = g->link4[(size_t)S->segs[(size_t)i * 5] * 4 + 3];
        if (wid < 0)
            continue;
        int32_t dw = -1;
        ssb_fsg_built_word(S->fsg, wid, &dw);
        if (dw < 0 || ssb_fsg_built_is_filler(S->fsg, wid))
            continue;
        if (!h.empty())
            h += ' ';
        h += basestr(S, dw);
    }
    free(S->base.hyp_str);
    S->base.hyp_str = nullptr;
    if (h.empty())
        return nullptr;
    set_hyp(S, h);
    return S->base.hyp_str;
}

int search_reinit(ssb_search_t *, void *, void *)
{
    // the aligner "does nothing, you need to make a new search for each utterance"
    // (ref: src/state_align_search.c:270-278); the grammar search would rebuild its lextree for
    // a new dictionary, which here means building a new object from a new ssb_fsg_built_t
    return 0;
}

void search_free(ssb_search_t *s)
{
    SearchImpl *S = impl(s);
    free(S->base.hyp_str);
    if (S->fsg)
        ssb_fsg_built_free(S->fsg);
    delete S;
}

// bestpath is off by default, so fsg_search_prob is 0 and there is no lattice
// (ref: src/fsg_search.c:1145-1170)
int32_t fsg_prob(ssb_search_t *) { return 0; }
void *fsg_lattice(ssb_search_t *) { return nullptr; }

ssb_searchfuncs_t g_align_funcs = {search_start, search_step, align_finish, search_reinit, search_free,
                                   nullptr, align_hyp, nullptr, search_seg_iter};
ssb_searchfuncs_t g_fsg_funcs = {search_start, search_step, fsg_finish, search_reinit, search_free,
                                 fsg_lattice, fsg_hyp, fsg_prob, search_seg_iter};

SearchImpl *new_search(int kind, const char *type, const char *name, ssb_model_t *m,
                       const ssb_lexicon_t *lx, ssb_feat_source_fn src, void *acmod)
{
    const HostModel *h = model_host(m);
    if (!h || !lx) {
        set_error("search init: NULL model or lexicon");
        return nullptr;
    }
    SearchImpl *S = new SearchImpl;
    memset(&S->base, 0, sizeof S->base);
    S->kind = kind;
    S->m = m;
    S->lx = lx;
    S->src = src;
    S->src_ctx = acmod;
    S->blk = h->blk;
    S->n_emit = h->n_emit;
    S->n_sen = h->n_sen;
    S->n_cs = h->kind == SSB_SCORER_CONT ? 0 : h->n_mgau * h->n_feat;
    S->type_s = type;
    S->name_s = name ? name : "";
    // search_module_init (ref: src/decoder.c:1276-1307)
    S->base.vt = kind == 0 ? &g_align_funcs : &g_fsg_funcs;
    S->base.type = const_cast<char *>(S->type_s.c_str());
    S->base.name = const_cast<char *>(S->name_s.c_str());
    S->base.acmod = acmod;
    S->base.dict = const_cast<ssb_lexicon_t *>(lx);
    S->base.d2p = const_cast<ssb_lexicon_t *>(lx);
    S->base.start_wid = ssb_lexicon_wordid(lx, "<s>");
    S->base.finish_wid = ssb_lexicon_wordid(lx, "</s>");
    S->base.silence_wid = ssb_lexicon_wordid(lx, "<sil>");
    S->base.n_words = ssb_lexicon_size(lx);
    return S;
}

}  // namespace

extern "C" ssb_search_t *ssb_state_align_search_init(const char *name, ssb_model_t *m,
                                                     const ssb_lexicon_t *lx, const int32_t *wids,
                                                     const int32_t *wstart, const int32_t *wdur,
                                                     int32_t n_words, ssb_feat_source_fn src,
                                                     void *acmod)
{
    SearchImpl *S = new_search(0, "state_align", name, m, lx, src, acmod);
    if (!S)
        return nullptr;
    if (n_words < 0 || (n_words > 0 && !wids)) {
        set_error("ssb_state_align_search_init: bad word list");
        search_free(&S->base);
        return nullptr;
    }
    const HostModel *h = model_host(m);
    for (int i = 0; i < n_words; ++i)
        S->word.push_back(Entry{wids[i], wstart ? wstart[i] : 0, wdur ? wdur[i] : 0, 0, -1});
    const int32_t np = ssb_chain_populate(lx, wids, n_words, nullptr, nullptr, nullptr, nullptr, 0);
    if (np < 0) {
        search_free(&S->base);
        return nullptr;
    }
    std::vector<int32_t> ci(np), parent(np);
    S->ssid.resize(np);
    S->tmat.resize(np);
    if (np > 0
        && ssb_chain_populate(lx, wids, n_words, S->ssid.data(), S->tmat.data(), ci.data(),
                              parent.data(), np) != np) {
        search_free(&S->base);
        return nullptr;
    }
    S->sf.resize(np);
    S->ef.resize(np);
    for (int i = 0; i < np; ++i) {
        const Entry &w = S->word[parent[i]];
        S->phone.push_back(Entry{ci[i], w.start, w.dur, 0, parent[i]});
        // ref: src/state_align_search.c:464-471
        S->sf[i] = w.start > 0 ? w.start : 0;
        S->ef[i] = w.dur > 0 ? w.start + w.dur : INT_MAX;
        for (int j = 0; j < h->n_emit; ++j)
            S->state.push_back(Entry{h->sseq[(size_t)S->ssid[i] * h->n_emit + j], w.start, w.dur, 0, i});
    }
    return &S->base;
}

extern "C" ssb_search_t *ssb_fsg_search_init(const char *name, ssb_model_t *m, const ssb_lexicon_t *lx,
                                             ssb_fsg_built_t *fsg, ssb_feat_source_fn src, void *acmod)
{
    if (!fsg) {
        set_error("ssb_fsg_search_init: NULL grammar");
        return nullptr;
    }
    SearchImpl *S = new_search(1, "fsg", name, m, lx, src, acmod);
    if (!S)
        return nullptr;
    S->fsg = fsg;  // consumed, like fsg_search_init consumes its fsg_model_t
    return &S->base;
}

extern "C" int ssb_search_feed(ssb_search_t *s, const float *feat, int32_t n_frames)
{
    if (!s || n_frames < 0 || (n_frames > 0 && !feat)) {
        set_error("ssb_search_feed: bad arguments");
        return -1;
    }
    SearchImpl *S = impl(s);
    S->feat.insert(S->feat.end(), feat, feat + (size_t)n_frames * S->blk);
    return (int)(S->feat.size() / S->blk);
}

extern "C" int ssb_search_final_active(const ssb_search_t *s, uint32_t *bits)
{
    if (!s || !bits) {
        set_error("ssb_search_final_active: bad arguments");
        return -1;
    }
    const SearchImpl *S = reinterpret_cast<const SearchImpl *>(s);
    const size_t nw = (size_t)(S->n_sen + 31) / 32;
    for (size_t i = 0; i < nw; ++i)
        bits[i] = i < S->flags.size() ? S->flags[i] : 0u;
    return 0;
}

extern "C" int ssb_search_set_init_active(ssb_search_t *s, const uint32_t *bits)
{
    if (!s) {
        set_error("NULL search");
        return -1;
    }
    SearchImpl *S = impl(s);
    const size_t nw = (size_t)(S->n_sen + 31) / 32;
    if (bits)
        S->flags.assign(bits, bits + nw);
    else
        S->flags.clear();
    return 0;
}

extern "C" int ssb_search_final_topn(const ssb_search_t *s, uint8_t *cw)
{
    if (!s || !cw) {
        set_error("ssb_search_final_topn: bad arguments");
        return -1;
    }
    const SearchImpl *S = reinterpret_cast<const SearchImpl *>(s);
    for (size_t i = 0; i < (size_t)S->n_cs * 4; ++i)
        cw[i] = i < S->carried.size() ? S->carried[i] : (uint8_t)(i & 3);
    return S->n_cs;
}

extern "C" int ssb_search_set_init_topn(ssb_search_t *s, const uint8_t *cw)
{
    if (!s) {
        set_error("NULL search");
        return -1;
    }
    SearchImpl *S = impl(s);
    if (cw)
        S->carried.assign(cw, cw + (size_t)S->n_cs * 4);
    else
        S->carried.clear();
    return 0;
}

extern "C" int32_t ssb_search_alignment(const ssb_search_t *s, int32_t level, int32_t *out5,
                                        int32_t max_entries)
{
    if (!s) {
        set_error("NULL search");
        return -1;
    }
    const SearchImpl *S = reinterpret_cast<const SearchImpl *>(s);
    if (S->kind != 0 || level < 0 || level > 2) {
        set_error("ssb_search_alignment: not an alignment search / bad level");
        return -1;
    }
    const std::vector<Entry> &v = level == 0 ? S->word : (level == 1 ? S->phone : S->state);
    if (out5)
        for (int i = 0; i < (int)v.size() && i < max_entries; ++i) {
            out5[i * 5] = v[i].id;
            out5[i * 5 + 1] = v[i].start;
            out5[i * 5 + 2] = v[i].dur;
            out5[i * 5 + 3] = v[i].score;
            out5[i * 5 + 4] = v[i].parent;
        }
    return (int32_t)v.size();
}
