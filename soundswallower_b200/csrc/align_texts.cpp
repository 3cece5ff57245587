// align_texts.cpp -- `soundswallower --align` for a batch: transcripts + features in, the
// word / phone / state alignment of every utterance and the reference's JSON line out, both
// passes batched on the GPU (the result surface).
//
//   pass 1  decoder_set_align_text + search_module_forward      ref: src/decoder.c:685-735, 935-957
//           -> ssb_fsg_build_align per transcript, ssb_fsg_batch in the reference's default mode
//   pass 2  decoder_alignment                                   ref: src/decoder.c:737-798
//           -> pass 1's words (null transitions dropped) with their frame windows,
//              ssb_chain_populate (alignment_populate), ssb_align_batch starting from the
//              acmod flags pass 1 left, alignment_propagate      ref: src/ps_alignment.c:317-355
//   JSON    decoder_result_json                                  ref: src/decoder.c:1339-1593
// Host code only: every number comes out of the batched kernels.
#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "model.h"

namespace ssb {
const HostModel *model_host(const ssb_model_t *m);  // api.cu
}
using namespace ssb;

namespace {
struct Ent {
    int32_t id, start, dur, score, parent;
};
struct UttResult {
    int32_t rv = -1;  // 0 ok, -1 no hypothesis (transcript does not match), -2 second pass failed
    int32_t hyp_score = 0, n_frames = 0;
    std::string hyp;
    std::vector<Ent> seg;  // pass 1: id = fsg word id (-1 null), start = sf, dur = ef, score = ascr, parent = lscr
    std::vector<std::string> seg_word;
    std::vector<Ent> word, phone, state;
    std::string json;
    bool json_ok = false;   // json holds the line for (json_start, json_level)
    double json_start = 0;
    int json_level = -1;
};
}  // namespace

struct ssb_text_align_s {
    ssb_model_t *m = nullptr;
    const ssb_lexicon_t *lx = nullptr;
    double logbase = 1.0001;
    int frate = 100;
    std::vector<UttResult> utt;
    float kernel_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

static void hyp_item(std::string &out, double b, double d, double p, const char *t)
{
    char buf[160];
    snprintf(buf, sizeof buf, "{\"b\":%.3f,\"d\":%.3f,\"p\":%.3f,\"t\":\"", b, d, p);  // HYP_FORMAT
    out += buf;
    out += t ? t : "";
    out += '"';
}

extern "C" void ssb_text_align_free(ssb_text_align_t *r) { delete r; }

// fn(u) for every utterance, on up to 16 host threads (utterances are independent; one thread per
// 64 utterances at least, so that small batches stay on the caller's thread)
template <class F>
static void for_each_utt(int U, F fn)
{
    const int nt = std::max(1, std::min<int>({16, (int)std::thread::hardware_concurrency(), U / 64}));
    auto work = [&](int t) {
        for (int u = t; u < U; u += nt)
            fn(u);
    };
    if (nt == 1) {
        work(0);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back(work, t);
    for (auto &x : th)
        x.join();
}

extern "C" ssb_text_align_t *ssb_align_texts(ssb_model_t *m, const ssb_lexicon_t *lx, const float *feat,
                                             const int64_t *frame_off, const char *const *texts,
                                             int32_t n_utts, const ssb_fsg_config_t *cfg,
                                             int32_t align_level, int32_t frate)
{
    const HostModel *h = model_host(m);
    if (!h || !lx || n_utts < 0 || (n_utts > 0 && (!frame_off || !texts))) {
        set_error("ssb_align_texts: bad arguments");
        return nullptr;
    }
    std::unique_ptr<ssb_text_align_s> R(new ssb_text_align_s);
    R->m = m;
    R->lx = lx;
    R->logbase = h->cfg.logbase;
    R->frate = frate > 0 ? frate : 100;
    R->utt.resize(n_utts);
    const int U = n_utts, E = h->n_emit, nw = (h->n_sen + 31) / 32;
    if (U == 0)
        return R.release();

    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point t) {
        return (float)std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count();
    };
    // ---- pass 1: one alignment grammar per distinct transcript
    std::vector<ssb_fsg_built_t *> built;
    std::vector<std::string> built_text;
    std::vector<int32_t> utt_graph(U, 0);
    struct Guard {
        std::vector<ssb_fsg_built_t *> &b;
        ~Guard()
        {
            for (auto *x : b)
                ssb_fsg_built_free(x);
        }
    } guard{built};
    for (int u = 0; u < U; ++u) {
        const std::string t = texts[u] ? texts[u] : "";
        int gi = -1;
        for (size_t k = 0; k < built_text.size(); ++k)
            if (built_text[k] == t) {
                gi = (int)k;
                break;
            }
        if (gi < 0) {
            ssb_fsg_built_t *b = ssb_fsg_build_align(lx, t.c_str(), cfg);
            if (!b)  // "Unknown word ..."
                return nullptr;
            built.push_back(b);
            built_text.push_back(t);
            gi = (int)built.size() - 1;
        }
        utt_graph[u] = gi;
    }
    std::vector<ssb_fsg_graph_t> graphs;
    for (auto *b : built)
        graphs.push_back(*ssb_fsg_built_graph(b));
    int max_T = 0;
    for (int u = 0; u < U; ++u)
        max_T = std::max<int64_t>(max_T, frame_off[u + 1] - frame_off[u]);
    // a segmentation has one entry per word plus the silences / fillers between them (several
    // per gap are possible; a second round follows if the estimate is too small):
    // size the buffer from the longest transcript so that the search runs once
    size_t max_words = 0;
    for (const std::string &t : built_text) {
        size_t nwd = 0;
        bool in_word = false;
        for (char ch : t) {
            const bool sp = ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r';
            if (!sp && !in_word)
                ++nwd;
            in_word = !sp;
        }
        max_words = std::max(max_words, nwd);
    }
    int max_seg = (int)std::max<size_t>(64, 4 * max_words + 64);
    ssb_fsg_in_t fin;
    std::vector<int32_t> segs, n_seg(U), hyp_score(U), exit_bp(U), rv1(U);
    std::vector<uint32_t> flags((size_t)U * nw, 0u);
    // the scorer's carried top-N codewords after pass 1 (PTM / semi-continuous models)
    const size_t CS = h->kind == SSB_SCORER_CONT ? 0 : (size_t)h->n_mgau * h->n_feat;
    std::vector<uint8_t> carried((size_t)U * CS * 4, 0);
    float ms1[4] = {0, 0, 0, 0};
    int hist_cap = std::max(4096, 8 * max_T);
    for (int attempt = 0; attempt < 6; ++attempt) {
        memset(&fin, 0, sizeof fin);
        fin.n_utts = U;
        fin.feat = feat;
        fin.frame_off = frame_off;
        fin.n_graphs = (int32_t)graphs.size();
        fin.graphs = graphs.data();
        fin.utt_graph = utt_graph.data();
        fin.hist_cap = hist_cap;
        fin.max_seg = max_seg;
        fin.active_lists = ssb_model_fsg_active_ok(m);
        segs.assign((size_t)U * max_seg * 5, 0);
        ssb_fsg_out_t fo;
        memset(&fo, 0, sizeof fo);
        fo.segs = segs.data();
        fo.n_seg = n_seg.data();
        fo.hyp_score = hyp_score.data();
        fo.exit_bp = exit_bp.data();
        fo.utt_rv = rv1.data();
        fo.kernel_ms = ms1;
        fo.final_active = flags.data();
        fo.final_topn = CS ? carried.data() : nullptr;
        if (ssb_fsg_batch(m, &fin, &fo) != 0)
            return nullptr;
        int need = 0;  // a segmentation longer than max_seg comes back as -length: ask again
        bool hist_over = false;  // (the reference's history table is unbounded: grow ours)
        for (int u = 0; u < U; ++u) {
            need = std::max(need, -n_seg[u]);
            hist_over = hist_over || rv1[u] == -2;
        }
        if (need <= max_seg && !hist_over)
            break;
        if (attempt == 5) {
            ssb::set_error("ssb_align_texts: history / segment capacity exceeded (hist_cap %d, max_seg %d)",
                           hist_cap, max_seg);
            return nullptr;
        }
        max_seg = std::max(max_seg, need);
        if (hist_over)
            hist_cap *= 2;
    }
    for (int k = 0; k < 4; ++k)
        R->kernel_ms[k] = ms1[k];
    R->kernel_ms[4] = since(t_begin);  // wall: grammars + first pass
    const auto t_chains = std::chrono::steady_clock::now();

    // ---- decoder_alignment: words of pass 1 -> chains.  Per utterance (worker threads), then the
    // batch arrays by concatenation.
    std::vector<int64_t> phone_off(U + 1, 0);
    std::vector<int32_t> ssid, tmat, sf, ef, st_start, st_dur, st_score;
    std::vector<std::vector<int32_t>> ci_u(U), parent_u(U), wid_u(U), ssid_u(U), tmat_u(U);
    std::vector<uint8_t> failed(U, 0);
    auto chain_of = [&](int u) {
        UttResult &r = R->utt[u];
        r.n_frames = (int32_t)(frame_off[u + 1] - frame_off[u]) + 1;  // decoder_n_frames (ref :1247-1250)
        r.hyp_score = hyp_score[u];
        const ssb_fsg_built_t *b = built[utt_graph[u]];
        const ssb_fsg_graph_t &g = graphs[utt_graph[u]];
        if (rv1[u] != 0 || exit_bp[u] <= 0 || n_seg[u] <= 0)
            return;  // no hypothesis: "does not match the grammar"
        std::vector<int32_t> wstart, wdur;
        for (int i = 0; i < n_seg[u]; ++i) {
            const int32_t *sg = &segs[((size_t)u * max_seg + i) * 5];
            const int32_t fw = g.link4[(size_t)sg[0] * 4 + 3];
            r.seg.push_back(Ent{fw, sg[1], sg[2], sg[3], sg[4]});
            int32_t dw = -1;
            const char *ws = fw >= 0 ? ssb_fsg_built_word(b, fw, &dw) : "(NULL)";
            r.seg_word.push_back(ws ? ws : "");
            if (fw < 0)
                continue;  // null transitions carry no word (ref: src/decoder.c:757-766)
            if (!ssb_fsg_built_is_filler(b, fw)) {
                if (!r.hyp.empty())
                    r.hyp += ' ';
                const int32_t base = ssb_lexicon_basewid(lx, dw);
                r.hyp += ssb_lexicon_wordstr(lx, base >= 0 ? base : dw);
            }
            wid_u[u].push_back(dw);
            wstart.push_back(sg[1]);
            wdur.push_back(sg[2] - sg[1] + 1);
        }
        r.rv = 0;
        if (!align_level || wid_u[u].empty())
            return;
        const int nwd = (int)wid_u[u].size();
        const int32_t np = ssb_chain_populate(lx, wid_u[u].data(), nwd, nullptr, nullptr, nullptr, nullptr, 0);
        if (np < 0) {
            failed[u] = 1;
            return;
        }
        ssid_u[u].resize(np);
        tmat_u[u].resize(np);
        ci_u[u].resize(np);
        parent_u[u].resize(np);
        if (np > 0
            && ssb_chain_populate(lx, wid_u[u].data(), nwd, ssid_u[u].data(), tmat_u[u].data(), ci_u[u].data(),
                                  parent_u[u].data(), np) != np) {
            failed[u] = 1;
            return;
        }
        for (int i = 0; i < nwd; ++i)
            r.word.push_back(Ent{wid_u[u][i], wstart[i], wdur[i], 0, -1});
        r.phone.reserve(np);
        r.state.reserve((size_t)np * E);
        for (int i = 0; i < np; ++i) {
            const Ent &w = r.word[parent_u[u][i]];
            r.phone.push_back(Ent{ci_u[u][i], w.start, w.dur, 0, parent_u[u][i]});
            for (int j = 0; j < E; ++j)
                // what alignment_populate leaves in the state entries (ref: src/ps_alignment.c:237-240)
                r.state.push_back(Ent{h->sseq[(size_t)ssid_u[u][i] * E + j], w.start, w.dur, 0, i});
        }
    };
    for_each_utt(U, chain_of);
    {
        size_t np_total = 0;
        for (int u = 0; u < U; ++u) {
            if (failed[u])
                return nullptr;  // (ssb_chain_populate has set the error)
            np_total += ssid_u[u].size();
        }
        ssid.reserve(np_total);
        tmat.reserve(np_total);
        sf.reserve(np_total);
        ef.reserve(np_total);
        st_start.reserve(np_total * E);
        st_dur.reserve(np_total * E);
        st_score.assign(np_total * E, 0);
        for (int u = 0; u < U; ++u) {
            const UttResult &r = R->utt[u];
            const size_t np = ssid_u[u].size();
            ssid.insert(ssid.end(), ssid_u[u].begin(), ssid_u[u].end());
            tmat.insert(tmat.end(), tmat_u[u].begin(), tmat_u[u].end());
            for (size_t i = 0; i < np; ++i) {
                const Ent &w = r.word[parent_u[u][i]];
                // ref: src/state_align_search.c:464-471
                sf.push_back(w.start > 0 ? w.start : 0);
                ef.push_back(w.dur > 0 ? w.start + w.dur : INT_MAX);
                for (int j = 0; j < E; ++j) {
                    st_start.push_back(w.start);
                    st_dur.push_back(w.dur);
                }
            }
            phone_off[u + 1] = phone_off[u] + (int64_t)np;
        }
    }

    R->kernel_ms[5] = since(t_chains);  // wall: chains on the host
    const auto t_p2 = std::chrono::steady_clock::now();
    // ---- pass 2
    if (align_level && phone_off[U] > 0) {
        std::vector<int32_t> rv2(U, 0), best(U, 0), ren(U, 0);
        ssb_align_in_t ain;
        memset(&ain, 0, sizeof ain);
        ain.n_utts = U;
        ain.feat = feat;
        ain.frame_off = frame_off;
        ain.phone_off = phone_off.data();
        ain.ssid = ssid.data();
        ain.tmat = tmat.data();
        ain.sf = sf.data();
        ain.ef = ef.data();
        ain.init_active = fin.active_lists ? flags.data() : nullptr;  // the acmod both passes share
        ain.init_topn = CS ? carried.data() : nullptr;
        ssb_align_out_t ao;
        memset(&ao, 0, sizeof ao);
        ao.st_start = st_start.data();
        ao.st_dur = st_dur.data();
        ao.st_score = st_score.data();
        ao.utt_rv = rv2.data();
        ao.utt_best = best.data();
        ao.utt_renorm = ren.data();
        if (ssb_align_batch(m, &ain, &ao) != 0)
            return nullptr;
        auto finish_of = [&](int u) {
            UttResult &r = R->utt[u];
            if (r.rv != 0 || r.state.empty())
                return;
            if (rv2[u] != 0) {
                r.rv = -2;  // "Failed to reach final state in alignment"
                return;
            }
            const size_t s0 = (size_t)phone_off[u] * E;
            for (size_t i = 0; i < r.state.size(); ++i) {
                r.state[i].start = st_start[s0 + i];
                r.state[i].dur = st_dur[s0 + i];
                r.state[i].score = st_score[s0 + i];
            }
            int last = -1;  // alignment_propagate (ref: src/ps_alignment.c:317-355)
            for (const Ent &e : r.state) {
                Ent &p = r.phone[e.parent];
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
            for (const Ent &p : r.phone) {
                Ent &w = r.word[p.parent];
                if (p.parent != last) {
                    w.start = p.start;
                    w.dur = 0;
                    w.score = 0;
                }
                w.dur += p.dur;
                w.score += p.score;
                last = p.parent;
            }
        };
        for_each_utt(U, finish_of);
    }
    R->kernel_ms[6] = since(t_p2);      // wall: second pass + propagate
    R->kernel_ms[7] = since(t_begin);
    return R.release();
}

extern "C" int32_t ssb_text_align_status(const ssb_text_align_t *r, int32_t u, int32_t *hyp_score,
                                         int32_t *n_frames)
{
    if (!r || u < 0 || u >= (int32_t)r->utt.size())
        return INT32_MIN;
    if (hyp_score)
        *hyp_score = r->utt[u].hyp_score;
    if (n_frames)
        *n_frames = r->utt[u].n_frames;
    return r->utt[u].rv;
}

extern "C" const char *ssb_text_align_hyp(const ssb_text_align_t *r, int32_t u)
{
    if (!r || u < 0 || u >= (int32_t)r->utt.size() || r->utt[u].rv == -1 || r->utt[u].hyp.empty())
        return nullptr;
    return r->utt[u].hyp.c_str();
}

extern "C" int32_t ssb_text_align_entries(const ssb_text_align_t *r, int32_t u, int32_t level,
                                          int32_t *out5, int32_t max_entries)
{
    if (!r || u < 0 || u >= (int32_t)r->utt.size() || level < 0 || level > 3) {
        set_error("ssb_text_align_entries: bad arguments");
        return -1;
    }
    const UttResult &x = r->utt[u];
    const std::vector<Ent> &v = level == 0 ? x.word : (level == 1 ? x.phone : (level == 2 ? x.state : x.seg));
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

extern "C" const char *ssb_text_align_json(ssb_text_align_t *r, int32_t u, double start, int32_t align_level)
{
    if (!r || u < 0 || u >= (int32_t)r->utt.size()) {
        set_error("ssb_text_align_json: bad arguments");
        return nullptr;
    }
    UttResult &x = r->utt[u];
    if (x.rv == -1 || (align_level && (x.rv != 0 || x.word.empty())))
        return nullptr;  // decoder_result_json returns NULL without an alignment (ref :1511-1515)
    if (x.json_ok && x.json_start == start && x.json_level == align_level)
        return x.json.c_str();  // rendered by ssb_text_align_render
    const HostModel *h = model_host(r->m);
    const double base = r->logbase, fr = r->frate;
    auto P = [&](int32_t score) { return pow(base, (double)score); };  // logmath_exp
    std::string &o = x.json;
    o.clear();
    hyp_item(o, start, x.n_frames / fr, P(0), x.hyp.c_str());  // fsg_search_prob = 0 (no bestpath)
    o += ",\"w\":[";
    bool first = true;
    if (align_level) {
        size_t pi = 0;
        for (size_t wi = 0; wi < x.word.size(); ++wi) {
            const Ent &w = x.word[wi];
            if (!first)
                o += ',';
            first = false;
            hyp_item(o, start + w.start / fr, w.dur / fr, P(w.score), ssb_lexicon_wordstr(r->lx, w.id));
            o += ",\"w\":[";
            bool pf = true;
            for (; pi < x.phone.size() && x.phone[pi].parent == (int32_t)wi; ++pi) {
                const Ent &p = x.phone[pi];
                if (!pf)
                    o += ',';
                pf = false;
                hyp_item(o, start + p.start / fr, p.dur / fr, P(p.score), ssb_model_ciphone_str(r->m, p.id));
                if (align_level > 1) {
                    o += ",\"w\":[";
                    for (int j = 0; j < h->n_emit; ++j) {
                        const Ent &s = x.state[pi * h->n_emit + j];
                        char name[16];
                        snprintf(name, sizeof name, "%u", (unsigned)s.id);  // alignment_iter_name of a state
                        if (j)
                            o += ',';
                        hyp_item(o, start + s.start / fr, s.dur / fr, P(s.score), name);
                        o += '}';
                    }
                    o += ']';
                }
                o += '}';
            }
            o += "]}";
        }
    } else {
        for (size_t i = 0; i < x.seg.size(); ++i) {
            const Ent &s = x.seg[i];  // start = sf, dur = ef, score = ascr, parent = lscr
            if (!first)
                o += ',';
            first = false;
            hyp_item(o, start + s.start / fr, (s.dur + 1 - s.start) / fr, P(s.score + s.parent),
                     x.seg_word[i].c_str());
            o += '}';
        }
    }
    o += "]}\n";
    x.json_ok = true;
    x.json_start = start;
    x.json_level = align_level;
    return o.c_str();
}

extern "C" int ssb_text_align_render(ssb_text_align_t *r, double start, int32_t align_level)
{
    if (!r) {
        set_error("NULL result");
        return -1;
    }
    const int U = (int)r->utt.size();
    for_each_utt(U, [&](int u) { ssb_text_align_json(r, u, start, align_level); });
    return 0;
}

extern "C" int ssb_text_align_kernel_ms(const ssb_text_align_t *r, float *ms8)
{
    if (!r || !ms8)
        return -1;
    for (int k = 0; k < 8; ++k)
        ms8[k] = r->kernel_ms[k];
    return 0;
}
