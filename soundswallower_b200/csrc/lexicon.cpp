// lexicon.cpp -- host-side graph preparation for the chain aligner: the
// pronunciation dictionary, context-dependent phone lookup and the word -> phone-chain
// expansion that feeds ssb_align_in_t.
//
// ref: src/dict.c:74-366 (dict_init: main dictionary, filler dictionary, <s> </s> <sil>,
//      alternate pronunciations "word(2)"), src/bin_mdef.c:596-716 (bin_mdef_phone_id and its
//      back-off bin_mdef_phone_id_nearest), src/dict2pid.c:236-480 (the word-initial,
//      word-final and single-phone triphone tables, filled in dictionary order),
//      src/ps_alignment.c:133-248 (alignment_populate).
//
// The reference compresses the word-final tables into (ssid list, cimap) pairs; the value it
// reads back, rssid->ssid[rssid->cimap[rc]], is the uncompressed entry kept here.
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "model.h"

namespace ssb {
const HostModel *model_host(const ssb_model_t *m);  // api.cu
}
using namespace ssb;

namespace {

enum { POS_INTERNAL = 0, POS_BEGIN = 1, POS_END = 2, POS_SINGLE = 3, N_POS = 4 };
constexpr uint16_t BAD_SSID = 0xffff;

struct Word {
    std::string str;
    std::vector<int16_t> ph;
    int32_t basewid, alt;
};

}  // namespace

struct ssb_lexicon_s {
    const HostModel *h = nullptr;
    std::vector<Word> word;
    std::unordered_map<std::string, int32_t> id;
    int32_t filler_start = 0, filler_end = -1, startwid = -1, finishwid = -1, silwid = -1;
    int n_ci = 0;
    // [b][r][l], [b][l][r], [b][l][r]
    std::vector<uint16_t> ldiph_lc, rdiph_rc, lrdiph_rc;

    size_t at(int a, int b, int c) const { return ((size_t)a * n_ci + b) * n_ci + c; }

    // ref: src/bin_mdef.c:596-664
    int phone_id(int ci, int lc, int rc, int wpos) const
    {
        const auto &t = h->cd_tree;
        if (t.empty() || lc < 0 || rc < 0)
            return -1;
        const int sil = h->sil;
        const int ctx[4] = {wpos, ci, (sil >= 0 && h->ci_filler[lc]) ? sil : lc,
                            (sil >= 0 && h->ci_filler[rc]) ? sil : rc};
        size_t base = 0;
        int max = N_POS;
        for (int level = 0; level < 4; ++level) {
            int i = 0;
            for (; i < max; ++i)
                if (t[base + i].ctx == ctx[level])
                    break;
            if (i == max)
                return -1;
            if (t[base + i].n_down == 0)
                return t[base + i].c;
            max = t[base + i].n_down;
            base = (size_t)t[base + i].c;
        }
        return -1;
    }

    // ref: src/bin_mdef.c:666-716
    int phone_id_nearest(int b, int l, int r, int pos) const
    {
        if (l < 0 || r < 0)
            return b;
        int p = phone_id(b, l, r, pos);
        if (p >= 0)
            return p;
        for (int tp = 0; tp < N_POS; ++tp)
            if (tp != pos && (p = phone_id(b, l, r, tp)) >= 0)
                return p;
        if (h->sil >= 0) {
            int nl = l, nr = r;
            if (h->ci_filler[l] || pos == POS_BEGIN || pos == POS_SINGLE)
                nl = h->sil;
            if (h->ci_filler[r] || pos == POS_END || pos == POS_SINGLE)
                nr = h->sil;
            if (nl != l || nr != r) {
                if ((p = phone_id(b, nl, nr, pos)) >= 0)
                    return p;
                for (int tp = 0; tp < N_POS; ++tp)
                    if (tp != pos && (p = phone_id(b, nl, nr, tp)) >= 0)
                        return p;
            }
        }
        return b;
    }

    uint16_t ssid_of(int pid) const { return (uint16_t)h->ph_ssid[pid]; }

    // ref: src/dict.c:74-132
    int32_t add_word(const std::string &w, const int16_t *p, int np)
    {
        Word e;
        e.str = w;
        const int32_t wid = (int32_t)word.size();
        e.basewid = wid;
        e.alt = -1;
        const size_t len = w.size();
        if (len > 0 && w[len - 1] == ')') {  // <baseword>(...)
            size_t i = len >= 2 ? len - 2 : 0;
            while (i > 0 && w[i] != '(')
                --i;
            if (i > 0) {
                auto it = id.find(w.substr(0, i));
                if (it == id.end())
                    return -1;  // "Missing base word"
                e.basewid = it->second;
                e.alt = word[it->second].alt;
                word[it->second].alt = wid;
            }
        }
        if (id.count(w))
            return -1;  // duplicate
        id[w] = wid;
        e.ph.assign(p, p + np);
        word.push_back(std::move(e));
        return wid;
    }

    int ciphone_id(const std::string &s) const
    {
        for (int i = 0; i < (int)h->ciname.size(); ++i)
            if (h->ciname[i] == s)
                return i;
        return -1;
    }

    // ref: src/dict.c:164-242 (comment lines "##" / ";;", whitespace-separated fields)
    bool read_dict(const char *path)
    {
        FILE *fh = fopen(path, "rb");
        if (!fh) {
            set_error("failed to read dictionary from %s", path);
            return false;
        }
        std::string txt;
        char buf[65536];
        size_t n;
        while ((n = fread(buf, 1, sizeof(buf), fh)) > 0)
            txt.append(buf, n);
        fclose(fh);
        std::vector<std::string> tok;
        std::vector<int16_t> ph;
        size_t pos = 0;
        while (pos < txt.size()) {
            size_t eol = txt.find('\n', pos);
            if (eol == std::string::npos)
                eol = txt.size();
            const size_t b = pos;
            pos = eol + 1;
            if (eol - b >= 2 && (txt.compare(b, 2, "##") == 0 || txt.compare(b, 2, ";;") == 0))
                continue;
            tok.clear();
            size_t i = b;
            while (i < eol) {
                while (i < eol && (txt[i] == ' ' || txt[i] == '\t' || txt[i] == '\r'))
                    ++i;
                size_t j = i;
                while (j < eol && !(txt[j] == ' ' || txt[j] == '\t' || txt[j] == '\r'))
                    ++j;
                if (j > i)
                    tok.emplace_back(txt, i, j - i);
                i = j;
            }
            if (tok.size() < 2)
                continue;  // empty line, or a word without pronunciation (ignored)
            ph.clear();
            bool ok = true;
            for (size_t k = 1; k < tok.size() && ok; ++k) {
                const int c = ciphone_id(tok[k]);
                ok = c >= 0;  // "phone is missing in the acoustic model; word ignored"
                ph.push_back((int16_t)c);
            }
            if (ok)
                add_word(tok[0], ph.data(), (int)ph.size());
        }
        return true;
    }

    // ref: src/dict2pid.c:236-266
    void populate_lrdiph(int b)
    {
        for (int l = 0; l < n_ci; ++l)
            for (int r = 0; r < n_ci; ++r) {
                const uint16_t s = ssid_of(phone_id_nearest(b, l, r, POS_SINGLE));
                lrdiph_rc[at(b, l, r)] = s;
                if (r == h->sil)
                    ldiph_lc[at(b, r, l)] = s;
                if (l == h->sil)
                    rdiph_rc[at(b, l, r)] = s;
            }
    }

    // ref: src/dict2pid.c:372-480, in dictionary order (later words never overwrite a
    // diphone that is already marked done)
    void build_tables()
    {
        n_ci = h->n_ciphone;
        const size_t n3 = (size_t)n_ci * n_ci * n_ci;
        ldiph_lc.assign(n3, BAD_SSID);
        rdiph_rc.assign(n3, BAD_SSID);
        lrdiph_rc.assign(n3, BAD_SSID);
        std::vector<uint8_t> ldone((size_t)n_ci * n_ci, 0), rdone((size_t)n_ci * n_ci, 0),
            single(n_ci, 0);
        for (const Word &w : word) {
            const int len = (int)w.ph.size();
            if (len >= 2) {
                int b = w.ph[0], r = w.ph[1];
                if (!ldone[(size_t)b * n_ci + r]) {
                    ldone[(size_t)b * n_ci + r] = 1;
                    for (int l = 0; l < n_ci; ++l)
                        ldiph_lc[at(b, r, l)] = ssid_of(phone_id_nearest(b, l, r, POS_BEGIN));
                }
                const int l = w.ph[len - 2];
                b = w.ph[len - 1];
                if (!rdone[(size_t)b * n_ci + l]) {
                    rdone[(size_t)b * n_ci + l] = 1;
                    for (r = 0; r < n_ci; ++r)
                        rdiph_rc[at(b, l, r)] = ssid_of(phone_id_nearest(b, l, r, POS_END));
                }
            } else if (len == 1) {
                const int b = w.ph[0];
                if (!single[b]) {
                    populate_lrdiph(b);
                    single[b] = 1;
                }
            }
        }
    }
};

extern "C" ssb_lexicon_t *ssb_lexicon_load(const ssb_model_t *m, const char *dictfile,
                                           const char *fdictfile)
{
    const HostModel *h = m ? model_host(m) : nullptr;
    if (!h) {
        set_error("ssb_lexicon_load: model is NULL");
        return nullptr;
    }
    if ((int)h->ci_filler.size() != h->n_ciphone) {
        set_error("ssb_lexicon_load: the model definition carries no CI phone attributes");
        return nullptr;
    }
    auto *lx = new ssb_lexicon_s();
    lx->h = h;
    auto fail = [&](const char *msg) {
        if (msg)
            set_error("%s", msg);
        delete lx;
        return (ssb_lexicon_t *)nullptr;
    };
    if (dictfile && dictfile[0] && !lx->read_dict(dictfile))
        return fail(nullptr);
    for (const char *w : {"<s>", "</s>", "<sil>"})
        if (lx->id.count(w))
            return fail("remove <s>, </s> and <sil> from the dictionary");  // dict.c:294-305
    lx->filler_start = (int32_t)lx->word.size();
    if (fdictfile && fdictfile[0] && !lx->read_dict(fdictfile))
        return fail(nullptr);
    const int16_t sil = (int16_t)(h->sil >= 0 ? h->sil : 0);
    for (const char *w : {"<s>", "</s>", "<sil>"})
        if (!lx->id.count(w))
            lx->add_word(w, &sil, 1);
    lx->filler_end = (int32_t)lx->word.size() - 1;
    lx->startwid = lx->id["<s>"];
    lx->finishwid = lx->id["</s>"];
    lx->silwid = lx->id["<sil>"];
    lx->build_tables();
    return lx;
}

extern "C" void ssb_lexicon_free(ssb_lexicon_t *lx) { delete lx; }

extern "C" int32_t ssb_lexicon_size(const ssb_lexicon_t *lx)
{
    return lx ? (int32_t)lx->word.size() : -1;
}

extern "C" int32_t ssb_lexicon_wordid(const ssb_lexicon_t *lx, const char *word)
{
    if (!lx || !word)
        return -1;
    auto it = lx->id.find(word);
    return it == lx->id.end() ? -1 : it->second;
}

extern "C" const char *ssb_lexicon_wordstr(const ssb_lexicon_t *lx, int32_t wid)
{
    if (!lx || wid < 0 || wid >= (int32_t)lx->word.size())
        return nullptr;
    return lx->word[wid].str.c_str();
}

extern "C" int32_t ssb_lexicon_pron(const ssb_lexicon_t *lx, int32_t wid, int32_t *ciphones,
                                    int32_t max)
{
    if (!lx || wid < 0 || wid >= (int32_t)lx->word.size())
        return -1;
    const Word &w = lx->word[wid];
    for (int i = 0; i < (int)w.ph.size() && i < max && ciphones; ++i)
        ciphones[i] = w.ph[i];
    return (int32_t)w.ph.size();
}

extern "C" int32_t ssb_lexicon_basewid(const ssb_lexicon_t *lx, int32_t wid)
{
    // dict_basewid: the first pronunciation of "word(2)" is "word"
    if (!lx || wid < 0 || wid >= (int32_t)lx->word.size())
        return -1;
    return lx->word[wid].basewid;
}

extern "C" int32_t ssb_lexicon_is_filler(const ssb_lexicon_t *lx, int32_t wid)
{
    // ref: src/dict.c:381-393
    if (!lx || wid < 0 || wid >= (int32_t)lx->word.size())
        return -1;
    const int32_t w = lx->word[wid].basewid;
    if (w == lx->startwid || w == lx->finishwid)
        return 0;
    return w >= lx->filler_start && w <= lx->filler_end;
}

extern "C" int32_t ssb_chain_populate(const ssb_lexicon_t *lx, const int32_t *wids, int32_t n_words,
                                      int32_t *ssid, int32_t *tmat, int32_t *cipid,
                                      int32_t *parent, int32_t max_phones)
{
    if (!lx || n_words < 0 || (n_words > 0 && !wids)) {
        set_error("ssb_chain_populate: bad arguments");
        return -1;
    }
    const HostModel &h = *lx->h;
    const int sil = h.sil;
    int32_t n = 0;
    int lc = sil;
    auto emit = [&](int ci, uint16_t s, int w) {
        if (s == BAD_SSID) {
            set_error("word %d: no senone sequence for phone %s in this context", w,
                      h.ciname[ci].c_str());
            return false;
        }
        if (n < max_phones) {
            if (ssid)
                ssid[n] = s;
            if (tmat)
                tmat[n] = h.ph_tmat[ci];
            if (cipid)
                cipid[n] = ci;
            if (parent)
                parent[n] = w;
        }
        ++n;
        return true;
    };
    for (int i = 0; i < n_words; ++i) {
        if (wids[i] < 0 || wids[i] >= (int32_t)lx->word.size()) {
            set_error("word %d: id %d is not in the dictionary", i, wids[i]);
            return -1;
        }
        const Word &w = lx->word[wids[i]];
        const int len = (int)w.ph.size();
        if (len == 0 || sil < 0) {
            set_error("word %d has no pronunciation / the model has no SIL phone", i);
            return -1;
        }
        int rc = sil;
        if (i < n_words - 1) {
            if (wids[i + 1] < 0 || wids[i + 1] >= (int32_t)lx->word.size()
                || lx->word[wids[i + 1]].ph.empty()) {
                set_error("word %d: id %d is not in the dictionary", i + 1, wids[i + 1]);
                return -1;
            }
            rc = lx->word[wids[i + 1]].ph[0];
        }
        // first phone (ref: src/ps_alignment.c:163-183)
        const int b0 = w.ph[0];
        if (!emit(b0, len == 1 ? lx->lrdiph_rc[lx->at(b0, lc, rc)]
                               : lx->ldiph_lc[lx->at(b0, w.ph[1], lc)], i))
            return -1;
        // internal phones (ref: src/dict2pid.c:352-370)
        for (int j = 1; j < len - 1; ++j)
            if (!emit(w.ph[j], lx->ssid_of(lx->phone_id_nearest(w.ph[j], w.ph[j - 1], w.ph[j + 1],
                                                                POS_INTERNAL)), i))
                return -1;
        // last phone (ref: src/ps_alignment.c:203-220)
        if (len > 1 && !emit(w.ph[len - 1], lx->rdiph_rc[lx->at(w.ph[len - 1], w.ph[len - 2], rc)], i))
            return -1;
        lc = w.ph[len - 1];
    }
    if (n > max_phones && (ssid || tmat || cipid || parent)) {
        set_error("ssb_chain_populate: %d phones do not fit the %d provided", n, max_phones);
        return -1;
    }
    return n;
}

// ====================================================================== grammar graphs
// The alignment grammar of decoder_set_align_text (ref: src/decoder.c:685-735), augmented as
// fsg_search_init does (silence / filler self-loops on every state, alternate pronunciations;
// ref: src/fsg_search.c:83-168, src/fsg_model.c:359-449), and its lextree
// (ref: src/fsg_lextree.c:83-716), flattened to the arrays ssb_fsg_graph_t takes.
//
// Link order and node order are part of the search's tie-breaking, so they are reproduced:
// the reference keeps the arcs of a state in a 101-bucket hash table keyed by destination
// state (binary key spelt as two letters per byte, ref: src/hash_table.c:171-223), every
// bucket chain grows behind its head, every (from, to) list grows at the front; lextree nodes
// are numbered state by state, newest first (the alloc_head lists).
namespace {

struct FLink {
    int32_t from, to, logp, wid;
};

struct ArcTable {  // the hash table of one state: to_state -> list of links (front = newest)
    struct Ent {
        int32_t to;
        std::vector<int> links;  // indices into the link pool, front = most recently added
    };
    std::vector<std::vector<Ent>> bucket;
    ArcTable() : bucket(101) {}
    static unsigned hash(int32_t to)
    {
        unsigned char b[4];
        std::memcpy(b, &to, 4);
        unsigned h = 0;
        int s = 0;
        for (int i = 0; i < 4; ++i)
            for (int half = 0; half < 2; ++half) {
                const unsigned char c = half == 0 ? (unsigned char)('A' + (b[i] & 15))
                                                  : (unsigned char)('J' + (b[i] >> 4));
                h += (unsigned)c << s;
                s += 5;
                if (s >= 25)
                    s -= 24;
            }
        return h % 101u;
    }
    Ent *find(int32_t to)
    {
        for (Ent &e : bucket[hash(to)])
            if (e.to == to)
                return &e;
        return nullptr;
    }
    Ent *enter(int32_t to)
    {
        std::vector<Ent> &bk = bucket[hash(to)];
        Ent e{to, {}};
        if (bk.empty()) {
            bk.push_back(e);
            return &bk[0];
        }
        bk.insert(bk.begin() + 1, e);  // behind the head, in front of older collisions
        return &bk[1];
    }
};

struct PNode {
    int32_t ssid, tmat, logs2prob, ci_ext, leaf, next, sibling, ppos;
    uint32_t ctxt[4];
};

struct FsgWork {
    const ssb_lexicon_s *lx;
    ssb_fsg_config_t cfg;
    int n_state = 0;
    std::vector<std::string> vocab;
    std::vector<uint8_t> silword, altword;
    std::vector<FLink> pool;
    std::vector<ArcTable> trans;
    std::vector<ArcTable> ntrans;   // null transitions: one link per (from, to) (ref: src/fsg_model.c:96-140)
    std::vector<int> nulls;         // the reference's glist of null links, front = newest

    int word_add(const std::string &w)
    {
        for (size_t i = 0; i < vocab.size(); ++i)
            if (vocab[i] == w)
                return (int)i;
        vocab.push_back(w);
        silword.push_back(0);
        altword.push_back(0);
        return (int)vocab.size() - 1;
    }
    // ref: src/fsg_model.c:62-95
    void trans_add(int from, int to, int logp, int wid)
    {
        ArcTable::Ent *e = trans[from].find(to);
        if (e)
            for (int li : e->links)
                if (pool[li].wid == wid) {
                    if (pool[li].logp < logp)
                        pool[li].logp = logp;
                    return;
                }
        if (!e)
            e = trans[from].enter(to);
        pool.push_back({from, to, logp, wid});
        e->links.insert(e->links.begin(), (int)pool.size() - 1);
    }
    // ref: src/fsg_model.c:359-386
    void add_silence(const std::string &w, float prob)
    {
        LogMath lm(lx->h->cfg.logbase);
        const int wid = word_add(w);
        const int logp = (int32_t)((float)lm.log((double)prob, 0) * cfg.lw);
        silword[wid] = 1;
        for (int s = 0; s < n_state; ++s)
            trans_add(s, s, logp, wid);
    }
    // ref: src/fsg_model.c:388-449
    void add_alt(const std::string &base, const std::string &alt)
    {
        int basewid = -1;
        for (size_t i = 0; i < vocab.size(); ++i)
            if (vocab[i] == base) {
                basewid = (int)i;
                break;
            }
        if (basewid < 0)
            return;
        const int altwid = word_add(alt);
        altword[altwid] = 1;
        if (silword[basewid])
            silword[altwid] = 1;
        for (int s = 0; s < n_state; ++s)
            for (auto &bk : trans[s].bucket)
                for (ArcTable::Ent &e : bk) {
                    const std::vector<int> old = e.links;  // the walk does not see its own additions
                    for (int li : old)
                        if (pool[li].wid == basewid) {
                            pool.push_back({pool[li].from, pool[li].to, pool[li].logp, altwid});
                            e.links.insert(e.links.begin(), (int)pool.size() - 1);
                        }
                }
    }
    // ref: src/fsg_model.c:96-140 -- 1 = new link, 0 = better probability kept, -1 = nothing changed
    int null_add(int from, int to, int logp)
    {
        if (from == to)
            return -1;  // self-loop null transitions are redundant
        if ((int)ntrans.size() < n_state)
            ntrans.resize(n_state);
        ArcTable::Ent *e = ntrans[from].find(to);
        if (e) {
            FLink &l = pool[e->links[0]];
            if (l.logp < logp) {
                l.logp = logp;
                return 0;
            }
            return -1;
        }
        e = ntrans[from].enter(to);
        pool.push_back({from, to, logp, -1});
        e->links.push_back((int)pool.size() - 1);
        return 1;
    }
    // transitive closure of the null transitions, the reference's sweep order (ref: src/fsg_model.c:
    // 146-213): the list is walked from its newest entry; links found during a sweep go to the front
    void null_closure()
    {
        if ((int)ntrans.size() < n_state)
            ntrans.resize(n_state);
        bool updated;
        do {
            updated = false;
            const std::vector<int> sweep = nulls;  // (additions are seen by the next sweep only)
            for (int li1 : sweep) {
                const int from = pool[li1].from, mid = pool[li1].to;
                for (const auto &bk : ntrans[mid].bucket)
                    for (const ArcTable::Ent &e2 : bk) {
                        const int to = pool[e2.links[0]].to;
                        const int k = null_add(from, to, pool[li1].logp + pool[e2.links[0]].logp);
                        if (k >= 0) {
                            updated = true;
                            if (k > 0)
                                nulls.insert(nulls.begin(), ntrans[from].find(to)->links[0]);
                        }
                    }
            }
        } while (updated);
    }
    // links of state s in fsg_model_arcs order: word transitions, then null ones (ref :249-300)
    std::vector<int> arcs(int s) const
    {
        std::vector<int> out;
        for (const auto &bk : trans[s].bucket)
            for (const ArcTable::Ent &e : bk)
                out.insert(out.end(), e.links.begin(), e.links.end());
        if (s < (int)ntrans.size())
            for (const auto &bk : ntrans[s].bucket)
                for (const ArcTable::Ent &e : bk)
                    out.insert(out.end(), e.links.begin(), e.links.end());
        return out;
    }
};

}  // namespace

struct ssb_fsg_built_s {
    ssb_fsg_graph_t g;
    std::vector<int32_t> link4, arc_off, root, pnode8, dictwid;
    std::vector<uint8_t> link_flag;
    std::vector<uint32_t> ctxt;
    std::vector<std::string> vocab;
    std::vector<uint8_t> silword;  // fsg_model_is_filler (ref: include/soundswallower/fsg_model.h)
};

extern "C" void ssb_fsg_config_defaults(ssb_fsg_config_t *c)
{
    // ref: include/soundswallower/config_defs.h:79-159
    c->beam = 1e-48;
    c->pbeam = 1e-48;
    c->wbeam = 7e-29;
    c->lw = 6.5f;
    c->wip = 0.65f;
    c->pip = 1.0f;
    c->silprob = 0.005f;
    c->fillprob = 1e-8f;
    c->maxhmmpf = 30000;
    c->fsgusefiller = 1;
    c->fsgusealtpron = 1;
}

static ssb_fsg_built_t *fsg_finish(FsgWork &W, int start, int final);

extern "C" ssb_fsg_built_t *ssb_fsg_build_align(const ssb_lexicon_t *lx, const char *text,
                                                const ssb_fsg_config_t *cfg)
{
    if (!lx || !text) {
        set_error("ssb_fsg_build_align: bad arguments");
        return nullptr;
    }
    FsgWork W;
    W.lx = lx;
    if (cfg)
        W.cfg = *cfg;
    else
        ssb_fsg_config_defaults(&W.cfg);
    // words separated by blanks (ref: src/decoder.c:693-707)
    std::vector<std::string> words;
    const std::string t(text);
    for (size_t i = 0; i < t.size();) {
        while (i < t.size() && std::strchr(" \t\n\r", t[i]))
            ++i;
        size_t j = i;
        while (j < t.size() && !std::strchr(" \t\n\r", t[j]))
            ++j;
        if (j > i)
            words.emplace_back(t, i, j - i);
        i = j;
    }
    for (const std::string &w : words)
        if (!lx->id.count(w)) {
            set_error("Unknown word %s", w.c_str());
            return nullptr;
        }
    W.n_state = (int)words.size() + 1;
    W.trans.resize(W.n_state);
    W.ntrans.resize(W.n_state);
    for (size_t i = 0; i < words.size(); ++i)
        W.trans_add((int)i, (int)i + 1, 0, W.word_add(words[i]));
    return fsg_finish(W, 0, (int)words.size());
}

// A general grammar from its transition list, in the order the reference's reader / JSGF
// compiler would add them (ref: src/fsg_model.c:506-690 fsg_model_read_s3file, :62-140
// fsg_model_trans_add / fsg_model_null_trans_add; link order decides ties in the search).
// word[i] == NULL (or "") is a null transition; prob is the linear transition probability,
// converted like the reference does: (int32)(logmath_log(p) * lw).
static ssb_fsg_built_t *fsg_build_impl(const ssb_lexicon_t *lx, int32_t n_state, int32_t start,
                                       int32_t final, int32_t n_trans, const int32_t *from,
                                       const int32_t *to, const float *prob, const int32_t *logs2prob,
                                       const char *const *word, int32_t null_closure,
                                       const ssb_fsg_config_t *cfg);

extern "C" ssb_fsg_built_t *ssb_fsg_build(const ssb_lexicon_t *lx, int32_t n_state, int32_t start,
                                          int32_t final, int32_t n_trans, const int32_t *from,
                                          const int32_t *to, const float *prob, const char *const *word,
                                          int32_t null_closure, const ssb_fsg_config_t *cfg)
{
    return fsg_build_impl(lx, n_state, start, final, n_trans, from, to, prob, nullptr, word, null_closure, cfg);
}

// the same from an fsg_model_t's own integer scores (fsg_link_t.logs2prob, already scaled by lw)
extern "C" ssb_fsg_built_t *ssb_fsg_build_logp(const ssb_lexicon_t *lx, int32_t n_state, int32_t start,
                                               int32_t final, int32_t n_trans, const int32_t *from,
                                               const int32_t *to, const int32_t *logs2prob,
                                               const char *const *word, int32_t null_closure,
                                               const ssb_fsg_config_t *cfg)
{
    return fsg_build_impl(lx, n_state, start, final, n_trans, from, to, nullptr, logs2prob, word, null_closure, cfg);
}

static ssb_fsg_built_t *fsg_build_impl(const ssb_lexicon_t *lx, int32_t n_state, int32_t start,
                                       int32_t final, int32_t n_trans, const int32_t *from,
                                       const int32_t *to, const float *prob, const int32_t *logs2prob,
                                       const char *const *word, int32_t null_closure,
                                       const ssb_fsg_config_t *cfg)
{
    if (!lx || n_state <= 0 || start < 0 || start >= n_state || final < 0 || final >= n_state || n_trans < 0
        || (n_trans > 0 && (!from || !to || (!prob && !logs2prob) || !word))) {
        set_error("ssb_fsg_build: bad arguments");
        return nullptr;
    }
    FsgWork W;
    W.lx = lx;
    if (cfg)
        W.cfg = *cfg;
    else
        ssb_fsg_config_defaults(&W.cfg);
    W.n_state = n_state;
    W.trans.resize(n_state);
    W.ntrans.resize(n_state);
    LogMath lm(lx->h->cfg.logbase);
    for (int i = 0; i < n_trans; ++i) {
        if (from[i] < 0 || from[i] >= n_state || to[i] < 0 || to[i] >= n_state) {
            set_error("ssb_fsg_build: transition %d: state out of range", i);
            return nullptr;
        }
        if (prob && (!(prob[i] > 0.f) || prob[i] > 1.f)) {
            set_error("ssb_fsg_build: transition %d: probability %g not in (0, 1]", i, (double)prob[i]);
            return nullptr;
        }
        if (!prob && logs2prob[i] > 0) {
            set_error("ssb_fsg_build: transition %d: log probability %d above 0", i, logs2prob[i]);
            return nullptr;
        }
        const int logp = prob ? (int32_t)((float)lm.log((double)prob[i], 0) * W.cfg.lw) : logs2prob[i];
        if (word[i] && word[i][0]) {
            if (!lx->id.count(word[i])) {
                set_error("Unknown word %s", word[i]);  // (fsg_search_check_dict, ref: src/fsg_search.c:120-139)
                return nullptr;
            }
            W.trans_add(from[i], to[i], logp, W.word_add(word[i]));
        } else if (W.null_add(from[i], to[i], logp) == 1) {
            W.nulls.insert(W.nulls.begin(), W.ntrans[from[i]].find(to[i])->links[0]);
        }
    }
    if (null_closure)
        W.null_closure();
    return fsg_finish(W, start, final);
}

// fsg_search_init's augmentation + fsg_lextree_init, then the flattening
static ssb_fsg_built_t *fsg_finish(FsgWork &W, int start, int final)
{
    const ssb_lexicon_s *lx = W.lx;
    const HostModel &h = *lx->h;
    const int n_ci = h.n_ciphone, sil = h.sil, NS = W.n_state;
    if (sil < 0 || n_ci > 128) {
        set_error("the grammar search needs a SIL phone and at most 128 CI phones");
        return nullptr;
    }
    LogMath lm(h.cfg.logbase);
    if (W.cfg.fsgusefiller) {  // ref: src/fsg_search.c:83-118
        W.add_silence("<sil>", W.cfg.silprob);
        for (int32_t wid = lx->filler_start; wid < lx->filler_end; ++wid) {
            if (wid == lx->startwid || wid == lx->finishwid)
                continue;
            W.add_silence(lx->word[wid].str, W.cfg.fillprob);
        }
    }
    if (W.cfg.fsgusealtpron) {  // ref: src/fsg_search.c:141-168
        const int n_word = (int)W.vocab.size();
        for (int i = 0; i < n_word; ++i) {
            auto it = lx->id.find(W.vocab[i]);
            if (it == lx->id.end())
                continue;
            const std::string base = W.vocab[i];
            for (int32_t a = lx->word[it->second].alt; a >= 0; a = lx->word[a].alt)
                W.add_alt(base, lx->word[a].str);
        }
    }
    const int wip = (int32_t)((float)lm.log((double)W.cfg.wip, 0) * W.cfg.lw) >> 10;
    const int pip = (int32_t)((float)lm.log((double)W.cfg.pip, 0) * W.cfg.lw) >> 10;

    auto *B = new ssb_fsg_built_s();
    // ---- links in arc order
    std::vector<std::vector<int>> arcs(NS);
    std::vector<int> link_id(W.pool.size(), -1);
    B->arc_off.assign(NS + 1, 0);
    std::vector<int32_t> dictwid_of_word(W.vocab.size());
    for (size_t i = 0; i < W.vocab.size(); ++i)
        dictwid_of_word[i] = lx->id.at(W.vocab[i]);
    for (int s = 0; s < NS; ++s) {
        arcs[s] = W.arcs(s);
        B->arc_off[s] = (int32_t)(B->link4.size() / 4);
        for (int li : arcs[s]) {
            const FLink &l = W.pool[li];
            link_id[li] = (int)(B->link4.size() / 4);
            B->link4.insert(B->link4.end(), {l.from, l.to, l.logp, l.wid});
            const bool filler = l.wid >= 0 && W.silword[l.wid];
            const bool single = l.wid >= 0 && lx->word[dictwid_of_word[l.wid]].ph.size() == 1;
            B->link_flag.push_back((uint8_t)(((filler || single) ? 1 : 0) | (filler ? 2 : 0)));
        }
    }
    B->arc_off[NS] = (int32_t)(B->link4.size() / 4);
    // ---- left / right context sets of every state (ref: src/fsg_lextree.c:83-214)
    std::vector<std::vector<uint8_t>> lcb(NS, std::vector<uint8_t>(n_ci, 0)), rcb = lcb;
    for (int s = 0; s < NS; ++s)
        for (int li : arcs[s]) {
            const FLink &l = W.pool[li];
            if (l.wid < 0)
                continue;
            if (W.silword[l.wid]) {
                rcb[l.from][sil] = 1;
                lcb[l.to][sil] = 1;
            } else {
                const Word &w = lx->word[dictwid_of_word[l.wid]];
                rcb[l.from][w.ph[0]] = 1;
                lcb[l.to][w.ph.back()] = 1;
            }
        }
    std::vector<std::vector<int>> lcl(NS), rcl(NS);
    for (int s = 0; s < NS; ++s)
        lcb[s][sil] = rcb[s][sil] = 1;
    // contexts travel across null transitions, one sweep in state / arc order (ref :157-186;
    // the grammar holds the closure of its null transitions)
    for (int s = 0; s < NS; ++s)
        for (int li : arcs[s]) {
            const FLink &l = W.pool[li];
            if (l.wid >= 0)
                continue;
            for (int i = 0; i < n_ci; ++i) {
                lcb[l.to][i] |= lcb[l.from][i];
                rcb[l.from][i] |= rcb[l.to][i];
            }
        }
    for (int s = 0; s < NS; ++s) {
        for (int i = 0; i < n_ci; ++i) {
            if (lcb[s][i])
                lcl[s].push_back(i);
            if (rcb[s][i])
                rcl[s].push_back(i);
        }
    }
    // ---- lextree, state by state (ref: src/fsg_lextree.c:352-716)
    std::vector<PNode> all;       // final numbering
    B->root.assign(NS, -1);
    for (int s = 0; s < NS; ++s) {
        std::vector<PNode> nd;    // allocation order; ids are local until the state is done
        int root = -1;
        struct Shared {
            int ci, rc;
            std::vector<int> lc_nodes;  // front = newest
        };
        std::vector<Shared> shared;  // front = newest
        auto alloc = [&](int ssid, int tmat, int prob, int ci_ext, int leaf, int next, int sibling,
                         int ppos) {
            nd.push_back({ssid, tmat, prob, ci_ext, leaf, next, sibling, ppos, {0, 0, 0, 0}});
            return (int)nd.size() - 1;
        };
        auto add_ctxt = [&](int p, int c) { nd[p].ctxt[c >> 5] |= 1u << (c & 31); };
        for (int li : arcs[s]) {
            const FLink &l = W.pool[li];
            if (l.wid < 0)
                continue;  // null transitions have no HMMs (the search follows them: null_prop)
            const int32_t dw = dictwid_of_word[l.wid];
            const Word &w = lx->word[dw];
            const int pronlen = (int)w.ph.size();
            const std::vector<int> &lclist = lcl[s], &rclist = rcl[l.to];
            const int lprob = l.logp >> 10;
            if (pronlen == 1) {
                const int ci = w.ph[0];
                if (!ssb_lexicon_is_filler(lx, dw)) {
                    std::vector<int> mine;  // front = newest
                    for (int lc : lclist) {
                        const int ssid = lx->lrdiph_rc[lx->at(ci, lc, sil)];
                        int found = -1;
                        for (int p : mine)
                            if (nd[p].ssid == ssid) {
                                found = p;
                                break;
                            }
                        if (found >= 0) {
                            add_ctxt(found, lc);
                            continue;
                        }
                        const int p = alloc(ssid, h.ph_tmat[ci], lprob + wip + pip, ci, 1, -2 - li,
                                            root, 0);
                        add_ctxt(p, lc);
                        root = p;
                        mine.insert(mine.begin(), p);
                    }
                } else {
                    const int p = alloc(h.ph_ssid[ci], h.ph_tmat[ci], lprob + wip + pip, sil, 1,
                                        -2 - li, root, 0);
                    for (int k = 0; k < 4; ++k)
                        nd[p].ctxt[k] = 0xffffffffu;
                    root = p;
                }
                continue;
            }
            std::vector<int> lc_nodes;  // front = newest
            int pred = -1;
            for (int p = 0; p < pronlen; ++p) {
                const int ci = w.ph[p];
                if (p == 0) {
                    const int rc = w.ph[1];
                    int hit = -1;
                    for (size_t k = 0; k < shared.size(); ++k)
                        if (shared[k].ci == ci && shared[k].rc == rc) {
                            hit = (int)k;
                            break;
                        }
                    if (hit >= 0) {
                        lc_nodes = shared[hit].lc_nodes;
                        pred = lc_nodes.front();
                        continue;
                    }
                    // the reference's ssid -> node scan keeps the LAST node it looked at when
                    // nothing matches, so once a node exists every further left context is
                    // folded into it (ref: src/fsg_lextree.c:504-536)
                    std::vector<int> map;
                    for (int lc : lclist) {
                        const int ssid = lx->ldiph_lc[lx->at(ci, rc, lc)];
                        int pn = map.empty() ? -1 : map[0];
                        for (size_t j = 0; j < map.size(); ++j) {
                            pn = map[j];
                            if (nd[pn].ssid == ssid)
                                break;
                        }
                        if (pn < 0) {
                            pn = alloc(ssid, h.ph_tmat[w.ph[0]], wip + pip, w.ph[0], 0, -1, root, 0);
                            root = pn;
                            lc_nodes.insert(lc_nodes.begin(), pn);
                            map.push_back(pn);
                        }
                        add_ctxt(pn, lc);
                    }
                    shared.insert(shared.begin(), Shared{ci, rc, lc_nodes});
                    pred = root;
                } else if (p != pronlen - 1) {
                    const int ssid = lx->ssid_of(lx->phone_id_nearest(ci, w.ph[p - 1], w.ph[p + 1],
                                                                      POS_INTERNAL));
                    int pn = nd[pred].next;
                    const int youngest = pn;
                    while (pn >= 0 && (nd[pn].ssid != ssid || nd[pn].leaf))
                        pn = nd[pn].sibling;
                    if (pn >= 0) {
                        pred = pn;
                        continue;
                    }
                    pn = alloc(ssid, h.ph_tmat[ci], pip, ci, 0, -1, youngest, p);
                    if (p == 1)
                        for (int r : lc_nodes) {
                            pred = r;
                            nd[r].next = pn;
                        }
                    else
                        nd[pred].next = pn;
                    pred = pn;
                } else {
                    const int lc = w.ph[p - 1];
                    std::vector<std::pair<int, int>> map;  // ssid -> node
                    std::vector<int> rc_nodes;             // front = newest
                    for (int rc : rclist) {
                        const int ssid = lx->rdiph_rc[lx->at(ci, lc, rc)];
                        int pn = -1;
                        for (auto &e : map)
                            if (e.first == ssid)
                                pn = e.second;
                        if (pn < 0) {
                            pn = alloc(ssid, h.ph_tmat[ci], lprob + pip, ci, 1, -2 - li,
                                       rc_nodes.empty() ? -1 : rc_nodes.front(), p);
                            rc_nodes.insert(rc_nodes.begin(), pn);
                            map.emplace_back(ssid, pn);
                        }
                        add_ctxt(pn, rc);
                    }
                    auto hook = [&](int pr) {  // true: appended to an existing chain
                        if (nd[pr].next < 0) {
                            nd[pr].next = rc_nodes.front();
                            return false;
                        }
                        int succ = nd[pr].next;
                        while (nd[succ].sibling >= 0)
                            succ = nd[succ].sibling;
                        nd[succ].sibling = rc_nodes.front();
                        return true;
                    };
                    if (p == 1) {
                        for (int r : lc_nodes) {
                            pred = r;
                            if (hook(r))
                                break;
                        }
                    } else
                        hook(pred);
                }
            }
        }
        // number the state's nodes newest first; leaf `next` (-2 - li) becomes a link id
        const int base = (int)all.size(), n = (int)nd.size();
        auto gid = [&](int local) { return local < 0 ? -1 : base + (n - 1 - local); };
        for (int k = n - 1; k >= 0; --k) {
            PNode q = nd[k];
            q.next = q.leaf ? link_id[-2 - q.next] : gid(q.next);
            q.sibling = gid(q.sibling);
            all.push_back(q);
        }
        B->root[s] = gid(root);
    }
    for (const PNode &q : all) {
        B->pnode8.insert(B->pnode8.end(),
                         {q.ssid, q.tmat, q.logs2prob, q.ci_ext, q.leaf, q.next, q.sibling, q.ppos});
        B->ctxt.insert(B->ctxt.end(), q.ctxt, q.ctxt + 4);
    }
    B->vocab = W.vocab;
    B->silword = W.silword;
    B->dictwid = dictwid_of_word;
    ssb_fsg_graph_t &g = B->g;
    g.n_state = NS;
    g.start = start;
    g.final = final;
    g.n_link = (int32_t)B->link_flag.size();
    g.n_pnode = (int32_t)all.size();
    g.n_ciphone = n_ci;
    g.sil = sil;
    g.beam = (int32_t)lm.log(W.cfg.beam, 0) >> 10;
    g.pbeam = (int32_t)lm.log(W.cfg.pbeam, 0) >> 10;
    g.wbeam = (int32_t)lm.log(W.cfg.wbeam, 0) >> 10;
    g.maxhmmpf = W.cfg.maxhmmpf;
    g.link4 = B->link4.data();
    g.link_flag = B->link_flag.data();
    g.arc_off = B->arc_off.data();
    g.root = B->root.data();
    g.pnode8 = B->pnode8.data();
    g.ctxt = B->ctxt.data();
    return B;
}

extern "C" const ssb_fsg_graph_t *ssb_fsg_built_graph(const ssb_fsg_built_t *b)
{
    return b ? &b->g : nullptr;
}

extern "C" int32_t ssb_fsg_built_n_words(const ssb_fsg_built_t *b)
{
    return b ? (int32_t)b->vocab.size() : -1;
}

extern "C" const char *ssb_fsg_built_word(const ssb_fsg_built_t *b, int32_t fsg_wid, int32_t *dict_wid)
{
    if (!b || fsg_wid < 0 || fsg_wid >= (int32_t)b->vocab.size())
        return nullptr;
    if (dict_wid)
        *dict_wid = b->dictwid[fsg_wid];
    return b->vocab[fsg_wid].c_str();
}

extern "C" int32_t ssb_fsg_built_is_filler(const ssb_fsg_built_t *b, int32_t fsg_wid)
{
    if (!b || fsg_wid < 0 || fsg_wid >= (int32_t)b->silword.size())
        return -1;
    return b->silword[fsg_wid];
}

extern "C" void ssb_fsg_built_free(ssb_fsg_built_t *b) { delete b; }
