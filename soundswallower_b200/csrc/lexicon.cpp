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
