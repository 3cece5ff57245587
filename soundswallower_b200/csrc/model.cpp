// model.cpp -- host loaders for the on-disk acoustic model formats.
//
// Formats: S3 binary containers ("means", "variances",
// "transition_matrices", optional "mixture_weights"), "sendump", binary "mdef".
// Behaviour follows the reference loaders cited at each function; the code is
// written from the format descriptions, array-based, with no mmap retention.
#include "model.h"

#include <sys/stat.h>

#include <algorithm>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <fstream>

namespace ssb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

const char *last_error() { return g_err; }

static thread_local int g_launches = 0;
void note_launch() { ++g_launches; }
int launch_count(bool reset)
{
    int n = g_launches;
    if (reset)
        g_launches = 0;
    return n;
}

// ---------------------------------------------------------------- log math
LogMath::LogMath(double b) : base(b), inv_log_base(1.0 / std::log(b)) {}

// ref: src/logmath.c:283-290
int32_t LogMath::log(double p, int shift) const
{
    if (p <= 0)
        return zero(shift);
    return (int32_t)(std::log(p) * inv_log_base) >> shift;
}

// ref: src/logmath.c:298-302
int32_t LogMath::ln_to_log(double lnp, int shift) const
{
    return (int32_t)(lnp * inv_log_base) >> shift;
}

// ref: src/logmath.c:89-160.  Slot (i >> shift) keeps the first non-zero value written
// to it, i.e. the one for the smallest difference that maps there.
bool LogMath::add_table8(int shift, uint8_t out[256]) const
{
    uint32_t span = (uint32_t)(std::log(2.0) / std::log(base) + 0.5) >> shift;
    if (span >= 256)
        return false;
    std::memset(out, 0, 256);
    double ratio = 1.0;  // base^(y-x), y-x = 0, -1, -2, ...
    for (uint32_t diff = 0;; ++diff) {
        double v = std::log(1.0 + ratio) * inv_log_base;
        int32_t q = (int32_t)(v + 0.5 * (1 << shift)) >> shift;
        uint32_t slot = diff >> shift;
        if (slot < 256 && out[slot] == 0)
            out[slot] = (uint8_t)q;
        if (q <= 0)
            break;
        ratio /= base;
    }
    return true;
}

// ---------------------------------------------------------------- byte reader
namespace {

struct Blob {
    std::vector<uint8_t> bytes;
    size_t at = 0;
    bool swapped = false;
    bool summing = false;
    uint32_t sum = 0;

    bool open(const std::string &path)
    {
        std::ifstream in(path, std::ios::binary | std::ios::ate);
        if (!in)
            return false;
        std::streamsize n = in.tellg();
        in.seekg(0);
        bytes.resize((size_t)n);
        return n == 0 || (bool)in.read((char *)bytes.data(), n);
    }
    static uint32_t flip(uint32_t v) { return __builtin_bswap32(v); }
    size_t left() const { return bytes.size() - at; }

    // 32-bit words with optional swap and the S3 running checksum
    // (ref: src/s3file.c:369-397: sum = rotl(sum, 20) + word).
    bool words(void *dst, size_t n)
    {
        if (left() < 4 * n)
            return false;
        uint32_t *o = (uint32_t *)dst;
        std::memcpy(o, bytes.data() + at, 4 * n);
        at += 4 * n;
        for (size_t i = 0; i < n; ++i) {
            if (swapped)
                o[i] = flip(o[i]);
            if (summing)
                sum = ((sum << 20) | (sum >> 12)) + o[i];
        }
        return true;
    }
    bool i32(int32_t &v) { return words(&v, 1); }

    // "s3\n" + "key value\n"* + "endhdr\n" + 0x11223344 (ref: src/s3file.c:210-326)
    bool s3_header()
    {
        if (bytes.size() < 3 || std::memcmp(bytes.data(), "s3\n", 3) != 0)
            return false;
        at = 3;
        bool want_sum = false;
        for (;;) {
            size_t b = at;
            while (at < bytes.size() && bytes[at] != '\n')
                ++at;
            if (at >= bytes.size())
                return false;
            std::string line((const char *)bytes.data() + b, at - b);
            ++at;
            size_t k = line.find_first_not_of(" \t");
            if (k == std::string::npos)
                return false;
            if (line[k] == '#')
                continue;
            if (line.compare(k, 6, "endhdr") == 0)
                break;
            if (line.compare(k, 7, "chksum0") == 0)
                want_sum = true;
        }
        uint32_t magic;
        if (!words(&magic, 1))
            return false;
        if (magic != 0x11223344u) {
            if (flip(magic) != 0x11223344u)
                return false;
            swapped = true;
        }
        summing = want_sum;
        sum = 0;
        return true;
    }
    // trailing checksum word (ref: src/s3file.c:551-570)
    bool s3_trailer()
    {
        if (!summing)
            return true;
        uint32_t have = sum, want;
        summing = false;
        return words(&want, 1) && want == have;
    }
};

}  // namespace

// ---------------------------------------------------------------- gaussians
// payload: n_mgau n_feat n_density veclen[n_feat] n data[n]  (ref: src/ms_gauden.c:126-195)
static bool read_gauden(const std::string &path, int32_t dims[3], int32_t *featlen,
                        std::vector<float> &data)
{
    Blob f;
    if (!f.open(path) || !f.s3_header()) {
        set_error("cannot read S3 file %s", path.c_str());
        return false;
    }
    int32_t n;
    if (!f.i32(dims[0]) || !f.i32(dims[1]) || !f.i32(dims[2]) || dims[1] < 1
        || dims[1] > SSB_MAX_FEAT || !f.words(featlen, dims[1]) || !f.i32(n)) {
        set_error("%s: bad gaussian header (at most %d streams supported)", path.c_str(),
                  SSB_MAX_FEAT);
        return false;
    }
    int64_t per = 0;
    for (int i = 0; i < dims[1]; ++i)
        per += featlen[i];
    if ((int64_t)n != (int64_t)dims[0] * dims[2] * per) {
        set_error("%s: #floats %d does not match dimensions", path.c_str(), n);
        return false;
    }
    data.resize(n);
    if (!f.words(data.data(), n) || !f.s3_trailer()) {
        set_error("%s: truncated or checksum mismatch", path.c_str());
        return false;
    }
    return true;
}

// ref: src/ms_gauden.c:217-258.  Floor, then det += (f32)log_b(1/sqrt(2 pi s2)) and
// var <- (f32)(int)(1/(2 s2) / ln b), both through double then int truncation.
static void precompute_gauden(HostModel &m)
{
    LogMath lm(m.cfg.logbase);
    const float vfloor = m.cfg.varfloor;
    m.det.assign((size_t)m.n_mgau * m.n_feat * m.n_density, 0.f);
    for (int c = 0; c < m.n_mgau; ++c)
        for (int f = 0; f < m.n_feat; ++f) {
            const int L = m.featlen[f];
            float *v = m.var.data() + m.gau_off[c * m.n_feat + f];
            float *dt = m.det.data() + (size_t)(c * m.n_feat + f) * m.n_density;
            for (int k = 0; k < m.n_density; ++k) {
                float acc = 0.f;
                for (int i = 0; i < L; ++i) {
                    float &s2 = v[k * L + i];
                    if (s2 < vfloor)
                        s2 = vfloor;
                    acc += (float)lm.log(1.0 / std::sqrt(s2 * 2.0 * M_PI), 0);
                    s2 = (float)lm.ln_to_log(1.0 / (s2 * 2.0), 0);
                }
                dt[k] = acc;
            }
        }
}

// ---------------------------------------------------------------- mdef
// ref: src/bin_mdef.c:333-520
static bool read_mdef(HostModel &m, const std::string &path)
{
    Blob f;
    if (!f.open(path) || f.bytes.size() < 8) {
        set_error("cannot read %s", path.c_str());
        return false;
    }
    if (std::memcmp(f.bytes.data(), "BMDF", 4) == 0)
        f.swapped = false;
    else if (std::memcmp(f.bytes.data(), "FDMB", 4) == 0)
        f.swapped = true;
    else {
        set_error("%s: not a binary mdef (text mdef is not supported)", path.c_str());
        return false;
    }
    f.at = 4;
    int32_t version, fmt_len, h[10];
    if (!f.i32(version) || version > 1 || !f.i32(fmt_len) || f.left() < (size_t)fmt_len) {
        set_error("%s: bad version/format block", path.c_str());
        return false;
    }
    f.at += fmt_len;
    if (!f.words(h, 10)) {
        set_error("%s: truncated header", path.c_str());
        return false;
    }
    m.n_ciphone = h[0];
    m.n_phone = h[1];
    m.n_emit = h[2];
    m.n_ci_sen = h[3];
    int32_t n_sen = h[4], n_sseq = h[6], n_cd_tree = h[8];
    m.sil = h[9];
    if (m.n_emit <= 0) {
        set_error("%s: heterogeneous topologies are not supported", path.c_str());
        return false;
    }
    // header sanity: every count indexes a table below
    if (m.n_ciphone <= 0 || m.n_ciphone > 255 || m.n_phone < m.n_ciphone || m.n_emit > 8 || n_sen <= 0
        || n_sen > 65535 || n_sseq <= 0 || n_cd_tree < 0 || m.n_ci_sen < 0 || m.n_ci_sen > n_sen) {
        set_error("%s: implausible header (ciphones %d, phones %d, states %d, senones %d, sseq %d, tree %d)",
                  path.c_str(), m.n_ciphone, m.n_phone, m.n_emit, n_sen, n_sseq, n_cd_tree);
        return false;
    }
    size_t names = f.at, p = f.at;
    for (int i = 0; i < m.n_ciphone; ++i) {
        size_t e = p;
        while (e < f.bytes.size() && f.bytes[e])
            ++e;
        if (e >= f.bytes.size())
            return false;
        m.ciname.emplace_back((const char *)f.bytes.data() + p, e - p);
        p = e + 1;
    }
    p = names + (((p - names) + 3) & ~(size_t)3);
    if (f.bytes.size() < p + (size_t)n_cd_tree * 8)
        return false;
    m.cd_tree.resize(n_cd_tree);  // cd_tree_t {int16 ctx, int16 n_down, int32 pid|down}
    for (int i = 0; i < n_cd_tree; ++i, p += 8) {
        uint16_t a, b;
        uint32_t c;
        std::memcpy(&a, &f.bytes[p], 2);
        std::memcpy(&b, &f.bytes[p + 2], 2);
        std::memcpy(&c, &f.bytes[p + 4], 4);
        if (f.swapped) {
            a = (uint16_t)((a >> 8) | (a << 8));
            b = (uint16_t)((b >> 8) | (b << 8));
            c = Blob::flip(c);
        }
        m.cd_tree[i] = {(int16_t)a, (int16_t)b, (int32_t)c};
    }
    if (f.bytes.size() < p + (size_t)m.n_phone * 12 + 4) {
        set_error("%s: truncated phone table", path.c_str());
        return false;
    }
    m.ph_ssid.resize(m.n_phone);
    m.ph_tmat.resize(m.n_phone);
    m.ph_ci.resize(m.n_phone);
    for (int i = 0; i < m.n_phone; ++i, p += 12) {
        uint32_t a, b;
        std::memcpy(&a, &f.bytes[p], 4);
        std::memcpy(&b, &f.bytes[p + 4], 4);
        m.ph_ssid[i] = (int32_t)(f.swapped ? Blob::flip(a) : a);
        m.ph_tmat[i] = (int32_t)(f.swapped ? Blob::flip(b) : b);
        // info bytes: CI entries {reserved, filler}; CD entries {wpos, base ci, lc, rc}
        m.ph_ci[i] = i < m.n_ciphone ? i : f.bytes[p + 9];
        if (i < m.n_ciphone)
            m.ci_filler.push_back(f.bytes[p + 8]);
    }
    f.at = p;
    int32_t sseq_size;
    if (!f.i32(sseq_size) || sseq_size != n_sseq * m.n_emit
        || f.left() < (size_t)sseq_size * 2) {
        set_error("%s: bad senone-sequence table", path.c_str());
        return false;
    }
    m.n_sseq = n_sseq;
    m.sseq.resize(sseq_size);
    std::memcpy(m.sseq.data(), &f.bytes[f.at], (size_t)sseq_size * 2);
    if (f.swapped)
        for (auto &s : m.sseq)
            s = (uint16_t)((s >> 8) | (s << 8));
    m.n_sen = n_sen;
    for (int i = 0; i < m.n_phone; ++i)
        if (m.ph_ssid[i] < 0 || m.ph_ssid[i] >= n_sseq || m.ph_tmat[i] < 0 || m.ph_ci[i] >= m.n_ciphone) {
            set_error("%s: phone %d: ssid %d / tmat %d / ci %d out of range", path.c_str(), i, m.ph_ssid[i],
                      m.ph_tmat[i], m.ph_ci[i]);
            return false;
        }
    // senone -> CI phone of the first phone (in id order) that uses it
    // (ref: src/bin_mdef.c:470-516); this is the PTM senone->codebook map.
    std::vector<int> owner(n_sen, -1);
    for (int i = 0; i < m.n_phone; ++i)
        for (int j = 0; j < m.n_emit; ++j) {
            int s = m.sseq[(size_t)m.ph_ssid[i] * m.n_emit + j];
            if (s < n_sen && owner[s] < 0)
                owner[s] = m.ph_ci[i];
        }
    m.sen2cb.resize(n_sen);
    for (int s = 0; s < n_sen; ++s) {
        if (owner[s] < 0) {
            // the reference leaves such a senone on codebook -1 and never scores it (no phone
            // lists it); it cannot be uploaded as an index
            set_error("%s: senone %d is used by no phone", path.c_str(), s);
            return false;
        }
        m.sen2cb[s] = (uint8_t)owner[s];
    }
    for (int i = 0; i < m.n_ciphone; ++i)
        if (m.ciname[i] == "SIL")
            m.sil = i;
    return true;
}

// ---------------------------------------------------------------- sendump
// ref: src/ptm_mgau.c:456-609 (layout) and :375-378 (4-bit lookup, taken literally:
// the nibble is chosen by the low bit of the packed byte itself).
static bool read_sendump(HostModel &m, const std::string &path)
{
    Blob f;
    if (!f.open(path))
        return false;  // caller falls back to mixture_weights
    int32_t n;
    auto skip_str = [&](bool must_nul) {
        if (n < 0 || f.left() < (size_t)n || (must_nul && (n == 0 || f.bytes[f.at + n - 1])))
            return false;
        f.at += n;
        return true;
    };
    if (!f.i32(n))
        return false;
    if (n < 1 || n > 999) {
        n = (int32_t)Blob::flip((uint32_t)n);
        if (n < 1 || n > 999) {
            set_error("%s: title length out of range", path.c_str());
            return false;
        }
        f.swapped = true;
    }
    if (!skip_str(true) || !f.i32(n) || !skip_str(true)) {
        set_error("%s: bad title/header strings", path.c_str());
        return false;
    }
    int n_feat = m.n_feat, n_density = m.n_density, n_sen = m.n_sen, n_clust = 0, n_bits = 8;
    for (;;) {
        if (!f.i32(n))
            return false;
        if (n == 0)
            break;
        if (n < 0 || f.left() < (size_t)n)
            return false;
        std::string s((const char *)&f.bytes[f.at], strnlen((const char *)&f.bytes[f.at], n));
        f.at += n;
        int v;
        if (sscanf(s.c_str(), "feature_count %d", &v) == 1)
            n_feat = v;
        else if (sscanf(s.c_str(), "mixture_count %d", &v) == 1)
            n_density = v;
        else if (sscanf(s.c_str(), "model_count %d", &v) == 1)
            n_sen = v;
        else if (sscanf(s.c_str(), "cluster_count %d", &v) == 1)
            n_clust = v;
        else if (sscanf(s.c_str(), "cluster_bits %d", &v) == 1)
            n_bits = v;
    }
    int32_t rows = n_density, cols = n_sen;
    if (n_clust == 0 && (!f.i32(rows) || !f.i32(cols)))
        return false;
    if (n_feat != m.n_feat || n_density != m.n_density || n_sen != m.n_sen) {
        set_error("%s: %d x %d x %d does not match the model (%d x %d x %d)", path.c_str(),
                  n_feat, n_density, n_sen, m.n_feat, m.n_density, m.n_sen);
        return false;
    }
    if (!(n_clust == 0 || n_clust == 15 || n_clust == 16) || !(n_bits == 8 || n_bits == 4)) {
        set_error("%s: cluster count %d / bits %d not supported", path.c_str(), n_clust, n_bits);
        return false;
    }
    if (n_clust == 15)
        n_clust = 16;
    const uint8_t *book = nullptr;
    if (n_clust) {
        if (f.left() < (size_t)n_clust)
            return false;
        book = &f.bytes[f.at];
        f.at += n_clust;
        // the reference's unrolled 4-bit loops keep weight + score in uint8 tables
        // (ref: src/s2_semi_mgau.c:445-731): entries above MAX_NEG_MIXW would wrap there
        for (int i = 0; i < n_clust; ++i)
            if (m.kind == SSB_SCORER_SEMI && book[i] > 159) {
                set_error("%s: cluster codebook entry %d > 159 overflows the reference's 8-bit "
                          "weight tables", path.c_str(), book[i]);
                return false;
            }
    }
    if (rows <= 0 || cols < n_sen) {
        set_error("%s: weight rows of %d columns for %d senones", path.c_str(), cols, n_sen);
        return false;
    }
    const size_t stride = n_bits == 4 ? (size_t)(cols + 1) / 2 : (size_t)cols;
    if (f.left() < stride * rows * n_feat) {
        set_error("%s: truncated weight rows", path.c_str());
        return false;
    }
    m.mixw.assign((size_t)n_feat * n_density * n_sen, 0);
    for (int ft = 0; ft < n_feat; ++ft)
        for (int r = 0; r < rows; ++r, f.at += stride) {
            if (r >= n_density)
                continue;
            const uint8_t *src = &f.bytes[f.at];
            uint8_t *dst = &m.mixw[((size_t)ft * n_density + r) * n_sen];
            if (!book) {
                std::memcpy(dst, src, n_sen);
                continue;
            }
            for (int s = 0; s < n_sen; ++s) {
                // ptm_mgau picks the nibble by the low bit of the packed byte itself
                // (ref: src/ptm_mgau.c:375-378), s2_semi by the senone's parity
                // (ref: src/s2_semi_mgau.c:733-757, 795-824)
                const int b = src[s / 2];
                const bool high = m.kind == SSB_SCORER_SEMI ? (s & 1) : (b & 1);
                dst[s] = book[high ? b >> 4 : b & 0x0f];
            }
        }
    return true;
}

// ref: src/ptm_mgau.c:611-692, src/vector.c (sum_norm / floor)
static bool read_mixw_float(HostModel &m, const std::string &path)
{
    Blob f;
    if (!f.open(path) || !f.s3_header()) {
        set_error("neither sendump nor mixture_weights readable in model directory");
        return false;
    }
    int32_t n_sen, n_feat, n_comp, n;
    if (!f.i32(n_sen) || !f.i32(n_feat) || !f.i32(n_comp) || !f.i32(n) || n_feat != m.n_feat
        || n_comp != m.n_density || n_sen != m.n_sen || n != n_sen * n_feat * n_comp) {
        set_error("%s: dimensions do not match the model", path.c_str());
        return false;
    }
    LogMath lm(m.cfg.logbase);
    std::vector<float> pdf(n_comp);
    m.mixw.assign((size_t)n_feat * n_comp * n_sen, 0);
    auto sum_norm = [&]() {
        double s = 0.0;
        for (float x : pdf)
            s += x;
        if (s != 0.0) {
            double r = 1.0 / s;
            for (float &x : pdf)
                x = (float)(x * r);
        }
    };
    for (int s = 0; s < n_sen; ++s)
        for (int ft = 0; ft < n_feat; ++ft) {
            if (!f.words(pdf.data(), n_comp))
                return false;
            sum_norm();
            for (float &x : pdf)
                if (x < m.cfg.mixwfloor)
                    x = (float)m.cfg.mixwfloor;
            sum_norm();
            for (int c = 0; c < n_comp; ++c) {
                int32_t q = -lm.log(pdf[c], 10);
                if (q > 159 || q < 0)
                    q = 159;
                m.mixw[((size_t)ft * n_comp + c) * n_sen + s] = (uint8_t)q;
            }
        }
    return f.s3_trailer();
}

// ref: src/ms_senone.c:103-190 (senone_mixw_read): the continuous scorer's own quantisation
// (shift-0 log, +511 before the >> 10, clamp 255), kept [sen][feat][density]
static bool read_mixw_cont(HostModel &m, const std::string &path)
{
    Blob f;
    if (!f.open(path) || !f.s3_header()) {
        set_error("cannot read S3 file %s", path.c_str());
        return false;
    }
    int32_t n_sen, n_feat, n_cw, n;
    if (!f.i32(n_sen) || !f.i32(n_feat) || !f.i32(n_cw) || !f.i32(n) || n_feat != m.n_feat
        || n_cw != m.n_density || n_sen != m.n_sen || n != n_sen * n_feat * n_cw) {
        set_error("%s: dimensions do not match the model", path.c_str());
        return false;
    }
    if (m.cfg.mixwfloor <= 0.0 || m.cfg.mixwfloor >= 1.0) {
        set_error("mixwfloor (%e) not in range (0, 1)", m.cfg.mixwfloor);
        return false;
    }
    LogMath lm(m.cfg.logbase);
    const float flr_f = (float)m.cfg.mixwfloor;  // the reference holds it in a float32
    const double flr = flr_f;
    std::vector<float> pdf(n_cw);
    m.mixw.assign((size_t)n_sen * n_feat * n_cw, 0);
    auto sum_norm = [&]() {
        double s = 0.0;
        for (float x : pdf)
            s += x;
        if (s != 0.0) {
            const double r = 1.0 / s;
            for (float &x : pdf)
                x = (float)(x * r);
        }
    };
    for (int s = 0; s < n_sen; ++s)
        for (int ft = 0; ft < n_feat; ++ft) {
            if (!f.words(pdf.data(), n_cw))
                return false;
            sum_norm();
            for (float &x : pdf)
                if (x < flr)
                    x = (float)flr;
            sum_norm();
            for (int c = 0; c < n_cw; ++c) {
                int32_t p = -lm.log(pdf[c], 0) + ((1 << 9) - 1);
                m.mixw[((size_t)s * n_feat + ft) * n_cw + c] = (uint8_t)(p < (255 << 10) ? p >> 10 : 255);
            }
        }
    if (!f.s3_trailer()) {
        set_error("%s: checksum mismatch", path.c_str());
        return false;
    }
    return true;
}

// ---------------------------------------------------------------- tmat
// ref: src/tmat.c:125-225: rows normalised, non-zero entries floored, renormalised,
// then min(255, (-log_b p) >> 10); zero probability -> 255.
static bool read_tmat(HostModel &m, const std::string &path)
{
    Blob f;
    int32_t n_tmat, n_src, n_dst, n;
    if (!f.open(path) || !f.s3_header() || !f.i32(n_tmat) || !f.i32(n_src) || !f.i32(n_dst)
        || !f.i32(n) || n_dst != n_src + 1 || n != n_tmat * n_src * n_dst) {
        set_error("%s: unreadable or inconsistent transition matrices", path.c_str());
        return false;
    }
    if (n_src != m.n_emit) {
        set_error("%s: %d emitting states but mdef says %d", path.c_str(), n_src, m.n_emit);
        return false;
    }
    LogMath lm(m.cfg.logbase);
    const double floor_p = m.cfg.tmatfloor;
    m.n_tmat = n_tmat;
    m.tp.resize((size_t)n);
    std::vector<float> mat((size_t)n_src * n_dst);
    for (int t = 0; t < n_tmat; ++t) {
        if (!f.words(mat.data(), mat.size()))
            return false;
        for (int i = 0; i < n_src; ++i) {
            float *row = &mat[(size_t)i * n_dst];
            for (int round = 0; round < 2; ++round) {
                double s = 0.0;
                for (int k = 0; k < n_dst; ++k)
                    s += row[k];
                if (s != 0.0) {
                    double r = 1.0 / s;
                    for (int k = 0; k < n_dst; ++k)
                        row[k] = (float)(row[k] * r);
                }
                if (round == 0)
                    for (int k = 0; k < n_dst; ++k)
                        if (row[k] != 0.0 && row[k] < floor_p)
                            row[k] = (float)floor_p;
            }
            for (int k = 0; k < n_dst; ++k) {
                int32_t q = (-lm.log(row[k], 0)) >> 10;
                m.tp[((size_t)t * n_src + i) * n_dst + k] = (uint8_t)(q > 255 ? 255 : q);
            }
        }
    }
    if (!f.s3_trailer()) {
        set_error("%s: checksum mismatch", path.c_str());
        return false;
    }
    return true;
}

// ---------------------------------------------------------------- top level
bool HostModel::load(const std::string &dir, const ssb_config_t &c)
{
    cfg = c;
    int32_t d1[3], d2[3], fl2[SSB_MAX_FEAT] = {0, 0, 0, 0};
    if (!read_gauden(dir + "/means", d1, featlen, mean)
        || !read_gauden(dir + "/variances", d2, fl2, var))
        return false;
    if (std::memcmp(d1, d2, sizeof(d1)) || std::memcmp(featlen, fl2, sizeof(fl2))) {
        set_error("means and variances have different shapes");
        return false;
    }
    n_mgau = d1[0];
    n_feat = d1[1];
    n_density = d1[2];
    if (n_density > 256) {
        set_error("at most 256 densities per codebook are supported");
        return false;
    }
    blk = 0;
    for (int f = 0; f < n_feat; ++f) {
        featoff[f] = blk;
        blk += featlen[f];
    }
    featoff[n_feat] = blk;
    gau_off.resize((size_t)n_mgau * n_feat);
    int64_t off = 0;
    for (int c2 = 0; c2 < n_mgau; ++c2)
        for (int f = 0; f < n_feat; ++f) {
            gau_off[c2 * n_feat + f] = off;
            off += (int64_t)n_density * featlen[f];
        }
    precompute_gauden(*this);
    if (!read_mdef(*this, dir + "/mdef"))
        return false;
    // acmod_load_am's order (ref: src/acmod.c:101-119): PTM needs one codebook per CI phone
    // (ref: src/ptm_mgau.c:760-764), s2_semi a single codebook (ref: src/s2_semi_mgau.c:947-949)
    if (n_mgau == n_ciphone)
        kind = SSB_SCORER_PTM;
    else if (n_mgau == 1) {
        kind = SSB_SCORER_SEMI;
        std::fill(sen2cb.begin(), sen2cb.end(), (uint8_t)0);
    } else if (n_mgau == n_sen) {
        // ms_mgau with one codebook per senone (ref: src/ms_senone.c:262-275, ".cont.")
        kind = SSB_SCORER_CONT;
        std::fill(sen2cb.begin(), sen2cb.end(), (uint8_t)0);
    } else {
        set_error("%d codebooks for %d CI phones / %d senones: not a PTM, semi-continuous or "
                  "continuous model", n_mgau, n_ciphone, n_sen);
        return false;
    }
    if (kind == SSB_SCORER_PTM && n_mgau > SSB_MAX_CB) {
        set_error("at most %d codebooks are supported (ref: src/ptm_mgau.c:754)", SSB_MAX_CB);
        return false;
    }
    if (kind == SSB_SCORER_CONT) {
        struct stat sb;
        if (stat((dir + "/senmgau").c_str(), &sb) == 0) {
            set_error("%s/senmgau: explicit senone-to-codebook maps are not supported", dir.c_str());
            return false;
        }
        if (!read_mixw_cont(*this, dir + "/mixture_weights"))
            return false;
    } else
    if (!read_sendump(*this, dir + "/sendump")) {
        if (!mixw.empty() || !read_mixw_float(*this, dir + "/mixture_weights"))
            return false;
    }
    if (!read_tmat(*this, dir + "/transition_matrices"))
        return false;
    LogMath lm(cfg.logbase);
    if (!lm.add_table8(10, lut8)) {
        // ref: src/ptm_mgau.c:740-744
        set_error("log base %f is too small to represent the add table in 8 bits", cfg.logbase);
        return false;
    }
    return true;
}

}  // namespace ssb
