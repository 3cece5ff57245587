// api.cu -- the C ABI of libssb200.so (include/ssb200.h): model image in HBM, the
// mgau_t drop-in, the host-side planner of a batch and the kernel pipeline
//   K1 gmm_topn -> K2 senone_mix -> K3 chain_viterbi -> backtrace.
// No CPU fallback: every compute entry point needs a CUDA device.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <condition_variable>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "device.cuh"
#include "host_util.cuh"

namespace ssb {
const char *last_error();
int launch_senone_mix_active(const DevModel &m, const DevPlan &p, const int4 *tn_score,
                             const uchar4 *tn_cw, int64_t n_frames, int max_union,
                             int max_frames_per_utt, int16_t *chain_scr, cudaStream_t st);
int launch_chain_viterbi(const DevModel &m, const DevPlan &p, const int16_t *chain_scr,
                         int2 *tokens, int32_t *spill, int64_t spill_stride, int32_t *utt_best,
                         int32_t *utt_renorm, int32_t *fin_hist, int32_t *fin_score,
                         int max_phones, int max_band, cudaStream_t st);
// cont_score.cu: fully continuous models (ref: src/ms_mgau.c:279-368)
int launch_cont_dense(const DevModel &m, const float *feat, int64_t g0, int64_t n, int16_t *dense,
                      cudaStream_t st);
int launch_cont_active(const DevModel &m, const DevPlan &p, const float *feat, int W,
                       int max_frames_per_utt, int16_t *scratch, int16_t *chain_scr,
                       cudaStream_t st);
int launch_topn_fixup(const DevModel &m, const DevPlan &p, const int32_t *seg_utts, int n_seg_utts,
                      const float *feat, int64_t n_frames, int4 *tn_s, uchar4 *tn_c,
                      const uint32_t *tie, int64_t tie_w, cudaStream_t st);
int launch_cont_frame(const DevModel &m, const float *x, const uint16_t *act_sen, int n_act,
                      int compallsen, int16_t *raw, int16_t *senscr, cudaStream_t st);
}  // namespace ssb

using namespace ssb;

// ------------------------------------------------------------------ small helpers
namespace ssb {
cudaError_t raise_dyn_smem_limit(const void *kernel, size_t bytes)
{
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, size_t> limit;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    size_t &cur = limit[std::make_pair(dev, kernel)];
    if (bytes <= cur)
        return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess)
        cur = bytes;
    return e;
}
}  // namespace ssb

namespace {

template <class T>
int upload(DBuf &b, const std::vector<T> &v, cudaStream_t st)
{
    size_t bytes = v.size() * sizeof(T);
    if (b.ensure(bytes ? bytes : 16) != 0)
        return -1;
    if (bytes)
        SSB_CUDA(cudaMemcpyAsync(b.p, v.data(), bytes, cudaMemcpyHostToDevice, st));
    return 0;
}

}  // namespace

// ------------------------------------------------------------------ model
struct ssb_model_s {
    HostModel h;
    DevModel d;
    int device = -1;
    std::vector<void *> owned;
    std::vector<uint32_t> tc_hot;  // [cs][4] densities with their own screening error bound
    int sm_count = 0;
};

namespace ssb {
const HostModel *model_host(const ssb_model_t *m) { return m ? &m->h : nullptr; }  // lexicon.cpp
}
extern "C" int ssb_version(void) { return 100; }
extern "C" const char *ssb_model_ciphone_str(const ssb_model_t *m, int32_t ci)
{
    if (!m || ci < 0 || ci >= (int32_t)m->h.ciname.size())
        return nullptr;
    return m->h.ciname[ci].c_str();
}
extern "C" int ssb_model_kind(const ssb_model_t *m)
{
    if (!m) {
        set_error("ssb_model_kind: model is NULL");
        return -1;
    }
    return m->h.kind;
}
extern "C" const char *ssb_last_error(void) { return ssb::last_error(); }

extern "C" int ssb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess
            && major == 10)
            ++ok;
    }
    return ok;
}

extern "C" int ssb_device_cache_trim(void)
{
    DevBlockCache &c = dev_block_cache();
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess)
        return -1;
    std::lock_guard<std::mutex> lk(c.mu);
    c.trim(dev & 15);
    return 0;
}

extern "C" void ssb_config_defaults(ssb_config_t *c)
{
    // ref: include/soundswallower/config_defs.h:78-257
    c->logbase = 1.0001;
    c->varfloor = 1e-4f;
    c->mixwfloor = 1e-7;
    c->tmatfloor = 1e-4;
    c->topn = 4;
    c->ds = 1;
    c->device = 0;
    for (int f = 0; f < SSB_MAX_FEAT; ++f)
        c->topn_beam[f] = 0;
}

template <class T>
static const T *to_device(ssb_model_s *m, const std::vector<T> &v)
{
    void *p = nullptr;
    size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
    if (cudaMalloc(&p, bytes) != cudaSuccess
        || cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("model upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (p)
            cudaFree(p);
        return nullptr;
    }
    m->owned.push_back(p);
    return reinterpret_cast<const T *>(p);
}

static int model_to_device(ssb_model_s *m)
{
    const HostModel &h = m->h;
    DevModel &d = m->d;
    int dev_major = 0;
    API_CUDA(cudaSetDevice(m->device), -1);
    API_CUDA(cudaDeviceGetAttribute(&dev_major, cudaDevAttrComputeCapabilityMajor, m->device), -1);
    if (dev_major != 10) {
        set_error("device %d has compute capability %d.x; libssb200 carries sm_100a code only",
                  m->device, dev_major);
        return -1;
    }
    API_CUDA(cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, m->device), -1);
    std::memset(&d, 0, sizeof(d));
    d.n_mgau = h.n_mgau;
    d.n_feat = h.n_feat;
    d.n_density = h.n_density;
    d.n_sen = h.n_sen;
    d.n_emit = h.n_emit;
    d.n_tmat = h.n_tmat;
    d.n_sseq = h.n_sseq;
    d.blk = h.blk;
    d.topn = h.cfg.topn;
    d.ds = h.cfg.ds;
    d.kind = h.kind;
    for (int f = 0; f < SSB_MAX_FEAT; ++f)  // the reference keeps the beams in uint8
        d.topn_beam[f] = h.kind == SSB_SCORER_SEMI ? (h.cfg.topn_beam[f] & 0xff) : 0;
    // packed Gaussian records: [det, mean[L], prec[L], 0-pad], codebook-major
    int64_t off = 0;
    for (int f = 0; f < h.n_feat; ++f) {
        d.featlen[f] = h.featlen[f];
        d.featoff[f] = h.featoff[f];
        d.rec_len[f] = (1 + 2 * h.featlen[f] + 3) & ~3;
        d.gau_base[f] = off;
        off += (int64_t)h.n_density * d.rec_len[f];
    }
    d.gau_cb_stride = off;
    std::vector<float> gau((size_t)off * h.n_mgau, 0.f);
    for (int c = 0; c < h.n_mgau; ++c)
        for (int f = 0; f < h.n_feat; ++f) {
            const int L = h.featlen[f], RL = d.rec_len[f];
            const float *mu = h.mean.data() + h.gau_off[c * h.n_feat + f];
            const float *pv = h.var.data() + h.gau_off[c * h.n_feat + f];
            const float *dt = h.det.data() + (size_t)(c * h.n_feat + f) * h.n_density;
            float *dst = gau.data() + d.gau_base[f] + (int64_t)c * d.gau_cb_stride;
            for (int k = 0; k < h.n_density; ++k) {
                float *r = dst + (size_t)k * RL;
                r[0] = dt[k];
                for (int i = 0; i < L; ++i) {
                    r[1 + i] = mu[k * L + i];
                    r[1 + L + i] = pv[k * L + i];
                }
            }
        }
    // senones grouped by codebook
    std::vector<int32_t> cb_off(h.n_mgau + 1, 0);
    std::vector<uint16_t> cb_sen(h.n_sen);
    for (int s = 0; s < h.n_sen; ++s)
        cb_off[h.sen2cb[s] + 1]++;
    int max_cb = 0;
    for (int c = 0; c < h.n_mgau; ++c) {
        max_cb = std::max(max_cb, cb_off[c + 1]);
        cb_off[c + 1] += cb_off[c];
    }
    {
        std::vector<int32_t> at(cb_off.begin(), cb_off.end() - 1);
        for (int s = 0; s < h.n_sen; ++s)
            cb_sen[at[h.sen2cb[s]]++] = (uint16_t)s;
    }
    d.max_cb_sen = max_cb;
    // operands of the tensor-core screening GEMM (gmm_topn_tc.cu), double precision -> TF32
    std::vector<float> gB, gBlo, gAux;
    bool tc_shape = h.n_density == 128;
    for (int f = 0; f < h.n_feat; ++f)
        tc_shape = tc_shape && h.featlen[f] == 13;
    if (tc_shape) {
        auto tf32 = [](double x) {
            float v = (float)x;
            uint32_t b;
            std::memcpy(&b, &v, 4);
            b = (b + 0x1000u) & 0xFFFFE000u;  // round to nearest (ties away), 10-bit mantissa
            std::memcpy(&v, &b, 4);
            return v;
        };
        const int CS = h.n_mgau * h.n_feat, L = 13, ND = 128, K = 32;
        gB.assign((size_t)CS * ND * K, 0.f);
        gBlo.assign((size_t)CS * ND * K, 0.f);
        gAux.assign((size_t)CS * SSB_TC_AUX, 0.f);
        m->tc_hot.assign((size_t)CS * 4, 0u);
        std::vector<float> col(ND);
        for (int cs = 0; cs < CS; ++cs) {
            const float *mu = h.mean.data() + h.gau_off[cs];
            const float *pv = h.var.data() + h.gau_off[cs];
            const float *dt = h.det.data() + (size_t)cs * ND;
            float *B = gB.data() + (size_t)cs * ND * K;
            float *Bl = gBlo.data() + (size_t)cs * ND * K;
            float *aux = gAux.data() + (size_t)cs * SSB_TC_AUX;
            uint32_t *hot = m->tc_hot.data() + (size_t)cs * 4;
            for (int j = 0; j < L; ++j) {
                double c = 0;
                for (int n = 0; n < ND; ++n)
                    c += mu[n * L + j];
                aux[j] = (float)(c / ND);  // the value the kernel subtracts, in fp32
            }
            // "hot" densities: a precision far above the codebook's typical one in some
            // dimension (floored variances reach 5e7 against a median of ~60).  They would
            // inflate the screening error bound of every other density, so they get their own.
            for (int j = 0; j < L; ++j) {
                for (int n = 0; n < ND; ++n)
                    col[n] = pv[n * L + j];
                std::nth_element(col.begin(), col.begin() + ND / 2, col.end());
                const float lim = 16.f * col[ND / 2];
                for (int n = 0; n < ND; ++n)
                    if (pv[n * L + j] > lim)
                        hot[n >> 5] |= 1u << (n & 31);
            }
            for (int n = 0; n < ND; ++n) {
                double cst = dt[n];
                for (int j = 0; j < L; ++j) {
                    double mc = (double)mu[n * L + j] - (double)aux[j], v = pv[n * L + j];
                    B[n * K + j] = tf32(2.0 * mc * v);
                    B[n * K + L + j] = tf32(-v);
                    Bl[n * K + j] = tf32(2.0 * mc * v - (double)B[n * K + j]);
                    Bl[n * K + L + j] = tf32(-v - (double)B[n * K + L + j]);
                    cst -= mc * mc * v;
                    // [13..38] maxima over all densities (v1 kernel); [40..65] over the regular
                    // ones, [67..92] over the hot ones (v2 kernel)
                    const int o = ((hot[n >> 5] >> (n & 31)) & 1u) ? 67 : 40;
                    aux[13 + j] = std::max(aux[13 + j], std::fabs(B[n * K + j]));
                    aux[26 + j] = std::max(aux[26 + j], std::fabs(B[n * K + L + j]));
                    aux[o + j] = std::max(aux[o + j], std::fabs(B[n * K + j]));
                    aux[o + 13 + j] = std::max(aux[o + 13 + j], std::fabs(B[n * K + L + j]));
                }
                float hi = tf32(cst);
                B[n * K + 26] = hi;
                B[n * K + 27] = tf32(cst - (double)hi);
                aux[39] = std::max(aux[39], (float)std::fabs(cst));
            }
        }
        if (!(d.gB = to_device(m, gB)) || !(d.gBlo = to_device(m, gBlo))
            || !(d.gAux = to_device(m, gAux)) || !(d.gHot = to_device(m, m->tc_hot)))
            return -1;
        // ---- frame-tiled kernel (gmm_scan_ft.cu): the same split operands around ONE centre per
        // stream (the mean of all the stream's means), stored in shared-memory swizzle order
        if (h.n_feat <= SSB_MAX_FEAT) {
            std::vector<float> gBft((size_t)CS * 2 * ND * K, 0.f), gAuxFt((size_t)CS * 64, 0.f);
            for (int f = 0; f < h.n_feat; ++f)
                for (int j = 0; j < L; ++j) {
                    double c = 0;
                    for (int cb = 0; cb < h.n_mgau; ++cb) {
                        const float *mu = h.mean.data() + h.gau_off[cb * h.n_feat + f];
                        for (int n = 0; n < ND; ++n)
                            c += mu[n * L + j];
                    }
                    d.ft_centre[f * 16 + j] = (float)(c / ((double)ND * h.n_mgau));
                }
            auto swz = [](int n, int k) {  // float index of element (row n, k) in a SWIZZLE_128B tile
                return ((n >> 3) * 64 + (n & 7) * 8 + ((k >> 2) ^ (n & 7))) * 4 + (k & 3);
            };
            for (int cs = 0; cs < CS; ++cs) {
                const int f = cs % h.n_feat;
                const float *mu = h.mean.data() + h.gau_off[cs];
                const float *pv = h.var.data() + h.gau_off[cs];
                const float *dt = h.det.data() + (size_t)cs * ND;
                const uint32_t *hot = m->tc_hot.data() + (size_t)cs * 4;
                float *Bh = gBft.data() + (size_t)cs * 2 * ND * K, *Bl = Bh + ND * K;
                float *aux = gAuxFt.data() + (size_t)cs * 64;
                bool any_hot = false;
                for (int n = 0; n < ND; ++n) {
                    const bool is_hot = (hot[n >> 5] >> (n & 31)) & 1u;
                    any_hot = any_hot || is_hot;
                    double cst = dt[n];
                    for (int j = 0; j < L; ++j) {
                        const double mc = (double)mu[n * L + j] - (double)d.ft_centre[f * 16 + j];
                        const double v = pv[n * L + j];
                        const float b1 = tf32(2.0 * mc * v), b2 = tf32(-v);
                        Bh[swz(n, j)] = b1;
                        Bh[swz(n, L + j)] = b2;
                        Bl[swz(n, j)] = tf32(2.0 * mc * v - (double)b1);
                        Bl[swz(n, L + j)] = tf32(-v - (double)b2);
                        cst -= mc * mc * v;
                        const int o = is_hot ? 32 : 0;
                        aux[o + j] = std::max(aux[o + j], std::fabs(b1));
                        aux[o + 16 + j] = std::max(aux[o + 16 + j], std::fabs(b2));
                    }
                    const float hi = tf32(cst);
                    Bh[swz(n, 26)] = hi;
                    Bh[swz(n, 27)] = tf32(cst - (double)hi);
                    const int o = is_hot ? 14 : 13;
                    aux[o] = std::max(aux[o], (float)std::fabs(cst));
                }
                aux[15] = any_hot ? 1.f : 0.f;
                aux[29] = aux[13];
                aux[30] = aux[14];
                aux[31] = aux[15];
            }
            if (!(d.gBft = to_device(m, gBft)) || !(d.gAuxFt = to_device(m, gAuxFt)))
                return -1;
        }
    }
    std::vector<uint8_t> lut(h.lut8, h.lut8 + 256);
    if (!(d.gau = to_device(m, gau)) || !(d.mixw = to_device(m, h.mixw))
        || !(d.sen2cb = to_device(m, h.sen2cb)) || !(d.sseq = to_device(m, h.sseq))
        || !(d.tp = to_device(m, h.tp)) || !(d.lut8 = to_device(m, lut))
        || !(d.cb_sen_off = to_device(m, cb_off)) || !(d.cb_sen = to_device(m, cb_sen)))
        return -1;
    return 0;
}

extern "C" ssb_model_t *ssb_model_load(const char *hmmdir, const ssb_config_t *cfg)
{
    ssb_config_t c;
    if (cfg)
        c = *cfg;
    else
        ssb_config_defaults(&c);
    if (c.topn < 1 || c.topn > SSB_MAX_TOPN) {
        set_error("topn %d out of range 1..%d", c.topn, SSB_MAX_TOPN);
        return nullptr;
    }
    if (c.ds < 1)
        c.ds = 1;
    std::unique_ptr<ssb_model_s> m(new ssb_model_s);
    if (!m->h.load(hmmdir ? hmmdir : "", c))
        return nullptr;
    if (m->h.n_sen > 65535) {
        set_error("%d senones exceed the 16-bit senone ids of the chain planner", m->h.n_sen);
        return nullptr;
    }
    m->device = c.device;
    if (c.device >= 0 && model_to_device(m.get()) != 0) {
        for (void *p : m->owned)
            cudaFree(p);
        return nullptr;
    }
    return m.release();
}

extern "C" void ssb_model_free(ssb_model_t *m)
{
    if (!m)
        return;
    if (m->device >= 0)
        cudaSetDevice(m->device);
    for (void *p : m->owned)
        cudaFree(p);
    delete m;
}

extern "C" int ssb_model_dims(const ssb_model_t *m, int32_t *o)
{
    if (!m || !o)
        return -1;
    const HostModel &h = m->h;
    int32_t v[16] = {h.n_mgau, h.n_feat, h.n_density, h.featlen[0], h.n_sen, h.n_sseq,
                     h.n_emit, h.n_tmat, h.n_ciphone, h.n_phone, h.sil,
                     h.featlen[0], h.featlen[1], h.featlen[2], h.featlen[3], h.blk};
    std::memcpy(o, v, sizeof(v));
    return 0;
}

extern "C" int ssb_model_copy(const ssb_model_t *m, float *mean, float *var, float *det,
                              uint8_t *mixw, uint8_t *sen2cb, uint8_t *tp, uint16_t *sseq,
                              uint8_t *lut8)
{
    if (!m)
        return -1;
    const HostModel &h = m->h;
    if (mean)
        std::memcpy(mean, h.mean.data(), h.mean.size() * 4);
    if (var)
        std::memcpy(var, h.var.data(), h.var.size() * 4);
    if (det)
        std::memcpy(det, h.det.data(), h.det.size() * 4);
    if (mixw)
        std::memcpy(mixw, h.mixw.data(), h.mixw.size());
    if (sen2cb)
        std::memcpy(sen2cb, h.sen2cb.data(), h.sen2cb.size());
    if (tp)
        std::memcpy(tp, h.tp.data(), h.tp.size());
    if (sseq)
        std::memcpy(sseq, h.sseq.data(), h.sseq.size() * 2);
    if (lut8)
        std::memcpy(lut8, h.lut8, 256);
    return 0;
}

extern "C" int ssb_model_phones(const ssb_model_t *m, int32_t *ssid, int32_t *tmat, int32_t *ci)
{
    if (!m)
        return -1;
    const HostModel &h = m->h;
    if (ssid)
        std::memcpy(ssid, h.ph_ssid.data(), h.ph_ssid.size() * 4);
    if (tmat)
        std::memcpy(tmat, h.ph_tmat.data(), h.ph_tmat.size() * 4);
    if (ci)
        std::memcpy(ci, h.ph_ci.data(), h.ph_ci.size() * 4);
    return 0;
}

static int need_device(const ssb_model_t *m)
{
    if (!m) {
        set_error("NULL model");
        return -1;
    }
    if (m->device < 0) {
        set_error("model was loaded with device = -1 (host tables only); "
                  "libssb200 has no CPU compute path");
        return -1;
    }
    API_CUDA(cudaSetDevice(m->device), -1);
    return 0;
}

// ------------------------------------------------------------------ active list decoding
// ref: src/acmod.c:947-999 -- the list is uint8 deltas; a gap above 255 is bridged with
// entries of 255 which make the scorer evaluate those in-between senones too.
static void decode_active(const uint8_t *list, int n, std::vector<uint16_t> &out)
{
    out.clear();
    int last = 0;
    for (int i = 0; i < n; ++i) {
        last += list[i];
        out.push_back((uint16_t)last);
    }
}

// ------------------------------------------------------------------ mgau drop-in
struct ssb_mgau_impl {
    ssb_mgau_t base;  // must be first: {vt, frame_idx}
    ssb_model_t *m;
    bool owns_model = false;  // ssb_mgau_own_model: freeing the scorer frees the model too
    FrameHist hist;
    DBuf hs[2], hc[2], ha[2], x, senscr, act, raw;
    float *h_x = nullptr;
    int16_t *h_senscr = nullptr;
    uint16_t *h_act = nullptr;
    uint8_t *h_cb = nullptr;
    cudaStream_t st = nullptr;
    std::vector<uint16_t> tmp;
};

static int mgau_frame_eval_vt(ssb_mgau_t *g, int16_t *senscr, uint8_t *act, int32_t n_act,
                              float **feat, int32_t frame, int32_t compallsen)
{
    return ssb_mgau_frame_eval(g, senscr, act, n_act, feat, frame, compallsen);
}
static int mgau_transform_vt(ssb_mgau_t *, void *)
{
    // ref: src/ptm_mgau.c:818-825 (MLLR re-estimates means/variances); out of scope
    set_error("MLLR transform is not supported by the B200 scorer");
    return -1;
}
static void mgau_free_vt(ssb_mgau_t *g) { ssb_mgau_free(g); }
// one vtable per scorer family, named like the reference's (ptm_mgau.c:57, s2_semi_mgau.c:57,
// ms_mgau.c:75)
static ssb_mgaufuncs_t g_mgau_funcs[3] = {
    {"ptm", mgau_frame_eval_vt, mgau_transform_vt, mgau_free_vt},
    {"s2_semi", mgau_frame_eval_vt, mgau_transform_vt, mgau_free_vt},
    {"ms", mgau_frame_eval_vt, mgau_transform_vt, mgau_free_vt}};

static int mgau_reset_device(ssb_mgau_impl *g)
{
    const HostModel &h = g->m->h;
    const int CS = h.n_mgau * h.n_feat;
    std::vector<int4> s(CS, make_int4(INT32_MIN, INT32_MIN, INT32_MIN, INT32_MIN));
    std::vector<uchar4> c(CS, make_uchar4(0, 1, 2, 3));
    std::vector<uint8_t> a(h.n_mgau, 1);
    for (int i = 0; i < 2; ++i) {
        API_CUDA(cudaMemcpyAsync(g->hs[i].p, s.data(), CS * sizeof(int4), cudaMemcpyHostToDevice, g->st), -1);
        API_CUDA(cudaMemcpyAsync(g->hc[i].p, c.data(), CS * sizeof(uchar4), cudaMemcpyHostToDevice, g->st), -1);
        API_CUDA(cudaMemcpyAsync(g->ha[i].p, a.data(), h.n_mgau, cudaMemcpyHostToDevice, g->st), -1);
    }
    API_CUDA(cudaStreamSynchronize(g->st), -1);
    return 0;
}

extern "C" ssb_mgau_t *ssb_mgau_init(ssb_model_t *m)
{
    if (need_device(m) != 0)
        return nullptr;
    const HostModel &h = m->h;
    const int CS = h.n_mgau * h.n_feat;
    ssb_mgau_impl *g = new ssb_mgau_impl;
    g->base.vt = &g_mgau_funcs[h.kind];
    g->base.frame_idx = 0;
    g->m = m;
    bool ok = cudaStreamCreateWithFlags(&g->st, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i) {
        ok = g->hs[i].ensure(CS * sizeof(int4)) == 0 && g->hc[i].ensure(CS * sizeof(uchar4)) == 0
             && g->ha[i].ensure(h.n_mgau) == 0;
        g->hist.score[i] = g->hs[i].as<int4>();
        g->hist.cw[i] = g->hc[i].as<uchar4>();
        g->hist.act[i] = g->ha[i].as<uint8_t>();
    }
    ok = ok && g->x.ensure(h.blk * sizeof(float)) == 0 && g->senscr.ensure(h.n_sen * 2) == 0
         && g->raw.ensure(h.n_sen * 2) == 0 && g->act.ensure((size_t)h.n_sen * 2 + 16) == 0;
    // the continuous scorer only writes the listed senones: the rest of the buffer starts as
    // the reference's calloc'ed senone_scores does
    ok = ok && cudaMemset(g->senscr.p, 0, h.n_sen * 2) == cudaSuccess;
    ok = ok && cudaMallocHost((void **)&g->h_x, h.blk * sizeof(float)) == cudaSuccess
         && cudaMallocHost((void **)&g->h_senscr, h.n_sen * 2) == cudaSuccess
         && cudaMallocHost((void **)&g->h_act, (size_t)h.n_sen * 2 + 16) == cudaSuccess
         && cudaMallocHost((void **)&g->h_cb, h.n_mgau) == cudaSuccess;
    if (!ok || mgau_reset_device(g) != 0) {
        if (ok)
            ;
        else
            set_error("ssb_mgau_init: device allocation failed: %s",
                      cudaGetErrorString(cudaGetLastError()));
        ssb_mgau_free(&g->base);
        return nullptr;
    }
    return &g->base;
}

extern "C" void ssb_mgau_own_model(ssb_mgau_t *gg, int own)
{
    ssb_mgau_impl *g = reinterpret_cast<ssb_mgau_impl *>(gg);
    if (g)
        g->owns_model = own != 0;
}

extern "C" ssb_model_t *ssb_mgau_model(const ssb_mgau_t *gg)
{
    const ssb_mgau_impl *g = reinterpret_cast<const ssb_mgau_impl *>(gg);
    return g ? g->m : nullptr;
}

extern "C" void ssb_mgau_reset(ssb_mgau_t *gg)
{
    ssb_mgau_impl *g = reinterpret_cast<ssb_mgau_impl *>(gg);
    if (!g)
        return;
    cudaSetDevice(g->m->device);
    mgau_reset_device(g);
    g->base.frame_idx = 0;
}

extern "C" void ssb_mgau_free(ssb_mgau_t *gg)
{
    ssb_mgau_impl *g = reinterpret_cast<ssb_mgau_impl *>(gg);
    if (!g)
        return;
    cudaSetDevice(g->m->device);
    for (int i = 0; i < 2; ++i) {
        g->hs[i].release();
        g->hc[i].release();
        g->ha[i].release();
    }
    g->x.release();
    g->senscr.release();
    g->act.release();
    g->raw.release();
    if (g->h_x)
        cudaFreeHost(g->h_x);
    if (g->h_senscr)
        cudaFreeHost(g->h_senscr);
    if (g->h_act)
        cudaFreeHost(g->h_act);
    if (g->h_cb)
        cudaFreeHost(g->h_cb);
    if (g->st)
        cudaStreamDestroy(g->st);
    ssb_model_t *owned = g->owns_model ? g->m : nullptr;
    delete g;
    if (owned)
        ssb_model_free(owned);
}

extern "C" int ssb_mgau_frame_eval(ssb_mgau_t *gg, int16_t *senscr, uint8_t *senone_active,
                                   int32_t n_act, float **feat, int32_t frame, int32_t compallsen)
{
    ssb_mgau_impl *g = reinterpret_cast<ssb_mgau_impl *>(gg);
    if (!g || !senscr || !feat || frame < 0) {
        set_error("ssb_mgau_frame_eval: bad arguments");
        return -1;
    }
    if (need_device(g->m) != 0)
        return -1;
    const HostModel &h = g->m->h;
    const DevModel &d = g->m->d;
    const int slot = frame % 2, prev = slot ? slot - 1 : 1;
    int n_list = 0;
    if (!compallsen) {
        decode_active(senone_active, n_act, g->tmp);
        n_list = (int)g->tmp.size();
        for (int i = 0; i < n_list; ++i) {
            if (g->tmp[i] >= h.n_sen) {
                set_error("active senone list runs past n_sen");
                return -1;
            }
            g->h_act[i] = g->tmp[i];
        }
        if (n_list)
            API_CUDA(cudaMemcpyAsync(g->act.p, g->h_act, n_list * 2, cudaMemcpyHostToDevice, g->st), -1);
    }
    if (h.kind == SSB_SCORER_CONT) {  // stateless: no history slots, no codebook flags
        for (int f = 0; f < h.n_feat; ++f)
            std::memcpy(g->h_x + h.featoff[f], feat[f], h.featlen[f] * sizeof(float));
        API_CUDA(cudaMemcpyAsync(g->x.p, g->h_x, h.blk * sizeof(float), cudaMemcpyHostToDevice, g->st), -1);
        if (launch_cont_frame(d, g->x.as<float>(), g->act.as<uint16_t>(), n_list, compallsen,
                              g->raw.as<int16_t>(), g->senscr.as<int16_t>(), g->st) != 0)
            return -1;
        API_CUDA(cudaMemcpyAsync(g->h_senscr, g->senscr.p, h.n_sen * 2, cudaMemcpyDeviceToHost, g->st), -1);
        API_CUDA(cudaStreamSynchronize(g->st), -1);
        std::memcpy(senscr, g->h_senscr, h.n_sen * 2);
        return 0;
    }
    const bool fresh = frame >= g->base.frame_idx;
    if (fresh) {
        // ptm_mgau_calc_cb_active (ref: src/ptm_mgau.c:297-321)
        if (compallsen || h.kind == SSB_SCORER_SEMI)  // s2_semi always scans its codebook
            std::memset(g->h_cb, 1, h.n_mgau);
        else {
            std::memset(g->h_cb, 0, h.n_mgau);
            for (int i = 0; i < n_list; ++i)
                g->h_cb[h.sen2cb[g->tmp[i]]] = 1;
        }
        for (int f = 0; f < h.n_feat; ++f)
            std::memcpy(g->h_x + h.featoff[f], feat[f], h.featlen[f] * sizeof(float));
        API_CUDA(cudaMemcpyAsync(g->hist.act[slot], g->h_cb, h.n_mgau, cudaMemcpyHostToDevice, g->st), -1);
        API_CUDA(cudaMemcpyAsync(g->x.p, g->h_x, h.blk * sizeof(float), cudaMemcpyHostToDevice, g->st), -1);
        if (launch_frame_topn(d, g->hist, slot, prev, g->x.as<float>(), frame % h.cfg.ds == 0, g->st) != 0)
            return -1;
    }
    if (launch_frame_senones(d, g->hist, slot, fresh ? 1 : 0, g->act.as<uint16_t>(), n_list,
                             compallsen, g->senscr.as<int16_t>(), g->st) != 0)
        return -1;
    API_CUDA(cudaMemcpyAsync(g->h_senscr, g->senscr.p, h.n_sen * 2, cudaMemcpyDeviceToHost, g->st), -1);
    API_CUDA(cudaStreamSynchronize(g->st), -1);
    std::memcpy(senscr, g->h_senscr, h.n_sen * 2);
    return 0;
}

// ------------------------------------------------------------------ batch
struct ssb_batch_s {
    ssb_model_t *m = nullptr;
    cudaStream_t st = nullptr;
    // plan (host)
    int n_utts = 0;
    int64_t n_frames = 0, n_phones = 0, n_states = 0, n_state_frames = 0;
    int64_t n_active_sen_frames = 0, n_scanned_cb_frames = 0, n_band_state_frames = 0;
    int64_t plan_us = 0;  // host time of the last upload's planning + staging calls
    int max_phones = 0, max_union = 0, max_T = 0, max_band = 0;
    int compallsen = 0;
    bool want_tokens_all = false;  // debug: dense token stack, pre-filled, downloadable
    bool want_dense = false;       // debug: dense chain scores / tokens (the reference's layout)
    bool banded = false;           // chain_scr / tokens keep only the evaluated band (DevPlan.banded)
    int64_t n_band_scr = 0, n_band_tok = 0;
    std::vector<int64_t> frame_off, phone_off, scr_off;
    std::vector<int32_t> enter;
    // device
    DBuf feat, featp, tn_s, tn_c, chain_scr, tokens, spill;
    DBuf d_frame_off, d_phone_off, d_scr_off, d_ssid, d_tmat, d_sf, d_ef, d_ep_off, d_ep_start,
        d_ep_cbmask, d_ep_slot_off, d_ep_slot, d_us_off, d_usen, d_st_slot, d_enter;
    DBuf st_start, st_dur, st_score, utt_rv, utt_best, utt_renorm, fin_hist, fin_score;
    DBuf dense, best_tmp;
    DevPlan plan{};
    int64_t spill_stride = 0;
    // K1 over time (topn_fixup.cu): long utterances of a small batch cut into segments
    bool k1_seg = false;
    DevPlan k1_plan{};
    DBuf d_k1_frame_off, d_k1_ep_off, d_k1_ep_start, d_k1_ep_cbmask, d_seg_utts, d_k1_tie, d_init_topn;
    int n_seg_utts = 0, n_k1_rows = 0;
    int64_t k1_tie_w = 0;
    // K1 frame-tiled (gmm_scan_ft.cu): tiles of 128 frames; every tie step goes to the fix-up
    bool ft = false;
    int n_tiles = 0;
    DBuf d_tile_utt, d_tile_t0, d_tile_ctr;
    DBuf d_scr_boff, d_tok_boff;
    // chain cutting: K3 + backtrace run on segments of the utterances (cut_chains)
    bool cut = false, cut_ran = false;  // planned / what the last run used
    int n_segs = 0, cut_max_phones = 0, cut_max_band = 0;
    std::vector<int32_t> seg_off;       // [U+1] segments of each utterance
    DevPlan cut_plan{};
    DBuf d_seg_frame_off, d_seg_phone_off, d_seg_scr_off, d_seg_enter, d_seg_sf, d_seg_ef, d_seg_scr_boff,
        d_seg_tok_boff, d_seg_t0, seg_rv, seg_best, seg_renorm, seg_fin_hist, seg_fin_score;
    // timing
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int n_launches = 0;
    bool ran = false;
    size_t bytes_held() const
    {
        const DBuf *all[] = {&feat, &featp, &tn_s, &tn_c, &chain_scr, &tokens, &spill, &d_frame_off,
                             &d_phone_off, &d_scr_off, &d_ssid, &d_tmat, &d_sf, &d_ef, &d_ep_off,
                             &d_ep_start, &d_ep_cbmask, &d_ep_slot_off, &d_ep_slot, &d_us_off,
                             &d_usen, &d_st_slot, &d_enter, &st_start, &st_dur, &st_score, &utt_rv,
                             &utt_best, &utt_renorm, &fin_hist, &fin_score, &dense, &best_tmp,
                             &d_k1_frame_off, &d_k1_ep_off, &d_k1_ep_start, &d_k1_ep_cbmask,
                             &d_seg_utts, &d_k1_tie, &d_init_topn, &d_tile_utt, &d_tile_t0,
                             &d_tile_ctr, &d_scr_boff, &d_tok_boff, &d_seg_frame_off, &d_seg_phone_off,
                             &d_seg_scr_off, &d_seg_enter, &d_seg_sf, &d_seg_ef, &d_seg_scr_boff,
                             &d_seg_tok_boff, &d_seg_t0, &seg_rv, &seg_best, &seg_renorm, &seg_fin_hist,
                             &seg_fin_score};
        size_t n = 0;
        for (const DBuf *b : all)
            n += b->cap;
        return n;
    }
    void release_all()
    {
        DBuf *all[] = {&feat, &featp, &tn_s, &tn_c, &chain_scr, &tokens, &spill, &d_frame_off,
                       &d_phone_off, &d_scr_off, &d_ssid, &d_tmat, &d_sf, &d_ef, &d_ep_off,
                       &d_ep_start, &d_ep_cbmask, &d_ep_slot_off, &d_ep_slot, &d_us_off, &d_usen,
                       &d_st_slot, &d_enter, &st_start, &st_dur, &st_score, &utt_rv, &utt_best,
                       &utt_renorm, &fin_hist, &fin_score, &dense, &best_tmp, &d_k1_frame_off,
                       &d_k1_ep_off, &d_k1_ep_start, &d_k1_ep_cbmask, &d_seg_utts, &d_k1_tie,
                       &d_init_topn, &d_tile_utt, &d_tile_t0, &d_tile_ctr, &d_scr_boff, &d_tok_boff,
                       &d_seg_frame_off, &d_seg_phone_off, &d_seg_scr_off, &d_seg_enter, &d_seg_sf,
                       &d_seg_ef, &d_seg_scr_boff, &d_seg_tok_boff, &d_seg_t0, &seg_rv, &seg_best,
                       &seg_renorm, &seg_fin_hist, &seg_fin_score};
        for (DBuf *b : all)
            b->release();
    }
};

// final senone scores of frames [g0, g0+n) of the uploaded batch, "compallsen" semantics
static int dense_scores(ssb_batch_t *b, int64_t g0, int64_t n, int16_t *dense)
{
    const DevModel &d = b->m->d;
    if (d.kind == SSB_SCORER_CONT)
        return launch_cont_dense(d, b->feat.as<float>(), g0, n, dense, b->st);
    return launch_senone_mix_all(d, b->tn_s.as<int4>(), b->tn_c.as<uchar4>(), b->n_frames, g0, n,
                                 dense, b->st);
}

extern "C" ssb_batch_t *ssb_batch_create(ssb_model_t *m, void *stream)
{
    if (need_device(m) != 0)
        return nullptr;
    ssb_batch_s *b = new ssb_batch_s;
    b->m = m;
    b->st = reinterpret_cast<cudaStream_t>(stream);
    for (auto &e : b->ev)
        if (cudaEventCreate(&e) != cudaSuccess) {
            set_error("cudaEventCreate failed");
            ssb_batch_free(b);
            return nullptr;
        }
    return b;
}

extern "C" void ssb_batch_free(ssb_batch_t *b)
{
    if (!b)
        return;
    cudaSetDevice(b->m->device);
    cudaStreamSynchronize(b->st);
    b->release_all();
    for (auto &e : b->ev)
        if (e)
            cudaEventDestroy(e);
    delete b;
}

// Data-independent schedule of one chain (see chain_viterbi.cu header):
//   enter[0] = 0;  enter[i] = max(enter[i-1], sf[i], 1) if that is <= max(enter[i-1], ef[i-1])
//   and <= T, else the phone (and every later one) is never entered (-1).
static void plan_enter(int np, int T, const int32_t *sf, const int32_t *ef, int32_t *enter)
{
    for (int i = 0; i < np; ++i)
        enter[i] = -1;
    if (np == 0 || T == 0)
        return;
    enter[0] = 0;
    for (int i = 1; i < np; ++i) {
        // transitions happen at the end of a step, so the earliest entry frame is 1
        int32_t e = std::max<int32_t>(std::max(enter[i - 1], sf[i]), 1);
        int32_t last_prev = std::min<int32_t>(std::max(enter[i - 1], ef[i - 1]), T);
        if (e > last_prev)
            break;
        enter[i] = e;
    }
}

extern "C" int ssb_plan_chain(int32_t np, int32_t T, const int32_t *sf, const int32_t *ef,
                              int32_t *enter)
{
    if (np < 0 || T < 0 || (np > 0 && (!sf || !ef || !enter))) {
        set_error("ssb_plan_chain: bad arguments");
        return -1;
    }
    for (int i = 1; i < np; ++i)
        if (ef[i] < ef[i - 1]) {
            set_error("phone %d: window ends must not decrease along the chain", i);
            return -1;
        }
    plan_enter(np, T, sf, ef, enter);
    return 0;
}

// K1 kernel choice.  Two tcgen05 kernels give the same lists:
//   * gmm_scan_ft.cu (frame-tiled: A tile in tensor memory built once per 128 frames, B streamed
//     by TMA bulk copies, no exact evaluation on the common path) -- measured faster when every
//     codebook is scanned (first pass / compallsen: 56 against 84 ms on 4096 x 10 s) and when
//     lists are carried in from a first pass (no second instantiation at the register limit);
//   * gmm_topn_tc2.cu (codebook-stream x 256 utterances per CTA, 16 warps per SM) -- measured
//     faster for the aligner's active lists on big batches (26.5 against 28 ms on config #2).
// $SSB_K1 = ft / tc2 / fp32 forces one (A/B runs, profiles/prof_gmm_scan_ft_r2*.txt).
static bool k1_frame_tiled(const DevModel &d, bool all_active, bool carried_lists, int n_utts)
{
    const char *k1 = getenv("SSB_K1");
    if (k1 && *k1)
        return strcmp(k1, "ft") == 0 && ft_supported(d);
    if (const char *seg = getenv("SSB_K1_SEG"))
        if (atoll(seg) > 0)
            return false;  // (tests: the segmented tc2 launch)
    // tc2's rows are utterances, 256 per CTA: below ~1000 utterances its CTAs are mostly idle
    // lanes (or, for a long recording, a few hundred threads walking thousands of frames each),
    // while the frame-tiled kernel fills the machine with 128-frame tiles of whatever there is:
    // K1 of the 1-hour utterance 10.7 -> 2.6 ms
    return ft_supported(d) && (all_active || carried_lists || n_utts < 1024);
}

// Planner threads of one batch: the host's cores are shared by the ranks / device threads of a
// box (8 GPUs x 16 threads on 32 cores cost 5 % end to end in round 1), so: cores / local ranks,
// at most 16.  $SSB_PLAN_THREADS overrides; torchrun exports LOCAL_WORLD_SIZE;
// ssb_align_batch_multi sets the divisor for its device threads.
static std::atomic<int> g_plan_divisor{1};
static int planner_threads()
{
    if (const char *e = getenv("SSB_PLAN_THREADS"))
        if (atoi(e) > 0)
            return atoi(e);
    int div = g_plan_divisor.load();
    if (const char *e = getenv("LOCAL_WORLD_SIZE"))
        div = std::max(div, atoi(e));
    const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
    return std::max(1, std::min(16, hw / std::max(1, div)));
}

// ---- host planner --------------------------------------------------------------------------
// Per-utterance plan pieces produced by one worker for a contiguous range of utterances;
// offsets are local to the piece and rebased when the pieces are concatenated.
namespace {
struct PlanPiece {
    std::vector<int32_t> ep_count, ep_start, ep_slot_len, us_count;
    std::vector<uint32_t> ep_cbmask;
    std::vector<uint16_t> ep_slot, usen;
    int max_union = 0;
    int64_t active_sen_frames = 0, scanned_cb_frames = 0, band_state_frames = 0;
    std::string error;
};

struct PlanScratch {
    std::vector<uint8_t> mark, umark;   // [n_sen] active / in-union flags, reset sparsely
    std::vector<uint16_t> act, uni, slot_of, ev;
    std::vector<int32_t> ev_off;
    explicit PlanScratch(int n_sen) : mark(n_sen, 0), umark(n_sen, 0), slot_of(n_sen, 0) {}
};

// senone ids the scorer evaluates for the sorted active set `act`: acmod_flags2list's uint8
// delta list bridges gaps above 255 with real entries (ref: src/acmod.c:947-999)
inline void eval_list_from_sorted(const std::vector<uint16_t> &act, std::vector<uint16_t> &out)
{
    int last = 0;
    for (uint16_t s : act) {
        int delta = (int)s - last;
        while (delta > 255) {
            last += 255;
            delta -= 255;
            out.push_back((uint16_t)last);
        }
        out.push_back(s);
        last = s;
    }
}

// Plans utterances [u0, u1).  enter[] and st_slot[] are written in place (disjoint ranges).
void plan_range(const HostModel &h, const ssb_align_in_t *in, const std::vector<int64_t> &frame_off,
                const std::vector<int64_t> &phone_off, int u0, int u1, int32_t *enter_all,
                uint16_t *st_slot_all, PlanPiece &out)
{
    const int E = h.n_emit, n_sen = h.n_sen, nw = (n_sen + 31) / 32;
    PlanScratch S(n_sen);
    // Utterances with the same chain, windows, length and initial flags get the same plan (many
    // readers of one text, the benchmark's tiled sentence): planned once per worker, then copied.
    struct Done {
        int u;
        size_t ep0, n_ep, sl0, n_sl, us0;
        int n_us;
        int64_t act, scan, band;
    };
    std::unordered_map<uint64_t, Done> seen;
    for (int u = u0; u < u1; ++u) {
        const int64_t p0 = phone_off[u];
        const int T = (int)(frame_off[u + 1] - frame_off[u]);
        const int np = (int)(phone_off[u + 1] - p0);
        const int32_t *ssid = in->ssid + p0, *sf = in->sf + p0, *ef = in->ef + p0;
        int32_t *enter = enter_all + p0;
        uint16_t *st_slot = st_slot_all + p0 * E;
        uint64_t key = 1469598103934665603ull;
        auto mix = [&key](const void *ptr, size_t n) {
            const unsigned char *b = static_cast<const unsigned char *>(ptr);
            for (size_t i = 0; i < n; ++i)
                key = (key ^ b[i]) * 1099511628211ull;
        };
        mix(&T, sizeof T);
        mix(&np, sizeof np);
        mix(ssid, (size_t)np * 4);
        mix(sf, (size_t)np * 4);
        mix(ef, (size_t)np * 4);
        if (in->init_active)
            mix(in->init_active + (size_t)u * nw, (size_t)nw * 4);
        if (!in->compallsen) {
            auto it = seen.find(key);
            if (it != seen.end()) {
                const Done &d = it->second;
                const int64_t q0 = phone_off[d.u];
                const bool same = (int)(phone_off[d.u + 1] - q0) == np
                                  && (int)(frame_off[d.u + 1] - frame_off[d.u]) == T
                                  && std::memcmp(in->ssid + q0, ssid, (size_t)np * 4) == 0
                                  && std::memcmp(in->sf + q0, sf, (size_t)np * 4) == 0
                                  && std::memcmp(in->ef + q0, ef, (size_t)np * 4) == 0
                                  && (!in->init_active
                                      || std::memcmp(in->init_active + (size_t)d.u * nw,
                                                     in->init_active + (size_t)u * nw, (size_t)nw * 4) == 0);
                if (same) {
                    std::memcpy(enter, enter_all + q0, (size_t)np * 4);
                    std::memcpy(st_slot, st_slot_all + q0 * E, (size_t)np * E * 2);
                    // (copy by index: the vectors may reallocate while they grow)
                    for (size_t k = 0; k < d.n_ep; ++k) {
                        out.ep_start.push_back(out.ep_start[d.ep0 + k]);
                        out.ep_slot_len.push_back(out.ep_slot_len[d.ep0 + k]);
                        for (int w = 0; w < 8; ++w)
                            out.ep_cbmask.push_back(out.ep_cbmask[(d.ep0 + k) * 8 + w]);
                    }
                    for (size_t k = 0; k < d.n_sl; ++k)
                        out.ep_slot.push_back(out.ep_slot[d.sl0 + k]);
                    for (int k = 0; k < d.n_us; ++k)
                        out.usen.push_back(out.usen[d.us0 + k]);
                    out.ep_count.push_back((int32_t)d.n_ep);
                    out.us_count.push_back(d.n_us);
                    out.active_sen_frames += d.act;
                    out.scanned_cb_frames += d.scan;
                    out.band_state_frames += d.band;
                    continue;
                }
            }
        }
        const size_t rec_ep0 = out.ep_start.size(), rec_sl0 = out.ep_slot.size(), rec_us0 = out.usen.size();
        const int64_t rec_act = out.active_sen_frames, rec_scan = out.scanned_cb_frames,
                      rec_band = out.band_state_frames;
        plan_enter(np, T, sf, ef, enter);
        // state-frames the chain Viterbi evaluates: phone i on frames [enter, max(enter, ef)]
        for (int i = 0; i < np; ++i)
            if (enter[i] >= 0 && enter[i] < T)
                out.band_state_frames += (int64_t)E * (std::min<int64_t>(std::max(enter[i], ef[i]), T - 1) - enter[i] + 1);
        if (in->compallsen) {
            out.ep_count.push_back(0);
            out.us_count.push_back(0);
            out.active_sen_frames += (int64_t)T * n_sen;
            out.scanned_cb_frames += (int64_t)T * h.n_mgau;
            for (int i = 0; i < np * E; ++i)
                st_slot[i] = 0;
            continue;
        }
        // epochs: the active senone set only grows (ref: src/state_align_search.c:186-188)
        S.act.clear();
        S.uni.clear();
        S.ev.clear();
        S.ev_off.assign(1, 0);
        if (in->init_active) {
            const uint32_t *bits = in->init_active + (size_t)u * nw;
            for (int w = 0; w < nw; ++w)
                for (uint32_t x = bits[w]; x; x &= x - 1) {
                    int s = w * 32 + __builtin_ctz(x);
                    if (s < n_sen && !S.mark[s]) {
                        S.mark[s] = 1;
                        S.act.push_back((uint16_t)s);
                    }
                }
        }
        const size_t ep_first = out.ep_start.size();
        int n_ep = 0;
        // enter[] is non-decreasing over the entered prefix of the chain
        int i = 0;
        while (i < np && enter[i] >= 0 && enter[i] < T) {
            const int32_t start = enter[i];
            bool grew = n_ep == 0;
            const size_t before = S.act.size();
            for (; i < np && enter[i] == start; ++i)
                for (int j = 0; j < E; ++j) {
                    const int s = h.sseq[(size_t)ssid[i] * E + j];
                    if (!S.mark[s]) {
                        S.mark[s] = 1;
                        S.act.push_back((uint16_t)s);
                        grew = true;
                    }
                }
            if (!grew)
                continue;
            if (S.act.size() != before || n_ep == 0) {
                // the list was sorted up to `before`: sort the few newcomers and merge them in
                std::sort(S.act.begin() + before, S.act.end());
                std::inplace_merge(S.act.begin(), S.act.begin() + before, S.act.end());
            }
            const size_t e0 = S.ev.size();
            eval_list_from_sorted(S.act, S.ev);
            S.ev_off.push_back((int32_t)S.ev.size());
            uint32_t mask[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (size_t k = e0; k < S.ev.size(); ++k) {
                const uint16_t s = S.ev[k];
                const int cb = h.sen2cb[s];
                mask[cb >> 5] |= 1u << (cb & 31);
                if (!S.umark[s]) {
                    S.umark[s] = 1;
                    S.uni.push_back(s);
                }
            }
            out.ep_start.push_back(start);
            for (int w = 0; w < 8; ++w)
                out.ep_cbmask.push_back(mask[w]);
            ++n_ep;
        }
        // union of everything ever evaluated for this utterance -> slots
        std::sort(S.uni.begin(), S.uni.end());
        const int n_us = (int)S.uni.size();
        for (int k = 0; k < n_us; ++k)
            S.slot_of[S.uni[k]] = (uint16_t)k;
        out.usen.insert(out.usen.end(), S.uni.begin(), S.uni.end());
        out.us_count.push_back(n_us);
        out.max_union = std::max(out.max_union, n_us);
        for (int e = 0; e < n_ep; ++e) {
            const int32_t a0 = S.ev_off[e], a1 = S.ev_off[e + 1];
            for (int32_t k = a0; k < a1; ++k)
                out.ep_slot.push_back(S.slot_of[S.ev[k]]);
            out.ep_slot_len.push_back(a1 - a0);
            const int32_t end = e + 1 < n_ep ? out.ep_start[ep_first + e + 1] : T;
            const int32_t len = end - out.ep_start[ep_first + e];
            int ncb = 0;
            for (int w = 0; w < 8; ++w)
                ncb += __builtin_popcount(out.ep_cbmask[(ep_first + e) * 8 + w]);
            out.active_sen_frames += (int64_t)len * (a1 - a0);
            out.scanned_cb_frames += (int64_t)len * ncb;
        }
        out.ep_count.push_back(n_ep);
        // chain state -> union slot; states whose phone never becomes active read the
        // always-zero slot n_us (the reference leaves such senone scores at 0 - best)
        for (int q = 0; q < np; ++q)
            for (int j = 0; j < E; ++j) {
                const int s = h.sseq[(size_t)ssid[q] * E + j];
                st_slot[q * E + j] = S.umark[s] ? S.slot_of[s] : (uint16_t)n_us;
            }
        for (uint16_t s : S.act)
            S.mark[s] = 0;
        for (uint16_t s : S.uni)
            S.umark[s] = 0;
        seen[key] = Done{u, rec_ep0, out.ep_start.size() - rec_ep0, rec_sl0, out.ep_slot.size() - rec_sl0,
                         rec_us0, n_us, out.active_sen_frames - rec_act, out.scanned_cb_frames - rec_scan,
                         out.band_state_frames - rec_band};
    }
}
}  // namespace

extern "C" int ssb_batch_upload(ssb_batch_t *b, const ssb_align_in_t *in)
{
    if (!b || !in || in->n_utts < 0 || (in->n_utts > 0 && (!in->frame_off || !in->phone_off))) {
        set_error("ssb_batch_upload: bad arguments");
        return -1;
    }
    if (need_device(b->m) != 0)
        return -1;
    const HostModel &h = b->m->h;
    const int U = in->n_utts, E = h.n_emit;
    b->ran = false;
    b->n_utts = U;
    b->compallsen = in->compallsen ? 1 : 0;
    if (U == 0) {
        b->frame_off.assign(1, 0);
        b->phone_off.assign(1, 0);
    } else {
        b->frame_off.assign(in->frame_off, in->frame_off + U + 1);
        b->phone_off.assign(in->phone_off, in->phone_off + U + 1);
    }
    b->n_frames = b->frame_off[U] - b->frame_off[0];
    b->n_phones = b->phone_off[U] - b->phone_off[0];
    if (b->frame_off[0] != 0 || b->phone_off[0] != 0) {
        set_error("frame_off[0] and phone_off[0] must be 0");
        return -1;
    }
    if (b->n_phones > 0 && (!in->ssid || !in->tmat || !in->sf || !in->ef)) {
        set_error("ssb_batch_upload: chain arrays are NULL");
        return -1;
    }
    b->n_states = b->n_phones * E;
    b->scr_off.assign(U + 1, 0);
    b->max_phones = b->max_union = b->max_T = 0;
    // ---- validation + sizes
    for (int u = 0; u < U; ++u) {
        const int64_t p0 = b->phone_off[u];
        const int64_t Tl = b->frame_off[u + 1] - b->frame_off[u], npl = b->phone_off[u + 1] - p0;
        if (Tl < 0 || npl < 0 || Tl > INT32_MAX - 2 || npl * E > 0x7fffffff) {
            set_error("utterance %d: offsets must be non-decreasing", u);
            return -1;
        }
        const int T = (int)Tl, np = (int)npl;
        b->max_phones = std::max(b->max_phones, np);
        b->max_T = std::max(b->max_T, T);
        b->scr_off[u + 1] = b->scr_off[u] + (int64_t)T * np * E;
        const int32_t *ssid = in->ssid + p0, *tmat = in->tmat + p0, *ef = in->ef + p0;
        for (int i = 0; i < np; ++i) {
            if (ssid[i] < 0 || ssid[i] >= h.n_sseq || tmat[i] < 0 || tmat[i] >= h.n_tmat) {
                set_error("utterance %d phone %d: ssid %d / tmat %d out of range", u, i, ssid[i],
                          tmat[i]);
                return -1;
            }
            if (i > 0 && ef[i] < ef[i - 1]) {
                // the reference's own windows never decrease (phones inherit word windows,
                // ref: src/ps_alignment.c:168-305); a decreasing one would make HMM activity
                // depend on scores, which the planned schedule cannot express
                set_error("utterance %d phone %d: window ends must not decrease along the chain",
                          u, i);
                return -1;
            }
            for (int j = 0; j < E; ++j)
                if (h.sseq[(size_t)ssid[i] * E + j] >= h.n_sen) {
                    set_error("utterance %d phone %d: senone id out of range", u, i);
                    return -1;
                }
        }
    }
    b->n_state_frames = b->scr_off[U];

    // chain scores / tokens: banded (only the frames a phone can be evaluated on) unless the caller
    // wants the reference's dense arrays back (debug outputs) or the scoring path is one that
    // writes dense rows (compallsen gather, the continuous scorer)
    {
        const char *e = getenv("SSB_BANDED");
        b->banded = !b->want_dense && !b->want_tokens_all && !b->compallsen && h.kind != SSB_SCORER_CONT
                    && !(e && *e == '0');
    }
    // ---- device buffers; the feature copy starts now and overlaps the planning below
    cudaStream_t st = b->st;
    // buffers that have to grow go back to the block cache: nothing of the previous run may
    // still be using them
    API_CUDA(cudaStreamSynchronize(st), -1);
    const int CS = h.kind == SSB_SCORER_CONT ? 0 : h.n_mgau * h.n_feat;  // no top-N stage
    const int64_t G = b->n_frames;
    if (b->feat.ensure(std::max<size_t>((size_t)G * h.blk * sizeof(float), 16)) != 0
        || (tc_supported(b->m->d) && b->featp.ensure(std::max<size_t>(tc2_featp_bytes(b->m->d, G), 16)) != 0)
        || b->tn_s.ensure(std::max<size_t>((size_t)G * CS * sizeof(int4), 16)) != 0
        || b->tn_c.ensure(std::max<size_t>((size_t)G * CS * sizeof(uchar4), 16)) != 0
        || b->st_start.ensure(std::max<size_t>((size_t)b->n_states * 4, 16)) != 0
        || b->st_dur.ensure(std::max<size_t>((size_t)b->n_states * 4, 16)) != 0
        || b->st_score.ensure(std::max<size_t>((size_t)b->n_states * 4, 16)) != 0
        || b->utt_rv.ensure(std::max<size_t>((size_t)U * 4, 16)) != 0
        || b->utt_best.ensure(std::max<size_t>((size_t)U * 4, 16)) != 0
        || b->utt_renorm.ensure(std::max<size_t>((size_t)U * 4, 16)) != 0
        || b->fin_hist.ensure(std::max<size_t>((size_t)U * 4, 16)) != 0
        || b->fin_score.ensure(std::max<size_t>((size_t)U * 4, 16)) != 0)
        return -1;
    // chains too long for shared memory keep their HMM state in HBM/L2
    {
        const size_t per_phone = (size_t)(2 * E + 2) * 4;
        const int cap = (int)((200 * 1024) / per_phone);
        b->spill_stride = 0;
        if (b->max_phones > cap) {
            b->spill_stride = (int64_t)b->max_phones * (2 * E + 2);
            if (b->spill.ensure((size_t)b->spill_stride * 4 * U) != 0)
                return -1;
        }
    }
    if (G > 0 && in->feat)
        API_CUDA(cudaMemcpyAsync(b->feat.p, in->feat, (size_t)G * h.blk * sizeof(float),
                                 cudaMemcpyDefault, st), -1);
    else if (G > 0) {
        set_error("ssb_batch_upload: feat is NULL");
        return -1;
    }

    // ---- plan: worker threads over contiguous utterance ranges, then concatenate
    const auto t_plan0 = std::chrono::steady_clock::now();
    b->enter.assign((size_t)b->n_phones, -1);
    std::vector<uint16_t> st_slot((size_t)b->n_states, 0);
    int n_workers = planner_threads();
    n_workers = std::max(1, std::min(n_workers, U / 64));
    std::vector<PlanPiece> pieces(n_workers);
    {
        std::vector<std::thread> th;
        for (int w = 0; w < n_workers; ++w) {
            const int u0 = (int)((int64_t)U * w / n_workers), u1 = (int)((int64_t)U * (w + 1) / n_workers);
            if (n_workers == 1)
                plan_range(h, in, b->frame_off, b->phone_off, u0, u1, b->enter.data(), st_slot.data(), pieces[w]);
            else
                th.emplace_back(plan_range, std::cref(h), in, std::cref(b->frame_off),
                                std::cref(b->phone_off), u0, u1, b->enter.data(), st_slot.data(),
                                std::ref(pieces[w]));
        }
        for (auto &t : th)
            t.join();
    }
    std::vector<int32_t> ep_off(1, 0), ep_start, ep_slot_off(1, 0), us_off(1, 0);
    std::vector<uint32_t> ep_cbmask;
    std::vector<uint16_t> ep_slot, usen;
    b->n_active_sen_frames = b->n_scanned_cb_frames = b->n_band_state_frames = 0;
    for (const PlanPiece &pc : pieces) {
        for (int32_t c : pc.ep_count)
            ep_off.push_back(ep_off.back() + c);
        for (int32_t c : pc.us_count)
            us_off.push_back(us_off.back() + c);
        for (int32_t c : pc.ep_slot_len)
            ep_slot_off.push_back(ep_slot_off.back() + c);
        ep_start.insert(ep_start.end(), pc.ep_start.begin(), pc.ep_start.end());
        ep_cbmask.insert(ep_cbmask.end(), pc.ep_cbmask.begin(), pc.ep_cbmask.end());
        ep_slot.insert(ep_slot.end(), pc.ep_slot.begin(), pc.ep_slot.end());
        usen.insert(usen.end(), pc.usen.begin(), pc.usen.end());
        b->max_union = std::max(b->max_union, pc.max_union);
        b->n_active_sen_frames += pc.active_sen_frames;
        b->n_band_state_frames += pc.band_state_frames;
        b->n_scanned_cb_frames += pc.scanned_cb_frames;
    }
    if (ep_slot.size() > (size_t)INT32_MAX - 65536) {
        set_error("active-list plan too large; split the batch");
        return -1;
    }
    // widest band of phones alive at once (evaluated, or entering at the end of the frame)
    b->max_band = 0;
    for (int u = 0; u < U; ++u) {
        const int64_t p0 = b->phone_off[u];
        const int np = (int)(b->phone_off[u + 1] - p0);
        int lo = 0;
        for (int i = 0; i < np && b->enter[p0 + i] >= 0; ++i) {
            const int32_t t = b->enter[p0 + i] - 1;  // phone i enters at the end of frame t
            while (lo < i && std::max(b->enter[p0 + lo], in->ef[p0 + lo]) < t)
                ++lo;
            b->max_band = std::max(b->max_band, i - lo + 1);
        }
    }
    // ---- chain scores and token stack
    b->n_band_scr = b->n_band_tok = 0;
    std::vector<int64_t> sb, tb;
    if (b->banded) {
        sb.assign((size_t)b->n_phones + 1, 0);
        tb.assign((size_t)b->n_phones + 1, 0);
        int64_t ns = 0, nt = 0;
        for (int u = 0; u < U; ++u) {
            const int64_t p0 = b->phone_off[u];
            const int np = (int)(b->phone_off[u + 1] - p0);
            const int T = (int)(b->frame_off[u + 1] - b->frame_off[u]);
            for (int i = 0; i < np; ++i) {
                const int32_t en = b->enter[p0 + i];
                sb[p0 + i] = ns;
                tb[p0 + i] = nt;
                if (en < 0 || T == 0)
                    continue;
                const int64_t last = std::min<int64_t>(std::max(en, in->ef[p0 + i]), T - 1);
                const int64_t first = std::max(en - 1, 0);
                sb[p0 + i] = ns - (int64_t)en * E;   // virtual bases: element (t, j) at base + t * E + j
                tb[p0 + i] = nt - first * E;
                if (en <= T - 1)
                    ns += (last - en + 1) * E;
                if (last >= first)
                    nt += (last - first + 1) * E;
            }
        }
        b->n_band_scr = ns;
        b->n_band_tok = nt;
        if (upload(b->d_scr_boff, sb, st) || upload(b->d_tok_boff, tb, st))
            return -1;
    }
    {
        const size_t n_scr = b->banded ? (size_t)b->n_band_scr : (size_t)b->n_state_frames;
        const size_t n_tok = b->banded ? (size_t)b->n_band_tok : (size_t)b->n_state_frames;
        if (b->chain_scr.ensure(std::max<size_t>(n_scr * 2, 16)) != 0
            || b->tokens.ensure(std::max<size_t>(n_tok * sizeof(int2), 16)) != 0)
            return -1;
    }
    std::vector<int32_t> v_ssid(in->ssid, in->ssid + b->n_phones),
        v_tmat(in->tmat, in->tmat + b->n_phones), v_sf(in->sf, in->sf + b->n_phones),
        v_ef(in->ef, in->ef + b->n_phones);
    if (upload(b->d_frame_off, b->frame_off, st) || upload(b->d_phone_off, b->phone_off, st)
        || upload(b->d_scr_off, b->scr_off, st) || upload(b->d_ssid, v_ssid, st)
        || upload(b->d_tmat, v_tmat, st) || upload(b->d_sf, v_sf, st) || upload(b->d_ef, v_ef, st)
        || upload(b->d_ep_off, ep_off, st) || upload(b->d_ep_start, ep_start, st)
        || upload(b->d_ep_cbmask, ep_cbmask, st) || upload(b->d_ep_slot_off, ep_slot_off, st)
        || upload(b->d_ep_slot, ep_slot, st) || upload(b->d_us_off, us_off, st)
        || upload(b->d_usen, usen, st) || upload(b->d_st_slot, st_slot, st)
        || upload(b->d_enter, b->enter, st))
        return -1;
    b->plan_us = (int64_t)std::chrono::duration<double, std::micro>(
                     std::chrono::steady_clock::now() - t_plan0).count();
    // the host vectors above die at return: make sure the copies have been staged
    API_CUDA(cudaStreamSynchronize(st), -1);
    DevPlan &p = b->plan;
    p.n_utts = U;
    p.frame_off = b->d_frame_off.as<int64_t>();
    p.phone_off = b->d_phone_off.as<int64_t>();
    p.scr_off = b->d_scr_off.as<int64_t>();
    p.ssid = b->d_ssid.as<int32_t>();
    p.tmat = b->d_tmat.as<int32_t>();
    p.sf = b->d_sf.as<int32_t>();
    p.ef = b->d_ef.as<int32_t>();
    p.ep_off = b->d_ep_off.as<int32_t>();
    p.ep_start = b->d_ep_start.as<int32_t>();
    p.ep_cbmask = b->d_ep_cbmask.as<uint32_t>();
    p.ep_slot_off = b->d_ep_slot_off.as<int32_t>();
    p.ep_slot = b->d_ep_slot.as<uint16_t>();
    p.us_off = b->d_us_off.as<int32_t>();
    p.usen = b->d_usen.as<uint16_t>();
    p.st_slot = b->d_st_slot.as<uint16_t>();
    p.enter_plan = b->d_enter.as<int32_t>();
    p.all_active = b->compallsen;
    p.banded = b->banded ? 1 : 0;
    p.scr_boff = b->banded ? b->d_scr_boff.as<int64_t>() : nullptr;
    p.tok_boff = b->banded ? b->d_tok_boff.as<int64_t>() : nullptr;
    p.seg_t0 = nullptr;
    p.tie_bits = nullptr;
    p.tie_w = 0;
    p.init_topn = nullptr;
    if (in->init_topn && U > 0 && CS > 0) {
        // the top-N lists the scorer carries in from a previous pass on the same decoder
        if (b->d_init_topn.ensure((size_t)U * CS * 4) != 0)
            return -1;
        API_CUDA(cudaMemcpyAsync(b->d_init_topn.p, in->init_topn, (size_t)U * CS * 4,
                                 cudaMemcpyDefault, st), -1);
        API_CUDA(cudaStreamSynchronize(st), -1);
        p.init_topn = b->d_init_topn.as<uchar4>();
    }

    // ---- chain cutting.  Where one word window ends exactly where the next begins, the chain can
    // only be crossed on that one frame (prune_hmms / phone_transition, ref: src/
    // state_align_search.c:88-133): what happens after the cut depends on what happened before
    // only through the score the first phone is entered with, and integer max-plus arithmetic
    // is shift invariant -- the phones after the cut see that score as their zero, exactly as
    // the first phone of an utterance does (hmm_enter(hmms, 0, 0, 0), ref :46-55).  So K3 and the
    // backtrace run on the SEGMENTS, one warp each, instead of one warp walking 360 000 frames of
    // an hour-long chain; state scores are differences and come out as they are, start frames are
    // shifted back by the backtrace, the utterance's best score is the sum of the segments' exit
    // scores.  Absolute scores matter to the reference in one place, the renormalisation at
    // best - 0x300000 < WORST_SCORE (ref :193-197): path scores only decrease, so a total above
    // that line proves no frame was renormalised; otherwise ssb_batch_download runs the uncut
    // kernels again.  Only with the banded layout (the dense debug stacks hold absolute scores),
    // and by default for chains of more than 128 phones and for batches of fewer than 512 utterances
    // ($SSB_K3_CUT = all | 0).
    b->cut = false;
    b->n_segs = 0;
    {
        const char *e = getenv("SSB_K3_CUT");
        const bool cut_all = e && strcmp(e, "all") == 0, cut_off = e && *e == '0';
        const bool few = U < 512;  // a batch that leaves the GPU idle: every cut adds a warp of work in flight
        if (b->banded && !cut_off && U > 0 && b->n_phones > 0 && (cut_all || few || b->max_phones > 128)) {
            std::vector<int64_t> sfo(1, b->frame_off[0]), spo(1, 0), sbo2(sb.begin(), sb.end()), tbo2(tb.begin(), tb.end());
            std::vector<int32_t> st0, en2(b->enter), sf2(in->sf, in->sf + b->n_phones), ef2(in->ef, in->ef + b->n_phones);
            b->seg_off.assign(1, 0);
            b->cut_max_phones = 0;
            auto shift = [&](int64_t pa, int64_t pb, int32_t t0) {
                if (t0 == 0)
                    return;
                for (int64_t j = pa; j < pb; ++j) {
                    if (en2[j] >= 0)
                        en2[j] -= t0;
                    sf2[j] = (int32_t)std::max<int64_t>((int64_t)sf2[j] - t0, INT32_MIN / 2);
                    if (ef2[j] < INT32_MAX / 2)
                        ef2[j] = (int32_t)std::max<int64_t>((int64_t)ef2[j] - t0, INT32_MIN / 2);
                    sbo2[j] += (int64_t)t0 * E;
                    tbo2[j] += (int64_t)t0 * E;
                }
            };
            for (int u = 0; u < U; ++u) {
                const int64_t p0 = b->phone_off[u];
                const int np = (int)(b->phone_off[u + 1] - p0);
                const int T = (int)(b->frame_off[u + 1] - b->frame_off[u]);
                int a = 0;
                int32_t t0 = 0;
                if (cut_all || few || np > 128)
                    for (int i = 1; i < np; ++i) {
                        const int32_t s = in->sf[p0 + i];
                        if (s > t0 && s < T && b->enter[p0 + i] == s && in->ef[p0 + i - 1] == s
                            && b->enter[p0 + i - 1] >= 0 && b->enter[p0 + i - 1] < s) {
                            shift(p0 + a, p0 + i, t0);
                            st0.push_back(t0);
                            sfo.push_back(b->frame_off[u] + s);
                            spo.push_back(p0 + i);
                            b->cut_max_phones = std::max(b->cut_max_phones, i - a);
                            a = i;
                            t0 = s;
                        }
                    }
                shift(p0 + a, p0 + np, t0);
                st0.push_back(t0);
                sfo.push_back(b->frame_off[u + 1]);
                spo.push_back(p0 + np);
                b->cut_max_phones = std::max(b->cut_max_phones, np - a);
                b->seg_off.push_back((int32_t)st0.size());
            }
            const int S = (int)st0.size();
            // (a segment too long for shared memory would index the per-utterance spill area by its
            // segment number: such batches stay uncut)
            const int smem_cap = (int)((200 * 1024) / ((size_t)(2 * E + 2) * 4));
            if (S > U && b->cut_max_phones <= smem_cap) {
                std::vector<int64_t> zero((size_t)S + 1, 0);
                if (upload(b->d_seg_frame_off, sfo, st) || upload(b->d_seg_phone_off, spo, st)
                    || upload(b->d_seg_scr_off, zero, st) || upload(b->d_seg_enter, en2, st)
                    || upload(b->d_seg_sf, sf2, st) || upload(b->d_seg_ef, ef2, st)
                    || upload(b->d_seg_scr_boff, sbo2, st) || upload(b->d_seg_tok_boff, tbo2, st)
                    || upload(b->d_seg_t0, st0, st) || b->seg_rv.ensure((size_t)S * 4) != 0
                    || b->seg_best.ensure((size_t)S * 4) != 0 || b->seg_renorm.ensure((size_t)S * 4) != 0
                    || b->seg_fin_hist.ensure((size_t)S * 4) != 0 || b->seg_fin_score.ensure((size_t)S * 4) != 0)
                    return -1;
                API_CUDA(cudaStreamSynchronize(st), -1);
                b->cut = true;
                b->n_segs = S;
                b->cut_max_band = std::min(b->max_band, b->cut_max_phones);
                b->cut_plan = p;
                b->cut_plan.n_utts = S;
                b->cut_plan.frame_off = b->d_seg_frame_off.as<int64_t>();
                b->cut_plan.phone_off = b->d_seg_phone_off.as<int64_t>();
                b->cut_plan.scr_off = b->d_seg_scr_off.as<int64_t>();
                b->cut_plan.enter_plan = b->d_seg_enter.as<int32_t>();
                b->cut_plan.sf = b->d_seg_sf.as<int32_t>();
                b->cut_plan.ef = b->d_seg_ef.as<int32_t>();
                b->cut_plan.scr_boff = b->d_seg_scr_boff.as<int64_t>();
                b->cut_plan.tok_boff = b->d_seg_tok_boff.as<int64_t>();
                b->cut_plan.seg_t0 = b->d_seg_t0.as<int32_t>();
            }
        }
    }

    // ---- K1 over time: with few utterances the top-N kernel (thread = utterance) has no rows
    // to fill its CTAs with, so long utterances are cut into segments that are scored
    // independently; the steps whose result depends on the carried list (ties) are replayed
    // afterwards (topn_fixup.cu).  $SSB_K1_SEG forces a segment length (tests).
    b->k1_seg = false;
    b->n_seg_utts = 0;
    b->ft = false;
    b->n_tiles = 0;
    if (k1_frame_tiled(b->m->d, b->compallsen != 0, p.init_topn != nullptr, U) && U > 0 && G > 0) {
        // frame-tiled K1: every utterance is cut into tiles of 128 frames that are scored
        // independently; all utterances are candidates for the tie fix-up
        std::vector<int32_t> tu, tt, all_utts(U);
        for (int u = 0; u < U; ++u) {
            all_utts[u] = u;
            const int T = (int)(b->frame_off[u + 1] - b->frame_off[u]);
            for (int t0 = 0; t0 < T; t0 += 128) {
                tu.push_back(u);
                tt.push_back(t0);
            }
        }
        if (getenv("SSB_FT_DIAG") && !b->compallsen) {
            // row utilisation of the frame-tiled kernel vs a codebook-major arrangement
            double own = 0, tile_rows = 0, warp_rows = 0, piece_rows = 0, n_runs = 0;
            for (int u = 0; u < U; ++u) {
                const int T = (int)(b->frame_off[u + 1] - b->frame_off[u]);
                std::vector<uint64_t> fm(T, 0);
                for (int e = ep_off[u]; e < ep_off[u + 1]; ++e) {
                    const int s = ep_start[e], en = e + 1 < ep_off[u + 1] ? ep_start[e + 1] : T;
                    const uint64_t mk = (uint64_t)ep_cbmask[(size_t)e * 8] | ((uint64_t)ep_cbmask[(size_t)e * 8 + 1] << 32);
                    for (int t = std::max(s, 0); t < en && t < T; ++t)
                        fm[t] = mk;
                }
                for (int t = 0; t < T; ++t)
                    own += __builtin_popcountll(fm[t]);
                for (int t0 = 0; t0 < T; t0 += 128) {
                    uint64_t un = 0;
                    for (int t = t0; t < std::min(T, t0 + 128); ++t)
                        un |= fm[t];
                    tile_rows += 128.0 * __builtin_popcountll(un);
                    for (int w = 0; w < 4; ++w) {
                        uint64_t wn = 0;
                        for (int t = t0 + 32 * w; t < std::min(T, t0 + 32 * w + 32); ++t)
                            wn |= fm[t];
                        warp_rows += 32.0 * __builtin_popcountll(wn);
                    }
                }
                for (int c = 0; c < 64; ++c) {
                    int run = 0;
                    for (int t = 0; t <= T; ++t) {
                        const bool on = t < T && ((fm[t] >> c) & 1u);
                        if (on)
                            ++run;
                        else if (run) {
                            piece_rows += 32.0 * ((run + 31) / 32);
                            n_runs += 1;
                            run = 0;
                        }
                    }
                }
            }
            fprintf(stderr, "[ft diag] frames %lld own cb-rows %.0f (%.2f/frame) tile-union rows %.0f (util %.3f) "
                            "warp-union rows %.0f (util %.3f) cb-major 32-row pieces %.0f (util %.3f) runs %.0f (mean %.1f)\n",
                    (long long)G, own, own / (double)G, tile_rows, own / tile_rows, warp_rows, own / warp_rows,
                    piece_rows, own / piece_rows, n_runs, own / n_runs);
        }
        b->n_tiles = (int)tu.size();
        b->k1_tie_w = (G + 31) / 32 + 1;
        if (b->n_tiles > 0) {
            if (upload(b->d_tile_utt, tu, st) || upload(b->d_tile_t0, tt, st)
                || upload(b->d_seg_utts, all_utts, st) || b->d_tile_ctr.ensure(16) != 0
                || b->d_k1_tie.ensure((size_t)CS * b->k1_tie_w * 4 + 16) != 0)
                return -1;
            API_CUDA(cudaStreamSynchronize(st), -1);
            b->ft = true;
            b->n_seg_utts = U;
        }
    }
    if (!b->ft) {
        const char *force = getenv("SSB_K1_SEG");
        const char *k1 = getenv("SSB_K1");
        int64_t seg = 0;
        if (tc_supported(b->m->d) && (!(k1 && *k1) || strcmp(k1, "tc2") == 0) && h.cfg.ds <= 1 && U > 0 && G > 0) {
            if (force && atoll(force) > 0)
                seg = atoll(force);
            else if (U <= 512 && b->max_T >= 8192)
                seg = std::min<int64_t>(4096, std::max<int64_t>(512, G / 2048));
        }
        if (seg > 0) {
            std::vector<int64_t> kfo(1, 0);
            std::vector<int32_t> kep_off(1, 0), kep_start, seg_utts;
            std::vector<uint32_t> kep_mask;
            bool any_cut = false;
            for (int u = 0; u < U; ++u) {
                const int64_t f0 = b->frame_off[u];
                const int T = (int)(b->frame_off[u + 1] - f0);
                const bool cut = T > 2 * seg;
                // (rows of the segmented launch start from the initial lists: with lists carried in
                // from a previous pass the uncut utterances need their tie steps replayed as well)
                any_cut = any_cut || cut;
                if (cut || p.init_topn)
                    seg_utts.push_back(u);
                const int64_t step = cut ? seg : std::max(T, 1);
                for (int64_t s0 = 0; s0 < std::max(T, 1); s0 += step) {
                    const int64_t s1 = std::min<int64_t>(T, s0 + step);
                    kfo.push_back(f0 + s1);
                    if (!b->compallsen) {
                        // the parent's epoch in force at s0, then those starting inside the segment
                        int last = -1;
                        for (int e = ep_off[u]; e < ep_off[u + 1] && ep_start[e] <= s0; ++e)
                            last = e;
                        if (last >= 0) {
                            kep_start.push_back(0);
                            kep_mask.insert(kep_mask.end(), ep_cbmask.begin() + (size_t)last * 8,
                                            ep_cbmask.begin() + (size_t)last * 8 + 8);
                        }
                        for (int e = std::max(last + 1, ep_off[u]); e < ep_off[u + 1]; ++e)
                            if (ep_start[e] > s0 && ep_start[e] < s1) {
                                kep_start.push_back((int32_t)(ep_start[e] - s0));
                                kep_mask.insert(kep_mask.end(), ep_cbmask.begin() + (size_t)e * 8,
                                                ep_cbmask.begin() + (size_t)e * 8 + 8);
                            }
                    }
                    kep_off.push_back((int32_t)kep_start.size());
                }
            }
            if (any_cut) {
                if (kep_start.empty()) {  // keep the uploads non-empty
                    kep_start.push_back(0);
                    kep_mask.assign(8, 0u);
                }
                b->k1_tie_w = (G + 31) / 32 + 1;
                if (upload(b->d_k1_frame_off, kfo, st) || upload(b->d_k1_ep_off, kep_off, st)
                    || upload(b->d_k1_ep_start, kep_start, st) || upload(b->d_k1_ep_cbmask, kep_mask, st)
                    || upload(b->d_seg_utts, seg_utts, st)
                    || b->d_k1_tie.ensure((size_t)CS * b->k1_tie_w * 4 + 16) != 0)
                    return -1;
                API_CUDA(cudaStreamSynchronize(st), -1);
                b->k1_seg = true;
                b->n_seg_utts = (int)seg_utts.size();
                b->n_k1_rows = (int)kfo.size() - 1;
                b->k1_plan = p;
                b->k1_plan.init_topn = nullptr;  // rows are segments; ties are replayed by the fix-up
                b->k1_plan.n_utts = b->n_k1_rows;
                b->k1_plan.frame_off = b->d_k1_frame_off.as<int64_t>();
                b->k1_plan.ep_off = b->d_k1_ep_off.as<int32_t>();
                b->k1_plan.ep_start = b->d_k1_ep_start.as<int32_t>();
                b->k1_plan.ep_cbmask = b->d_k1_ep_cbmask.as<uint32_t>();
            }
        }
    }
    return 0;
}

// K1 for the uploaded batch.  `tie` (optional, zeroed by the caller, [CS][tie_w] words) receives
// the tie flags; when the batch's long utterances were cut into segments the flagged steps are
// replayed afterwards unless the caller does that itself (`fixup` = false: the grammar search in
// its default mode never reads them).
static int batch_topn(ssb_batch_t *b, uint32_t *tie, int64_t tie_w, bool fixup, int exact = 0,
                      TcDebug dbg = TcDebug{nullptr, nullptr, nullptr})
{
    const DevModel &d = b->m->d;
    cudaStream_t st = b->st;
    const int64_t G = b->n_frames;
    if (b->n_utts == 0 || G == 0)
        return 0;
    DevPlan p = b->k1_seg ? b->k1_plan : b->plan;
    if ((b->k1_seg || b->ft) && !tie) {
        tie = b->d_k1_tie.as<uint32_t>();
        tie_w = b->k1_tie_w;
        if (cudaMemsetAsync(tie, 0, (size_t)d.n_mgau * d.n_feat * tie_w * 4, st) != cudaSuccess) {
            set_error("cudaMemsetAsync failed");
            return -1;
        }
    }
    p.tie_bits = tie;
    p.tie_w = tie_w;
    if (b->ft) {
        if (launch_gmm_scan_ft(d, p, b->feat.as<float>(), G, b->tn_s.as<int4>(), b->tn_c.as<uchar4>(),
                               b->d_tile_utt.as<int32_t>(), b->d_tile_t0.as<int32_t>(), b->n_tiles,
                               b->d_tile_ctr.as<int>(), exact, dbg, st) != 0)
            return -1;
    } else if (launch_gmm_topn(d, p, b->feat.as<float>(), G, b->tn_s.as<int4>(), b->tn_c.as<uchar4>(),
                               b->featp.as<float>(), st) != 0)
        return -1;
    if ((b->k1_seg || b->ft) && fixup)
        return launch_topn_fixup(d, b->plan, b->d_seg_utts.as<int32_t>(), b->n_seg_utts,
                                 b->feat.as<float>(), G, b->tn_s.as<int4>(), b->tn_c.as<uchar4>(),
                                 tie, tie_w, st);
    return 0;
}

// K3 + backtrace of the uploaded batch, on the utterances or (cut) on their segments
static int run_chain(ssb_batch_t *b, bool cut, bool timed)
{
    const DevModel &d = b->m->d;
    cudaStream_t st = b->st;
    const int U = b->n_utts;
    const DevPlan &p = cut ? b->cut_plan : b->plan;
    if (U > 0) {
        if (b->want_tokens_all && b->n_state_frames > 0)
            API_CUDA(cudaMemsetAsync(b->tokens.p, 0xff, (size_t)b->n_state_frames * sizeof(int2), st), -1);
        if (launch_chain_viterbi(d, p, b->chain_scr.as<int16_t>(), b->tokens.as<int2>(),
                                 b->spill.as<int32_t>(), b->spill_stride,
                                 (cut ? b->seg_best : b->utt_best).as<int32_t>(),
                                 (cut ? b->seg_renorm : b->utt_renorm).as<int32_t>(),
                                 (cut ? b->seg_fin_hist : b->fin_hist).as<int32_t>(),
                                 (cut ? b->seg_fin_score : b->fin_score).as<int32_t>(),
                                 cut ? b->cut_max_phones : b->max_phones, cut ? b->cut_max_band : b->max_band,
                                 st) != 0)
            return -1;
    }
    if (timed)
        API_CUDA(cudaEventRecord(b->ev[3], st), -1);
    if (U > 0) {
        // duration -1 marks "state not on the best path"
        if (b->n_states > 0) {
            API_CUDA(cudaMemsetAsync(b->st_dur.p, 0xff, (size_t)b->n_states * 4, st), -1);
            // (start frames of states off the path are never read back into the caller's arrays;
            // zeroed so that the download copies initialised memory)
            API_CUDA(cudaMemsetAsync(b->st_start.p, 0, (size_t)b->n_states * 4, st), -1);
            // 0x80808080 marks "score never written" (the first state of an utterance)
            API_CUDA(cudaMemsetAsync(b->st_score.p, 0x80, (size_t)b->n_states * 4, st), -1);
        }
        if (launch_backtrace(d, p, b->tokens.as<int2>(), (cut ? b->seg_fin_hist : b->fin_hist).as<int32_t>(),
                             (cut ? b->seg_fin_score : b->fin_score).as<int32_t>(), b->st_start.as<int32_t>(),
                             b->st_dur.as<int32_t>(), b->st_score.as<int32_t>(),
                             (cut ? b->seg_rv : b->utt_rv).as<int32_t>(), st) != 0)
            return -1;
    }
    return 0;
}

// frames per dense slab in compallsen mode (whole utterances)
static const int64_t kSlabFrames = 32768;

extern "C" int ssb_batch_run(ssb_batch_t *b)
{
    if (!b) {
        set_error("NULL batch");
        return -1;
    }
    if (need_device(b->m) != 0)
        return -1;
    const DevModel &d = b->m->d;
    const DevPlan &p = b->plan;
    cudaStream_t st = b->st;
    const int U = b->n_utts;
    b->n_launches = 0;
    launch_count(true);
    API_CUDA(cudaEventRecord(b->ev[0], st), -1);
    if (U > 0 && b->n_frames > 0) {
        if (batch_topn(b, nullptr, 0, true) != 0)
            return -1;
    }
    API_CUDA(cudaEventRecord(b->ev[1], st), -1);
    if (U > 0 && b->n_frames > 0 && b->n_states > 0) {
        if (!b->compallsen && d.kind == SSB_SCORER_CONT) {
            const int W = std::max(b->max_union, 1);
            if (b->dense.ensure((size_t)b->n_frames * W * 2) != 0
                || launch_cont_active(d, p, b->feat.as<float>(), W, b->max_T,
                                      b->dense.as<int16_t>(), b->chain_scr.as<int16_t>(), st) != 0)
                return -1;
        } else if (!b->compallsen) {
            if (launch_senone_mix_active(d, p, b->tn_s.as<int4>(), b->tn_c.as<uchar4>(),
                                         b->n_frames, b->max_union + 1, b->max_T,
                                         b->chain_scr.as<int16_t>(), st) != 0)
                return -1;
        } else {
            int u0 = 0;
            while (u0 < U) {
                int u1 = u0 + 1;
                while (u1 < U && b->frame_off[u1 + 1] - b->frame_off[u0] <= kSlabFrames)
                    ++u1;
                const int64_t g0 = b->frame_off[u0], n = b->frame_off[u1] - g0;
                if (n > 0) {
                    if (b->dense.ensure((size_t)n * d.n_sen * 2) != 0)
                        return -1;
                    if (dense_scores(b, g0, n, b->dense.as<int16_t>()) != 0
                        || launch_gather_chain(d, p, b->dense.as<int16_t>(), u0, u1, g0,
                                               b->chain_scr.as<int16_t>(), st) != 0)
                        return -1;
                }
                u0 = u1;
            }
        }
    }
    API_CUDA(cudaEventRecord(b->ev[2], st), -1);
    if (run_chain(b, b->cut, true) != 0)
        return -1;
    b->cut_ran = b->cut;
    b->n_launches = launch_count(true);
    API_CUDA(cudaEventRecord(b->ev[4], st), -1);
    b->ran = true;
    return 0;
}

extern "C" int ssb_batch_download(ssb_batch_t *b, ssb_align_out_t *out)
{
    if (!b || !out) {
        set_error("ssb_batch_download: bad arguments");
        return -1;
    }
    if (!b->ran) {
        set_error("ssb_batch_download: ssb_batch_run has not been called on this upload");
        return -1;
    }
    if (need_device(b->m) != 0)
        return -1;
    cudaStream_t st = b->st;
    const int U = b->n_utts;
    const size_t ns = (size_t)b->n_states;
    // ---- cut chains: the utterances' verdicts from their segments'
    std::vector<int32_t> c_rv, c_best;
    if (b->cut_ran && U > 0) {
        const int S = b->n_segs;
        std::vector<int32_t> srv(S), sbest(S), sfin(S);
        API_CUDA(cudaMemcpyAsync(srv.data(), b->seg_rv.p, (size_t)S * 4, cudaMemcpyDeviceToHost, st), -1);
        API_CUDA(cudaMemcpyAsync(sbest.data(), b->seg_best.p, (size_t)S * 4, cudaMemcpyDeviceToHost, st), -1);
        API_CUDA(cudaMemcpyAsync(sfin.data(), b->seg_fin_score.p, (size_t)S * 4, cudaMemcpyDeviceToHost, st), -1);
        API_CUDA(cudaStreamSynchronize(st), -1);
        c_rv.assign(U, 0);
        c_best.assign(U, 0);
        bool renorm = false;
        for (int u = 0; u < U; ++u) {
            int64_t total = 0;
            for (int s = b->seg_off[u]; s < b->seg_off[u + 1]; ++s) {
                if (srv[s] != 0)
                    c_rv[u] = -1;
                if (s + 1 < b->seg_off[u + 1])
                    total += sfin[s];
            }
            const int last = b->seg_off[u + 1] - 1;
            c_best[u] = (int32_t)(total + sbest[last]);
            // no frame of the utterance saw best - 0x300000 < WORST_SCORE: scores along a path only
            // decrease and every frame's best is at least the final path's score
            if (c_rv[u] == 0 && total + sfin[last] - 0x300000 < (int64_t)WORST_SCORE)
                renorm = true;
            // a segment that cannot be crossed: the utterance fails ("Failed to reach final state
            // in alignment"); what the reference leaves behind then (best score of the last
            // frame, entries written before the backtrace gave up) comes from the uncut kernels
            if (c_rv[u] != 0)
                renorm = true;
        }
        if (renorm) {
            // the reference renormalised somewhere (practically unreachable: ~10^8 frames) or an
            // utterance failed: the uncut kernels reproduce either
            if (run_chain(b, false, false) != 0)
                return -1;
            b->cut_ran = false;
            c_rv.clear();
            c_best.clear();
        }
    }
    std::vector<int32_t> s_start, s_dur, s_score;
    if (ns && (out->st_start || out->st_dur || out->st_score)) {
        s_start.resize(ns);
        s_dur.resize(ns);
        s_score.resize(ns);
        API_CUDA(cudaMemcpyAsync(s_start.data(), b->st_start.p, ns * 4, cudaMemcpyDeviceToHost, st), -1);
        API_CUDA(cudaMemcpyAsync(s_dur.data(), b->st_dur.p, ns * 4, cudaMemcpyDeviceToHost, st), -1);
        API_CUDA(cudaMemcpyAsync(s_score.data(), b->st_score.p, ns * 4, cudaMemcpyDeviceToHost, st), -1);
    }
    if (b->cut_ran) {
        for (int u = 0; u < U; ++u) {
            if (out->utt_rv)
                out->utt_rv[u] = c_rv[u];
            if (out->utt_best)
                out->utt_best[u] = c_best[u];
            if (out->utt_renorm)
                out->utt_renorm[u] = 0;
        }
    } else {
        if (U && out->utt_rv)
            API_CUDA(cudaMemcpyAsync(out->utt_rv, b->utt_rv.p, (size_t)U * 4, cudaMemcpyDeviceToHost, st), -1);
        if (U && out->utt_best)
            API_CUDA(cudaMemcpyAsync(out->utt_best, b->utt_best.p, (size_t)U * 4, cudaMemcpyDeviceToHost, st), -1);
        if (U && out->utt_renorm)
            API_CUDA(cudaMemcpyAsync(out->utt_renorm, b->utt_renorm.p, (size_t)U * 4, cudaMemcpyDeviceToHost, st), -1);
    }
    if (b->banded && b->n_state_frames && (out->chain_scr || out->tokens)) {
        set_error("dense chain scores / tokens were not kept (banded layout): call "
                  "ssb_batch_debug_tokens(b, 1) before ssb_batch_upload");
        return -1;
    }
    if (b->n_state_frames && out->chain_scr)
        API_CUDA(cudaMemcpyAsync(out->chain_scr, b->chain_scr.p, (size_t)b->n_state_frames * 2,
                                 cudaMemcpyDeviceToHost, st), -1);
    if (b->n_state_frames && out->tokens) {
        if (!b->want_tokens_all) {
            set_error("token stack requested but the batch was not run with debug tokens "
                      "(call ssb_batch_debug_tokens(b, 1) before ssb_batch_run)");
            return -1;
        }
        API_CUDA(cudaMemcpyAsync(out->tokens, b->tokens.p, (size_t)b->n_state_frames * sizeof(int2),
                                 cudaMemcpyDeviceToHost, st), -1);
    }
    API_CUDA(cudaStreamSynchronize(st), -1);
    // states off the best path keep the caller's values, like the reference's alignment
    // entries keep what alignment_populate put there (ref: src/state_align_search.c:236-263)
    for (size_t i = 0; i < s_dur.size(); ++i) {
        if (s_dur[i] < 0)
            continue;
        if (out->st_start)
            out->st_start[i] = s_start[i];
        if (out->st_dur)
            out->st_dur[i] = s_dur[i];
        // the first state of an utterance keeps its score (ref :256-261)
        if (out->st_score && s_score[i] != (int32_t)0x80808080)
            out->st_score[i] = s_score[i];
    }
    return 0;
}

extern "C" int ssb_batch_debug_tokens(ssb_batch_t *b, int on)
{
    if (!b)
        return -1;
    // bit 0: dense token stack, pre-filled with the reference's 0xff (downloadable);
    // bit 1: dense chain scores only.  Either keeps the reference's dense [T][states] layout.
    b->want_tokens_all = (on & 1) != 0;
    b->want_dense = on != 0;
    return 0;
}

extern "C" int ssb_batch_kernel_ms(ssb_batch_t *b, float *ms)
{
    if (!b || !ms || !b->ran) {
        set_error("ssb_batch_kernel_ms: no completed run");
        return -1;
    }
    API_CUDA(cudaEventSynchronize(b->ev[4]), -1);
    for (int i = 0; i < 8; ++i)
        ms[i] = 0.f;
    for (int i = 0; i < 4; ++i)
        API_CUDA(cudaEventElapsedTime(&ms[i], b->ev[i], b->ev[i + 1]), -1);
    API_CUDA(cudaEventElapsedTime(&ms[4], b->ev[0], b->ev[4]), -1);
    return 0;
}

extern "C" int ssb_batch_n_launches(const ssb_batch_t *b) { return b ? b->n_launches : -1; }

extern "C" int ssb_batch_stats(const ssb_batch_t *b, int64_t *o)
{
    if (!b || !o)
        return -1;
    for (int i = 0; i < 8; ++i)
        o[i] = 0;
    o[0] = b->n_frames;
    o[1] = b->n_state_frames;
    o[2] = b->n_active_sen_frames;
    o[3] = b->n_scanned_cb_frames;
    o[4] = (int64_t)b->bytes_held();
    o[5] = b->max_union;
    o[6] = b->max_phones;
    o[7] = b->plan_us;
    return 0;
}

extern "C" int32_t ssb_batch_n_segments(const ssb_batch_t *b)
{
    return b ? (b->cut ? b->n_segs : b->n_utts) : -1;
}

extern "C" int64_t ssb_batch_band_state_frames(const ssb_batch_t *b)
{
    return b ? b->n_band_state_frames : -1;
}

// ------------------------------------------------------------------ pipeline
// Utterances are independent, so a large batch is cut into chunks of whole utterances and the
// chunks travel through `n_lanes` batches, each on its own stream and driven by its own worker
// thread: while one chunk is on the SMs the next one is being planned and copied in and the
// previous one copied out.  Results are those of one big batch, utterance by utterance.
// Batches are submitted and collected separately, so consecutive batches overlap as well: with
// whole batches as chunks and kernels kept apart (overlap_kernels = 0) batch i+1 is planned
// and uploaded while batch i computes.
namespace {
struct PipeJob {
    int64_t ticket = -1, seq = -1;
    int chunk = 0, u0 = 0;
    ssb_align_in_t in;
    ssb_align_out_t out;
    std::vector<int64_t> fo, po;
};
struct PipeTicket {
    int n_chunks = 0, n_done = 0, launches = 0;
    bool failed = false;
    std::string err;
    std::vector<double> trace;
    std::chrono::steady_clock::time_point t0;
};
}  // namespace

struct ssb_pipeline_s {
    ssb_model_t *m = nullptr;
    int n_lanes = 0;
    int64_t chunk_frames = 0;
    bool overlap_kernels = true;
    std::vector<ssb_batch_t *> lane;
    std::vector<cudaStream_t> st;
    std::vector<std::thread> workers;
    std::mutex mu, compute_mu;
    std::condition_variable cv_job, cv_done, cv_up;
    std::deque<std::unique_ptr<PipeJob>> jobs;
    std::map<int64_t, PipeTicket> tickets;
    int64_t next_ticket = 0, next_seq = 0, up_turn = 0;
    bool stop = false;
    // of the last collected batch
    int n_launches = 0, n_chunks = 0;
    std::vector<double> trace;
};

static void pipe_worker(ssb_pipeline_s *p, int li)
{
    cudaSetDevice(p->m->device);
    ssb_batch_t *b = p->lane[li];
    for (;;) {
        std::unique_ptr<PipeJob> job;
        {
            std::unique_lock<std::mutex> lk(p->mu);
            p->cv_job.wait(lk, [&] { return p->stop || !p->jobs.empty(); });
            if (p->jobs.empty())
                return;  // stop requested and nothing left
            job = std::move(p->jobs.front());
            p->jobs.pop_front();
        }
        double tr[8] = {(double)li, (double)job->u0, 0, 0, 0, 0, 0, 0};
        std::chrono::steady_clock::time_point t0;
        bool skip;
        {
            std::unique_lock<std::mutex> lk(p->mu);
            PipeTicket &tk = p->tickets[job->ticket];
            t0 = tk.t0;
            skip = tk.failed;
        }
        auto now_ms = [&]() {
            return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        };
        int rv = 0;
        {
            // one chunk on the PCIe link at a time, in submission order: the first chunk
            // reaches the SMs after 1/n of the copy time and the lanes stay out of phase
            std::unique_lock<std::mutex> lk(p->mu);
            p->cv_up.wait(lk, [&] { return p->up_turn == job->seq; });
            lk.unlock();
            tr[2] = now_ms();
            if (!skip) {
                ssb_batch_debug_tokens(b, (job->out.tokens ? 1 : 0) | (job->out.chain_scr ? 2 : 0));
                rv = ssb_batch_upload(b, &job->in);
            }
            lk.lock();
            ++p->up_turn;
            lk.unlock();
            p->cv_up.notify_all();
        }
        tr[3] = now_ms();
        if (!skip && rv == 0) {
            if (p->overlap_kernels) {
                rv = ssb_batch_run(b);
            } else {
                std::lock_guard<std::mutex> g(p->compute_mu);
                rv = ssb_batch_run(b);
                if (rv == 0 && cudaStreamSynchronize(b->st) != cudaSuccess) {
                    set_error("pipeline: kernels failed: %s", cudaGetErrorString(cudaGetLastError()));
                    rv = -1;
                }
            }
        }
        if (!skip && rv == 0)
            rv = ssb_batch_download(b, &job->out);
        tr[4] = now_ms();
        if (!skip && rv == 0) {
            float ms[8];
            if (ssb_batch_kernel_ms(b, ms) == 0) {
                tr[5] = ms[0];
                tr[6] = ms[4];
            }
        }
        {
            std::lock_guard<std::mutex> lk(p->mu);
            PipeTicket &tk = p->tickets[job->ticket];
            if (rv != 0 && !tk.failed) {
                tk.failed = true;
                tk.err = ssb::last_error();
            }
            if (!skip && rv == 0)
                tk.launches += b->n_launches;
            for (int k = 0; k < 8; ++k)
                tk.trace[(size_t)job->chunk * 8 + k] = tr[k];
            ++tk.n_done;
        }
        p->cv_done.notify_all();
    }
}

extern "C" void ssb_pipeline_free(ssb_pipeline_t *p)
{
    if (!p)
        return;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->stop = true;
    }
    p->cv_job.notify_all();
    for (auto &t : p->workers)
        if (t.joinable())
            t.join();
    cudaSetDevice(p->m->device);
    for (ssb_batch_t *b : p->lane)
        ssb_batch_free(b);
    for (cudaStream_t s : p->st)
        if (s)
            cudaStreamDestroy(s);
    delete p;
}

extern "C" ssb_pipeline_t *ssb_pipeline_create(ssb_model_t *m, int32_t n_lanes, int64_t chunk_frames)
{
    if (need_device(m) != 0)
        return nullptr;
    if (n_lanes <= 0) {
        const char *e = getenv("SSB_PIPE_LANES");
        n_lanes = e ? atoi(e) : 4;
    }
    if (chunk_frames <= 0) {
        const char *e = getenv("SSB_PIPE_CHUNK_FRAMES");
        chunk_frames = e ? atoll(e) : 1024000;
    }
    n_lanes = std::max(1, std::min(n_lanes, 8));
    ssb_pipeline_s *p = new ssb_pipeline_s;
    p->m = m;
    p->n_lanes = n_lanes;
    p->chunk_frames = std::max<int64_t>(chunk_frames, 1);
    for (int i = 0; i < n_lanes; ++i) {
        cudaStream_t s = nullptr;
        if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) {
            set_error("cudaStreamCreate failed");
            ssb_pipeline_free(p);
            return nullptr;
        }
        p->st.push_back(s);
        ssb_batch_t *b = ssb_batch_create(m, s);
        if (!b) {
            ssb_pipeline_free(p);
            return nullptr;
        }
        p->lane.push_back(b);
    }
    for (int i = 0; i < n_lanes; ++i)
        p->workers.emplace_back(pipe_worker, p, i);
    return p;
}

extern "C" int ssb_pipeline_set_overlap(ssb_pipeline_t *p, int32_t overlap_kernels)
{
    if (!p)
        return -1;
    std::lock_guard<std::mutex> lk(p->mu);
    p->overlap_kernels = overlap_kernels != 0;
    return 0;
}

extern "C" int ssb_pipeline_n_launches(const ssb_pipeline_t *p) { return p ? p->n_launches : -1; }
extern "C" int ssb_pipeline_n_chunks(const ssb_pipeline_t *p) { return p ? p->n_chunks : -1; }
extern "C" int ssb_pipeline_trace(const ssb_pipeline_t *p, double *out, int32_t max_chunks)
{
    if (!p || !out)
        return -1;
    const int n = std::min<int>(max_chunks, p->n_chunks);
    for (int i = 0; i < n * 8; ++i)
        out[i] = p->trace[i];
    return n;
}

extern "C" int64_t ssb_pipeline_submit(ssb_pipeline_t *p, const ssb_align_in_t *in, ssb_align_out_t *out)
{
    if (!p || !in || !out || in->n_utts < 0 || (in->n_utts > 0 && (!in->frame_off || !in->phone_off))) {
        set_error("ssb_pipeline_submit: bad arguments");
        return -1;
    }
    if (need_device(p->m) != 0)
        return -1;
    const HostModel &h = p->m->h;
    const int U = in->n_utts, E = h.n_emit;
    if (U > 0 && (in->frame_off[0] != 0 || in->phone_off[0] != 0)) {
        set_error("frame_off[0] and phone_off[0] must be 0");
        return -1;
    }
    // chunk boundaries (whole utterances, about chunk_frames frames each; a multiple of 256
    // utterances where that many fit, the row count of one top-N CTA) and where each chunk's
    // per-state-frame debug outputs start
    std::vector<int> cut(1, 0);
    std::vector<int64_t> scr0(1, 0);
    {
        int64_t scr = 0;
        int u0 = 0;
        for (int u = 0; u < U; ++u) {
            const int64_t T = in->frame_off[u + 1] - in->frame_off[u];
            const int64_t np = in->phone_off[u + 1] - in->phone_off[u];
            if (T < 0 || np < 0) {
                set_error("utterance %d: offsets must be non-decreasing", u);
                return -1;
            }
            scr += T * np * E;
            const int64_t fr = in->frame_off[u + 1] - in->frame_off[u0];
            const int nu = u + 1 - u0;
            if (u + 1 == U || (fr >= p->chunk_frames && (nu < 256 || nu % 256 == 0))
                || fr >= 2 * p->chunk_frames) {
                cut.push_back(u + 1);
                scr0.push_back(scr);
                u0 = u + 1;
            }
        }
    }
    const int n_chunks = (int)cut.size() - 1;
    const int nw = (h.n_sen + 31) / 32;
    const size_t CSb = h.kind == SSB_SCORER_CONT ? 0 : (size_t)h.n_mgau * h.n_feat;
    std::vector<std::unique_ptr<PipeJob>> jobs;
    for (int c = 0; c < n_chunks; ++c) {
        std::unique_ptr<PipeJob> j(new PipeJob);
        const int u0 = cut[c], u1 = cut[c + 1];
        const int64_t f0 = in->frame_off[u0], p0 = in->phone_off[u0];
        j->chunk = c;
        j->u0 = u0;
        j->fo.resize(u1 - u0 + 1);
        j->po.resize(u1 - u0 + 1);
        for (int u = u0; u <= u1; ++u) {
            j->fo[u - u0] = in->frame_off[u] - f0;
            j->po[u - u0] = in->phone_off[u] - p0;
        }
        j->in = *in;
        j->in.n_utts = u1 - u0;
        j->in.feat = in->feat ? in->feat + f0 * h.blk : nullptr;
        j->in.frame_off = j->fo.data();
        j->in.phone_off = j->po.data();
        j->in.ssid = in->ssid ? in->ssid + p0 : nullptr;
        j->in.tmat = in->tmat ? in->tmat + p0 : nullptr;
        j->in.sf = in->sf ? in->sf + p0 : nullptr;
        j->in.ef = in->ef ? in->ef + p0 : nullptr;
        j->in.init_active = in->init_active ? in->init_active + (size_t)u0 * nw : nullptr;
        j->in.init_topn = in->init_topn ? in->init_topn + (size_t)u0 * CSb * 4 : nullptr;
        j->out = *out;
        j->out.st_start = out->st_start ? out->st_start + p0 * E : nullptr;
        j->out.st_dur = out->st_dur ? out->st_dur + p0 * E : nullptr;
        j->out.st_score = out->st_score ? out->st_score + p0 * E : nullptr;
        j->out.utt_rv = out->utt_rv ? out->utt_rv + u0 : nullptr;
        j->out.utt_best = out->utt_best ? out->utt_best + u0 : nullptr;
        j->out.utt_renorm = out->utt_renorm ? out->utt_renorm + u0 : nullptr;
        j->out.chain_scr = out->chain_scr ? out->chain_scr + scr0[c] : nullptr;
        j->out.tokens = out->tokens ? out->tokens + 2 * scr0[c] : nullptr;
        jobs.push_back(std::move(j));
    }
    int64_t ticket;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        ticket = p->next_ticket++;
        PipeTicket &tk = p->tickets[ticket];
        tk.n_chunks = n_chunks;
        tk.trace.assign((size_t)n_chunks * 8, 0.0);
        tk.t0 = std::chrono::steady_clock::now();
        for (auto &j : jobs) {
            j->ticket = ticket;
            j->seq = p->next_seq++;
            p->jobs.push_back(std::move(j));
        }
    }
    p->cv_job.notify_all();
    return ticket;
}

extern "C" int ssb_pipeline_collect(ssb_pipeline_t *p, int64_t ticket)
{
    if (!p) {
        set_error("NULL pipeline");
        return -1;
    }
    std::unique_lock<std::mutex> lk(p->mu);
    auto it = p->tickets.find(ticket);
    if (it == p->tickets.end()) {
        set_error("ssb_pipeline_collect: unknown ticket %lld", (long long)ticket);
        return -1;
    }
    p->cv_done.wait(lk, [&] { return it->second.n_done >= it->second.n_chunks; });
    PipeTicket tk = std::move(it->second);
    p->tickets.erase(it);
    p->n_launches = tk.launches;
    p->n_chunks = tk.n_chunks;
    p->trace = std::move(tk.trace);
    lk.unlock();
    if (tk.failed) {
        set_error("%s", tk.err.c_str());
        return -1;
    }
    return 0;
}

extern "C" int ssb_pipeline_align(ssb_pipeline_t *p, const ssb_align_in_t *in, ssb_align_out_t *out)
{
    const int64_t t = ssb_pipeline_submit(p, in, out);
    if (t < 0)
        return -1;
    return ssb_pipeline_collect(p, t);
}

extern "C" int ssb_align_batch(ssb_model_t *m, const ssb_align_in_t *in, ssb_align_out_t *out)
{
    // a batch of several chunks goes through the pipeline (copies overlap the kernels)
    if (in && out && in->n_utts > 1 && in->frame_off) {
        const char *e = getenv("SSB_PIPE_CHUNK_FRAMES");
        const int64_t chunk = e ? atoll(e) : 1024000;
        if (in->frame_off[in->n_utts] >= 2 * chunk) {
            ssb_pipeline_t *p = ssb_pipeline_create(m, 0, 0);
            if (!p)
                return -1;
            const int rv = ssb_pipeline_align(p, in, out);
            std::string keep = rv ? ssb::last_error() : "";
            ssb_pipeline_free(p);
            if (rv)
                set_error("%s", keep.c_str());
            return rv;
        }
    }
    if (need_device(m) != 0)
        return -1;
    ssb_batch_t *b = ssb_batch_create(m, nullptr);
    if (!b)
        return -1;
    if (out && (out->tokens || out->chain_scr))
        ssb_batch_debug_tokens(b, (out->tokens ? 1 : 0) | (out->chain_scr ? 2 : 0));
    int rv = ssb_batch_upload(b, in);
    if (rv == 0)
        rv = ssb_batch_run(b);
    if (rv == 0)
        rv = ssb_batch_download(b, out);
    ssb_batch_free(b);
    return rv;
}

// ------------------------------------------------------------------ several GPUs of one box
// The path shards by utterance and has no exchange step: one host thread per GPU,
// each with its device's model, a contiguous range of utterances balanced by frames (slices of
// the caller's arrays: nothing is copied or re-packed), its own pipeline; results land in the
// caller's output arrays directly.  No collective.
extern "C" int ssb_align_batch_multi(ssb_model_t *const *models, int32_t n_models,
                                     const ssb_align_in_t *in, ssb_align_out_t *out)
{
    if (!models || n_models <= 0 || !in || !out || in->n_utts < 0 || (in->n_utts > 0 && (!in->frame_off || !in->phone_off))) {
        set_error("ssb_align_batch_multi: bad arguments");
        return -1;
    }
    for (int d = 0; d < n_models; ++d)
        if (!models[d] || models[d]->h.n_emit != models[0]->h.n_emit || models[d]->h.n_sen != models[0]->h.n_sen) {
            set_error("ssb_align_batch_multi: model %d missing or different from model 0", d);
            return -1;
        }
    if (n_models == 1 || in->n_utts < 2)
        return ssb_align_batch(models[0], in, out);
    if (out->chain_scr || out->tokens) {
        set_error("ssb_align_batch_multi: debug outputs (chain_scr, tokens) are single-device only");
        return -1;
    }
    const int U = in->n_utts, E = models[0]->h.n_emit, nw = (models[0]->h.n_sen + 31) / 32;
    const int CS = models[0]->h.n_mgau * models[0]->h.n_feat, blk = models[0]->h.blk;
    // contiguous ranges with about the same number of frames
    std::vector<int> cut(n_models + 1, U);
    cut[0] = 0;
    const int64_t G = in->frame_off[U];
    for (int d = 1, u = 0; d < n_models; ++d) {
        while (u < U && in->frame_off[u] < G * d / n_models)
            ++u;
        cut[d] = u;
    }
    std::vector<int> rv(n_models, 0);
    std::vector<std::string> err(n_models);
    std::vector<std::thread> th;
    g_plan_divisor.store(n_models);
    for (int d = 0; d < n_models; ++d) {
        th.emplace_back([&, d] {
            const int a = cut[d], b = cut[d + 1];
            if (b <= a)
                return;
            const int n = b - a;
            std::vector<int64_t> fo(n + 1), po(n + 1);
            for (int i = 0; i <= n; ++i) {
                fo[i] = in->frame_off[a + i] - in->frame_off[a];
                po[i] = in->phone_off[a + i] - in->phone_off[a];
            }
            const int64_t f0 = in->frame_off[a], p0 = in->phone_off[a];
            ssb_align_in_t si = *in;
            si.n_utts = n;
            si.feat = in->feat + f0 * blk;
            si.frame_off = fo.data();
            si.phone_off = po.data();
            si.ssid = in->ssid + p0;
            si.tmat = in->tmat + p0;
            si.sf = in->sf + p0;
            si.ef = in->ef + p0;
            si.init_active = in->init_active ? in->init_active + (size_t)a * nw : nullptr;
            si.init_topn = in->init_topn ? in->init_topn + (size_t)a * CS * 4 : nullptr;
            ssb_align_out_t so = *out;
            so.st_start = out->st_start ? out->st_start + p0 * E : nullptr;
            so.st_dur = out->st_dur ? out->st_dur + p0 * E : nullptr;
            so.st_score = out->st_score ? out->st_score + p0 * E : nullptr;
            so.utt_rv = out->utt_rv ? out->utt_rv + a : nullptr;
            so.utt_best = out->utt_best ? out->utt_best + a : nullptr;
            so.utt_renorm = out->utt_renorm ? out->utt_renorm + a : nullptr;
            cudaSetDevice(models[d]->device);
            rv[d] = ssb_align_batch(models[d], &si, &so);
            if (rv[d] != 0)
                err[d] = ssb::last_error();  // (the error text is per thread)
        });
    }
    for (auto &t : th)
        t.join();
    g_plan_divisor.store(1);
    for (int d = 0; d < n_models; ++d)
        if (rv[d] != 0) {
            set_error("device %d: %s", models[d]->device, err[d].c_str());
            return -1;
        }
    return 0;
}

// ------------------------------------------------------------------ dense scoring
static int score_prepare(ssb_batch_t *b, const float *feat, const int64_t *frame_off, int32_t U)
{
    std::vector<int64_t> poff(U + 1, 0);
    ssb_align_in_t in;
    std::memset(&in, 0, sizeof(in));
    in.n_utts = U;
    in.feat = feat;
    in.frame_off = frame_off;
    in.phone_off = poff.data();
    int32_t dummy = 0;
    in.ssid = in.tmat = in.sf = in.ef = &dummy;
    in.compallsen = 1;
    return ssb_batch_upload(b, &in);
}

extern "C" int64_t ssb_score_batch(ssb_model_t *m, const float *feat, const int64_t *frame_off,
                                   int32_t n_utts, int16_t *senscr)
{
    ssb_batch_t *b = ssb_batch_create(m, nullptr);
    if (!b)
        return -1;
    int64_t rv = -1;
    do {
        if (score_prepare(b, feat, frame_off, n_utts) != 0)
            break;
        const DevModel &d = m->d;
        const int64_t G = b->n_frames;
        if (G == 0) {
            rv = 0;
            break;
        }
        if (batch_topn(b, nullptr, 0, true) != 0)
            break;
        bool ok = true;
        for (int64_t g0 = 0; g0 < G && ok; g0 += kSlabFrames) {
            const int64_t n = std::min(kSlabFrames, G - g0);
            ok = b->dense.ensure((size_t)n * d.n_sen * 2) == 0
                 && dense_scores(b, g0, n, b->dense.as<int16_t>()) == 0;
            if (ok && senscr
                && cudaMemcpyAsync(senscr + g0 * d.n_sen, b->dense.p, (size_t)n * d.n_sen * 2,
                                   cudaMemcpyDeviceToHost, b->st) != cudaSuccess) {
                set_error("senone score download failed: %s", cudaGetErrorString(cudaGetLastError()));
                ok = false;
            }
        }
        if (!ok)
            break;
        cudaError_t e = cudaStreamSynchronize(b->st);
        if (e != cudaSuccess) {
            set_error("ssb_score_batch: %s", cudaGetErrorString(e));
            break;
        }
        rv = G;
    } while (0);
    ssb_batch_free(b);
    return rv;
}

static int64_t topn_impl(ssb_model_t *m, const float *feat, const int64_t *frame_off,
                         int32_t n_utts, uint8_t *cw, int32_t *score, float *approx, float *eps,
                         int64_t *counters)
{
    ssb_batch_t *b = ssb_batch_create(m, nullptr);
    if (!b)
        return -1;
    const bool probe = approx || eps || counters;
    DBuf d_approx, d_eps, d_cnt;
    int64_t rv = -1;
    do {
        if (score_prepare(b, feat, frame_off, n_utts) != 0)
            break;
        const DevModel &d = m->d;
        const int64_t G = b->n_frames;
        const int CS = d.n_mgau * d.n_feat, N = d.topn, ND = d.n_density;
        if (G == 0) {
            rv = 0;
            break;
        }
        if (!probe) {
            // exact raw scores (the frame-tiled kernel's production output is scores >> 10 only)
            if (batch_topn(b, nullptr, 0, true, 1) != 0)
                break;
        } else if (b->ft) {
            if (d_approx.ensure((size_t)G * CS * ND * 4) || d_eps.ensure((size_t)G * CS * 8)
                || d_cnt.ensure(32))
                break;
            cudaMemsetAsync(d_cnt.p, 0, 32, b->st);
            cudaMemsetAsync(d_eps.p, 0, (size_t)G * CS * 8, b->st);
            const char *ex = getenv("SSB_FT_EXACT");
            if (batch_topn(b, nullptr, 0, true, !(ex && *ex == '0'),
                           TcDebug{d_approx.as<float>(), d_eps.as<float>(),
                                   d_cnt.as<unsigned long long>()}) != 0)
                break;
        } else {
            if (!tc_supported(d)) {
                set_error("ssb_tc_probe: model shape not supported by the tensor-core scorer");
                break;
            }
            if (d_approx.ensure((size_t)G * CS * ND * 4) || d_eps.ensure((size_t)G * CS * 8)
                || d_cnt.ensure(32))
                break;
            cudaMemsetAsync(d_cnt.p, 0, 32, b->st);
            if (launch_gmm_topn_tc(d, b->plan, b->feat.as<float>(), G, b->tn_s.as<int4>(),
                                   b->tn_c.as<uchar4>(), b->featp.as<float>(), d_approx.as<float>(),
                                   d_eps.as<float>(),
                                   d_cnt.as<unsigned long long>(), b->st) != 0)
                break;
        }
        std::vector<int4> hs((size_t)G * CS);
        std::vector<uchar4> hc((size_t)G * CS);
        if (cudaMemcpyAsync(hs.data(), b->tn_s.p, hs.size() * sizeof(int4), cudaMemcpyDeviceToHost, b->st) != cudaSuccess
            || cudaMemcpyAsync(hc.data(), b->tn_c.p, hc.size() * sizeof(uchar4), cudaMemcpyDeviceToHost, b->st) != cudaSuccess
            || cudaStreamSynchronize(b->st) != cudaSuccess) {
            set_error("ssb_topn_batch: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        // device layout [cs][frame] -> [frame][cs][topn]
        for (int64_t g = 0; g < G; ++g)
            for (int cs = 0; cs < CS; ++cs) {
                const int4 s = hs[(size_t)cs * G + g];
                const uchar4 c = hc[(size_t)cs * G + g];
                const int32_t sv[4] = {s.x, s.y, s.z, s.w};
                const uint8_t cv[4] = {c.x, c.y, c.z, c.w};
                for (int k = 0; k < N; ++k) {
                    if (score)
                        score[((size_t)g * CS + cs) * N + k] = sv[k];
                    if (cw)
                        cw[((size_t)g * CS + cs) * N + k] = cv[k];
                }
            }
        if (probe) {
            std::vector<float> ha((size_t)G * CS * ND), he((size_t)G * CS * 2);
            unsigned long long hcnt[4] = {0, 0, 0, 0};
            if (cudaMemcpy(ha.data(), d_approx.p, ha.size() * 4, cudaMemcpyDeviceToHost) != cudaSuccess
                || cudaMemcpy(he.data(), d_eps.p, he.size() * 4, cudaMemcpyDeviceToHost) != cudaSuccess
                || cudaMemcpy(hcnt, d_cnt.p, 32, cudaMemcpyDeviceToHost) != cudaSuccess) {
                set_error("ssb_tc_probe: %s", cudaGetErrorString(cudaGetLastError()));
                break;
            }
            for (int64_t g = 0; g < G; ++g)
                for (int cs = 0; cs < CS; ++cs) {
                    if (approx)
                        std::memcpy(approx + ((size_t)g * CS + cs) * ND,
                                    ha.data() + ((size_t)cs * G + g) * ND, (size_t)ND * 4);
                    if (eps)
                        for (int q = 0; q < 2; ++q)
                            eps[((size_t)g * CS + cs) * 2 + q] = he[((size_t)cs * G + g) * 2 + q];
                }
            if (counters) {
                counters[0] = (int64_t)hcnt[0];
                counters[1] = (int64_t)hcnt[1];
                counters[2] = (int64_t)hcnt[2];
            }
        }
        rv = G;
    } while (0);
    d_approx.release();
    d_eps.release();
    d_cnt.release();
    ssb_batch_free(b);
    return rv;
}

extern "C" int64_t ssb_topn_batch(ssb_model_t *m, const float *feat, const int64_t *frame_off,
                                  int32_t n_utts, uint8_t *cw, int32_t *score)
{
    return topn_impl(m, feat, frame_off, n_utts, cw, score, nullptr, nullptr, nullptr);
}

extern "C" int ssb_tc_hot_mask(const ssb_model_t *m, uint32_t *out)
{
    if (!m || !out || m->tc_hot.empty()) {
        set_error("ssb_tc_hot_mask: model has no tensor-core operands (device = -1 or unsupported shape)");
        return -1;
    }
    std::memcpy(out, m->tc_hot.data(), m->tc_hot.size() * 4);
    return 0;
}

extern "C" int64_t ssb_tc_probe(ssb_model_t *m, const float *feat, const int64_t *frame_off,
                                int32_t n_utts, uint8_t *cw, int32_t *score, float *approx,
                                float *eps, int64_t *counters)
{
    static float dummy_eps;
    if (!approx && !eps && !counters)
        eps = &dummy_eps, (void)0;
    if (eps == &dummy_eps) {
        set_error("ssb_tc_probe: nothing requested");
        return -1;
    }
    return topn_impl(m, feat, frame_off, n_utts, cw, score, approx, eps, counters);
}

extern "C" int ssb_hmm_vit_eval(ssb_model_t *m, int32_t n_emit, int32_t tmatid,
                                const uint16_t *senid, const int16_t *senscr, int32_t *st12,
                                int32_t *best)
{
    if (need_device(m) != 0)
        return -1;
    if (!senid || !senscr || !st12 || tmatid < 0 || tmatid >= m->h.n_tmat
        || n_emit != m->h.n_emit) {
        set_error("ssb_hmm_vit_eval: bad arguments (n_emit must equal the model's %d)", m->h.n_emit);
        return -1;
    }
    for (int j = 0; j < n_emit; ++j)
        if (senid[j] >= m->h.n_sen) {
            set_error("ssb_hmm_vit_eval: senone id out of range");
            return -1;
        }
    DBuf a, s, t, o;
    int rv = -1;
    do {
        if (a.ensure(16) || s.ensure((size_t)m->h.n_sen * 2) || t.ensure(48) || o.ensure(16))
            break;
        if (cudaMemcpy(a.p, senid, n_emit * 2, cudaMemcpyHostToDevice) != cudaSuccess
            || cudaMemcpy(s.p, senscr, (size_t)m->h.n_sen * 2, cudaMemcpyHostToDevice) != cudaSuccess
            || cudaMemcpy(t.p, st12, 48, cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("ssb_hmm_vit_eval: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        if (launch_hmm_eval(m->d, n_emit, tmatid, a.as<uint16_t>(), s.as<int16_t>(),
                            t.as<int32_t>(), o.as<int32_t>(), nullptr) != 0)
            break;
        int32_t bb = 0;
        if (cudaMemcpy(st12, t.p, 48, cudaMemcpyDeviceToHost) != cudaSuccess
            || cudaMemcpy(&bb, o.p, 4, cudaMemcpyDeviceToHost) != cudaSuccess) {
            set_error("ssb_hmm_vit_eval: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        if (best)
            *best = bb;
        rv = 0;
    } while (0);
    a.release();
    s.release();
    t.release();
    o.release();
    return rv;
}

extern "C" int ssb_hmm_vit_eval_tp(ssb_model_t *m, int32_t n_emit, int32_t n_cases, const uint8_t *tp,
                                   const int16_t *senscr, int32_t *st12, int32_t *best)
{
    if (need_device(m) != 0)
        return -1;
    if ((n_emit != 3 && n_emit != 5) || n_cases < 0 || (n_cases > 0 && (!tp || !senscr || !st12 || !best))) {
        set_error("ssb_hmm_vit_eval_tp: bad arguments (3 or 5 emitting states)");
        return -1;
    }
    if (n_cases == 0)
        return 0;
    const size_t nt = (size_t)n_cases * n_emit * (n_emit + 1), ns = (size_t)n_cases * n_emit * 2,
                 nst = (size_t)n_cases * 48, nb = (size_t)n_cases * 4;
    DBuf a, s, t, o;
    int rv = -1;
    do {
        if (a.ensure(nt) || s.ensure(ns) || t.ensure(nst) || o.ensure(nb))
            break;
        if (cudaMemcpy(a.p, tp, nt, cudaMemcpyHostToDevice) != cudaSuccess
            || cudaMemcpy(s.p, senscr, ns, cudaMemcpyHostToDevice) != cudaSuccess
            || cudaMemcpy(t.p, st12, nst, cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("ssb_hmm_vit_eval_tp: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        if (launch_hmm_eval_tp(n_emit, n_cases, a.as<uint8_t>(), s.as<int16_t>(), t.as<int32_t>(),
                               o.as<int32_t>(), nullptr) != 0)
            break;
        if (cudaMemcpy(st12, t.p, nst, cudaMemcpyDeviceToHost) != cudaSuccess
            || cudaMemcpy(best, o.p, nb, cudaMemcpyDeviceToHost) != cudaSuccess) {
            set_error("ssb_hmm_vit_eval_tp: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        rv = 0;
    } while (0);
    a.release();
    s.release();
    t.release();
    o.release();
    return rv;
}

// ------------------------------------------------------------------ FSG search (K4)
// frames per dense-score slab of the FSG path: whole utterances, <= ~12 GB of int16 scores
static const int64_t kFsgSlabFrames = 1200000;

extern "C" int ssb_model_fsg_active_ok(const ssb_model_t *m)
{
    const char *k1 = getenv("SSB_K1");  // the earlier A/B kernels do not flag tie steps
    return m && m->device >= 0 && tc_supported(m->d) && (!(k1 && *k1) || strcmp(k1, "ft") == 0 || strcmp(k1, "tc2") == 0);
}

extern "C" int ssb_fsg_batch(ssb_model_t *m, const ssb_fsg_in_t *in, ssb_fsg_out_t *out)
{
    if (!in || !out || in->n_utts < 0 || in->n_graphs < 0 || in->hist_cap < 2 || in->max_seg < 1
        || (in->n_utts > 0 && (!in->frame_off || !in->graphs || !in->utt_graph || in->n_graphs < 1))) {
        set_error("ssb_fsg_batch: bad arguments");
        return -1;
    }
    if (need_device(m) != 0)
        return -1;
    const HostModel &h = m->h;
    const int U = in->n_utts;
    if (U == 0)
        return 0;
    // ---- validate + concatenate the graphs
    std::vector<DevFsg> hdr(in->n_graphs);
    std::vector<int32_t> link4, arc_off, root, pnode8;
    std::vector<uint8_t> link_flag;
    std::vector<uint32_t> ctxt;
    int tent_cap = 16;
    for (int gi = 0; gi < in->n_graphs; ++gi) {
        const ssb_fsg_graph_t &g = in->graphs[gi];
        if (g.n_state < 1 || g.n_link < 0 || g.n_pnode < 0 || g.n_ciphone < 1 || g.n_ciphone > 128
            || g.start < 0 || g.start >= g.n_state || g.final < 0 || g.final >= g.n_state
            || g.sil < 0 || g.sil >= g.n_ciphone || !g.arc_off || !g.root
            || (g.n_link && (!g.link4 || !g.link_flag)) || (g.n_pnode && (!g.pnode8 || !g.ctxt))) {
            set_error("ssb_fsg_batch: graph %d is malformed", gi);
            return -1;
        }
        if (g.arc_off[0] != 0 || g.arc_off[g.n_state] != g.n_link) {
            set_error("ssb_fsg_batch: graph %d: arc_off must span the %d links", gi, g.n_link);
            return -1;
        }
        for (int i = 0; i < g.n_link; ++i) {
            const int32_t *l = g.link4 + i * 4;
            if (l[0] < 0 || l[0] >= g.n_state || l[1] < 0 || l[1] >= g.n_state) {
                set_error("ssb_fsg_batch: graph %d link %d: state out of range", gi, i);
                return -1;
            }
        }
        for (int i = 0; i < g.n_pnode; ++i) {
            const int32_t *p = g.pnode8 + i * 8;
            const bool leaf = p[4] != 0;
            if (p[0] < 0 || p[0] >= h.n_sseq || p[1] < 0 || p[1] >= h.n_tmat || p[3] < 0
                || p[3] >= g.n_ciphone || p[6] < -1 || p[6] >= g.n_pnode
                || (leaf ? (p[5] < 0 || p[5] >= g.n_link) : (p[5] < -1 || p[5] >= g.n_pnode))) {
                set_error("ssb_fsg_batch: graph %d pnode %d is malformed", gi, i);
                return -1;
            }
        }
        for (int s = 0; s < g.n_state; ++s)
            if (g.root[s] < -1 || g.root[s] >= g.n_pnode || g.arc_off[s] > g.arc_off[s + 1]) {
                set_error("ssb_fsg_batch: graph %d state %d is malformed", gi, s);
                return -1;
            }
        DevFsg &d = hdr[gi];
        d.n_state = g.n_state;
        d.start = g.start;
        d.final = g.final;
        d.n_link = g.n_link;
        d.n_pnode = g.n_pnode;
        d.n_ciphone = g.n_ciphone;
        d.sil = g.sil;
        d.beam = g.beam;
        d.pbeam = g.pbeam;
        d.wbeam = g.wbeam;
        d.maxhmmpf = g.maxhmmpf;
        d.link_off = (int32_t)(link4.size() / 4);
        d.arcoff_off = (int32_t)arc_off.size();
        d.root_off = (int32_t)root.size();
        d.pnode_off = (int32_t)(pnode8.size() / 8);
        link4.insert(link4.end(), g.link4, g.link4 + (size_t)g.n_link * 4);
        link_flag.insert(link_flag.end(), g.link_flag, g.link_flag + g.n_link);
        arc_off.insert(arc_off.end(), g.arc_off, g.arc_off + g.n_state + 1);
        root.insert(root.end(), g.root, g.root + g.n_state);
        pnode8.insert(pnode8.end(), g.pnode8, g.pnode8 + (size_t)g.n_pnode * 8);
        ctxt.insert(ctxt.end(), g.ctxt, g.ctxt + (size_t)g.n_pnode * 4);
        tent_cap = std::max(tent_cap, g.n_pnode + g.n_link + 16);
    }
    std::vector<int64_t> ws_off(U + 1, 0);
    std::vector<int32_t> utt_graph(in->utt_graph, in->utt_graph + U);
    for (int u = 0; u < U; ++u) {
        if (utt_graph[u] < 0 || utt_graph[u] >= in->n_graphs) {
            set_error("ssb_fsg_batch: utterance %d names graph %d of %d", u, utt_graph[u], in->n_graphs);
            return -1;
        }
        const DevFsg &d = hdr[utt_graph[u]];
        ws_off[u + 1] = ws_off[u] + (int64_t)d.n_pnode * (FSG_PH + 2) + (int64_t)d.n_state * d.n_ciphone
                        + tent_cap + (int64_t)tent_cap * FSG_TE;
    }
    // mode "compallsen = no": scores are computed inside the search kernel
    const bool active = in->active_lists != 0;
    std::vector<int64_t> aws_off(U + 1, 0);
    if (active) {
        if (!ssb_model_fsg_active_ok(m)) {
            set_error("ssb_fsg_batch: active lists need the tensor-core top-N kernel (a PTM model "
                      "with 128 densities, SSB_K1 unset); use active_lists = 0 (compallsen)");
            return -1;
        }
        for (int u = 0; u < U; ++u)
            aws_off[u + 1] = aws_off[u] + (int64_t)fsg_active_ws_ints(m->d, hdr[utt_graph[u]].n_pnode);
    }
    if (need_device(m) != 0)
        return -1;
    ssb_batch_t *b = ssb_batch_create(m, nullptr);
    if (!b)
        return -1;
    DBuf d_hdr, d_link4, d_flag, d_arc, d_root, d_pnode, d_ctxt, d_ug, d_wsoff, d_ws, d_hist, d_nhist,
        d_neval, d_frames, d_rv, d_exit, d_score, d_segs, d_nseg, d_awsoff, d_aws, d_tie, d_fact, d_nsen,
        d_ftopn;
    int rv = -1;
    do {
        if (score_prepare(b, in->feat, in->frame_off, U) != 0)
            break;
        cudaStream_t st = b->st;
        const DevModel &d = m->d;
        const int64_t G = b->n_frames;
        const size_t hist_ints = (size_t)U * in->hist_cap * 9;
        if (upload(d_hdr, hdr, st) || upload(d_link4, link4, st) || upload(d_flag, link_flag, st)
            || upload(d_arc, arc_off, st) || upload(d_root, root, st) || upload(d_pnode, pnode8, st)
            || upload(d_ctxt, ctxt, st) || upload(d_ug, utt_graph, st) || upload(d_wsoff, ws_off, st)
            || d_ws.ensure(std::max<size_t>((size_t)ws_off[U] * 4, 16)) || d_hist.ensure(hist_ints * 4)
            || d_nhist.ensure((size_t)U * 4) || d_neval.ensure((size_t)U * 8) || d_frames.ensure((size_t)U * 4)
            || d_rv.ensure((size_t)U * 4) || d_exit.ensure((size_t)U * 4) || d_score.ensure((size_t)U * 4)
            || d_segs.ensure((size_t)U * in->max_seg * 5 * 4) || d_nseg.ensure((size_t)U * 4))
            break;
        DevFsgSet gs;
        gs.graph = d_hdr.as<DevFsg>();
        gs.link4 = d_link4.as<int32_t>();
        gs.link_flag = d_flag.as<uint8_t>();
        gs.arc_off = d_arc.as<int32_t>();
        gs.root = d_root.as<int32_t>();
        gs.pnode8 = d_pnode.as<int32_t>();
        gs.ctxt = d_ctxt.as<uint32_t>();
        const int nw_sen = (h.n_sen + 31) / 32;
        const int64_t tie_w = (G + 31) / 32 + 1;
        DevPlan plan = b->plan;
        if (active) {
            const size_t tie_bytes = (size_t)h.n_mgau * h.n_feat * tie_w * 4;
            if (upload(d_awsoff, aws_off, st) || d_aws.ensure(std::max<size_t>((size_t)aws_off[U] * 4, 16))
                || d_tie.ensure(tie_bytes) || d_fact.ensure((size_t)U * nw_sen * 4)
                || d_nsen.ensure((size_t)U * 8)
                || cudaMemsetAsync(d_tie.p, 0, tie_bytes, st) != cudaSuccess)
                break;
            plan.tie_bits = d_tie.as<uint32_t>();
            plan.tie_w = tie_w;
        }
        const size_t ftopn_bytes = (size_t)U * h.n_mgau * h.n_feat * 4;
        if (out->final_topn && h.kind != SSB_SCORER_CONT && d_ftopn.ensure(std::max<size_t>(ftopn_bytes, 16)))
            break;
        launch_count(true);
        cudaEventRecord(b->ev[0], st);
        // (the default-mode search replays tie steps from its own carried lists: no fix-up pass)
        if (G > 0 && batch_topn(b, plan.tie_bits, plan.tie_w, !active) != 0)
            break;
        cudaEventRecord(b->ev[1], st);
        // dense senone scores slab by slab (whole utterances), searched as soon as they exist
        bool ok = true;
        float ms_mix = 0.f, ms_search = 0.f;
        int u0 = 0;
        if (active) {
            cudaEventRecord(b->ev[3], st);
            ok = launch_fsg_search_active(d, gs, b->d_frame_off.as<int64_t>(), d_ug.as<int32_t>(),
                                          d_wsoff.as<int64_t>(), d_ws.as<int32_t>(), b->feat.as<float>(),
                                          b->tn_s.as<int4>(), b->tn_c.as<uchar4>(), d_tie.as<uint32_t>(),
                                          G, tie_w, d_awsoff.as<int64_t>(), d_aws.as<int32_t>(),
                                          d_fact.as<uint32_t>(), d_nsen.as<int64_t>(),
                                          out->final_topn ? d_ftopn.as<uchar4>() : nullptr, U,
                                          d_hist.as<int32_t>(), in->hist_cap, tent_cap,
                                          d_nhist.as<int32_t>(), d_neval.as<int64_t>(),
                                          d_frames.as<int32_t>(), d_rv.as<int32_t>(), st) == 0;
            cudaEventRecord(b->ev[4], st);
            if (ok && cudaEventSynchronize(b->ev[4]) == cudaSuccess)
                cudaEventElapsedTime(&ms_search, b->ev[3], b->ev[4]);
            u0 = U;
        }
        while (u0 < U && ok) {
            int u1 = u0 + 1;
            while (u1 < U && b->frame_off[u1 + 1] - b->frame_off[u0] <= kFsgSlabFrames)
                ++u1;
            const int64_t g0 = b->frame_off[u0], n = b->frame_off[u1] - g0;
            cudaEventRecord(b->ev[2], st);
            if (n > 0)
                ok = b->dense.ensure((size_t)n * d.n_sen * 2) == 0
                     && dense_scores(b, g0, n, b->dense.as<int16_t>()) == 0;
            cudaEventRecord(b->ev[3], st);
            ok = ok
                 && launch_fsg_search(d, gs, b->d_frame_off.as<int64_t>(), d_ug.as<int32_t>(),
                                      d_wsoff.as<int64_t>(), d_ws.as<int32_t>(), b->dense.as<int16_t>(), g0,
                                      u0, u1 - u0, d_hist.as<int32_t>(), in->hist_cap, tent_cap,
                                      d_nhist.as<int32_t>(), d_neval.as<int64_t>(), d_frames.as<int32_t>(),
                                      d_rv.as<int32_t>(), st) == 0;
            cudaEventRecord(b->ev[4], st);
            if (ok && cudaEventSynchronize(b->ev[4]) == cudaSuccess) {
                float a = 0.f, c = 0.f;
                cudaEventElapsedTime(&a, b->ev[2], b->ev[3]);
                cudaEventElapsedTime(&c, b->ev[3], b->ev[4]);
                ms_mix += a;
                ms_search += c;
            }
            u0 = u1;
        }
        if (!ok)
            break;
        if (!active && out->final_topn && h.kind != SSB_SCORER_CONT && G > 0
            && launch_fsg_final_topn_dense(d, b->d_frame_off.as<int64_t>(), U, b->tn_c.as<uchar4>(), G,
                                           d_ftopn.as<uchar4>(), st) != 0)
            break;
        cudaEventRecord(b->ev[2], st);
        if (launch_fsg_backtrace(gs, d_ug.as<int32_t>(), 0, U, d_hist.as<int32_t>(), in->hist_cap,
                                 d_nhist.as<int32_t>(), d_frames.as<int32_t>(), d_exit.as<int32_t>(),
                                 d_score.as<int32_t>(), d_segs.as<int32_t>(), in->max_seg,
                                 d_nseg.as<int32_t>(), in->partial ? 0 : 1, st) != 0)
            break;
        cudaEventRecord(b->ev[3], st);
        struct { void *dst; DBuf *src; size_t bytes; } copies[] = {
            {out->segs, &d_segs, (size_t)U * in->max_seg * 5 * 4}, {out->n_seg, &d_nseg, (size_t)U * 4},
            {out->hyp_score, &d_score, (size_t)U * 4}, {out->exit_bp, &d_exit, (size_t)U * 4},
            {out->utt_rv, &d_rv, (size_t)U * 4}, {out->n_hist, &d_nhist, (size_t)U * 4},
            {out->n_hmm_eval, &d_neval, (size_t)U * 8}, {out->hist9, &d_hist, hist_ints * 4},
            {active ? (void *)out->final_active : nullptr, &d_fact, (size_t)U * nw_sen * 4},
            {active ? (void *)out->n_sen_eval : nullptr, &d_nsen, (size_t)U * 8},
            {(h.kind != SSB_SCORER_CONT && G > 0) ? (void *)out->final_topn : nullptr, &d_ftopn, ftopn_bytes}};
        for (auto &c : copies)
            if (c.dst && cudaMemcpyAsync(c.dst, c.src->p, c.bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess)
                ok = false;
        cudaError_t e = cudaStreamSynchronize(st);
        if (!ok || e != cudaSuccess) {
            set_error("ssb_fsg_batch: %s", cudaGetErrorString(e != cudaSuccess ? e : cudaGetLastError()));
            break;
        }
        if (out->kernel_ms) {
            float k1 = 0.f, bt = 0.f;
            cudaEventElapsedTime(&k1, b->ev[0], b->ev[1]);
            cudaEventElapsedTime(&bt, b->ev[2], b->ev[3]);
            out->kernel_ms[0] = k1;
            out->kernel_ms[1] = ms_mix;
            out->kernel_ms[2] = ms_search;
            out->kernel_ms[3] = bt;
        }
        out->n_launches = launch_count(true);
        rv = 0;
    } while (0);
    cudaStreamSynchronize(b->st);  // (error paths: nothing may still be running on these buffers)
    DBuf *all[] = {&d_hdr, &d_link4, &d_flag, &d_arc, &d_root, &d_pnode, &d_ctxt, &d_ug, &d_wsoff, &d_ws,
                   &d_hist, &d_nhist, &d_neval, &d_frames, &d_rv, &d_exit, &d_score, &d_segs, &d_nseg,
                   &d_awsoff, &d_aws, &d_tie, &d_fact, &d_nsen, &d_ftopn};
    for (DBuf *x : all)
        x->release();
    ssb_batch_free(b);
    return rv;
}
