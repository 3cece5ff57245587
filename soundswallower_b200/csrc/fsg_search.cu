// fsg_search.cu -- K4: FSG token passing over a flattened lextree, one warp per utterance.
//
// Replaces fsg_search_start / step / finish and their helpers (ref: src/fsg_search.c:309-851),
// the history table (ref: src/fsg_history.c:129-232) and the backtrace behind
// fsg_search_hyp / fsg_search_seg_iter (ref: src/fsg_search.c:853-924, 1030-1142).  The graph
// (FSG + per-state phonetic lextrees with their cross-word context sets) is built on the host
// by the caller -- fsg_lextree.c / fsg_model.c are graph preparation, not the per-frame path --
// and arrives as flat arrays (include/ssb200.h: ssb_fsg_graph_t).  Senone scores: either dense
// ("compallsen = yes", computed by K1 + K2' for the whole batch beforehand), or -- the
// reference's default -- computed inside this kernel frame by frame for the senones of the
// active HMMs only (fsg_search_sen_active + acmod_score with an active list, ref:
// src/fsg_search.c:309-328, 686-690; src/acmod.c:822-999; src/ptm_mgau.c:264-403): the beam
// decides the active set, the active set decides the normalisers, so scoring and search cannot
// be separated there.  What CAN be computed beforehand is the Gaussian top-N of every
// (frame, codebook, stream): the reference's list after a scan is "the N best, sorted" whatever
// list it carried in, unless integer scores tie -- K1 flags those steps and they are replayed
// here literally from the list this search carried (see replay_tie).
//
// The search is beam pruned and pointer chasing: a handful of HMMs (~5) are alive per frame
// whatever the grammar size, history entries are inserted into per-(state, left-context)
// lists whose order decides ties.  All of that is kept literally, so it runs sequentially on
// lane 0; the warp's other lanes evaluate the active HMMs in parallel (the only data-parallel
// piece) and the machine is filled by utterances: 4096 utterances = 4096 warps.  Integer
// work: results are bit-exact (history table, scores, segmentation).
#include <cstring>

#include "hmm_step.cuh"
#include "tc_common.cuh"

namespace ssb {

struct FsgUtt {  // per-utterance view, set up once by every lane
    const DevFsg *g;
    const int32_t *link4, *arc_off, *root, *pnode8;
    const uint8_t *link_flag;
    const uint32_t *ctxt;
    int32_t *phmm, *act, *nxt, *heads, *touched, *tent;
    int32_t *hist;
    int cap, tent_cap;
    // lane-0 state
    int n_act, n_nxt, n_hist, n_tent, n_touched, overflow;
    int32_t frame, bestscore, bpidx_start, beam, pbeam, wbeam;
    float beam_factor;
};

constexpr int PH = FSG_PH;  // phmm: score[5] hist[5] out_score out_hist frame bestscore
constexpr int TE = FSG_TE;  // tentative entry: link score pred frame lc rc[4] next

__device__ __forceinline__ void ph_clear(int32_t *h)
{
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        h[i] = WORST_SCORE;
        h[5 + i] = -1;
    }
    h[10] = WORST_SCORE;
    h[11] = -1;
    h[12] = -1;
    h[13] = WORST_SCORE;
}

__device__ __forceinline__ void ph_enter(int32_t *h, int32_t score, int32_t hist, int32_t frame)
{
    h[0] = score;
    h[5] = hist;
    h[12] = frame;
}

__device__ void hist_append(FsgUtt &s, const int32_t *e9)
{
    if (s.n_hist >= s.cap) {
        s.overflow = 1;
        return;
    }
    int32_t *o = s.hist + (size_t)s.n_hist * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k)
        o[k] = e9[k];
    ++s.n_hist;
}

// ref: src/fsg_history.c:129-202
__device__ void entry_add(FsgUtt &s, int link, int32_t frame, int32_t score, int32_t pred, int32_t lc,
                          const uint32_t *rc_in)
{
    uint32_t rc[4] = {rc_in[0], rc_in[1], rc_in[2], rc_in[3]};
    if (frame < 0) {
        int32_t e[9] = {link, score, pred, frame, lc, (int32_t)rc[0], (int32_t)rc[1], (int32_t)rc[2],
                        (int32_t)rc[3]};
        hist_append(s, e);
        return;
    }
    const int hidx = s.link4[link * 4 + 1] * s.g->n_ciphone + lc;
    int prev = -1, gn;
    for (gn = s.heads[hidx]; gn >= 0; gn = s.tent[gn * TE + 9]) {
        const int32_t *e = s.tent + gn * TE;
        if (score > e[1])
            break;
        uint32_t left = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            left |= (rc[k] = ~(uint32_t)e[5 + k] & rc[k]);
        if (left == 0)
            return;
        prev = gn;
    }
    if (s.n_tent >= s.tent_cap) {
        s.overflow = 1;
        return;
    }
    const int k = s.n_tent++;
    int32_t *ne = s.tent + k * TE;
    ne[0] = link;
    ne[1] = score;
    ne[2] = pred;
    ne[3] = frame;
    ne[4] = lc;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        ne[5 + q] = (int32_t)rc[q];
    ne[9] = gn;
    if (prev < 0) {
        if (s.heads[hidx] < 0)
            s.touched[s.n_touched++] = hidx;
        s.heads[hidx] = k;
    } else
        s.tent[prev * TE + 9] = k;
    prev = k;
    while (gn >= 0) {
        int32_t *e = s.tent + gn * TE;
        uint32_t left = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t v = ~rc[q] & (uint32_t)e[5 + q];
            e[5 + q] = (int32_t)v;
            left |= v;
        }
        const int nx = e[9];
        if (left == 0)
            s.tent[prev * TE + 9] = nx;  // pruned
        else
            prev = gn;
        gn = nx;
    }
}

// ref: src/fsg_history.c:208-232 -- (state, left context) lists in ascending order
__device__ void end_frame(FsgUtt &s)
{
    // insertion sort of the few touched list heads
    for (int i = 1; i < s.n_touched; ++i) {
        const int v = s.touched[i];
        int j = i - 1;
        for (; j >= 0 && s.touched[j] > v; --j)
            s.touched[j + 1] = s.touched[j];
        s.touched[j + 1] = v;
    }
    for (int i = 0; i < s.n_touched; ++i) {
        const int hidx = s.touched[i];
        for (int gn = s.heads[hidx]; gn >= 0; gn = s.tent[gn * TE + 9])
            hist_append(s, s.tent + gn * TE);
        s.heads[hidx] = -1;
    }
    s.n_touched = 0;
    s.n_tent = 0;
}

// ref: src/fsg_search.c:543-591
__device__ void null_prop(FsgUtt &s)
{
    const int32_t thresh = s.bestscore + s.wbeam;
    const int n = s.n_hist;
    for (int bp = s.bpidx_start; bp < n; ++bp) {
        int32_t he[9];
#pragma unroll
        for (int k = 0; k < 9; ++k)
            he[k] = s.hist[(size_t)bp * 9 + k];
        const int st = he[0] >= 0 ? s.link4[he[0] * 4 + 1] : s.g->start;
        for (int a = s.arc_off[st]; a < s.arc_off[st + 1]; ++a) {
            if (s.link4[a * 4 + 3] != -1)
                continue;
            const int32_t newscore = he[1] + (s.link4[a * 4 + 2] >> SENSCR_SHIFT);
            if (newscore >= thresh)
                entry_add(s, a, he[3], newscore, bp, he[4], reinterpret_cast<const uint32_t *>(he + 5));
        }
    }
}

// ref: src/fsg_search.c:597-662
__device__ void word_trans(FsgUtt &s)
{
    const int32_t thresh = s.bestscore + s.beam, nf = s.frame + 1;
    const int n = s.n_hist;
    for (int bp = s.bpidx_start; bp < n; ++bp) {
        const int32_t *he = s.hist + (size_t)bp * 9;
        const int d = he[0] >= 0 ? s.link4[he[0] * 4 + 1] : s.g->start;
        const int lc = he[4];
        const int32_t score = he[1];
        for (int root = s.root[d]; root >= 0; root = s.pnode8[root * 8 + 6]) {
            const int rc = s.pnode8[root * 8 + 3];
            if ((s.ctxt[root * 4 + (lc >> 5)] & (1u << (lc & 31)))
                && ((uint32_t)he[5 + (rc >> 5)] & (1u << (rc & 31)))) {
                const int32_t newscore = score + s.pnode8[root * 8 + 2];
                int32_t *h = s.phmm + root * PH;
                if (newscore > thresh && newscore > h[0]) {
                    if (h[12] < nf)
                        s.nxt[s.n_nxt++] = root;
                    ph_enter(h, newscore, bp, nf);
                }
            }
        }
    }
}

// ---------------------------------------------------------------- active-list scoring
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int warp_incl_scan(int v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(FULL, v, o);
        if (lane >= o)
            v += n;
    }
    return v;
}

struct FsgScore {            // per-utterance view of the scoring state (every lane holds it)
    const float *feat;       // the utterance's features [T][blk]
    const int4 *tn_s;        // dense top-N of K1, [cs][G]; g0 = the utterance's first frame
    const uchar4 *tn_c;
    const uint32_t *tie;     // [cs][tie_w] bit (global frame): integer scores tie on that step
    int64_t G, g0, tie_w;
    uint32_t *bits;          // acmod's senone_active_vec
    uint16_t *srt, *ev;      // active senones ascending; the evaluated list (with bridging entries)
    int16_t *scr;            // [n_sen] raw senone scores of the frame (listed entries only)
    int4 *cur_s;             // [CS] raw top-N scores of the frame (scanned codebooks only)
    uchar4 *cur_c;           // [CS] codewords of the last scan = the list the reference carries
    uchar4 *cur_n;           // [CS] normalised scores of the frame
    int32_t *last_t;         // [CS] frame of the codebook-stream's last scan (-1: never)
    uchar4 *snap_c;          // [CS] cur_c / last_t as they stood after the last odd frame: the
    int32_t *snap_t;         //      reference's history slot 1, which a second pass starts from
    float *sd;               // shared, [128]: exact distances of one codebook-stream
    const uint8_t *lut;      // shared, [256]
};

__device__ __forceinline__ float fsg_gau_dist(const float *__restrict__ rec, const float *__restrict__ x, int L)
{
    // the reference's operation order, no contraction (ref: src/ptm_mgau.c:63-68, 150-225)
    float d = rec[0];
    for (int j = 0; j < L; ++j) {
        const float diff = __fsub_rn(x[j], rec[1 + j]);
        const float sq = __fmul_rn(diff, diff);
        const float c = __fmul_rn(sq, rec[1 + L + j]);
        d = __fsub_rn(d, c);
    }
    return d;
}

// A step on which integer scores tie: the reference's result depends on the list it carried in
// (ref: src/ptm_mgau.c:70-84, 139-148: eval_topn re-scores the carried codewords in their
// order, eval_cb inserts newcomers in front of equals).  Replayed literally.  eval_topn runs for
// EVERY codebook on every frame (ref :234-237), scanned or not, so the carried list has been
// re-scored and stably re-sorted on each frame since the last scan: lane 0 first catches it up
// through those frames, then all densities of frame t are evaluated across the warp and lane 0
// runs the insertion procedure.
template <int N>
__device__ void replay_tie(const DevModel &m, const FsgScore &q, int cs, int t, int lane)
{
    const int cb = cs / m.n_feat, f = cs - cb * m.n_feat;
    const int RL = m.rec_len[f], L = m.featlen[f], ND = m.n_density;
    const float *rec = m.gau + gau_offset(m, cb, f);
    const float *x = q.feat + (int64_t)t * m.blk + m.featoff[f];
    for (int n = lane; n < ND; n += 32)
        q.sd[n] = fsg_gau_dist(rec + (int64_t)n * RL, x, L);
    __syncwarp();
    if (lane == 0) {
        TcTopN<N> tn;
        const uchar4 c = q.cur_c[cs];
        const int cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int k = 0; k < N; ++k) {
            tn.c[k] = cc[k];
            tn.s[k] = INT32_MIN;
        }
#pragma unroll 1
        for (int tt = q.last_t[cs] + 1; tt < t; ++tt) {  // eval_topn of the frames in between
            const float *xx = q.feat + (int64_t)tt * m.blk + m.featoff[f];
#pragma unroll 1
            for (int i = 0; i < N; ++i) {
                int32_t ci = tn.c[0];
#pragma unroll
                for (int k = 1; k < N; ++k)
                    if (k == i)
                        ci = tn.c[k];
                const int32_t sc = __float2int_rz(fsg_gau_dist(rec + (int64_t)ci * RL, xx, L));
#pragma unroll
                for (int k = 0; k < N; ++k)
                    if (k == i)
                        tn.s[k] = sc;
                tn.settle(i);
            }
        }
#pragma unroll 1
        for (int i = 0; i < N; ++i) {
            int32_t ci = tn.c[0];
#pragma unroll
            for (int k = 1; k < N; ++k)
                if (k == i)
                    ci = tn.c[k];
            const int32_t sc = __float2int_rz(q.sd[ci]);
#pragma unroll
            for (int k = 0; k < N; ++k)
                if (k == i)
                    tn.s[k] = sc;
            tn.settle(i);
        }
#pragma unroll 1
        for (int cw = 0; cw < ND; ++cw) {
            const float d = q.sd[cw];
            if (d < __int2float_rn(tn.s[N - 1]))
                continue;
            if (tn.has(cw))
                continue;
            tn.insert(__float2int_rz(d), cw);
        }
        int4 sv = make_int4(INT32_MIN, INT32_MIN, INT32_MIN, INT32_MIN);
        uchar4 cv = make_uchar4(0, 0, 0, 0);
        sv.x = tn.s[0];
        cv.x = (unsigned char)tn.c[0];
        if (N > 1) {
            sv.y = tn.s[N > 1 ? 1 : 0];
            cv.y = (unsigned char)tn.c[N > 1 ? 1 : 0];
        }
        if (N > 2) {
            sv.z = tn.s[N > 2 ? 2 : 0];
            cv.z = (unsigned char)tn.c[N > 2 ? 2 : 0];
        }
        if (N > 3) {
            sv.w = tn.s[N > 3 ? 3 : 0];
            cv.w = (unsigned char)tn.c[N > 3 ? 3 : 0];
        }
        q.cur_s[cs] = sv;
        q.cur_c[cs] = cv;
        q.last_t[cs] = t;
    }
    __syncwarp();
}

// Scores of frame t for the senones of the `n_act` active HMMs (+ the bridging entries of the
// uint8 delta list).  Returns the best (smallest) raw score; the value the search sees for
// senone s is (int16)(scr[s] - best) (ref: src/ptm_mgau.c:398-400).
// FIX = the bundled models' shape at compile time (3 streams, top-4, 128 densities).
template <bool FIX>
__device__ int score_active_frame(const DevModel &m, const FsgUtt &s, const FsgScore &q, int t,
                                  int n_act, int lane, int &n_ev_out)
{
    const int NF = FIX ? 3 : m.n_feat, N = FIX ? 4 : m.topn;
    const int nw = (m.n_sen + 31) >> 5, E = m.n_emit, CS = m.n_mgau * NF;
    // fsg_search_sen_active: acmod_clear_active + acmod_activate_hmm (ref :309-328)
    for (int w = lane; w < nw; w += 32)
        q.bits[w] = 0u;
    __syncwarp();
    for (int i = lane; i < n_act; i += 32) {
        const uint16_t *sid = m.sseq + (size_t)s.pnode8[s.act[i] * 8 + 0] * E;
        for (int j = 0; j < E; ++j)
            atomicOr(&q.bits[sid[j] >> 5], 1u << (sid[j] & 31));
    }
    __syncwarp();
    // acmod_flags2list (ref: src/acmod.c:947-999): ascending, gaps above 255 bridged
    const int per = (nw + 31) >> 5;
    int cnt = 0;
    for (int k = 0; k < per; ++k) {
        const int w = lane * per + k;
        if (w < nw)
            cnt += __popc(q.bits[w]);
    }
    int off = warp_incl_scan(cnt, lane);
    const int n_srt = __shfl_sync(FULL, off, 31);
    off -= cnt;
    for (int k = 0; k < per; ++k) {
        const int w = lane * per + k;
        if (w < nw)
            for (uint32_t x = q.bits[w]; x; x &= x - 1)
                q.srt[off++] = (uint16_t)(w * 32 + __ffs((int)x) - 1);
    }
    __syncwarp();
    int n_ev = 0;
    for (int k0 = 0; k0 < n_srt; k0 += 32) {
        const int k = k0 + lane;
        int c = 0, sen = 0, prev = 0;
        if (k < n_srt) {
            sen = q.srt[k];
            prev = k ? q.srt[k - 1] : 0;
            const int delta = sen - prev;
            c = (delta > 255 ? (delta - 1) / 255 : 0) + 1;
        }
        const int pos = warp_incl_scan(c, lane);
        if (k < n_srt) {
            int o = n_ev + pos - c;
            for (int b = 1; b < c; ++b)
                q.ev[o++] = (uint16_t)(prev + 255 * b);
            q.ev[o] = (uint16_t)sen;
        }
        n_ev += __shfl_sync(FULL, pos, 31);
    }
    __syncwarp();
    n_ev_out = n_ev;
    if (n_ev == 0)
        return 0;
    // ptm_mgau_calc_cb_active (ref: src/ptm_mgau.c:297-321)
    unsigned long long cbm0 = 0ull, cbm1 = 0ull;  // up to 128 codebooks
    for (int i = lane; i < n_ev; i += 32) {
        const int cb = m.sen2cb[q.ev[i]];
        if (cb < 64)
            cbm0 |= 1ull << cb;
        else
            cbm1 |= 1ull << (cb - 64);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cbm0 |= __shfl_xor_sync(FULL, cbm0, o);
        cbm1 |= __shfl_xor_sync(FULL, cbm1, o);
    }
    // the scanned codebooks' top-N lists of this frame
    const int64_t g = q.g0 + t;
    bool any_tie = false;
    for (int cs = lane; cs < CS; cs += 32) {
        const int cb = cs / NF;
        const bool on = cb < 64 ? (cbm0 >> cb) & 1ull : (cbm1 >> (cb - 64)) & 1ull;
        if (!on)
            continue;
        const bool tie = (q.tie[(int64_t)cs * q.tie_w + (g >> 5)] >> (g & 31)) & 1u;
        if (tie) {
            any_tie = true;
            continue;
        }
        q.cur_s[cs] = q.tn_s[(int64_t)cs * q.G + g];
        q.cur_c[cs] = q.tn_c[(int64_t)cs * q.G + g];
        q.last_t[cs] = t;
    }
    if (__any_sync(FULL, any_tie)) {
        for (int cs = 0; cs < CS; ++cs) {  // warp-uniform walk; ties are rare
            const int cb = cs / NF;
            const bool on = cb < 64 ? (cbm0 >> cb) & 1ull : (cbm1 >> (cb - 64)) & 1ull;
            if (!on || !((q.tie[(int64_t)cs * q.tie_w + (g >> 5)] >> (g & 31)) & 1u))
                continue;
            switch (N) {
            case 1: replay_tie<1>(m, q, cs, t, lane); break;
            case 2: replay_tie<2>(m, q, cs, t, lane); break;
            case 3: replay_tie<3>(m, q, cs, t, lane); break;
            default: replay_tie<4>(m, q, cs, t, lane); break;
            }
        }
    }
    __syncwarp();
    // ptm_mgau_codebook_norm (ref :264-295): per stream, over the active codebooks
    int nm[SSB_MAX_FEAT];
#pragma unroll
    for (int f = 0; f < SSB_MAX_FEAT; ++f)
        nm[f] = WORST_SCORE;
    for (int cs = lane; cs < CS; cs += 32) {
        const int cb = cs / NF, f = cs - cb * NF;
        const bool on = cb < 64 ? (cbm0 >> cb) & 1ull : (cbm1 >> (cb - 64)) & 1ull;
        if (!on)
            continue;
        const int v = q.cur_s[cs].x >> SENSCR_SHIFT;
#pragma unroll
        for (int ff = 0; ff < SSB_MAX_FEAT; ++ff)
            if (ff == f)
                nm[ff] = max(nm[ff], v);
    }
#pragma unroll
    for (int f = 0; f < SSB_MAX_FEAT; ++f)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            nm[f] = max(nm[f], __shfl_xor_sync(FULL, nm[f], o));
    for (int cs = lane; cs < CS; cs += 32) {
        const int cb = cs / NF, f = cs - cb * NF;
        const bool on = cb < 64 ? (cbm0 >> cb) & 1ull : (cbm1 >> (cb - 64)) & 1ull;
        if (!on)
            continue;
        int nf = nm[0];
#pragma unroll
        for (int ff = 1; ff < SSB_MAX_FEAT; ++ff)
            if (ff == f)
                nf = nm[ff];
        const int4 r = q.cur_s[cs];
        q.cur_n[cs] = make_uchar4((unsigned char)min(MAX_NEG_ASCR, nf - (r.x >> SENSCR_SHIFT)),
                                  (unsigned char)min(MAX_NEG_ASCR, nf - (r.y >> SENSCR_SHIFT)),
                                  (unsigned char)min(MAX_NEG_ASCR, nf - (r.z >> SENSCR_SHIFT)),
                                  (unsigned char)min(MAX_NEG_ASCR, nf - (r.w >> SENSCR_SHIFT)));
    }
    __syncwarp();
    // ptm_mgau_senone_eval (ref :326-403)
    int best = INT32_MAX;
    const int ND = FIX ? 128 : m.n_density;
    for (int i = lane; i < n_ev; i += 32) {
        const int sen = q.ev[i];
        const int cb = m.sen2cb[sen];
        int ascore = 0;
        for (int f = 0; f < NF; ++f) {
            const uchar4 sv = q.cur_n[cb * NF + f];
            const uchar4 cv = q.cur_c[cb * NF + f];
            const int sc[4] = {sv.x, sv.y, sv.z, sv.w};
            const int cw[4] = {cv.x, cv.y, cv.z, cv.w};
            int fden = 0;
            for (int k = 0; k < N; ++k) {
                const int v = m.mixw[(int64_t)(f * ND + cw[k]) * m.n_sen + sen] + sc[k];
                if (k == 0)
                    fden = v;
                else {
                    // min(a, b) - lut[|a - b|]  (= the reference's branchy form, senone_mix.cu: logadd8)
                    fden = min(fden, v) - (int)q.lut[__sad(fden, v, 0u)];
                }
            }
            ascore += fden;
        }
        q.scr[sen] = (int16_t)ascore;
        best = min(best, ascore);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        best = min(best, __shfl_xor_sync(FULL, best, o));
    __syncwarp();
    return best;
}

// one warp per utterance; dense = senone scores of frames [g0, ...), row-major [frame][n_sen]
struct FsgActiveArgs {       // mode "compallsen = no"
    const float *feat;       // [G][blk]
    const int4 *tn_s;
    const uchar4 *tn_c;
    const uint32_t *tie;
    int64_t G, tie_w;
    const int64_t *aws_off;  // [U+1] offsets (int32 units) into aws
    int32_t *aws;            // scoring workspace
    uint32_t *final_active;  // [U][(n_sen+31)/32] or null
    int64_t *n_sen_eval;     // [U] or null
    uchar4 *final_topn;      // [U][CS] or null: the carried top-N codewords after the search
};

template <bool ACTIVE, bool FIX, bool FIVE>
__global__ void __launch_bounds__(128)
fsg_search_kernel(DevModel m, DevFsgSet gs, FsgActiveArgs aa, const int64_t *__restrict__ frame_off,
                  const int32_t *__restrict__ utt_graph, const int64_t *__restrict__ ws_off,
                  int32_t *__restrict__ ws, const int16_t *__restrict__ dense, int64_t g0, int u0,
                  int n_utts, int32_t *__restrict__ hist_all, int hist_cap, int tent_cap,
                  int32_t *__restrict__ n_hist_out, int64_t *__restrict__ n_eval_out,
                  int32_t *__restrict__ frames_out, int32_t *__restrict__ rv_out)
{
    const int lane = threadIdx.x & 31;
    const int u = u0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (u >= u0 + n_utts)
        return;
    __shared__ float sh_sd[4][128];
    __shared__ uint8_t sh_lut[4][256];
    if (ACTIVE) {  // a private copy per warp: warps are independent, no block barrier anywhere
        for (int i = lane; i < 256; i += 32)
            sh_lut[threadIdx.x >> 5][i] = m.lut8[i];
        __syncwarp();
    }
    FsgUtt s;
    const DevFsg *g = gs.graph + utt_graph[u];
    s.g = g;
    s.link4 = gs.link4 + (size_t)g->link_off * 4;
    s.link_flag = gs.link_flag + g->link_off;
    s.arc_off = gs.arc_off + g->arcoff_off;
    s.root = gs.root + g->root_off;
    s.pnode8 = gs.pnode8 + (size_t)g->pnode_off * 8;
    s.ctxt = gs.ctxt + (size_t)g->pnode_off * 4;
    int32_t *w = ws + ws_off[u];
    const int NP = g->n_pnode, NH = g->n_state * g->n_ciphone;
    s.phmm = w;
    s.act = s.phmm + (size_t)NP * PH;
    s.nxt = s.act + NP;
    s.heads = s.nxt + NP;
    s.touched = s.heads + NH;
    s.tent = s.touched + tent_cap;
    s.tent_cap = tent_cap;
    s.hist = hist_all + (size_t)u * hist_cap * 9;
    s.cap = hist_cap;
    const int T = (int)(frame_off[u + 1] - frame_off[u]);
    const int16_t *scr = ACTIVE ? nullptr : dense + (frame_off[u] - g0) * m.n_sen;
    const int E = m.n_emit;
    FsgScore q;
    long long n_sen_eval = 0;
    if (ACTIVE) {
        const int nw = (m.n_sen + 31) >> 5, CS = m.n_mgau * m.n_feat;
        const int cap_ev = (min(m.n_emit * NP, m.n_sen) + m.n_sen / 255 + 8 + 1) & ~1;  // (one senone per emitting state)
        int32_t *a = aa.aws + aa.aws_off[u];
        q.feat = aa.feat + frame_off[u] * m.blk;
        q.tn_s = aa.tn_s;
        q.tn_c = aa.tn_c;
        q.tie = aa.tie;
        q.G = aa.G;
        q.g0 = frame_off[u];
        q.tie_w = aa.tie_w;
        q.cur_s = reinterpret_cast<int4 *>(a);              // 16-byte aligned first
        a += 4 * CS;
        q.cur_c = reinterpret_cast<uchar4 *>(a);
        a += CS;
        q.cur_n = reinterpret_cast<uchar4 *>(a);
        a += CS;
        q.last_t = a;
        a += CS;
        q.snap_c = reinterpret_cast<uchar4 *>(a);
        a += CS;
        q.snap_t = a;
        a += CS;
        q.bits = reinterpret_cast<uint32_t *>(a);
        a += nw;
        q.srt = reinterpret_cast<uint16_t *>(a);
        a += cap_ev / 2;
        q.ev = reinterpret_cast<uint16_t *>(a);
        a += cap_ev / 2;
        q.scr = reinterpret_cast<int16_t *>(a);
        q.sd = sh_sd[threadIdx.x >> 5];
        q.lut = sh_lut[threadIdx.x >> 5];
        // the list a codebook carries before its first scan (ref: src/ptm_mgau.c:694-720)
        for (int i = lane; i < CS; i += 32) {
            q.cur_c[i] = make_uchar4(0, 1, 2, 3);
            q.last_t[i] = -1;
            q.snap_c[i] = make_uchar4(0, 1, 2, 3);
            q.snap_t[i] = -1;
        }
        for (int i = lane; i < nw; i += 32)
            q.bits[i] = 0u;
    }

    for (int i = lane; i < NP; i += 32)
        ph_clear(s.phmm + i * PH);
    for (int i = lane; i < NH; i += 32)
        s.heads[i] = -1;
    __syncwarp();
    s.n_act = s.n_nxt = s.n_hist = s.n_tent = s.n_touched = s.overflow = 0;
    long long n_eval = 0;
    if (lane == 0) {
        // fsg_search_start (ref :746-798)
        const uint32_t all[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
        s.beam_factor = 1.0f;
        s.beam = g->beam;
        s.pbeam = g->pbeam;
        s.wbeam = g->wbeam;
        s.frame = -1;
        s.bestscore = 0;
        entry_add(s, -1, -1, 0, -1, g->sil, all);
        s.bpidx_start = 0;
        null_prop(s);
        word_trans(s);
        for (int i = 0; i < s.n_nxt; ++i)
            s.act[i] = s.nxt[i];
        s.n_act = s.n_nxt;
        s.n_nxt = 0;
        s.frame = 0;
    }
    __syncwarp();
    for (int t = 0; t < T; ++t) {
        const int n_act = __shfl_sync(0xffffffffu, s.n_act, 0);
        const int16_t *ss = ACTIVE ? q.scr : scr + (size_t)t * m.n_sen;
        int sbest = 0;
        if (ACTIVE) {
            int n_ev = 0;
            sbest = score_active_frame<FIX>(m, s, q, t, n_act, lane, n_ev);
            n_sen_eval += n_ev;
            if (aa.final_topn && t == ((T & 1) ? T - 2 : T - 1)) {
                // t is the last odd frame: the state of the reference's history slot 1
                const int CS = m.n_mgau * m.n_feat;
                for (int i = lane; i < CS; i += 32) {
                    q.snap_c[i] = q.cur_c[i];
                    q.snap_t[i] = q.last_t[i];
                }
                __syncwarp();
            }
        }
        // fsg_search_hmm_eval (ref :330-398): the active HMMs in parallel
        int32_t best = WORST_SCORE;
        for (int i = lane; i < n_act; i += 32) {
            const int pn = s.act[i];
            int32_t *h = s.phmm + pn * PH;
            const uint16_t *sid = m.sseq + (size_t)s.pnode8[pn * 8 + 0] * E;
            constexpr int EN = FIVE ? 5 : 3;  // emitting states (hmm_vit_eval_3st_lr / _5st_lr)
            int32_t sc[EN], hi[EN], o_s = h[10], o_h = h[11];
            int sv[EN];
#pragma unroll
            for (int j = 0; j < EN; ++j) {
                sc[j] = h[j];
                hi[j] = h[5 + j];
                sv[j] = ACTIVE ? (int)(int16_t)(ss[sid[j]] - sbest) : (int)ss[sid[j]];
            }
            const int32_t b = hmm_step<EN>(m.tp + (size_t)s.pnode8[pn * 8 + 1] * EN * (EN + 1), sv, sc, hi,
                                           o_s, o_h);
#pragma unroll
            for (int j = 0; j < EN; ++j) {
                h[j] = sc[j];
                h[5 + j] = hi[j];
            }
            h[10] = o_s;
            h[11] = o_h;
            h[13] = b;
            best = max(best, b);
        }
        for (int o = 16; o > 0; o >>= 1)
            best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
        __syncwarp();
        if (lane == 0) {
            s.bpidx_start = s.n_hist;
            if (s.n_act > 0) {
                n_eval += s.n_act;
                if (g->maxhmmpf != -1 && s.n_act > g->maxhmmpf) {
                    if (s.beam_factor > 0.1) {
                        s.beam_factor *= 0.9f;
                        s.beam = (int32_t)(g->beam * s.beam_factor);
                        s.pbeam = (int32_t)(g->pbeam * s.beam_factor);
                        s.wbeam = (int32_t)(g->wbeam * s.beam_factor);
                    }
                } else {
                    s.beam_factor = 1.0f;
                    s.beam = g->beam;
                    s.pbeam = g->pbeam;
                    s.wbeam = g->wbeam;
                }
                s.bestscore = best;
            }
            // fsg_search_hmm_prune_prop (ref :497-538); "prepended" lists are walked backwards
            const int32_t thresh = s.bestscore + s.beam, phone_thresh = s.bestscore + s.pbeam,
                          word_thresh = s.bestscore + s.wbeam;
            const int32_t nf = s.frame + 1;
            for (int i = s.n_act - 1; i >= 0; --i) {
                const int pn = s.act[i];
                int32_t *h = s.phmm + pn * PH;
                if (h[13] >= thresh) {
                    if (h[12] == s.frame) {
                        h[12] = nf;
                        s.nxt[s.n_nxt++] = pn;
                    }
                    if (!s.pnode8[pn * 8 + 4]) {
                        if (h[10] >= phone_thresh) {
                            // fsg_search_pnode_trans (ref :400-428)
                            for (int child = s.pnode8[pn * 8 + 5]; child >= 0; child = s.pnode8[child * 8 + 6]) {
                                const int32_t newscore = h[10] + s.pnode8[child * 8 + 2];
                                int32_t *c = s.phmm + child * PH;
                                if (newscore > thresh && newscore > c[0]) {
                                    if (c[12] < nf)
                                        s.nxt[s.n_nxt++] = child;
                                    ph_enter(c, newscore, h[11], nf);
                                }
                            }
                        }
                    } else if (h[10] >= word_thresh) {
                        // fsg_search_pnode_exit (ref :430-489)
                        const uint32_t all[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
                        const int link = s.pnode8[pn * 8 + 5];
                        const uint32_t *rc = (s.link_flag[link] & 1) ? all : s.ctxt + pn * 4;
                        entry_add(s, link, s.frame, h[10], h[11], s.pnode8[pn * 8 + 3], rc);
                    }
                }
            }
            end_frame(s);
            null_prop(s);
            end_frame(s);
            word_trans(s);
            for (int i = s.n_act - 1; i >= 0; --i) {
                int32_t *h = s.phmm + s.act[i] * PH;
                if (h[12] == s.frame)
                    ph_clear(h);
            }
            int32_t *tmp = s.act;
            s.act = s.nxt;
            s.nxt = tmp;
            s.n_act = s.n_nxt;
            s.n_nxt = 0;
            ++s.frame;
        }
        // every lane follows the list swap
        s.act = reinterpret_cast<int32_t *>(__shfl_sync(0xffffffffu, (unsigned long long)s.act, 0));
        __syncwarp();
    }
    if (lane == 0) {
        n_hist_out[u] = s.n_hist;
        n_eval_out[u] = n_eval;
        frames_out[u] = s.frame;
        rv_out[u] = s.overflow ? -2 : 0;
        if (ACTIVE && aa.n_sen_eval)
            aa.n_sen_eval[u] = n_sen_eval;
    }
    if (ACTIVE && aa.final_topn) {
        // What a second pass on the same decoder starts from: frame 0 copies history slot
        // n_fast_hist-1 = 1 (ref: src/ptm_mgau.c:426-440), i.e. the lists as they stood after the
        // last ODD frame tl -- the codewords of the last scan up to tl, re-scored and stably
        // re-sorted by eval_topn on every frame since (replayed from the most recent frame on
        // which their scores are pairwise distinct: the order is forced there).
        const int CS = m.n_mgau * m.n_feat, N = m.topn;
        const int tl = (T & 1) ? T - 2 : T - 1;
        __syncwarp();
        for (int cs = lane; cs < CS; cs += 32) {
            uchar4 c = make_uchar4(0, 1, 2, 3);
            if (tl >= 0) {
                c = q.snap_c[cs];
                const int lt = q.snap_t[cs];
                if (lt < tl) {
                    const int cb = cs / m.n_feat, f = cs - cb * m.n_feat;
                    const int RL = m.rec_len[f], L = m.featlen[f];
                    const float *rec = m.gau + gau_offset(m, cb, f);
                    int cc[4] = {c.x, c.y, c.z, c.w};
                    int tt0 = lt + 1;
                    for (int tt = tl; tt > lt + 1; --tt) {
                        const float *xx = q.feat + (int64_t)tt * m.blk + m.featoff[f];
                        int32_t sc[4];
                        for (int k = 0; k < N; ++k)
                            sc[k] = __float2int_rz(fsg_gau_dist(rec + (int64_t)cc[k] * RL, xx, L));
                        bool distinct = true;
                        for (int i = 0; i < N; ++i)
                            for (int j = i + 1; j < N; ++j)
                                distinct = distinct && sc[i] != sc[j];
                        if (distinct) {
                            tt0 = tt;
                            break;
                        }
                    }
                    for (int tt = tt0; tt <= tl; ++tt) {  // eval_topn: re-score in list order, settle
                        const float *xx = q.feat + (int64_t)tt * m.blk + m.featoff[f];
                        int32_t sc[4];
                        for (int i = 0; i < N; ++i) {
                            const int32_t v = __float2int_rz(fsg_gau_dist(rec + (int64_t)cc[i] * RL, xx, L));
                            int j = i;
                            const int ci = cc[i];
                            for (; j > 0 && v > sc[j - 1]; --j) {
                                sc[j] = sc[j - 1];
                                cc[j] = cc[j - 1];
                            }
                            sc[j] = v;
                            cc[j] = ci;
                        }
                    }
                    c = make_uchar4((unsigned char)cc[0], (unsigned char)cc[1], (unsigned char)cc[2],
                                    (unsigned char)cc[3]);
                }
            }
            aa.final_topn[(size_t)u * CS + cs] = c;
        }
    }
    if (ACTIVE && aa.final_active) {
        // acmod's flags as the last frame left them: what the second pass starts from
        // (ref: src/state_align_search.c:186-188 never clears them)
        const int nw = (m.n_sen + 31) >> 5;
        __syncwarp();
        for (int i = lane; i < nw; i += 32)
            aa.final_active[(size_t)u * nw + i] = q.bits[i];
    }
}

// ref: src/fsg_search.c:853-924 (find_exit; `final` = 0 while the utterance is running) and
// :1030-1142 (segmentation).
// One thread per utterance.  segs [u][max_seg][5] = link sf ef ascr lscr, first word first.
__global__ void fsg_backtrace_kernel(DevFsgSet gs, const int32_t *__restrict__ utt_graph, int u0,
                                     int n_utts, const int32_t *__restrict__ hist_all, int hist_cap,
                                     const int32_t *__restrict__ n_hist, const int32_t *__restrict__ frames,
                                     int32_t *__restrict__ exit_bp, int32_t *__restrict__ hyp_score,
                                     int32_t *__restrict__ segs, int max_seg, int32_t *__restrict__ n_seg, int final)
{
    const int u = u0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= u0 + n_utts)
        return;
    const DevFsg *g = gs.graph + utt_graph[u];
    const int32_t *link4 = gs.link4 + (size_t)g->link_off * 4;
    const int32_t *hist = hist_all + (size_t)u * hist_cap * 9;
    const int frame_idx = frames[u];
    int bpidx = n_hist[u] - 1, frm = frame_idx, last_frm = frame_idx, besthist = -1;
    int32_t bestscore = INT32_MIN;
    const int32_t *e = nullptr;
    n_seg[u] = 0;
    hyp_score[u] = 0;
    while (bpidx > 0) {
        e = hist + (size_t)bpidx * 9;
        if (e[3] <= frame_idx) {
            frm = last_frm = e[3];
            break;
        }
        --bpidx;
    }
    if (bpidx <= 0) {
        exit_bp[u] = bpidx;
        return;
    }
    while (frm == last_frm) {
        const int link = e[0];
        const int32_t score = e[1];
        if (link < 0)
            break;
        if (score == bestscore && link4[link * 4 + 1] == g->final)
            besthist = bpidx;
        else if (score > bestscore && (!final || link4[link * 4 + 1] == g->final)) {
            bestscore = score;
            besthist = bpidx;
        }
        --bpidx;
        if (bpidx < 0)
            break;
        e = hist + (size_t)bpidx * 9;
        frm = e[3];
    }
    exit_bp[u] = besthist;
    if (besthist == -1)
        return;  // "Final result does not match the grammar"
    hyp_score[u] = bestscore;
    int n = 0;
    for (int bp = besthist; bp > 0; bp = hist[(size_t)bp * 9 + 2])
        ++n;
    if (n > max_seg) {
        n_seg[u] = -n;
        return;
    }
    n_seg[u] = n;
    int32_t *so = segs + (size_t)u * max_seg * 5;
    int cur = n - 1;
    for (int bp = besthist; bp > 0; bp = hist[(size_t)bp * 9 + 2], --cur) {
        const int32_t *h = hist + (size_t)bp * 9;
        const int32_t *ph = h[2] >= 0 ? hist + (size_t)h[2] * 9 : nullptr;
        int32_t sf = ph ? ph[3] + 1 : 0;
        const int32_t ef = h[3];
        const int32_t lscr = link4[h[0] * 4 + 2] >> SENSCR_SHIFT;
        if (sf > ef)
            sf = ef;
        so[cur * 5 + 0] = h[0];
        so[cur * 5 + 1] = sf;
        so[cur * 5 + 2] = ef;
        so[cur * 5 + 3] = ph ? h[1] - ph[1] - lscr : h[1] - lscr;
        so[cur * 5 + 4] = lscr;
    }
}

int launch_fsg_search(const DevModel &m, const DevFsgSet &gs, const int64_t *frame_off,
                      const int32_t *utt_graph, const int64_t *ws_off, int32_t *ws,
                      const int16_t *dense, int64_t g0, int u0, int n_utts, int32_t *hist,
                      int hist_cap, int tent_cap, int32_t *n_hist, int64_t *n_eval, int32_t *frames,
                      int32_t *rv, cudaStream_t st)
{
    if (n_utts <= 0)
        return 0;
    if (m.n_emit != 3 && m.n_emit != 5) {
        set_error("FSG search supports 3- and 5-state HMMs, model has %d", m.n_emit);
        return -1;
    }
    const int wpb = 4;
    FsgActiveArgs none;
    memset(&none, 0, sizeof none);
    if (m.n_emit == 3)
        fsg_search_kernel<false, false, false><<<(n_utts + wpb - 1) / wpb, wpb * 32, 0, st>>>(
            m, gs, none, frame_off, utt_graph, ws_off, ws, dense, g0, u0, n_utts, hist, hist_cap, tent_cap,
            n_hist, n_eval, frames, rv);
    else
        fsg_search_kernel<false, false, true><<<(n_utts + wpb - 1) / wpb, wpb * 32, 0, st>>>(
            m, gs, none, frame_off, utt_graph, ws_off, ws, dense, g0, u0, n_utts, hist, hist_cap, tent_cap,
            n_hist, n_eval, frames, rv);
    SSB_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

size_t fsg_active_ws_ints(const DevModel &m, int n_pnode)
{
    const size_t nw = (m.n_sen + 31) / 32, CS = (size_t)m.n_mgau * m.n_feat;
    const size_t cap_ev = (std::min<size_t>((size_t)m.n_emit * (size_t)n_pnode, (size_t)m.n_sen) + m.n_sen / 255 + 8 + 1) & ~(size_t)1;
    size_t n = 4 * CS + CS + CS + CS + CS + CS + nw + cap_ev / 2 + cap_ev / 2 + ((size_t)m.n_sen + 1) / 2;
    return (n + 3) & ~(size_t)3;  // keeps every utterance's int4 block 16-byte aligned
}

int launch_fsg_search_active(const DevModel &m, const DevFsgSet &gs, const int64_t *frame_off,
                             const int32_t *utt_graph, const int64_t *ws_off, int32_t *ws,
                             const float *feat, const int4 *tn_s, const uchar4 *tn_c,
                             const uint32_t *tie, int64_t G, int64_t tie_w, const int64_t *aws_off,
                             int32_t *aws, uint32_t *final_active, int64_t *n_sen_eval,
                             uchar4 *final_topn, int n_utts, int32_t *hist, int hist_cap, int tent_cap, int32_t *n_hist,
                             int64_t *n_eval, int32_t *frames, int32_t *rv, cudaStream_t st)
{
    if (n_utts <= 0)
        return 0;
    if ((m.n_emit != 3 && m.n_emit != 5) || m.kind != SSB_SCORER_PTM || m.n_density > 128 || m.n_mgau > 128) {
        set_error("FSG search with active lists needs a PTM model with 3- or 5-state HMMs, at most 128 "
                  "densities and 128 codebooks");
        return -1;
    }
    FsgActiveArgs aa;
    aa.feat = feat;
    aa.tn_s = tn_s;
    aa.tn_c = tn_c;
    aa.tie = tie;
    aa.G = G;
    aa.tie_w = tie_w;
    aa.aws_off = aws_off;
    aa.aws = aws;
    aa.final_active = final_active;
    aa.n_sen_eval = n_sen_eval;
    aa.final_topn = final_topn;
    const int wpb = 4;
#define SSB_K4(FX, FV)                                                                          \
    fsg_search_kernel<true, FX, FV><<<(n_utts + wpb - 1) / wpb, wpb * 32, 0, st>>>(                 \
        m, gs, aa, frame_off, utt_graph, ws_off, ws, nullptr, 0, 0, n_utts, hist, hist_cap, tent_cap, \
        n_hist, n_eval, frames, rv)
    const bool fix = m.n_feat == 3 && m.topn == 4 && m.n_density == 128;
    if (m.n_emit == 3) {
        if (fix) SSB_K4(true, false); else SSB_K4(false, false);
    } else {
        if (fix) SSB_K4(true, true); else SSB_K4(false, true);
    }
#undef SSB_K4
    SSB_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

__global__ void fsg_final_topn_dense_kernel(DevModel m, const int64_t *__restrict__ frame_off, int n_utts,
                                            const uchar4 *__restrict__ tn_c, int64_t G,
                                            uchar4 *__restrict__ final_topn)
{
    const int CS = m.n_mgau * m.n_feat;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_utts * CS)
        return;
    const int u = (int)(i / CS), cs = (int)(i - (int64_t)u * CS);
    const int T = (int)(frame_off[u + 1] - frame_off[u]);
    const int tl = (T & 1) ? T - 2 : T - 1;  // history slot 1 = the last odd frame
    final_topn[i] = tl >= 0 ? tn_c[(int64_t)cs * G + frame_off[u] + tl] : make_uchar4(0, 1, 2, 3);
}

int launch_fsg_final_topn_dense(const DevModel &m, const int64_t *frame_off, int n_utts,
                                const uchar4 *tn_c, int64_t G, uchar4 *final_topn, cudaStream_t st)
{
    if (n_utts <= 0)
        return 0;
    const int64_t n = (int64_t)n_utts * m.n_mgau * m.n_feat;
    fsg_final_topn_dense_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m, frame_off, n_utts, tn_c, G,
                                                                             final_topn);
    SSB_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

int launch_fsg_backtrace(const DevFsgSet &gs, const int32_t *utt_graph, int u0, int n_utts,
                         const int32_t *hist, int hist_cap, const int32_t *n_hist,
                         const int32_t *frames, int32_t *exit_bp, int32_t *hyp_score, int32_t *segs,
                         int max_seg, int32_t *n_seg, int final, cudaStream_t st)
{
    if (n_utts <= 0)
        return 0;
    fsg_backtrace_kernel<<<(n_utts + 63) / 64, 64, 0, st>>>(gs, utt_graph, u0, n_utts, hist, hist_cap,
                                                           n_hist, frames, exit_bp, hyp_score, segs,
                                                           max_seg, n_seg, final);
    SSB_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

}  // namespace ssb
