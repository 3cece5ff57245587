// chain_viterbi.cu -- K3: left-to-right chain Viterbi with token stack, and the
// backtrace that turns the stack into state segmentations.
//
// Replaces state_align_search_start/step/finish and their helpers
// (ref: src/state_align_search.c:46-268) together with hmm_vit_eval_3st_lr /
// hmm_vit_eval_5st_lr (ref: src/hmm.c:166-304, 482-567), hmm_clear / hmm_enter /
// hmm_normalize (ref: src/hmm.c:121-161).  Integer max-plus arithmetic only; every
// tie-break, clamp and "stale" value of the reference is kept.
//
// What the reference decides frame by frame with `hmm_frame(hmm)` bookkeeping is
// data independent when the window ends `ef` do not decrease along the chain (the
// reference's own case: phones inherit their word's window, ref:
// src/ps_alignment.c:168-305): phone i is evaluated on frames [enter[i], last[i]]
// with  enter[i] = max(enter[i-1], sf[i], 1)  and  last[i] = max(enter[i], ef[i]).
// The host planner computes enter[] once per utterance (api.cu: plan_utterance);
// the kernel only moves scores.
//
// B200 mapping: one CTA per utterance -- a single warp for ordinary sentences, up
// to 32 warps for book-length chains -- with the whole HMM state (score, history,
// exit score/history per phone) resident in shared memory.  Only the band of
// phones that is active on frame t is touched.  HBM traffic per state-frame is
// one int16 senone score in (pre-gathered, coalesced) and one {history, score}
// token out (the reference's own 8-byte record, ref: state_align_search.h:61-64).
#include "hmm_step.cuh"

namespace ssb {

__device__ __forceinline__ int32_t block_max(int32_t v, int32_t *red, int nwarps)
{
    for (int o = 16; o > 0; o >>= 1)
        v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (nwarps == 1)
        return v;
    if ((threadIdx.x & 31) == 0)
        red[threadIdx.x >> 5] = v;
    __syncthreads();
    int32_t r = red[0];
    for (int w = 1; w < nwarps; ++w)
        r = max(r, red[w]);
    __syncthreads();
    return r;
}

__device__ __forceinline__ void cta_sync(int nwarps)
{
    if (nwarps == 1)
        __syncwarp();
    else
        __syncthreads();
}

// the synchronisation domain of one utterance: a group of L lanes (L < 32), or the CTA
template <int L>
__device__ __forceinline__ void group_sync(int nwarps)
{
    if (L < 32)
        __syncwarp();  // (the whole warp: its groups stay converged)
    else
        cta_sync(nwarps);
}

template <int L>
__device__ __forceinline__ int32_t group_max(int32_t v, int32_t *red, int nwarps)
{
    if (L < 32) {
#pragma unroll
        for (int o = L / 2; o > 0; o >>= 1)
            v = max(v, __shfl_xor_sync(0xffffffffu, v, o));  // (xor < L stays inside the group)
        return v;
    }
    return block_max(v, red, nwarps);
}

// Shared (or, for very long chains, global) state of one utterance's chain:
//   sc[E][np] hi[E][np] osc[np] ohi[np]
constexpr int K3_RING = 64;  // ring of HMM states in shared memory (band narrower than this)
#define IX(i) (RING ? ((i) & (K3_RING - 1)) : (i))

// int32 words between the shared-memory states of two groups of a warp: the groups' lanes touch the
// same offsets at the same time, so the regions are staggered by L banks (no two-way conflicts)
__host__ __device__ inline int k3_group_stride(int words, int lanes)
{
    return words + ((lanes - words % 32) + 32) % 32;
}

template <int E, bool RING, int L>
__global__ void __launch_bounds__(L < 32 ? 128 : 1024)
chain_viterbi_kernel(DevModel m, DevPlan p, const int16_t *__restrict__ chain_scr,
                     int2 *__restrict__ tokens, int32_t *__restrict__ spill,
                     int64_t spill_stride, int32_t *__restrict__ utt_best,
                     int32_t *__restrict__ utt_renorm, int32_t *__restrict__ fin_hist,
                     int32_t *__restrict__ fin_score, int smem_phones)
{
    extern __shared__ int32_t sh_all[];
    __shared__ int32_t red[32];
    // L < 32: a GROUP of L lanes per utterance, 32 / L utterances per warp -- the evaluated band is
    // a handful of phones, so that a whole warp per utterance leaves most lanes idle.  The groups
    // of a warp stay CONVERGED (that is the point: one instruction serves four utterances): they
    // walk the frames in step up to the longest of them, a group past its last frame only
    // keeps the warp-wide synchronisations company.  L == 32: the CTA.
    const int tid = L < 32 ? (int)(threadIdx.x % L) : (int)threadIdx.x;
    const int nthr = L < 32 ? L : (int)blockDim.x;
    const int u_raw = L < 32 ? (int)(blockIdx.x * (blockDim.x / L) + threadIdx.x / L) : (int)blockIdx.x;
    const bool live = u_raw < p.n_utts;
    const int u = live ? u_raw : 0;
    const int nwarps = L < 32 ? 1 : (int)(blockDim.x >> 5);
    int32_t *sh = sh_all + (L < 32 ? (size_t)(threadIdx.x / L) * (size_t)k3_group_stride(smem_phones * (2 * E + 2), L) : (size_t)0);
    const int64_t g0 = p.frame_off[u];
    const int64_t ph0 = p.phone_off[u];
    const int np = (int)(p.phone_off[u + 1] - ph0);
    const int ns = np * E;
    const bool valid = live && np > 0;
    if (live && np == 0 && tid == 0) {
        utt_best[u] = 0;
        utt_renorm[u] = 0;
        fin_hist[u] = -1;
        fin_score[u] = WORST_SCORE;
    }
    if (L == 32 && !valid)
        return;
    const int T = valid ? (int)(p.frame_off[u + 1] - g0) : 0;
    int Tloop = T;
    if (L < 32)
        for (int o = 16; o > 0; o >>= 1)
            Tloop = max(Tloop, __shfl_xor_sync(0xffffffffu, Tloop, o));
    // RING: only the band of phones that is alive (a word window's worth) is kept, phone i in slot
    // i mod 64 -- a phone that has left the band is never looked at again (its successor enters
    // while it is still being evaluated: api.cu plan_enter) -- so that book-length chains run in
    // shared memory on one warp instead of spilling to L2 behind block barriers
    int32_t *st = RING ? sh : (np <= smem_phones ? sh : spill + (int64_t)u * spill_stride);
    const int W = RING ? K3_RING : np;
    int32_t *sc = st;                  // [E][W]
    int32_t *hi = st + (size_t)E * W;  // [E][W]
    int32_t *osc = hi + (size_t)E * W;
    int32_t *ohi = osc + W;
    const int32_t *enter = p.enter_plan + ph0;
    const int32_t *sf = p.sf + ph0, *ef = p.ef + ph0, *tmat = p.tmat + ph0;
    const int16_t *scr = chain_scr + p.scr_off[u];
    int2 *tok = tokens + p.scr_off[u];

    // hmm_clear on every phone, hmm_enter(hmms, 0, 0, 0) (ref: state_align_search.c:46-55)
    for (int i = tid; i < (valid ? W : 0); i += nthr) {
#pragma unroll
        for (int j = 0; j < E; ++j) {
            sc[j * W + i] = WORST_SCORE;
            hi[j * W + i] = -1;
        }
        osc[i] = WORST_SCORE;
        ohi[i] = -1;
    }
    group_sync<L>(nwarps);
    if (tid == 0 && valid) {
        sc[0] = 0;
        hi[0] = 0;
    }
    group_sync<L>(nwarps);

    int32_t best = 0, n_renorm = 0;
    int lo = 0, hiq = 0;  // phones [lo, hiq] are evaluated on frame t
    // the two plan values the band bookkeeping looks at on every frame, kept in registers:
    // the entry frame of the next phone (-1: none / never) and the last frame of phone lo
    int nx_en = valid && np > 1 ? enter[1] : -1;
    int lo_end = valid ? max(enter[0], ef[0]) : 0;
    for (int t = 0; t < Tloop; ++t) {
        const int nf = t + 1;
        const bool on = t < T;  // (L < 32: this group's utterance is still running)
        if (on) {
            while (nx_en >= 0 && nx_en <= t) {
                ++hiq;
                nx_en = hiq + 1 < np ? enter[hiq + 1] : -1;
            }
            while (lo < hiq && lo_end < t) {
                ++lo;
                lo_end = max(enter[lo], ef[lo]);
            }
        }
        const bool lo_alive = on && lo_end >= t;  // lo == hiq may have expired too
        // renormalize_hmms (ref: state_align_search.c:57-64,193-197; hmm.c:150-161):
        // every phone, alive or not, whose scores are above WORST_SCORE
        if (on && best - 0x300000 < WORST_SCORE) {
            for (int i = (RING ? lo : 0) + tid; i <= (RING ? hiq : np - 1); i += nthr) {
#pragma unroll
                for (int j = 0; j < E; ++j)
                    if (sc[j * W + IX(i)] > WORST_SCORE)
                        sc[j * W + IX(i)] -= best;
                if (osc[IX(i)] > WORST_SCORE)
                    osc[IX(i)] -= best;
            }
            ++n_renorm;
            if (L == 32)
                group_sync<L>(nwarps);
        }
        if (L < 32)
            group_sync<L>(nwarps);
        // evaluate_hmms (ref :66-86)
        int32_t lb = WORST_SCORE;
        const int16_t *scr_t = scr + (int64_t)t * ns;
        int2 *tok_t = tok + (int64_t)t * ns;
        if (lo_alive) {
            for (int i = lo + tid; i <= hiq; i += nthr) {
                int32_t s[E], h[E], o_s = osc[IX(i)], o_h = ohi[IX(i)];
                int ss[E];
#pragma unroll
                for (int j = 0; j < E; ++j) {
                    s[j] = sc[j * W + IX(i)];
                    h[j] = hi[j * W + IX(i)];
                }
                if (p.banded) {
                    const int16_t *sp = chain_scr + p.scr_boff[ph0 + i] + (int64_t)t * E;
#pragma unroll
                    for (int j = 0; j < E; ++j)
                        ss[j] = sp[j];
                } else {
#pragma unroll
                    for (int j = 0; j < E; ++j)
                        ss[j] = scr_t[i * E + j];
                }
                int32_t b = hmm_step<E>(m.tp + (size_t)tmat[i] * E * (E + 1), ss, s, h, o_s, o_h);
                lb = max(lb, b);
#pragma unroll
                for (int j = 0; j < E; ++j) {
                    sc[j * W + IX(i)] = s[j];
                    hi[j * W + IX(i)] = h[j];
                }
                osc[IX(i)] = o_s;
                ohi[IX(i)] = o_h;
            }
        }
        group_sync<L>(nwarps);
        {
            const int32_t gm = group_max<L>(lb, red, nwarps);
            if (on)
                best = gm;
        }
        // prune_hmms + phone_transition + record_transitions (ref :88-175), one target phone
        // per thread.  Targets: the evaluated band plus every phone the plan enters at nf
        // (several when the reference's transition loop cascades along the chain).
        const int first = lo_alive ? lo : hiq + 1;
        int last_t = on ? hiq : first - 1;
        if (on && nx_en == nf) {
            ++last_t;
            while (last_t + 1 < np && enter[last_t + 1] == nf)
                ++last_t;
        }
        if (RING && last_t > hiq) {
            // the slots of the phones entering now: their previous tenants have left the band
            // (hmm_clear), before anybody reads a neighbour's exit score
            for (int i = hiq + 1 + tid; i <= last_t; i += nthr) {
#pragma unroll
                for (int j = 0; j < E; ++j) {
                    sc[j * W + IX(i)] = WORST_SCORE;
                    hi[j * W + IX(i)] = -1;
                }
                osc[IX(i)] = WORST_SCORE;
                ohi[IX(i)] = -1;
            }
            group_sync<L>(nwarps);
        }
        for (int i = first + tid; i <= last_t; i += nthr) {
            const bool was_active = i <= hiq;  // evaluated on frame t
            bool now = was_active;             // hmm_frame(hmm) >= t after this step
            if (i > 0) {
                const int hprev = i - 1;
                if (!was_active) {
                    // first entry: unconditional, with whatever exit score the previous
                    // phone holds (WORST_SCORE / -1 if it has never been evaluated)
                    sc[IX(i)] = osc[IX(hprev)];
                    hi[IX(i)] = ohi[IX(hprev)];
                    now = true;
                } else {
                    const bool prev_eval = hprev >= first && hprev <= hiq;
                    // hmm_frame(prev) == nf: kept by prune_hmms, or entered in this pass
                    const bool prev_nf = (prev_eval && nf <= ef[hprev]) || enter[hprev] == nf;
                    if (prev_nf && nf >= sf[i]) {
                        const int32_t o = osc[IX(hprev)];
                        if (o > sc[IX(i)]) {
                            sc[IX(i)] = o;  // state 0
                            hi[IX(i)] = ohi[IX(hprev)];
                        }
                    }
                }
            }
            if (now) {
                int2 *tk = tok_t + i * E;
                if (p.banded)
                    tk = tokens + p.tok_boff[ph0 + i] + (int64_t)t * E;
#pragma unroll
                for (int j = 0; j < E; ++j) {
                    const int si = i * E + j;
                    tk[j] = make_int2(hi[j * W + IX(i)], sc[j * W + IX(i)]);
                    hi[j * W + IX(i)] = si;
                }
            }
        }
        group_sync<L>(nwarps);
    }
    if (tid == 0 && valid) {
        utt_best[u] = best;
        utt_renorm[u] = n_renorm;
        // (a last phone that was never entered still holds hmm_clear's values)
        const bool reached = !RING || (enter[np - 1] >= 0 && enter[np - 1] <= T);
        fin_hist[u] = reached ? ohi[IX(np - 1)] : -1;
        fin_score[u] = reached ? osc[IX(np - 1)] : WORST_SCORE;
    }
}

int launch_chain_viterbi(const DevModel &m, const DevPlan &p, const int16_t *chain_scr,
                         int2 *tokens, int32_t *spill, int64_t spill_stride, int32_t *utt_best,
                         int32_t *utt_renorm, int32_t *fin_hist, int32_t *fin_score,
                         int max_phones, int max_band, cudaStream_t st)
{
    if (p.n_utts == 0)
        return 0;
    const int E = m.n_emit;
    if (E != 3 && E != 5) {
        // ref: src/hmm.c:741-759 also has an any-topology evaluator; the bundled models are 3-state
        set_error("chain Viterbi supports 3- and 5-state HMMs, model has %d", E);
        return -1;
    }
    // The evaluated band is usually a handful of phones (one word window); a single warp
    // keeps every step warp-synchronous.  Long chains whose band stays narrow (max_band, from the
    // planner) run in the shared-memory ring on one warp; more warps only pay for unconstrained
    // long chains, whose state then lives in shared memory or, beyond 6400 phones, in L2.
    const bool ring = max_phones > 128 && max_band > 0 && max_band <= K3_RING - 4;
    int threads = (max_phones <= 128 || ring) ? 32 : (max_phones < 4096 ? 128 : 512);
    const size_t per_phone = (size_t)(2 * E + 2) * sizeof(int32_t);
    int smem_phones = (int)((200 * 1024) / per_phone);
    if (max_phones < smem_phones)
        smem_phones = max_phones;
    // Short chains with a narrow band (ordinary sentences with word windows: ~4 phones alive):
    // 16 lanes per utterance, 8 utterances per 128-thread CTA, instead of a warp each with most of
    // its lanes idle (config #2: 1.66 -> 1.25 ms; the frame loop stays latency-bound: one dependent
    // score load per frame).  $SSB_K3_LANES forces 8 / 16 / 32.
    int lanes = 32;
    if (!ring && max_phones <= 128 && max_band > 0) {
        lanes = max_band <= 20 ? 16 : 32;  // (8 measured slower than 16 on config #2: 1.80 / 1.25 / 1.66 ms)
        if (const char *e = getenv("SSB_K3_LANES"))
            if (atoi(e) == 8 || atoi(e) == 16 || atoi(e) == 32)
                lanes = atoi(e);
    }
    const int groups = lanes < 32 ? 128 / lanes : 1;
    const size_t smem = lanes < 32 ? (size_t)k3_group_stride(smem_phones * (2 * E + 2), lanes) * 4 * groups
                                   : per_phone * (ring ? K3_RING : smem_phones);
    const unsigned grid = lanes < 32 ? (unsigned)((p.n_utts + groups - 1) / groups) : (unsigned)p.n_utts;
    if (lanes < 32)
        threads = 128;
#define SSB_K3(EE, RR, LL)                                                                           \
    do {                                                                                             \
        SSB_DYN_SMEM((chain_viterbi_kernel<EE, RR, LL>), smem);                                      \
        chain_viterbi_kernel<EE, RR, LL><<<grid, threads, smem, st>>>(                               \
            m, p, chain_scr, tokens, spill, spill_stride, utt_best, utt_renorm, fin_hist, fin_score, \
            smem_phones);                                                                            \
    } while (0)
    if (E == 3) {
        if (ring) SSB_K3(3, true, 32);
        else if (lanes == 8) SSB_K3(3, false, 8);
        else if (lanes == 16) SSB_K3(3, false, 16);
        else SSB_K3(3, false, 32);
    } else {
        if (ring) SSB_K3(5, true, 32);
        else if (lanes == 8) SSB_K3(5, false, 8);
        else if (lanes == 16) SSB_K3(5, false, 16);
        else SSB_K3(5, false, 32);
    }
#undef SSB_K3
    SSB_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// ---------------------------------------------------------------- backtrace
// ref: src/state_align_search.c:215-268.  The walk t = T-2 .. 0 is a chain of dependent 8-byte
// loads, cur = tokens[t][cur.id] -- but the id changes only ~once per state (every 6-12 frames),
// so one WARP per utterance reads the tokens of 32 consecutive frames at the current state's
// column speculatively and consumes them up to the first frame on which the stored id differs
// (or the token is missing): one memory latency per state run instead of one per frame.
// States that are not on the best path keep the sentinel duration -1 so the host leaves the
// caller's entries untouched (the reference never writes them either).
__global__ void __launch_bounds__(128)
backtrace_kernel(DevModel m, DevPlan p, const int2 *__restrict__ tokens,
                 const int32_t *__restrict__ fin_hist, const int32_t *__restrict__ fin_score,
                 int32_t *__restrict__ st_start, int32_t *__restrict__ st_dur,
                 int32_t *__restrict__ st_score, int32_t *__restrict__ utt_rv)
{
    const int lane = threadIdx.x & 31;
    const int u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (u >= p.n_utts)
        return;
    const int T = (int)(p.frame_off[u + 1] - p.frame_off[u]);
    const int64_t s0 = p.phone_off[u] * m.n_emit;
    const int ns = (int)(p.phone_off[u + 1] - p.phone_off[u]) * m.n_emit;
    const int2 *tok = tokens + p.scr_off[u];
    if (ns == 0) {
        if (lane == 0)
            utt_rv[u] = -1;
        return;
    }
    int32_t last_id = fin_hist[u], last_score = fin_score[u];
    if (last_id == -1) {
        if (lane == 0)
            utt_rv[u] = -1;  // "Failed to reach final state in alignment"
        return;
    }
    const int32_t *enter = p.enter_plan + p.phone_off[u];
    const int32_t *ef = p.ef + p.phone_off[u];
    // a segment of a cut chain: frames are reported in the real utterance's numbering
    const int t0s = p.seg_t0 ? p.seg_t0[u] : 0;
    int32_t cur_id = last_id, last_frame = T;
    int cur_frame = T - 2;
    while (cur_frame >= 0) {
        const int fr = cur_frame - lane;
        // a token exists for (frame, state) only if the phone was evaluated on that frame or
        // entered at its end; the reference reads {-1,-1} otherwise (its 0xff-filled stack)
        const int ph = cur_id / m.n_emit;
        const int32_t en = enter[ph], efp = ef[ph];
        int2 tk = make_int2(cur_id, 0);
        bool event = false;
        if (fr >= 0) {
            const bool missing = en < 0 || en > fr + 1 || (en <= fr && fr > max(en, efp));
            if (missing)
                tk.x = -1;
            else if (p.banded)
                tk = tokens[p.tok_boff[p.phone_off[u] + ph] + (int64_t)fr * m.n_emit + (cur_id - ph * m.n_emit)];
            else
                tk = tok[(int64_t)fr * ns + cur_id];
            event = tk.x != cur_id;
        }
        const unsigned ev = __ballot_sync(0xffffffffu, event);
        if (ev == 0u) {  // 32 more frames in the same state (or the beginning of the utterance)
            cur_frame -= 32;
            continue;
        }
        const int k = __ffs((int)ev) - 1;
        const int32_t nid = __shfl_sync(0xffffffffu, tk.x, k), nsc = __shfl_sync(0xffffffffu, tk.y, k);
        const int f = cur_frame - k;
        if (nid == -1) {
            if (lane == 0)
                utt_rv[u] = -1;
            return;
        }
        if (lane == 0) {
            st_start[s0 + last_id] = f + 1 + t0s;
            st_dur[s0 + last_id] = last_frame - (f + 1);
            st_score[s0 + last_id] = last_score - nsc;
        }
        last_id = cur_id = nid;
        last_score = nsc;
        last_frame = f + 1;
        cur_frame = f - 1;
    }
    if (lane == 0) {
        // first state: score left alone by the reference (ref :256-261); report 0.  The first
        // state of a later segment of a cut chain is an ordinary state boundary of the whole
        // utterance: last.score - cur.score with cur = the entry token, whose score is this
        // segment's zero.
        st_start[s0] = t0s;
        st_dur[s0] = last_frame;
        if (t0s > 0)
            st_score[s0] = last_score;
        utt_rv[u] = 0;
    }
}

int launch_backtrace(const DevModel &m, const DevPlan &p, const int2 *tokens,
                     const int32_t *fin_hist, const int32_t *fin_score, int32_t *st_start,
                     int32_t *st_dur, int32_t *st_score, int32_t *utt_rv, cudaStream_t st)
{
    if (p.n_utts == 0)
        return 0;
    backtrace_kernel<<<(p.n_utts + 3) / 4, 128, 0, st>>>(m, p, tokens, fin_hist, fin_score,
                                                         st_start, st_dur, st_score, utt_rv);
    SSB_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

// ---------------------------------------------------------------- single HMM (vtable-level test hook)
// ref: src/hmm.c:741-759 hmm_vit_eval on one non-multiplex HMM.
__global__ void hmm_eval_kernel(DevModel m, int n_emit, int tmatid,
                                const uint16_t *__restrict__ senid,
                                const int16_t *__restrict__ senscr, int32_t *st12, int32_t *best)
{
    if (threadIdx.x != 0 || blockIdx.x != 0)
        return;
    const uint8_t *tp = m.tp + (size_t)tmatid * n_emit * (n_emit + 1);
    if (n_emit == 3) {
        int32_t s[3], h[3], o_s = st12[10], o_h = st12[11];
        int ss[3];
        for (int j = 0; j < 3; ++j) {
            s[j] = st12[j];
            h[j] = st12[5 + j];
            ss[j] = senscr[senid[j]];
        }
        *best = hmm_step3(tp, ss, s, h, o_s, o_h);
        for (int j = 0; j < 3; ++j) {
            st12[j] = s[j];
            st12[5 + j] = h[j];
        }
        st12[10] = o_s;
        st12[11] = o_h;
    } else {
        int32_t s[5], h[5], o_s = st12[10], o_h = st12[11];
        int ss[5];
        for (int j = 0; j < 5; ++j) {
            s[j] = st12[j];
            h[j] = st12[5 + j];
            ss[j] = senscr[senid[j]];
        }
        *best = hmm_step5(tp, ss, s, h, o_s, o_h);
        for (int j = 0; j < 5; ++j) {
            st12[j] = s[j];
            st12[5 + j] = h[j];
        }
        st12[10] = o_s;
        st12[11] = o_h;
    }
}

// Many independent HMM steps on caller-provided transition matrices (known-answer tests of
// hmm_step3 / hmm_step5, the 5-state evaluator included: no bundled model has 5-state HMMs).
// Case i: tp[i][E][E+1], senscr[i][E] (the scores of the HMM's own states), st12[i] in place.
__global__ void hmm_eval_tp_kernel(int n_emit, int n_cases, const uint8_t *__restrict__ tp,
                                   const int16_t *__restrict__ senscr, int32_t *st12, int32_t *best)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cases)
        return;
    const uint8_t *t = tp + (size_t)i * n_emit * (n_emit + 1);
    int32_t *st = st12 + (size_t)i * 12;
    if (n_emit == 3) {
        int32_t s[3], h[3], o_s = st[10], o_h = st[11];
        int ss[3];
        for (int j = 0; j < 3; ++j) {
            s[j] = st[j];
            h[j] = st[5 + j];
            ss[j] = senscr[(size_t)i * 3 + j];
        }
        best[i] = hmm_step3(t, ss, s, h, o_s, o_h);
        for (int j = 0; j < 3; ++j) {
            st[j] = s[j];
            st[5 + j] = h[j];
        }
        st[10] = o_s;
        st[11] = o_h;
    } else {
        int32_t s[5], h[5], o_s = st[10], o_h = st[11];
        int ss[5];
        for (int j = 0; j < 5; ++j) {
            s[j] = st[j];
            h[j] = st[5 + j];
            ss[j] = senscr[(size_t)i * 5 + j];
        }
        best[i] = hmm_step5(t, ss, s, h, o_s, o_h);
        for (int j = 0; j < 5; ++j) {
            st[j] = s[j];
            st[5 + j] = h[j];
        }
        st[10] = o_s;
        st[11] = o_h;
    }
}

int launch_hmm_eval_tp(int n_emit, int n_cases, const uint8_t *tp, const int16_t *senscr, int32_t *st12,
                       int32_t *best, cudaStream_t st)
{
    if (n_emit != 3 && n_emit != 5) {
        set_error("hmm_vit_eval: %d emitting states not supported (3 or 5)", n_emit);
        return -1;
    }
    if (n_cases <= 0)
        return 0;
    hmm_eval_tp_kernel<<<(n_cases + 127) / 128, 128, 0, st>>>(n_emit, n_cases, tp, senscr, st12, best);
    SSB_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

int launch_hmm_eval(const DevModel &m, int n_emit, int tmatid, const uint16_t *senid,
                    const int16_t *senscr, int32_t *st12, int32_t *best, cudaStream_t st)
{
    if (n_emit != 3 && n_emit != 5) {
        set_error("hmm_vit_eval: %d emitting states not supported (3 or 5)", n_emit);
        return -1;
    }
    hmm_eval_kernel<<<1, 32, 0, st>>>(m, n_emit, tmatid, senid, senscr, st12, best);
    SSB_CUDA(cudaGetLastError());
    note_launch();
    return 0;
}

}  // namespace ssb
