"""BASELINE configs #3, #4, #5 and the compallsen rate, as functions bench.py calls inside its
one run (same process, same clock sampler), each returning a dict for the JSON line:

  config3           goforward.gram decode (fsg_search), 4096 x 278 frames, en-us, active lists
  config4           fr-fr long-form: ONE utterance of an hour, windows given / from the transcript
  config5_two_pass  65 536 utterances, audio + transcript -> JSON (frontend, both passes), split
                    utt % n_gpu over the ranks: STRONG scaling
  compallsen        config #2 with every senone scored on every frame (SURVEY 8d's 2.1e10 unit)

`cpu_*` functions time the reference's own implementation of the same operation on a bounded
sample (oracle/_ref through oracle/refshim.py -- bench.py's CPU legs are the one place outside
tests/ that may execute oracle/)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "tests", "data")
MODELS = os.path.join(ROOT, "soundswallower_b200", "model")
JSGF = os.path.join(DATA, "goforward.gram")


def _graph_from_golden(g, name):
    keys = ("n_state", "start", "final", "n_ciphone", "sil", "beam", "pbeam", "wbeam", "maxhmmpf",
            "link", "link_flag", "arc_off", "root", "pnode", "ctxt")
    return {k: g["%s_%s" % (name, k)] for k in keys}


# ------------------------------------------------------------------ config #3
def config3(ssb, model, utts=4096, steps=2, peaks=None):
    g = np.load(os.path.join(GOLD, "fsg_en-us.npz"))
    feat = np.load(os.path.join(GOLD, "align_en-us.npz"))["feat"]
    T = feat.shape[0]
    rng = np.random.Generator(np.random.Philox(1234))
    feats = feat[None] + rng.standard_normal((utts,) + feat.shape, dtype=np.float32) * np.float32(0.05)
    feats = [f for f in feats]
    graph = _graph_from_golden(g, "jsgf")
    active = bool(model.fsg_active_ok)
    ssb.fsg_batch(model, feats[:64], [graph], compallsen=not active)
    best = None
    for _ in range(steps):
        t0 = time.perf_counter()
        res = ssb.fsg_batch(model, feats, [graph], max_seg=32, compallsen=not active)
        wall = time.perf_counter() - t0
        if best is None or wall < best[0]:
            best = (wall, res[0]["kernel_ms"], res)
    wall, ms, res = best
    dev_ms = float(sum(ms.values()))
    audio_s = utts * T / 100.0
    out = {"workload": "config#3: goforward.gram (JSGF) decode, %d x %d frames, en-us, %s"
                       % (utts, T, "active lists (reference default)" if active else "dense scores"),
           "ms": dev_ms, "kernel_ms": ms, "audio_s_per_s": audio_s / (dev_ms * 1e-3),
           "e2e_ms": wall * 1e3, "audio_s_per_s_e2e": audio_s / wall,
           "decoded": int(sum(1 for r in res if r["exit"] > 0 and r["rv"] == 0)),
           "hmm_evals_per_frame": float(np.mean([r["n_hmm_eval"] for r in res])) / T}
    if active:
        out["senones_per_frame"] = float(np.mean([r["n_sen_eval"] for r in res])) / T
    # K1 scans every codebook of every frame here: tensor roofline by SURVEY 8d's algorithmic FLOPs
    flops = utts * T * model.n_mgau * model.n_feat * model.n_density * 2 * (2 * model.veclen + 1)
    peak = float((peaks or {}).get("bf16_tflops_sustained", 1400.0))
    ach = flops / (ms["gmm_topn"] * 1e-3) / 1e12
    out["roofline"] = {"kernel": "K1 (all codebooks)", "bound": "tensor", "achieved": ach, "peak": peak,
                       "unit": "TFLOP/s", "frac": ach / peak}
    return out


def cpu_config3(n_utts):
    """The reference's fsg_search on goforward.gram, one worker; returns seconds per utterance."""
    from oracle.refshim import Ref
    feat = np.load(os.path.join(GOLD, "align_en-us.npz"))["feat"]
    r = Ref(os.path.join(MODELS, "en-us"))
    jsgf = open(JSGF).read()
    r.fsg_decode(feat, jsgf=jsgf)
    t0 = time.perf_counter()
    for u in range(n_utts):
        rng = np.random.Generator(np.random.Philox(1234 + u))
        r.fsg_decode(feat + rng.standard_normal(feat.shape, dtype=np.float32) * np.float32(0.05), jsgf=jsgf)
    return (time.perf_counter() - t0) / n_utts, feat.shape[0] / 100.0


# ------------------------------------------------------------------ config #4
def _longform(reps, frames_per_rep):
    from bench_longform import build
    g = np.load(os.path.join(GOLD, "align_fr-fr.npz"))
    return build(g, reps, frames_per_rep)


def config4(ssb, reps=714, frames_per_rep=504, from_text=True, peaks=None):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    hmm = os.path.join(MODELS, "fr-fr")
    x, chain = _longform(reps, frames_per_rep)
    m = ssb.AcousticModel(hmm)
    b = ssb.StateAlignBatch(m)
    b.upload([x], [chain]); b.run(); b.download()          # warm-up (allocations, module load)
    t0 = time.perf_counter()
    b.upload([x], [chain])
    b.run()
    res = b.per_utt(b.download())[0]
    wall = time.perf_counter() - t0
    ms = b.kernel_ms()
    st = b.stats()
    on = res["dur"] > 0
    start, dur = res["start"][on], res["dur"][on]
    sf, ef = np.repeat(chain["sf"], 3)[on], np.repeat(chain["ef"], 3)[on]
    ok = bool(res["rv"] == 0 and start[0] == 0 and (start[1:] == start[:-1] + dur[:-1]).all()
              and start[-1] + dur[-1] == x.shape[0] and (start >= sf).all() and (start + dur <= ef).all())
    audio_s = x.shape[0] / 100.0
    hbm = float((peaks or {}).get("hbm_gbs", 6650.0))
    k3 = ms["chain_viterbi"] * 1e-3
    out = {"workload": "config#4: fr-fr long-form, 1 utterance, %d frames (%.0f min), %d phones / %d states, "
                       "word windows given" % (x.shape[0], x.shape[0] / 6000.0, len(chain["ssid"]),
                                               3 * len(chain["ssid"])),
           "ms": ms["total"], "kernel_ms": ms, "audio_s_per_s": audio_s / (ms["total"] * 1e-3),
           "e2e_ms": wall * 1e3, "audio_s_per_s_e2e": audio_s / wall, "invariants_ok": ok,
           "device_bytes": st["device_bytes"], "state_frames_dense": st["state_frames"],
           "band_state_frames": st["band_state_frames"], "chain_segments": st["segments"],
           "plan_us": st["plan_us"],
           "roofline": {"kernel": "chain_viterbi_kernel (K3)", "bound": "hbm", "unit": "GB/s", "peak": hbm,
                        "achieved": st["band_state_frames"] * 10 / k3 / 1e9,
                        "frac": st["band_state_frames"] * 10 / k3 / 1e9 / hbm,
                        "reference_equivalent_dense_10B_frac": st["state_frames"] * 10 / k3 / 1e9 / hbm,
                        "note": "10 B per EVALUATED state-frame (the word-window band); the dense T x states "
                                "figure is what the reference's token stack would hold"}}
    b.close()
    if from_text:
        lx = ssb.Lexicon(m, hmmdir=hmm)
        text = " ".join(["avance de dix mètres"] * reps)
        t0 = time.perf_counter()
        ta = ssb.TextAlignment(m, lx, [x], [text], align_level=1)
        t1 = time.perf_counter()
        rv, hyp, nfr = ta.status(0)
        wd = ta.entries(0, "words")
        j = ta.json(0, align_level=1)
        t2 = time.perf_counter()
        okt = bool(rv == 0 and wd[0, 1] == 0 and (wd[1:, 1] == wd[:-1, 1] + wd[:-1, 2]).all()
                   and wd[-1, 1] + wd[-1, 2] == len(x))
        real = [w for w in json.loads(j)["w"] if not w["t"].startswith("<")]
        out["from_text"] = {"workload": "the same hour from its %d-word transcript: grammar, first pass "
                                        "(default mode), chains, second pass, JSON" % (4 * reps),
                            "ms": (t2 - t0) * 1e3, "audio_s_per_s": audio_s / (t2 - t0), "invariants_ok": okt,
                            "words_aligned": len(real), "pass1_kernel_ms": ta.kernel_ms()}
        ta.close()
        lx.close()
    m.close()
    return out


def cpu_config4(minutes=2.0):
    """The reference's second pass on a prefix of the long-form utterance; seconds per audio-second."""
    from oracle.refshim import Ref
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    reps = max(1, int(minutes * 6000 / 504))
    g = np.load(os.path.join(GOLD, "align_fr-fr.npz"))
    from bench_longform import build
    x, _chain = build(g, reps, 504)
    words = g["words"]
    wid, ws, wd = [], [], []
    P = 504
    for k in range(reps):
        for w in range(len(words) - 1):
            s, d = int(words[w, 1]) + k * P, int(words[w, 2])
            if w == 0 and k > 0:
                s = int(words[-1, 1]) + (k - 1) * P
                d = k * P + int(words[0, 2]) - s
            wid.append(int(words[w, 0])); ws.append(s); wd.append(d)
    s = int(words[-1, 1]) + (reps - 1) * P
    wid.append(int(words[-1, 0])); ws.append(s); wd.append(reps * P - s)
    r = Ref(os.path.join(MODELS, "fr-fr"))
    t0 = time.perf_counter()
    res = r.state_align(x, np.array(wid, np.int32), np.array(ws, np.int32), np.array(wd, np.int32))
    dt = time.perf_counter() - t0
    return dt, x.shape[0] / 100.0, res["rv"] == 0


# ------------------------------------------------------------------ config #5 (two passes, audio in)
def two_pass_pool(utts, pinned=None):
    """`utts` distinct 10 s utterances back to back: goforward.raw x3 + trailing silence + noise."""
    pcm = np.frombuffer(open(os.path.join(DATA, "goforward.raw"), "rb").read(), np.int16)
    n = 160000
    sil = pcm[33760:]
    base = np.concatenate([pcm] * 3)
    base = np.concatenate([base, np.tile(sil, (n - len(base)) // len(sil) + 1)])[:n].astype(np.int16)
    flat = pinned if pinned is not None else np.empty(utts * n, np.int16)
    rng = np.random.Generator(np.random.Philox(99))
    blk = 256
    for u0 in range(0, utts, blk):
        k = min(blk, utts - u0)
        noise = rng.integers(-30, 31, (k, n), dtype=np.int16)
        np.add(base[None], noise, out=flat[u0 * n:(u0 + k) * n].reshape(k, n))
    return flat, n


def config5_two_pass(ssb, model, total_utts=65536, rank=0, world=1, chunk=4096, pinned=None, max_chunks=None):
    """This rank's share (utt % world == rank) of `total_utts` utterances, in chunks of `chunk`:
    every chunk = H2D of its audio, frontend, ssb_align_texts (both passes), JSON for every
    utterance.  The host keeps ONE pool of `chunk` distinct utterances and sends it again for
    every chunk (generating 21 GB of distinct audio would dominate the run)."""
    hmm = os.path.join(MODELS, "en-us")
    mine = (total_utts - rank + world - 1) // world
    n_chunks = (mine + chunk - 1) // chunk
    if max_chunks:
        n_chunks = min(n_chunks, max_chunks)
    flat, n = two_pass_pool(chunk, pinned)
    text = " ".join(["go forward ten meters"] * 3)
    n_workers = max(1, int(os.environ.get("SSB_TWO_PASS_WORKERS", "2")))

    class Worker:
        """One host thread's objects: while its chunk is in host code (grammars, chains, JSON) the
        other thread's chunk is on the GPU."""

        def __init__(self):
            self.lx = ssb.Lexicon(model, hmmdir=hmm)
            # (several workers: each frontend on a stream of its own, so that one worker's audio
            # copy and kernels do not queue behind the other's on the shared default stream)
            self.stream = None
            if n_workers > 1:
                import torch
                self.stream = torch.cuda.Stream(device=model.device)
            self.fe = ssb.Frontend(hmm, device=model.device,
                                   stream=self.stream.cuda_stream if self.stream else None)

        def one(self, k_utts):
            off = np.arange(k_utts + 1, dtype=np.int64) * n
            feats = self.fe.run_raw(flat[:k_utts * n], off)
            ta = ssb.TextAlignment(model, self.lx, feats, [text] * k_utts, align_level=1)
            ta.render(align_level=1)
            js0 = ta.json(0, align_level=1)
            ok = sum(1 for u in range(k_utts) if ta.status(u)[0] == 0)
            ms1 = ta.kernel_ms()
            ta.close()
            return ok, js0, ms1

        def close(self):
            self.fe.close()
            self.lx.close()

    workers = [Worker() for _ in range(n_workers)]
    for w in workers:
        w.one(min(chunk, 256))  # warm-up
    sizes, done = [], 0
    for c in range(n_chunks):
        k = min(chunk, mine - done)
        sizes.append(k)
        done += k
    import concurrent.futures as cf
    import threading
    local = threading.local()
    free = list(workers)
    lock = threading.Lock()

    def run(k):
        with lock:
            w = free.pop()
        try:
            return w.one(k)
        finally:
            with lock:
                free.append(w)

    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(n_workers) as ex:
        res = list(ex.map(run, sizes))
    wall = time.perf_counter() - t0
    for w in workers:
        w.close()
    ok_total = sum(r[0] for r in res)
    js0, ms1 = res[-1][1], res[-1][2]
    return {"utts": done, "aligned": ok_total, "wall_s": wall, "audio_s": done * n / 16000.0,
            "first_json_words": len(json.loads(js0)["w"]) if js0 else 0, "pass1_kernel_ms_last_chunk": ms1,
            "h2d_bytes": int(done * n * 2), "host_threads": n_workers}


def cpu_two_pass(n_utts):
    """The reference's own two-pass alignment (decoder_process_int16(full_utt) + decoder_alignment,
    through ref.align_pcm) on utterances of the same workload; seconds per utterance."""
    from oracle.refshim import Ref
    flat, n = two_pass_pool(max(n_utts, 1))
    r = Ref(os.path.join(MODELS, "en-us"))
    text = " ".join(["go forward ten meters"] * 3)
    r.align_pcm(flat[:n], text)
    t0 = time.perf_counter()
    for u in range(n_utts):
        r.align_pcm(flat[u * n:(u + 1) * n], text)
    return (time.perf_counter() - t0) / max(n_utts, 1), n / 16000.0


# ------------------------------------------------------------------ CPU samples on all cores
def _cpu_task(task):
    name, n = task
    if name == "config3":
        sec, audio = cpu_config3(n)
        return sec, audio
    if name == "two_pass":
        sec, audio = cpu_two_pass(n)
        return sec, audio
    raise ValueError(name)


def cpu_rate(name, cores, n_per_worker):
    """audio-s/s of the reference on `cores` worker processes (each its own decoder)."""
    import multiprocessing as mp
    with mp.get_context("fork").Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_cpu_task, [(name, n_per_worker)] * cores)
        wall = time.perf_counter() - t0
    sec = float(np.mean([r[0] for r in res]))
    audio = res[0][1]
    return {"value": cores * audio / sec, "unit": "audio-s/s", "cores": cores, "kind": "reference",
            "sample": "%d utterances per worker x %d workers (%.1f s wall incl. model load)"
                      % (n_per_worker, cores, wall)}
