#!/usr/bin/env python
"""bench.py -- forced-alignment throughput of the B200 path (BASELINE.json metric:
"forced-align audio-sec/sec", quoted on config #2: en-us senone scoring +
state_align_search on 4096 synthetic 10 s utterances per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
  python bench.py --impl reference [--gpus N] --steps K ...      the reference's CPU path

One process per GPU (torchrun for N > 1); utterances are sharded across ranks with no
collective on the data path (weak scaling: every rank aligns its own 4096 utterances).
A step = one pass of the hot path over one batch:
  * `value`: features resident in HBM, CUDA-event time of K1..backtrace on the launch stream;
  * `e2e`  : the same batch through the C ABI with HOST buffers (pinned features in,
             segmentations out), plan + H2D + kernels + D2H inside the timed region.
Inputs (639 MB of features per rank) are larger than L2 and every step rewrites > 15 GB of
intermediates, so no explicit L2 flush is needed between timed steps.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "align_en-us.npz")
MODEL = os.path.join(ROOT, "soundswallower_b200", "model", "en-us")
FRAMES = 1000          # 10 s utterances
FRAME_RATE = 100.0     # frames per audio second (ref: config_defs.h "frate")


# ------------------------------------------------------------------ workload
def config2_template(g, frames=FRAMES):
    """goforward features tiled 3x and padded with their trailing silence to `frames` frames;
    chain = ("<sil> go forward ten meters") x3 + "<sil>": 16 words / 52 phones / 156 states,
    word windows from the reference's first pass on the single sentence."""
    import soundswallower_b200 as ssb
    feat = g["feat"]                       # [278][39]
    T1 = feat.shape[0]
    words, phones = g["words"], g["phones"]
    sil_tail = feat[int(words[-1, 1]):]    # trailing silence frames 211..277
    reps = 3
    base = np.concatenate([feat] * reps)
    pad = frames - base.shape[0]
    assert pad >= 0
    tail = np.concatenate([sil_tail] * (pad // sil_tail.shape[0] + 1))[:pad]
    base = np.concatenate([base, tail]).astype(np.float32)
    n_ph_sentence = len(phones) - 1        # everything but the final silence
    ssid, tmat, wstart, wdur = [], [], [], []
    for k in range(reps):
        for i in range(n_ph_sentence):
            w = int(phones[i, 6])
            s, d = int(words[w, 1]) + k * T1, int(words[w, 2])
            if w == 0 and k > 0:           # leading silence merges with the previous trailing one
                s = int(words[-1, 1]) + (k - 1) * T1
                d = k * T1 + int(words[0, 2]) - s
            ssid.append(int(phones[i, 1]))
            tmat.append(int(phones[i, 2]))
            wstart.append(s)
            wdur.append(d)
    s = int(words[-1, 1]) + (reps - 1) * T1
    ssid.append(int(phones[-1, 1]))
    tmat.append(int(phones[-1, 2]))
    wstart.append(s)
    wdur.append(frames - s)
    sf, ef = ssb.windows(np.array(wstart, np.int32), np.array(wdur, np.int32))
    chain = dict(ssid=np.array(ssid, np.int32), tmat=np.array(tmat, np.int32), sf=sf, ef=ef)
    return base, chain


def make_config2_batch(g, n_utts, noise=0.05, seed=1234, frames=FRAMES, out=None, ids=None):
    """Per-utterance additive N(0, noise^2) on the tiled features, Philox seed = seed + utt
    (utt = ids[u], the utterance's index in the global list, when a shard is generated)."""
    base, chain = config2_template(g, frames)
    feats = []
    for u in range(n_utts):
        rng = np.random.Generator(np.random.Philox(seed + (int(ids[u]) if ids is not None else u)))
        x = out[u] if out is not None else np.empty_like(base)
        np.add(base, rng.standard_normal(base.shape, dtype=np.float32) * np.float32(noise), out=x)
        feats.append(x)
    return feats, [chain] * n_utts


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock + clock-event (throttle) reasons sampled every ~20 ms while the timed region runs,
    through NVML (nvidia-ml-py); falls back to polling nvidia-smi when NVML is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.sm, self.reasons, self.sm_max = [], set(), None
        self.stop = threading.Event()
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20,
                "hw_thermal_slowdown": 0x40}
        while not self.stop.is_set():
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            r = int(get_reasons(h))
            for name, bit in bits.items():
                if r & bit:
                    self.reasons.add(name)
            self.stop.wait(0.02)
        nv.nvmlShutdown()

    def _run_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                r = [c.strip() for c in out.split(",")]
                if len(r) >= 7:
                    self.sm.append(float(r[0]))
                    self.sm_max = float(r[1])
                    for i in range(4):
                        if r[3 + i].lower().startswith("active"):
                            self.reasons.add(names[i])
            except Exception:
                pass
            self.stop.wait(0.05)

    def _run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def __enter__(self):
        self.th.start()
        time.sleep(0.03)
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ------------------------------------------------------------------ CPU legs (reference / oracle)
_W = {}


def _cpu_init(kind):
    if kind == "reference":
        from oracle.refshim import Ref
        _W["ref"] = Ref(MODEL)
    else:
        from oracle.oracle import Oracle
        _W["orc"] = Oracle(MODEL)
    g = np.load(GOLDEN)
    _W["g"] = g
    _W["base"], _W["chain"] = config2_template(g)
    words = g["words"]
    T1 = g["feat"].shape[0]
    wid, ws, wd = [], [], []
    for k in range(3):
        for w in range(len(words) - 1):
            s, d = int(words[w, 1]) + k * T1, int(words[w, 2])
            if w == 0 and k > 0:
                s = int(words[-1, 1]) + (k - 1) * T1
                d = k * T1 + int(words[0, 2]) - s
            wid.append(int(words[w, 0]))
            ws.append(s)
            wd.append(d)
    s = int(words[-1, 1]) + 2 * T1
    wid.append(int(words[-1, 0]))
    ws.append(s)
    wd.append(FRAMES - s)
    _W["words"] = (np.array(wid, np.int32), np.array(ws, np.int32), np.array(wd, np.int32))


def _cpu_align(u):
    """Second pass of the reference (re-score + chain Viterbi) on utterance u of the workload."""
    rng = np.random.Generator(np.random.Philox(1234 + u))
    base = _W["base"]
    x = base + rng.standard_normal(base.shape, dtype=np.float32) * np.float32(0.05)
    t0 = time.perf_counter()
    if "ref" in _W:
        wid, ws, wd = _W["words"]
        r = _W["ref"].state_align(x, wid, ws, wd, clear_active=True)
        ok = r["rv"] == 0
        fp = int(r["states"][:, 2].astype(np.int64).sum())
    else:
        c = _W["chain"]
        r = _W["orc"].state_align(x, c["ssid"], c["tmat"], c["sf"], c["ef"])
        ok = r["rv"] == 0
        fp = int(r["dur"].astype(np.int64).sum())
    return time.perf_counter() - t0, ok, fp


def cpu_kind():
    from oracle import refshim
    return "reference" if refshim.available() else "port"


class CpuPool:
    """`cores` worker processes, each with its own reference decoder (the reference is
    single-threaded and not re-entrant)."""

    def __init__(self, cores, kind):
        import multiprocessing as mp
        self.cores = cores
        self.pool = mp.get_context("fork").Pool(cores, initializer=_cpu_init, initargs=(kind,))
        self.pool.map(_cpu_align, range(cores))  # warm the workers (model load, page-in)

    def run(self, n_utts, first=0):
        """Align `n_utts` utterances; returns (audio-s/s, all valid, wall seconds)."""
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_align, range(first, first + n_utts), chunksize=1)
        wall = time.perf_counter() - t0
        ok = all(r[1] for r in res) and all(r[2] == FRAMES for r in res)
        return n_utts * FRAMES / FRAME_RATE / wall, ok, wall

    def close(self):
        self.pool.close()
        self.pool.join()


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--utts", type=int, default=4096, help="utterances per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--compallsen", action="store_true", help="score every senone every frame")
    ap.add_argument("--no-extra", action="store_true",
                    help="config #2 only (skip configs #3, #4, #5 and the compallsen rate)")
    ap.add_argument("--total-utts", type=int, default=65536, help="config #5: utterances over all GPUs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    workload = ("config#2(i): en-us, %d x 10 s utterances per GPU, goforward features tiled 3x "
                "+ N(0,0.05^2), 16 words / 52 phones / 156 states, word windows, "
                "senone scoring (%s) + state_align_search" %
                (args.utts, "compallsen" if args.compallsen else "active lists"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        cores = host_cores()
        kind = cpu_kind()
        per_step = cores * 16
        pool = CpuPool(cores, kind)
        for _ in range(args.warmup):
            pool.run(cores)
        vals, oks = [], True
        t_all = time.perf_counter()
        for s in range(args.steps):
            v, ok, _ = pool.run(per_step, first=s * per_step)
            vals.append(v)
            oks = oks and ok
        wall = time.perf_counter() - t_all
        pool.close()
        v = float(np.mean(vals))
        line = {"impl": "reference", "metric": "forced-align audio-sec/sec", "value": v,
                "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+i32",
                "data": "synthetic", "config": {"workload": workload, "sample_utts_per_step": per_step},
                "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": cores, "kind": kind,
                                 "sample": "%d utterances (%.0f audio-s) per step, second pass "
                                           "(re-score + state_align_search), all results valid=%s"
                                           % (per_step, per_step * FRAMES / FRAME_RATE, oks)},
                "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import soundswallower_b200 as ssb
    from soundswallower_b200 import _build
    _build.build_lib()
    if not torch.cuda.is_available() or ssb.device_count() == 0:
        raise SystemExit("bench.py needs a B200: the product has no CPU path")
    torch.cuda.set_device(local)
    # One rank per GPU on a two-socket host: keep the rank's threads -- and with them the pinned
    # staging buffers it allocates (first touch) -- on the CPUs next to its GPU, so that eight
    # ranks' H2D copies do not cross the socket interconnect.  ($SSB_BENCH_NUMA=0: leave it alone.)
    numa_note = None
    if world > 1 and os.environ.get("SSB_BENCH_NUMA", "1") != "0" and hasattr(os, "sched_setaffinity"):
        try:
            import pynvml
            pynvml.nvmlInit()
            ncpu = os.cpu_count() or 1
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
            cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
            cpus &= set(os.sched_getaffinity(0))
            if cpus and len(cpus) < len(os.sched_getaffinity(0)):
                os.sched_setaffinity(0, cpus)
                numa_note = "rank pinned to the %d CPUs local to its GPU" % len(cpus)
        except Exception as e:  # noqa: BLE001 -- a box without NVML affinity data runs unpinned
            numa_note = "not pinned (%s)" % type(e).__name__
    dist = None
    if world > 1:
        import torch.distributed as dist
        # keep stdout to the one JSON line: NCCL prints its version banner there otherwise
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    g = np.load(GOLDEN)
    model = ssb.AcousticModel(MODEL, device=local)
    U = args.utts
    # this rank's shard of the global list of world*U utterances: utt % n_gpu == rank
    from soundswallower_b200 import shard
    my_ids = shard.shard_indices(U * world, rank, world)
    pinned = torch.empty((U, FRAMES, model.blk), dtype=torch.float32, pin_memory=True)
    feat_np = pinned.numpy()
    feats, chains = make_config2_batch(g, U, seed=1234, out=feat_np, ids=my_ids)
    chain = chains[0]
    frame_off = np.arange(U + 1, dtype=np.int64) * FRAMES
    phone_off = np.arange(U + 1, dtype=np.int64) * len(chain["ssid"])
    flat = {k: np.tile(chain[k], U) for k in ("ssid", "tmat", "sf", "ef")}
    batch = ssb.StateAlignBatch(model)

    def upload():
        batch.upload_raw(feat_np.reshape(-1, model.blk), frame_off, phone_off, flat["ssid"],
                         flat["tmat"], flat["sf"], flat["ef"], None, args.compallsen)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    upload()
    for _ in range(args.warmup):
        batch.run()
    barrier()
    kms = {k: 0.0 for k in ("gmm_topn", "senone_mix", "chain_viterbi", "backtrace", "total")}
    launches = 0
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            batch.run()
            launches += batch.n_launches()
            ms = batch.kernel_ms()  # CUDA events on the launch stream; synchronises
            for k in kms:
                kms[k] += ms[k]
        barrier()
        wall_dev = time.perf_counter() - t0
        clocks = clk.summary()
    dev_ms = kms["total"] / args.steps
    stats = batch.stats()
    res = batch.download()
    n_fail = int((res["rv"] != 0).sum())
    h2d = feat_np.nbytes + sum(a.nbytes for a in flat.values()) + frame_off.nbytes + phone_off.nbytes
    d2h = 3 * res["start"].nbytes + 3 * res["rv"].nbytes

    # end to end through the public call, host buffers in and out: (a) one resident batch,
    # upload -> run -> download back to back; (b) the pipeline (ssb_pipeline_align: chunks of
    # whole utterances through 4 lanes, copies and planning overlap the kernels) -- the headline
    upload(); batch.run(); res_one = batch.download()
    barrier()
    t0 = time.perf_counter()
    host_split = [0.0, 0.0, 0.0]  # seconds inside upload (plan + staging), run (launches), download
    for _ in range(args.steps):
        ta = time.perf_counter()
        upload()
        tb = time.perf_counter()
        batch.run()
        tc = time.perf_counter()
        res_one = batch.download()
        td = time.perf_counter()
        host_split[0] += tb - ta
        host_split[1] += tc - tb
        host_split[2] += td - tc
    barrier()
    e2e_one_s = (time.perf_counter() - t0) / args.steps
    batch.close()

    pipe = ssb.AlignPipeline(model)

    def pipe_args(pp):
        pp.upload_raw(feat_np.reshape(-1, model.blk), frame_off, phone_off, flat["ssid"],
                      flat["tmat"], flat["sf"], flat["ef"], None, args.compallsen)

    def pipe_once():
        pipe_args(pipe)
        return pipe.align()

    for _ in range(2):
        res = pipe_once()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = pipe_once()
    barrier()
    e2e_call_s = (time.perf_counter() - t0) / args.steps
    pipe_chunks, pipe_launches = pipe.n_chunks(), pipe.n_launches()
    if os.environ.get("SSB_PIPE_TRACE") and rank == 0:
        for row in pipe.trace():
            print("chunk lane=%d utt0=%d upload %.1f..%.1f done %.1f  K1 %.1f ms kernels %.1f ms"
                  % tuple(row[:7]), file=sys.stderr)
    pipe_same = all(np.array_equal(res[k], res_one[k]) for k in ("start", "dur", "score", "rv", "best_score"))
    pipe.close()

    # (c) a stream of batches, the headline: every step is one whole batch submitted through
    # ssb_pipeline_submit / collected through ssb_pipeline_collect, two in flight -- batch i+1 is
    # planned and copied in while batch i computes (kernels of different batches never share the
    # GPU); every step's H2D and D2H happen inside the timed region
    stream = ssb.AlignPipeline(model, n_lanes=2, chunk_frames=1 << 40, overlap_kernels=False)

    def stream_run(n):
        tickets, out = [], None
        for _ in range(n):
            pipe_args(stream)
            tickets.append(stream.submit())
            if len(tickets) > 1:
                out = stream.collect(tickets.pop(0))
        while tickets:
            out = stream.collect(tickets.pop(0))
        return out

    stream_run(3)
    barrier()
    t0 = time.perf_counter()
    res = stream_run(args.steps)
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    stream_launches = stream.n_launches()
    stream_same = all(np.array_equal(res[k], res_one[k]) for k in ("start", "dur", "score", "rv", "best_score"))
    stream.close()

    # ---- the other BASELINE configs, same run, same clock sampler (tools/bench_configs.py)
    feat_nbytes = feat_np.nbytes
    extra = {}
    c5 = None
    if not args.no_extra and not args.compallsen:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_configs as bc
        peaks_x = {}
        try:
            peaks_x = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        feat_nbytes = feat_np.nbytes
        del pinned, feat_np, feats
        with ClockSampler(local) as clk_x:
            # config #5: 65 536 utterances, audio + text -> JSON, utt % n_gpu: STRONG scaling
            pin5 = torch.empty(4096 * 160000, dtype=torch.int16, pin_memory=True).numpy()
            barrier()
            c5 = bc.config5_two_pass(ssb, model, total_utts=args.total_utts, rank=rank, world=world,
                                     chunk=4096, pinned=pin5)
            barrier()
            del pin5
            if world == 1:
                # compallsen: every senone of every frame (2.1e10 senone scores)
                Uc = min(U, 1024)
                fc = np.empty((Uc, FRAMES, model.blk), np.float32)
                make_config2_batch(g, Uc, seed=1234, out=fc)
                bcall = ssb.StateAlignBatch(model)
                bcall.upload_raw(fc.reshape(-1, model.blk), frame_off[:Uc + 1], phone_off[:Uc + 1],
                                 flat["ssid"][:Uc * len(chain["ssid"])], flat["tmat"][:Uc * len(chain["ssid"])],
                                 flat["sf"][:Uc * len(chain["ssid"])], flat["ef"][:Uc * len(chain["ssid"])],
                                 None, True)
                bcall.run()
                bcall.run()
                msc = bcall.kernel_ms()
                stc = bcall.stats()
                score_ms = msc["gmm_topn"] + msc["senone_mix"]
                extra["compallsen"] = {
                    "workload": "config#2 in compallsen mode: %d x 10 s, all %d senones of every frame "
                                "(dense int16 scores in slabs of 32768 frames) + gather + state_align"
                                % (Uc, model.n_sen),
                    "kernel_ms": msc, "senone_scores": stc["active_senone_frames"],
                    "senone_scores_per_s": stc["active_senone_frames"] / (score_ms * 1e-3),
                    "audio_s_per_s": Uc * FRAMES / FRAME_RATE / (msc["total"] * 1e-3),
                    "roofline": {"bound": "tensor", "unit": "TFLOP/s",
                                 "achieved": stc["scanned_cb_frames"] * model.n_feat * model.n_density * 2
                                             * (2 * model.veclen + 1) / (msc["gmm_topn"] * 1e-3) / 1e12,
                                 "peak": float(peaks_x.get("bf16_tflops_sustained", 1400.0))}}
                extra["compallsen"]["roofline"]["frac"] = (extra["compallsen"]["roofline"]["achieved"]
                                                           / extra["compallsen"]["roofline"]["peak"])
                bcall.close()
                del fc
                extra["config3"] = bc.config3(ssb, model, utts=4096, steps=2, peaks=peaks_x)
                extra["config4"] = bc.config4(ssb, peaks=peaks_x)
            extra["clocks"] = clk_x.summary()

    audio_s = U * FRAMES / FRAME_RATE
    t_dev = torch.tensor([dev_ms, e2e_s * 1e3, c5["wall_s"] if c5 else 0.0], dtype=torch.float64, device="cuda")
    # the only exchange of the job: result summaries (failures, a checksum of all segmentations)
    chk = torch.tensor([n_fail, int(res["start"].astype(np.int64).sum() % (1 << 40)),
                        int(res["dur"].astype(np.int64).sum()), c5["utts"] if c5 else 0,
                        c5["aligned"] if c5 else 0, c5["h2d_bytes"] if c5 else 0],
                       dtype=torch.int64, device="cuda")
    if dist is not None:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        dist.all_reduce(chk, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_ms_max, c5_wall_max = float(t_dev[0]), float(t_dev[1]), float(t_dev[2])
    if c5:
        tot_utts, tot_ok, tot_h2d = int(chk[3]), int(chk[4]), int(chk[5])
        extra["config5_two_pass"] = {
            "workload": "config#5: %d utterances of 10 s (16 kHz int16 audio + 12-word transcript) -> the "
                        "reference CLI's JSON for every utterance: frontend, alignment grammar + first "
                        "pass (default mode), chains, second pass, decoder_result_json; utterances "
                        "split utt %% n_gpu, chunks of 4096 per rank (one pool of 4096 distinct noisy "
                        "utterances sent again for every chunk)" % tot_utts,
            "scaling": "strong", "n_gpus": world, "utts": tot_utts, "aligned": tot_ok,
            "wall_s": c5_wall_max, "ms": c5_wall_max * 1e3,
            "audio_s_per_s": tot_utts * 10.0 / c5_wall_max if c5_wall_max > 0 else None,
            "h2d_bytes": tot_h2d, "timing": "host wall clock around this rank's chunks (H2D, kernels, host "
                                            "graph/chain/JSON work, D2H), barrier on both sides, max over ranks",
            "rank0_last_chunk_pass1_kernel_ms": c5["pass1_kernel_ms_last_chunk"]}
    n_fail = int(chk[0])
    frames_covered = int(chk[2])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # DRAM traffic of the dominant kernel per launch: dram__bytes_read.sum + dram__bytes_write.sum of
    # one `ncu --set full` capture of this very launch shape (tools/profile_r2j.sh r2l)
    traffic = None
    # ... and, from the same capture, what the kernel keeps busy: it is bound by issue slots of
    # the integer / FP32 epilogue, not by the tensor pipe the roofline below is stated against
    ncu_pct = {"smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_pct",
               "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
               "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
               "smsp__inst_executed.sum": "warp_instructions"}
    ncu_k1 = {}
    try:
        if U == 4096 and not args.compallsen:
            vals = {}
            unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            for ln in open(os.path.join(ROOT, "profiles", "prof_gmm_topn_r2l.txt")):
                f = ln.split()
                if len(f) >= 3 and f[0] in ncu_pct:
                    ncu_k1[ncu_pct[f[0]]] = float(f[2] if f[1] in ("%", "inst") else f[1])
                if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    # (tools/ncu_digest.py prints name unit value, tools/summarise_profiles.py name value unit)
                    vals[f[0]] = float(f[2]) * unit[f[1]] if f[1] in unit else float(f[1]) * unit[f[2]]
            if len(vals) == 2:
                traffic = sum(vals.values())
    except Exception:
        traffic = None
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback"
    # K1 algorithmic FLOPs: scanned codebook-frames x streams x densities x 2(2D+1)  
    flop_per_cbframe = model.n_feat * model.n_density * 2 * (2 * model.veclen + 1)
    k1_flops = stats["scanned_cb_frames"] * flop_per_cbframe
    k1_ms = kms["gmm_topn"] / args.steps
    achieved_tf = k1_flops / (k1_ms * 1e-3) / 1e12
    # FP32-pipe view of the same kernel: 4 dependent-rounding FP32 ops per (density, dim)
    fp32_ops = stats["scanned_cb_frames"] * model.n_feat * model.n_density * model.veclen * 4
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    fp32_peak = 148 * 128 * sm_mhz * 1e6
    line = {
        "metric": "forced-align audio-sec/sec", "value": world * audio_s / (dev_ms_max * 1e-3),
        "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+i32", "data": "synthetic",
        "config": {"workload": workload, "utts_per_gpu": U, "frames_per_utt": FRAMES,
                   "l2": "inputs (%.0f MB/rank) and per-step intermediates exceed the 126 MB L2; "
                         "no flush needed" % (feat_nbytes / 1e6),
                   "failed_alignments": n_fail,
                   "frames_covered_by_state_segments": frames_covered,
                   "frames_total": world * U * FRAMES, "sharding": "utt % n_gpu, no data-path collective"},
        "e2e": {"value": world * audio_s / (e2e_ms_max * 1e-3), "unit": "audio-s/s",
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_ms_max,
                "path": "ssb_pipeline_submit/collect: a stream of whole batches, two in flight (batch i+1 is "
                        "planned and copied in while batch i computes; kernels of different batches never "
                        "overlap), pinned host features in, segmentations out, %d kernel launches per step; "
                        "results identical to the single resident batch: %s" % (stream_launches, stream_same),
                "single_call_ms_per_step": e2e_call_s * 1e3,
                "single_call_path": "ssb_pipeline_align on one batch: %d chunks of whole utterances through 4 "
                                    "lanes (stream + host thread each), %d launches; identical results: %s"
                                    % (pipe_chunks, pipe_launches, pipe_same),
                "single_batch_ms_per_step": e2e_one_s * 1e3,
                "single_batch_host_call_ms": {"upload(plan+H2D issue)": 1e3 * host_split[0] / args.steps,
                                 "run(launch)": 1e3 * host_split[1] / args.steps,
                                 "download(wait+D2H+scatter)": 1e3 * host_split[2] / args.steps}},
        "gpu_launches": launches,
        "kernel_ms_per_step": {k: v / args.steps for k, v in kms.items()},
        "clocks": clocks,
        "host_affinity": numa_note,
        "roofline": {"kernel": "gmm_topn_tc2_kernel (K1: tcgen05 3xTF32 screening GEMM in TMEM + exact "
                               "FP32 survivors + top-N)" if not os.environ.get("SSB_K1") else
                               "K1 variant SSB_K1=%s" % os.environ.get("SSB_K1"),
                     "bound": "tensor",
                     "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / peak_tf, "traffic": traffic, "ncu": ncu_k1 or None,
                     "traffic_source": "profiles/prof_gmm_topn_r2l.txt (ncu --set full, same launch shape, this round's build), bytes per launch",
                     "peak_source": peak_src,
                     "note": "achieved = ALGORITHMIC flops (SURVEY 8d: scanned codebook-frames x 3 streams x "
                             "128 densities x 2(2*13+1)) / CUDA-event time of the kernel; the MMAs actually "
                             "executed are 3 split products on K padded to 32 (mma_tflops_executed); the "
                             "kernel is bound by the integer/FP32 epilogue (top-N selection + exact "
                             "re-scoring), see DESIGN.md; scalar_equiv_fp32_frac = what the reference's "
                             "scalar scan of the same densities would need of the FP32 lanes",
                     "mma_tflops_executed": stats["scanned_cb_frames"] * model.n_feat * 3 * 2 * 128 * 32
                                            / (k1_ms * 1e-3) / 1e12,
                     "scalar_equiv_fp32_frac": fp32_ops / (k1_ms * 1e-3) / fp32_peak},
        "roofline_other": [
            {"kernel": "chain_viterbi_kernel (K3)", "bound": "hbm", "unit": "GB/s",
             "achieved": stats["band_state_frames"] * 10 / (kms["chain_viterbi"] / args.steps * 1e-3) / 1e9,
             "peak": float(peaks.get("hbm_gbs", 6650.0)),
             "frac": stats["band_state_frames"] * 10 / (kms["chain_viterbi"] / args.steps * 1e-3) / 1e9
                     / float(peaks.get("hbm_gbs", 6650.0)),
             "band_state_frames": stats["band_state_frames"], "dense_state_frames": stats["state_frames"],
             "reference_equivalent_dense_10B_frac": stats["state_frames"] * 10
                     / (kms["chain_viterbi"] / args.steps * 1e-3) / 1e9 / float(peaks.get("hbm_gbs", 6650.0)),
             "note": "bytes = 10 B (2 B score in + 8 B token out, SURVEY 8d) x the state-frames the kernel "
                     "EVALUATES (the word-window band the planner computed: band_state_frames); the dense "
                     "T x states count the reference allocates tokens for is reported separately as "
                     "reference_equivalent_dense_10B_frac and is NOT a bandwidth; the kernel is issue / "
                     "latency bound (profiles/prof_chain_viterbi_r1j.txt)"},
            {"kernel": "senone_mix_active_kernel (K2)", "bound": "hbm", "unit": "GB/s",
             "achieved": (stats["scanned_cb_frames"] * model.n_feat * 20 + stats["state_frames"] * 2)
                         / (kms["senone_mix"] / args.steps * 1e-3) / 1e9,
             "peak": float(peaks.get("hbm_gbs", 6650.0)),
             "frac": (stats["scanned_cb_frames"] * model.n_feat * 20 + stats["state_frames"] * 2)
                     / (kms["senone_mix"] / args.steps * 1e-3) / 1e9 / float(peaks.get("hbm_gbs", 6650.0)),
             "note": "algorithmic bytes: 20 B of top-N list per scanned codebook-stream-frame in + 2 B per "
                     "state-frame out (ncu: 5.0 + 1.3 GB per launch); integer mixing, issue-bound"}],
        "senone_scores_per_s": stats["active_senone_frames"] / ((kms["gmm_topn"] + kms["senone_mix"]) / args.steps * 1e-3),
        "dp_state_frames_per_s": stats["band_state_frames"] / (kms["chain_viterbi"] / args.steps * 1e-3),
        "dp_hbm_frac_10B": stats["band_state_frames"] * 10 / (kms["chain_viterbi"] / args.steps * 1e-3) / 1e9
                           / float(peaks.get("hbm_gbs", 6650.0)),
        "wall_ms_per_step_device_loop": 1e3 * wall_dev / args.steps,
    }
    for k, v in extra.items():
        line[k] = v
    if world == 1 and not args.no_cpu_baseline and extra and cpu_kind() == "reference":
        import bench_configs as bc
        cores = host_cores()
        if "config3" in line:
            line["config3"]["cpu_baseline"] = bc.cpu_rate("config3", cores, 12)
        if "config5_two_pass" in line:
            line["config5_two_pass"]["cpu_baseline"] = bc.cpu_rate("two_pass", cores, 2)
        if "config4" in line:
            dt, audio, ok = bc.cpu_config4(2.0)
            line["config4"]["cpu_baseline"] = {
                "value": audio / dt, "unit": "audio-s/s", "cores": 1, "kind": "reference",
                "sample": "second pass of the reference on the first %.0f s of the utterance (the reference "
                          "is single-threaded per utterance; its token stack for the whole hour would be "
                          "86 GB), valid=%s" % (audio, ok)}
    if world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        kind = cpu_kind()
        n = cores * 32
        pool = CpuPool(cores, kind)
        v, ok, wall = pool.run(n)
        pool.close()
        line["cpu_baseline"] = {"value": v, "unit": "audio-s/s", "cores": cores, "kind": kind,
                                "sample": "%d utterances of the same workload (%.0f audio-s, %.1f s wall), "
                                          "second pass of the reference (re-score + state_align_search), "
                                          "valid=%s" % (n, n * FRAMES / FRAME_RATE, wall, ok)}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
