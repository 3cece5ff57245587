#!/usr/bin/env python
"""Audio in, the reference CLI's JSON out, for a batch (BASELINE config #5's "end-to-end" on one
GPU): frontend (PCM -> features, stays in HBM) -> ssb_align_texts (alignment grammar, first pass
in the reference's default mode, chains, second pass, decoder_result_json).  Audio = goforward.raw
three times + its trailing silence to 10 s + per-utterance noise; transcript = the sentence
three times.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import soundswallower_b200 as ssb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--align-level", type=int, default=1)
    args = ap.parse_args()
    hmm = os.path.join(ROOT, "soundswallower_b200/model/en-us")
    pcm = np.frombuffer(open(os.path.join(ROOT, "tests/data/goforward.raw"), "rb").read(), np.int16)
    n = 160000
    sil = pcm[33760:]                      # the trailing silence (frames 211..)
    base = np.concatenate([pcm] * 3)
    base = np.concatenate([base, np.tile(sil, (n - len(base)) // len(sil) + 1)])[:n].astype(np.int32)
    # all utterances back to back in one host buffer (pinned when torch is there), as a reader
    # thread would leave them
    try:
        import torch
        flat = torch.empty(args.utts * n, dtype=torch.int16, pin_memory=True).numpy()
    except Exception:
        flat = np.empty(args.utts * n, np.int16)
    for u in range(args.utts):
        rng = np.random.Generator(np.random.Philox(99 + u))
        flat[u * n:(u + 1) * n] = (base + rng.integers(-30, 31, n)).astype(np.int16)
    samp_off = np.arange(args.utts + 1, dtype=np.int64) * n
    texts = [" ".join(["go forward ten meters"] * 3)] * args.utts
    m = ssb.AcousticModel(hmm)
    lx = ssb.Lexicon(m, hmmdir=hmm)
    fe = ssb.Frontend(hmm)
    best = None
    for _ in range(args.steps + 1):
        t0 = time.perf_counter()
        feats = fe.run_raw(flat, samp_off)
        t1 = time.perf_counter()
        ta = ssb.TextAlignment(m, lx, feats, texts, align_level=args.align_level)
        t2 = time.perf_counter()
        ta.render(align_level=args.align_level)
        js = [ta.json(u, align_level=args.align_level) for u in range(args.utts)]
        t3 = time.perf_counter()
        ok = sum(1 for u in range(args.utts) if ta.status(u)[0] == 0)
        cur = (t3 - t0, t1 - t0, t2 - t1, t3 - t2, ok, ta.kernel_ms(), js[0])
        ta.close()
        if best is None or cur[0] < best[0]:
            best = cur
    wall, t_fe, t_al, t_js, ok, ms1, j0 = best
    audio_s = args.utts * n / 16000.0
    words = json.loads(j0)["w"]
    print(json.dumps({"workload": "audio -> JSON: %d x 10 s utterances (16 kHz int16), 12-word transcript, "
                                  "align_level %d, en-us" % (args.utts, args.align_level),
                      "wall_s": wall, "audio_s_per_s": audio_s / wall,
                      "split_s": {"frontend (H2D + kernels)": t_fe, "ssb_align_texts (both passes + host)": t_al,
                                  "decoder_result_json": t_js},
                      "pass1_kernel_ms": ms1, "aligned": ok,
                      "first_utterance": {"text": json.loads(j0)["t"], "n_words": len(words),
                                          "last_word": words[-1]}}))


if __name__ == "__main__":
    main()
