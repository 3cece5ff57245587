"""Utterance sharding across the GPUs of one box.

The path partitions by utterance (every utterance is an independent scorer + DP instance,
SURVEY.md §8e): rank r of n aligns utterances r, r+n, r+2n, ... with its own model replica and
no collective on the data path.  The only exchange is the optional gather of the (tiny,
fixed-size) per-utterance results onto rank 0.

Works with any initialised torch.distributed backend: NCCL on the GPU box (tensors are moved
to the rank's device), gloo in the CPU tests.
"""
import numpy as np


def shard_indices(n_utts, rank, world):
    """Indices of the utterances rank `rank` of `world` owns (`utt % n_gpu` sharding)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return np.arange(rank, n_utts, world, dtype=np.int64)


def balanced_shards(lengths, world):
    """Length-balanced alternative for ragged batches: greedy longest-first bin packing on
    frames x phones.  Returns a list of index arrays, one per rank."""
    order = np.argsort(-np.asarray(lengths, np.int64), kind="stable")
    load = np.zeros(world, np.int64)
    bins = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        bins[r].append(int(i))
        load[r] += int(lengths[i])
    return [np.asarray(sorted(b), np.int64) for b in bins]


def gather_results(local, n_utts, rank, world, dist=None, device=None):
    """Gather per-utterance results onto rank 0.

    `local` = {utterance index: dict of int32 numpy arrays (start, dur, score) + ints (rv,
    best_score)} for the utterances this rank aligned.  Returns the full list on rank 0 and
    None elsewhere.  Records are flattened to one int32 tensor per rank and exchanged with a
    single all_gather of sizes plus one gather of payloads."""
    if world == 1 or dist is None:
        return [local[u] for u in range(n_utts)]
    import torch
    dev = device if device is not None else torch.device("cpu")
    parts = []
    for u in sorted(local):
        r = local[u]
        n = len(r["start"])
        parts.append(np.concatenate([np.array([u, n, r["rv"], r["best_score"]], np.int32),
                                     np.asarray(r["start"], np.int32), np.asarray(r["dur"], np.int32),
                                     np.asarray(r["score"], np.int32)]))
    flat = np.concatenate(parts) if parts else np.zeros(0, np.int32)
    size = torch.tensor([flat.size], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, size)
    cap = int(max(int(s.item()) for s in sizes))
    buf = torch.zeros(max(cap, 1), dtype=torch.int32, device=dev)
    if flat.size:
        buf[:flat.size] = torch.from_numpy(flat).to(dev)
    gathered = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, gathered, dst=0)
    if rank != 0:
        return None
    out = [None] * n_utts
    for r in range(world):
        a = gathered[r].cpu().numpy()[:int(sizes[r].item())]
        i = 0
        while i < len(a):
            u, n, rv, best = (int(x) for x in a[i:i + 4])
            i += 4
            out[u] = dict(rv=rv, best_score=best, start=a[i:i + n].copy(),
                          dur=a[i + n:i + 2 * n].copy(), score=a[i + 2 * n:i + 3 * n].copy())
            i += 3 * n
    if any(o is None for o in out):
        raise RuntimeError("gather_results: some utterances were not aligned by any rank")
    return out
