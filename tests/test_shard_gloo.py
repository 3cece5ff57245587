"""The N > 1 host logic on CPU: two processes (gloo, world_size 2) shard a batch by utterance,
align their shares (with the oracle standing in for the GPU), and rank 0 gathers everything.
The result must equal the single-process alignment of the whole batch."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_batch():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import model_dir, model_features, random_chain
    from oracle.oracle import Oracle
    o = Oracle(model_dir("fr-fr"))
    rs = np.random.RandomState(123)
    feats, chains = [], []
    for u in range(7):
        T = int(rs.randint(5, 40))
        feats.append(model_features(rs, o.model_arrays(), T))
        chains.append(random_chain(rs, o, int(rs.randint(1, 8)), T, windowed=u % 2 == 0))
    return o, feats, chains


def _align(o, f, c):
    r = o.state_align(f, c["ssid"], c["tmat"], c["sf"], c["ef"])
    return dict(rv=r["rv"], best_score=r["best_score"], start=r["start"], dur=r["dur"], score=r["score"])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from soundswallower_b200 import shard
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        o, feats, chains = _make_batch()
        mine = shard.shard_indices(len(feats), rank, world)
        local = {int(u): _align(o, feats[u], chains[u]) for u in mine}
        out = shard.gather_results(local, len(feats), rank, world, dist=dist)
        dist.barrier()
        if rank == 0:
            q.put([(r["rv"], r["best_score"], r["start"].tolist(), r["dur"].tolist(), r["score"].tolist())
                   for r in out])
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o, feats, chains = _make_batch()
    want = [_align(o, f, c) for f, c in zip(feats, chains)]
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g[0] == w["rv"] and g[1] == w["best_score"]
        assert g[2] == w["start"].tolist() and g[3] == w["dur"].tolist() and g[4] == w["score"].tolist()


def test_shard_indices_partition():
    from soundswallower_b200 import shard
    for n in (0, 1, 7, 64):
        for world in (1, 2, 4, 8):
            allidx = np.concatenate([shard.shard_indices(n, r, world) for r in range(world)])
            assert sorted(allidx.tolist()) == list(range(n))
    with pytest.raises(ValueError):
        shard.shard_indices(4, 2, 2)
    bins = shard.balanced_shards([100, 5, 5, 90, 50, 50], 2)
    assert sorted(np.concatenate(bins).tolist()) == list(range(6))
    loads = [sum([100, 5, 5, 90, 50, 50][i] for i in b) for b in bins]
    assert abs(loads[0] - loads[1]) <= 10
