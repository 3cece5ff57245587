"""Several GPUs behind the C ABI: ssb_align_batch_multi = one host thread per GPU,
utterance ranges, no collective.  Needs two devices (gpurun --gpus 2); with one device the call
is checked in its single-device form."""
import numpy as np
import pytest

import soundswallower_b200 as ssb
from conftest import chain_from_golden, model_dir

pytestmark = pytest.mark.gpu


def _batch(g, n):
    rs = np.random.RandomState(5)
    feat = g["feat"]
    feats, chains = [], []
    for u in range(n):
        T = int(rs.randint(150, len(feat) + 1)) if u % 3 else len(feat)
        feats.append((feat[:T] + rs.normal(0, 0.05, (T, feat.shape[1]))).astype(np.float32))
        chains.append(chain_from_golden(g, windows=(T == len(feat))))
    return feats, chains


def test_align_batch_multi_equals_single_device(golden):
    g = golden["en-us"]
    n_dev = min(ssb.device_count(), 4)
    models = [ssb.AcousticModel(model_dir("en-us"), device=d) for d in range(n_dev)]
    feats, chains = _batch(g, 37)
    want = ssb.align_batch(models[0], feats, chains)
    got = ssb.align_batch_multi(models, feats, chains)
    assert len(got) == len(want) == 37
    for a, b in zip(got, want):
        assert a["rv"] == b["rv"] and a["best_score"] == b["best_score"]
        for k in ("start", "dur", "score"):
            assert np.array_equal(a[k], b[k]), k
    # the shard helper's gather on the same results (what a multi-process caller does instead)
    from soundswallower_b200 import shard
    parts = [shard.shard_indices(37, r, max(n_dev, 2)) for r in range(max(n_dev, 2))]
    assert sorted(int(i) for p in parts for i in p) == list(range(37))
    for m in models:
        m.close()


def test_align_batch_multi_needs_matching_models(golden):
    m = ssb.AcousticModel(model_dir("en-us"), device=0)
    f = ssb.AcousticModel(model_dir("fr-fr"), device=0)
    feats, chains = _batch(golden["en-us"], 4)
    with pytest.raises(ssb.SsbError, match="different from model 0"):
        ssb.align_batch_multi([m, f], feats, chains)
    m.close()
    f.close()
