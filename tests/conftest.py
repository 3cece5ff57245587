"""Shared fixtures.  Markers: `gpu` = needs a B200 (parity tests proper, through the C ABI).

Only tests/ (plus __graft_entry__.smoke and bench.py's CPU legs) may import oracle/.
Nothing here reads /root/reference: the fixtures under tests/golden/ were generated
from the reference by tools/make_golden.py and are committed.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "tests", "data")
MODELS = os.path.join(ROOT, "soundswallower_b200", "model")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


def _has_gpu():
    # (a library that does not load -- stale build, missing symbol -- must fail the run, not skip
    # every GPU test: only "no device" is a reason to skip)
    import soundswallower_b200 as ssb
    from soundswallower_b200 import _build, _lib
    if not os.path.exists(_lib.LIB_PATH):   # a fresh checkout: compile first (nvcc needs no GPU)
        _build.build_lib()
    return ssb.device_count() > 0


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no sm_100a device visible")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The product library and the oracle are compiled once per session."""
    from soundswallower_b200 import _build
    _build.build_lib()
    from oracle import oracle as orc
    orc.build()


def model_dir(lang):
    return os.path.join(MODELS, lang)


@pytest.fixture(scope="session")
def golden():
    return {lang: np.load(os.path.join(GOLDEN, "align_%s.npz" % lang)) for lang in ("en-us", "fr-fr")}


@pytest.fixture(scope="session")
def synthetic():
    return np.load(os.path.join(GOLDEN, "synthetic_en-us.npz"))


@pytest.fixture(scope="session")
def oracles():
    from oracle.oracle import Oracle
    cache = {}

    def get(lang):
        if lang not in cache:
            cache[lang] = Oracle(model_dir(lang))
        return cache[lang]
    return get


@pytest.fixture(scope="session")
def models():
    """AcousticModel on cuda:0 (gpu tests only)."""
    import soundswallower_b200 as ssb
    cache = {}

    def get(lang, **kw):
        key = (lang, tuple(sorted(kw.items())))
        if key not in cache:
            cache[key] = ssb.AcousticModel(model_dir(lang), device=0, **kw)
        return cache[key]
    return get


def chain_from_golden(g, windows=True):
    """Per-phone chain arrays of the golden utterance (what state_align_search_init derives)."""
    import soundswallower_b200 as ssb
    ph = g["phones"]
    if windows:
        sf, ef = ssb.windows(g["ph_start"], g["ph_dur"])
    else:
        sf, ef = ssb.windows(np.zeros(len(ph), np.int32), np.zeros(len(ph), np.int32))
    return dict(ssid=ph[:, 1].astype(np.int32), tmat=ph[:, 2].astype(np.int32), sf=sf, ef=ef)


def random_chain(rs, orc_or_model, n_phones, T, windowed=True):
    """A random phone chain (random triphone ids) with optional monotone word-like windows."""
    ssid_t, tmat_t, _ = orc_or_model.phone_table()
    pid = rs.randint(0, len(ssid_t), n_phones)
    ssid = ssid_t[pid].astype(np.int32)
    tmat = tmat_t[pid].astype(np.int32)
    if windowed and n_phones > 1 and T > 4:
        # split the phones into words, the frames into word windows
        n_words = max(1, min(n_phones, int(rs.randint(1, max(2, n_phones // 2 + 1)))))
        cuts = np.sort(rs.choice(np.arange(1, n_phones), n_words - 1, replace=False)) if n_words > 1 else np.zeros(0, int)
        word_of = np.searchsorted(cuts, np.arange(n_phones), side="right")
        fcuts = np.sort(rs.choice(np.arange(1, T), n_words - 1, replace=False)) if n_words > 1 else np.zeros(0, int)
        wstart = np.concatenate([[0], fcuts]).astype(np.int32)
        wend = np.concatenate([fcuts, [T]]).astype(np.int32)
        import soundswallower_b200 as ssb
        sf, ef = ssb.windows(wstart[word_of], (wend - wstart)[word_of])
    else:
        sf = np.zeros(n_phones, np.int32)
        ef = np.full(n_phones, 2**31 - 1, np.int32)
    return dict(ssid=ssid, tmat=tmat, sf=sf, ef=ef)


def model_features(rs, arrays, T, noise=0.7):
    """Frames drawn around randomly chosen Gaussians of the model (realistic score ranges)."""
    mean = arrays["mean"]
    n_mgau, n_feat, n_den, _ = mean.shape
    cb = rs.randint(0, n_mgau, T)
    dn = rs.randint(0, n_den, T)
    x = np.stack([np.concatenate([mean[cb[t], f, dn[t]] for f in range(n_feat)]) for t in range(T)]) \
        if T else np.zeros((0, n_feat * mean.shape[3]), np.float32)
    return (x + rs.normal(0, noise, x.shape)).astype(np.float32)
