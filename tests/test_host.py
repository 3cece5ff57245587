"""Host-side logic of libssb200.so that needs no GPU: the C ABI loads and exports what
include/ssb200.h declares, the model loaders reproduce the reference's in-memory tables
bit for bit, the chain planner agrees with the oracle's frame bookkeeping, and compute
entry points refuse to run without a device (no CPU fallback)."""
import ctypes as C
import hashlib
import os
import re

import numpy as np
import pytest

import soundswallower_b200 as ssb
from soundswallower_b200 import _lib
from conftest import ROOT, chain_from_golden, model_dir, random_chain


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "ssb200.h")).read()
    declared = set(re.findall(r"\b(ssb_[a-z0-9_]+)\s*\(", hdr))
    L = _lib.load()
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert declared == set(_lib.SYMBOLS)
    assert L.ssb_version() >= 100


def test_struct_layouts_match_header():
    # mgau_t prefix: {vt pointer, int frame_idx} (ref: include/soundswallower/acmod.h:108-111)
    assert _lib.MgauBase.vt.offset == 0 and _lib.MgauBase.frame_idx.offset == C.sizeof(C.c_void_p)
    assert [f[0] for f in _lib.MgauFuncs._fields_] == ["name", "frame_eval", "transform", "free"]
    assert C.sizeof(_lib.Config) == 64 and _lib.Config.topn_beam.offset == 44
    assert C.sizeof(_lib.FeConfig) == 72
    assert C.sizeof(_lib.AlignIn) == 88 and C.sizeof(_lib.AlignOut) == 64


@pytest.mark.parametrize("lang", ["en-us", "fr-fr"])
def test_loader_matches_reference_tables(golden, oracles, lang):
    m = ssb.AcousticModel(model_dir(lang), device=-1)
    g, o = golden[lang], oracles(lang)
    assert [m.n_mgau, m.n_feat, m.n_density, m.veclen, m.n_sen, m.n_sseq, m.n_emit, m.n_tmat,
            m.n_ciphone, m.n_phone, m.sil] == g["dims"].tolist()
    assert m.blk == 39 and m.featlen == [13, 13, 13]
    for k, v in m.arrays().items():
        assert sha(v) == str(g["model_sha_" + k]), k
    for a, b in zip(m.phone_table(), o.phone_table()):
        assert np.array_equal(a, b)
    m.close()


def test_loader_errors_are_reported(tmp_path):
    with pytest.raises(ssb.SsbError, match="cannot read"):
        ssb.AcousticModel(str(tmp_path), device=-1)
    with pytest.raises(ssb.SsbError, match="topn"):
        ssb.AcousticModel(model_dir("en-us"), device=-1, topn=9)
    # a truncated means file must fail its size/checksum test, not load garbage
    src = model_dir("en-us")
    for f in os.listdir(src):
        data = open(os.path.join(src, f), "rb").read()
        open(os.path.join(tmp_path, f), "wb").write(data[:-8] if f == "means" else data)
    with pytest.raises(ssb.SsbError):
        ssb.AcousticModel(str(tmp_path), device=-1)


def test_no_cpu_fallback():
    """Without a device handle every compute call fails loudly."""
    m = ssb.AcousticModel(model_dir("en-us"), device=-1)
    with pytest.raises(ssb.SsbError, match="no CPU compute path"):
        ssb.PtmMgau(m)
    with pytest.raises(ssb.SsbError, match="no CPU compute path"):
        ssb.StateAlignBatch(m)
    with pytest.raises(ssb.SsbError, match="no CPU compute path"):
        ssb.score_batch(m, [np.zeros((3, 39), np.float32)])
    with pytest.raises(ssb.SsbError, match="no CPU compute path"):
        ssb.hmm_vit_eval(m, 0, [0, 1, 2], np.zeros(m.n_sen, np.int16), np.zeros(12, np.int32))
    m.close()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "soundswallower_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.lower().replace("oracle/_ref", "") or f == "__init__.py" and False, \
                    os.path.join(dirpath, f)


def _oracle_activity(o, feat, chain):
    """(T, n_phones) bool: phone has a recorded token on frame t, from the oracle's stack."""
    r = o.state_align(feat, chain["ssid"], chain["tmat"], chain["sf"], chain["ef"],
                      compallsen=True, want_tokens=True)
    tok = r["tokens"]  # [T][ns][2]; {-1,-1} where nothing was recorded
    rec = ~((tok[:, :, 0] == -1) & (tok[:, :, 1] == -1))
    return rec.reshape(tok.shape[0], -1, o.n_emit).any(2)


def _planned_activity(T, chain):
    enter = ssb.plan_chain(T, chain["sf"], chain["ef"])
    act = np.zeros((T, len(enter)), bool)
    for i, e in enumerate(enter):
        if e < 0:
            continue
        last = max(int(e), int(chain["ef"][i]))
        for t in range(T):
            # recorded on frame t: evaluated on t, or entered at the end of t
            act[t, i] = (e <= t <= last) or (e == t + 1)
    return act


def _evaluated(rec):
    """evaluated on t <=> recorded on t and already in the stack on t-1 (phone 0: from t=0)."""
    ev = rec.copy()
    ev[1:] &= rec[:-1]
    if rec.shape[0]:
        ev[0, 1:] = False
    return ev


def _planned_evaluated(T, chain):
    enter = ssb.plan_chain(T, chain["sf"], chain["ef"])
    ev = np.zeros((T, len(enter)), bool)
    for i, e in enumerate(enter):
        if e >= 0:
            ev[int(e):max(int(e), min(int(chain["ef"][i]), T - 1)) + 1, i] = True
    return ev


def test_planner_matches_oracle_bookkeeping_golden(golden, oracles):
    for lang in ("en-us", "fr-fr"):
        g, o = golden[lang], oracles(lang)
        for win in (True, False):
            chain = chain_from_golden(g, windows=win)
            T = g["feat"].shape[0]
            rec = _oracle_activity(o, g["feat"], chain)
            assert np.array_equal(_planned_activity(T, chain), rec)
            assert np.array_equal(_planned_evaluated(T, chain), _evaluated(rec))


def test_planner_matches_oracle_bookkeeping_random(oracles):
    o = oracles("en-us")
    rs = np.random.RandomState(11)
    from conftest import model_features
    arrays = o.model_arrays()
    for case in range(12):
        T = int(rs.randint(1, 40))
        npn = int(rs.randint(1, 12))
        chain = random_chain(rs, o, npn, T, windowed=case % 3 != 0)
        if case % 4 == 1 and npn > 2:
            chain["sf"][npn // 2:] = np.maximum(chain["sf"][npn // 2:], T // 2)  # late starts
        feat = model_features(rs, arrays, T)
        rec = _oracle_activity(o, feat, chain)
        assert np.array_equal(_planned_activity(T, chain), rec), case
        assert np.array_equal(_planned_evaluated(T, chain), _evaluated(rec)), case


def test_planner_rejects_decreasing_window_ends():
    with pytest.raises(ssb.SsbError, match="must not decrease"):
        ssb.plan_chain(10, [0, 0, 0], [5, 9, 7])
    assert ssb.plan_chain(0, [0, 0], [ssb.INT_MAX] * 2).tolist() == [-1, -1]
    # unconstrained chain: everything is entered after the first step
    assert ssb.plan_chain(5, [0] * 4, [ssb.INT_MAX] * 4).tolist() == [0, 1, 1, 1]
