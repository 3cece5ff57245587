"""Chain cutting (csrc/api.cu: cut_chains): where one word window ends exactly where the next begins
the chain can only be crossed on that frame (prune_hmms / phone_transition, ref:
src/state_align_search.c:88-133), so K3 + backtrace run on the segments in parallel.  Results must
be those of the uncut chain -- i.e. the oracle's / the reference's -- state for state."""
import numpy as np
import pytest

import soundswallower_b200 as ssb
from conftest import model_features, random_chain
from test_gpu_parity import _tiled_fr

pytestmark = pytest.mark.gpu


def _run(m, feats, chains, init_active=None):
    b = ssb.StateAlignBatch(m)
    b.upload(feats, chains, init_active=init_active)
    b.run()
    res = b.per_utt(b.download())
    st = b.stats()
    b.close()
    return res, st


def _check(o, feats, chains, res, init_active=None):
    n_ok = n_fail = 0
    for u, (f, c, r) in enumerate(zip(feats, chains, res)):
        w = o.state_align(f, c["ssid"], c["tmat"], c["sf"], c["ef"],
                          init_active=None if init_active is None else init_active[u])
        assert r["rv"] == w["rv"], u
        if w["rv"] != 0:
            n_fail += 1
            continue
        n_ok += 1
        assert r["best_score"] == w["best_score"], u
        assert r["n_renorm"] == w.get("n_renorm", 0), u
        for k in ("start", "dur", "score"):
            assert np.array_equal(r[k], w[k]), (u, k)
    return n_ok, n_fail


def test_every_cuttable_chain_cut_equals_oracle(models, oracles, monkeypatch):
    """Random ragged batch with random word windows (feasible and not), every eligible cut made."""
    monkeypatch.setenv("SSB_K3_CUT", "all")
    m, o = models("en-us"), oracles("en-us")
    rs = np.random.RandomState(4242)
    arrays = o.model_arrays()
    feats, chains = [], []
    for u in range(72):
        T = int(rs.randint(30, 160))
        feats.append(model_features(rs, arrays, T))
        chains.append(random_chain(rs, o, int(rs.randint(1, 22)), T, windowed=u % 5 != 0))
    res, st = _run(m, feats, chains)
    assert st["segments"] > 2 * len(feats)          # the cut really happened
    n_ok, n_fail = _check(o, feats, chains, res)
    assert n_ok >= 25 and n_fail >= 3               # windows too short for their words fail in both
    # ... and the same batch uncut
    monkeypatch.setenv("SSB_K3_CUT", "0")
    res0, st0 = _run(m, feats, chains)
    assert st0["segments"] == len(feats)
    for r, r0 in zip(res, res0):
        assert r["rv"] == r0["rv"]
        if r0["rv"] == 0:
            assert r["best_score"] == r0["best_score"]
            for k in ("start", "dur", "score"):
                assert np.array_equal(r[k], r0[k])


def test_cut_with_flags_carried_in(models, oracles, monkeypatch):
    """init_active (what a first pass left in acmod) reaches every segment through the plan."""
    monkeypatch.setenv("SSB_K3_CUT", "all")
    m, o = models("en-us"), oracles("en-us")
    rs = np.random.RandomState(99)
    arrays = o.model_arrays()
    feats, chains, init = [], [], []
    for u in range(16):
        T = int(rs.randint(40, 120))
        feats.append(model_features(rs, arrays, T))
        chains.append(random_chain(rs, o, int(rs.randint(4, 16)), T, windowed=True))
        init.append(sorted(set(int(x) for x in rs.randint(0, m.n_sen, size=rs.randint(0, 6)))))
    res, st = _run(m, feats, chains, init_active=init)
    assert st["segments"] > len(feats)
    n_ok, _ = _check(o, feats, chains, res, init_active=init)
    assert n_ok >= 6


def test_long_chain_is_cut_by_default(models, oracles, golden):
    """169 phones (> 128): cut at its word windows without being asked; 39 segments."""
    m, o = models("fr-fr"), oracles("fr-fr")
    x, chain = _tiled_fr(golden, 12)
    res, st = _run(m, [x], [chain])
    assert st["segments"] > 30
    r = res[0]
    w = o.state_align(x, chain["ssid"], chain["tmat"], chain["sf"], chain["ef"])
    assert r["rv"] == w["rv"] == 0 and r["best_score"] == w["best_score"]
    for k in ("start", "dur", "score"):
        assert np.array_equal(r[k], w[k]), k
    on = r["dur"] > 0
    assert r["start"][on][0] == 0 and (r["start"][on][1:] == (r["start"][on] + r["dur"][on])[:-1]).all()


def test_unwindowed_long_chain_is_not_cut(models, oracles, golden):
    m = models("fr-fr")
    x, chain = _tiled_fr(golden, 12)
    chain = dict(chain, sf=chain["sf"] * 0, ef=chain["ef"] * 0 + ssb.INT_MAX)
    _, st = _run(m, [x], [chain])
    assert st["segments"] == 1


@pytest.mark.parametrize("lanes", ["8", "16", "32"])
def test_lane_groups_and_cuts_on_french(models, oracles, monkeypatch, lanes):
    """K3 with 8 / 16 / 32 lanes per utterance (several utterances per warp, kept converged), every
    cut made, fr-fr, ragged lengths so that the groups of a warp finish at different frames."""
    monkeypatch.setenv("SSB_K3_CUT", "all")
    monkeypatch.setenv("SSB_K3_LANES", lanes)
    m, o = models("fr-fr"), oracles("fr-fr")
    rs = np.random.RandomState(31 + int(lanes))
    arrays = o.model_arrays()
    feats, chains = [], []
    for u in range(37):
        T = int(rs.randint(30, 200))
        feats.append(model_features(rs, arrays, T))
        chains.append(random_chain(rs, o, int(rs.randint(1, 26)), T, windowed=u % 4 != 0))
    res, st = _run(m, feats, chains)
    assert st["segments"] > len(feats)
    n_ok, _ = _check(o, feats, chains, res)
    assert n_ok >= 12


def test_cut_chains_with_five_state_hmms(tmp_path_factory, golden, monkeypatch):
    """The synthetic 5-state model (skip arcs: states off the best path keep the caller's values):
    cut and uncut give the same entries, and they are the oracle's."""
    import model_variants as mv
    from conftest import model_dir
    from oracle.oracle import Oracle
    d = mv.write_five_state_model(model_dir("en-us"), str(tmp_path_factory.mktemp("five_cut") / "five"))
    m = ssb.AcousticModel(d, device=0)
    o = Oracle(d)
    lx = ssb.Lexicon(m, hmmdir=d)
    g = golden["en-us"]
    words = g["words"]
    chain = lx.populate(words[:, 0], words[:, 1], words[:, 2])
    rs = np.random.RandomState(8)
    feats = [g["feat"], g["feat"] + rs.randn(*g["feat"].shape).astype(np.float32) * np.float32(0.2)]
    chains = [dict(ssid=chain["ssid"], tmat=chain["tmat"], sf=chain["sf"], ef=chain["ef"])] * 2
    out = {}
    for mode in ("all", "0"):
        monkeypatch.setenv("SSB_K3_CUT", mode)
        out[mode], st = _run(m, feats, chains)
        assert (st["segments"] > 2) == (mode == "all")
    for u, (f, c) in enumerate(zip(feats, chains)):
        w = o.state_align(f, c["ssid"], c["tmat"], c["sf"], c["ef"])
        for mode in ("all", "0"):
            r = out[mode][u]
            assert r["rv"] == w["rv"] == 0 and r["best_score"] == w["best_score"], (u, mode)
            for k in ("start", "dur", "score"):
                assert np.array_equal(r[k], w[k]), (u, mode, k)
    lx.close()
    m.close()
