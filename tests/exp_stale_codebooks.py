#!/usr/bin/env python
"""Offline experiment (CPU, oracle): how many of the codebooks the aligner scans per frame are
"stale" (flagged by an HMM that is no longer evaluated; pass 2 never clears acmod's flags) and
could be PROVED irrelevant -- not the per-stream normaliser's maximum, none of their senones the
frame's best -- from their top-1 score alone."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import soundswallower_b200 as ssb
from oracle.oracle import Oracle
import bench

g = np.load(bench.GOLDEN)
o = Oracle(bench.MODEL)
feats, chains = bench.make_config2_batch(g, 2)
arr = o.model_arrays()
sseq = np.asarray(arr["sseq"]); s2c = np.asarray(arr["sen2cb"]); mixw = np.asarray(arr["mixw"]); lut = np.asarray(arr["lut"]).astype(int)
print("mixw", mixw.shape, "sseq", sseq.shape)
n_feat, n_den, n_sen = mixw.shape
EPS = 12
tot = dict(frames=0, flagged_cb=0, live_cb=0, stale_cb=0, stale_ok_norm=0, stale_ok_both=0, stale_ok_both_m=0, norm_changed=0, best_changed=0)
# per (cb, f): smallest mixture weight any of the codebook's senones has for any density
mmin = np.zeros((int(s2c.max()) + 1, n_feat), int)
for c in range(mmin.shape[0]):
    sel = np.nonzero(s2c == c)[0]
    for f in range(n_feat):
        mmin[c, f] = mixw[f][:, sel].min() if len(sel) else 0
for feat, ch in zip(feats, chains):
    T = feat.shape[0]
    enter = ssb.plan_chain(T, ch["sf"], ch["ef"])
    cw, sc = o.topn_all(feat)           # [T][cb][f][4]
    top1 = sc[..., 0] >> 10             # [T][cb][f]
    top1hi = (sc[..., 0] + EPS) >> 10
    ssid, ef = ch["ssid"], ch["ef"]
    last = np.maximum(enter, np.minimum(ef, T - 1))
    def ascore(s, t, N):
        c = s2c[s]; a = 0
        for f in range(n_feat):
            v = [int(mixw[f][cw[t, c, f, k]][s]) + min(96, int(N[f]) - int(sc[t, c, f, k] >> 10)) for k in range(4)]
            fd = v[0]
            for k in range(1, 4):
                d = fd - v[k]
                fd = (v[k] if d > 0 else fd) - lut[min(abs(d), 255)]
            a += fd
        return a
    for t in range(0, T, 7):
        act = [i for i in range(len(ssid)) if 0 <= enter[i] <= t]
        live = [i for i in act if t <= last[i]]
        if not live:
            continue
        flagged_sen = np.unique(sseq[ssid[act]].ravel()); live_sen = np.unique(sseq[ssid[live]].ravel())
        fcb = np.unique(s2c[flagged_sen]); lcb = np.unique(s2c[live_sen])
        stale = np.setdiff1d(fcb, lcb)
        N_live = top1[t, lcb].max(0); N_all = top1[t, fcb].max(0)
        B_live = min(ascore(s, t, N_all) for s in live_sen)
        stale_sen = np.setdiff1d(flagged_sen, live_sen)
        B_all = min([B_live] + [ascore(s, t, N_all) for s in stale_sen])
        tot["frames"] += 1; tot["flagged_cb"] += len(fcb); tot["live_cb"] += len(lcb); tot["stale_cb"] += len(stale)
        tot["norm_changed"] += int((N_all != N_live).any()); tot["best_changed"] += int(B_all != B_live)
        # B_live as the kernel would know it: with the live normaliser (valid when every stale cb passes the norm test)
        for c in stale:
            okn = bool((top1hi[t, c] <= N_live).all())
            d = np.minimum(96, N_live - top1hi[t, c])
            lb = int(d.sum()) - 63
            lbm = int((d + mmin[c]).sum()) - 63
            tot["stale_ok_norm"] += okn
            tot["stale_ok_both"] += okn and lb >= B_live
            tot["stale_ok_both_m"] += okn and lbm >= B_live
print(tot)
print("stale fraction of scanned codebooks %.2f; provable (norm) %.2f, (norm+best) %.2f, with min-weight bound %.2f"
      % (tot["stale_cb"] / tot["flagged_cb"], tot["stale_ok_norm"] / max(1, tot["stale_cb"]),
         tot["stale_ok_both"] / max(1, tot["stale_cb"]), tot["stale_ok_both_m"] / max(1, tot["stale_cb"])))
