import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
import soundswallower_b200 as ssb
g = np.load('tests/golden/align_en-us.npz')
hmm = 'soundswallower_b200/model/en-us'
feat = g['feat']
res = {}
for k1 in ('tc2', 'ft'):
    os.environ['SSB_K1'] = k1
    model = ssb.AcousticModel(hmm, device=0)
    lex = ssb.Lexicon(model, hmmdir=hmm)
    graph = lex.align_graph("go forward ten meters")
    p1 = ssb.fsg_batch(model, [feat], [graph], want_hist=True, compallsen=False)[0]
    res[k1] = p1
    print(k1, p1['rv'], p1['hyp_score'], p1['n_sen_eval'], len(p1['hist']))
a, b = res['tc2'], res['ft']
print('hist equal', np.array_equal(a['hist'], b['hist']))
if not np.array_equal(a['hist'], b['hist']):
    n = min(len(a['hist']), len(b['hist']))
    d = np.nonzero((a['hist'][:n] != b['hist'][:n]).any(1))[0]
    print('first diff rows', d[:5]); print(a['hist'][d[0]], b['hist'][d[0]])
