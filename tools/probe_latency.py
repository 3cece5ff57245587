import time, os, sys
sys.path.insert(0,'/root/repo')
import soundswallower_b200 as ssb
d=ssb.Decoder('/root/repo/soundswallower_b200/model/en-us')
d.set_align_text("go forward ten meters")
f='/root/repo/tests/data/goforward.raw'
for i in range(3): d.decode_file(f); d.dumps(align_level=1)
ts=[]
for i in range(20):
    t0=time.perf_counter(); d.decode_file(f); j=d.dumps(align_level=1); ts.append(time.perf_counter()-t0)
ts.sort(); print("single utterance (2.8 s audio) decode_file + dumps(align_level=1): median %.1f ms, min %.1f ms" % (1e3*ts[len(ts)//2], 1e3*ts[0]))
