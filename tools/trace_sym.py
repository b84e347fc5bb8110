"""Per-phase medians of the symmetric variant (VND_TM_SHAPE=15) from gpurun_out/tm_trace.npy: per warp (quarter, group)."""
import os
import numpy as np
t = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "tm_trace.npy")).astype(np.int64)
tiles = slice(8, 40)
med = lambda a: int(np.median(a))
for w in range(16):
    q, g = w & 3, w >> 2
    st, fl, wt, nd, fd, sf, sg, en = (t[w, tiles, e] for e in (0, 6, 1, 2, 3, 5, 4, 7))
    print(f"q{q} g{g}: period {med(np.diff(t[w, 8:41, 0]))}  fill(+waits) {med(fl - st)}  wait others' fill {med(wt - fl)}  tmem {med(nd - wt)}  far {med(fd - nd)}  stfree {med(sf - fd)}  stage {med(sg - sf)}  store/load duty {med(en - sg)}")
