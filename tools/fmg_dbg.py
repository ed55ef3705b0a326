import os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pde-based-image-processing_b200"))
from oracle import pipelines, oracle as o
from pdegpu import synth, lib
ctx = lib.Context(0)
be = o.RefBackend() if o.have_ref() else o.OracleBackend()
def epe(a, b, c, d): return float(np.mean(np.sqrt((a.astype(np.float64) - c) ** 2 + (b.astype(np.float64) - d) ** 2)))
import time
for (nr, nc, C, seed) in [(120, 160, 1, 5), (96, 128, 3, 3), (480, 640, 3, 5)]:
    I0, I1, u, v = synth.image_pair(seed, nr, nc, nframes=C, scale=255.0, max_flow=0.8)
    I0 = I0.reshape(nr, nc, C); I1 = I1.reshape(nr, nc, C)
    for kw in [dict(), dict(omega=1.0), dict(omega=1.2), dict(omega=1.4), dict(omega=1.6), dict(iter=300, omega=1.3, firstLoop=2, max_scales=4)]:
        t0 = time.time(); Ug, Vg = ctx.flow_fmg(I0, I1, **kw); t1 = time.time()
        Uo, Vo = pipelines.flow_fmg(I0, I1, be, **({} if 'max_scales' not in kw else kw)); t2 = time.time()
        m = 8; sl = (slice(m, nr - m), slice(m, nc - m))
        print((nr, nc, C), kw, "EPE gpu-ref %.4f  AEE gpu %.4f  AEE ref %.4f   (gpu %.2fs ref %.2fs)" % (epe(Ug, Vg, Uo, Vo), epe(Ug[sl], Vg[sl], u[sl], v[sl]), epe(Uo[sl], Vo[sl], u[sl], v[sl]), t1 - t0, t2 - t1), flush=True)
