import os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pde-based-image-processing_b200"))
from oracle import pipelines, oracle as o
from pdegpu import synth, lib
ctx = lib.Context(0)
be = o.RefBackend() if o.have_ref() else o.OracleBackend()
def epe(a, b, c, d): return float(np.mean(np.sqrt((a.astype(np.float64) - c) ** 2 + (b.astype(np.float64) - d) ** 2)))
nr, nc = 64, 80
for C in (1, 3):
    I0, I1, u, v = synth.image_pair(31, nr, nc, nframes=C, scale=255.0, max_flow=0.8)
    I0 = I0.reshape(nr, nc, C); I1 = I1.reshape(nr, nc, C)
    for kw in [dict(iter=1600, omega=1.8, alpha=0.02), dict(iter=1600, omega=1.8, alpha=0.02, max_scales=1), dict(iter=1600, omega=1.8, alpha=0.02, max_scales=2),
               dict(iter=3000, omega=1.0, alpha=0.02, solver=1), dict(iter=1600, omega=1.8, alpha=0.002), dict(iter=6400, omega=1.5, alpha=0.02)]:
        Ug, Vg = ctx.flow_hs(I0, I1, **kw)
        Uo, Vo = pipelines.flow_hs(I0, I1, be, **kw)
        d = np.sqrt((Ug - Uo) ** 2 + (Vg - Vo) ** 2)
        print(C, kw, "EPE %.2e max %.2e at %s  meanU gpu %.4f ref %.4f" % (epe(Ug, Vg, Uo, Vo), d.max(), np.unravel_index(d.argmax(), d.shape), Ug.mean(), Uo.mean()), flush=True)
