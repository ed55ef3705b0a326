"""Residual norm after k ALR iterations at omega = 1.9: GPU zebra ordering against the reference's lexicographic ordering
(same linear system, same start)."""
import os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pde-based-image-processing_b200"))
from oracle import oracle as o
from pdegpu import synth, mex
be = o.RefBackend() if o.have_ref() else o.OracleBackend()
F32 = np.float32
nr, nc = int(sys.argv[1]) if len(sys.argv) > 1 else 120, int(sys.argv[2]) if len(sys.argv) > 2 else 160
s = synth.flow_system(77, nr, nc, late=False)
names = ("U", "V", "M", "Cu", "Cv", "Du", "Dv", "wW", "wN", "wE", "wS")
A = [np.asarray(s[k], dtype=F32).reshape(nr, nc) for k in names]
def resid(U, V, call):
    out = call("Oflow_sor_elin4_2d", [U, V] + A[2:] + [F32(0), F32(1.9), F32(2)], 4)
    return float(np.sqrt(np.mean(out[2].astype(np.float64) ** 2 + out[3].astype(np.float64) ** 2)))
gpu = mex.GpuBackend()
print("start residual", resid(A[0], A[1], be.call))
for om in (1.9, 1.5, 1.0):
    for solver in (2, 1):
        for it in (1, 2, 4, 8, 16, 64):
            Ur, Vr = be.call("Oflow_sor_elin4_2d", A + [F32(it), F32(om), F32(solver)], 2)
            Ug, Vg = gpu.call("Oflow_sor_elin4_2d", A + [F32(it), F32(om), F32(solver)], 2)
            print(f"omega {om} solver {solver} iter {it:3d}: residual ref {resid(Ur, Vr, be.call):.3e}  gpu {resid(Ug, Vg, be.call):.3e}")
