#!/usr/bin/env python
"""Urban3 crop, Horn-Schunck with (nearly) converged solves: GPU (zebra order, several iteration counts; reference
order) against the restated driver on the reference MEX code. Diagnostic for tests/test_gpu_configs_fixtures.py."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pde-based-image-processing_b200"))
from pdegpu import lib  # noqa: E402
from oracle import oracle as orc, pipelines  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "urban3_pair.npz"))
I0, I1 = z["frame07"].astype(np.float32), z["frame08"].astype(np.float32)
c = (slice(160, 320), slice(216, 424))
I0, I1 = np.ascontiguousarray(I0[c]), np.ascontiguousarray(I1[c])
be = orc.RefBackend() if orc.have_ref() else orc.OracleBackend()
epe = lambda a, b, c_, d: float(np.mean(np.sqrt((a.astype(np.float64) - c_) ** 2 + (b.astype(np.float64) - d) ** 2)))
ctx = lib.Context(0)
for alpha, omega in ((0.002, 1.8), (0.02, 1.8)):
    ref = {}
    for it in (400, 1600):
        ref[it] = pipelines.flow_hs(I0, I1, be, iter=it, omega=omega, alpha=alpha)
    print(f"alpha {alpha}: reference 400 vs 1600 iterations: {epe(*ref[400], *ref[1600]):.2e} px")
    Uo, Vo = ref[1600]
    for order in (lib.ORDER_FAST, lib.ORDER_REFERENCE):
        ctx.set_sweep_order(order)
        for it in ((400, 1600, 6400, 25600) if order == lib.ORDER_FAST else (400, 1600)):
            Ug, Vg = ctx.flow_hs(I0, I1, iter=it, omega=omega, alpha=alpha)
            print(f"  order {order} iter {it}: EPE vs reference(1600) {epe(Ug, Vg, Uo, Vo):.3e} px, vs reference({min(it, 1600)}) {epe(Ug, Vg, *ref[min(it, 1600)]):.3e}")
