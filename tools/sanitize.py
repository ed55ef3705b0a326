#!/usr/bin/env python
"""Small invocations of every kernel added in round 2, for compute-sanitizer --tool memcheck (tools/gpu_r3b.sh)."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pde-based-image-processing_b200"))
import torch  # noqa: E402
from pdegpu import bands, lib, mex, synth  # noqa: E402

os.environ["PDEGPU_POINT_WINDOW"] = "1"
gpu = mex.GpuBackend()
for order in ("fast", "reference"):
    os.environ["PDEGPU_ORDER"] = order
    for nr, nc in ((37, 53), (120, 164), (203, 270), (48, 64)):
        for fn, s, nout in (("Oflow_sor_llin4_2d", synth.flow_system(1, nr, nc, late=True), 2), ("Oflow_sor_elin4_2d", synth.flow_system(2, nr, nc), 2),
                            ("Oflow_sor_llin8_2d", synth.flow_system(3, nr, nc, late=True, eight=True), 2), ("Disp_sor_llin4_2d", synth.disp_system(4, nr, nc), 1),
                            ("PDEsolver4", synth.pde_system(5, nr, nc, nframes=2), 1), ("PDEsolver8", synth.pde_system(6, nr, nc, nframes=2, eight=True), 1)):
            for solver in (1, 2):
                out = gpu.call(fn, synth.mex_args(fn, s, 4, 1.5, solver), nout)
                assert all(np.isfinite(o).all() for o in out), (fn, order, nr, nc, solver)
    print("sweeps", order, "ok", flush=True)
os.environ["PDEGPU_ORDER"] = "fast"
# long lines: segments (zebra) and whole lines (reference order)
s = synth.flow_system(7, 1080, 900, late=True)
for order in ("fast", "reference"):
    os.environ["PDEGPU_ORDER"] = order
    gpu.call("Oflow_sor_llin4_2d", synth.mex_args("Oflow_sor_llin4_2d", s, 1, 1.5, 2), 2)
os.environ["PDEGPU_ORDER"] = "fast"
gpu.call("Oflow_sor_llin4_2d", synth.mex_args("Oflow_sor_llin4_2d", s, 4, 1.0, 1), 2)       # window kernel, strips
print("long lines ok", flush=True)
# bands in one process
KEYS = ("U", "V", "dU", "dV", "M", "Cu", "Cv", "Du", "Dv", "wW", "wN", "wE", "wS")
nr, nc, world, T = 128, 192, 3, 2
s = synth.flow_system(77, nr, nc, late=True)
bs = []
for r in range(world):
    plan = bands.BandPlan(nr, nc, r, world, sweeps_per_exchange=T)
    f = {k: torch.from_numpy(np.ascontiguousarray(plan.take_local(np.ascontiguousarray(s[k].T)))).cuda(0) for k in KEYS}
    bs.append(bands.GpuBand(lib.Context(0), plan, lib.FLOW_LLIN4, f, transport="p2p"))
torch.cuda.synchronize()
for r, b in enumerate(bs):
    b.connect_local(bs[r - 1] if r > 0 else None, bs[r + 1] if r < world - 1 else None)
for _ in range(2):
    for b in bs:
        b.exchange_p2p()
    for b in bs:
        b.ctx.relax(b.sys, T, 1.0, 1)
for b in bs:
    b.ctx.sync()
print("bands ok", flush=True)
# pipelines with lanes + fused inner solve
c = lib.Context(0)
ps = [synth.image_pair(30 + k, 72, 88, nframes=1, scale=255.0, max_flow=0.8) for k in range(3)]
I0 = np.stack([p[0].reshape(72, 88, 1) for p in ps]); I1 = np.stack([p[1].reshape(72, 88, 1) for p in ps])
c.flow_fmg(I0, I1); c.flow_hs(I0, I1)
P0, P1, _, _ = synth.image_pair(11, 96, 128, nframes=3, scale=255.0, max_flow=2.0)
os.environ["PDEGPU_FUSE"] = "1"
c.flow_llin(P0, P1)
print("pipelines ok", flush=True)
