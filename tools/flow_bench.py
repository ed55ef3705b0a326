"""Time the device-resident flow pipeline (pdegpu_dev_flow_llin_2d) on synthetic 640x480 pairs and print
the per-kernel breakdown.  python tools/flow_bench.py [batch] [reps]"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pde-based-image-processing_b200"))
import numpy as np
import torch
from pdegpu import lib, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
nr, nc, C = 480, 640, 3
I0, I1, u, v = synth.image_pair(1, nr, nc, nframes=C, scale=255.0, max_flow=3.0)
a0 = torch.from_numpy(np.stack([I0.reshape(-1, order="F")] * B)).cuda()
a1 = torch.from_numpy(np.stack([I1.reshape(-1, order="F")] * B)).cuda()
U = torch.empty(B, nr * nc, device="cuda"); V = torch.empty(B, nr * nc, device="cuda")
ctx = lib.Context(0)
L = lib.dll()
p = lib.FlowLlinParams(); L.pdegpu_flow_llin_default_params(ctypes.byref(p))
def run():
    ctx._chk(L.pdegpu_dev_flow_llin_2d(ctx.h, U.data_ptr(), V.data_ptr(), a0.data_ptr(), a1.data_ptr(), nr, nc, C, B, ctypes.byref(p)))
for _ in range(3):      # direct run, graph capture, first replay
    run()
ctx.sync()
n0 = ctx.launches
t0 = time.perf_counter()
for _ in range(reps):
    run()
ctx.sync()
dt = (time.perf_counter() - t0) / reps
print("batch %d: %.1f ms per batch, %.1f flows/s, %d launches per batch" % (B, dt * 1e3, B / dt, (ctx.launches - n0) // reps))
Uh = U[0].cpu().numpy().reshape(nr, nc, order="F"); Vh = V[0].cpu().numpy().reshape(nr, nc, order="F")
s = (slice(8, -8), slice(8, -8))
print("AEE vs ground truth %.4f px" % float(np.mean(np.sqrt((Uh[s] - u[s]) ** 2 + (Vh[s] - v[s]) ** 2))))
ctx.profile(True); run(); ctx.sync()
rep = sorted(ctx.profile_report(), key=lambda k: -k["ms_total"])
tot = sum(k["ms_total"] for k in rep)
for k in rep[:12]:
    print("  %-40s %5d launches %8.2f ms %5.1f%%  %7.0f GB/s" % (k["kernel"], k["launches"], k["ms_total"], 100 * k["ms_total"] / tot,
          k["bytes_total"] / k["ms_total"] / 1e6 if k["ms_total"] > 0 else 0))
print("  sum of kernel times %.1f ms" % tot)
