"""Time pdegpu_dev_flow_fmg_2d (1080p pairs, library default order = the reference's) against the number of pairs that run
side by side on lanes.  PDEGPU_LANES=128 python tools/fmg_lanes.py PAIRS [reps] [order]"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pde-based-image-processing_b200"))
import numpy as np
import torch
from pdegpu import lib, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
order = {"auto": lib.ORDER_AUTO, "fast": lib.ORDER_FAST, "reference": lib.ORDER_REFERENCE}[sys.argv[3] if len(sys.argv) > 3 else "auto"]
nr, nc, C = 1080, 1920, 1
I0, I1, u, v = synth.image_pair(300, nr, nc, nframes=C, scale=255.0, max_flow=0.8)
a0 = torch.from_numpy(np.stack([I0.reshape(-1, order="F")] * B)).cuda()
a1 = torch.from_numpy(np.stack([I1.reshape(-1, order="F")] * B)).cuda()
U = torch.empty(B, nr * nc, device="cuda"); V = torch.empty(B, nr * nc, device="cuda")
ctx = lib.Context(0)
ctx.set_sweep_order(order)
L = lib.dll()
p = lib.FlowFmgParams(); L.pdegpu_flow_fmg_default_params(ctypes.byref(p))
def run():
    ctx._chk(L.pdegpu_dev_flow_fmg_2d(ctx.h, U.data_ptr(), V.data_ptr(), a0.data_ptr(), a1.data_ptr(), nr, nc, C, B, ctypes.byref(p)))
for _ in range(3):                 # direct run, graph capture, first replay
    run()
ctx.sync()
t0 = time.perf_counter()
for _ in range(reps):
    run()
ctx.sync()
dt = (time.perf_counter() - t0) / reps
s = (slice(8, -8), slice(8, -8))
aee = []
for b in (0, B - 1):
    Uh = U[b].cpu().numpy().reshape(nr, nc, order="F"); Vh = V[b].cpu().numpy().reshape(nr, nc, order="F")
    aee.append(float(np.mean(np.sqrt((Uh[s] - u[s]) ** 2 + (Vh[s] - v[s]) ** 2))))
print("pairs %d lanes %s maxconn %s order %s: %.1f ms per call, %.2f flows/s, AEE first/last pair %.5f %.5f, mem %.1f GB" % (
    B, os.environ.get("PDEGPU_LANES", "32"), os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS", "-"), sys.argv[3] if len(sys.argv) > 3 else "auto",
    dt * 1e3, B / dt, aee[0], aee[1], torch.cuda.mem_get_info()[0] and (torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9))
