#!/usr/bin/env python
"""Phase probes of alr_window2_kernel on the bench workload (needs a library built with -DW2_PROBE):

    PDEGPU_BUILD_SUFFIX=_probe PDEGPU_BUILD_NAME=libpdegpu_probe.so PDEGPU_NVCC_EXTRA=-DW2_PROBE python pde-based-image-processing_b200/build.py
    PDEGPU_LIB=pde-based-image-processing_b200/libpdegpu_probe.so python tools/w2_probe.py [batch]

Prints, per warp role, the share of a warp's cycles spent in each phase (summed over all warps of all CTAs).
"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pde-based-image-processing_b200"))
import torch  # noqa: E402
from pdegpu import lib, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
NR, NC = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (480, 640)
dev = torch.device("cuda", 0)
ctx = lib.Context(0)
keys = ("U", "V", "dU", "dV", "M", "Cu", "Cv", "Du", "Dv", "wW", "wN", "wE", "wS")
base = [synth.flow_system(1235 + k, NR, NC, late=True) for k in range(2)]
d = {k: torch.from_numpy(np.stack([base[b % 2][k].reshape(-1, order="F") for b in range(B)])).to(dev) for k in keys}
sysd = lib.make_system(lib.FLOW_LLIN4, NR, NC, batch=B, batch_stride=NR * NC,
                       x=(d["dU"].data_ptr(), d["dV"].data_ptr()), x0=(d["U"].data_ptr(), d["V"].data_ptr()),
                       m=d["M"].data_ptr(), c=(d["Cu"].data_ptr(), d["Cv"].data_ptr()), d=(d["Du"].data_ptr(), d["Dv"].data_ptr()),
                       w=[d[k].data_ptr() for k in ("wW", "wN", "wE", "wS")])
L = lib.dll()
out = (ctypes.c_ulonglong * 16)()
GEN3 = False
probe_fn = getattr(L, "pdegpu_debug_w2_probe", None)
have = probe_fn is not None
SOLVER = int(os.environ.get('W2_SOLVER', '2'))
ctx.relax(sysd, 1, 1.9, SOLVER)
ctx.sync()
if have:
    probe_fn(out)            # discard the warm-up pass
ctx.profile(True)
ctx.relax(sysd, 4, 1.9, SOLVER)
ctx.sync()
if have:
    probe_fn(out)
v = [int(x) for x in out]
an = ("other", "wait freed buffer", "wait ring slot", "load issue", "wait solved (odd)", "wait loads", "rows + st.shared")
sn = ("other", "wait filled buffer", "pick-up + solve", "ring write (+ wait odd readers)", "stores to X_out") if GEN3 else ("other", "wait filled buffer", "pick-up + solve + ring", "block write-out")
if not have:
    v = [0] * 16
print("assembler warps: total cycles", v[7])
for k, name in enumerate(an):
    print("   %-26s %5.1f %%" % (name, 100.0 * v[k] / max(v[7], 1)))
print("solver warps: total cycles", v[15])
for k, name in enumerate(sn):
    print("   %-26s %5.1f %%" % (name, 100.0 * v[8 + k] / max(v[15], 1)))
rep = ctx.profile_report()
for k in (rep["kernels"] if isinstance(rep, dict) and "kernels" in rep else rep):
    print(k)
