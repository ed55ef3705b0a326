#!/usr/bin/env python
"""Phase probes of tline_pass_kernel (needs a library built with -DTL_PROBE):

    PDEGPU_BUILD_SUFFIX=_probe PDEGPU_BUILD_NAME=libpdegpu_probe.so PDEGPU_NVCC_EXTRA=-DTL_PROBE python pde-based-image-processing_b200/build.py
    PDEGPU_LIB=pde-based-image-processing_b200/libpdegpu_probe.so python tools/tl_probe.py [batch nr nc]

Prints, per warp role, the share of its cycles spent in each phase (summed over all warps of all CTAs) per direction pass.
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
out = (ctypes.c_ulonglong * 32)()
ctx.relax(sysd, 1, 1.9, 2)
ctx.sync()
L.pdegpu_debug_tl_probe(out)            # discard the warm-up call
ITER = 2
ctx.relax(sysd, ITER, 1.9, 2)
ctx.sync()
L.pdegpu_debug_tl_probe(out)
v = [int(x) for x in out]
npass = 2 * ITER
cons = ["wait slab", "wait ring", "wait even nbrs", "rows (both unknowns)", "solves", "relax + store", "-", "publish", "block write / loop"]
tot = v[15] or 1
print(f"consumers: {v[14]} tasks, {tot / max(1, v[14]):.0f} warp-cycles per task incl. waits (all passes); env={ {k: x for k, x in os.environ.items() if k.startswith('PDEGPU_TL')} }")
for k, name in enumerate(cons):
    print(f"  {name:22s} {100.0 * v[k] / tot:5.1f} %   {v[k] / max(1, v[14]):8.0f} cycles/task")
print(f"producer: {v[22]} tasks; wait empty {100.0 * v[16] / max(1, v[23]):.1f} %  ({v[16] / max(1, v[22]):.0f} cyc/task), issue {100.0 * v[17] / max(1, v[23]):.1f} % ({v[17] / max(1, v[22]):.0f} cyc/task)")
print(f"loader:   {v[30]} lines; wait free {100.0 * v[24] / max(1, v[31]):.1f} %  ({v[24] / max(1, v[30]):.0f} cyc/line), issue {100.0 * v[25] / max(1, v[31]):.1f} % ({v[25] / max(1, v[30]):.0f} cyc/line)")
print(f"cycles per CTA and pass (producer total / passes / CTAs): {v[23] / npass / 148:.0f}")
ctx.close()
