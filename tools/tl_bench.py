#!/usr/bin/env python
"""Micro-benchmark / cross-check of one relax call (pdegpu_dev_relax) on synthetic systems.

    python tools/tl_bench.py [--fam llin4] [--nr 480 --nc 640] [--batch 64] [--iter 4] [--solver 2] [--check] [--reps 5]

Prints one JSON line: per-kernel CUDA-event times (the library's profile), Mpix*iter/s of the call, and with --check the
difference between the streaming kernels (kernel path 1) and generation 0 (kernel path 0) after the same call.
Geometry knobs of the generation-3 line kernel are environment variables read by the library (PDEGPU_TL_*,
PDEGPU_ALR_GEN), so a sweep runs this script once per setting.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pde-based-image-processing_b200"))
import torch  # noqa: E402
from pdegpu import lib, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--fam", default="llin4", choices=["elin4", "llin4", "llin8", "disp", "pde4", "pde8"])
ap.add_argument("--nr", type=int, default=480)
ap.add_argument("--nc", type=int, default=640)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iter", type=int, default=4)
ap.add_argument("--omega", type=float, default=1.9)
ap.add_argument("--solver", type=int, default=2)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--check", action="store_true")
ap.add_argument("--resid", action="store_true", help="flow families: reference-defined residual norm before / after the call")
ap.add_argument("--tag", default="")
a = ap.parse_args()

dev = torch.device("cuda", 0)
ctx = lib.Context(0)
NR, NC, B = a.nr, a.nc, a.batch
n = NR * NC
nd = min(B, 3)


def stack(systems, k):
    return torch.from_numpy(np.stack([systems[b % nd][k].reshape(-1, order="F") for b in range(B)])).to(dev)


if a.fam in ("elin4", "llin4", "llin8"):
    late, eight = a.fam != "elin4", a.fam == "llin8"
    base = [synth.flow_system(1235 + k, NR, NC, late=late, eight=eight) for k in range(nd)]
    wk = ("wW", "wN", "wE", "wS") + (("wNW", "wNE", "wSE", "wSW") if eight else ())
    d = {k: stack(base, k) for k in base[0]}
    unk = ("dU", "dV") if late else ("U", "V")
    fam = {"elin4": lib.FLOW_ELIN4, "llin4": lib.FLOW_LLIN4, "llin8": lib.FLOW_LLIN8}[a.fam]
    bpp = {"elin4": 52.0, "llin4": 60.0, "llin8": 76.0}[a.fam]

    def mk(x):
        return lib.make_system(fam, NR, NC, batch=B, batch_stride=n, x=(x[0].data_ptr(), x[1].data_ptr()),
                               x0=(d["U"].data_ptr(), d["V"].data_ptr()) if late else (), m=d["M"].data_ptr(),
                               c=(d["Cu"].data_ptr(), d["Cv"].data_ptr()), d=(d["Du"].data_ptr(), d["Dv"].data_ptr()),
                               w=[d[k].data_ptr() for k in wk])
    init = [d[unk[0]], d[unk[1]]]
elif a.fam == "disp":
    base = [synth.disp_system(44 + k, NR, NC) for k in range(nd)]
    d = {k: stack(base, k) for k in base[0]}
    bpp = 36.0

    def mk(x):
        return lib.make_system(lib.DISP_LLIN4, NR, NC, batch=B, batch_stride=n, x=(x[0].data_ptr(),), x0=(d["U"].data_ptr(),),
                               c=(d["Cu"].data_ptr(),), d=(d["Du"].data_ptr(),), w=[d[k].data_ptr() for k in ("wW", "wN", "wE", "wS")])
    init = [d["dU"]]
else:
    eight = a.fam == "pde8"
    base = [synth.pde_system(45 + k, NR, NC, nframes=1, eight=eight) for k in range(nd)]
    d = {k: stack(base, k) for k in base[0]}
    wk = ("wW", "wN", "wE", "wS") + (("wNW", "wNE", "wSE", "wSW") if eight else ())
    bpp = 48.0 if eight else 32.0

    def mk(x):
        return lib.make_system(lib.PDE8 if eight else lib.PDE4, NR, NC, batch=B, batch_stride=n, x=(x[0].data_ptr(),),
                               c=(d["B"].data_ptr(),), d=(d["TRACE"].data_ptr(),), w=[d[k].data_ptr() for k in wk])
    init = [d["X"]]

out = {"tag": a.tag, "fam": a.fam, "nr": NR, "nc": NC, "batch": B, "iter": a.iter, "solver": a.solver,
       "env": {k: v for k, v in os.environ.items() if k.startswith("PDEGPU_")}}

if a.check:
    res = []
    for path in (0, 1):
        ctx.set_kernel_path(path)
        x = [t.clone() for t in init]
        torch.cuda.synchronize()
        ctx.relax(mk(x), a.iter, a.omega, a.solver)
        ctx.sync()
        res.append([t.cpu().numpy() for t in x])
    errs = []
    for p, q in zip(res[0], res[1]):
        den = float(np.max(np.abs(p))) or 1.0
        errs.append(float(np.nanmax(np.abs(p - q))) / den)
    out["check_rel_err_vs_gen0"] = errs
    out["finite"] = bool(all(np.isfinite(q).all() for q in res[1]))
    ctx.set_kernel_path(1)

if a.resid and a.fam in ("elin4", "llin4"):
    def resnorm(xs):
        RU = torch.empty(B * n, device=dev); RV = torch.empty(B * n, device=dev)
        ctx.residual(mk(xs), 1, RU.data_ptr(), RV.data_ptr())
        ctx.sync()
        return float(torch.sqrt(torch.nanmean(RU.double() ** 2 + RV.double() ** 2)))
    xs = [t.clone() for t in init]
    r = [resnorm(xs)]
    for _ in range(3):
        ctx.relax(mk(xs), a.iter, a.omega, a.solver)
        ctx.sync()
        r.append(resnorm(xs))
    out["residual_norms_per_call"] = r

x = [t.clone() for t in init]
sysd = mk(x)
for _ in range(2):
    ctx.relax(sysd, a.iter, a.omega, a.solver)
ctx.sync()
for t, t0 in zip(x, init):
    t.copy_(t0)
torch.cuda.synchronize()
ctx.profile(True)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(stream)
for _ in range(a.reps):
    ctx.relax(sysd, a.iter, a.omega, a.solver)
ev1.record(stream)
ctx.sync()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / a.reps
prof = ctx.profile_report()
ctx.profile(False)
out["ms_per_call"] = ms
out["mpix_iter_s"] = B * n * a.iter / 1e6 / (ms / 1e3)
out["kernels"] = [{"kernel": p["kernel"], "launches": p["launches"], "ms_avg": p["ms_total"] / p["launches"],
                   "GBs_algorithmic": (p["bytes_total"] / p["launches"]) / (p["ms_total"] / p["launches"] * 1e-3) / 1e9 if p["bytes_total"] else None}
                  for p in prof]
print(json.dumps(out))
ctx.close()
