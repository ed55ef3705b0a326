#!/usr/bin/env python
"""At which `iter` does the zebra-ordered GPU smoother reach the accuracy of the reference's lexicographic one?

The early-linearisation drivers (FMG: one FAS cycle per level; Horn-Schunck: one solve per level) do not re-warp, so the
iterate after a FIXED number of sweeps decides the flow. Runs the reference-MEX pipeline (oracle/pipelines.py around
oracle/_ref) at the drivers' defaults and the GPU pipeline over a range of `iter`, on synthetic pairs with known flow,
and prints AEE against the ground truth + time per pair. Used to set pdegpu_flow_*_default_params.

    python tools/accuracy_match.py [fmg|hs] [nr nc]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pde-based-image-processing_b200"))
from pdegpu import lib, synth  # noqa: E402
from oracle import oracle as orc, pipelines  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "fmg"
nr, nc = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (480, 640)
be = orc.RefBackend() if orc.have_ref() else orc.OracleBackend()
ctx = lib.Context(0)
sl = (slice(8, -8), slice(8, -8))
out = {"driver": which, "nrows": nr, "ncols": nc, "pairs": []}
for seed in (300, 311):
    I0, I1, u, v = synth.image_pair(seed, nr, nc, nframes=1, scale=255.0, max_flow=0.8)     # both drivers take 0..255 frames (HS divides by 255 itself, FlowEminHS_elin_2D_v10.m:67)
    I0, I1 = I0.reshape(nr, nc, 1), I1.reshape(nr, nc, 1)

    def aee(U, V):
        return float(np.mean(np.sqrt((U[sl] - u[sl]) ** 2 + (V[sl] - v[sl]) ** 2)))

    t0 = time.perf_counter()
    Uo, Vo = (pipelines.flow_fmg if which == "fmg" else pipelines.flow_hs)(I0, I1, be)
    rec = {"seed": seed, "reference_default": {"aee": aee(Uo, Vo), "s": time.perf_counter() - t0}, "gpu": []}
    fn = ctx.flow_fmg if which == "fmg" else ctx.flow_hs
    for it in ((4, 6, 8, 12, 16, 24, 32, 48, 64) if which == "fmg" else (20, 30, 40, 60, 80, 120, 160, 240)):
        fn(I0, I1, iter=it)
        t0 = time.perf_counter()
        Ug, Vg = fn(I0, I1, iter=it)
        rec["gpu"].append({"iter": it, "aee": aee(Ug, Vg), "s": time.perf_counter() - t0,
                           "epe_vs_reference_default": float(np.mean(np.sqrt((Ug - Uo) ** 2 + (Vg - Vo) ** 2)))})
    out["pairs"].append(rec)
print(json.dumps(out))
ctx.close()
