#!/usr/bin/env python
"""Symmetric stereo driver in the reference's sweep order: GPU pipeline against the restated driver on the reference MEX
code, at several sizes / disparity ranges (diagnostic for tests/test_gpu_fullsize.py)."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pde-based-image-processing_b200"))
from pdegpu import lib, synth  # noqa: E402
from oracle import oracle as orc, pipelines  # noqa: E402

be = orc.RefBackend() if orc.have_ref() else orc.OracleBackend()
ctx = lib.Context(0)
ctx.set_sweep_order(lib.ORDER_REFERENCE)
for nr, nc, mf, u8 in ((96, 128, 3.0, True), (135, 256, 4.0 / 0.55, True), (270, 512, 4.0 / 0.55, True), (270, 512, 4.0 / 0.55, False), (540, 1024, 8.0 / 0.55, True), (540, 1024, 3.0, True)):
    Il, Ir, u, _ = synth.image_pair(303, nr, nc, nframes=1, scale=255.0, max_flow=mf, horizontal=True)
    for ms in (None, 2):
        kw = dict(uint8_input=u8) if ms is None else dict(uint8_input=u8, max_scales=ms)
        U0, U1 = ctx.disp_sym(Il, Ir, **{k: (int(v) if isinstance(v, bool) else v) for k, v in kw.items()})
        O0, O1 = pipelines.disp_sym(Il, Ir, be, **kw)
        s = (slice(16, -16), slice(16, -16))
        print(f"{nr}x{nc} max|u| {np.abs(u).max():.1f} u8={u8} max_scales={ms}: nan equal {np.array_equal(np.isnan(U0), np.isnan(O0))}, "
              f"mean|dU0| {np.nanmean(np.abs(U0 - O0)):.2e} median {np.nanmedian(np.abs(U0 - O0)):.2e} max {np.nanmax(np.abs(U0 - O0)):.2e}; "
              f"err vs truth gpu {np.nanmean(np.abs(U0[s] - u[s])):.3f} ref {np.nanmean(np.abs(O0[s] - u[s])):.3f}", flush=True)
