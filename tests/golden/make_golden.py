#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref, compiled from
/root/reference by `make -C oracle ref`) on seeded synthetic inputs.

    python tests/golden/make_golden.py

Run where /root/reference exists. Each file stores the inputs (so the vectors do not depend on
the generator staying unchanged) and the reference's outputs. The reference ships no golden
vectors or tests of its own, so these are the known-answer tests.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pde-based-image-processing_b200"))
from oracle.oracle import RefBackend  # noqa: E402
from pdegpu import synth  # noqa: E402

NR, NC = 23, 31


def cases():
    out = []
    for solver in (1, 2):
        for it in (0, 3):
            s = synth.flow_system(101, NR, NC, nframes=2)
            out.append((f"elin4_s{solver}_it{it}", "Oflow_sor_elin4_2d", synth.mex_args("Oflow_sor_elin4_2d", s, it, 1.9, solver), 4))
            s = synth.flow_system(102, NR, NC, late=True, nframes=2)
            out.append((f"llin4_s{solver}_it{it}", "Oflow_sor_llin4_2d", synth.mex_args("Oflow_sor_llin4_2d", s, it, 1.9, solver), 4))
        s = synth.flow_system(103, NR, NC, late=True, eight=True)
        out.append((f"llin8_s{solver}", "Oflow_sor_llin8_2d", synth.mex_args("Oflow_sor_llin8_2d", s, 3, 1.9, solver), 2))
        s = synth.disp_system(104, NR, NC)
        out.append((f"disp_s{solver}", "Disp_sor_llin4_2d", synth.mex_args("Disp_sor_llin4_2d", s, 3, 1.9, solver), 2))
        ss = {"f0": synth.disp_system(105, NR, NC), "f1": synth.disp_system(106, NR, NC)}
        out.append((f"dispsym_s{solver}", "Disp_sor_llin_sym4_2d", synth.mex_args("Disp_sor_llin_sym4_2d", ss, 3, 1.9, solver), 2))
        for eight in (False, True):
            fn = "PDEsolver8" if eight else "PDEsolver4"
            s = synth.pde_system(107, NR, NC, nframes=2, eight=eight)
            out.append((f"{fn}_s{solver}", fn, synth.mex_args(fn, s, 3, 1.75, solver), 1))
    s = synth.flow_system(108, NR, NC, nframes=2)
    out.append(("lhs_elin4", "Oflow_lhs_elin4_2d", synth.mex_args("Oflow_lhs_elin4_2d", s), 2))
    s = synth.flow_system(109, NR, NC, late=True, nframes=2)
    out.append(("lhs_llin4", "Oflow_lhs_llin4_2d", synth.mex_args("Oflow_lhs_llin4_2d", s), 2))
    I0, I1, _, _ = synth.image_pair(110, NR, NC, nframes=2)
    out.append(("fst", "FstDerivatives5", [I0, I1], 3))
    out.append(("snd", "SndDerivatives5", [I0, I1], 5))
    out.append(("ddiff", "DdiffWeights", [synth.f32(I0 * 10), synth.f32([[1e-3]])], 4))
    return out


def main():
    R = RefBackend()
    for name, fn, args, nlhs in cases():
        outs = R.call(fn, args, nlhs)
        d = {"fn": np.array(fn), "nlhs": np.array(nlhs)}
        for k, a in enumerate(args):
            d[f"in{k}"] = np.asarray(a)
        for k, o in enumerate(outs):
            d[f"out{k}"] = o
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    # warp: through the explicit 5-argument prototype, both out-of-image conventions
    I0, _, u, v = synth.image_pair(111, NR, NC, nframes=2)
    X, Y = np.meshgrid(np.arange(1, NC + 1, dtype=np.float32), np.arange(1, NR + 1, dtype=np.float32))
    X = synth.f32(X + 4 * u)
    Y = synth.f32(Y + 4 * v)
    X[3, 4] = np.nan
    Y[5, 6] = -5e9
    X[9, 9] = NC
    Y[9, 9] = NR
    np.savez_compressed(os.path.join(HERE, "warp.npz"), I=I0, X=X, Y=Y,
                        out_nan=R.bilin(I0, X, Y, float("nan")), out_zero=R.bilin(I0, X, Y, 0.0))
    print("wrote", len(cases()) + 1, "golden files to", HERE)


if __name__ == "__main__":
    main()
