"""Pins the oracle: the plain-C restatement (oracle/pde_oracle.c) against the UNMODIFIED reference
compiled from /root/reference (oracle/_ref). Bit-exact everywhere except the 8-neighbour flow line
solver (summation order differs from border case to border case in the reference; <= 2e-6)."""
import numpy as np
import pytest

from pdegpu import synth
from util import assert_bitwise, rel_err

SHAPES = [(37, 53), (16, 9), (5, 5)]


def both(oracle, ref, fn, args, nlhs):
    return oracle.call(fn, args, nlhs), ref.call(fn, args, nlhs)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("solver", [1, 2])
@pytest.mark.parametrize("it", [0, 1, 4])
def test_flow_elin4(oracle, ref, shape, solver, it):
    s = synth.flow_system(11, *shape, nframes=3)
    a, b = both(oracle, ref, "Oflow_sor_elin4_2d", synth.mex_args("Oflow_sor_elin4_2d", s, it, 1.9, solver), 4)
    for k, (x, y) in enumerate(zip(a, b)):
        assert_bitwise(x, y, f"out{k}")


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("solver", [1, 2])
@pytest.mark.parametrize("it", [0, 1, 4])
def test_flow_llin4(oracle, ref, shape, solver, it):
    s = synth.flow_system(12, *shape, late=True, nframes=3)
    a, b = both(oracle, ref, "Oflow_sor_llin4_2d", synth.mex_args("Oflow_sor_llin4_2d", s, it, 1.9, solver), 4)
    for k, (x, y) in enumerate(zip(a, b)):
        assert_bitwise(x, y, f"out{k}")


@pytest.mark.parametrize("solver", [1, 2])
def test_flow_llin8(oracle, ref, solver):
    s = synth.flow_system(13, 37, 53, late=True, eight=True)
    a, b = both(oracle, ref, "Oflow_sor_llin8_2d", synth.mex_args("Oflow_sor_llin8_2d", s, 4, 1.9, solver), 4)
    for k, (x, y) in enumerate(zip(a, b)):
        if solver == 1 or k >= 2:
            assert_bitwise(x, y, f"out{k}")
        else:
            assert rel_err(x, y) < 2e-6


@pytest.mark.parametrize("fn,late", [("Oflow_lhs_elin4_2d", False), ("Oflow_lhs_llin4_2d", True)])
@pytest.mark.parametrize("nframes", [1, 3])
def test_flow_lhs(oracle, ref, fn, late, nframes):
    s = synth.flow_system(14, 23, 31, late=late, nframes=nframes)
    a, b = both(oracle, ref, fn, synth.mex_args(fn, s), 2)
    for k, (x, y) in enumerate(zip(a, b)):
        assert_bitwise(x, y, f"out{k}")


@pytest.mark.parametrize("solver", [1, 2])
@pytest.mark.parametrize("it", [0, 1, 4])
def test_disparity(oracle, ref, solver, it):
    s = synth.disp_system(15, 37, 53)
    a, b = both(oracle, ref, "Disp_sor_llin4_2d", synth.mex_args("Disp_sor_llin4_2d", s, it, 1.9, solver), 2)
    for k, (x, y) in enumerate(zip(a, b)):
        assert_bitwise(x, y, f"out{k}")
    ss = {"f0": synth.disp_system(16, 37, 53), "f1": synth.disp_system(17, 37, 53)}
    a, b = both(oracle, ref, "Disp_sor_llin_sym4_2d", synth.mex_args("Disp_sor_llin_sym4_2d", ss, it, 1.9, solver), 2)
    for k, (x, y) in enumerate(zip(a, b)):
        assert_bitwise(x, y, f"sym out{k}")


@pytest.mark.parametrize("eight", [False, True])
@pytest.mark.parametrize("solver", [1, 2])
@pytest.mark.parametrize("it", [0, 1, 4])
def test_pde(oracle, ref, eight, solver, it):
    fn = "PDEsolver8" if eight else "PDEsolver4"
    s = synth.pde_system(18, 29, 41, nframes=2, eight=eight)
    a, b = both(oracle, ref, fn, synth.mex_args(fn, s, it, 1.75, solver), 1)
    assert_bitwise(a[0], b[0], fn)


@pytest.mark.parametrize("nframes", [1, 3])
def test_derivatives_and_weights(oracle, ref, nframes):
    I0, I1, _, _ = synth.image_pair(19, 37, 53, nframes=nframes)
    for fn, nl in (("FstDerivatives5", 3), ("SndDerivatives5", 5)):
        a, b = both(oracle, ref, fn, [I0, I1], nl)
        for k, (x, y) in enumerate(zip(a, b)):
            assert_bitwise(x, y, f"{fn} out{k}")
    a, b = both(oracle, ref, "DdiffWeights", [synth.f32(I0 * 10), synth.f32([[1e-3]])], 4)
    for k, (x, y) in enumerate(zip(a, b)):
        assert_bitwise(x, y, f"DdiffWeights out{k}")


@pytest.mark.parametrize("oob", [float("nan"), 0.0])
def test_warp(oracle, ref, oob):
    nr, nc = 37, 53
    I0, _, u, v = synth.image_pair(20, nr, nc, nframes=3)
    X, Y = np.meshgrid(np.arange(1, nc + 1, dtype=np.float32), np.arange(1, nr + 1, dtype=np.float32))
    X = synth.f32(X + 3 * u)
    Y = synth.f32(Y + 3 * v)
    X[3, 4] = np.nan            # UB conversions as gcc/x86-64 evaluates them
    Y[5, 6] = -5e9
    X[7, 7] = 4294967297.5
    X[9, 9] = nc                # exactly on the last column: valid, +1 tap clamped
    Y[9, 9] = nr
    assert_bitwise(oracle.bilin(I0, X, Y, oob), ref.bilin(I0, X, Y, oob), "bilin")
