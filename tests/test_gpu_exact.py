"""GPU parity, non-iterative kernels: bit-exact against the oracle (and the golden vectors of the
unmodified reference). Called through the MEX gateways -> C ABI -> CUDA."""
import glob
import os

import numpy as np
import pytest

from pdegpu import synth
from util import assert_bitwise

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SHAPES = [(37, 53), (5, 5), (64, 8), (130, 71), (480, 640)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("nframes", [1, 3])
@pytest.mark.parametrize("late", [False, True])
def test_residual_and_lhs(gpu, oracle, shape, nframes, late):
    s = synth.flow_system(21, *shape, late=late, nframes=nframes, nan_frac=0.02)
    fn = "Oflow_sor_llin4_2d" if late else "Oflow_sor_elin4_2d"
    a = synth.mex_args(fn, s, 0, 1.9, 2)          # iter = 0: only the residual path
    got, want = gpu.call(fn, a, 4), oracle.call(fn, a, 4)
    for k in range(4):
        assert_bitwise(got[k], want[k], f"{fn} out{k}")
    assert not got[0].any() and not got[1].any()  # iter<=0 leaves the solution outputs zero
    fn = "Oflow_lhs_llin4_2d" if late else "Oflow_lhs_elin4_2d"
    a = synth.mex_args(fn, s)
    got, want = gpu.call(fn, a, 2), oracle.call(fn, a, 2)
    for k in range(2):
        assert_bitwise(got[k], want[k], f"{fn} out{k}")


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("nframes", [1, 3])
def test_derivatives(gpu, oracle, shape, nframes):
    I0, I1, _, _ = synth.image_pair(22, *shape, nframes=nframes, scale=255.0)
    for fn, nl in (("FstDerivatives5", 3), ("SndDerivatives5", 5)):
        got, want = gpu.call(fn, [I0, I1], nl), oracle.call(fn, [I0, I1], nl)
        for k in range(nl):
            assert_bitwise(got[k], want[k], f"{fn} out{k}")


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("nframes", [1, 3])
def test_diffusion_weights(gpu, oracle, shape, nframes):
    D = synth.f32(synth.image_pair(23, *shape, nframes=nframes, scale=16.0)[0])
    if nframes > 1:
        D[2, 3, 1] = np.nan
    a = [D, synth.f32([[1e-3]])]
    got, want = gpu.call("DdiffWeights", a, 4), oracle.call("DdiffWeights", a, 4)
    for k in range(4):
        assert_bitwise(got[k], want[k], f"DdiffWeights out{k}")


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("nframes", [1, 3])
def test_warp(gpu, oracle, shape, nframes):
    nr, nc = shape
    I0, _, u, v = synth.image_pair(24, nr, nc, nframes=nframes)
    X, Y = np.meshgrid(np.arange(1, nc + 1, dtype=np.float32), np.arange(1, nr + 1, dtype=np.float32))
    X = synth.f32(X + 4 * u)
    Y = synth.f32(Y + 4 * v)
    X[1, 2] = np.nan
    Y[2, 1] = -5e9
    X[3, 3] = 4294967297.5
    X[4, 4], Y[4, 4] = nc, nr          # last pixel: valid, +1 taps clamped
    X[0, 0], Y[0, 0] = 0.999, 1.0      # just outside
    got = gpu.call("BilinInterp_2d", [I0, X, Y], 1)[0]
    assert_bitwise(got, oracle.bilin(I0, X, Y, float("nan")), "BilinInterp_2d (oob=NaN)")


def test_warp_oob_value_through_c_abi(built, oracle):
    """The C ABI takes the out-of-image value explicitly."""
    import ctypes
    from pdegpu import lib
    ctx = lib.Context(0)
    nr, nc = 37, 53
    I0, _, u, v = synth.image_pair(25, nr, nc)
    X, Y = np.meshgrid(np.arange(1, nc + 1, dtype=np.float32), np.arange(1, nr + 1, dtype=np.float32))
    X, Y = synth.f32(X + 6 * u), synth.f32(Y + 6 * v)
    out = np.zeros((nr, nc), np.float32, order="F")
    fp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    L = lib.dll()
    L.pdegpu_bilin_interp_2d.restype = ctypes.c_int
    L.pdegpu_bilin_interp_2d.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int] * 3 + [ctypes.c_float]
    for oob in (0.0, -7.5):
        assert L.pdegpu_bilin_interp_2d(ctx.h, fp(out), fp(I0), fp(X), fp(Y), nr, nc, 1, oob) == 0
        assert_bitwise(out, oracle.bilin(I0, X, Y, oob), f"oob={oob}")
    ctx.close()


GOLD_EXACT = [f for f in sorted(glob.glob(os.path.join(GOLD, "*.npz")))
              if os.path.basename(f).split("_")[0] in ("lhs", "fst.npz", "snd.npz", "ddiff.npz")
              or "_it0" in f]


@pytest.mark.parametrize("path", GOLD_EXACT, ids=[os.path.basename(f)[:-4] for f in GOLD_EXACT])
def test_gpu_matches_reference_golden(gpu, path):
    z = np.load(path)
    fn, nlhs = str(z["fn"]), int(z["nlhs"])
    nin = len([k for k in z.files if k.startswith("in")])
    got = gpu.call(fn, [np.asfortranarray(z[f"in{k}"]) for k in range(nin)], nlhs)
    for k in range(nlhs):
        assert_bitwise(got[k], np.asfortranarray(z[f"out{k}"]), f"{fn} out{k}")


def test_gpu_warp_matches_reference_golden(gpu):
    z = np.load(os.path.join(GOLD, "warp.npz"))
    got = gpu.call("BilinInterp_2d", [z["I"], z["X"], z["Y"]], 1)[0]
    assert_bitwise(got, np.asfortranarray(z["out_nan"]), "warp golden")
