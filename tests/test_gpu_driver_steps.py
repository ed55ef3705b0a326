"""GPU parity of the driver-side stencils against oracle/matlab_steps.py, the CPU
restatement of the reference's .m formulas. Called through the C ABI (pdegpu_dev_*).

Tolerances: formulas evaluated in single by the drivers are restated with one rounding per operation on
both sides -> bit-exact or 1e-6; double-precision formulas cast to single -> 1e-6 relative (north_star:
"warps and pyramids must match within 1e-6", non-iterative kernels 1e-5)."""
import os

import numpy as np
import pytest

from oracle import matlab_steps as ms
from pdegpu import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def steps(built):
    from pdegpu.steps import Steps
    return Steps()


def close(a, b, tol):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b))
    return float(np.nanmax(np.abs(a - b)) / (np.nanmax(np.abs(b)) + 1e-30)) < tol


def rnd(seed, *shape, scale=1.0):
    return (np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float32)


@pytest.mark.parametrize("shape", [(37, 53), (64, 48), (3, 5)])
def test_op_diff_weights(steps, shape):
    U, V = rnd(1, *shape), rnd(2, *shape)
    g = steps.op_diff_weights(U, V)
    o = ms.op_diff_weights(U, V)
    for a, b in zip(g, o):
        assert close(a, b.astype(np.float32), 1e-6)
    # every edge is computed once and stored on both of its sides (returned order: wW, wN, wS, wE)
    assert np.array_equal(g[0], np.roll(g[3], 1, axis=1)) and np.array_equal(g[1], np.roll(g[2], 1, axis=0))


@pytest.mark.parametrize("second", ["none", "first", "gradmag"])
@pytest.mark.parametrize("shape,c1,c2", [((37, 53), 6, 3), ((16, 24), 1, 1)])
def test_llin_terms(steps, second, shape, c1, c2):
    d1 = [rnd(10 + k, *shape, c1, scale=0.3) for k in range(3)]
    d1[0][3, 4, 0] = np.nan                      # out-of-image warp -> NaN derivative -> nansum skips the channel
    d2 = None
    if second == "first":
        d2 = [rnd(20 + k, *shape, c2, scale=0.3) for k in range(3)]
    elif second == "gradmag":
        d2 = [rnd(30 + k, *shape, c2, scale=0.3) for k in range(5)]
        d2[2][5, 6, 0] = np.nan
    dU, dV = rnd(40, *shape, scale=0.5), rnd(41, *shape, scale=0.5)
    g = steps.llin_terms(d1, d2, dU, dV, 1.4843, 0.2915, 0.042, second == "gradmag")
    o = ms.llin_terms(d1, d2, dU, dV, 1.4843, 0.2915, 0.042, second == "gradmag")
    for a, b in zip(g, o):
        assert close(a, b, 2e-6)


@pytest.mark.parametrize("summed", [True, False])
def test_elin_terms(steps, summed):
    shape, ch = (33, 41), 3
    der = [rnd(50 + k, *shape, ch, scale=0.2) for k in range(8)]
    coef = [rnd(60 + k, *shape, ch, scale=0.2) for k in range(5)]
    U, V = rnd(70, *shape), rnd(71, *shape)
    g = steps.elin_terms(der, coef, U, V, 0.03, 0.97, 0.035, summed)
    o = ms.elin_terms(der, coef, U, V, 0.03, 0.97, 0.035, summed)
    for a, b in zip(g, o):
        assert close(a, b, 2e-6)


def test_disp_sym_terms(steps):
    shape, ch = (29, 37), 3
    d = [rnd(80 + k, *shape, ch, scale=0.2) for k in range(6)]
    dU, Udt, Udx = rnd(90, *shape, scale=0.5), rnd(91, *shape, scale=0.5), rnd(92, *shape, scale=0.1)
    g = steps.disp_sym_terms(d, dU, Udt, Udx, 0.25, 0.72, 0.035, 0.4, 1.125)
    o = ms.disp_sym_terms(d, dU, Udt, Udx, 0.25, 0.72, 0.035, 0.4, 1.125)
    for a, b in zip(g, o):
        assert close(a, b, 2e-6)


def test_fas_rhs(steps):
    R, A = rnd(1, 21, 17, 3), rnd(2, 21, 17, 3)
    gd = np.abs(rnd(3, 21, 17, 3)) + 0.5
    assert np.array_equal(steps.fas_rhs(R, A, gd), ms.fas_rhs(R, A, gd))


@pytest.mark.parametrize("shape", [(37, 53), (480, 640), (5, 7)])
def test_gaussian_imfilter(steps, shape):
    A = rnd(5, *shape)
    G = ms.fspecial_gaussian(5, 1.25)
    assert close(steps.imfilter(A, G), ms.imfilter(A, G, "replicate"), 1e-6)


def test_rgb2grad(steps):
    A = rnd(6, 31, 45, 3)
    assert close(steps.rgb2grad(A), ms.rgb2grad(A), 1e-6)


@pytest.mark.parametrize("shape", [(37, 53), (64, 48), (11, 10)])
def test_lpf_pyramid_and_restriction(steps, shape):
    A = rnd(7, *shape)
    lpf = np.array([[1, 4, 6, 4, 1]], dtype=np.float64) / 16.0
    g = steps.imfilter(steps.imfilter(A, lpf, conv=True), lpf.T, conv=True, step=2)
    assert close(g, ms.lpf_decimate(A), 1e-6)
    fw = np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=np.float64) / 16.0
    g = steps.imfilter(A, fw, conv=True, step=2, prescale=0.5)
    assert close(g, ms.fw_restrict(A, 0.5), 1e-6)


@pytest.mark.parametrize("shape", [(37, 53), (480, 640), (20, 27)])
def test_imresize_pyramid_down_and_up(steps, shape):
    A = rnd(8, *shape)
    d_g, d_o = steps.imresize_bilinear(A, scale=0.75), ms.imresize_bilinear(A, scale=0.75)
    assert d_g.shape == d_o.shape and close(d_g, d_o, 1e-6)
    u_g, u_o = steps.imresize_bilinear(d_o, output_size=shape), ms.imresize_bilinear(d_o, output_size=shape)
    assert close(u_g, u_o, 1e-6)
    # FAS prolongation: factor ~2, non-integer ratio for odd sizes
    c = rnd(9, (shape[0] + 1) // 2, (shape[1] + 1) // 2)
    assert close(steps.imresize_bilinear(c, output_size=shape), ms.imresize_bilinear(c, output_size=shape), 1e-6)


@pytest.mark.parametrize("shape", [(37, 53), (3, 3), (128, 96)])
def test_medfilt3(steps, shape):
    A = rnd(10, *shape)
    assert np.array_equal(steps.medfilt3(A), ms.medfilt2_symmetric(A))


def test_warp_coords(steps):
    U, V = rnd(11, 23, 31), rnd(12, 23, 31)
    X, Y = steps.warp_coords(U, V)
    jj, ii = np.meshgrid(np.arange(1, 32, dtype=np.float32), np.arange(1, 24, dtype=np.float32))
    assert np.array_equal(X, jj + U) and np.array_equal(Y, ii + V)


@pytest.mark.parametrize("shape,frames", [((37, 53), 1), ((64, 48), 3), ((120, 160), 1)])
def test_ad_diff_weights_and_tv_terms(steps, shape, frames):
    """ADdiffWeights incl. the device-side quantile select, and the TV data terms (TVdenoise8.m:83-85,119-231)."""
    D = np.abs(rnd(13, *shape, frames, scale=0.3)) + 0.2
    D[5:9, 7:12] = 0.5                                          # a flat patch: zero gradients are excluded from the quantile
    Iin = (D + rnd(14, *shape, frames, scale=0.05)).astype(np.float32)
    if frames == 1:
        D, Iin = D[:, :, 0], Iin[:, :, 0]
    alpha = 500.0
    o = ms.ad_diff_weights(D)
    g = steps.ad_diff_weights(D, Iin=Iin, scale=alpha)
    assert abs(g[8] - o[8]) <= 1e-12 * o[8], f"lambda {g[8]} vs {o[8]}"
    tr, b, ws = ms.tv_terms(D, Iin, o[:8], alpha)
    for k in range(8):
        assert close(g[k], ws[k], 1e-6)
    assert close(g[9], tr, 1e-6) and close(g[10], b, 1e-6)
    g1 = steps.ad_diff_weights(D)
    for k in range(8):
        assert close(g1[k], o[k].astype(np.float32), 1e-6)


@pytest.mark.parametrize("shape", [(96, 120), (203, 270), (37, 53)])
@pytest.mark.parametrize("order", ["fast", "reference"])
def test_fused_inner_solve_equals_the_three_steps(steps, shape, order):
    """pdegpu_dev_llin_solve (FlowEminND_llin_2D_v10.m:278-348 in one call; weights and terms computed inside the line
    kernels' preparation, north_star subsystem 3) against OPdiffWeights, the term assembly and Oflow_sor_llin4_2d called
    one after the other: the same bits. (37 x 53 takes the shared-memory resident kernel: the unfused path inside.)
    The fused preparation is opt-in (PDEGPU_FUSE=1, read once per process: set in tests/conftest.py)."""
    from pdegpu import lib, mex
    nr, nc = shape
    r = np.random.default_rng(17)
    c1, c2 = 6, 3
    d1 = [(r.standard_normal((nr, nc, c1)) * 0.1).astype(np.float32) for _ in range(3)]
    d2 = [(r.standard_normal((nr, nc, c2)) * 0.1).astype(np.float32) for _ in range(5)]
    d1[0][3, 4, 1] = np.nan                                    # a warped pixel that left the image
    U, V = synth.smooth_field(r, nr, nc, 2.0).astype(np.float32), synth.smooth_field(r, nr, nc, 2.0).astype(np.float32)
    dU, dV = (0.05 * r.standard_normal((nr, nc))).astype(np.float32), (0.05 * r.standard_normal((nr, nc))).astype(np.float32)
    steps.ctx.set_sweep_order(lib.ORDER_REFERENCE if order == "reference" else lib.ORDER_FAST)
    os.environ["PDEGPU_ORDER"] = order
    try:
        fu, fv = steps.llin_solve(d1, d2, U, V, dU, dV, 1.4843, 0.2915, 0.042, True, 4, 1.9)
        wW, wN, wS, wE = steps.op_diff_weights((U + dU).astype(np.float32), (V + dV).astype(np.float32))
        M, Cu, Cv, Du, Dv = steps.llin_terms(d1, d2, dU, dV, 1.4843, 0.2915, 0.042, True)
        su, sv = mex.Oflow_sor_llin4_2d(U, V, dU, dV, M, Cu, Cv, Du, Dv, wW, wN, wE, wS, np.float32(4), np.float32(1.9), np.float32(2))
    finally:
        os.environ["PDEGPU_ORDER"] = "fast"
        steps.ctx.set_sweep_order(lib.ORDER_FAST)
    assert np.array_equal(fu, su, equal_nan=True) and np.array_equal(fv, sv, equal_nan=True)
