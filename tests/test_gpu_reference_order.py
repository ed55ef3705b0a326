"""Solver 2 in the REFERENCE'S line order (pdegpu_set_sweep_order(ctx, PDEGPU_ORDER_REFERENCE) / PDEGPU_ORDER=reference).

The reference relaxes the lines of a direction one after the other (GS_ALR_SOR_*: opticalflowSolvers.c:196,690,1677;
disparitySolvers.c:154,452; pdeSolvers.c:277,344). In this order the GPU iterates are held to agreement with the
reference SWEEP BY SWEEP -- BASELINE.json north_star's 1e-5 bar for orderings that can be reproduced -- not only at
convergence (the zebra order of the throughput kernels: tests/test_gpu_sweeps.py, test_gpu_parity_fullsize.py):

  * every family, through the MEX gateways, after 1 and 4 iterations at the drivers' omega, at sizes with one and
    several lane chunk lengths, even and odd line lengths, and lines longer than the zebra kernel's segments;
  * the drivers that do not re-warp (Horn-Schunck on Urban3 as runme.m:74 runs it, FMG on Yosemite as runme.m:90 runs
    it) at the drivers' DEFAULT parameters against the restated .m drivers on the unmodified reference MEX code.
"""
import os

import numpy as np
import pytest

from oracle import pipelines
from pdegpu import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_SWEEP = 1e-5          # max |gpu - reference| / max |reference| per unknown field, after `iter` complete iterations


@pytest.fixture()
def reference_order():
    os.environ["PDEGPU_ORDER"] = "reference"
    yield
    os.environ["PDEGPU_ORDER"] = "fast"                      # the test session's default (conftest.py)


def _backend():
    from oracle import oracle as o
    return o.RefBackend() if o.have_ref() else o.OracleBackend()


def _rel(g, o):
    return float(np.nanmax(np.abs(g.astype(np.float64) - o))) / (float(np.nanmax(np.abs(o))) or 1.0)


CASES = [
    ("Oflow_sor_elin4_2d", lambda s, r, c: synth.flow_system(s, r, c, late=False, nan_frac=0.02), 2),
    ("Oflow_sor_llin4_2d", lambda s, r, c: synth.flow_system(s, r, c, late=True, nan_frac=0.02), 2),
    ("Oflow_sor_llin8_2d", lambda s, r, c: synth.flow_system(s, r, c, late=True, eight=True, nan_frac=0.02), 2),
    ("Disp_sor_llin4_2d", lambda s, r, c: synth.disp_system(s, r, c, nan_frac=0.02), 1),
    ("Disp_sor_llin_sym4_2d", lambda s, r, c: {"f0": synth.disp_system(s, r, c), "f1": synth.disp_system(s + 1, r, c)}, 2),
    ("PDEsolver4", lambda s, r, c: synth.pde_system(s, r, c, nframes=2, nan_frac=0.02), 1),
    ("PDEsolver8", lambda s, r, c: synth.pde_system(s, r, c, nframes=2, eight=True, nan_frac=0.02), 1),
]


@pytest.mark.parametrize("shape", [(37, 53), (120, 160), (203, 270), (480, 640)])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("iters", [1, 4])
@pytest.mark.parametrize("solver", [2, 1])
def test_sweep_by_sweep_agreement(gpu, reference_order, case, shape, iters, solver):
    """solver 2: the reference's line order (lex_pass_kernel); solver 1: its lexicographic point order, run as a
    wavefront over anti-diagonals (lex_point_kernel). The point flow solver is unstable for omega >~ 1.3 on both sides
    (DESIGN.md section 2): omega = 1 there."""
    fn, mk, nout = case
    s = mk(71, *shape)
    omega = 1.75 if fn.startswith("PDE") else (1.9 if solver == 2 else 1.0)
    a = synth.mex_args(fn, s, iters, omega, solver)
    g, o = gpu.call(fn, a, nout), _backend().call(fn, a, nout)
    for k in range(nout):
        e = _rel(g[k], o[k])
        assert np.isfinite(g[k]).all() and e < TOL_SWEEP, f"{fn} {shape} iter={iters} solver={solver}: output {k} differs by {e:.2e} of its range"


@pytest.mark.parametrize("fn,mk", [(CASES[0][0], CASES[0][1]), (CASES[3][0], CASES[3][1])], ids=["elin4_1080x1920", "disp_1080x1920"])
def test_long_lines_are_solved_uncut(gpu, reference_order, fn, mk):
    """lines of 1080 and 1920 elements (finest level of BASELINE configs[2]): the zebra kernel cuts them into segments,
    the reference order solves them whole"""
    s = mk(72, 1080, 1920)
    nout = 2 if fn.startswith("Oflow") else 1
    a = synth.mex_args(fn, s, 2, 1.9, 2)
    g, o = gpu.call(fn, a, nout), _backend().call(fn, a, nout)
    for k in range(nout):
        assert _rel(g[k], o[k]) < TOL_SWEEP


def test_gateways_follow_the_environment_and_auto_picks_by_family(gpu):
    """PDEGPU_ORDER is read at every gateway call. "auto" (what an unset variable means): the early-linearisation family
    runs in the reference's order, the late-linearisation one on the zebra kernels (different iterate, same fixed point)"""
    late = synth.flow_system(73, 64, 80, late=True)
    early = synth.flow_system(74, 64, 80, late=False)
    al = synth.mex_args("Oflow_sor_llin4_2d", late, 2, 1.9, 2)
    ae = synth.mex_args("Oflow_sor_elin4_2d", early, 2, 1.9, 2)
    ol, oe = _backend().call("Oflow_sor_llin4_2d", al, 2), _backend().call("Oflow_sor_elin4_2d", ae, 2)
    try:
        os.environ["PDEGPU_ORDER"] = "reference"
        assert _rel(gpu.call("Oflow_sor_llin4_2d", al, 2)[0], ol[0]) < TOL_SWEEP
        os.environ["PDEGPU_ORDER"] = "auto"
        assert _rel(gpu.call("Oflow_sor_elin4_2d", ae, 2)[0], oe[0]) < TOL_SWEEP
        assert _rel(gpu.call("Oflow_sor_llin4_2d", al, 2)[0], ol[0]) > 1e-3
        os.environ["PDEGPU_ORDER"] = "fast"
        assert _rel(gpu.call("Oflow_sor_elin4_2d", ae, 2)[0], oe[0]) > 1e-3
    finally:
        os.environ["PDEGPU_ORDER"] = "fast"


# ------------------------------------------------------------------------------------------------------------------
# the early-linearisation drivers at their DEFAULT parameters, on the reference's own demo data
# ------------------------------------------------------------------------------------------------------------------
def _epe(u0, v0, u1, v1):
    return float(np.mean(np.sqrt((u0.astype(np.float64) - u1) ** 2 + (v0.astype(np.float64) - v1) ** 2)))


@pytest.fixture()
def ref_ctx(built):
    from pdegpu import lib
    c = lib.Context(0)
    c.set_sweep_order(lib.ORDER_REFERENCE)
    yield c
    c.close()


@pytest.fixture()
def auto_ctx(built):
    """the library's default order (what a caller gets who sets nothing)"""
    from pdegpu import lib
    c = lib.Context(0)
    c.set_sweep_order(lib.ORDER_AUTO)
    yield c
    c.close()


def test_horn_schunck_urban3_defaults_give_the_reference_flow(auto_ctx):
    """FlowEminHS_elin_2D_v10(cat(3, I7, I8), 3) exactly as runme.m:74 calls it (iter = 20, omega = 1.9, alpha = 0.2), on a
    context with the library's default settings"""
    z = np.load(os.path.join(GOLD, "urban3_pair.npz"))
    I0, I1 = z["frame07"].astype(np.float32), z["frame08"].astype(np.float32)
    Ug, Vg = auto_ctx.flow_hs(I0, I1)
    Uo, Vo = pipelines.flow_hs(I0, I1, _backend())
    e = _epe(Ug, Vg, Uo, Vo)
    mag = float(np.mean(np.sqrt(Uo.astype(np.float64) ** 2 + Vo ** 2)))
    print(f"\nUrban3 Horn-Schunck at the driver's defaults, reference order: mean EPE GPU vs reference {e:.2e} px (mean |flow| {mag:.2f} px)")
    assert np.isfinite(Ug).all() and e < 1e-3


def test_fmg_yosemite_defaults_give_the_reference_flow(auto_ctx):
    """FlowEminNDFASFMG_elin_2D_v10(Y.I, 1) as runme.m:90 calls it; both sides also scored against Utrue / Vtrue"""
    z = np.load(os.path.join(GOLD, "yosemite.npz"))
    I = z["I"].astype(np.float32)
    I0, I1, ut, vt = I[:, :, 0:1], I[:, :, 1:2], z["Utrue"], z["Vtrue"]
    Ug, Vg = auto_ctx.flow_fmg(I0, I1)
    Uo, Vo = pipelines.flow_fmg(I0, I1, _backend())
    e, ag, ao = _epe(Ug, Vg, Uo, Vo), _epe(Ug, Vg, ut, vt), _epe(Uo, Vo, ut, vt)
    print(f"\nYosemite FMG at the driver's defaults, reference order: mean EPE GPU vs reference {e:.2e} px; "
          f"AEE vs ground truth GPU {ag:.4f} px, reference {ao:.4f} px")
    assert np.isfinite(Ug).all() and e < 1e-3 and abs(ag - ao) < 1e-3


def test_llin_flow_defaults_give_the_reference_flow(ref_ctx):
    """FlowEminND_llin_2D_v10 at its defaults (BASELINE configs[1]) on a synthetic 120 x 160 RGB pair"""
    I0, I1, u, v = synth.image_pair(11, 120, 160, nframes=3, scale=255.0, max_flow=2.0)
    Ug, Vg = ref_ctx.flow_llin(I0, I1)
    Uo, Vo = pipelines.flow_llin(I0, I1, _backend())
    e = _epe(Ug, Vg, Uo, Vo)
    print(f"\nllin flow at the driver's defaults, reference order: mean EPE GPU vs reference {e:.2e} px")
    assert np.isfinite(Ug).all() and e < 1e-3
