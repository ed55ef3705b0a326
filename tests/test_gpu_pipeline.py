"""The device-resident late-linearisation flow pipeline (pdegpu_flow_llin_2d, BASELINE configs[1]) against
the CPU restatement of the reference's driver built on the UNMODIFIED reference C code (oracle/pipelines.py
with RefBackend, or the C restatement when oracle/_ref is absent).

The reference relaxes lexicographically, libpdegpu in zebra order: iterates differ at finite `iter`, so
  * with every inner linear solve run to convergence (iter = 80, omega = 1.3) the two pipelines must agree
    to <= 1e-3 px mean end-point error (north_star's bar for Gauss-Seidel solvers);
  * with the driver's defaults (iter = 4, omega = 1.9) both must recover the known synthetic flow equally
    well (average end-point error against ground truth within 0.02 px of each other)."""
import numpy as np
import pytest

from oracle import pipelines
from pdegpu import synth

pytestmark = pytest.mark.gpu


def backend():
    from oracle import oracle as o
    return o.RefBackend() if o.have_ref() else o.OracleBackend()


def pair(seed, nr, nc, C=3):
    I0, I1, u, v = synth.image_pair(seed, nr, nc, nframes=C, scale=255.0, max_flow=2.0)
    return I0.reshape(nr, nc, C), I1.reshape(nr, nc, C), u, v


def epe(u0, v0, u1, v1, margin=8):
    s = (slice(margin, u0.shape[0] - margin), slice(margin, u0.shape[1] - margin))
    return float(np.mean(np.sqrt((u0[s].astype(np.float64) - u1[s]) ** 2 + (v0[s].astype(np.float64) - v1[s]) ** 2)))


@pytest.fixture(scope="module")
def ctx(built):
    from pdegpu import lib
    return lib.Context(0)


def test_pipeline_converged_solves_match_reference(ctx):
    nr, nc = 96, 128
    I0, I1, _, _ = pair(3, nr, nc)
    kw = dict(iter=80, omega=1.3, firstLoop=2, secondLoop=2)
    Ug, Vg = ctx.flow_llin(I0, I1, **kw)
    Uo, Vo = pipelines.flow_llin(I0, I1, backend(), **kw)
    assert np.isfinite(Ug).all() and np.isfinite(Vg).all()
    e = epe(Ug, Vg, Uo, Vo, margin=0)
    assert e < 1e-3, f"mean EPE between GPU and reference pipelines {e}"


def test_pipeline_default_parameters_quality(ctx):
    nr, nc = 120, 160
    I0, I1, u, v = pair(5, nr, nc)
    Ug, Vg = ctx.flow_llin(I0, I1)
    Uo, Vo = pipelines.flow_llin(I0, I1, backend())
    eg, eo = epe(Ug, Vg, u, v), epe(Uo, Vo, u, v)
    assert abs(eg - eo) < 0.02 and eg < 0.5, f"AEE vs ground truth: GPU {eg}, reference {eo}"


def test_pipeline_batch_equals_single(ctx):
    nr, nc = 64, 80
    ps = [pair(10 + k, nr, nc) for k in range(3)]
    I0 = np.stack([p[0] for p in ps]); I1 = np.stack([p[1] for p in ps])
    Ub, Vb = ctx.flow_llin(I0, I1)
    for k in range(3):
        U1, V1 = ctx.flow_llin(ps[k][0], ps[k][1])
        assert np.array_equal(U1, Ub[k]) and np.array_equal(V1, Vb[k])


def noisy_image(seed, nr, nc, fr):
    rng = np.random.default_rng(seed)
    ii, jj = np.meshgrid(np.arange(nr), np.arange(nc), indexing="ij")
    clean = np.stack([0.5 + 0.3 * np.sign(np.sin(ii / (9.0 + k)) * np.cos(jj / (11.0 + 2 * k))) for k in range(fr)], axis=2)
    return clean.astype(np.float32), (clean + 0.08 * rng.standard_normal(clean.shape)).astype(np.float32)


def test_tvdenoise8_converged_inner_solves_match_reference(ctx):
    """TVdenoise8 with the point solver run to convergence in every lagged-diffusivity step (the ALR solver of the
    8-neighbour family is hard-wired to one iteration): GPU (4-colour) and reference (lexicographic)
    pipelines must agree."""
    clean, noisy = noisy_image(1, 64, 80, 3)
    kw = dict(solver=1, inner_iter=150, omega=1.0, outer_iter=3)
    g = ctx.tvdenoise8(noisy, **kw)
    o = pipelines.tvdenoise8(noisy, backend(), **kw)
    assert np.isfinite(g).all()
    rel = float(np.abs(g - o).max() / np.abs(o).max())
    assert rel < 1e-3, f"max relative difference {rel}"


def test_tvdenoise8_default_parameters_quality(ctx):
    clean, noisy = noisy_image(2, 96, 120, 3)
    g = ctx.tvdenoise8(noisy)
    o = pipelines.tvdenoise8(noisy, backend())
    # one ALR iteration per lagged-diffusivity step (Q4): zebra and lexicographic iterates differ slightly; both must
    # land on the same image (the driver's alpha = 500 smooths heavily, so "quality" is agreement, not PSNR)
    rm = lambda a: float(np.sqrt(np.mean((a - clean) ** 2)))
    assert np.isfinite(g).all()
    assert abs(rm(g) - rm(o)) < 0.03 * rm(o), f"RMSE vs clean: GPU {rm(g)}, reference {rm(o)}, noisy {rm(noisy)}"
    assert float(np.mean(np.abs(g - o))) < 2e-2, f"mean |GPU - reference| = {float(np.mean(np.abs(g - o)))}"


# ---- FlowEminNDFASFMG_elin_2D_v10 (BASELINE configs[2]): early linearisation + full multigrid ----
def small_pair(seed, nr, nc, C):
    """early linearisation = small displacements (the driver's own note): sub-pixel synthetic flow"""
    I0, I1, u, v = synth.image_pair(seed, nr, nc, nframes=C, scale=255.0, max_flow=0.8)
    return I0.reshape(nr, nc, C), I1.reshape(nr, nc, C), u, v


@pytest.mark.parametrize("C,cycle", [(1, 1), (3, 1), (1, 2)])
def test_fmg_converged_solves_match_reference(ctx, C, cycle):
    """Every smoother call solved to convergence (iter = 600, omega = 1.6; the reference side is then within 5e-6 px of
    its own limit): zebra (GPU) and lexicographic (reference)
    line relaxation reach the same fixed point, so the whole FMG/FAS pipeline must agree to <= 1e-3 px mean EPE.
    Four levels: with the 6x8 fifth level the reference's FAS iteration is itself unstable at converged inner solves
    (the restatement on the reference MEX code diverges there too), which is not a property of the sweep."""
    nr, nc = 96, 128
    I0, I1, _, _ = small_pair(3, nr, nc, C)
    kw = dict(iter=600, omega=1.6, firstLoop=2, max_scales=4 if cycle == 1 else 3, cycle_index=cycle)
    Ug, Vg = ctx.flow_fmg(I0, I1, **kw)
    Uo, Vo = pipelines.flow_fmg(I0, I1, backend(), **kw)
    assert np.isfinite(Ug).all() and np.isfinite(Vg).all()
    e = epe(Ug, Vg, Uo, Vo, margin=0)
    assert e < 1e-3, f"mean EPE between GPU and reference FMG pipelines {e}"


def test_fmg_driver_iteration_counts(ctx):
    """At the driver's defaults (iter = 4, omega = 1.9) the two orderings are far from converged and differ: the
    reference's lexicographic line sweeps carry information across the whole image in one sweep, zebra sweeps only
    between neighbouring lines, so after 4 iterations the zebra iterate is the less accurate one (measured: AEE 0.10
    against 0.03 px on this pair). Checked here: the default call is sane, and a moderate number
    of sweeps (iter = 64 on four levels, ~0.1 s at 480x640) already agrees with the reference run the same way."""
    nr, nc = 120, 160
    I0, I1, u, v = small_pair(5, nr, nc, 1)
    Ug, Vg = ctx.flow_fmg(I0, I1)
    mag = float(np.mean(np.sqrt(u ** 2 + v ** 2)))
    assert np.isfinite(Ug).all() and epe(Ug, Vg, u, v) < 0.6 * mag
    kw = dict(iter=64, omega=1.6, max_scales=4)
    Ug, Vg = ctx.flow_fmg(I0, I1, **kw)
    Uo, Vo = pipelines.flow_fmg(I0, I1, backend(), **kw)
    assert epe(Ug, Vg, Uo, Vo, margin=0) < 5e-3
    assert abs(epe(Ug, Vg, u, v) - epe(Uo, Vo, u, v)) < 5e-3


def test_fmg_batch_equals_single(ctx):
    nr, nc = 64, 80
    ps = [small_pair(20 + k, nr, nc, 1) for k in range(2)]
    I0 = np.stack([p[0] for p in ps]); I1 = np.stack([p[1] for p in ps])
    Ub, Vb = ctx.flow_fmg(I0, I1)
    for k in range(2):
        U1, V1 = ctx.flow_fmg(ps[k][0], ps[k][1])
        assert np.array_equal(U1, Ub[k]) and np.array_equal(V1, Vb[k])


@pytest.mark.parametrize("driver", ["fmg", "hs"])
def test_pairs_on_parallel_lanes_equal_single_pairs(built, driver):
    """the pairs of a batch run side by side on the context's lanes (child streams with their own workspace, one branch
    per lane in the captured graph): direct run, capture and replay of a batch of 5 must all equal the pairs run alone"""
    from pdegpu import lib
    c = lib.Context(0)
    nr, nc = 72, 88
    ps = [small_pair(30 + k, nr, nc, 1) for k in range(5)]
    I0 = np.stack([p[0] for p in ps]); I1 = np.stack([p[1] for p in ps])
    fn = c.flow_fmg if driver == "fmg" else c.flow_hs
    single = [fn(p[0], p[1]) for p in ps]
    for _ in range(3):
        Ub, Vb = fn(I0, I1)
        for k in range(5):
            assert np.array_equal(single[k][0], Ub[k]) and np.array_equal(single[k][1], Vb[k])
    c.close()


# ---- FlowEminHS_elin_2D_v10 (BASELINE configs[0]): Horn-Schunck, one linear solve per pyramid level ----
@pytest.mark.parametrize("C", [1, 3])
def test_hs_converged_solves_match_reference(ctx, C):
    """The driver's alpha = 0.2 on frames scaled to 0..1 makes the system a nearly pure Laplacian (thousands of sweeps to
    converge, for either ordering; the zebra iterate keeps a constant offset longest: measured 1.2e-2 px at alpha = 0.02 after
    1600 iterations, 2e-6 px after 6400); with alpha = 0.002 both orderings reach the solution quickly and the pipelines
    must agree."""
    nr, nc = 64, 80
    I0, I1, _, _ = small_pair(31, nr, nc, C)
    kw = dict(iter=1600, omega=1.8, alpha=0.002)
    Ug, Vg = ctx.flow_hs(I0, I1, **kw)
    Uo, Vo = pipelines.flow_hs(I0, I1, backend(), **kw)
    assert np.isfinite(Ug).all() and np.isfinite(Vg).all()
    e = epe(Ug, Vg, Uo, Vo, margin=0)
    assert e < 1e-3, f"mean EPE between GPU and reference Horn-Schunck pipelines {e}"


def test_hs_driver_iteration_counts(ctx):
    """Driver defaults (iter = 20, omega = 1.9, ALR, alpha = 0.2): far from converged for both orderings, and the zebra
    iterate trails the lexicographic one (see test_fmg_driver_iteration_counts; measured AEE 0.24 against 0.16 px here).
    The default call must be sane, and with a data term that lets 100 sweeps converge (alpha = 0.002) both must have
    recovered the flow equally well."""
    nr, nc = 120, 160
    I0, I1, u, v = small_pair(32, nr, nc, 3)
    mag = float(np.mean(np.sqrt(u ** 2 + v ** 2)))
    Ug, Vg = ctx.flow_hs(I0, I1)
    assert np.isfinite(Ug).all() and epe(Ug, Vg, u, v) < 1.5 * mag
    Ug, Vg = ctx.flow_hs(I0, I1, iter=100, alpha=0.002)
    Uo, Vo = pipelines.flow_hs(I0, I1, backend(), iter=100, alpha=0.002)
    eg, eo = epe(Ug, Vg, u, v), epe(Uo, Vo, u, v)
    assert abs(eg - eo) < 0.02, f"AEE vs ground truth at iter=100, alpha=0.002: GPU {eg}, reference {eo}"


# ---- DispEminND_llin_sym_2D (BASELINE configs[3]): symmetric-constraint stereo ----
def stereo_pair(seed, nr, nc, C, max_disp=3.0):
    I0, I1, u, _ = synth.image_pair(seed, nr, nc, nframes=C, scale=255.0, max_flow=max_disp, horizontal=True)
    return I0.reshape(nr, nc, C), I1.reshape(nr, nc, C), u


@pytest.mark.parametrize("C,u8", [(1, 1), (3, 0)])
def test_disp_sym_converged_solves_match_reference(ctx, C, u8):
    nr, nc = 96, 128
    Il, Ir, _ = stereo_pair(41, nr, nc, C)
    kw = dict(iter=150, omega=1.3, firstLoop=2, secondLoop=2, uint8_input=u8)
    g0, g1 = ctx.disp_sym(Il, Ir, **kw)
    o0, o1 = pipelines.disp_sym(Il, Ir, backend(), **kw)
    assert np.array_equal(np.isnan(g0), np.isnan(o0)) and np.array_equal(np.isnan(g1), np.isnan(o1))
    e = float(np.nanmean(np.abs(g0 - o0)) + np.nanmean(np.abs(g1 - o1)))
    assert e < 1e-3, f"mean |GPU - reference| over both disparity fields {e}"


def test_disp_sym_default_parameters_quality(ctx):
    nr, nc = 120, 160
    Il, Ir, u = stereo_pair(42, nr, nc, 3)
    g0, g1 = ctx.disp_sym(Il, Ir)
    o0, o1 = pipelines.disp_sym(Il, Ir, backend())
    s = (slice(10, -10), slice(10, -10))
    eg = float(np.nanmean(np.abs(g0[s] - u[s]))); eo = float(np.nanmean(np.abs(o0[s] - u[s])))
    assert abs(eg - eo) < 0.02 and eg < 0.2, f"mean abs disparity error: GPU {eg}, reference {eo}"
    # symmetry: the right -> left field mirrors the left -> right one
    assert float(np.nanmean(np.abs(g1[s] + u[s]))) < 0.2


# ---- CUDA-graph replay of the pipelines (pdegpu_graph_run): call 1 runs the launches directly, call 2 captures them,
#      call 3 replays the graph; all three must return the same bits, and a different input must not hit a stale graph ----
def test_pipeline_graph_replay_is_bitwise(built):
    from pdegpu import lib
    c = lib.Context(0)
    nr, nc = 64, 80
    I0, I1, _, _ = small_pair(51, nr, nc, 1)
    J0, J1, _, _ = small_pair(52, nr, nc, 1)
    l0 = c.launches
    a = c.flow_fmg(I0, I1)
    per_call = c.launches - l0
    b = c.flow_fmg(I0, I1)
    d = c.flow_fmg(I0, I1)
    assert c.launches - l0 == 3 * per_call                       # replays are counted like the launches they stand for
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[0], d[0]) and np.array_equal(a[1], d[1])
    e = c.flow_fmg(J0, J1)                                       # same buffers, new content: the graph reads the new data
    assert not np.array_equal(e[0], a[0])
    f = c.flow_fmg(J0, J1, iter=8)                               # other parameters: another key
    assert not np.array_equal(f[0], e[0])
    P0, P1, _, _ = pair(53, nr, nc)
    r = [c.flow_llin(P0, P1) for _ in range(3)]
    assert np.array_equal(r[0][0], r[1][0]) and np.array_equal(r[0][0], r[2][0])
    Sl, Sr, _ = stereo_pair(54, nr, nc, 1)
    q = [c.disp_sym(Sl, Sr) for _ in range(3)]
    assert np.array_equal(q[0][0], q[2][0], equal_nan=True) and np.array_equal(q[0][1], q[2][1], equal_nan=True)
    h = [c.flow_hs(P0, P1) for _ in range(3)]
    assert np.array_equal(h[0][0], h[2][0])
