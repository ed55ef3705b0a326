"""CPU checks of the driver restatements in oracle/pipelines.py (they are the checker of tests/test_gpu_pipeline.py):
run around the compiled reference MEX code (or the C restatement when oracle/_ref is absent) they must recover a known
synthetic flow. Small sizes: the whole file runs in seconds."""
import numpy as np

from oracle import pipelines
from pdegpu import synth


def backend():
    from oracle import oracle as o
    return o.RefBackend() if o.have_ref() else o.OracleBackend()


def aee(U, V, u, v, m=8):
    s = (slice(m, U.shape[0] - m), slice(m, U.shape[1] - m))
    return float(np.mean(np.sqrt((U[s] - u[s]) ** 2 + (V[s] - v[s]) ** 2)))


def test_fmg_restatement_recovers_subpixel_flow():
    nr, nc = 96, 128
    I0, I1, u, v = synth.image_pair(3, nr, nc, nframes=1, scale=255.0, max_flow=0.8)
    U, V = pipelines.flow_fmg(I0.reshape(nr, nc, 1), I1.reshape(nr, nc, 1), backend())
    assert U.dtype == np.float32 and U.shape == (nr, nc) and np.isfinite(U).all()
    e, mag = aee(U, V, u, v), float(np.mean(np.sqrt(u ** 2 + v ** 2)))
    assert e < 0.05 and e < 0.25 * mag, f"AEE {e} for a mean displacement of {mag}"


def test_fmg_pyramid_stops_at_ten_pixels():
    # FlowEminNDFASFMG_elin_2D_v10.m:113-117: the level that reaches <= 10 pixels is the last one
    nr, nc = 96, 128
    I0, I1, _, _ = synth.image_pair(4, nr, nc, nframes=1, scale=255.0, max_flow=0.5)
    a = pipelines.flow_fmg(I0.reshape(nr, nc, 1), I1.reshape(nr, nc, 1), backend())
    b = pipelines.flow_fmg(I0.reshape(nr, nc, 1), I1.reshape(nr, nc, 1), backend(), max_scales=5)    # 96 48 24 12 6
    c = pipelines.flow_fmg(I0.reshape(nr, nc, 1), I1.reshape(nr, nc, 1), backend(), max_scales=9)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[0], c[0])


def test_llin_restatement_recovers_flow():
    nr, nc = 64, 80
    I0, I1, u, v = synth.image_pair(7, nr, nc, nframes=3, scale=255.0, max_flow=2.0)
    U, V = pipelines.flow_llin(I0.reshape(nr, nc, 3), I1.reshape(nr, nc, 3), backend())
    assert aee(U, V, u, v) < 0.25


def test_hs_restatement_recovers_subpixel_flow():
    """Horn-Schunck with the driver's alpha = 0.2 over-smooths frames scaled to 0..1 (the parameters are tuned for the
    Middlebury pair of runme.m:74); with a weaker smoothness term the restatement must beat the zero flow clearly."""
    nr, nc = 96, 128
    I0, I1, u, v = synth.image_pair(31, nr, nc, nframes=3, scale=255.0, max_flow=0.8)
    U, V = pipelines.flow_hs(I0.reshape(nr, nc, 3), I1.reshape(nr, nc, 3), backend(), alpha=0.002, iter=100)
    assert U.dtype == np.float32 and np.isfinite(U).all()
    e, mag = aee(U, V, u, v), float(np.mean(np.sqrt(u ** 2 + v ** 2)))
    assert e < 0.8 * mag, f"AEE {e} for a mean displacement of {mag}"


def test_disp_sym_restatement_recovers_disparity():
    nr, nc = 96, 128
    Il, Ir, u, _ = synth.image_pair(41, nr, nc, nframes=3, scale=255.0, max_flow=3.0, horizontal=True)
    U0, U1 = pipelines.disp_sym(Il, Ir, backend())
    s = (slice(10, -10), slice(10, -10))
    assert float(np.nanmean(np.abs(U0[s] - u[s]))) < 0.05 and float(np.nanmean(np.abs(U1[s] + u[s]))) < 0.05


def test_interp2_rows_matches_definition():
    from oracle import matlab_steps as ms
    V = np.arange(12, dtype=np.float32).reshape(3, 4)          # rows x cols, value = 4*i + j
    Xq = np.array([[1.0, 2.5, 4.0, 4.01], [0.99, 1.25, 3.75, np.nan], [1.0, 1.0, 4.0, 2.0]])
    out = ms.interp2_rows(V, Xq)
    exp = np.array([[0.0, 1.5, 3.0, np.nan], [np.nan, 4.25, 6.75, np.nan], [8.0, 8.0, 11.0, 9.0]], dtype=np.float32)
    assert np.array_equal(np.isnan(out), np.isnan(exp)) and np.allclose(np.nan_to_num(out), np.nan_to_num(exp))


# ---- sibling drivers: TVdenoise4, DispEminND_llin_2D, FlowEminAD_llin_2D_v10 ----
def test_tvdenoise4_restatement_removes_noise():
    rng = np.random.default_rng(5)
    nr, nc = 64, 80
    # a smooth image: the driver relaxes towards its input smoothed by a 7x7 Gaussian (TVdenoise4.m:57,66), so sharp
    # edges of a test image would be counted against it
    ii, jj = np.meshgrid(np.arange(nr), np.arange(nc), indexing="ij")
    clean = (0.5 + 0.3 * np.sin(ii / 17.0) * np.cos(jj / 23.0)).astype(np.float32)
    noisy = (clean + 0.08 * rng.standard_normal((nr, nc))).astype(np.float32)
    out = pipelines.tvdenoise4(noisy, backend(), outer_iter=4)
    assert out.shape == noisy.shape and out.dtype == np.float32 and np.isfinite(out).all()
    rm = lambda a: float(np.sqrt(np.mean((a - clean) ** 2)))
    assert rm(out) < 0.6 * rm(noisy), (rm(out), rm(noisy))


def test_tv4_diff_weights_against_the_written_formula():
    """DiffWeights (TVdenoise4.m:116-148) evaluated pixel by pixel in double precision from its text"""
    from oracle import matlab_steps as ms
    rng = np.random.default_rng(6)
    D = rng.random((7, 9, 2)).astype(np.float32)
    wW, wN, wE, wS = ms.tv4_diff_weights(D)
    r, c, f = D.shape
    cl = lambda v, n: min(max(v, 0), n - 1)
    ver = lambda i, j, k: 0.25 * D[cl(i - 1, r), j, k] - 0.25 * D[cl(i + 1, r), j, k]      # imfilter = correlation, replicate
    hor = lambda i, j, k: 0.25 * D[i, cl(j - 1, c), k] - 0.25 * D[i, cl(j + 1, c), k]
    for (i, j) in ((1, 1), (3, 4), (5, 7), (2, 3), (0, 5)):
        jw, is_ = (j - 1) % c, (i + 1) % r                                                   # circshift wraps
        eW = max((float(D[i, jw, k]) - D[i, j, k]) ** 2 + (ver(i, j, k) + ver(i, jw, k)) ** 2 for k in range(f))
        eS = max((float(D[is_, j, k]) - D[i, j, k]) ** 2 + (hor(i, j, k) + hor(is_, j, k)) ** 2 for k in range(f))
        assert abs(wW[i, j] - 1 / np.sqrt(eW + 1e-5)) < 2e-5 * wW[i, j]
        assert abs(wS[i, j] - 1 / np.sqrt(eS + 1e-5)) < 2e-5 * wS[i, j]
    # :145-148: the edges that would leave the image carry no weight
    assert not wW[:, 0].any() and not wE[:, -1].any() and not wN[0, :].any() and not wS[-1, :].any()
    assert wE[:, :-1].all() and wN[1:, :].all()


def test_disp_llin_restatement_recovers_disparity():
    nr, nc = 96, 128
    Il, Ir, u, _ = synth.image_pair(43, nr, nc, nframes=3, scale=255.0, max_flow=3.0, horizontal=True)
    U = pipelines.disp_llin(Il, Ir, backend())
    s = (slice(10, -10), slice(10, -10))
    assert U.dtype == np.float32 and float(np.nanmean(np.abs(U[s] - u[s]))) < 0.08


def test_flow_ad_restatement_recovers_flow():
    nr, nc = 64, 80
    I0, I1, u, v = synth.image_pair(9, nr, nc, nframes=3, scale=255.0, max_flow=2.0)
    for diffusion in ("image", "flow"):
        U, V = pipelines.flow_ad(I0.reshape(nr, nc, 3), I1.reshape(nr, nc, 3), backend(), diffusion=diffusion)
        assert np.isfinite(U).all() and aee(U, V, u, v) < 0.3, (diffusion, aee(U, V, u, v))
