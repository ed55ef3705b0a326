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
