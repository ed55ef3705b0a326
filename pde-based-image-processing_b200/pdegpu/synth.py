"""Seeded synthetic inputs for the solver core.

Structurally valid systems only: edge-symmetric non-negative weights (wE(p) == wW(p+east),
wS(p) == wN(p+south)), outward border edges 0, positive semi-definite data term
(M^2 <= Du*Dv), a small fraction of NaN data terms to exercise the "no data term" branch.
Random unstructured systems diverge at omega = 1.9.

All arrays are float32, column-major (order='F'), shaped (nrows, ncols[, nframes]).
"""
from __future__ import annotations

import numpy as np


def f32(x) -> np.ndarray:
    return np.asfortranarray(np.asarray(x, dtype=np.float32))


def edge_weights(rng, nrows, ncols, lo=0.2, span=3.0, eight=False):
    """Edge-symmetric diffusion weights; returns dict with wW,wN,wE,wS (and diagonals)."""
    h = lo + span * rng.random((nrows, ncols - 1))          # edge between (i,j) and (i,j+1)
    v = lo + span * rng.random((nrows - 1, ncols))          # edge between (i,j) and (i+1,j)
    w = {k: np.zeros((nrows, ncols), np.float32) for k in ("wW", "wN", "wE", "wS")}
    w["wE"][:, :-1] = h
    w["wW"][:, 1:] = h
    w["wS"][:-1, :] = v
    w["wN"][1:, :] = v
    if eight:
        d1 = 0.5 * (lo + span * rng.random((nrows - 1, ncols - 1)))   # (i,j)-(i+1,j+1)
        d2 = 0.5 * (lo + span * rng.random((nrows - 1, ncols - 1)))   # (i+1,j)-(i,j+1)
        for k in ("wNW", "wNE", "wSE", "wSW"):
            w[k] = np.zeros((nrows, ncols), np.float32)
        w["wSE"][:-1, :-1] = d1
        w["wNW"][1:, 1:] = d1
        w["wNE"][1:, :-1] = d2
        w["wSW"][:-1, 1:] = d2
    return {k: f32(a) for k, a in w.items()}


def data_terms(rng, nrows, ncols, nframes=1, nan_frac=0.01, alpha=1.0):
    shape = (nrows, ncols) if nframes == 1 else (nrows, ncols, nframes)
    Ix = rng.random(shape) - 0.5
    Iy = rng.random(shape) - 0.5
    It = rng.random(shape) - 0.5
    g = np.minimum(1.0 / (alpha * np.sqrt(It * It + 1e-5)), 50.0)
    M, Du, Dv = g * Ix * Iy, g * Ix * Ix, g * Iy * Iy
    Cu, Cv = -g * It * Ix, -g * It * Iy
    if nan_frac > 0:
        mask = rng.random(shape) < nan_frac
        for a in (M, Du, Dv, Cu, Cv):
            a[mask] = np.nan
    return {k: f32(a) for k, a in dict(M=M, Cu=Cu, Cv=Cv, Du=Du, Dv=Dv).items()}


def smooth_field(rng, nrows, ncols, amp=1.0, nwaves=4):
    ii, jj = np.meshgrid(np.arange(nrows), np.arange(ncols), indexing="ij")
    out = np.zeros((nrows, ncols))
    for _ in range(nwaves):
        fx, fy = rng.uniform(0.5, 4.0, 2) * 2 * np.pi
        ph = rng.uniform(0, 2 * np.pi)
        out += rng.uniform(0.2, 1.0) * np.sin(fx * jj / ncols + fy * ii / nrows + ph)
    return f32(amp * out / nwaves)


def flow_system(seed, nrows, ncols, late=False, eight=False, nframes=1, nan_frac=0.01):
    """Inputs of Oflow_sor_elin4_2d / Oflow_sor_llin4_2d / Oflow_sor_llin8_2d as a dict."""
    rng = np.random.default_rng(seed)
    s = {}
    s.update(edge_weights(rng, nrows, ncols, eight=eight))
    s.update(data_terms(rng, nrows, ncols, nframes=nframes, nan_frac=nan_frac))
    if late:
        s["U"] = smooth_field(rng, nrows, ncols, 2.0)
        s["V"] = smooth_field(rng, nrows, ncols, 2.0)
        s["dU"] = f32(0.1 * (rng.random((nrows, ncols)) - 0.5))
        s["dV"] = f32(0.1 * (rng.random((nrows, ncols)) - 0.5))
    else:
        s["U"] = f32(0.1 * (rng.random((nrows, ncols)) - 0.5))
        s["V"] = f32(0.1 * (rng.random((nrows, ncols)) - 0.5))
    return s


def disp_system(seed, nrows, ncols, nan_frac=0.01):
    rng = np.random.default_rng(seed)
    s = edge_weights(rng, nrows, ncols)
    d = data_terms(rng, nrows, ncols, nan_frac=nan_frac)
    s["Cu"], s["Du"] = d["Cu"], d["Du"]
    s["U"] = smooth_field(rng, nrows, ncols, 4.0)
    s["dU"] = f32(0.1 * (rng.random((nrows, ncols)) - 0.5))
    return s


def pde_system(seed, nrows, ncols, nframes=1, eight=False, nan_frac=0.01):
    """TV-denoising-like system: TRACE = sum of weights + data weight, B = data weight * image."""
    rng = np.random.default_rng(seed)
    keys = ("wW", "wN", "wE", "wS") + (("wNW", "wNE", "wSE", "wSW") if eight else ())
    frames = [edge_weights(rng, nrows, ncols, eight=eight) for _ in range(nframes)]
    shape = (nrows, ncols) if nframes == 1 else (nrows, ncols, nframes)
    s = {}
    for k in keys:
        a = np.stack([fr[k] for fr in frames], axis=2)
        s[k] = f32(a.reshape(shape))
    img = rng.random(shape)
    psi = 0.5 + rng.random(shape)
    sw = sum(s[k].astype(np.float64) for k in keys)
    TRACE = sw + psi
    B = psi * img
    if nan_frac > 0:
        mask = rng.random(shape) < nan_frac
        TRACE[mask] = np.nan
    s["TRACE"], s["B"] = f32(TRACE), f32(B)
    s["X"] = f32(img + 0.05 * rng.standard_normal(shape))
    return s


def image_pair(seed, nrows, ncols, nframes=1, scale=1.0, max_flow=3.0, horizontal=False):
    """Smooth random texture and a second frame shifted by a known smooth flow (u along columns,
    v along rows), sampled analytically so the pair is exact. horizontal: v = 0 (a stereo pair)."""
    rng = np.random.default_rng(seed)
    ii, jj = np.meshgrid(np.arange(nrows, dtype=np.float64), np.arange(ncols, dtype=np.float64), indexing="ij")
    waves = [(rng.uniform(2, 30) * 2 * np.pi / ncols, rng.uniform(2, 30) * 2 * np.pi / nrows,
              rng.uniform(0, 2 * np.pi), rng.uniform(0.3, 1.0)) for _ in range(6 * nframes)]

    def tex(x, y, k):
        out = np.zeros_like(x)
        for (fx, fy, ph, am) in waves[6 * k:6 * k + 6]:
            out += am * np.sin(fx * x + fy * y + ph)
        return out

    u = max_flow * (0.5 * (jj / ncols - 0.5) + 0.3 * np.sin(2 * np.pi * ii / nrows))
    v = max_flow * (0.4 * (ii / nrows - 0.5) + 0.3 * np.cos(2 * np.pi * jj / ncols))
    if horizontal:
        v = np.zeros_like(u)
    I0 = np.stack([tex(jj, ii, k) for k in range(nframes)], axis=2)
    I1 = np.stack([tex(jj - u, ii - v, k) for k in range(nframes)], axis=2)
    lo, hi = I0.min(), I0.max()
    I0 = scale * (I0 - lo) / (hi - lo)
    I1 = scale * (I1 - lo) / (hi - lo)
    if nframes == 1:
        I0, I1 = I0[:, :, 0], I1[:, :, 0]
    return f32(I0), f32(I1), f32(u), f32(v)


def mex_args(fn: str, s: dict, iter_=4, omega=1.9, solver=2):
    """Argument list of MEX function `fn` from a system dict (scalars as single, like the drivers pass them)."""
    one = lambda v: f32([[v]])
    tail = [one(iter_), one(omega), one(solver)]
    if fn == "Oflow_sor_elin4_2d":
        return [s[k] for k in ("U", "V", "M", "Cu", "Cv", "Du", "Dv", "wW", "wN", "wE", "wS")] + tail
    if fn == "Oflow_sor_llin4_2d":
        return [s[k] for k in ("U", "V", "dU", "dV", "M", "Cu", "Cv", "Du", "Dv", "wW", "wN", "wE", "wS")] + tail
    if fn == "Oflow_sor_llin8_2d":
        return [s[k] for k in ("U", "V", "dU", "dV", "M", "Cu", "Cv", "Du", "Dv",
                               "wW", "wNW", "wN", "wNE", "wE", "wSE", "wS", "wSW")] + tail
    if fn == "Oflow_lhs_elin4_2d":
        return [s[k] for k in ("U", "V", "M", "Du", "Dv", "wW", "wN", "wE", "wS")]
    if fn == "Oflow_lhs_llin4_2d":
        return [s[k] for k in ("U", "V", "dU", "dV", "M", "Du", "Dv", "wW", "wN", "wE", "wS")]
    if fn == "Disp_sor_llin4_2d":
        return [s[k] for k in ("U", "dU", "Cu", "Du", "wW", "wN", "wE", "wS")] + tail
    if fn == "Disp_sor_llin_sym4_2d":
        s0, s1 = s["f0"], s["f1"]
        keys = ("U", "dU", "Cu", "Du", "wW", "wN", "wE", "wS")
        return [s0[k] for k in keys] + [s1[k] for k in keys] + tail
    if fn == "PDEsolver4":
        return [s[k] for k in ("X", "TRACE", "B", "wW", "wN", "wE", "wS")] + tail
    if fn == "PDEsolver8":
        return [s[k] for k in ("X", "TRACE", "B", "wW", "wNW", "wN", "wNE", "wE", "wSE", "wS", "wSW")] + tail
    raise KeyError(fn)
