"""CPU restatement of the stencil arithmetic the reference's MATLAB DRIVERS do between MEX calls
(SURVEY.md section 8a rows 17-21).  TEST INFRASTRUCTURE ONLY (see oracle/oracle.py for who may import it).

Two kinds of functions live here:

* formulas WRITTEN in the reference's .m files (OPdiffWeights, robust data weights and term assembly,
  ADdiffWeights, FAS restriction / right-hand side, lpf pyramid, rgb2grad): restated literally, each citing
  the lines it follows. These are pinned to the reference's source text.
* Image Processing Toolbox calls whose source is NOT in /root/reference (imresize, imfilter, medfilt2,
  fspecial): restated from their documented behaviour. PARITY UNPINNED: no reference fixture pins them;
  the GPU path is checked against this restatement, and the CPU and GPU pipelines share it so it cancels
  out of pipeline comparisons.

Arrays are numpy, indexed [row, col(, channel)] like Matlab; precision follows the .m code (double where
the driver computes in double, single where it computes in single).
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32


# ---------------------------------------------------------------------------------------------
# Toolbox functions (parity unpinned)
# ---------------------------------------------------------------------------------------------
def _pad_index(idx, n, mode):
    """Map out-of-range indices into [0, n) the way imfilter/medfilt2 pad: 'replicate' or 'symmetric'."""
    if mode == "replicate":
        return np.clip(idx, 0, n - 1)
    if mode == "symmetric":
        p = 2 * n
        m = np.mod(idx, p)
        return np.where(m < n, m, p - 1 - m)
    if mode == "circular":
        return np.mod(idx, n)
    raise ValueError(mode)


def imfilter(A, h, boundary="replicate", conv=False):
    """imfilter(A, h, boundary[, 'conv']): correlation (default) or convolution, 'same' size, accumulated in
    double, result in the class of A."""
    A = np.asarray(A)
    h = np.atleast_2d(np.asarray(h, dtype=np.float64))
    if conv:
        h = h[::-1, ::-1]
    kr, kc = h.shape
    assert kr % 2 == 1 and kc % 2 == 1, "odd-sized kernels only (all the drivers use)"
    cr, cc = (kr - 1) // 2, (kc - 1) // 2
    rows, cols = A.shape[:2]
    out = np.zeros(A.shape, dtype=np.float64)
    Ad = A.astype(np.float64)
    for a in range(kr):
        ri = _pad_index(np.arange(rows) + a - cr, rows, boundary)
        for b in range(kc):
            if h[a, b] == 0.0:
                continue
            ci = _pad_index(np.arange(cols) + b - cc, cols, boundary)
            out += h[a, b] * Ad[np.ix_(ri, ci)]
    return out.astype(A.dtype) if A.dtype in (np.float32, np.float64) else out


def fspecial_gaussian(size, sigma):
    """fspecial('gaussian', [size size], sigma)."""
    r = (size - 1) / 2.0
    y, x = np.mgrid[-r:r + 1, -r:r + 1]
    h = np.exp(-(x * x + y * y) / (2.0 * sigma * sigma))
    h[h < np.finfo(np.float64).eps * h.max()] = 0
    return h / h.sum()


def medfilt2_symmetric(A):
    """medfilt2(A, [3 3], 'symmetric')."""
    A = np.asarray(A)
    rows, cols = A.shape
    st = []
    for a in (-1, 0, 1):
        ri = _pad_index(np.arange(rows) + a, rows, "symmetric")
        for b in (-1, 0, 1):
            ci = _pad_index(np.arange(cols) + b, cols, "symmetric")
            st.append(A[np.ix_(ri, ci)])
    return np.sort(np.stack(st, axis=-1), axis=-1)[..., 4]


def imresize_contributions(in_len, out_len, scale, antialias=True, cubic=False):
    """indices (0-based) and weights of imresize's triangle ('bilinear') kernel along one dimension:
    output sample x (1-based) sits at u = x/scale + 0.5(1 - 1/scale); when shrinking with antialiasing the
    kernel is stretched by 1/scale; weights are normalised; out-of-range taps are mirrored."""
    kw = 4.0 if cubic else 2.0
    if scale < 1 and antialias:
        kw = kw / scale

    def base(x):
        ax = np.abs(x)
        if cubic:                            # imresize's cubic convolution kernel (a = -0.5)
            ax2, ax3 = ax * ax, ax * ax * ax
            return (1.5 * ax3 - 2.5 * ax2 + 1.0) * (ax <= 1) + (-0.5 * ax3 + 2.5 * ax2 - 4.0 * ax + 2.0) * ((1 < ax) & (ax <= 2))
        return np.maximum(0.0, 1.0 - ax)

    def h(x):
        if scale < 1 and antialias:
            return scale * base(x * scale)
        return base(x)

    x = np.arange(1, out_len + 1, dtype=np.float64)
    u = x / scale + 0.5 * (1.0 - 1.0 / scale)
    left = np.floor(u - kw / 2.0)
    P = int(math.ceil(kw)) + 2
    ind = left[:, None] + np.arange(P)[None, :]              # 1-based
    w = h(u[:, None] - ind)
    w = w / w.sum(axis=1, keepdims=True)
    aux = np.concatenate([np.arange(1, in_len + 1), np.arange(in_len, 0, -1)])
    ind = aux[np.mod(ind.astype(np.int64) - 1, aux.size)]    # mirror
    return ind - 1, w


def imresize_bicubic(A, scale=None, output_size=None, antialias=True):
    """imresize(A, ...) with the default 'bicubic' method."""
    return imresize_bilinear(A, scale, output_size, antialias, cubic=True)


def imresize_bilinear(A, scale=None, output_size=None, antialias=True, cubic=False):
    """imresize(A, scale, 'bilinear') / imresize(A, 'OutputSize', [r c], 'Method', 'triangle'|'bilinear').
    The dimension with the smaller scale factor is resized first (rows first on a tie)."""
    A = np.asarray(A)
    rows, cols = A.shape[:2]
    if output_size is None:
        orows, ocols = int(math.ceil(rows * scale)), int(math.ceil(cols * scale))
        sr = sc = float(scale)
    else:
        orows, ocols = int(output_size[0]), int(output_size[1])
        sr, sc = orows / rows, ocols / cols
    dt = A.dtype if A.dtype in (np.float32, np.float64) else np.float64
    out = A.astype(dt)
    order = [0, 1] if sr <= sc else [1, 0]
    for dim in order:                                        # each pass: double accumulation, result in the class of A
        if dim == 0:
            ind, w = imresize_contributions(rows, orows, sr, antialias, cubic)
            acc = np.zeros((orows,) + out.shape[1:], dtype=np.float64)
            for p in range(w.shape[1]):
                acc += w[:, p].reshape((-1,) + (1,) * (out.ndim - 1)) * out[ind[:, p]].astype(np.float64)
        else:
            ind, w = imresize_contributions(cols, ocols, sc, antialias, cubic)
            acc = np.zeros((out.shape[0], ocols) + out.shape[2:], dtype=np.float64)
            for p in range(w.shape[1]):
                acc += w[:, p].reshape((1, -1) + (1,) * (out.ndim - 2)) * out[:, ind[:, p]].astype(np.float64)
        out = acc.astype(dt)
    return out


# ---------------------------------------------------------------------------------------------
# Formulas written in the reference's drivers (pinned to the .m text)
# ---------------------------------------------------------------------------------------------
def rgb2grad(IN):
    """rgb2grad, matlab/optical_flow/FlowEminND_llin_2D_v10.m:374-384: per channel, [1 0 -1] correlation along
    columns then along rows, replicate border; output channels interleaved (dx, dy)."""
    IN = np.asarray(IN)
    if IN.ndim == 2:
        IN = IN[:, :, None]
    out = np.zeros(IN.shape[:2] + (2 * IN.shape[2],), dtype=IN.dtype)
    odx = np.array([[1.0, 0.0, -1.0]])
    for i in range(IN.shape[2]):
        out[:, :, 2 * i] = imfilter(IN[:, :, i], odx, "replicate")
        out[:, :, 2 * i + 1] = imfilter(IN[:, :, i], odx.T, "replicate")
    return out


def op_diff_weights(U, V):
    """OPdiffWeights, FlowEminND_llin_2D_v10.m:389-433 (duplicate: FlowEminNDFASFMG_elin_2D_v10.m:469-514).
    Computed in double with circshift (periodic wrap). Returns (wW, wN, wS, wE) as the .m function does."""
    U = np.asarray(U, dtype=np.float64)
    V = np.asarray(V, dtype=np.float64)
    k = np.array([[0.25, 0.0, -0.25]])
    Uver, Vver = imfilter(U, k.T, "replicate"), imfilter(V, k.T, "replicate")
    Uhor, Vhor = imfilter(U, k, "replicate"), imfilter(V, k, "replicate")

    def cs(a, dr, dc):                    # circshift(a, [dr dc])
        return np.roll(np.roll(a, dr, axis=0), dc, axis=1)

    wW = (cs(U, 0, 1) - U) ** 2 + (Uver + cs(Uver, 0, 1)) ** 2 + (cs(V, 0, 1) - V) ** 2 + (Vver + cs(Vver, 0, 1)) ** 2
    wE = (cs(U, 0, -1) - U) ** 2 + (Uver + cs(Uver, 0, -1)) ** 2 + (cs(V, 0, -1) - V) ** 2 + (Vver + cs(Vver, 0, -1)) ** 2
    wN = (cs(U, 1, 0) - U) ** 2 + (Uhor + cs(Uhor, 1, 0)) ** 2 + (cs(V, 1, 0) - V) ** 2 + (Vhor + cs(Vhor, 1, 0)) ** 2
    wS = (cs(U, -1, 0) - U) ** 2 + (Uhor + cs(Uhor, -1, 0)) ** 2 + (cs(V, -1, 0) - V) ** 2 + (Vhor + cs(Vhor, -1, 0)) ** 2
    f = lambda w: 1.0 / np.sqrt(w + 0.00001)
    return f(wW), f(wN), f(wS), f(wE)


def _nansum3(parts):
    """nansum(cat(3, parts...), 3) in single precision, channels accumulated in order."""
    acc = None
    for p in parts:
        p = np.asarray(p, dtype=F32)
        if p.ndim == 2:
            p = p[:, :, None]
        for c in range(p.shape[2]):
            t = np.where(np.isnan(p[:, :, c]), F32(0), p[:, :, c])
            acc = t.copy() if acc is None else (acc + t).astype(F32)
    return acc


def llin_terms(d1, d2, dU, dV, b1, b2, alpha, gradmag):
    """Robust data weights and data-term assembly of the late-linearisation flow driver,
    FlowEminND_llin_2D_v10.m:235-258 (products), :289-299 (gD1, gD2), :323-327 (nansum over channels).
    d1 = (I1dt, I1dx, I1dy), d2 = (I2dt, I2dx, I2dy) or, with gradmag, (I2dxt, I2dyt, I2dxx, I2dyy, I2dxy),
    or None. Single precision throughout, as in the driver. Returns (M, Cu, Cv, Du, Dv)."""
    dU = np.asarray(dU, dtype=F32)[:, :, None]
    dV = np.asarray(dV, dtype=F32)[:, :, None]
    a = F32(alpha)
    I1dt, I1dx, I1dy = [np.asarray(x, dtype=F32).reshape(dU.shape[0], dU.shape[1], -1) for x in d1]
    M1, Cu1, Cv1, Du1, Dv1 = I1dy * I1dx, I1dt * I1dx, I1dt * I1dy, I1dx * I1dx, I1dy * I1dy
    OP = (I1dt - I1dx * dU - I1dy * dV) ** 2
    g1 = (F32(b1) / (a * np.sqrt(OP + F32(0.00001)))).astype(F32)
    parts = {"M": [M1 * g1], "Cu": [Cu1 * g1], "Cv": [Cv1 * g1], "Du": [Du1 * g1], "Dv": [Dv1 * g1]}
    if d2 is not None:
        d2 = [np.asarray(x, dtype=F32).reshape(dU.shape[0], dU.shape[1], -1) for x in d2]
        if not gradmag:
            I2dt, I2dx, I2dy = d2
            M2, Cu2, Cv2, Du2, Dv2 = I2dy * I2dx, I2dt * I2dx, I2dt * I2dy, I2dx * I2dx, I2dy * I2dy
            OP = (I2dt - I2dx * dU - I2dy * dV) ** 2
        else:
            xt, yt, xx, yy, xy = d2
            M2 = xy * (xx + yy)
            Cu2 = xt * xx + yt * xy
            Cv2 = xt * xy + yt * yy
            Du2 = xx * xx + xy * xy
            Dv2 = xy * xy + yy * yy
            OP = (xt - xx * dU - xy * dV) ** 2 + (yt - xy * dU - yy * dV) ** 2
        g2 = (F32(b2) / (a * np.sqrt(OP + F32(0.00001)))).astype(F32)
        for k, v in zip(("M", "Cu", "Cv", "Du", "Dv"), (M2, Cu2, Cv2, Du2, Dv2)):
            parts[k].append(v * g2)
    return tuple(_nansum3(parts[k]) for k in ("M", "Cu", "Cv", "Du", "Dv"))


def elin_terms(der, coef, U, V, b1, b2, alpha, summed):
    """Robust weight gd and weighted terms of the FMG early-linearisation smoother,
    FlowEminNDFASFMG_elin_2D_v10.m:375-396 (summed over channels, gd carries 1/channels) and :421-440
    (per channel, for the residual). der = (Idt, Idx, Idy, Idxt, Idyt, Idxx, Idyy, Idxy),
    coef = (M, Cu, Cv, Du, Dv); all rows x cols x channels. Returns (gd, M, Cu, Cv, Du, Dv)."""
    U = np.asarray(U, dtype=F32)[:, :, None]
    V = np.asarray(V, dtype=F32)[:, :, None]
    Idt, Idx, Idy, Idxt, Idyt, Idxx, Idyy, Idxy = [np.asarray(x, dtype=F32).reshape(U.shape[0], U.shape[1], -1) for x in der]
    ch = Idx.shape[2]
    OP = F32(b1) * (Idt - Idx * U - Idy * V) ** 2 + F32(b2) * ((Idxt - Idxx * U - Idxy * V) ** 2 + (Idyt - Idxy * U - Idyy * V) ** 2)
    fac = F32((ch if summed else 1) * alpha)
    gd = (F32(1) / (fac * np.sqrt(OP + F32(0.00001)))).astype(F32)
    out = []
    for c in coef:
        c = np.asarray(c, dtype=F32).reshape(U.shape[0], U.shape[1], -1)
        t = (c * gd).astype(F32)
        if summed:
            acc = t[:, :, 0].copy()
            for k in range(1, ch):
                acc = (acc + t[:, :, k]).astype(F32)
            t = acc
        out.append(t)
    return (gd,) + tuple(out)


def fw_restrict(A, scl_factor):
    """Full-weighting restriction of the FAS cycle, FlowEminNDFASFMG_elin_2D_v10.m:199,212-217:
    imfilter(A*scl_factor, fw, 'replicate', 'conv') then (1:2:end, 1:2:end)."""
    fw = np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=np.float64) / 16.0
    A = np.asarray(A)
    s = (A * A.dtype.type(scl_factor)) if A.dtype == F32 else A * scl_factor
    return imfilter(s, fw, "replicate", conv=True)[::2, ::2]


def lpf_decimate(A):
    """One pyramid level of the FMG driver, FlowEminNDFASFMG_elin_2D_v10.m:98,107-110: [1 4 6 4 1]/16 along
    columns, then along rows, replicate border, then (1:2:end, 1:2:end)."""
    lpf = np.array([[1, 4, 6, 4, 1]], dtype=np.float64) / 16.0
    t = imfilter(imfilter(A, lpf, "replicate", conv=True), lpf.T, "replicate", conv=True)
    return t[::2, ::2]


def fas_rhs(R, A, gd):
    """Coarse-grid right-hand side of the FAS cycle, FlowEminNDFASFMG_elin_2D_v10.m:250-251: (R + A)./gd."""
    return ((np.asarray(R, dtype=F32) + np.asarray(A, dtype=F32)) / np.asarray(gd, dtype=F32)).astype(F32)


def ad_diff_weights(D, quantile=0.5, flow_variant=False):
    """ADdiffWeights, matlab/denoising/TVdenoise8.m:119-231 (Alvarez derivatives, lambda = the `quantile`
    order statistic of the non-zero squared gradient norms). Double precision.
    flow_variant: the copy in matlab/optical_flow/FlowEminAD_llin_2D_v10.m:416-488 -- the order statistic is
    round(numel*quantile) without the eps (:459) and the weights towards the missing neighbours of border pixels are
    NOT zeroed (:472-479; the line solvers never read them).
    Returns (W, NW, N, NE, E, SE, S, SW, lambda), each rows x cols."""
    D = np.asarray(D, dtype=np.float64)
    if D.ndim == 2:
        D = D[:, :, None]
    s2 = math.sqrt(2.0)
    O_dx = np.array([[1, 0, -1], [s2, 0, -s2], [1, 0, -1]]) / (4 + math.sqrt(8.0))
    O_dy = np.array([[1, s2, 1], [0, 0, 0], [-1, -s2, -1]]) / (4 + math.sqrt(8.0))
    Ddx = np.stack([imfilter(D[:, :, k], O_dx, "replicate", conv=True) for k in range(D.shape[2])], axis=2)
    Ddy = np.stack([imfilter(D[:, :, k], O_dy, "replicate", conv=True) for k in range(D.shape[2])], axis=2)
    if D.shape[2] > 1:
        ind = np.argmax(Ddx ** 2 + Ddy ** 2, axis=2)          # first maximum, like Matlab's max
        mx = np.take_along_axis(Ddx, ind[:, :, None], axis=2)[:, :, 0]
        my = np.take_along_axis(Ddy, ind[:, :, None], axis=2)[:, :, 0]
    else:
        mx, my = Ddx[:, :, 0], Ddy[:, :, 0]
    nrm = mx ** 2 + my ** 2
    srt = np.sort(nrm.reshape(-1, order="F"))
    srt = srt[srt != 0]
    if srt.size:
        k = int(np.floor(srt.size * quantile + (0.0 if flow_variant else np.finfo(np.float64).eps) + 0.5))   # Matlab round()
        lam = srt[k - 1]
    else:
        lam = 1.0
    mul = 1.0 / (nrm + 2 * lam)
    dyy, dxx, dxy = mul * (my ** 2 + lam), mul * (mx ** 2 + lam), -mul * (mx * my)

    def cs(a, dr, dc):
        return np.roll(np.roll(a, dr, axis=0), dc, axis=1)

    if flow_variant:
        return (0.5 * (dyy + cs(dyy, 0, 1)), 0.25 * (dxy + cs(dxy, 1, 1)), 0.5 * (dxx + cs(dxx, 1, 0)), -0.25 * (dxy + cs(dxy, 1, -1)),
                0.5 * (dyy + cs(dyy, 0, -1)), 0.25 * (dxy + cs(dxy, -1, -1)), 0.5 * (dxx + cs(dxx, -1, 0)), -0.25 * (dxy + cs(dxy, -1, 1)), lam)
    W = 0.5 * (dyy + cs(dyy, 0, 1));      W[:, 0] = 0
    NW = 0.25 * (dxy + cs(dxy, 1, 1));    NW[:, 0] = 0;   NW[0, :] = 0
    N = 0.5 * (dxx + cs(dxx, 1, 0));      N[0, :] = 0
    NE = -0.25 * (dxy + cs(dxy, 1, -1));  NE[:, -1] = 0;  NE[0, :] = 0
    E = 0.5 * (dyy + cs(dyy, 0, -1));     E[:, -1] = 0
    SE = 0.25 * (dxy + cs(dxy, -1, -1));  SE[:, -1] = 0;  SE[-1, :] = 0
    S = 0.5 * (dxx + cs(dxx, -1, 0));     S[-1, :] = 0
    SW = -0.25 * (dxy + cs(dxy, -1, 1));  SW[-1, :] = 0;  SW[:, 0] = 0
    return W, NW, N, NE, E, SE, S, SW, lam


def tv_terms(Iout, Iin, weights, alpha):
    """TVdenoise8.m:83-85 for SINGLE images (runme.m:118,144 passes single): PsiData and B are computed in single,
    TRACE = single PsiData + double alpha*(sum of weights) -> single; plus the single(alpha*w) handed to PDEsolver8."""
    Iout = np.asarray(Iout, dtype=F32)
    Iin = np.asarray(Iin, dtype=F32)
    df = (Iout - Iin).astype(F32)
    psi = (F32(1) / np.sqrt((df * df).astype(F32) + F32(np.finfo(np.float64).eps))).astype(F32)
    sw = sum(weights)
    if psi.ndim == 3:                      # ADdiffWeights repmat's its weights over the frames (:221-230)
        sw = sw[:, :, None]
    tr = (psi + (alpha * sw).astype(F32)).astype(F32)
    return tr, (psi * Iin).astype(F32), [np.asarray(alpha * w, dtype=F32) for w in weights]


def disp_sym_terms(d, dU, Udt, Udx, b1, b2, alpha, beta, srDiff):
    """Robust data + symmetry weights and term assembly of the symmetric stereo driver for ONE view,
    matlab/disparity/DispEminND_llin_sym_2D.m:172-180 (CuD, DuD, CuS, DuS), :197-210 (gD, gSYM), :222-225 (sums).
    d = (Idt, Idx, Idxt, Idyt, Idxx, Idxy) of that view (rows x cols x channels). Returns (CuG, DuG)."""
    dU = np.asarray(dU, dtype=F32)
    Idt, Idx, Idxt, Idyt, Idxx, Idxy = [np.asarray(x, dtype=F32).reshape(dU.shape[0], dU.shape[1], -1) for x in d]
    ch = Idx.shape[2]
    Udt = np.asarray(Udt, dtype=F32)
    Udx = np.asarray(Udx, dtype=F32)
    CuD = F32(b1) * Idt * Idx + F32(b2) * (Idxt * Idxx + Idyt * Idxy)
    DuD = F32(b1) * Idx * Idx + F32(b2) * (Idxx * Idxx + Idxy * Idxy)
    CuS = Udt * (F32(1) + Udx)
    DuS = F32(1) + Udx + Udx + Udx * Udx
    d3 = dU[:, :, None]
    OP = F32(b1) * (Idt - Idx * d3) ** 2 + F32(b2) * ((Idxt - Idxx * d3) ** 2 + (Idyt - Idxy * d3) ** 2)
    gD = (F32(1) / (F32(alpha) * np.sqrt(OP + F32(0.00001)))).astype(F32)
    Sn = (dU + Udt + Udx * dU) ** 2
    gS = (F32(ch * beta / alpha) / (F32(1) + Sn / F32(srDiff ** 2))).astype(F32)

    def s3(parts):
        acc = None
        for p in parts:
            p = p if p.ndim == 3 else p[:, :, None]
            for c in range(p.shape[2]):
                acc = p[:, :, c].astype(F32).copy() if acc is None else (acc + p[:, :, c]).astype(F32)
        return acc
    return s3([gD * CuD, -gS * CuS]), s3([gD * DuD, gS * DuS])


def fmg_derivatives(It0, It1):
    """Derivative stacks of one level of the FMG driver, FlowEminNDFASFMG_elin_2D_v10.m:81-141 (single images 0..255).
    Returns (Idt, Idx, Idy, Idxt, Idyt, Idxx, Idyy, Idxy)."""
    pre = np.array([[0.037659, 0.249724, 0.439911, 0.249724, 0.037659]])
    odx = np.array([[0.104550, 0.292315, 0.0, -0.292315, -0.104550]])
    odx_s = odx / 255
    oxx = np.array([[0.232905, 0.002668, -0.471147, 0.002668, 0.232905]])
    It0 = np.asarray(It0, dtype=F32); It1 = np.asarray(It1, dtype=F32)
    f = lambda A, h: imfilter(A, h, "replicate", conv=True)
    Ist = (((It0 + It1).astype(F32) * F32(0.55)).astype(F32) / F32(255)).astype(F32)
    Idt = ((It0 - It1).astype(F32) / F32(255)).astype(F32)
    Idx = f(f(Ist, pre.T), odx); Idy = f(f(Ist, pre), odx.T)
    Idxx = f(f(Ist, pre.T), oxx); Idyy = f(f(Ist, pre), oxx.T)
    Idxy = f(f(Ist, odx), odx.T)
    Idxt = (f(f(It0, pre.T), odx_s) - f(f(It1, pre.T), odx_s)).astype(F32)
    Idyt = (f(f(It0, pre), odx_s.T) - f(f(It1, pre), odx_s.T)).astype(F32)
    return Idt, Idx, Idy, Idxt, Idyt, Idxx, Idyy, Idxy


def fmg_terms(der, b1, b2):
    """M, Cu, Cv, Du, Dv of FlowEminNDFASFMG_elin_2D_v10.m:143-149 (single, evaluated left to right)."""
    Idt, Idx, Idy, Idxt, Idyt, Idxx, Idyy, Idxy = [np.asarray(x, dtype=F32) for x in der]
    b1, b2 = F32(b1), F32(b2)
    M = b1 * Idy * Idx + b2 * Idxy * (Idxx + Idyy)
    Cu = b1 * Idt * Idx + b2 * (Idxt * Idxx + Idyt * Idxy)
    Cv = b1 * Idt * Idy + b2 * (Idxt * Idxy + Idyt * Idyy)
    Du = b1 * Idx * Idx + b2 * (Idxx * Idxx + Idxy * Idxy)
    Dv = b1 * Idy * Idy + b2 * (Idxy * Idxy + Idyy * Idyy)
    return M, Cu, Cv, Du, Dv


def interp2_rows(Vals, Xq):
    """interp2(X, Y, Vals, Xq, Y) with X, Y = meshgrid(1:cols, 1:rows) and the query on the grid's own rows: linear
    interpolation along each row at the 1-based column positions Xq, NaN outside [1, cols] (and for NaN queries).
    DispEminND_llin_sym_2D.m:143-144. Toolbox function: parity unpinned; double arithmetic, result in the class of Vals."""
    Vals = np.asarray(Vals)
    rows, cols = Vals.shape
    xq = np.asarray(Xq, dtype=np.float64)
    ok = (xq >= 1.0) & (xq <= cols)
    xs = np.where(ok, xq, 1.0)
    j0 = np.minimum(np.floor(xs).astype(np.int64), cols - 1)          # 1-based left sample; the last column uses t = 1
    t = xs - j0
    r = np.arange(rows)[:, None]
    v0 = Vals[r, j0 - 1].astype(np.float64)
    v1 = Vals[r, j0].astype(np.float64)
    out = v0 * (1.0 - t) + v1 * t
    return np.where(ok, out, np.nan).astype(Vals.dtype)


def round_uint8(A):
    """class uint8 after a toolbox call (imresize, imfilter on uint8 images): round half away from zero, saturate."""
    return np.clip(np.floor(np.asarray(A, dtype=np.float64) + 0.5), 0, 255).astype(F32)


def tv4_diff_weights(D):
    """DiffWeights of matlab/denoising/TVdenoise4.m:116-148: the 4-neighbour edge weights of OPdiffWeights for ONE
    field with several frames, maximum over the frames (:131-134), outward border edges zeroed (:145-148). Computed in the class of D (single for the images
    runme.m:143 passes): one rounding per operation. Returns (wW, wN, wE, wS), rows x cols."""
    D = np.asarray(D, dtype=F32)
    if D.ndim == 2:
        D = D[:, :, None]
    q = np.array([[0.25, 0.0, -0.25]])
    Dver = np.stack([imfilter(D[:, :, k], q.T, "replicate") for k in range(D.shape[2])], axis=2).astype(F32)
    Dhor = np.stack([imfilter(D[:, :, k], q, "replicate") for k in range(D.shape[2])], axis=2).astype(F32)

    def cs(a, dr, dc):
        return np.roll(np.roll(a, dr, axis=0), dc, axis=1)

    def edge(dr, dc, G):
        a = (cs(D, dr, dc) - D).astype(F32); b = (G + cs(G, dr, dc)).astype(F32)
        return ((a * a).astype(F32) + (b * b).astype(F32)).astype(F32).max(axis=2)

    wW, wE = edge(0, 1, Dver), edge(0, -1, Dver)
    wN, wS = edge(1, 0, Dhor), edge(-1, 0, Dhor)
    f = lambda w: (F32(1) / np.sqrt((w + F32(0.00001)).astype(F32))).astype(F32)
    wW, wN, wE, wS = f(wW), f(wN), f(wE), f(wS)
    wW[:, 0] = 0; wE[:, -1] = 0; wN[0, :] = 0; wS[-1, :] = 0                # :145-148: no edge leaves the image
    return wW, wN, wE, wS


def disp_terms(d1, d2, dU, b1, b2, alpha, gradmag):
    """Robust weights and channel sums of matlab/disparity/DispEminND_llin_2D.m:251-280 (single precision; plain sum over
    the channels, :279-280: a NaN of a warped pixel reaches CuGd / DuGd). d1 = (I1dt, I1dx, I1dy); d2 = None, the three
    first derivatives or the five second derivatives (gradmag). Returns (CuGd, DuGd)."""
    dU = np.asarray(dU, dtype=F32)[:, :, None]
    I1dt, I1dx = d1[0], d1[1]
    r = (I1dt - (I1dx * dU).astype(F32)).astype(F32)
    g1 = (F32(b1) / (F32(alpha) * np.sqrt(((r * r).astype(F32) + F32(0.00001)).astype(F32))).astype(F32)).astype(F32)
    Cu = [((I1dt * I1dx).astype(F32) * g1).astype(F32)]
    Du = [((I1dx * I1dx).astype(F32) * g1).astype(F32)]
    if d2 is not None:
        if gradmag:
            xt, yt, xx, _, xy = d2
            n2 = (((xt - (xx * dU).astype(F32)).astype(F32) ** 2).astype(F32) + ((yt - (xy * dU).astype(F32)).astype(F32) ** 2).astype(F32)).astype(F32)
            c2 = ((xt * xx).astype(F32) + (yt * xy).astype(F32)).astype(F32)
            e2 = ((xx * xx).astype(F32) + (xy * xy).astype(F32)).astype(F32)
        else:
            t, x = d2[0], d2[1]
            n2 = ((t - (x * dU).astype(F32)).astype(F32) ** 2).astype(F32)
            c2, e2 = (t * x).astype(F32), (x * x).astype(F32)
        g2 = (F32(b2) / (F32(alpha) * np.sqrt((n2 + F32(0.00001)).astype(F32))).astype(F32)).astype(F32)
        Cu.append((c2 * g2).astype(F32)); Du.append((e2 * g2).astype(F32))

    def s3(parts):
        acc = None
        for p_ in parts:
            for k in range(p_.shape[2]):
                acc = p_[:, :, k].copy() if acc is None else (acc + p_[:, :, k]).astype(F32)
        return acc

    return s3(Cu), s3(Du)
