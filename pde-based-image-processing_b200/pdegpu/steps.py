"""numpy-in / numpy-out wrappers of libpdegpu's driver-side stencils (include/pdegpu.h, "Driver-side
stencils"): the steps the reference's Matlab drivers perform between MEX calls, SURVEY.md 8a rows 17-21.
Used by the parity tests; torch only allocates the device buffers. Arrays are indexed [row, col(, channel)]
and travel in Matlab's column-major layout."""
from __future__ import annotations

import ctypes
import math

import numpy as np

from . import lib


def _dev(a):
    """numpy [r, c(, ch)] -> flat device tensor in column-major (channel-slowest) order."""
    import torch
    a = np.asarray(a, dtype=np.float32)
    return torch.from_numpy(np.ascontiguousarray(a.reshape(-1, order="F"))).to("cuda:0")


def _host(t, shape):
    return t.cpu().numpy().reshape(shape, order="F")


def _empty(n):
    import torch
    return torch.empty(int(n), dtype=torch.float32, device="cuda:0")


class Steps:
    def __init__(self, ctx: lib.Context | None = None):
        self.ctx = ctx or lib.Context(0)
        self.L = lib.dll()

    def _run(self, rc):
        self.ctx._chk(rc)
        self.ctx.sync()

    def op_diff_weights(self, U, V):
        nr, nc = U.shape
        u, v = _dev(U), _dev(V)
        o = [_empty(nr * nc) for _ in range(4)]
        self._run(self.L.pdegpu_dev_op_diff_weights(self.ctx.h, *[t.data_ptr() for t in o], u.data_ptr(), v.data_ptr(), nr, nc, 1, nr * nc))
        return tuple(_host(t, (nr, nc)) for t in o)          # wW, wN, wS, wE

    def llin_terms(self, d1, d2, dU, dV, b1, b2, alpha, gradmag):
        nr, nc = dU.shape
        t = lib.LlinTerms()
        keep = []

        def stack(x):
            x = np.asarray(x, dtype=np.float32).reshape(nr, nc, -1)
            keep.append(_dev(x))
            return keep[-1].data_ptr(), x.shape[2]
        for k, x in enumerate(d1):
            t.d1[k], t.channels1 = stack(x)
        t.channels2 = 0
        if d2 is not None:
            for k, x in enumerate(d2):
                t.d2[k], t.channels2 = stack(x)
        du, dv = _dev(dU), _dev(dV)
        out = [_empty(nr * nc) for _ in range(5)]
        t.nrows, t.ncols, t.batch, t.gradmag = nr, nc, 1, 1 if gradmag else 0
        t.b1, t.b2, t.alpha = b1, b2, alpha
        t.dU, t.dV = du.data_ptr(), dv.data_ptr()
        for k in range(5):
            t.out[k] = out[k].data_ptr()
        t.batch_stride1, t.batch_stride2, t.batch_stride = nr * nc * t.channels1, nr * nc * max(t.channels2, 1), nr * nc
        self._run(self.L.pdegpu_dev_llin_terms(self.ctx.h, ctypes.byref(t)))
        return tuple(_host(o, (nr, nc)) for o in out)

    def llin_solve(self, d1, d2, U, V, dU, dV, b1, b2, alpha, gradmag, iters, omega, solver=2):
        """pdegpu_dev_llin_solve: OPdiffWeights(U+dU, V+dV) + robust weights / channel sums + Oflow_sor_llin4_2d in one
        call (fused into the line kernels' preparation where they run from packed lines). Returns the new (dU, dV)."""
        nr, nc = dU.shape
        t = lib.LlinTerms()
        keep = []

        def stack(x):
            x = np.asarray(x, dtype=np.float32).reshape(nr, nc, -1)
            keep.append(_dev(x))
            return keep[-1].data_ptr(), x.shape[2]
        for k, x in enumerate(d1):
            t.d1[k], t.channels1 = stack(x)
        t.channels2 = 0
        if d2 is not None:
            for k, x in enumerate(d2):
                t.d2[k], t.channels2 = stack(x)
        u, v, du, dv = _dev(U), _dev(V), _dev(dU), _dev(dV)
        work = _empty(11 * nr * nc)
        t.nrows, t.ncols, t.batch, t.gradmag = nr, nc, 1, 1 if gradmag else 0
        t.b1, t.b2, t.alpha = b1, b2, alpha
        t.batch_stride1, t.batch_stride2, t.batch_stride = nr * nc * t.channels1, nr * nc * max(t.channels2, 1), nr * nc
        self.L.pdegpu_dev_llin_solve.restype = ctypes.c_int
        self.L.pdegpu_dev_llin_solve.argtypes = [ctypes.c_void_p, ctypes.POINTER(lib.LlinTerms)] + [ctypes.c_void_p] * 5 + \
            [ctypes.c_int, ctypes.c_float, ctypes.c_int]
        self._run(self.L.pdegpu_dev_llin_solve(self.ctx.h, ctypes.byref(t), u.data_ptr(), v.data_ptr(), du.data_ptr(), dv.data_ptr(),
                                               work.data_ptr(), iters, omega, solver))
        return _host(du, (nr, nc)), _host(dv, (nr, nc))

    def elin_terms(self, der, coef, U, V, b1, b2, alpha, summed):
        nr, nc = U.shape
        t = lib.ElinTerms()
        keep = []
        ch = np.asarray(der[0]).reshape(nr, nc, -1).shape[2]
        for k, x in enumerate(der):
            keep.append(_dev(np.asarray(x).reshape(nr, nc, ch)))
            t.der[k] = keep[-1].data_ptr()
        for k, x in enumerate(coef):
            keep.append(_dev(np.asarray(x).reshape(nr, nc, ch)))
            t.coef[k] = keep[-1].data_ptr()
        u, v = _dev(U), _dev(V)
        oc = 1 if summed else ch
        gd = _empty(nr * nc * ch)
        out = [_empty(nr * nc * oc) for _ in range(5)]
        t.nrows, t.ncols, t.channels, t.summed = nr, nc, ch, 1 if summed else 0
        t.b1, t.b2, t.alpha = b1, b2, alpha
        t.U, t.V, t.gd = u.data_ptr(), v.data_ptr(), gd.data_ptr()
        for k in range(5):
            t.out[k] = out[k].data_ptr()
        self._run(self.L.pdegpu_dev_elin_terms(self.ctx.h, ctypes.byref(t)))
        shp = (nr, nc) if summed else (nr, nc, ch)
        return (_host(gd, (nr, nc, ch)),) + tuple(_host(o, shp) for o in out)

    def disp_sym_terms(self, d, dU, Udt, Udx, b1, b2, alpha, beta, srdiff):
        nr, nc = dU.shape
        t = lib.DispSymTerms()
        keep = []
        ch = np.asarray(d[0]).reshape(nr, nc, -1).shape[2]
        for k, x in enumerate(d):
            keep.append(_dev(np.asarray(x).reshape(nr, nc, ch)))
            t.d[k] = keep[-1].data_ptr()
        a, b, c = _dev(dU), _dev(Udt), _dev(Udx)
        cu, du = _empty(nr * nc), _empty(nr * nc)
        t.nrows, t.ncols, t.channels = nr, nc, ch
        t.b1, t.b2, t.alpha, t.alpha_d, t.beta, t.srdiff = b1, b2, alpha, alpha, beta, srdiff
        t.dU, t.Udt, t.Udx, t.CuG, t.DuG = a.data_ptr(), b.data_ptr(), c.data_ptr(), cu.data_ptr(), du.data_ptr()
        self._run(self.L.pdegpu_dev_disp_sym_terms(self.ctx.h, ctypes.byref(t)))
        return _host(cu, (nr, nc)), _host(du, (nr, nc))

    def fas_rhs(self, R, A, gd):
        shp = np.asarray(R).shape
        r, a, g = _dev(R), _dev(A), _dev(gd)
        f = _empty(r.numel())
        self._run(self.L.pdegpu_dev_fas_rhs(self.ctx.h, f.data_ptr(), r.data_ptr(), a.data_ptr(), g.data_ptr(), r.numel()))
        return _host(f, shp)

    def imfilter(self, A, h, conv=False, step=1, prescale=1.0):
        A = np.asarray(A, dtype=np.float32)
        nr, nc = A.shape[:2]
        planes = A.size // (nr * nc)
        h = np.atleast_2d(np.asarray(h, dtype=np.float64))
        if conv:
            h = h[::-1, ::-1]
        kr, kc = h.shape
        hh = np.ascontiguousarray(h.reshape(-1, order="F"))
        onr, onc = (nr + step - 1) // step, (nc + step - 1) // step
        a = _dev(A)
        o = _empty(onr * onc * planes)
        self._run(self.L.pdegpu_dev_imfilter(self.ctx.h, o.data_ptr(), a.data_ptr(), nr, nc, planes, nr * nc, onr * onc,
                                             hh.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), kr, kc, step, prescale))
        return _host(o, (onr, onc) + A.shape[2:])

    def imresize_bilinear(self, A, scale=None, output_size=None, antialias=True):
        A = np.asarray(A, dtype=np.float32)
        nr, nc = A.shape[:2]
        planes = A.size // (nr * nc)
        if output_size is None:
            onr, onc = int(math.ceil(nr * scale)), int(math.ceil(nc * scale))
            sr = sc = float(scale)
        else:
            onr, onc = int(output_size[0]), int(output_size[1])
            sr, sc = onr / nr, onc / nc
        a = _dev(A)
        o = _empty(onr * onc * planes)
        tmp = _empty(max(nr, onr) * max(nc, onc) * planes)
        self._run(self.L.pdegpu_dev_imresize_bilinear(self.ctx.h, o.data_ptr(), tmp.data_ptr(), a.data_ptr(), nr, nc, onr, onc,
                                                      sr, sc, 1 if antialias else 0, planes))
        return _host(o, (onr, onc) + A.shape[2:])

    def medfilt3(self, A):
        A = np.asarray(A, dtype=np.float32)
        nr, nc = A.shape
        a = _dev(A)
        o = _empty(nr * nc)
        self._run(self.L.pdegpu_dev_medfilt3(self.ctx.h, o.data_ptr(), a.data_ptr(), nr, nc, 1, nr * nc))
        return _host(o, (nr, nc))

    def warp_coords(self, U, V):
        nr, nc = U.shape
        u, v = _dev(U), _dev(V)
        x, y = _empty(nr * nc), _empty(nr * nc)
        self._run(self.L.pdegpu_dev_warp_coords(self.ctx.h, x.data_ptr(), y.data_ptr(), u.data_ptr(), v.data_ptr(), nr, nc, 1, nr * nc))
        return _host(x, (nr, nc)), _host(y, (nr, nc))

    def ad_diff_weights(self, D, Iin=None, quantile=0.5, scale=1.0):
        """(W, NW, N, NE, E, SE, S, SW, lambda[, TRACE, B])"""
        import torch
        D = np.asarray(D, dtype=np.float32)
        D3 = D.reshape(D.shape[0], D.shape[1], -1)
        nr, nc, fr = D3.shape
        d = _dev(D3)
        w = [_empty(nr * nc) for _ in range(8)]
        arr = (ctypes.c_void_p * 8)(*[t.data_ptr() for t in w])
        lam = torch.zeros(1, dtype=torch.float64, device="cuda:0")
        tr = bb = iin = None
        if Iin is not None:
            iin = _dev(np.asarray(Iin, dtype=np.float32).reshape(nr, nc, fr))
            tr, bb = _empty(nr * nc * fr), _empty(nr * nc * fr)
        self._run(self.L.pdegpu_dev_ad_diff_weights(self.ctx.h, ctypes.byref(arr), tr.data_ptr() if tr is not None else None,
                                                    bb.data_ptr() if bb is not None else None, d.data_ptr(),
                                                    iin.data_ptr() if iin is not None else None, nr, nc, fr, quantile, scale, lam.data_ptr()))
        out = tuple(_host(t, (nr, nc)) for t in w) + (float(lam.item()),)
        if tr is not None:
            out += (_host(tr, D.shape), _host(bb, D.shape))
        return out

    def rgb2grad(self, IN):
        IN = np.asarray(IN, dtype=np.float32).reshape(IN.shape[0], IN.shape[1], -1)
        out = np.zeros(IN.shape[:2] + (2 * IN.shape[2],), dtype=np.float32)
        for i in range(IN.shape[2]):
            out[:, :, 2 * i] = self.imfilter(IN[:, :, i], [[1.0, 0.0, -1.0]])
            out[:, :, 2 * i + 1] = self.imfilter(IN[:, :, i], [[1.0], [0.0], [-1.0]])
        return out
