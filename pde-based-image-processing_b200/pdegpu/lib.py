"""ctypes binding of the libpdegpu C ABI (include/pdegpu.h).

Device memory is addressed by integer pointers (e.g. ``torch.Tensor.data_ptr()``); torch is only
plumbing for allocation and streams, never part of the compute path.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_longlong, c_size_t, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.environ.get("PDEGPU_LIB") or os.path.join(os.path.dirname(_HERE), "libpdegpu.so")

FLOW_ELIN4, FLOW_LLIN4, FLOW_LLIN8, DISP_LLIN4, PDE4, PDE8 = range(6)
ORDER_FAST, ORDER_REFERENCE, ORDER_AUTO = 0, 1, 2
W_W, W_N, W_E, W_S, W_NW, W_NE, W_SE, W_SW = range(8)

OK = 0
ERR_NODEVICE = -5


class PdegpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libpdegpu error {code}: {msg}")
        self.code = code


class System(ctypes.Structure):
    """Mirror of `pdegpu_system`."""
    _fields_ = [("family", c_int), ("nrows", c_int), ("ncols", c_int), ("batch", c_int),
                ("batch_stride", c_longlong),
                ("x", c_void_p * 2), ("x0", c_void_p * 2), ("m", c_void_p),
                ("c", c_void_p * 2), ("d", c_void_p * 2), ("w", c_void_p * 8)]


class LlinTerms(ctypes.Structure):
    """Mirror of `pdegpu_llin_terms`."""
    _fields_ = [("nrows", c_int), ("ncols", c_int), ("batch", c_int),
                ("channels1", c_int), ("channels2", c_int), ("gradmag", c_int),
                ("b1", c_float), ("b2", c_float), ("alpha", c_float),
                ("d1", c_void_p * 3), ("d2", c_void_p * 5), ("dU", c_void_p), ("dV", c_void_p),
                ("out", c_void_p * 5),
                ("batch_stride1", c_longlong), ("batch_stride2", c_longlong), ("batch_stride", c_longlong)]


class ElinTerms(ctypes.Structure):
    """Mirror of `pdegpu_elin_terms`."""
    _fields_ = [("nrows", c_int), ("ncols", c_int), ("channels", c_int), ("summed", c_int),
                ("b1", c_float), ("b2", c_float), ("alpha", c_float),
                ("der", c_void_p * 8), ("coef", c_void_p * 5), ("U", c_void_p), ("V", c_void_p),
                ("gd", c_void_p), ("out", c_void_p * 5)]


class DispSymTerms(ctypes.Structure):
    """Mirror of `pdegpu_disp_sym_terms`."""
    _fields_ = [("nrows", c_int), ("ncols", c_int), ("channels", c_int),
                ("b1", c_float), ("b2", c_float), ("alpha", c_float),
                ("alpha_d", ctypes.c_double), ("beta", ctypes.c_double), ("srdiff", ctypes.c_double),
                ("d", c_void_p * 6), ("dU", c_void_p), ("Udt", c_void_p), ("Udx", c_void_p),
                ("CuG", c_void_p), ("DuG", c_void_p)]


class TvParams(ctypes.Structure):
    """Mirror of `pdegpu_tvdenoise8_params`."""
    _fields_ = [("alpha", ctypes.c_double), ("omega", ctypes.c_double), ("scl_factor", ctypes.c_double),
                ("outer_iter", c_int), ("inner_iter", c_int), ("solver", c_int)]


class FlowLlinParams(ctypes.Structure):
    """Mirror of `pdegpu_flow_llin_params`."""
    _fields_ = [("alpha", ctypes.c_double), ("omega", ctypes.c_double), ("b1", ctypes.c_double), ("b2", ctypes.c_double),
                ("scl_factor", ctypes.c_double),
                ("firstLoop", c_int), ("secondLoop", c_int), ("iter", c_int), ("solver", c_int),
                ("fst_grad", c_int), ("snd_term", c_int), ("max_scales", c_int), ("oob_value", c_float)]


class FlowFmgParams(ctypes.Structure):
    """Mirror of `pdegpu_flow_fmg_params`."""
    _fields_ = [("alpha", ctypes.c_double), ("omega", ctypes.c_double), ("b1", ctypes.c_double), ("b2", ctypes.c_double),
                ("scl_factor", ctypes.c_double),
                ("firstLoop", c_int), ("iter", c_int), ("solver", c_int), ("cycle_index", c_int), ("max_scales", c_int)]


class FlowHsParams(ctypes.Structure):
    """Mirror of `pdegpu_flow_hs_params`."""
    _fields_ = [("alpha", ctypes.c_double), ("omega", ctypes.c_double), ("b1", ctypes.c_double), ("b2", ctypes.c_double),
                ("scl_factor", ctypes.c_double), ("iter", c_int), ("solver", c_int), ("max_scales", c_int)]


class DispSymParams(ctypes.Structure):
    """Mirror of `pdegpu_disp_sym_params`."""
    _fields_ = [("alpha", ctypes.c_double), ("beta", ctypes.c_double), ("omega", ctypes.c_double), ("b1", ctypes.c_double),
                ("b2", ctypes.c_double), ("scl_factor", ctypes.c_double),
                ("firstLoop", c_int), ("secondLoop", c_int), ("iter", c_int), ("solver", c_int),
                ("max_scales", c_int), ("uint8_input", c_int), ("oob_value", c_float)]


_dll = None


def dll() -> ctypes.CDLL:
    global _dll
    if _dll is None:
        if not os.path.exists(LIBPATH):
            raise ImportError(f"{LIBPATH} not built: run `python pde-based-image-processing_b200/build.py` "
                              "(libpdegpu has no CPU fallback)")
        L = ctypes.CDLL(LIBPATH)
        L.pdegpu_device_count.restype = c_int
        L.pdegpu_init.restype = c_int
        L.pdegpu_init.argtypes = [c_int, POINTER(c_void_p)]
        L.pdegpu_free.restype = None
        L.pdegpu_free.argtypes = [c_void_p]
        L.pdegpu_last_error.restype = c_char_p
        L.pdegpu_last_error.argtypes = [c_void_p]
        L.pdegpu_version.restype = c_char_p
        L.pdegpu_sync.restype = c_int
        L.pdegpu_sync.argtypes = [c_void_p]
        L.pdegpu_stream.restype = c_void_p
        L.pdegpu_stream.argtypes = [c_void_p]
        L.pdegpu_launch_count.restype = c_ulonglong
        L.pdegpu_launch_count.argtypes = [c_void_p]
        L.pdegpu_set_kernel_path.restype = c_int
        L.pdegpu_set_kernel_path.argtypes = [c_void_p, c_int]
        L.pdegpu_set_sweep_order.restype = c_int
        L.pdegpu_set_sweep_order.argtypes = [c_void_p, c_int]
        L.pdegpu_get_sweep_order.restype = c_int
        L.pdegpu_get_sweep_order.argtypes = [c_void_p]
        L.pdegpu_dev_relax.restype = c_int
        L.pdegpu_dev_relax.argtypes = [c_void_p, POINTER(System), c_int, c_float, c_int]
        L.pdegpu_band_create.restype = c_int
        L.pdegpu_band_create.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, POINTER(c_void_p)]
        L.pdegpu_band_export.restype = c_int
        L.pdegpu_band_export.argtypes = [c_void_p, c_void_p]
        L.pdegpu_band_connect.restype = c_int
        L.pdegpu_band_connect.argtypes = [c_void_p, c_void_p, c_void_p]
        L.pdegpu_band_connect_local.restype = c_int
        L.pdegpu_band_connect_local.argtypes = [c_void_p, c_void_p, c_void_p]
        L.pdegpu_band_exchange.restype = c_int
        L.pdegpu_band_exchange.argtypes = [c_void_p, POINTER(c_void_p), c_int, c_int]
        L.pdegpu_band_bytes_sent.restype = c_ulonglong
        L.pdegpu_band_bytes_sent.argtypes = [c_void_p]
        L.pdegpu_band_free.restype = None
        L.pdegpu_band_free.argtypes = [c_void_p]
        L.pdegpu_dev_residual.restype = c_int
        L.pdegpu_dev_residual.argtypes = [c_void_p, POINTER(System), c_int, c_void_p, c_void_p]
        L.pdegpu_dev_lhs.restype = c_int
        L.pdegpu_dev_lhs.argtypes = [c_void_p, POINTER(System), c_int, c_void_p, c_void_p]
        L.pdegpu_dev_bilin_interp_2d.restype = c_int
        L.pdegpu_dev_bilin_interp_2d.argtypes = [c_void_p] + [c_void_p] * 4 + [c_int] * 3 + [c_float]
        L.pdegpu_dev_fst_derivatives5.restype = c_int
        L.pdegpu_dev_fst_derivatives5.argtypes = [c_void_p] + [c_void_p] * 5 + [c_int] * 3
        L.pdegpu_dev_snd_derivatives5.restype = c_int
        L.pdegpu_dev_snd_derivatives5.argtypes = [c_void_p] + [c_void_p] * 7 + [c_int] * 3
        L.pdegpu_dev_ddiff_weights.restype = c_int
        L.pdegpu_dev_ddiff_weights.argtypes = [c_void_p] + [c_void_p] * 5 + [c_int] * 3 + [c_float]
        L.pdegpu_profile_enable.restype = c_int
        L.pdegpu_profile_enable.argtypes = [c_void_p, c_int]
        L.pdegpu_profile_report.restype = c_int
        L.pdegpu_profile_report.argtypes = [c_void_p, c_char_p, c_size_t]
        c_double = ctypes.c_double
        L.pdegpu_dev_op_diff_weights.restype = c_int
        L.pdegpu_dev_op_diff_weights.argtypes = [c_void_p] + [c_void_p] * 6 + [c_int] * 3 + [c_longlong]
        L.pdegpu_dev_llin_terms.restype = c_int
        L.pdegpu_dev_llin_terms.argtypes = [c_void_p, POINTER(LlinTerms)]
        L.pdegpu_dev_elin_terms.restype = c_int
        L.pdegpu_dev_elin_terms.argtypes = [c_void_p, POINTER(ElinTerms)]
        L.pdegpu_dev_disp_sym_terms.restype = c_int
        L.pdegpu_dev_disp_sym_terms.argtypes = [c_void_p, POINTER(DispSymTerms)]
        L.pdegpu_dev_fas_rhs.restype = c_int
        L.pdegpu_dev_fas_rhs.argtypes = [c_void_p] + [c_void_p] * 4 + [c_longlong]
        L.pdegpu_dev_imfilter.restype = c_int
        L.pdegpu_dev_imfilter.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_longlong, c_longlong,
                                          POINTER(c_double), c_int, c_int, c_int, c_float]
        L.pdegpu_dev_imresize_bilinear.restype = c_int
        L.pdegpu_dev_imresize_bilinear.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                                   c_double, c_double, c_int, c_int]
        L.pdegpu_dev_medfilt3.restype = c_int
        L.pdegpu_dev_medfilt3.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_longlong]
        L.pdegpu_dev_axpby.restype = c_int
        L.pdegpu_dev_axpby.argtypes = [c_void_p, c_void_p, c_float, c_void_p, c_float, c_void_p, c_longlong]
        L.pdegpu_dev_warp_coords.restype = c_int
        L.pdegpu_dev_warp_coords.argtypes = [c_void_p] + [c_void_p] * 4 + [c_int] * 3 + [c_longlong]
        L.pdegpu_dev_ad_diff_weights.restype = c_int
        L.pdegpu_dev_ad_diff_weights.argtypes = [c_void_p, POINTER(c_void_p * 8), c_void_p, c_void_p, c_void_p, c_void_p,
                                                 c_int, c_int, c_int, c_double, c_double, c_void_p]
        L.pdegpu_tvdenoise8_default_params.restype = None
        L.pdegpu_tvdenoise8_default_params.argtypes = [POINTER(TvParams)]
        for fn in (L.pdegpu_dev_tvdenoise8, L.pdegpu_tvdenoise8):
            fn.restype = c_int
            fn.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, POINTER(TvParams)]
        L.pdegpu_flow_llin_default_params.restype = None
        L.pdegpu_flow_llin_default_params.argtypes = [POINTER(FlowLlinParams)]
        for fn in (L.pdegpu_dev_flow_llin_2d, L.pdegpu_flow_llin_2d):
            fn.restype = c_int
            fn.argtypes = [c_void_p] + [c_void_p] * 4 + [c_int] * 4 + [POINTER(FlowLlinParams)]
        L.pdegpu_flow_fmg_default_params.restype = None
        L.pdegpu_flow_fmg_default_params.argtypes = [POINTER(FlowFmgParams)]
        for fn in (L.pdegpu_dev_flow_fmg_2d, L.pdegpu_flow_fmg_2d):
            fn.restype = c_int
            fn.argtypes = [c_void_p] + [c_void_p] * 4 + [c_int] * 4 + [POINTER(FlowFmgParams)]
        L.pdegpu_flow_hs_default_params.restype = None
        L.pdegpu_flow_hs_default_params.argtypes = [POINTER(FlowHsParams)]
        for fn in (L.pdegpu_dev_flow_hs_2d, L.pdegpu_flow_hs_2d):
            fn.restype = c_int
            fn.argtypes = [c_void_p] + [c_void_p] * 4 + [c_int] * 4 + [POINTER(FlowHsParams)]
        L.pdegpu_disp_sym_default_params.restype = None
        L.pdegpu_disp_sym_default_params.argtypes = [POINTER(DispSymParams)]
        for fn in (L.pdegpu_dev_disp_sym_2d, L.pdegpu_disp_sym_2d):
            fn.restype = c_int
            fn.argtypes = [c_void_p] + [c_void_p] * 3 + [c_int] * 4 + [POINTER(DispSymParams)]
        L.pdegpu_upload.restype = c_int
        L.pdegpu_upload.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t]
        L.pdegpu_download.restype = c_int
        L.pdegpu_download.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t]
        _dll = L
    return _dll


class Context:
    """One libpdegpu context (= one GPU + one stream). Not thread-safe, like the C object."""

    def __init__(self, device: int = 0):
        L = dll()
        h = c_void_p()
        rc = L.pdegpu_init(device, ctypes.byref(h))
        if rc != OK:
            raise PdegpuError(rc, L.pdegpu_last_error(None).decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            dll().pdegpu_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != OK:
            raise PdegpuError(rc, dll().pdegpu_last_error(self.h).decode())

    @property
    def stream(self) -> int:
        return int(dll().pdegpu_stream(self.h) or 0)

    @property
    def launches(self) -> int:
        return int(dll().pdegpu_launch_count(self.h))

    def sync(self):
        self._chk(dll().pdegpu_sync(self.h))

    def set_kernel_path(self, path: int):
        self._chk(dll().pdegpu_set_kernel_path(self.h, path))

    def set_sweep_order(self, order: int):
        """ORDER_FAST (zebra) or ORDER_REFERENCE (the reference's lexicographic line order), solver 2 only"""
        self._chk(dll().pdegpu_set_sweep_order(self.h, order))

    def profile(self, on: bool):
        self._chk(dll().pdegpu_profile_enable(self.h, 1 if on else 0))

    def profile_report(self):
        import json
        buf = ctypes.create_string_buffer(1 << 16)
        self._chk(dll().pdegpu_profile_report(self.h, buf, len(buf)))
        return json.loads(buf.value.decode())

    def relax(self, sys: System, iters: int, omega: float, solver: int):
        self._chk(dll().pdegpu_dev_relax(self.h, ctypes.byref(sys), iters, c_float(omega), solver))

    def residual(self, sys: System, nframes: int, RU: int, RV: int):
        self._chk(dll().pdegpu_dev_residual(self.h, ctypes.byref(sys), nframes, RU, RV))

    def lhs(self, sys: System, nframes: int, AU: int, AV: int):
        self._chk(dll().pdegpu_dev_lhs(self.h, ctypes.byref(sys), nframes, AU, AV))

    def oflow_sor_batch(self, fields: dict, late: bool, iters: int, omega: float, solver: int):
        """pdegpu_oflow_sor_{llin4,elin4}_2d_batch: `fields` maps the gateway's argument names (U, V, (dU, dV,) M, Cu, Cv,
        Du, Dv, wW, wN, wE, wS) to numpy arrays [batch, rows, cols] (single). Returns the two relaxed unknowns with the
        same shape. Host-pointer entry point: chunked uploads, sweeps and downloads overlap inside the call."""
        import numpy as np
        keys = (("U", "V", "dU", "dV") if late else ("U", "V")) + ("M", "Cu", "Cv", "Du", "Dv", "wW", "wN", "wE", "wS")
        B, nr, nc = fields["U"].shape
        # system b of every array at offset b*nr*nc, each system column-major like the MEX arrays
        flat = [np.ascontiguousarray(np.asarray(fields[k], np.float32).transpose(0, 2, 1)).reshape(-1) for k in keys]
        assert all(a.size == B * nr * nc for a in flat)
        o0, o1 = np.empty(B * nr * nc, np.float32), np.empty(B * nr * nc, np.float32)
        fn = dll().pdegpu_oflow_sor_llin4_2d_batch if late else dll().pdegpu_oflow_sor_elin4_2d_batch
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_void_p] * (3 + len(keys)) + [ctypes.c_int] * 3 + [c_float] * 2 + [ctypes.c_int]
        self._chk(fn(self.h, o0.ctypes.data, o1.ctypes.data, *[a.ctypes.data for a in flat], nr, nc, B,
                     c_float(iters), c_float(omega), solver))
        back = lambda a: a.reshape(B, nc, nr).transpose(0, 2, 1)
        return back(o0), back(o1)

    def tvdenoise8(self, I, **overrides):
        """Iout = TVdenoise8(I): I numpy [rows, cols(, frames)] single. Host-pointer entry point."""
        import numpy as np
        I = np.asarray(I, dtype=np.float32)
        I3 = I.reshape(I.shape[0], I.shape[1], -1)
        nr, nc, fr = I3.shape
        p = TvParams()
        dll().pdegpu_tvdenoise8_default_params(ctypes.byref(p))
        for k, v in overrides.items():
            setattr(p, k, v)
        a = np.ascontiguousarray(I3.reshape(-1, order="F"))
        out = np.empty_like(a)
        self._chk(dll().pdegpu_tvdenoise8(self.h, out.ctypes.data, a.ctypes.data, nr, nc, fr, ctypes.byref(p)))
        return out.reshape((nr, nc, fr), order="F").reshape(I.shape)

    def flow_llin(self, I0, I1, params: "FlowLlinParams | None" = None, **overrides):
        """[U V] = FlowEminND_llin_2D_v10 for one pair or a batch: I0, I1 numpy [rows, cols, channels] or
        [batch, rows, cols, channels], values 0..255. Host-pointer entry point (H2D, pipeline, D2H)."""
        import numpy as np
        I0 = np.asarray(I0, dtype=np.float32)
        I1 = np.asarray(I1, dtype=np.float32)
        single = I0.ndim == 3
        if single:
            I0, I1 = I0[None], I1[None]
        B, nr, nc, C = I0.shape
        p = params or FlowLlinParams()
        if params is None:
            dll().pdegpu_flow_llin_default_params(ctypes.byref(p))
        for k, v in overrides.items():
            setattr(p, k, v)
        # Matlab layout per pair: column-major rows x cols x channels
        a0 = np.ascontiguousarray(np.stack([I0[b].reshape(-1, order="F") for b in range(B)]))
        a1 = np.ascontiguousarray(np.stack([I1[b].reshape(-1, order="F") for b in range(B)]))
        U = np.empty((B, nr * nc), dtype=np.float32)
        V = np.empty((B, nr * nc), dtype=np.float32)
        self._chk(dll().pdegpu_flow_llin_2d(self.h, U.ctypes.data, V.ctypes.data, a0.ctypes.data, a1.ctypes.data,
                                            nr, nc, C, B, ctypes.byref(p)))
        U = np.stack([U[b].reshape((nr, nc), order="F") for b in range(B)])
        V = np.stack([V[b].reshape((nr, nc), order="F") for b in range(B)])
        return (U[0], V[0]) if single else (U, V)

    def disp_sym(self, Il, Ir, **overrides):
        """(U0, U1) = DispEminND_llin_sym_2D(Il, Ir) for one stereo pair: Il, Ir numpy [rows, cols(, channels)], 0..255."""
        import numpy as np
        Il = np.asarray(Il, dtype=np.float32); Ir = np.asarray(Ir, dtype=np.float32)
        Il = Il.reshape(Il.shape[0], Il.shape[1], -1); Ir = Ir.reshape(Ir.shape[0], Ir.shape[1], -1)
        nr, nc, C = Il.shape
        p = DispSymParams()
        dll().pdegpu_disp_sym_default_params(ctypes.byref(p))
        for k, v in overrides.items():
            setattr(p, k, v)
        a0 = np.ascontiguousarray(Il.reshape(-1, order="F")); a1 = np.ascontiguousarray(Ir.reshape(-1, order="F"))
        U = np.empty(2 * nr * nc, dtype=np.float32)
        self._chk(dll().pdegpu_disp_sym_2d(self.h, U.ctypes.data, a0.ctypes.data, a1.ctypes.data, nr, nc, C, 1, ctypes.byref(p)))
        U = U.reshape((nr, nc, 2), order="F")
        return U[:, :, 0].copy(), U[:, :, 1].copy()

    def flow_hs(self, I0, I1, params: "FlowHsParams | None" = None, **overrides):
        """[U V] = FlowEminHS_elin_2D_v10 (Horn-Schunck) for one pair or a batch, same conventions as flow_fmg."""
        return self.flow_fmg(I0, I1, params, _hs=True, **overrides)

    def flow_fmg(self, I0, I1, params: "FlowFmgParams | None" = None, _hs=False, **overrides):
        """[U V] = FlowEminNDFASFMG_elin_2D_v10 for one pair or a batch: I0, I1 numpy [rows, cols, channels] or
        [batch, rows, cols, channels], values 0..255. Host-pointer entry point (H2D, pipeline, D2H)."""
        import numpy as np
        cls, dflt, run = ((FlowHsParams, dll().pdegpu_flow_hs_default_params, dll().pdegpu_flow_hs_2d) if _hs else
                          (FlowFmgParams, dll().pdegpu_flow_fmg_default_params, dll().pdegpu_flow_fmg_2d))
        I0 = np.asarray(I0, dtype=np.float32)
        I1 = np.asarray(I1, dtype=np.float32)
        single = I0.ndim == 3
        if single:
            I0, I1 = I0[None], I1[None]
        B, nr, nc, C = I0.shape
        p = params or cls()
        if params is None:
            dflt(ctypes.byref(p))
        for k, v in overrides.items():
            setattr(p, k, v)
        a0 = np.ascontiguousarray(np.stack([I0[b].reshape(-1, order="F") for b in range(B)]))
        a1 = np.ascontiguousarray(np.stack([I1[b].reshape(-1, order="F") for b in range(B)]))
        U = np.empty((B, nr * nc), dtype=np.float32)
        V = np.empty((B, nr * nc), dtype=np.float32)
        self._chk(run(self.h, U.ctypes.data, V.ctypes.data, a0.ctypes.data, a1.ctypes.data, nr, nc, C, B, ctypes.byref(p)))
        U = np.stack([U[b].reshape((nr, nc), order="F") for b in range(B)])
        V = np.stack([V[b].reshape((nr, nc), order="F") for b in range(B)])
        return (U[0], V[0]) if single else (U, V)

    def upload(self, dst_dev: int, src_host: int, nbytes: int):
        self._chk(dll().pdegpu_upload(self.h, dst_dev, src_host, nbytes))

    def download(self, dst_host: int, src_dev: int, nbytes: int):
        self._chk(dll().pdegpu_download(self.h, dst_host, src_dev, nbytes))


def make_system(family: int, nrows: int, ncols: int, batch: int = 1, batch_stride: int | None = None,
                x=(), x0=(), m=None, c=(), d=(), w=()) -> System:
    """Build a `pdegpu_system` from integer device pointers."""
    s = System()
    s.family, s.nrows, s.ncols, s.batch = family, nrows, ncols, batch
    s.batch_stride = nrows * ncols if batch_stride is None else batch_stride
    for k, p in enumerate(x):
        s.x[k] = p
    for k, p in enumerate(x0):
        s.x0[k] = p
    s.m = m
    for k, p in enumerate(c):
        s.c[k] = p
    for k, p in enumerate(d):
        s.d[k] = p
    for k, p in enumerate(w):
        s.w[k] = p
    return s


class BandExchange:
    """pdegpu_band_* (include/pdegpu.h): the halo exchange of one band over peer memory. `export()` gives the 64-byte
    CUDA IPC handle of this band's mailbox; `connect(left, right)` takes the neighbours' handles (bytes; one process per
    GPU) and `connect_local(left, right)` other BandExchange objects of the same process."""

    def __init__(self, ctx: "Context", nrows: int, halo_cols: int, nunknowns: int, has_left: bool, has_right: bool):
        self.ctx = ctx
        self.h = c_void_p()
        ctx._chk(dll().pdegpu_band_create(ctx.h, nrows, halo_cols, nunknowns, int(has_left), int(has_right), ctypes.byref(self.h)))

    def export(self) -> bytes:
        buf = ctypes.create_string_buffer(64)
        self.ctx._chk(dll().pdegpu_band_export(self.h, buf))
        return buf.raw

    def connect(self, left: "bytes | None", right: "bytes | None"):
        l = ctypes.create_string_buffer(left, 64) if left else None
        r = ctypes.create_string_buffer(right, 64) if right else None
        self.ctx._chk(dll().pdegpu_band_connect(self.h, l, r))

    def connect_local(self, left: "BandExchange | None", right: "BandExchange | None"):
        self.ctx._chk(dll().pdegpu_band_connect_local(self.h, left.h if left else None, right.h if right else None))

    def exchange(self, unknown_ptrs, own0: int, own1: int):
        arr = (c_void_p * len(unknown_ptrs))(*unknown_ptrs)
        self.ctx._chk(dll().pdegpu_band_exchange(self.h, arr, own0, own1))

    @property
    def bytes_sent(self) -> int:
        return int(dll().pdegpu_band_bytes_sent(self.h))

    def close(self):
        if self.h:
            dll().pdegpu_band_free(self.h)
            self.h = c_void_p()
