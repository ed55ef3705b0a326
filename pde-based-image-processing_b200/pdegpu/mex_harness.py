"""Drive MEX gateways that were built as plain shared objects against gateways/mex_shim.

A gateway keeps its Matlab calling convention
``mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])``; this module
plays the part of the Matlab interpreter: it wraps numpy arrays (column-major, like Matlab)
into the shim's ``mxArray``, calls the gateway through ``shim_call`` (which turns
``mexErrMsgTxt`` into an error return) and copies the outputs back into numpy arrays.

Nothing here computes anything: it is marshalling only.
"""
from __future__ import annotations

import ctypes
import os
from typing import Iterable, List, Sequence

import numpy as np

MX_DOUBLE = 6
MX_SINGLE = 7
_MAXDIMS = 8


class MexError(RuntimeError):
    """Raised when a gateway calls mexErrMsgTxt (Matlab would throw an error)."""


def _as_matlab(a) -> np.ndarray:
    """numpy view of `a` with Matlab semantics: >=2-D, column-major. dtype is preserved
    (the gateways insist on single and we want their type checks to see what the caller passed);
    python scalars become double 1x1, like a Matlab literal."""
    if isinstance(a, (int, float)):
        a = np.array([[a]], dtype=np.float64)
    a = np.asarray(a)
    if a.ndim == 0:
        a = a.reshape(1, 1)
    elif a.ndim == 1:
        a = a.reshape(-1, 1)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    return np.asfortranarray(a)


class MexLibrary:
    """One shared object holding one or more renamed mexFunction entry points + the shim."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path = path
        self.lib = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
        L = self.lib
        L.shim_wrap.restype = ctypes.c_void_p
        L.shim_wrap.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_void_p]
        L.shim_ndims.restype = ctypes.c_int
        L.shim_ndims.argtypes = [ctypes.c_void_p]
        L.shim_dim.restype = ctypes.c_ulonglong
        L.shim_dim.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.shim_data.restype = ctypes.c_void_p
        L.shim_data.argtypes = [ctypes.c_void_p]
        L.shim_classid.restype = ctypes.c_int
        L.shim_classid.argtypes = [ctypes.c_void_p]
        L.shim_sizeof_mwsize.restype = ctypes.c_int
        L.shim_call.restype = ctypes.c_int
        L.shim_call.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int,
                                ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p, ctypes.c_int]
        L.mxDestroyArray.restype = None
        L.mxDestroyArray.argtypes = [ctypes.c_void_p]

    @property
    def mwsize_bytes(self) -> int:
        return int(self.lib.shim_sizeof_mwsize())

    def has(self, entry: str) -> bool:
        return hasattr(self.lib, entry)

    def call(self, entry: str, args: Sequence, nlhs: int) -> List[np.ndarray]:
        """Call gateway `entry` (e.g. 'mex_Oflow_sor_elin4_2d') like Matlab would:
        ``[out1..out_nlhs] = entry(args...)``."""
        fn = getattr(self.lib, entry)
        fn_ptr = ctypes.cast(fn, ctypes.c_void_p)
        keep = []          # keep numpy buffers alive during the call
        prhs = (ctypes.c_void_p * max(len(args), 1))()
        wrapped = []
        for k, a in enumerate(args):
            m = _as_matlab(a)
            keep.append(m)
            dims = (ctypes.c_ulonglong * _MAXDIMS)(*([int(d) for d in m.shape] + [1] * (_MAXDIMS - m.ndim)))
            cls = MX_SINGLE if m.dtype == np.float32 else MX_DOUBLE
            h = self.lib.shim_wrap(cls, m.ndim, dims, m.ctypes.data_as(ctypes.c_void_p))
            wrapped.append(h)
            prhs[k] = h
        # Matlab always hands a gateway room for at least one output
        plhs = (ctypes.c_void_p * max(nlhs, 1))()
        err = ctypes.create_string_buffer(512)
        rc = self.lib.shim_call(fn_ptr, nlhs, plhs, len(args), prhs, err, 512)
        try:
            if rc != 0:
                raise MexError(err.value.decode("utf-8", "replace"))
            outs = []
            for k in range(nlhs):
                h = plhs[k]
                if not h:
                    outs.append(None)
                    continue
                nd = self.lib.shim_ndims(h)
                shape = tuple(int(self.lib.shim_dim(h, d)) for d in range(nd))
                cls = self.lib.shim_classid(h)
                dt = np.float32 if cls == MX_SINGLE else np.float64
                n = int(np.prod(shape)) if shape else 1
                ptr = self.lib.shim_data(h)
                if n == 0 or not ptr:
                    outs.append(np.zeros(shape, dtype=dt, order="F"))
                else:
                    buf = (ctypes.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr)
                    outs.append(np.frombuffer(buf, dtype=dt).reshape(shape, order="F").copy(order="F"))
            return outs
        finally:
            if rc == 0:
                for k in range(max(nlhs, 1)):
                    if plhs[k]:
                        self.lib.mxDestroyArray(plhs[k])
            for h in wrapped:
                self.lib.mxDestroyArray(h)     # frees only the header (data is borrowed)


def f32(x) -> np.ndarray:
    """`single(x)` for scalars/arrays (column-major)."""
    return np.asfortranarray(np.asarray(x, dtype=np.float32))
