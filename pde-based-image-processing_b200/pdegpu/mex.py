"""The reference's 13 hot-path MEX functions, served by libpdegpu through its C gateways.

    from pdegpu import mex
    U, V = mex.Oflow_sor_elin4_2d(U, V, M, Cu, Cv, Du, Dv, wW, wN, wE, wS, iter, omega, solver)
    U, V, RU, RV = mex.Oflow_sor_elin4_2d(..., nargout=4)

Arguments are exactly what the Matlab drivers pass (reference matlab/optical_flow/*.m etc.):
float32 arrays in Matlab (column-major) shape, scalars as float32 too. Anything else raises
MexError, like mexErrMsgTxt would in Matlab.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import numpy as np

from .mex_harness import MexLibrary, MexError

_HERE = os.path.dirname(os.path.abspath(__file__))
_MEXLIB = os.path.join(os.path.dirname(_HERE), "gateways", "pdegpu_mex.so")

NAMES = ("Oflow_sor_elin4_2d", "Oflow_sor_llin4_2d", "Oflow_sor_llin8_2d", "Oflow_lhs_elin4_2d", "Oflow_lhs_llin4_2d",
         "Disp_sor_llin4_2d", "Disp_sor_llin_sym4_2d", "PDEsolver4", "PDEsolver8",
         "BilinInterp_2d", "FstDerivatives5", "SndDerivatives5", "DdiffWeights")

_lib = None


def library() -> MexLibrary:
    """The gateway shared object; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_MEXLIB):
            raise ImportError(f"{_MEXLIB} not built: run `python pde-based-image-processing_b200/build.py` "
                              "(libpdegpu has no CPU fallback)")
        _lib = MexLibrary(_MEXLIB)
    return _lib


def call(fn: str, args: Sequence, nlhs: int) -> List[np.ndarray]:
    if fn not in NAMES:
        raise KeyError(fn)
    return library().call("mex_" + fn, args, nlhs)


class GpuBackend:
    """Same calling convention as oracle.oracle.{OracleBackend,RefBackend}."""
    name = "libpdegpu"

    def call(self, fn: str, args: Sequence, nlhs: int) -> List[np.ndarray]:
        return call(fn, args, nlhs)

    def bilin(self, Iin, X, Y, oob: float) -> np.ndarray:
        """BilinInterp_2d as the drivers call it; the gateway passes NaN for out-of-image pixels"""
        if not np.isnan(oob):
            raise ValueError("the BilinInterp_2d gateway marks out-of-image pixels with NaN")
        return call("BilinInterp_2d", [np.asarray(Iin, dtype=np.float32), np.asarray(X, dtype=np.float32), np.asarray(Y, dtype=np.float32)], 1)[0]


def _make(fn: str, default_nargout: int):
    def f(*args, nargout: int = default_nargout):
        out = call(fn, args, nargout)
        return out[0] if nargout == 1 else tuple(out)
    f.__name__ = fn
    f.__doc__ = f"{fn}(...) -- see the header of gateways/{fn}.c for the signature."
    return f


Oflow_sor_elin4_2d = _make("Oflow_sor_elin4_2d", 2)
Oflow_sor_llin4_2d = _make("Oflow_sor_llin4_2d", 2)
Oflow_sor_llin8_2d = _make("Oflow_sor_llin8_2d", 2)
Oflow_lhs_elin4_2d = _make("Oflow_lhs_elin4_2d", 2)
Oflow_lhs_llin4_2d = _make("Oflow_lhs_llin4_2d", 2)
Disp_sor_llin4_2d = _make("Disp_sor_llin4_2d", 1)
Disp_sor_llin_sym4_2d = _make("Disp_sor_llin_sym4_2d", 2)
PDEsolver4 = _make("PDEsolver4", 1)
PDEsolver8 = _make("PDEsolver8", 1)
BilinInterp_2d = _make("BilinInterp_2d", 1)
FstDerivatives5 = _make("FstDerivatives5", 3)
SndDerivatives5 = _make("SndDerivatives5", 5)
DdiffWeights = _make("DdiffWeights", 4)
