"""CPU checkers for libpdegpu.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. Two interchangeable back-ends expose the 13 hot-path MEX functions through
one calling convention, ``backend.call("Oflow_sor_elin4_2d", [args...], nlhs) -> [arrays]``
(the same convention `pdegpu.mex` uses for the GPU gateways):

* ``OracleBackend``  -- our plain-C restatement (oracle/pde_oracle.c, built by `make -C oracle oracle`),
  driven by a Python restatement of what each reference gateway does around the library call
  (output shapes, iter<=0 behaviour, residuals of the inputs, ...). Always available.
* ``RefBackend``     -- the UNMODIFIED reference sources compiled against the mex.h shim
  (oracle/_ref/*.so, built by `make -C oracle ref` where /root/reference exists). The prebuilt
  shared objects travel to the GPU box; nothing here reads /root/reference at run time.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
from typing import List, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "pde-based-image-processing_b200"))
from pdegpu.mex_harness import MexLibrary, MexError  # noqa: E402  (marshalling only, no compute)

FP = ctypes.POINTER(ctypes.c_float)


def build(ref: bool = True) -> None:
    """Compile the C restatement and, where /root/reference exists, the reference itself."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref and os.path.isdir("/root/reference/mex/source"):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


def have_ref() -> bool:
    return all(os.path.exists(os.path.join(HERE, "_ref", f + ".so"))
               for f in ("ref_oflow", "ref_disp", "ref_pde", "ref_interp", "ref_deriv", "ref_dweights"))


# ---------------------------------------------------------------------------------------------
class RefBackend:
    """The reference's own gateways (compiled unmodified) behind the common calling convention."""
    name = "reference"
    _family = {
        "Oflow_sor_elin4_2d": "ref_oflow", "Oflow_sor_llin4_2d": "ref_oflow", "Oflow_sor_llin8_2d": "ref_oflow",
        "Oflow_lhs_elin4_2d": "ref_oflow", "Oflow_lhs_llin4_2d": "ref_oflow",
        "Disp_sor_llin4_2d": "ref_disp", "Disp_sor_llin_sym4_2d": "ref_disp",
        "PDEsolver4": "ref_pde", "PDEsolver8": "ref_pde",
        "BilinInterp_2d": "ref_interp", "FstDerivatives5": "ref_deriv", "SndDerivatives5": "ref_deriv",
        "DdiffWeights": "ref_dweights",
    }

    def __init__(self):
        if not have_ref():
            raise FileNotFoundError("oracle/_ref/*.so not built (run `make -C oracle ref` where /root/reference exists)")
        self._libs = {}

    def _lib(self, fam: str) -> MexLibrary:
        if fam not in self._libs:
            self._libs[fam] = MexLibrary(os.path.join(HERE, "_ref", fam + ".so"))
        return self._libs[fam]

    def call(self, fn: str, args: Sequence, nlhs: int) -> List[np.ndarray]:
        return self._lib(self._family[fn]).call("mex_" + fn, args, nlhs)

    def bilin(self, Iin, X, Y, oob: float) -> np.ndarray:
        """bilinInterp2 through an explicit 5-argument prototype (oracle/ref_wrappers.c)."""
        lib = self._lib("ref_interp").lib
        Iin = np.asfortranarray(Iin, dtype=np.float32)
        X = np.asfortranarray(X, dtype=np.float32)
        Y = np.asfortranarray(Y, dtype=np.float32)
        out = np.zeros(Iin.shape, dtype=np.float32, order="F")
        nf = Iin.shape[2] if Iin.ndim > 2 else 1
        lib.ref_bilinInterp2.restype = None
        lib.ref_bilinInterp2.argtypes = [FP, FP, FP, FP, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, ctypes.c_float]
        lib.ref_bilinInterp2(_p(out), _p(Iin), _p(X), _p(Y), Iin.shape[0], Iin.shape[1], nf, ctypes.c_float(oob))
        return out


# ---------------------------------------------------------------------------------------------
def _p(a: np.ndarray):
    return a.ctypes.data_as(FP)


def _f(a) -> np.ndarray:
    a = np.asarray(a)
    if a.dtype != np.float32:
        raise MexError("must be a noncomplex single-valued matrix")
    if a.ndim < 2:
        a = a.reshape(-1, 1) if a.ndim == 1 else a.reshape(1, 1)
    return np.asfortranarray(a)


def _scalar(a) -> float:
    a = np.asarray(a)
    if a.dtype != np.float32:
        raise MexError("must be a noncomplex, single-type scalar")
    return float(a.reshape(-1)[0])


class OracleBackend:
    """Plain-C restatement + a restatement of each gateway's marshalling."""
    name = "oracle"

    def __init__(self):
        path = os.path.join(HERE, "libpde_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = ctypes.CDLL(path)
        for fn in ("orc_flow_gs", "orc_flow_alr", "orc_flow_alr8", "orc_flow_operator", "orc_disp_gs", "orc_disp_gs_sym",
                   "orc_disp_alr", "orc_pde_gs", "orc_pde_alr", "orc_bilin", "orc_fst", "orc_snd", "orc_ddiff"):
            getattr(self.lib, fn).restype = None

    # -- helpers ---------------------------------------------------------------------------
    @staticmethod
    def _dims(a):
        return a.shape[0], a.shape[1], (a.shape[2] if a.ndim > 2 else 1)

    def call(self, fn: str, args: Sequence, nlhs: int) -> List[np.ndarray]:
        return getattr(self, "_" + fn)(list(args), nlhs)

    # -- optical flow ------------------------------------------------------------------------
    def _flow_sor(self, late, fields, iter_, omega, solver, nlhs, name):
        if late:
            U, V, X0, X1, M, Cu, Cv, Du, Dv, wW, wN, wE, wS = fields
        else:
            X0, X1, M, Cu, Cv, Du, Dv, wW, wN, wE, wS = fields
            U = V = None
        if nlhs < 2:
            raise MexError(name + " insufficient number of outputs")
        if solver not in (1, 2):
            raise MexError(name + ": no such solver")
        nr, nc, nf = self._dims(M) if late else (X0.shape[0], X0.shape[1], self._dims(M)[2])
        it = int(iter_)
        o0 = np.zeros(X0.shape, np.float32, order="F")
        o1 = np.zeros(X1.shape, np.float32, order="F")
        if it > 0:
            o0[...] = X0
            o1[...] = X1
            f0 = _p(U) if late else None
            f1 = _p(V) if late else None
            fnc = self.lib.orc_flow_gs if solver == 1 else self.lib.orc_flow_alr
            fnc(ctypes.c_int(int(late)), _p(o0), _p(o1), f0, f1, _p(M), _p(Cu), _p(Cv), _p(Du), _p(Dv),
                _p(wW), _p(wN), _p(wE), _p(wS), nr, nc, it, ctypes.c_float(omega))
        outs = [o0, o1]
        if nlhs >= 4:
            RU = np.zeros(M.shape, np.float32, order="F")
            RV = np.zeros(M.shape, np.float32, order="F")
            self.lib.orc_flow_operator(int(late), 0, 1, _p(RU), _p(RV), _p(X0), _p(X1),
                                       _p(U) if late else None, _p(V) if late else None,
                                       _p(M), _p(Cu), _p(Cv), _p(Du), _p(Dv), _p(wW), _p(wN), _p(wE), _p(wS), nr, nc, nf)
            outs += [RU, RV]
        return outs

    def _Oflow_sor_elin4_2d(self, a, nlhs):
        if len(a) != 14:
            raise MexError("Oflow_sor_elin4_2d parameter error: wrong number of input parameters!")
        f = [_f(x) for x in a[:11]]
        return self._flow_sor(False, f, _scalar(a[11]), _scalar(a[12]), int(_scalar(a[13])), nlhs, "Oflow_sor_elin4_2d")

    def _Oflow_sor_llin4_2d(self, a, nlhs):
        if len(a) != 16:
            raise MexError("Oflow_sor_llin4_2d parameter error: wrong number of input parameters!")
        f = [_f(x) for x in a[:13]]
        return self._flow_sor(True, f, _scalar(a[13]), _scalar(a[14]), int(_scalar(a[15])), nlhs, "Oflow_sor_llin4_2d")

    def _Oflow_sor_llin8_2d(self, a, nlhs):
        if len(a) != 20:
            raise MexError("Oflow_sor_llin8_2d parameter error: wrong number of input parameters!")
        f = [_f(x) for x in a[:17]]
        U, V, dU, dV, M, Cu, Cv, Du, Dv, wW, wNW, wN, wNE, wE, wSE, wS, wSW = f
        it, omega, solver = int(_scalar(a[17])), _scalar(a[18]), int(_scalar(a[19]))
        if nlhs < 2:
            raise MexError("Oflow_sor_llin8_2d insufficient number of outputs")
        if solver not in (1, 2):
            raise MexError("Oflow_sor_llin8_2d: no such solver")
        nr, nc, _ = self._dims(M)
        o0 = np.zeros(dU.shape, np.float32, order="F")
        o1 = np.zeros(dV.shape, np.float32, order="F")
        if it > 0:
            o0[...] = dU
            o1[...] = dV
            if solver == 1:   # the point solver never reads the diagonal weights
                self.lib.orc_flow_gs(1, _p(o0), _p(o1), _p(U), _p(V), _p(M), _p(Cu), _p(Cv), _p(Du), _p(Dv),
                                     _p(wW), _p(wN), _p(wE), _p(wS), nr, nc, it, ctypes.c_float(omega))
            else:
                w8 = (FP * 8)(_p(wW), _p(wN), _p(wE), _p(wS), _p(wNW), _p(wNE), _p(wSE), _p(wSW))
                self.lib.orc_flow_alr8(_p(o0), _p(o1), _p(U), _p(V), _p(M), _p(Cu), _p(Cv), _p(Du), _p(Dv), w8,
                                       nr, nc, it, ctypes.c_float(omega))
        outs = [o0, o1]
        if nlhs >= 4:   # created, never filled (Oflow_sor_llin8_2d.c:465-488)
            outs += [np.zeros(M.shape, np.float32, order="F"), np.zeros(M.shape, np.float32, order="F")]
        return outs

    def _flow_lhs(self, late, f, nlhs, name):
        if late:
            U, V, X0, X1, M, Du, Dv, wW, wN, wE, wS = f
        else:
            X0, X1, M, Du, Dv, wW, wN, wE, wS = f
            U = V = None
        if nlhs < 2:
            raise MexError(name + " insufficient number of outputs")
        nr, nc, nf = self._dims(M)
        AU = np.zeros(M.shape, np.float32, order="F")
        AV = np.zeros(M.shape, np.float32, order="F")
        self.lib.orc_flow_operator(int(late), 1, 1, _p(AU), _p(AV), _p(X0), _p(X1),
                                   _p(U) if late else None, _p(V) if late else None,
                                   _p(M), _p(Du), _p(Dv), _p(Du), _p(Dv), _p(wW), _p(wN), _p(wE), _p(wS), nr, nc, nf)
        return [AU, AV]

    def _Oflow_lhs_elin4_2d(self, a, nlhs):
        if len(a) != 9:
            raise MexError("Oflow_lhs_elin4_2d parameter error: wrong number of input parameters!")
        return self._flow_lhs(False, [_f(x) for x in a], nlhs, "Oflow_lhs_elin4_2d")

    def _Oflow_lhs_llin4_2d(self, a, nlhs):
        if len(a) != 11:
            raise MexError("Oflow_lhs_llin4_2d parameter error: wrong number of input parameters!")
        return self._flow_lhs(True, [_f(x) for x in a], nlhs, "Oflow_lhs_llin4_2d")

    # -- disparity ---------------------------------------------------------------------------
    def _Disp_sor_llin4_2d(self, a, nlhs):
        if len(a) != 11:
            raise MexError("Disp_sor_llin4_2d parameter error: wrong number of input parameters!")
        U, dU, Cu, Du, wW, wN, wE, wS = [_f(x) for x in a[:8]]
        it, omega, solver = int(_scalar(a[8])), _scalar(a[9]), int(_scalar(a[10]))
        if nlhs < 1:
            raise MexError("Disp_sor_llin4_2d insufficient number of outputs")
        if solver not in (1, 2):
            raise MexError("Disp_sor_llin4_2d: no such solver")
        nr, nc, _ = self._dims(Cu)
        o = np.zeros(dU.shape, np.float32, order="F")
        if it > 0:
            o[...] = dU
            fnc = self.lib.orc_disp_gs if solver == 1 else self.lib.orc_disp_alr
            fnc(_p(o), _p(U), _p(Cu), _p(Du), _p(wW), _p(wN), _p(wE), _p(wS), nr, nc, it, ctypes.c_float(omega))
        outs = [o]
        if nlhs >= 2:   # allocated, never filled (Disp_sor_llin4_2d.c:251-269)
            outs.append(np.zeros(U.shape, np.float32, order="F"))
        return outs

    def _Disp_sor_llin_sym4_2d(self, a, nlhs):
        if len(a) != 19:
            raise MexError("Disp_sor_llin_sym4_2d parameter error: wrong number of input parameters!")
        f = [_f(x) for x in a[:16]]
        it, omega, solver = int(_scalar(a[16])), _scalar(a[17]), int(_scalar(a[18]))
        if nlhs < 2:
            raise MexError("Disp_sor_llin_sym4_2d insufficient number of outputs")
        if solver not in (1, 2):
            raise MexError("Disp_sor_llin_sym4_2d: no such solver")
        nr, nc, _ = self._dims(f[2])
        o0 = f[1].copy(order="F")
        o1 = f[9].copy(order="F")
        if it > 0:
            if solver == 1:
                self.lib.orc_disp_gs_sym(_p(o0), _p(f[0]), _p(f[2]), _p(f[3]), _p(f[4]), _p(f[5]), _p(f[6]), _p(f[7]),
                                         _p(o1), _p(f[8]), _p(f[10]), _p(f[11]), _p(f[12]), _p(f[13]), _p(f[14]), _p(f[15]),
                                         nr, nc, it, ctypes.c_float(omega))
            else:       # the two fields never interact (disparitySolvers.c:510-539)
                self.lib.orc_disp_alr(_p(o0), _p(f[0]), _p(f[2]), _p(f[3]), _p(f[4]), _p(f[5]), _p(f[6]), _p(f[7]),
                                      nr, nc, it, ctypes.c_float(omega))
                self.lib.orc_disp_alr(_p(o1), _p(f[8]), _p(f[10]), _p(f[11]), _p(f[12]), _p(f[13]), _p(f[14]), _p(f[15]),
                                      nr, nc, it, ctypes.c_float(omega))
        return [o0, o1]

    # -- generic PDE -------------------------------------------------------------------------
    def _pde(self, eight, a, nlhs, name):
        nin = 11 if eight else 7
        if len(a) != nin + 3:
            raise MexError("error: wrong number of input parameters!")
        f = [_f(x) for x in a[:nin]]
        it, omega, solver = int(_scalar(a[nin])), _scalar(a[nin + 1]), int(_scalar(a[nin + 2]))
        if nlhs < 1:
            raise MexError("error insufficient number of outputs.")
        if solver not in (1, 2):
            raise MexError("error: no such solver")
        X, TRACE, B = f[0], f[1], f[2]
        if eight:
            wW, wNW, wN, wNE, wE, wSE, wS, wSW = f[3:]
            w = (FP * 8)(_p(wW), _p(wN), _p(wE), _p(wS), _p(wNW), _p(wNE), _p(wSE), _p(wSW))
        else:
            wW, wN, wE, wS = f[3:]
            w = (FP * 8)(_p(wW), _p(wN), _p(wE), _p(wS), None, None, None, None)
        nr, nc, nf = self._dims(X)
        o = X.copy(order="F")
        fnc = self.lib.orc_pde_gs if solver == 1 else self.lib.orc_pde_alr
        fnc(int(eight), _p(o), _p(TRACE), _p(B), w, nr, nc, nf, it, ctypes.c_float(omega))
        return [o]

    def _PDEsolver4(self, a, nlhs):
        return self._pde(False, a, nlhs, "PDEsolver4")

    def _PDEsolver8(self, a, nlhs):
        return self._pde(True, a, nlhs, "PDEsolver8")

    # -- streaming kernels -------------------------------------------------------------------
    def bilin(self, Iin, X, Y, oob: float) -> np.ndarray:
        Iin, X, Y = _f(Iin), _f(X), _f(Y)
        nr, nc, nf = self._dims(Iin)
        out = np.zeros(Iin.shape, np.float32, order="F")
        self.lib.orc_bilin(_p(out), _p(Iin), _p(X), _p(Y), nr, nc, nf, ctypes.c_float(oob))
        return out

    def _BilinInterp_2d(self, a, nlhs):
        if len(a) != 3:
            raise MexError("proper function call is 'bilinInterp2( Iin, X, Y)'")
        if nlhs < 1:
            raise MexError("insufficient number of outputs")
        return [self.bilin(a[0], a[1], a[2], float("nan"))]

    def _FstDerivatives5(self, a, nlhs):
        if len(a) != 2:
            raise MexError("fstDerivatives: wrong number of input parameters!")
        if nlhs < 3:
            raise MexError("fstDerivatives: insufficient number of outputs")
        I0, I1 = _f(a[0]), _f(a[1])
        nr, nc, nf = self._dims(I0)
        o = [np.zeros(I0.shape, np.float32, order="F") for _ in range(3)]
        self.lib.orc_fst(_p(o[0]), _p(o[1]), _p(o[2]), _p(I0), _p(I1), nr, nc, nf)
        return o

    def _SndDerivatives5(self, a, nlhs):
        if len(a) != 2:
            raise MexError("sndDerivatives: wrong number of input parameters!")
        if nlhs < 5:
            raise MexError("sndDerivatives: insufficient number of outputs")
        I0, I1 = _f(a[0]), _f(a[1])
        nr, nc, nf = self._dims(I0)
        o = [np.zeros(I0.shape, np.float32, order="F") for _ in range(5)]
        self.lib.orc_snd(_p(o[0]), _p(o[1]), _p(o[2]), _p(o[3]), _p(o[4]), _p(I0), _p(I1), nr, nc, nf)
        return o

    def _DdiffWeights(self, a, nlhs):
        if len(a) != 2:
            raise MexError("diffusion6_2d parameter error: wrong number of input parameters!")
        if nlhs < 4:
            raise MexError("diffusion6_2d error insufficient number of outputs")
        D = _f(a[0])
        eps = _scalar(a[1])
        nr, nc, nf = self._dims(D)
        o = [np.zeros(D.shape, np.float32, order="F") for _ in range(4)]
        tmp = [np.zeros((nr, nc), np.float32, order="F") for _ in range(4)]
        self.lib.orc_ddiff(_p(tmp[0]), _p(tmp[1]), _p(tmp[2]), _p(tmp[3]), _p(D), nr, nc, nf, ctypes.c_float(eps))
        for k in range(4):
            if D.ndim > 2:
                o[k][:, :, 0] = tmp[k]
            else:
                o[k][...] = tmp[k]
        return o
