"""CPU restatement of the reference's DRIVERS (matlab/*/*.m) around the MEX calls.  TEST INFRASTRUCTURE ONLY.

`flow_llin` follows matlab/optical_flow/FlowEminND_llin_2D_v10.m line by line; every MEX call goes through a
backend with the common calling convention (oracle.RefBackend = the unmodified reference C code,
oracle.OracleBackend = the C restatement), every Matlab-side formula through oracle/matlab_steps.py.
Toolbox steps are "parity unpinned" (see matlab_steps.py); the optional spatial a-priori inputs are omitted."""
from __future__ import annotations

import math

import numpy as np

from . import matlab_steps as ms

F32 = np.float32


def flow_llin(I0, I1, backend, alpha=0.0420, omega=1.9, firstLoop=4, secondLoop=4, iter=4, b1=1.4843, b2=0.2915,
              scl_factor=0.75, solver=2, fst_grad=True, snd_term="gradmag", max_scales=None, oob=np.nan):
    """[U V] = FlowEminND_llin_2D_v10(cat(3, I0, I1), channels, fstTerm, sndTerm). I0, I1: rows x cols x channels, 0..255."""
    I0 = (np.asarray(I0, dtype=F32).reshape(I0.shape[0], I0.shape[1], -1) / F32(255)).astype(F32)      # :73
    I1 = (np.asarray(I1, dtype=F32).reshape(I1.shape[0], I1.shape[1], -1) / F32(255)).astype(F32)
    G = ms.fspecial_gaussian(5, 1.25)                                                                   # :99
    It0, It1 = [I0], [I1]
    scales = max_scales or (1 << 30)
    while len(It0) < scales:                                                                            # :105-127
        n0, n1 = ms.imresize_bilinear(It0[-1], scale=scl_factor), ms.imresize_bilinear(It1[-1], scale=scl_factor)
        It0[-1], It1[-1] = ms.imfilter(It0[-1], G), ms.imfilter(It1[-1], G)
        It0.append(n0); It1.append(n1)
        if n0.shape[0] <= 20 or n0.shape[1] <= 20:
            It0[-1], It1[-1] = ms.imfilter(It0[-1], G), ms.imfilter(It1[-1], G)
            break
    S = len(It0)
    F0 = [ms.rgb2grad(x) if fst_grad else x for x in It0]                                               # :135-147
    F1 = [ms.rgb2grad(x) if fst_grad else x for x in It1]
    U = V = None
    for s in range(S - 1, -1, -1):                                                                      # :195
        rows, cols = It0[s].shape[:2]
        Xg, Yg = np.meshgrid(np.arange(1, cols + 1, dtype=F32), np.arange(1, rows + 1, dtype=F32))
        if U is None:
            U = np.zeros((rows, cols), F32); V = np.zeros((rows, cols), F32)
        for _ in range(firstLoop):                                                                      # :217
            X, Y = (Xg + U).astype(F32), (Yg + V).astype(F32)
            W1 = backend.bilin(F1[s], X, Y, oob)                                                        # :222
            d1 = backend.call("FstDerivatives5", [F0[s], W1], 3)                                        # :234
            d2 = None
            if snd_term != "none":
                W2 = backend.bilin(It1[s], X, Y, oob)                                                   # :228
                d2 = backend.call("SndDerivatives5", [It0[s], W2], 5) if snd_term == "gradmag" else \
                    backend.call("FstDerivatives5", [It0[s], W2], 3)                                    # :245,253
            dU = np.zeros((rows, cols), F32); dV = np.zeros((rows, cols), F32)
            for _ in range(secondLoop):                                                                 # :278
                wW, wN, wS, wE = [w.astype(F32) for w in ms.op_diff_weights((U + dU).astype(F32), (V + dV).astype(F32))]   # :321
                M, Cu, Cv, Du, Dv = ms.llin_terms(d1, d2, dU, dV, b1, b2, alpha, snd_term == "gradmag")  # :289-327
                dU, dV = backend.call("Oflow_sor_llin4_2d", [U, V, dU, dV, M, Cu, Cv, Du, Dv, wW, wN, wE, wS,
                                                             F32(iter), F32(omega), F32(solver)], 2)    # :332-348
            U = ms.medfilt2_symmetric((U + dU).astype(F32))                                             # :354-355
            V = ms.medfilt2_symmetric((V + dV).astype(F32))
        if s > 0:                                                                                       # :364-367
            up = F32(1.0 / scl_factor)
            size = It0[s - 1].shape[:2]
            U = ms.imresize_bilinear((U * up).astype(F32), output_size=size)
            V = ms.imresize_bilinear((V * up).astype(F32), output_size=size)
    return U, V


def tvdenoise8(I_in, backend, alpha=500.0, omega=1.75, outer_iter=20, inner_iter=4, solver=2, scl_factor=0.75):
    """Iout = TVdenoise8(I_in), matlab/denoising/TVdenoise8.m:36-115, for a single image rows x cols (x frames)."""
    I_in = np.asarray(I_in, dtype=F32)
    shape = I_in.shape
    I0 = I_in.reshape(shape[0], shape[1], -1)
    G = ms.fspecial_gaussian(5, 1.25)
    I1 = ms.imresize_bilinear(I0, scale=scl_factor)                    # :60  (from the unsmoothed image)
    I0 = ms.imfilter(I0, G)                                            # :65
    # :68-74: the size test is met at scl = 2; the smoothed coarsest level is assigned to `Itin`, i.e. dropped
    Iin = [I0, I1]
    Iout = Iin[1].copy()
    for s in (1, 0):
        for _ in range(outer_iter + 1):                                # :79  iter = 0:outer_iter
            w = ms.ad_diff_weights(Iout)                               # :81
            TRACE, B, aw = ms.tv_terms(Iout, Iin[s], w[:8], alpha)     # :83-85
            fr = Iout.shape[2]
            aw3 = [np.repeat(x[:, :, None], fr, axis=2) for x in aw]
            Iout = backend.call("PDEsolver8", [Iout, TRACE, B] + aw3 + [F32(inner_iter), F32(omega), F32(solver)], 1)[0]   # :87-100
            Iout = np.asarray(Iout, dtype=F32).reshape(Iin[s].shape)
        if s > 0:
            Iout = ms.imresize_bilinear(Iout, output_size=Iin[0].shape[:2])   # :112
    return Iout.reshape(shape)


def flow_fmg(I0, I1, backend, alpha=0.035, omega=1.9, firstLoop=4, iter=4, b1=0.03, b2=0.97, scl_factor=0.5, solver=2,
             cycle_index=1, max_scales=None):
    """[U V] = FlowEminNDFASFMG_elin_2D_v10(cat(3, I0, I1), channels): full multigrid (FMG) over a lpf pyramid, one FAS
    V-cycle (cycle_index = 1) or W-cycle (2) per level. I0, I1: rows x cols x channels, 0..255.
    matlab/optical_flow/FlowEminNDFASFMG_elin_2D_v10.m, line numbers in the comments."""
    I0 = np.asarray(I0, dtype=F32).reshape(I0.shape[0], I0.shape[1], -1)                               # :72
    I1 = np.asarray(I1, dtype=F32).reshape(I1.shape[0], I1.shape[1], -1)
    G = ms.fspecial_gaussian(5, 1.0)                                                                    # :97
    It0, It1 = [ms.imfilter(I0, G, conv=True)], [ms.imfilter(I1, G, conv=True)]                         # :103-104
    scales = max_scales or (1 << 30)
    while len(It0) < scales:                                                                            # :106-118
        It0.append(ms.lpf_decimate(It0[-1])); It1.append(ms.lpf_decimate(It1[-1]))
        if It0[-1].shape[0] <= 10 or It0[-1].shape[1] <= 10:
            break
    S = len(It0)
    der = [ms.fmg_derivatives(It0[s], It1[s]) for s in range(S)]                                        # :123-141
    coef = [ms.fmg_terms(der[s], b1, b2) for s in range(S)]                                             # :143-149
    P = dict(alpha=alpha, omega=omega, firstLoop=firstLoop, iter=iter, b1=b1, b2=b2, scl_factor=scl_factor,
             solver=solver, cycle_index=cycle_index)

    def smooth(U, V, s, Cu, Cv, want_res):                                                              # :367-464
        M, _, _, Du, Dv = coef[s]
        for _ in range(P["firstLoop"]):
            _, MG, CuG, CvG, DuG, DvG = ms.elin_terms(der[s], (M, Cu, Cv, Du, Dv), U, V, b1, b2, alpha, True)
            wW, wN, wS, wE = [w.astype(F32) for w in ms.op_diff_weights(U, V)]
            U, V = backend.call("Oflow_sor_elin4_2d", [U, V, MG, CuG, CvG, DuG, DvG, wW, wN, wE, wS,
                                                       F32(P["iter"]), F32(omega), F32(solver)], 2)
        if not want_res:
            return U, V, None, None
        _, MG, CuG, CvG, DuG, DvG = ms.elin_terms(der[s], (M, Cu, Cv, Du, Dv), U, V, b1, b2, alpha, False)
        wW, wN, wS, wE = [w.astype(F32) for w in ms.op_diff_weights(U, V)]
        out = backend.call("Oflow_sor_elin4_2d", [U, V, MG, CuG, CvG, DuG, DvG, wW, wN, wE, wS,
                                                  F32(0), F32(omega), F32(solver)], 4)                # iter = 0: residuals only (:446-460)
        ch = der[s][0].shape[2]
        return U, V, out[2].reshape(U.shape + (ch,)), out[3].reshape(U.shape + (ch,))

    def fas_cycle(U, V, Cu, Cv, s):                                                                     # :193-273
        Uc = None
        if s < S - 1:
            for _ in range(P["cycle_index"]):
                U, V, RU, RV = smooth(U, V, s, Cu, Cv, True)                                            # :207
                RUr, RVr = ms.fw_restrict(RU, scl_factor), ms.fw_restrict(RV, scl_factor)               # :212-213
                Ur, Vr = ms.fw_restrict(U, scl_factor), ms.fw_restrict(V, scl_factor)                   # :216-217
                Mc, _, _, Duc, Dvc = coef[s + 1]
                gd, MG, _, _, DuG, DvG = ms.elin_terms(der[s + 1], (Mc, Mc, Mc, Duc, Dvc), Ur, Vr, b1, b2, alpha, False)   # :225-237
                wW, wN, wS, wE = [w.astype(F32) for w in ms.op_diff_weights(Ur, Vr)]                    # :234
                Au, Av = backend.call("Oflow_lhs_elin4_2d", [Ur, Vr, MG, DuG, DvG, wW, wN, wE, wS], 2)  # :239-248
                ch = gd.shape[2]
                fu = ms.fas_rhs(RUr.reshape(gd.shape), Au.reshape(gd.shape), gd)                        # :250-251
                fv = ms.fas_rhs(RVr.reshape(gd.shape), Av.reshape(gd.shape), gd)
                Uc, Vc = fas_cycle(Ur, Vr, fu, fv, s + 1)                                               # :253
                up = F32(1.0 / scl_factor)
                U = (U + ms.imresize_bilinear(((Uc - Ur) * up).astype(F32), output_size=U.shape)).astype(F32)   # :256-257
                V = (V + ms.imresize_bilinear(((Vc - Vr) * up).astype(F32), output_size=V.shape)).astype(F32)
        else:
            U, V, _, _ = smooth(U, V, s, Cu, Cv, False)                                                 # :261-263
        if Uc is not None:
            U, V, _, _ = smooth(U, V, s, Cu, Cv, False)                                                 # :269
        return U, V

    U = V = None
    for s in range(S - 1, -1, -1):                                                                      # :158
        rows, cols = It0[s].shape[:2]
        if U is None:
            U = np.zeros((rows, cols), F32); V = np.zeros((rows, cols), F32)
        U, V = fas_cycle(U, V, coef[s][1], coef[s][2], s)                                               # :174
        if s > 0:                                                                                       # :179-182
            up = F32(1.0 / scl_factor)
            size = It0[s - 1].shape[:2]
            U = ms.imresize_bicubic((U * up).astype(F32), output_size=size)
            V = ms.imresize_bicubic((V * up).astype(F32), output_size=size)
    return U, V


def flow_hs(I0, I1, backend, alpha=0.2, omega=1.9, iter=20, b1=0.25, b2=0.75, scl_factor=0.75, solver=2, max_scales=None):
    """[U V] = FlowEminHS_elin_2D_v10(cat(3, I0, I1), channels): Horn-Schunck (quadratic data and smoothness terms, one
    linear solve per pyramid level), BASELINE configs[0]. I0, I1: rows x cols x channels, 0..255.
    matlab/optical_flow/FlowEminHS_elin_2D_v10.m, line numbers in the comments."""
    I0 = (np.asarray(I0, dtype=F32).reshape(I0.shape[0], I0.shape[1], -1) / F32(255)).astype(F32)     # :66
    I1 = (np.asarray(I1, dtype=F32).reshape(I1.shape[0], I1.shape[1], -1) / F32(255)).astype(F32)
    ch = I0.shape[2]
    G = ms.fspecial_gaussian(5, 1.25)                                                                   # :88
    It0, It1 = [I0], [I1]
    scales = max_scales or (1 << 30)
    while len(It0) < scales:                                                                            # :96-115
        n0, n1 = ms.imresize_bilinear(It0[-1], scale=scl_factor), ms.imresize_bilinear(It1[-1], scale=scl_factor)
        It0[-1], It1[-1] = ms.imfilter(It0[-1], G), ms.imfilter(It1[-1], G)
        It0.append(n0); It1.append(n1)
        if n0.shape[0] <= 20 or n0.shape[1] <= 20:
            It0[-1], It1[-1] = ms.imfilter(It0[-1], G), ms.imfilter(It1[-1], G)
            break
    # (with max_scales reached before the size test the last level stays unsmoothed, as in the driver)
    pre = np.array([[0.037659, 0.249724, 0.439911, 0.249724, 0.037659]])
    odx = np.array([[0.104550, 0.292315, 0.0, -0.292315, -0.104550]])
    oxx = np.array([[0.232905, 0.002668, -0.471147, 0.002668, 0.232905]])
    f = lambda A, h: ms.imfilter(A, h, "replicate", conv=True)
    U = V = None
    for s in range(len(It0) - 1, -1, -1):                                                               # :123
        A0, A1 = It0[s], It1[s]
        rows, cols = A0.shape[:2]
        W = np.full((rows, cols), alpha * ch, dtype=np.float64).astype(F32)                             # :127
        if U is None:
            U = np.zeros((rows, cols), F32); V = np.zeros((rows, cols), F32)
        Ist = ((A0 + A1).astype(F32) * F32(0.55)).astype(F32)                                           # :139
        Idt = (A0 - A1).astype(F32)
        Idx = f(f(Ist, pre.T), odx); Idy = f(f(Ist, pre), odx.T)                                        # :142-146
        Idxx = f(f(Ist, pre.T), oxx); Idyy = f(f(Ist, pre), oxx.T)
        Idxy = f(f(Ist, odx), odx.T)
        Idxt = (f(f(A0, pre.T), odx) - f(f(A1, pre.T), odx)).astype(F32)                                # :148-154
        Idyt = (f(f(A0, pre), odx.T) - f(f(A1, pre), odx.T)).astype(F32)
        terms = ms.fmg_terms((Idt, Idx, Idy, Idxt, Idyt, Idxx, Idyy, Idxy), b1, b2)                     # :159-163 (same formulas)
        summed = []
        for t in terms:                                                                                 # :168-172 sum(.,3)
            acc = t[:, :, 0].copy()
            for k in range(1, ch):
                acc = (acc + t[:, :, k]).astype(F32)
            summed.append(acc)
        U, V = backend.call("Oflow_sor_elin4_2d", [U, V] + summed + [W, W, W, W, F32(iter), F32(omega), F32(solver)], 2)   # :174-188
        if s > 0:                                                                                       # :193-196
            up = F32(1.0 / scl_factor)
            size = It0[s - 1].shape[:2]
            U = ms.imresize_bicubic(ms.medfilt2_symmetric((U * up).astype(F32)), output_size=size)
            V = ms.imresize_bicubic(ms.medfilt2_symmetric((V * up).astype(F32)), output_size=size)
    return U, V


def disp_sym(Il, Ir, backend, alpha=0.035, beta=0.4, omega=1.9, firstLoop=3, secondLoop=4, iter=4, b1=0.25, b2=0.72,
             scl_factor=0.75, solver=2, max_scales=None, uint8_input=True, oob=np.nan):
    """U = DispEminND_llin_sym_2D(Il, Ir): symmetric stereo, two disparity fields U(:,:,1) (left -> right) and U(:,:,2)
    (right -> left) coupled by a symmetry term (BASELINE configs[3]). Il, Ir: rows x cols x channels, 0..255; with
    uint8_input the pyramid is kept in class uint8 as for the imread images of runme.m:17-28 (toolbox results rounded).
    matlab/disparity/DispEminND_llin_sym_2D.m, line numbers in the comments. Returns (U0, U1)."""
    q = ms.round_uint8 if uint8_input else (lambda A: np.asarray(A, dtype=F32))
    It0 = [q(np.asarray(Il, dtype=F32).reshape(Il.shape[0], Il.shape[1], -1))]                          # :86-87
    It1 = [q(np.asarray(Ir, dtype=F32).reshape(Ir.shape[0], Ir.shape[1], -1))]
    G = ms.fspecial_gaussian(3, 1.0)                                                                    # :81
    scales = max_scales or (1 << 30)
    while len(It0) < scales:                                                                            # :89-103
        n0, n1 = q(ms.imresize_bilinear(It0[-1], scale=scl_factor)), q(ms.imresize_bilinear(It1[-1], scale=scl_factor))
        It0[-1], It1[-1] = q(ms.imfilter(It0[-1], G)), q(ms.imfilter(It1[-1], G))
        It0.append(n0); It1.append(n1)
        if n0.shape[0] <= 10 or n0.shape[1] <= 10:
            break
    S = len(It0)
    pre = np.array([[0.037659, 0.249724, 0.439911, 0.249724, 0.037659]])
    odx = np.array([[0.104550, 0.292315, 0.0, -0.292315, -0.104550]])
    f = lambda A, h: ms.imfilter(A, h, "replicate", conv=True)
    U0 = U1 = None
    for s in range(S - 1, -1, -1):                                                                      # :111
        rows, cols = It0[s].shape[:2]
        Xg, Yg = np.meshgrid(np.arange(1, cols + 1, dtype=F32), np.arange(1, rows + 1, dtype=F32))
        if U0 is None:
            U0 = np.zeros((rows, cols), F32); U1 = np.zeros((rows, cols), F32)
        srDiff = 2.0 * (1.0 / scl_factor) ** -s                                                         # :127 (scl = s + 1)
        for _ in range(firstLoop):                                                                      # :133
            It0w = backend.bilin(It0[s], (Xg + U1).astype(F32), Yg, oob)                                # :138-139
            It1w = backend.bilin(It1[s], (Xg + U0).astype(F32), Yg, oob)
            U0w = ms.interp2_rows(U0, Xg.astype(np.float64) + U1)                                       # :144-145 (double)
            U1w = ms.interp2_rows(U1, Xg.astype(np.float64) + U0)
            Idt0, Idx1, _ = backend.call("FstDerivatives5", [It0[s], It1w], 3)                          # :150-154
            Idxt0, Idyt0, Idxx1, _, Idxy1 = backend.call("SndDerivatives5", [It0[s], It1w], 5)
            Idt1, Idx0, _ = backend.call("FstDerivatives5", [It1[s], It0w], 3)
            Idxt1, Idyt1, Idxx0, _, Idxy0 = backend.call("SndDerivatives5", [It1[s], It0w], 5)
            Udt0 = ((U0 + U1w) * F32(0.5)).astype(F32)                                                  # :159-165
            Udx1 = f(f(U1w, pre.T), odx)
            Udt1 = ((U1 + U0w) * F32(0.5)).astype(F32)
            Udx0 = f(f(U0w, pre.T), odx)
            dU0 = np.zeros((rows, cols), F32); dU1 = np.zeros((rows, cols), F32)
            for _ in range(secondLoop):                                                                 # :188
                CuG0, DuG0 = ms.disp_sym_terms((Idt0, Idx1, Idxt0, Idyt0, Idxx1, Idxy1), dU0, Udt0, Udx1, b1, b2, alpha, beta, srDiff)
                CuG1, DuG1 = ms.disp_sym_terms((Idt1, Idx0, Idxt1, Idyt1, Idxx0, Idxy0), dU1, Udt1, Udx0, b1, b2, alpha, beta, srDiff)
                w0 = backend.call("DdiffWeights", [(U0 + dU0).astype(F32), F32(0.00001)], 4)            # :219-220 [wW wN wE wS]
                w1 = backend.call("DdiffWeights", [(U1 + dU1).astype(F32), F32(0.00001)], 4)
                dU0, dU1 = backend.call("Disp_sor_llin_sym4_2d", [U0, dU0, CuG0, DuG0] + list(w0) + [U1, dU1, CuG1, DuG1] + list(w1)
                                        + [F32(iter), F32(omega), F32(solver)], 2)                       # :227-247
            U0 = ms.medfilt2_symmetric((U0 + dU0).astype(F32))                                          # :255-256
            U1 = ms.medfilt2_symmetric((U1 + dU1).astype(F32))
        if s > 0:                                                                                       # :265-267
            up = F32(1.0 / scl_factor)
            size = It0[s - 1].shape[:2]
            U0 = ms.imresize_bilinear((U0 * up).astype(F32), output_size=size)
            U1 = ms.imresize_bilinear((U1 * up).astype(F32), output_size=size)
    return U0, U1


# ------------------------------------------------------------------------------------------------------------------
# sibling drivers runme.m also calls: same gateways, other Matlab glue
# ------------------------------------------------------------------------------------------------------------------
def tvdenoise4(I_in, backend, alpha=5.0, omega=1.75, outer_iter=10, inner_iter=5, solver=2, scl=0.5, scl_factor=0.75):
    """Iout = TVdenoise4(I_in), matlab/denoising/TVdenoise4.m:36-112: multi-scale lagged-diffusivity TV denoising with
    the 4-neighbour weights of its own DiffWeights (:116) and PDEsolver4. I_in: rows x cols (x frames), single."""
    I_in = np.asarray(I_in, dtype=F32)
    shape = I_in.shape
    I0 = I_in.reshape(shape[0], shape[1], -1)
    ds_rows, ds_cols = math.ceil(shape[0] * scl), math.ceil(shape[1] * scl)                            # :52-53
    G = ms.fspecial_gaussian(7, 2.0)                                                                    # :57
    Iin = [I0]
    while True:                                                                                         # :61-77
        nxt = ms.imresize_bilinear(Iin[-1], scale=scl_factor)
        Iin[-1] = ms.imfilter(Iin[-1], G)
        Iin.append(nxt)
        if nxt.shape[0] <= ds_rows or nxt.shape[1] <= ds_cols:
            Iin[-1] = ms.imfilter(Iin[-1], G)
            break
    Iout = Iin[-1].copy()                                                                               # :78
    eps = F32(np.finfo(np.float64).eps)
    for s in range(len(Iin) - 1, -1, -1):                                                               # :80
        for _ in range(outer_iter + 1):                                                                 # :82  iter = 0:outer_iter
            df = (Iout - Iin[s]).astype(F32)
            psi = (F32(1) / np.sqrt(((df * df).astype(F32) + eps).astype(F32))).astype(F32)             # :84
            wW, wN, wE, wS = ms.tv4_diff_weights(Iout)                                                  # :85
            sw = (((wW + wN).astype(F32) + wE).astype(F32) + wS).astype(F32)
            TRACE = (psi + (F32(alpha) * sw).astype(F32)[:, :, None]).astype(F32)                       # :87
            B = (psi * Iin[s]).astype(F32)                                                              # :88
            fr = Iout.shape[2]
            aw = [np.repeat((F32(alpha) * w).astype(F32)[:, :, None], fr, axis=2) for w in (wW, wN, wE, wS)]
            Iout = backend.call("PDEsolver4", [Iout, TRACE, B] + aw + [F32(inner_iter), F32(omega), F32(solver)], 1)[0]   # :90-99
            Iout = np.asarray(Iout, dtype=F32).reshape(Iin[s].shape)
        if s > 0:                                                                                       # :108-110
            Iout = ms.imresize_bilinear(Iout, output_size=Iin[s - 1].shape[:2])
    return Iout.reshape(shape)


def disp_llin(Il, Ir, backend, alpha=0.042, omega=1.9, firstLoop=4, secondLoop=6, iter=4, b1=1.48, b2=0.29, scl_factor=0.75,
              solver=2, fst_grad=True, snd_term="gradmag", max_scales=None, oob=np.nan):
    """U = DispEminND_llin_2D(Il, Ir, fstTerm, sndTerm) as runme.m:20 calls it ('grad', 'gradmag'): late-linearisation
    stereo with ONE disparity field, DdiffWeights and Disp_sor_llin4_2d. Il, Ir: rows x cols x channels, 0..255.
    matlab/disparity/DispEminND_llin_2D.m, line numbers in the comments (spatial a-priori inputs omitted)."""
    I0 = (np.asarray(Il, dtype=F32).reshape(Il.shape[0], Il.shape[1], -1) / F32(255)).astype(F32)      # :74-75
    I1 = (np.asarray(Ir, dtype=F32).reshape(Ir.shape[0], Ir.shape[1], -1) / F32(255)).astype(F32)
    G = ms.fspecial_gaussian(5, 1.25)                                                                   # :98
    It0, It1 = [I0], [I1]
    scales = max_scales or (1 << 30)
    while len(It0) < scales:                                                                            # :106-125
        n0, n1 = ms.imresize_bilinear(It0[-1], scale=scl_factor), ms.imresize_bilinear(It1[-1], scale=scl_factor)
        It0[-1], It1[-1] = ms.imfilter(It0[-1], G), ms.imfilter(It1[-1], G)
        It0.append(n0); It1.append(n1)
        if n0.shape[0] <= 10 or n0.shape[1] <= 10:
            It0[-1], It1[-1] = ms.imfilter(It0[-1], G), ms.imfilter(It1[-1], G)
            break
    S = len(It0)
    F0 = [ms.rgb2grad(x) if fst_grad else x for x in It0]                                               # :130-143
    F1 = [ms.rgb2grad(x) if fst_grad else x for x in It1]
    U = None
    for s in range(S - 1, -1, -1):                                                                      # :184
        rows, cols = It0[s].shape[:2]
        Xg, Yg = np.meshgrid(np.arange(1, cols + 1, dtype=F32), np.arange(1, rows + 1, dtype=F32))
        if U is None:
            U = np.zeros((rows, cols), F32)
        for _ in range(firstLoop):                                                                      # :207
            X = (Xg + U).astype(F32)
            W1 = backend.bilin(F1[s], X, Yg, oob)                                                       # :212
            d1 = backend.call("FstDerivatives5", [F0[s], W1], 3)                                        # :224
            d2 = None
            if snd_term != "none":
                W2 = backend.bilin(It1[s], X, Yg, oob)                                                  # :218
                d2 = backend.call("SndDerivatives5", [It0[s], W2], 5) if snd_term == "gradmag" else \
                    backend.call("FstDerivatives5", [It0[s], W2], 3)                                    # :233,237
            dU = np.zeros((rows, cols), F32)
            for _ in range(secondLoop):                                                                 # :254
                CuGd, DuGd = ms.disp_terms(d1, d2, dU, b1, b2, alpha, snd_term == "gradmag")            # :259-280
                w = backend.call("DdiffWeights", [(U + dU).astype(F32), F32(0.00001)], 4)               # :277  [wW wN wE wS]
                dU = backend.call("Disp_sor_llin4_2d", [U, dU, CuGd, DuGd] + list(w) + [F32(iter), F32(omega), F32(solver)], 1)[0]   # :285-296
            U = ms.medfilt2_symmetric((U + dU).astype(F32))                                             # :303
        if s > 0:                                                                                       # :313-315
            U = ms.imresize_bilinear((U * F32(1.0 / scl_factor)).astype(F32), output_size=It0[s - 1].shape[:2])
    return U


def flow_ad(I0, I1, backend, quantile=0.9, diffusion="image", alpha=0.0420, omega=1.9, firstLoop=4, secondLoop=4, iter=4,
            b1=1.4843, b2=0.2915, scl_factor=0.75, solver=2, fst_grad=True, snd_term="gradmag", max_scales=None, oob=np.nan):
    """[U V] = FlowEminAD_llin_2D_v10(cat(3, I0, I1), channels, fstTerm, sndTerm, 'diffusion', diffusion) as runme.m:54,64
    calls it: the late-linearisation flow driver with anisotropic (Nagel-Enkelmann type) diffusion -- 8 edge weights from
    its own ADdiffWeights (:416-488), computed once per level from the first frame ('image') or per inner iteration from
    U+dU+V+dV ('flow'), and Oflow_sor_llin8_2d. matlab/optical_flow/FlowEminAD_llin_2D_v10.m, line numbers in the comments."""
    I0 = (np.asarray(I0, dtype=F32).reshape(I0.shape[0], I0.shape[1], -1) / F32(255)).astype(F32)      # :79
    I1 = (np.asarray(I1, dtype=F32).reshape(I1.shape[0], I1.shape[1], -1) / F32(255)).astype(F32)
    G = ms.fspecial_gaussian(5, 1.25)                                                                   # :102
    It0, It1 = [I0], [I1]
    scales = max_scales or (1 << 30)
    while len(It0) < scales:                                                                            # :110-131
        n0, n1 = ms.imresize_bilinear(It0[-1], scale=scl_factor), ms.imresize_bilinear(It1[-1], scale=scl_factor)
        It0[-1], It1[-1] = ms.imfilter(It0[-1], G), ms.imfilter(It1[-1], G)
        It0.append(n0); It1.append(n1)
        if n0.shape[0] <= 20 or n0.shape[1] <= 20:
            It0[-1], It1[-1] = ms.imfilter(It0[-1], G), ms.imfilter(It1[-1], G)
            break
    S = len(It0)
    F0 = [ms.rgb2grad(x) if fst_grad else x for x in It0]
    F1 = [ms.rgb2grad(x) if fst_grad else x for x in It1]
    w8 = lambda D: [np.asarray(x, dtype=F32) for x in ms.ad_diff_weights(D, quantile, flow_variant=True)[:8]]   # W NW N NE E SE S SW
    U = V = None
    for s in range(S - 1, -1, -1):                                                                      # :198
        rows, cols = It0[s].shape[:2]
        Xg, Yg = np.meshgrid(np.arange(1, cols + 1, dtype=F32), np.arange(1, rows + 1, dtype=F32))
        if diffusion == "image":
            w = w8(It0[s])                                                                              # :214
        if U is None:
            U = np.zeros((rows, cols), F32); V = np.zeros((rows, cols), F32)
        for _ in range(firstLoop):                                                                      # :233
            X, Y = (Xg + U).astype(F32), (Yg + V).astype(F32)
            W1 = backend.bilin(F1[s], X, Y, oob)                                                        # :238
            d1 = backend.call("FstDerivatives5", [F0[s], W1], 3)
            d2 = None
            if snd_term != "none":
                W2 = backend.bilin(It1[s], X, Y, oob)
                d2 = backend.call("SndDerivatives5", [It0[s], W2], 5) if snd_term == "gradmag" else \
                    backend.call("FstDerivatives5", [It0[s], W2], 3)
            dU = np.zeros((rows, cols), F32); dV = np.zeros((rows, cols), F32)
            for _ in range(secondLoop):                                                                 # :295
                if diffusion == "flow":
                    w = w8((((U + dU).astype(F32) + V).astype(F32) + dV).astype(F32))                   # :343
                M, Cu, Cv, Du, Dv = ms.llin_terms(d1, d2, dU, dV, b1, b2, alpha, snd_term == "gradmag")  # :305-352 (same formulas)
                dU, dV = backend.call("Oflow_sor_llin8_2d", [U, V, dU, dV, M, Cu, Cv, Du, Dv] + w +
                                      [F32(iter), F32(omega), F32(solver)], 2)                          # :357-377
            U = ms.medfilt2_symmetric((U + dU).astype(F32))                                             # :383-384
            V = ms.medfilt2_symmetric((V + dV).astype(F32))
        if s > 0:                                                                                       # :393-396
            up = F32(1.0 / scl_factor)
            size = It0[s - 1].shape[:2]
            U = ms.imresize_bilinear((U * up).astype(F32), output_size=size)
            V = ms.imresize_bilinear((V * up).astype(F32), output_size=size)
    return U, V
