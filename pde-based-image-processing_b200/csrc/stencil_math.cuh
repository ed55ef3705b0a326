// stencil_math.cuh -- per-pixel arithmetic of the relaxation sweeps, shared by the simple and the
// streaming kernels. One function per "what the reference computes at a pixel":
//
//   point_update<FAM>()  one SOR point update of all unknowns of the family at pixel (i,j)
//                        (GS_SOR_elin4_2d opticalflowSolvers.c:89-152, GS_SOR_llin4_2d :563-647,
//                         disparity GS_SOR_llin4_2d disparitySolvers.c:89-118, GS_SOR_4_2d
//                         pdeSolvers.c:94-118, GS_SOR_8_2d :208-240). Both unknowns of a flow
//                         family are computed from the OLD pair before either is stored
//                         (Jacobi inside the point, as the reference does).
//   line_eq<FAM>()       the tridiagonal row (a,b,c,d) of unknown q at pixel (i,j) for a line
//                        solve along `dir` with the reference's one-sided border rows
//                        (e.g. middleColumn_llin4 opticalflowSolvers.c:2548-2639, westColumn_elin4
//                         :1790-1792, middleColumn4 disparitySolvers.c:1500-1587, TDMA_mcolumn_ALR_4
//                         pdeSolvers.c:530-650, TDMAcolumn_ALR_8 :1132-1266).
//
// Sweeps are only required to reach the reference's fixed point (the parallel ordering already
// changes the iterates), so FMA contraction is allowed here.
#pragma once
#include "pdegpu_internal.cuh"

// 1/x to 1 ulp (MUFU.RCP). A relaxation step is a contraction towards the reference's fixed point, so a 1-ulp
// reciprocal changes nothing that can be observed at convergence; an IEEE division costs ~10 instructions.
__device__ __forceinline__ float sweep_rcp(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ---------------------------------------------------------------------------------------------
// point update (interior pixels only: all neighbours exist)
// ---------------------------------------------------------------------------------------------
// One SOR point update from VALUES (4-neighbour stencils): w = {wW, wN, wE, wS}, xn[q][n] the unknown q at
// neighbour n (same order), xc[q] at the pixel, x0n/x0c the fixed fields of the late-linearisation families.
// Shared by the global-memory kernel (point_update) and the tiled kernel (sweeps_point.cu) so that both
// evaluate the same expressions.
template <int FAM>
__device__ __forceinline__ void point_formula(const float (&w)[4], const float (&xn)[2][4], const float (&xc)[2],
                                              const float (&x0n)[2][4], const float (&x0c)[2],
                                              const float (&C)[2], const float (&D)[2], float M, float omega, float (&out)[2])
{
    using F = Fam<FAM>;
    const float wW = w[W_W], wN = w[W_N], wE = w[W_E], wS = w[W_S];
    const float sw = (wW + wE) + (wN + wS);
    if (F::PDE) {
        const float nb = xn[0][W_E] * wE + xn[0][W_W] * wW + xn[0][W_S] * wS + xn[0][W_N] * wN;
        const float tr = D[0];
        float inv, bt;
        if (!is_nan(tr)) { inv = sweep_rcp(tr); bt = C[0]; }
        else             { inv = sweep_rcp(sw); bt = 0.0f; }
        out[0] = (1.0f - omega) * xc[0] + omega * (bt + nb) * inv;
        return;
    }
    float nb[2];
#pragma unroll
    for (int q = 0; q < F::NUNK; q++) {
        if (F::LATE) {
            const float c0 = x0c[q];
            nb[q] = (xn[q][W_W] + x0n[q][W_W] - c0) * wW + (xn[q][W_E] + x0n[q][W_E] - c0) * wE
                  + (xn[q][W_N] + x0n[q][W_N] - c0) * wN + (xn[q][W_S] + x0n[q][W_S] - c0) * wS;
        } else {
            nb[q] = xn[q][W_W] * wW + xn[q][W_E] * wE + xn[q][W_N] * wN + xn[q][W_S] * wS;
        }
    }
#pragma unroll
    for (int q = 0; q < F::NUNK; q++) {
        float inv, val;
        if (F::NUNK == 2) {
            // flow: divisor tests isnan(D), right-hand side tests isnan(C) (opticalflowSolvers.c:118-149)
            inv = sweep_rcp(is_nan(D[q]) ? sw : sw + D[q]);
            val = is_nan(C[q]) ? nb[q] : (nb[q] + C[q] - M * xc[1 - q]);
        } else {
            // disparity: both test isnan(Cu) (disparitySolvers.c:96-112)
            const bool t = !is_nan(C[q]);
            inv = sweep_rcp(t ? (D[q] + sw) : sw);
            val = t ? (nb[q] + C[q]) : nb[q];
        }
        out[q] = (1.0f - omega) * xc[q] + omega * (val * inv);
    }
}

template <int FAM>
__device__ __forceinline__ void point_update(const SysView &s, long long pos, float omega)
{
    using F = Fam<FAM>;
    const int nr = s.nrows;
    if (F::PDE && F::EIGHT) {
        const float wW = s.w[W_W][pos], wN = s.w[W_N][pos], wE = s.w[W_E][pos], wS = s.w[W_S][pos];
        float sw = (wW + wE) + (wN + wS);
        float *X = s.x[0];
        float nb = X[pos + nr] * wE + X[pos - nr] * wW + X[pos + 1] * wS + X[pos - 1] * wN;
        const float wNW = s.w[W_NW][pos], wNE = s.w[W_NE][pos], wSE = s.w[W_SE][pos], wSW = s.w[W_SW][pos];
        nb += X[pos - nr + 1] * wSW + X[pos - nr - 1] * wNW + X[pos + nr + 1] * wSE + X[pos + nr - 1] * wNE;
        sw += (wSW + wNW) + (wSE + wNE);
        const float tr = s.d[0][pos];
        float inv, bt;
        if (!is_nan(tr)) { inv = 1.0f / tr; bt = s.c[0][pos]; }
        else             { inv = 1.0f / sw; bt = 0.0f; }
        X[pos] = (1.0f - omega) * X[pos] + omega * (bt + nb) * inv;
        return;
    }
    // flow / disparity / 4-neighbour PDE; the 8-neighbour flow point solver ignores its diagonal weights
    // (GS_SOR_llin8_2d opticalflowSolvers.c:1550-1598) => identical to llin4.
    const long long off[4] = {-(long long)nr, -1, (long long)nr, 1};
    float w[4], xn[2][4], xc[2], x0n[2][4], x0c[2], C[2], D[2], out[2];
#pragma unroll
    for (int n = 0; n < 4; n++) w[n] = s.w[n][pos];
#pragma unroll
    for (int q = 0; q < F::NUNK; q++) {
        xc[q] = s.x[q][pos];
        C[q] = s.c[q][pos]; D[q] = s.d[q][pos];
        if (F::LATE) x0c[q] = s.x0[q][pos];
#pragma unroll
        for (int n = 0; n < 4; n++) {
            xn[q][n] = s.x[q][pos + off[n]];
            if (F::LATE) x0n[q][n] = s.x0[q][pos + off[n]];
        }
    }
    const float M = F::NUNK == 2 ? s.m[pos] : 0.0f;
    point_formula<FAM>(w, xn, xc, x0n, x0c, C, D, M, omega, out);
#pragma unroll
    for (int q = 0; q < F::NUNK; q++) s.x[q][pos] = out[q];
}

// ---------------------------------------------------------------------------------------------
// tridiagonal row of unknown q at (i,j) for a line along dir (0: along i = Matlab column,
// 1: along j = Matlab row). `pos` already includes the batch offset.
// ---------------------------------------------------------------------------------------------
template <int FAM, int DIR>
__device__ __forceinline__ void line_eq(const SysView &s, long long pos, int q, int i, int j,
                                        float &a, float &b, float &c, float &d)
{
    using F = Fam<FAM>;
    const int nr = s.nrows, nc = s.ncols;
    const bool eN = i > 0, eS = i < nr - 1, eW = j > 0, eE = j < nc - 1;
    const bool ex[8] = {eW, eN, eE, eS, eN && eW, eN && eE, eS && eE, eS && eW};
    const long long off[8] = {-(long long)nr, -1, (long long)nr, 1, -(long long)nr - 1, (long long)nr - 1, (long long)nr + 1, -(long long)nr + 1};
    constexpr int NN = F::EIGHT ? 8 : 4;
    constexpr int prev = DIR == 0 ? W_N : W_W, next = DIR == 0 ? W_S : W_E;
    const float *X = s.x[q];
    float w[NN];
#pragma unroll
    for (int n = 0; n < NN; n++) w[n] = s.w[n][pos];

    a = ex[prev] ? -w[prev] : 0.0f;
    c = ex[next] ? -w[next] : 0.0f;
    float dsum = 0.0f, bsum = 0.0f;
#pragma unroll
    for (int n = 0; n < NN; n++) {
        if (!ex[n]) continue;
        bsum += w[n];
        const bool inline_nb = (n == prev) || (n == next);
        if (F::LATE) {
            const float *X0 = s.x0[q];
            float t = X0[pos + off[n]] - X0[pos];
            if (!inline_nb) t += X[pos + off[n]];
            dsum += w[n] * t;
        } else if (!inline_nb) {
            dsum += w[n] * X[pos + off[n]];
        }
    }
    if (F::PDE) {
        const float tr = s.d[0][pos];
        if (!is_nan(tr)) { b = tr; d = dsum + s.c[0][pos]; }
        else {
            if (F::EIGHT) {
                // reference's NaN-TRACE diagonal: all 4 axial weights + wNW twice + wSW + wSE, whatever the
                // position (pdeSolvers.c:1179,1209,1238)
                b = (s.w[W_N][pos] + s.w[W_S][pos] + s.w[W_W][pos] + s.w[W_E][pos])
                  + (s.w[W_NW][pos] + s.w[W_NW][pos] + s.w[W_SW][pos] + s.w[W_SE][pos]);
            } else b = bsum;
            d = dsum;
        }
        return;
    }
    const float C = s.c[q][pos];
    b = bsum;
    d = dsum;
    if (!is_nan(C)) {
        b += s.d[q][pos];
        d += C;
        if (F::NUNK == 2) d -= s.m[pos] * s.x[1 - q][pos];
    }
}
