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

// ---------------------------------------------------------------------------------------------
// point update (interior pixels only: all neighbours exist)
// ---------------------------------------------------------------------------------------------
template <int FAM>
__device__ __forceinline__ void point_update(const SysView &s, long long pos, float omega)
{
    using F = Fam<FAM>;
    const int nr = s.nrows;
    const float wW = s.w[W_W][pos], wN = s.w[W_N][pos], wE = s.w[W_E][pos], wS = s.w[W_S][pos];
    float sw = (wW + wE) + (wN + wS);
    if (F::PDE) {
        float *X = s.x[0];
        float nb = X[pos + nr] * wE + X[pos - nr] * wW + X[pos + 1] * wS + X[pos - 1] * wN;
        if (F::EIGHT) {
            const float wNW = s.w[W_NW][pos], wNE = s.w[W_NE][pos], wSE = s.w[W_SE][pos], wSW = s.w[W_SW][pos];
            nb += X[pos - nr + 1] * wSW + X[pos - nr - 1] * wNW + X[pos + nr + 1] * wSE + X[pos + nr - 1] * wNE;
            sw += (wSW + wNW) + (wSE + wNE);
        }
        const float tr = s.d[0][pos];
        float inv, bt;
        if (!is_nan(tr)) { inv = 1.0f / tr; bt = s.c[0][pos]; }
        else             { inv = 1.0f / sw; bt = 0.0f; }
        X[pos] = (1.0f - omega) * X[pos] + omega * (bt + nb) * inv;
        return;
    }
    // flow / disparity families; the 8-neighbour flow point solver ignores its diagonal weights
    // (GS_SOR_llin8_2d opticalflowSolvers.c:1550-1598) => identical to llin4.
    float nb[2], xc[2];
#pragma unroll
    for (int q = 0; q < F::NUNK; q++) {
        const float *X = s.x[q];
        xc[q] = X[pos];
        if (F::LATE) {
            const float *X0 = s.x0[q];
            const float c0 = X0[pos];
            nb[q] = (X[pos - nr] + X0[pos - nr] - c0) * wW + (X[pos + nr] + X0[pos + nr] - c0) * wE
                  + (X[pos - 1] + X0[pos - 1] - c0) * wN + (X[pos + 1] + X0[pos + 1] - c0) * wS;
        } else {
            nb[q] = X[pos - nr] * wW + X[pos + nr] * wE + X[pos - 1] * wN + X[pos + 1] * wS;
        }
    }
#pragma unroll
    for (int q = 0; q < F::NUNK; q++) {
        const float C = s.c[q][pos], D = s.d[q][pos];
        float inv, val;
        if (F::NUNK == 2) {
            // flow: divisor tests isnan(D), right-hand side tests isnan(C) (opticalflowSolvers.c:118-149)
            inv = 1.0f / (is_nan(D) ? sw : sw + D);
            val = is_nan(C) ? nb[q] : (nb[q] + C - s.m[pos] * xc[1 - q]);
        } else {
            // disparity: both test isnan(Cu) (disparitySolvers.c:96-112)
            const bool t = !is_nan(C);
            inv = 1.0f / (t ? (D + sw) : sw);
            val = t ? (nb[q] + C) : nb[q];
        }
        s.x[q][pos] = (1.0f - omega) * xc[q] + omega * (val * inv);
    }
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
