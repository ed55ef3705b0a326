/*
 * pde_oracle.c -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY
 * (see pde_oracle.h). Never linked into libpdegpu; libpdegpu has no CPU path.
 *
 * All arrays are fp32, column-major: element (i,j,k) at k*nrows*ncols + j*nrows + i.
 * Neighbour names follow the reference: W/E = column j-1/j+1, N/S = row i-1/i+1.
 * Every function cites the reference lines (under /root/reference/mex/source/library/) it restates.
 */
#include "pde_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ISNAN(x) ((x) != (x))
#define P(i, j) ((long)(j) * nrows + (i))

/* reference border fill: rows first, then columns (opticalflowSolvers.c:161-179) */
static void fill_border(float *X, int nrows, int ncols)
{
    int i, j;
    for (j = 0; j < ncols; j++) {
        X[P(0, j)] = X[P(1, j)];
        X[P(nrows - 1, j)] = X[P(nrows - 2, j)];
    }
    for (i = 0; i < nrows; i++) {
        X[P(i, 0)] = X[P(i, 1)];
        X[P(i, ncols - 1)] = X[P(i, ncols - 2)];
    }
}

/* ============================================================================================
 * Point-wise lexicographic Gauss-Seidel SOR, flow families
 *   late=0: GS_SOR_elin4_2d  opticalflowSolvers.c:41-186
 *   late=1: GS_SOR_llin4_2d  opticalflowSolvers.c:504-680  (GS_SOR_llin8_2d :1487 is the same
 *           arithmetic: it never reads its diagonal weights)
 * ============================================================================================ */
void orc_flow_gs(int late, float *X0, float *X1, const float *F0, const float *F1,
                 const float *M, const float *Cu, const float *Cv, const float *Du, const float *Dv,
                 const float *wW, const float *wN, const float *wE, const float *wS,
                 int nrows, int ncols, int iter, float omega)
{
    int it, i, j;
    for (it = 0; it < iter; it++) {
        for (j = 1; j < ncols - 1; j++) {
            for (i = 1; i < nrows - 1; i++) {
                const long p = P(i, j), w = p - nrows, e = p + nrows, n = p - 1, s = p + 1;
                float nu, nv, t1, t2, t3, sw, divu, divv, unew, vnew;
                if (!late) {
                    nu = X0[w] * wW[p]; t1 = X0[e] * wE[p]; nu += t1;
                    t2 = X0[n] * wN[p]; t3 = X0[s] * wS[p]; t2 += t3; nu += t2;
                    nv = X1[w] * wW[p]; t1 = X1[e] * wE[p]; nv += t1;
                    t2 = X1[n] * wN[p]; t3 = X1[s] * wS[p]; t2 += t3; nv += t2;
                } else {
                    nu = X0[w] + F0[w]; t1 = X0[e] + F0[e]; t2 = X0[n] + F0[n]; t3 = X0[s] + F0[s];
                    nu -= F0[p]; t1 -= F0[p]; t2 -= F0[p]; t3 -= F0[p];
                    nu *= wW[p]; t1 *= wE[p]; t2 *= wN[p]; t3 *= wS[p];
                    nu += t1; t2 += t3; nu += t2;
                    nv = X1[w] + F1[w]; t1 = X1[e] + F1[e]; t2 = X1[n] + F1[n]; t3 = X1[s] + F1[s];
                    nv -= F1[p]; t1 -= F1[p]; t2 -= F1[p]; t3 -= F1[p];
                    nv *= wW[p]; t1 *= wE[p]; t2 *= wN[p]; t3 *= wS[p];
                    nv += t1; t2 += t3; nv += t2;
                }
                sw = wW[p] + wE[p]; t2 = wN[p] + wS[p]; sw += t2;
                divu = ISNAN(Du[p]) ? 1.0f / sw : 1.0f / (sw + Du[p]);
                divv = ISNAN(Dv[p]) ? 1.0f / sw : 1.0f / (sw + Dv[p]);
                if (ISNAN(Cu[p])) unew = nu * divu;
                else { t1 = nu + Cu[p]; t2 = M[p] * X1[p]; t1 = t1 - t2; unew = t1 * divu; }
                if (ISNAN(Cv[p])) vnew = nv * divv;
                else { t1 = nv + Cv[p]; t2 = M[p] * X0[p]; t1 = t1 - t2; vnew = t1 * divv; }
                X0[p] = (1.0f - omega) * X0[p] + omega * unew;
                X1[p] = (1.0f - omega) * X1[p] + omega * vnew;
            }
        }
        fill_border(X0, nrows, ncols);
        fill_border(X1, nrows, ncols);
    }
}

/* ============================================================================================
 * Thomas solve of one assembled line + SOR, exactly as the reference's line solvers do it
 * (e.g. middleColumn_elin4 opticalflowSolvers.c:1910-1970): forward elimination with
 * div = 1/(b - cp*a), true division on the first and last rows, back substitution through the
 * un-relaxed value, relaxation one step behind.
 * a,b,c,d: assembled rows; x: line in place with element stride `st`.
 * ============================================================================================ */
static void thomas_sor(float *x, long st, int n, const float *a, const float *b, const float *c, const float *d,
                       float *cp, float *dp, float omega, int true_div)
{
    int k;
    float div, temp1, temp2;
    cp[0] = c[0] / b[0];
    dp[0] = d[0] / b[0];
    for (k = 1; k <= n - 2; k++) {
        if (true_div) {
            /* southRow_llin4 (opticalflowSolvers.c:3059-3060), southRow_llin8 (:3871), southRow4
             * (disparitySolvers.c:1986-1987) divide twice instead of multiplying by a reciprocal */
            cp[k] = c[k] / (b[k] - cp[k - 1] * a[k]);
            dp[k] = (d[k] - dp[k - 1] * a[k]) / (b[k] - cp[k - 1] * a[k]);
            continue;
        }
        div = 1.0f / (b[k] - cp[k - 1] * a[k]);
        cp[k] = c[k] * div;
        dp[k] = (d[k] - dp[k - 1] * a[k]) * div;
    }
    k = n - 1;
    dp[k] = (d[k] - dp[k - 1] * a[k]) / (b[k] - cp[k - 1] * a[k]);
    temp1 = x[k * st];
    x[k * st] = dp[k];
    for (k = n - 2; k >= 0; k--) {
        temp2 = x[k * st];
        x[k * st] = dp[k] - cp[k] * x[(k + 1) * st];
        x[(k + 1) * st] = omega * x[(k + 1) * st] + (1.0f - omega) * temp1;
        temp1 = temp2;
    }
    x[0] = omega * x[0] + (1.0f - omega) * temp1;
}

typedef struct {
    float *a, *b, *c, *d, *cp, *dp;
} linebuf;

static int linebuf_alloc(linebuf *L, int n)
{
    L->a = (float *)calloc((size_t)6 * n, sizeof(float));
    if (!L->a) return 0;
    L->b = L->a + n; L->c = L->b + n; L->d = L->c + n; L->cp = L->d + n; L->dp = L->cp + n;
    return 1;
}

/* add a term to an accumulator whose first term ASSIGNS (the reference writes b = w1 + w2 + ...) */
#define ACC(acc, have, v) do { if (have) (acc) += (v); else { (acc) = (v); (have) = 1; } } while (0)

/* One line of a 4-neighbour flow / disparity system.
 *   dir 0: line along i at column j (west/middle/eastColumn_*), dir 1: along j at row i (north/middle/southRow_*)
 *   X: unknown being relaxed, Y: coupled unknown (NULL for disparity), F: fixed field (NULL for early lin.)
 * b sums the existing weights in the order N,S,E,W; d sums its terms in the order W,E,S,N
 * (opticalflowSolvers.c:1899-1901, 2138-2140, 2586-2591; disparitySolvers.c:1531-1536). */
static void flow_line(int dir, int line, float *X, const float *Y, const float *F,
                      const float *M, const float *C, const float *D,
                      const float *wW, const float *wN, const float *wE, const float *wS,
                      int nrows, int ncols, float omega, linebuf *L)
{
    const int n = dir == 0 ? nrows : ncols;
    int k;
    for (k = 0; k < n; k++) {
        const int i = dir == 0 ? k : line, j = dir == 0 ? line : k;
        const long p = P(i, j), w = p - nrows, e = p + nrows, nn = p - 1, s = p + 1;
        const int eN = i > 0, eS = i < nrows - 1, eW = j > 0, eE = j < ncols - 1;
        float b = 0.0f, d = 0.0f;
        int hb = 0, hd = 0;
        if (eN) ACC(b, hb, wN[p]);
        if (eS) ACC(b, hb, wS[p]);
        if (eE) ACC(b, hb, wE[p]);
        if (eW) ACC(b, hb, wW[p]);
        if (dir == 0) { L->a[k] = eN ? -wN[p] : 0.0f; L->c[k] = eS ? -wS[p] : 0.0f; }
        else          { L->a[k] = eW ? -wW[p] : 0.0f; L->c[k] = eE ? -wE[p] : 0.0f; }
        if (!F) {
            if (dir == 0) { if (eW) ACC(d, hd, wW[p] * X[w]); if (eE) ACC(d, hd, wE[p] * X[e]); }
            else          { if (eS) ACC(d, hd, wS[p] * X[s]); if (eN) ACC(d, hd, wN[p] * X[nn]); }
        } else {
            if (dir == 0) {
                if (eW) ACC(d, hd, wW[p] * (F[w] - F[p] + X[w]));
                if (eE) ACC(d, hd, wE[p] * (F[e] - F[p] + X[e]));
                if (eS) ACC(d, hd, wS[p] * (F[s] - F[p]));
                if (eN) ACC(d, hd, wN[p] * (F[nn] - F[p]));
            } else {
                if (eW) ACC(d, hd, wW[p] * (F[w] - F[p]));
                if (eE) ACC(d, hd, wE[p] * (F[e] - F[p]));
                if (eS) ACC(d, hd, wS[p] * (F[s] - F[p] + X[s]));
                if (eN) ACC(d, hd, wN[p] * (F[nn] - F[p] + X[nn]));
            }
        }
        if (!ISNAN(C[p])) {
            b += D[p];
            d += C[p];
            if (Y) d -= M[p] * Y[p];
        }
        L->b[k] = b; L->d[k] = d;
    }
    thomas_sor(X + (dir == 0 ? P(0, line) : P(line, 0)), dir == 0 ? 1 : nrows, n, L->a, L->b, L->c, L->d, L->cp, L->dp, omega,
               F != NULL && dir == 1 && line == nrows - 1);
}

/* GS_ALR_SOR_elin4_2d opticalflowSolvers.c:196-262 (late=0), GS_ALR_SOR_llin4_2d :690-759 (late=1):
 * per iteration: columns for unknown 0, columns for unknown 1, rows for unknown 1, rows for unknown 0,
 * lines visited in increasing order (lexicographic block Gauss-Seidel). */
void orc_flow_alr(int late, float *X0, float *X1, const float *F0, const float *F1,
                  const float *M, const float *Cu, const float *Cv, const float *Du, const float *Dv,
                  const float *wW, const float *wN, const float *wE, const float *wS,
                  int nrows, int ncols, int iter, float omega)
{
    linebuf L;
    int it, i, j;
    if (!linebuf_alloc(&L, nrows > ncols ? nrows : ncols)) return;
    if (!late) { F0 = NULL; F1 = NULL; }
    for (it = 0; it < iter; it++) {
        for (j = 0; j < ncols; j++) flow_line(0, j, X0, X1, F0, M, Cu, Du, wW, wN, wE, wS, nrows, ncols, omega, &L);
        for (j = 0; j < ncols; j++) flow_line(0, j, X1, X0, F1, M, Cv, Dv, wW, wN, wE, wS, nrows, ncols, omega, &L);
        for (i = 0; i < nrows; i++) flow_line(1, i, X1, X0, F1, M, Cv, Dv, wW, wN, wE, wS, nrows, ncols, omega, &L);
        for (i = 0; i < nrows; i++) flow_line(1, i, X0, X1, F0, M, Cu, Du, wW, wN, wE, wS, nrows, ncols, omega, &L);
    }
    free(L.a);
}

/* GS_ALR_SOR_llin4_2d disparitySolvers.c:154-211 with westColumn4..southRow4 :1376-2029 */
void orc_disp_alr(float *dU, const float *U, const float *Cu, const float *Du,
                  const float *wW, const float *wN, const float *wE, const float *wS,
                  int nrows, int ncols, int iter, float omega)
{
    linebuf L;
    int it, i, j;
    if (!linebuf_alloc(&L, nrows > ncols ? nrows : ncols)) return;
    for (it = 0; it < iter; it++) {
        for (j = 0; j < ncols; j++) flow_line(0, j, dU, NULL, U, NULL, Cu, Du, wW, wN, wE, wS, nrows, ncols, omega, &L);
        for (i = 0; i < nrows; i++) flow_line(1, i, dU, NULL, U, NULL, Cu, Du, wW, wN, wE, wS, nrows, ncols, omega, &L);
    }
    free(L.a);
}

/* GS_ALR_SOR_llin8_2d opticalflowSolvers.c:1677-1750 with *_llin8 line solvers :3104-3914.
 * Terms are summed in the fixed order W,N,E,S,NW,NE,SE,SW (the reference's order varies from
 * border case to border case), so agreement with the reference is to rounding, not bit-exact. */
static void flow8_line(int dir, int line, float *X, const float *Y, const float *F,
                       const float *M, const float *C, const float *D, const float *const w8[8],
                       int nrows, int ncols, float omega, linebuf *L)
{
    const int n = dir == 0 ? nrows : ncols;
    const int di[8] = {0, -1, 0, 1, -1, -1, 1, 1}, dj[8] = {-1, 0, 1, 0, -1, 1, 1, -1};
    const int prev = dir == 0 ? 1 : 0, next = dir == 0 ? 3 : 2;
    int k, q;
    for (k = 0; k < n; k++) {
        const int i = dir == 0 ? k : line, j = dir == 0 ? line : k;
        const long p = P(i, j);
        float b = 0.0f, d = 0.0f;
        L->a[k] = 0.0f; L->c[k] = 0.0f;
        for (q = 0; q < 8; q++) {
            const int ii = i + di[q], jj = j + dj[q];
            float t;
            if (ii < 0 || ii >= nrows || jj < 0 || jj >= ncols) continue;
            b += w8[q][p];
            t = F[P(ii, jj)] - F[p];
            if (q == prev) L->a[k] = -w8[q][p];
            else if (q == next) L->c[k] = -w8[q][p];
            else t += X[P(ii, jj)];
            d += w8[q][p] * t;
        }
        if (!ISNAN(C[p])) { b += D[p]; d += C[p]; d -= M[p] * Y[p]; }
        L->b[k] = b; L->d[k] = d;
    }
    thomas_sor(X + (dir == 0 ? P(0, line) : P(line, 0)), dir == 0 ? 1 : nrows, n, L->a, L->b, L->c, L->d, L->cp, L->dp, omega,
               dir == 1 && line == nrows - 1);
}

void orc_flow_alr8(float *dU, float *dV, const float *U, const float *V,
                   const float *M, const float *Cu, const float *Cv, const float *Du, const float *Dv,
                   const float *const w8[8], int nrows, int ncols, int iter, float omega)
{
    linebuf L;
    int it, i, j;
    if (!linebuf_alloc(&L, nrows > ncols ? nrows : ncols)) return;
    for (it = 0; it < iter; it++) {
        for (j = 0; j < ncols; j++) flow8_line(0, j, dU, dV, U, M, Cu, Du, w8, nrows, ncols, omega, &L);
        for (j = 0; j < ncols; j++) flow8_line(0, j, dV, dU, V, M, Cv, Dv, w8, nrows, ncols, omega, &L);
        for (i = 0; i < nrows; i++) flow8_line(1, i, dV, dU, V, M, Cv, Dv, w8, nrows, ncols, omega, &L);
        for (i = 0; i < nrows; i++) flow8_line(1, i, dU, dV, U, M, Cu, Du, w8, nrows, ncols, omega, &L);
    }
    free(L.a);
}

/* ============================================================================================
 * Residuals_elin4_2d :269-380, LHS_elin4_2d :387-496, Residuals_llin4_2d :766-916,
 * LHS_llin4_2d :923-1070 (opticalflowSolvers.c)
 * ============================================================================================ */
void orc_flow_operator(int late, int lhs, int quirks, float *RU, float *RV,
                       const float *X0, const float *X1, const float *F0, const float *F1,
                       const float *M, const float *Cu, const float *Cv, const float *Du, const float *Dv,
                       const float *wW, const float *wN, const float *wE, const float *wS,
                       int nrows, int ncols, int nframes)
{
    const long fsz = (long)nrows * ncols;
    int i, j, k;
    long stale_pos = 0;
    memset(RU, 0, sizeof(float) * fsz * nframes);
    memset(RV, 0, sizeof(float) * fsz * nframes);
    for (k = 0; k < nframes; k++) {
        const long fo = k * fsz;
        for (j = 1; j < ncols - 1; j++) {
            for (i = 1; i < nrows - 1; i++) {
                const long p = P(i, j), q = p + fo, w = p - nrows, e = p + nrows, n = p - 1, s = p + 1;
                float nu, nv, t1, t2, t3, sw;
                stale_pos = p;
                if (!late) {
                    nu = X0[w] * wW[p] + X0[e] * wE[p] + X0[n] * wN[p] + X0[s] * wS[p];
                    nv = X1[w] * wW[p] + X1[e] * wE[p] + X1[n] * wN[p] + X1[s] * wS[p];
                } else {
                    nu = X0[w] + F0[w]; t1 = X0[e] + F0[e]; t2 = X0[n] + F0[n]; t3 = X0[s] + F0[s];
                    nu -= F0[p]; t1 -= F0[p]; t2 -= F0[p]; t3 -= F0[p];
                    nu *= wW[p]; t1 *= wE[p]; t2 *= wN[p]; t3 *= wS[p];
                    nu += t1; t2 += t3; nu += t2;
                    nv = X1[w] + F1[w]; t1 = X1[e] + F1[e]; t2 = X1[n] + F1[n]; t3 = X1[s] + F1[s];
                    nv -= F1[p]; t1 -= F1[p]; t2 -= F1[p]; t3 -= F1[p];
                    nv *= wW[p]; t1 *= wE[p]; t2 *= wN[p]; t3 *= wS[p];
                    nv += t1; t2 += t3; nv += t2;
                }
                sw = wW[p] + wE[p]; t2 = wN[p] + wS[p]; sw += t2;
                if (!lhs) {
                    if (!ISNAN(Cu[q])) RU[q] = Cu[q] - M[q] * X1[p] + nu - (Du[q] + sw) * X0[p];
                    else RU[q] = nu - sw * X0[p];
                    if (!ISNAN(Cv[q])) RV[q] = Cv[q] - M[q] * X0[p] + nv - (Dv[q] + sw) * X1[p];
                    else RV[q] = nv - sw * X1[p];
                } else {
                    if (!ISNAN(Du[q])) RU[q] = M[q] * X1[p] - nu + (Du[q] + sw) * X0[p];
                    else RU[q] = -nu + sw * X0[p];
                    if (!ISNAN(Dv[q])) RV[q] = M[q] * X0[p] - nv + (Dv[q] + sw) * X1[p];
                    else RV[q] = -nv + sw * X1[p];
                }
            }
        }
    }
    for (k = 0; k < nframes; k++) {
        const long fo = k * fsz;
        for (j = 0; j < ncols; j++) {
            const long q = P(0, j) + fo;
            RU[q] = RU[q + 1];
            RU[q + nrows - 1] = RU[q + nrows - 2];
            /* LHS_llin4_2d reads AV through a stale, frame-less `pos` (:1056) */
            RV[q] = (quirks && late && lhs) ? RV[stale_pos + 1] : RV[q + 1];
            RV[q + nrows - 1] = RV[q + nrows - 2];
        }
        for (i = 0; i < nrows; i++) {
            const long q = i + fo;
            RU[q] = RU[q + nrows];
            RU[q + (long)(ncols - 1) * nrows] = RU[q + (long)(ncols - 2) * nrows];
            /* Residuals_llin4_2d drops the frame offset on the source (:912) */
            RV[q] = (quirks && late && !lhs) ? RV[i + nrows] : RV[q + nrows];
            RV[q + (long)(ncols - 1) * nrows] = RV[q + (long)(ncols - 2) * nrows];
        }
    }
}

/* ============================================================================================
 * Disparity point solvers: GS_SOR_llin4_2d disparitySolvers.c:41-146, GS_SOR_llinsym4_2d :301-455
 * ============================================================================================ */
static float disp_neigh(const float *U, const float *dU, const float *wW, const float *wN, const float *wE, const float *wS,
                        long p, int nrows)
{
    const long w = p - nrows, e = p + nrows, n = p - 1, s = p + 1;
    return (U[e] + dU[e] - U[p]) * wE[p] + (U[w] + dU[w] - U[p]) * wW[p]
         + (U[s] + dU[s] - U[p]) * wS[p] + (U[n] + dU[n] - U[p]) * wN[p];
}

void orc_disp_gs(float *dU, const float *U, const float *Cu, const float *Du,
                 const float *wW, const float *wN, const float *wE, const float *wS,
                 int nrows, int ncols, int iter, float omega)
{
    int it, i, j;
    for (it = 0; it < iter; it++) {
        for (j = 1; j < ncols - 1; j++)
            for (i = 1; i < nrows - 1; i++) {
                const long p = P(i, j);
                const float wn = disp_neigh(U, dU, wW, wN, wE, wS, p, nrows);
                float dividend, div, A, B;
                if (!ISNAN(Cu[p])) { dividend = Cu[p]; div = 1.0f / (Du[p] + wE[p] + wW[p] + wS[p] + wN[p]); }
                else               { dividend = 0.0f;  div = 1.0f / (wE[p] + wW[p] + wS[p] + wN[p]); }
                A = (1.0f - omega) * dU[p];
                B = omega * (wn + dividend) * div;
                dU[p] = A + B;
            }
        fill_border(dU, nrows, ncols);
    }
}

void orc_disp_gs_sym(float *dU0, const float *U0, const float *Cu0, const float *Du0,
                     const float *wW0, const float *wN0, const float *wE0, const float *wS0,
                     float *dU1, const float *U1, const float *Cu1, const float *Du1,
                     const float *wW1, const float *wN1, const float *wE1, const float *wS1,
                     int nrows, int ncols, int iter, float omega)
{
    int it, i, j, f;
    float *dU[2]; const float *U[2], *Cu[2], *Du[2], *wW[2], *wN[2], *wE[2], *wS[2];
    dU[0] = dU0; dU[1] = dU1; U[0] = U0; U[1] = U1; Cu[0] = Cu0; Cu[1] = Cu1; Du[0] = Du0; Du[1] = Du1;
    wW[0] = wW0; wW[1] = wW1; wN[0] = wN0; wN[1] = wN1; wE[0] = wE0; wE[1] = wE1; wS[0] = wS0; wS[1] = wS1;
    for (it = 0; it < iter; it++) {
        for (j = 1; j < ncols - 1; j++)
            for (i = 1; i < nrows - 1; i++) {
                const long p = P(i, j);
                float wn[2];
                /* both neighbour sums first, then both updates (:373-383, :433-437); the fields are independent */
                for (f = 0; f < 2; f++) wn[f] = disp_neigh(U[f], dU[f], wW[f], wN[f], wE[f], wS[f], p, nrows);
                for (f = 0; f < 2; f++) {
                    float dividend = 0.0f, div, approx;
                    if (!ISNAN(Cu[f][p])) { dividend = Cu[f][p]; div = 1.0f / (Du[f][p] + wE[f][p] + wW[f][p] + wS[f][p] + wN[f][p]); }
                    else div = 1.0f / (wE[f][p] + wW[f][p] + wS[f][p] + wN[f][p]);
                    approx = (wn[f] + dividend) * div;
                    dU[f][p] = (1.0f - omega) * dU[f][p] + omega * approx;
                }
            }
        fill_border(dU0, nrows, ncols);
        fill_border(dU1, nrows, ncols);
    }
}

/* ============================================================================================
 * Generic PDE solvers: GS_SOR_4_2d pdeSolvers.c:44-146, GS_SOR_8_2d :153-268,
 * GS_ALR_SOR_4_2d :277-335 (+ TDMA_*_ALR_4 :409-1130), GS_ALR_SOR_8_2d :344-402
 * (+ TDMAcolumn_ALR_8 :1132-1266, TDMArow_ALR_8 :1267-1393)
 * w = {wW,wN,wE,wS,wNW,wNE,wSE,wSW}
 * ============================================================================================ */
void orc_pde_gs(int eight, float *X, const float *TRACE, const float *B, const float *const w[8],
                int nrows, int ncols, int nframes, int iter, float omega)
{
    const long fsz = (long)nrows * ncols;
    int it, i, j, k;
    for (it = 0; it < iter; it++)
        for (k = 0; k < nframes; k++) {
            for (j = 1; j < ncols - 1; j++)
                for (i = 1; i < nrows - 1; i++) {
                    const long p = P(i, j) + k * fsz, wp = p - nrows, ep = p + nrows, np = p - 1, sp = p + 1;
                    float wn, inv, bt;
                    wn = X[ep] * w[2][p] + X[wp] * w[0][p];
                    wn += X[sp] * w[3][p] + X[np] * w[1][p];
                    if (eight) {
                        wn += X[wp + 1] * w[7][p] + X[wp - 1] * w[4][p];
                        wn += X[ep + 1] * w[6][p] + X[ep - 1] * w[5][p];
                    }
                    if (!ISNAN(TRACE[p])) { inv = 1.0f / TRACE[p]; bt = B[p]; }
                    else {
                        inv = w[2][p] + w[0][p];
                        inv += w[3][p] + w[1][p];
                        if (eight) { inv += w[7][p] + w[4][p]; inv += w[6][p] + w[5][p]; }
                        inv = 1.0f / inv;
                        bt = 0.0f;
                    }
                    X[p] = (1.0f - omega) * X[p];
                    X[p] += omega * (bt + wn) * inv;
                }
            fill_border(X + k * fsz, nrows, ncols);
        }
}

static void pde_line(int eight, int dir, int line, float *X, const float *TRACE, const float *B, const float *const w[8],
                     int nrows, int ncols, float omega, linebuf *L)
{
    const int n = dir == 0 ? nrows : ncols;
    int k;
    for (k = 0; k < n; k++) {
        const int i = dir == 0 ? k : line, j = dir == 0 ? line : k;
        const long p = P(i, j), wp = p - nrows, ep = p + nrows, np = p - 1, sp = p + 1;
        const int eN = i > 0, eS = i < nrows - 1, eW = j > 0, eE = j < ncols - 1;
        float b = 0.0f, d = 0.0f;
        int hb = 0, hd = 0;
        if (dir == 0) {
            L->a[k] = eN ? -w[1][p] : 0.0f; L->c[k] = eS ? -w[3][p] : 0.0f;
            if (eW) ACC(d, hd, w[0][p] * X[wp]);
            if (eE) ACC(d, hd, w[2][p] * X[ep]);
            if (eight) {   /* interior columns only: W and E always exist (pdeSolvers.c:1155) */
                if (eS) d += w[7][p] * X[wp + 1] + w[6][p] * X[ep + 1];
                if (eN) d += w[4][p] * X[wp - 1] + w[5][p] * X[ep - 1];
            }
        } else {
            L->a[k] = eW ? -w[0][p] : 0.0f; L->c[k] = eE ? -w[2][p] : 0.0f;
            if (eS) ACC(d, hd, w[3][p] * X[sp]);
            if (eN) ACC(d, hd, w[1][p] * X[np]);
            if (eight) {   /* interior rows only: N and S always exist (pdeSolvers.c:1290) */
                if (eW) d += w[7][p] * X[sp - nrows] + w[4][p] * X[np - nrows];
                if (eE) d += w[6][p] * X[sp + nrows] + w[5][p] * X[np + nrows];
            }
        }
        if (!ISNAN(TRACE[p])) { b = TRACE[p]; d += B[p]; }
        else if (eight) {
            /* wNW twice, wNE never, whatever the position (pdeSolvers.c:1179) */
            b = w[1][p] + w[3][p] + w[0][p] + w[2][p];
            b += w[4][p] + w[4][p] + w[7][p] + w[6][p];
        } else {
            if (eN) ACC(b, hb, w[1][p]);
            if (eS) ACC(b, hb, w[3][p]);
            if (eW) ACC(b, hb, w[0][p]);
            if (eE) ACC(b, hb, w[2][p]);
        }
        L->b[k] = b; L->d[k] = d;
    }
    thomas_sor(X + (dir == 0 ? P(0, line) : P(line, 0)), dir == 0 ? 1 : nrows, n, L->a, L->b, L->c, L->d, L->cp, L->dp, omega, 0);
}

void orc_pde_alr(int eight, float *X, const float *TRACE, const float *B, const float *const w[8],
                 int nrows, int ncols, int nframes, int iter, float omega)
{
    const long fsz = (long)nrows * ncols;
    linebuf L;
    int it, i, j, k, q;
    const float *wk[8];
    if (!linebuf_alloc(&L, nrows > ncols ? nrows : ncols)) return;
    if (eight) iter = 1;                                   /* pdeSolvers.c:362 */
    for (it = 0; it < iter; it++) {
        /* each pass loops over the frames itself (pdeSolvers.c:547, :1145) */
        for (k = 0; k < nframes; k++) {
            for (q = 0; q < (eight ? 8 : 4); q++) wk[q] = w[q] + k * fsz;
            for (j = eight ? 1 : 0; j <= (eight ? ncols - 2 : ncols - 1); j++)
                pde_line(eight, 0, j, X + k * fsz, TRACE + k * fsz, B + k * fsz, wk, nrows, ncols, omega, &L);
        }
        for (k = 0; k < nframes; k++) {
            for (q = 0; q < (eight ? 8 : 4); q++) wk[q] = w[q] + k * fsz;
            for (i = eight ? 1 : 0; i <= (eight ? nrows - 2 : nrows - 1); i++)
                pde_line(eight, 1, i, X + k * fsz, TRACE + k * fsz, B + k * fsz, wk, nrows, ncols, omega, &L);
        }
    }
    free(L.a);
}

/* ============================================================================================
 * bilinInterp2 imageInterpolation.c:44-140. (unsigned int)floor(v) is evaluated the way gcc/x86-64
 * does (cvttsd2si to 64 bits, low half kept; NaN/overflow -> 0x8000000000000000 -> 0), SURVEY Q3.
 * ============================================================================================ */
static unsigned int floor_to_uint(float v)
{
    const double f = floor((double)v);
    long long q;
    if (!(fabs(f) < 9.2233720368547758e18)) q = (long long)0x8000000000000000ULL;
    else q = (long long)f;
    return (unsigned int)(unsigned long long)q;
}

void orc_bilin(float *Iout, const float *Iin, const float *X, const float *Y,
               int nrows, int ncols, int nframes, float oob)
{
    const long fsz = (long)nrows * ncols;
    int i, j, k;
    for (j = 0; j < ncols; j++)
        for (i = 0; i < nrows; i++) {
            const long pos = P(i, j);
            const unsigned int x = floor_to_uint(X[pos] - 1.0f), y = floor_to_uint(Y[pos] - 1.0f);
            if (x < (unsigned)ncols && y < (unsigned)nrows) {
                const float xf = X[pos] - 1.0f - (float)x, yf = Y[pos] - 1.0f - (float)y;
                const float w00 = (1.0f - xf) * (1.0f - yf), w10 = xf * (1.0f - yf);
                const float w01 = (1.0f - xf) * yf, w11 = xf * yf;
                for (k = 0; k < nframes; k++) {
                    const long o = k * fsz + (long)nrows * x + y;
                    long ox1 = o, oy1 = o, oxy = o;
                    if (x < (unsigned)ncols - 1) ox1 += nrows;
                    if (y < (unsigned)nrows - 1) oy1 += 1;
                    if (x < (unsigned)ncols - 1 && y < (unsigned)nrows - 1) oxy += nrows + 1;
                    Iout[k * fsz + pos] = w00 * Iin[o] + w10 * Iin[ox1] + w01 * Iin[oy1] + w11 * Iin[oxy];
                }
            } else {
                for (k = 0; k < nframes; k++) Iout[k * fsz + pos] = oob;
            }
        }
}

/* ============================================================================================
 * Simoncelli derivatives: VerticalConvWO5 imageDerivatives.c:66-120, HorizontalConvWO5 :126-211,
 * TemporalConvWO2 :44-60, fstSimoncelli_c :309-384, sndSimoncelli_c :391-482.
 * Both convolutions are correlations with replicated borders, accumulated left to right.
 * ============================================================================================ */
static const float k_smooth[5] = {0.037659f, 0.249724f, 0.439911f, 0.249724f, 0.037659f};    /* FstDerivatives5.c:60 */
static const float k_d1[5] = {-0.104550f, -0.292315f, 0.0f, 0.292315f, 0.104550f};            /* FstDerivatives5.c:61 */
static const float k_d2[5] = {0.232905f, 0.002668f, -0.471147f, 0.002668f, 0.232905f};        /* SndDerivatives5.c:67 */

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static void conv5(int along_j, float *R, const float *I, const float *op, int nrows, int ncols)
{
    int i, j, t;
    for (j = 0; j < ncols; j++)
        for (i = 0; i < nrows; i++) {
            float tmp[5];
            for (t = 0; t < 5; t++) {
                const int ii = along_j ? i : clampi(i + t - 2, 0, nrows - 1);
                const int jj = along_j ? clampi(j + t - 2, 0, ncols - 1) : j;
                tmp[t] = I[P(ii, jj)] * op[t];
            }
            R[P(i, j)] = tmp[0] + tmp[1] + tmp[2] + tmp[3] + tmp[4];
        }
}

static void temporal2(float *R, const float *A, const float *Bm, long n)
{
    long p;
    for (p = 0; p < n; p++) R[p] = A[p] * 0.5f + Bm[p] * -0.5f;
}

void orc_fst(float *Idt, float *Idx, float *Idy, const float *It0, const float *It1,
             int nrows, int ncols, int nframes)
{
    const long fsz = (long)nrows * ncols;
    float *tmp = (float *)malloc(sizeof(float) * fsz);
    int k;
    if (!tmp) return;
    for (k = 0; k < nframes; k++) {
        const long o = k * fsz;
        temporal2(Idt + o, It0 + o, It1 + o, fsz);
        conv5(0, tmp, It1 + o, k_smooth, nrows, ncols);
        conv5(1, Idx + o, tmp, k_d1, nrows, ncols);
        conv5(1, tmp, It1 + o, k_smooth, nrows, ncols);
        conv5(0, Idy + o, tmp, k_d1, nrows, ncols);
    }
    free(tmp);
}

void orc_snd(float *Idxt, float *Idyt, float *Idxx, float *Idyy, float *Idxy,
             const float *It0, const float *It1, int nrows, int ncols, int nframes)
{
    const long fsz = (long)nrows * ncols;
    float *t1 = (float *)malloc(sizeof(float) * fsz * 3), *t2, *t3;
    int k;
    if (!t1) return;
    t2 = t1 + fsz; t3 = t2 + fsz;
    for (k = 0; k < nframes; k++) {
        const long o = k * fsz;
        conv5(0, t2, It0 + o, k_smooth, nrows, ncols); conv5(1, t1, t2, k_d1, nrows, ncols);
        conv5(0, t3, It1 + o, k_smooth, nrows, ncols); conv5(1, t2, t3, k_d1, nrows, ncols);
        temporal2(Idxt + o, t1, t2, fsz);
        conv5(1, t2, It0 + o, k_smooth, nrows, ncols); conv5(0, t1, t2, k_d1, nrows, ncols);
        conv5(1, t3, It1 + o, k_smooth, nrows, ncols); conv5(0, t2, t3, k_d1, nrows, ncols);
        temporal2(Idyt + o, t1, t2, fsz);
        conv5(0, t1, It1 + o, k_smooth, nrows, ncols); conv5(1, Idxx + o, t1, k_d2, nrows, ncols);
        conv5(1, t1, It1 + o, k_smooth, nrows, ncols); conv5(0, Idyy + o, t1, k_d2, nrows, ncols);
        conv5(1, t1, It1 + o, k_d1, nrows, ncols);     conv5(0, Idxy + o, t1, k_d1, nrows, ncols);
    }
    free(t1);
}

/* ============================================================================================
 * diffWeights6_2D_c imageDiffusionWeights.c:341-378: Dver :32, Dhor :73, Calc_wW/N/E/S :111-339
 * ============================================================================================ */
void orc_ddiff(float *wW, float *wN, float *wE, float *wS, const float *D,
               int nrows, int ncols, int nframes, float eps)
{
    const long fsz = (long)nrows * ncols;
    float *ver = (float *)malloc(sizeof(float) * fsz * nframes * 2), *hor;
    float *tW = (float *)calloc((size_t)fsz * 4, sizeof(float)), *tN, *tE, *tS;
    int i, j, k;
    if (!ver || !tW) { free(ver); free(tW); return; }
    hor = ver + fsz * nframes;
    tN = tW + fsz; tE = tN + fsz; tS = tE + fsz;
    for (k = 0; k < nframes; k++)
        for (j = 0; j < ncols; j++)
            for (i = 0; i < nrows; i++) {
                const long p = P(i, j) + k * fsz;
                const float *Dk = D + k * fsz;
                float A, B;
                A = 0.25f * Dk[P(clampi(i - 1, 0, nrows - 1), j)]; B = -0.25f * Dk[P(clampi(i + 1, 0, nrows - 1), j)];
                ver[p] = A + B;
                A = 0.25f * Dk[P(i, clampi(j - 1, 0, ncols - 1))]; B = -0.25f * Dk[P(i, clampi(j + 1, 0, ncols - 1))];
                hor[p] = A + B;
            }
    for (k = 0; k < nframes; k++)
        for (j = 0; j < ncols; j++)
            for (i = 0; i < nrows; i++) {
                const long q = P(i, j), p = q + k * fsz;
                float A, B, t;
                if (j >= 1) {
                    A = D[p] - D[p - nrows]; B = ver[p] + ver[p - nrows]; A = A * A; B = B * B; t = A + B;
                    if (k == 0) tW[q] = t; else if (t > tW[q]) tW[q] = t;
                }
                if (i >= 1) {
                    A = D[p] - D[p - 1]; B = hor[p] + hor[p - 1]; A = A * A; B = B * B; t = A + B;
                    if (k == 0) tN[q] = t; else if (t > tN[q]) tN[q] = t;
                }
                if (j <= ncols - 2) {
                    A = D[p] - D[p + nrows]; B = ver[p] + ver[p + nrows]; A = A * A; B = B * B; t = A + B;
                    if (k == 0) tE[q] = t; else if (t > tE[q]) tE[q] = t;
                }
                if (i <= nrows - 2) {
                    A = D[p] - D[p + 1]; B = hor[p] + hor[p + 1]; A = A * A; B = B * B; t = A + B;
                    if (k == 0) tS[q] = t; else if (t > tS[q]) tS[q] = t;
                }
            }
    for (j = 0; j < ncols; j++)
        for (i = 0; i < nrows; i++) {
            const long q = P(i, j);
            wW[q] = (j >= 1) ? 1.0f / (float)sqrt(tW[q] + eps) : 0.0f;
            wN[q] = (i >= 1) ? 1.0f / (float)sqrt(tN[q] + eps) : 0.0f;
            wE[q] = (j <= ncols - 2) ? 1.0f / (float)sqrt(tE[q] + eps) : 0.0f;
            wS[q] = (i <= nrows - 2) ? 1.0f / (float)sqrt(tS[q] + eps) : 0.0f;
        }
    free(ver);
    free(tW);
}
