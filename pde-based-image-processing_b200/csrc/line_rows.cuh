// line_rows.cuh -- the tridiagonal row of one pixel for a line solve, shared by the line-relaxation
// kernels (sweeps_line.cu, sweeps_window.cu). Follows the reference's line solvers with their
// one-sided border rows (e.g. middleColumn_llin4 opticalflowSolvers.c:2548-2639, westColumn_elin4
// :1790-1792, middleColumn4 disparitySolvers.c:1500-1587, TDMA_mcolumn_ALR_4 pdeSolvers.c:530-650,
// TDMAcolumn_ALR_8 :1132-1266).
#pragma once
#include "stencil_math.cuh"

// Raw operands of one pixel's tridiagonal rows. load() issues every global load unconditionally
// (neighbours that do not exist are redirected to the pixel itself), so that a thread can have the
// loads of several pixels in flight before it touches any of them: phase A is latency-bound otherwise.
// DIR: 0 = lines along i, first unknown first; 1 = lines along j, second unknown first (the reference's
// row pass); 2 = a row pass executed as lines along i of the TRANSPOSED problem (see alr_run).
template <int FAM, int DIR>
struct PixelRaw {
    using F = Fam<FAM>;
    static constexpr int NN = F::EIGHT ? 8 : 4;
    static constexpr int prev = (DIR & 1) == 0 ? W_N : W_W, next = (DIR & 1) == 0 ? W_S : W_E;
    static constexpr int qa = (F::NUNK == 2 && DIR != 0) ? 1 : 0, qb = 1 - qa;
    float w[NN];
    float xn[F::NUNK][NN];      // unknowns at the perpendicular neighbours
    float x0n[F::LATE ? F::NUNK : 1][NN], x0c[F::LATE ? F::NUNK : 1];
    float C[F::NUNK], D[F::NUNK], xo[F::NUNK], M;
    unsigned exmask;

    // `ip` = pixel index inside the problem (int), pointers in `s` already point at the problem.
    // LOADXN = false: the unknowns at the perpendicular neighbours are NOT read from global memory (the
    // caller supplies xn[][] from somewhere else, see sweeps_window.cu).
    template <bool LOADXN = true>
    __device__ __forceinline__ void load(const SysView &s, int ip, int i, int j)
    {
        const int nr = s.nrows, nc = s.ncols;
        const bool eN = i > 0, eS = i < nr - 1, eW = j > 0, eE = j < nc - 1;
        const bool ex[8] = {eW, eN, eE, eS, eN && eW, eN && eE, eS && eE, eS && eW};
        const int off[8] = {-nr, -1, nr, 1, -nr - 1, nr - 1, nr + 1, -nr + 1};
        exmask = 0;
#pragma unroll
        for (int n = 0; n < NN; n++) {
            exmask |= ex[n] ? (1u << n) : 0u;
            const int pn = ex[n] ? ip + off[n] : ip;
            w[n] = s.w[n][ip];
            const bool inl = (n == prev) || (n == next);
#pragma unroll
            for (int q = 0; q < F::NUNK; q++) {
                if (!inl && LOADXN) xn[q][n] = s.x[q][pn];
                if (F::LATE) x0n[q][n] = s.x0[q][pn];
            }
        }
#pragma unroll
        for (int q = 0; q < F::NUNK; q++) {
            if (F::LATE) x0c[q] = s.x0[q][ip];
            C[q] = s.c[q][ip];
            D[q] = s.d[q][ip];
            xo[q] = s.x[q][ip];
        }
        M = (F::NUNK == 2) ? s.m[ip] : 0.0f;
    }

    // a,c: sub/super diagonal; b[q],d[q]: diagonal and right-hand side of unknown q. d[qa] is complete
    // (coupling taken with the other unknown's current value); d[qb] lacks the coupling term, which is
    // m * x_qa(new) and is added by the solver.
    __device__ __forceinline__ void rows(float &a, float &c, float (&b)[2], float (&d)[2], float &m) const
    {
        a = (exmask >> prev) & 1 ? -w[prev] : 0.0f;
        c = (exmask >> next) & 1 ? -w[next] : 0.0f;
        float bsum = 0.0f, dsum[2] = {0.0f, 0.0f};
#pragma unroll
        for (int n = 0; n < NN; n++) {
            const bool e = (exmask >> n) & 1;
            const bool inl = (n == prev) || (n == next);
            bsum += e ? w[n] : 0.0f;
#pragma unroll
            for (int q = 0; q < F::NUNK; q++) {
                float t = 0.0f;
                if (F::LATE) t = x0n[q][n] - x0c[q];
                if (!inl) t += xn[q][n];
                if (F::LATE || !inl) dsum[q] += e ? w[n] * t : 0.0f;
            }
        }
        m = 0.0f;
        b[1] = 1.0f; d[1] = 0.0f;
        if (F::PDE) {
            if (!is_nan(D[0])) { b[0] = D[0]; d[0] = dsum[0] + C[0]; }
            else {
                if (F::EIGHT)   // pdeSolvers.c:1179: wNW twice, wNE never (on a transposed problem SW <-> NE)
                    b[0] = (w[W_N] + w[W_S] + w[W_W] + w[W_E])
                         + (w[W_NW % NN] + w[W_NW % NN] + w[(DIR == 2 ? W_NE : W_SW) % NN] + w[W_SE % NN]);
                else b[0] = bsum;
                d[0] = dsum[0];
            }
            return;
        }
#pragma unroll
        for (int q = 0; q < F::NUNK; q++) {
            b[q] = bsum; d[q] = dsum[q];
            if (!is_nan(C[q])) {
                b[q] += D[q];
                d[q] += C[q];
                if (F::NUNK == 2) {
                    if (q == qa) d[q] -= M * xo[qb];
                    else m = M;
                }
            }
        }
    }
};

// 1/x to 1 ulp (MUFU.RCP). The sweeps only have to reach the reference's fixed point, and a relaxation
// step is a contraction, so a 1-ulp reciprocal changes nothing that can be observed; it shortens the
// serial dependency chain of the elimination from ~80 to ~25 cycles per row.
__device__ __forceinline__ float fast_rcp(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
