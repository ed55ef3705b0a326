// sweeps_point.cu -- kernel generation 1, solver 1: red-black point SOR, one launch per full sweep.
//
// Generation 0 runs one launch per colour; each of them drags every coefficient sector through the SMs and
// uses half of it. Here a CTA owns a 128 x 8 tile (fast axis = Matlab row index i) and does the whole sweep
// on it:
//   1. the unknowns (and the fixed U,V of the late-linearisation families) of the tile plus a 2-pixel halo go
//      to shared memory; every thread loads the coefficients of its own 4 consecutive pixels ONCE (float4);
//   2. red half-sweep on the tile plus a 1-pixel ring (the ring's red pixels are recomputed here so that the
//      black pixels on the tile edge see new red values without any inter-CTA synchronisation);
//   3. black half-sweep on the tile;
//   4. the tile goes to X_out. The sweep is out of place (neighbouring CTAs read old values from X_in), the
//      two arrays swap roles from sweep to sweep, and the reference's border fill (opticalflowSolvers.c:161-179)
//      runs as generation 0's border kernel on X_out.
// HBM traffic per sweep = algorithmic (every field once, unknowns written once) plus the halo re-reads, which
// hit L2. Same ordering and the same arithmetic (point_formula) as generation 0.
#include "stencil_math.cuh"
#include <stdlib.h>

namespace {

constexpr int PT_I = 128, PT_J = 8, PT_H = 2;                 // tile and halo
constexpr int PT_C0 = 4;                                      // shared-memory column of the tile's first pixel (16-B aligned)
constexpr int PS_I = PT_I + 2 * PT_C0, PS_J = PT_J + 2 * PT_H; // shared-memory extent (columns PT_C0-2 .. PT_C0+PT_I+1 are used)

__device__ __forceinline__ float f4c(const float4 &v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }

// ALIGNED is a compile-time flag: a run-time one makes ptxas predicate the vector and the scalar loads into one
// instruction stream, where the (predicated-off) scalar loads wait for the vector loads that share their registers.
template <int FAM, bool ALIGNED>
#ifndef PDEGPU_PT_MINB
#define PDEGPU_PT_MINB 2
#endif
__global__ void __launch_bounds__(256, PDEGPU_PT_MINB)
rb_tile_kernel(SysView s, float *__restrict__ xo0, float *__restrict__ xo1, float omega)
{
    using F = Fam<FAM>;
    constexpr int NUNK = F::NUNK;
    constexpr int NF = NUNK * (F::LATE ? 2 : 1);              // fields kept in shared memory: x[q], then x0[q]
    __shared__ __align__(16) float sm[NF][PS_J][PS_I];
    const int nr = s.nrows, nc = s.ncols;
    const int i0 = blockIdx.x * PT_I, j0 = blockIdx.y * PT_J;
    const long long base = (long long)blockIdx.z * s.bstride;
    const int tid = threadIdx.x;
    const float *fld[4] = {s.x[0] + base, NUNK == 2 ? s.x[1] + base : nullptr,
                           F::LATE ? s.x0[0] + base : nullptr, (F::LATE && NUNK == 2) ? s.x0[1] + base : nullptr};
    auto field = [&](int f) -> const float * { return fld[f < NUNK ? f : 2 + (NUNK == 2 ? f - NUNK : 0)]; };
    // coefficients of this thread's 4 pixels
    const int ti = (tid & 31) * 4, tj = tid >> 5;
    const int gi0 = i0 + ti, gj = j0 + tj;
    float4 w4[4], C4[2], D4[2], M4;
    const bool row_ok = gj < nc && gi0 < nr;
    {
        const int gic = min(gi0, ALIGNED ? nr - 4 : nr - 1);     // (aligned: nr is a multiple of 4, the clamp keeps 16-B alignment)
        const long long p = base + (long long)min(gj, nc - 1) * nr + gic;
        const int room = nr - 1 - gic;
        auto ldv = [&](const float *f) -> float4 {
            if (ALIGNED) return *reinterpret_cast<const float4 *>(f + p);
            float4 v;
            v.x = f[p]; v.y = f[p + min(1, room)]; v.z = f[p + min(2, room)]; v.w = f[p + min(3, room)];
            return v;
        };
#pragma unroll
        for (int n = 0; n < 4; n++) w4[n] = ldv(s.w[n]);
#pragma unroll
        for (int q = 0; q < NUNK; q++) { C4[q] = ldv(s.c[q]); D4[q] = ldv(s.d[q]); }
        if (NUNK == 2) M4 = ldv(s.m);
    }
    // the red pixels of the 1-pixel ring around the tile (recomputed here): 138 of them, one per thread, coefficients
    // loaded now so that their latency overlaps everything else. Frame coordinates (origin = tile - 1): a pixel is
    // red when li + lj is even (i0 and j0 are even).
    int rli = -1, rlj = 0;
    if (tid < 65) { rli = 2 * tid; rlj = 0; }
    else if (tid < 130) { rli = 2 * (tid - 65) + 1; rlj = PT_J + 1; }
    else if (tid < 134) { rli = 0; rlj = 2 * (tid - 130) + 2; }
    else if (tid < 138) { rli = PT_I + 1; rlj = 2 * (tid - 134) + 1; }
    float rw[4], rC[2], rD[2], rM = 0.f;
    bool ring_ok = false;
    if (rli >= 0) {
        const int gi = i0 + rli - 1, gjr = j0 + rlj - 1;
        ring_ok = gi >= 1 && gi <= nr - 2 && gjr >= 1 && gjr <= nc - 2;
        const long long p = base + (long long)min(max(gjr, 0), nc - 1) * nr + min(max(gi, 0), nr - 1);
#pragma unroll
        for (int n = 0; n < 4; n++) rw[n] = s.w[n][p];
        rC[0] = s.c[0][p]; rD[0] = s.d[0][p];
        rC[1] = NUNK == 2 ? s.c[NUNK - 1][p] : 0.f; rD[1] = NUNK == 2 ? s.d[NUNK - 1][p] : 0.f;
        if (NUNK == 2) rM = s.m[p];
    }
    // 1. unknowns + halo (indices clamped into the image: clamped copies are never used by an interior update).
    //    The 128 tile columns of a row are one aligned run: float4 in, float4 out; the 2+2 halo elements are scalar.
    if (ALIGNED && i0 + PT_I <= nr) {
        for (int t = tid; t < (PT_I / 4) * PS_J; t += 256) {
            const int v = t % (PT_I / 4), lj = t / (PT_I / 4);
            const int gj = min(max(j0 + lj - PT_H, 0), nc - 1);
            const long long p = (long long)gj * nr + i0 + 4 * v;
#pragma unroll
            for (int f = 0; f < NF; f++)
                *reinterpret_cast<float4 *>(&sm[f][lj][PT_C0 + 4 * v]) = *reinterpret_cast<const float4 *>(field(f) + p);
        }
    } else {
        for (int t = tid; t < PT_I * PS_J; t += 256) {
            const int li = t % PT_I, lj = t / PT_I;
            const int gi = min(i0 + li, nr - 1), gj = min(max(j0 + lj - PT_H, 0), nc - 1);
            const long long p = (long long)gj * nr + gi;
#pragma unroll
            for (int f = 0; f < NF; f++) sm[f][lj][PT_C0 + li] = field(f)[p];
        }
    }
    for (int t = tid; t < 2 * PT_H * PS_J; t += 256) {
        const int h = t % (2 * PT_H), lj = t / (2 * PT_H);
        const int c = h < PT_H ? PT_C0 - PT_H + h : PT_C0 + PT_I + h - PT_H;       // columns 2,3 and 132,133
        const int gi = min(max(i0 + c - PT_C0, 0), nr - 1), gj = min(max(j0 + lj - PT_H, 0), nc - 1);
        const long long p = (long long)gj * nr + gi;
#pragma unroll
        for (int f = 0; f < NF; f++) sm[f][lj][c] = field(f)[p];
    }
    __syncthreads();

    // scalar update of the pixel at shared-memory position (c, lj) (ring pixels)
    auto update = [&](int c, int lj, const float (&w)[4], const float (&C)[2], const float (&D)[2], float M) {
        float xn[2][4], xc[2], x0n[2][4], x0c[2], out[2];
#pragma unroll
        for (int q = 0; q < NUNK; q++) {
            xc[q] = sm[q][lj][c];
            xn[q][W_W] = sm[q][lj - 1][c]; xn[q][W_E] = sm[q][lj + 1][c];
            xn[q][W_N] = sm[q][lj][c - 1]; xn[q][W_S] = sm[q][lj][c + 1];
            if (F::LATE) {
                x0c[q] = sm[NUNK + q][lj][c];
                x0n[q][W_W] = sm[NUNK + q][lj - 1][c]; x0n[q][W_E] = sm[NUNK + q][lj + 1][c];
                x0n[q][W_N] = sm[NUNK + q][lj][c - 1]; x0n[q][W_S] = sm[NUNK + q][lj][c + 1];
            }
        }
        point_formula<FAM>(w, xn, xc, x0n, x0c, C, D, M, omega, out);
#pragma unroll
        for (int q = 0; q < NUNK; q++) sm[q][lj][c] = out[q];
    };
    // half-sweep on this thread's 4 pixels: one float4 per row and field out of shared memory (conflict-free),
    // the updated unknowns go back as one float4; xk[q] keeps the thread's current values
    float4 xk[2];
    auto own = [&](int colour) {
        const int lj = tj + PT_H, c = PT_C0 + ti;
        float4 vc[NF], vw[NF], ve[NF];
        float lf[NF], rt[NF];
#pragma unroll
        for (int f = 0; f < NF; f++) {
            vc[f] = *reinterpret_cast<const float4 *>(&sm[f][lj][c]);
            vw[f] = *reinterpret_cast<const float4 *>(&sm[f][lj - 1][c]);
            ve[f] = *reinterpret_cast<const float4 *>(&sm[f][lj + 1][c]);
            lf[f] = sm[f][lj][c - 1]; rt[f] = sm[f][lj][c + 4];
        }
#pragma unroll
        for (int q = 0; q < NUNK; q++) xk[q] = vc[q];
        if (!row_ok || gj < 1 || gj > nc - 2) return;
        float res[2][4];
#pragma unroll
        for (int q = 0; q < NUNK; q++) { res[q][0] = vc[q].x; res[q][1] = vc[q].y; res[q][2] = vc[q].z; res[q][3] = vc[q].w; }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int gi = gi0 + k;
            if (gi < 1 || gi > nr - 2 || ((gi + gj) & 1) != colour) continue;
            float xn[2][4], xc[2], x0n[2][4], x0c[2], out[2];
#pragma unroll
            for (int q = 0; q < NUNK; q++) {
                xc[q] = f4c(vc[q], k);
                xn[q][W_W] = f4c(vw[q], k); xn[q][W_E] = f4c(ve[q], k);
                xn[q][W_N] = k > 0 ? f4c(vc[q], k - 1) : lf[q];
                xn[q][W_S] = k < 3 ? f4c(vc[q], k + 1) : rt[q];
                if (F::LATE) {
                    x0c[q] = f4c(vc[NUNK + q], k);
                    x0n[q][W_W] = f4c(vw[NUNK + q], k); x0n[q][W_E] = f4c(ve[NUNK + q], k);
                    x0n[q][W_N] = k > 0 ? f4c(vc[NUNK + q], k - 1) : lf[NUNK + q];
                    x0n[q][W_S] = k < 3 ? f4c(vc[NUNK + q], k + 1) : rt[NUNK + q];
                }
            }
            const float w[4] = {f4c(w4[0], k), f4c(w4[1], k), f4c(w4[2], k), f4c(w4[3], k)};
            const float C[2] = {f4c(C4[0], k), NUNK == 2 ? f4c(C4[NUNK - 1], k) : 0.f};
            const float D[2] = {f4c(D4[0], k), NUNK == 2 ? f4c(D4[NUNK - 1], k) : 0.f};
            point_formula<FAM>(w, xn, xc, x0n, x0c, C, D, NUNK == 2 ? f4c(M4, k) : 0.f, omega, out);
#pragma unroll
            for (int q = 0; q < NUNK; q++) res[q][k] = out[q];
        }
#pragma unroll
        for (int q = 0; q < NUNK; q++) {
            xk[q] = make_float4(res[q][0], res[q][1], res[q][2], res[q][3]);
            *reinterpret_cast<float4 *>(&sm[q][lj][c]) = xk[q];
        }
    };
    // 2. red: own pixels, then the red pixels of the 1-pixel ring around the tile
    own(0);
    if (ring_ok) update(rli + PT_C0 - 1, rlj + PT_H - 1, rw, rC, rD, rM);
    __syncthreads();
    // 3. black
    own(1);
    // 4. own pixels -> X_out straight from registers
    if (row_ok) {
        float *xo[2] = {xo0 + base, NUNK == 2 ? xo1 + base : nullptr};
        const long long p = (long long)gj * nr + gi0;
#pragma unroll
        for (int q = 0; q < NUNK; q++) {
            if (ALIGNED) *reinterpret_cast<float4 *>(xo[q] + p) = xk[q];
            else for (int k = 0; k < 4 && gi0 + k < nr; k++) xo[q][p + k] = f4c(xk[q], k);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Same sweep with every operand staged through shared memory by cp.async (aligned problems only).
// rb_tile_kernel keeps the coefficients of a thread's 4 pixels in registers from the first instruction to the last
// (36 + 11 registers), which holds it at 126 registers = 2 CTAs per SM: when both CTAs compute, nothing is in flight
// (ncu: 24 % occupancy, long_scoreboard 4.6, DRAM 45 % busy). Here a CTA fires all its copies (unknown tile + halo,
// coefficient tile, ring coefficients) without holding a register, waits once, and computes out of shared memory.
// Measured (B200, 64 x 480x640 llin4, us per sweep): register-staged 319, cp.async with 2 CTAs per SM 311, with 3 CTAs
// per SM (80 registers, 192 B of spills) 365 -- the sweep is bound by the per-warp instruction latency of the update
// (~640 instructions per thread and tile, 16 warps per SM), not by the bytes in flight; kept because it is the faster one.
// Same ordering, same point_formula, same bits as rb_tile_kernel.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *sm, const void *g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(sm)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async4(void *sm, const void *g)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(sm)), "l"(g) : "memory");
}

constexpr int PT_RING = 144;                                  // ring pixels (138) padded

template <int FAM> constexpr int pt_async_floats()
{
    using F = Fam<FAM>;
    return F::NUNK * (F::LATE ? 2 : 1) * PS_J * PS_I + (4 + 2 * F::NUNK + (F::NUNK == 2 ? 1 : 0)) * (PT_J * PT_I + PT_RING);
}

#ifndef PDEGPU_PTA_MINB
#define PDEGPU_PTA_MINB 2
#endif
template <int FAM>
__global__ void __launch_bounds__(256, PDEGPU_PTA_MINB)
rb_tile_async_kernel(SysView s, float *__restrict__ xo0, float *__restrict__ xo1, float omega)
{
    using F = Fam<FAM>;
    constexpr int NUNK = F::NUNK;
    constexpr int NF = NUNK * (F::LATE ? 2 : 1);              // x[q], then x0[q]
    constexpr int NC = 4 + 2 * NUNK + (NUNK == 2 ? 1 : 0);    // w[0..3], C[q], D[q], M
    constexpr int CI_C = 4, CI_D = 4 + NUNK, CI_M = 4 + 2 * NUNK;
    extern __shared__ __align__(16) float dsm[];
    float (*sm)[PS_J][PS_I] = reinterpret_cast<float (*)[PS_J][PS_I]>(dsm);
    float (*cf)[PT_J][PT_I] = reinterpret_cast<float (*)[PT_J][PT_I]>(dsm + NF * PS_J * PS_I);
    float *rg = dsm + NF * PS_J * PS_I + NC * PT_J * PT_I;    // ring coefficients: rg[f * PT_RING + tid]
    const int nr = s.nrows, nc = s.ncols;
    const int i0 = blockIdx.x * PT_I, j0 = blockIdx.y * PT_J;
    const long long base = (long long)blockIdx.z * s.bstride;
    const int tid = threadIdx.x;
    const float *fld[4] = {s.x[0] + base, NUNK == 2 ? s.x[1] + base : nullptr,
                           F::LATE ? s.x0[0] + base : nullptr, (F::LATE && NUNK == 2) ? s.x0[1] + base : nullptr};
    auto field = [&](int f) -> const float * { return fld[f < NUNK ? f : 2 + (NUNK == 2 ? f - NUNK : 0)]; };
    const float *cptr[9] = {s.w[0], s.w[1], s.w[2], s.w[3], nullptr, nullptr, nullptr, nullptr, nullptr};
#pragma unroll
    for (int q = 0; q < NUNK; q++) { cptr[CI_C + q] = s.c[q]; cptr[CI_D + q] = s.d[q]; }
    if (NUNK == 2) cptr[CI_M] = s.m;
    const int ti = (tid & 31) * 4, tj = tid >> 5;
    const int gi0 = i0 + ti, gj = j0 + tj;
    const bool row_ok = gj < nc && gi0 < nr;
    // ---- all copies, nothing held in registers ----
    {
        const long long p = base + (long long)min(gj, nc - 1) * nr + min(gi0, nr - 4);   // nr is a multiple of 4
#pragma unroll
        for (int f = 0; f < NC; f++) cp_async16(&cf[f][tj][ti], cptr[f] + p);
    }
    int rli = -1, rlj = 0;                                    // the red pixels of the 1-pixel ring (see rb_tile_kernel)
    if (tid < 65) { rli = 2 * tid; rlj = 0; }
    else if (tid < 130) { rli = 2 * (tid - 65) + 1; rlj = PT_J + 1; }
    else if (tid < 134) { rli = 0; rlj = 2 * (tid - 130) + 2; }
    else if (tid < 138) { rli = PT_I + 1; rlj = 2 * (tid - 134) + 1; }
    bool ring_ok = false;
    if (rli >= 0) {
        const int gi = i0 + rli - 1, gjr = j0 + rlj - 1;
        ring_ok = gi >= 1 && gi <= nr - 2 && gjr >= 1 && gjr <= nc - 2;
        const long long p = base + (long long)min(max(gjr, 0), nc - 1) * nr + min(max(gi, 0), nr - 1);
#pragma unroll
        for (int f = 0; f < NC; f++) cp_async4(&rg[f * PT_RING + tid], cptr[f] + p);
    }
    for (int t = tid; t < (PT_I / 4) * PS_J; t += 256) {
        const int v = t % (PT_I / 4), lj = t / (PT_I / 4);
        const int gjc = min(max(j0 + lj - PT_H, 0), nc - 1);
        const long long p = (long long)gjc * nr + min(i0 + 4 * v, nr - 4);
#pragma unroll
        for (int f = 0; f < NF; f++) cp_async16(&sm[f][lj][PT_C0 + 4 * v], field(f) + p);
    }
    for (int t = tid; t < 2 * PT_H * PS_J; t += 256) {
        const int h = t % (2 * PT_H), lj = t / (2 * PT_H);
        const int c = h < PT_H ? PT_C0 - PT_H + h : PT_C0 + PT_I + h - PT_H;       // columns 2,3 and 132,133
        const int gi = min(max(i0 + c - PT_C0, 0), nr - 1), gjc = min(max(j0 + lj - PT_H, 0), nc - 1);
        const long long p = (long long)gjc * nr + gi;
#pragma unroll
        for (int f = 0; f < NF; f++) cp_async4(&sm[f][lj][c], field(f) + p);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();

    auto update = [&](int c, int lj) {                        // scalar update of a ring pixel
        float xn[2][4], xc[2], x0n[2][4], x0c[2], out[2], w[4], C[2] = {0.f, 0.f}, D[2] = {0.f, 0.f};
#pragma unroll
        for (int n = 0; n < 4; n++) w[n] = rg[n * PT_RING + tid];
#pragma unroll
        for (int q = 0; q < NUNK; q++) {
            C[q] = rg[(CI_C + q) * PT_RING + tid]; D[q] = rg[(CI_D + q) * PT_RING + tid];
            xc[q] = sm[q][lj][c];
            xn[q][W_W] = sm[q][lj - 1][c]; xn[q][W_E] = sm[q][lj + 1][c];
            xn[q][W_N] = sm[q][lj][c - 1]; xn[q][W_S] = sm[q][lj][c + 1];
            if (F::LATE) {
                x0c[q] = sm[NUNK + q][lj][c];
                x0n[q][W_W] = sm[NUNK + q][lj - 1][c]; x0n[q][W_E] = sm[NUNK + q][lj + 1][c];
                x0n[q][W_N] = sm[NUNK + q][lj][c - 1]; x0n[q][W_S] = sm[NUNK + q][lj][c + 1];
            }
        }
        point_formula<FAM>(w, xn, xc, x0n, x0c, C, D, NUNK == 2 ? rg[CI_M * PT_RING + tid] : 0.f, omega, out);
#pragma unroll
        for (int q = 0; q < NUNK; q++) sm[q][lj][c] = out[q];
    };
    float4 xk[2];
    auto own = [&](int colour) {
        const int lj = tj + PT_H, c = PT_C0 + ti;
        float4 vc[NF], vw[NF], ve[NF];
        float lf[NF], rt[NF];
#pragma unroll
        for (int f = 0; f < NF; f++) {
            vc[f] = *reinterpret_cast<const float4 *>(&sm[f][lj][c]);
            vw[f] = *reinterpret_cast<const float4 *>(&sm[f][lj - 1][c]);
            ve[f] = *reinterpret_cast<const float4 *>(&sm[f][lj + 1][c]);
            lf[f] = sm[f][lj][c - 1]; rt[f] = sm[f][lj][c + 4];
        }
#pragma unroll
        for (int q = 0; q < NUNK; q++) xk[q] = vc[q];
        if (!row_ok || gj < 1 || gj > nc - 2) return;
        float4 k4[NC];
#pragma unroll
        for (int f = 0; f < NC; f++) k4[f] = *reinterpret_cast<const float4 *>(&cf[f][tj][ti]);
        float res[2][4];
#pragma unroll
        for (int q = 0; q < NUNK; q++) { res[q][0] = vc[q].x; res[q][1] = vc[q].y; res[q][2] = vc[q].z; res[q][3] = vc[q].w; }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int gi = gi0 + k;
            if (gi < 1 || gi > nr - 2 || ((gi + gj) & 1) != colour) continue;
            float xn[2][4], xc[2], x0n[2][4], x0c[2], out[2];
#pragma unroll
            for (int q = 0; q < NUNK; q++) {
                xc[q] = f4c(vc[q], k);
                xn[q][W_W] = f4c(vw[q], k); xn[q][W_E] = f4c(ve[q], k);
                xn[q][W_N] = k > 0 ? f4c(vc[q], k - 1) : lf[q];
                xn[q][W_S] = k < 3 ? f4c(vc[q], k + 1) : rt[q];
                if (F::LATE) {
                    x0c[q] = f4c(vc[NUNK + q], k);
                    x0n[q][W_W] = f4c(vw[NUNK + q], k); x0n[q][W_E] = f4c(ve[NUNK + q], k);
                    x0n[q][W_N] = k > 0 ? f4c(vc[NUNK + q], k - 1) : lf[NUNK + q];
                    x0n[q][W_S] = k < 3 ? f4c(vc[NUNK + q], k + 1) : rt[NUNK + q];
                }
            }
            const float w[4] = {f4c(k4[0], k), f4c(k4[1], k), f4c(k4[2], k), f4c(k4[3], k)};
            const float C[2] = {f4c(k4[CI_C], k), NUNK == 2 ? f4c(k4[CI_C + NUNK - 1], k) : 0.f};
            const float D[2] = {f4c(k4[CI_D], k), NUNK == 2 ? f4c(k4[CI_D + NUNK - 1], k) : 0.f};
            point_formula<FAM>(w, xn, xc, x0n, x0c, C, D, NUNK == 2 ? f4c(k4[NC - 1], k) : 0.f, omega, out);
#pragma unroll
            for (int q = 0; q < NUNK; q++) res[q][k] = out[q];
        }
#pragma unroll
        for (int q = 0; q < NUNK; q++) {
            xk[q] = make_float4(res[q][0], res[q][1], res[q][2], res[q][3]);
            *reinterpret_cast<float4 *>(&sm[q][lj][c]) = xk[q];
        }
    };
    own(0);
    if (ring_ok) update(rli + PT_C0 - 1, rlj + PT_H - 1);
    __syncthreads();
    own(1);
    if (row_ok) {
        float *xo[2] = {xo0 + base, NUNK == 2 ? xo1 + base : nullptr};
        const long long p = (long long)gj * nr + gi0;
#pragma unroll
        for (int q = 0; q < NUNK; q++) *reinterpret_cast<float4 *>(xo[q] + p) = xk[q];
    }
}

template <int NUNK>
__global__ void border_fill_tile_kernel(float *x0, float *x1, int nr, int nc, long long bstride)
{
    // border pixel := nearest interior pixel (rows first, then columns: corners take the diagonal neighbour)
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int per = 2 * nr + 2 * nc;
    if (t >= per) return;
    int i, j;
    if (t < nr)               { i = t;            j = 0; }
    else if (t < 2 * nr)      { i = t - nr;       j = nc - 1; }
    else if (t < 2 * nr + nc) { i = 0;            j = t - 2 * nr; }
    else                      { i = nr - 1;       j = t - 2 * nr - nc; }
    const int ic = min(max(i, 1), nr - 2), jc = min(max(j, 1), nc - 2);
    const long long base = (long long)blockIdx.y * bstride;
    float *x[2] = {x0, x1};
#pragma unroll
    for (int q = 0; q < NUNK; q++) x[q][base + (long long)j * nr + i] = x[q][base + (long long)jc * nr + ic];
}

}  // namespace
// two sweeps per pass over HBM (sweeps_tpoint.cu)
int relax_tpoint(pdegpu_ctx *ctx, const pdegpu_system *sys, float *const cur[2], float *const alt[2], int pairs, float omega, bool *result_in_alt);
namespace {

template <int FAM>
int run_point_tiles(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega)
{
    using F = Fam<FAM>;
    constexpr int NUNK = F::NUNK;
    const long long npix = (long long)sys->nrows * sys->ncols;
    const long long fld = ((long long)sys->batch * sys->batch_stride + 3) & ~3ll;
    int rc = pdegpu_scratch_reserve(ctx, (size_t)NUNK * fld * sizeof(float));
    if (rc) return rc;
    float *alt[2] = {(float *)ctx->scratch, (float *)ctx->scratch + fld};
    bool al = (sys->nrows % 4 == 0) && (sys->batch_stride % 4 == 0);
    auto a16 = [](const void *q) { return ((uintptr_t)q & 15) == 0; };
    for (int n = 0; n < 4; n++) al = al && a16(sys->w[n]);
    for (int q = 0; q < NUNK; q++) al = al && a16(sys->x[q]) && a16(sys->c[q]) && a16(sys->d[q]);
    if (NUNK == 2) al = al && a16(sys->m);
    static const int use_async = getenv("PDEGPU_POINT_ASYNC") ? atoi(getenv("PDEGPU_POINT_ASYNC")) : 1;
    if (al && use_async) {
        {   // every launch: the attribute is per device and the call is cheap (no static per-ordinal bookkeeping)
        cudaError_t e = cudaFuncSetAttribute(rb_tile_async_kernel<FAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(pt_async_floats<FAM>() * sizeof(float)));
        if (e != cudaSuccess) return pdegpu_check_cuda(ctx, e, "cudaFuncSetAttribute(rb_tile_async_kernel)");
        }
    }
    SysView v = make_view(sys);
    dim3 grid((sys->nrows + PT_I - 1) / PT_I, (sys->ncols + PT_J - 1) / PT_J, sys->batch);
    const int per = 2 * sys->nrows + 2 * sys->ncols;
    dim3 bgrid((per + 127) / 128, sys->batch);
    float *cur[2] = {sys->x[0], sys->x[1]}, *nxt[2] = {alt[0], alt[1]};
    // pairs of sweeps: the temporally blocked kernel (every field read once per TWO sweeps) where it wins (enough strips
    // to fill the SMs three times: sweeps_tpoint.cu; PDEGPU_POINT_WINDOW=1 / 0 forces it on / off); a last odd sweep and
    // everything else: one sweep per pass as before.
    if (al && iter >= 2) {
        bool in_alt = false;
        rc = relax_tpoint(ctx, sys, cur, nxt, iter / 2, omega, &in_alt);
        if (rc == PDEGPU_OK) {
            if (in_alt) for (int q = 0; q < 2; q++) { float *t = cur[q]; cur[q] = nxt[q]; nxt[q] = t; }
            iter &= 1;
        } else if (rc != PDEGPU_ERR_UNSUPPORTED) return rc;
    }
    for (int it = 0; it < iter; it++) {
        v.x[0] = cur[0]; v.x[1] = cur[1];
        PDEGPU_PROF(ctx, "rb_tile_kernel", sweep_bytes<FAM>() * (double)npix * sys->batch);
        if (al && use_async) rb_tile_async_kernel<FAM><<<grid, 256, pt_async_floats<FAM>() * sizeof(float), ctx->stream>>>(v, nxt[0], nxt[1], omega);
        else if (al) rb_tile_kernel<FAM, true><<<grid, 256, 0, ctx->stream>>>(v, nxt[0], nxt[1], omega);
        else rb_tile_kernel<FAM, false><<<grid, 256, 0, ctx->stream>>>(v, nxt[0], nxt[1], omega);
        PDEGPU_LAUNCH_CHECK(ctx, "rb_tile_kernel");
        PDEGPU_PROF(ctx, "border_fill_kernel", 0);
        border_fill_tile_kernel<NUNK><<<bgrid, 128, 0, ctx->stream>>>(nxt[0], nxt[1], sys->nrows, sys->ncols, sys->batch_stride);
        PDEGPU_LAUNCH_CHECK(ctx, "border_fill_kernel");
        for (int q = 0; q < 2; q++) { float *t = cur[q]; cur[q] = nxt[q]; nxt[q] = t; }
    }
    if (cur[0] != sys->x[0]) {            // odd number of sweeps: the result sits in the scratch copy
        for (int q = 0; q < NUNK; q++)
            PDEGPU_CUDA_OK(ctx, cudaMemcpy2DAsync(sys->x[q], (size_t)sys->batch_stride * sizeof(float), cur[q], (size_t)sys->batch_stride * sizeof(float),
                                                  (size_t)npix * sizeof(float), sys->batch, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return PDEGPU_OK;
}

}  // namespace

int relax_stream_point(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega)
{
    static const int use_tiles = getenv("PDEGPU_POINT_TILES") ? atoi(getenv("PDEGPU_POINT_TILES")) : 1;
    if (!use_tiles || sys->nrows < 3 || sys->ncols < 3) return PDEGPU_ERR_UNSUPPORTED;
    switch (sys->family) {
    case PDEGPU_FLOW_ELIN4: return run_point_tiles<PDEGPU_FLOW_ELIN4>(ctx, sys, iter, omega);
    case PDEGPU_FLOW_LLIN4:
    case PDEGPU_FLOW_LLIN8: return run_point_tiles<PDEGPU_FLOW_LLIN4>(ctx, sys, iter, omega);
    case PDEGPU_DISP_LLIN4: return run_point_tiles<PDEGPU_DISP_LLIN4>(ctx, sys, iter, omega);
    case PDEGPU_PDE4:       return run_point_tiles<PDEGPU_PDE4>(ctx, sys, iter, omega);
    default: return PDEGPU_ERR_UNSUPPORTED;                    // PDE8: 4-colour ordering, generation 0
    }
}
