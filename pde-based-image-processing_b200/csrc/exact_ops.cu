// exact_ops.cu -- the non-iterative kernels of the hot path: residual / LHS operators, bilinear
// warp, Simoncelli derivatives, diffusion weights.
//
// These are evaluated ONCE per call in the reference, so the parity bar is per-call
// (BASELINE.json: 1e-5 relative for operators, 1e-6 for warps). We go further: every kernel here
// reproduces the reference's fp32 evaluation order with explicitly rounded intrinsics
// (__fmul_rn/__fadd_rn/__fsub_rn never contract into FMA), so results are BIT-IDENTICAL to the
// reference compiled with `gcc -O2` (SSE2 scalar math, no FMA). All of them are HBM-bound
// streaming kernels, so giving up FMA costs nothing.
#include "pdegpu_internal.cuh"

#define MUL(a, b) __fmul_rn((a), (b))
#define ADD(a, b) __fadd_rn((a), (b))
#define SUB(a, b) __fsub_rn((a), (b))

// =============================================================================================
// Residual r = b - A x  and  LHS A x  for the flow families
//   elin4: Residuals_elin4_2d (opticalflowSolvers.c:269-380), LHS_elin4_2d (:387-496)
//   llin4: Residuals_llin4_2d (:766-916),                     LHS_llin4_2d (:923-1070)
// The reference computes interior pixels and then replicates them into the border (rows first,
// then columns) == every pixel takes the value of the nearest interior pixel.
// Data-term arrays (m,c,d) carry `nframes` channels; x, x0 and the weights have one.
// =============================================================================================
template <bool LATE, bool LHS>
__global__ void __launch_bounds__(256)
flow_operator_kernel(SysView s, int nframes, float *__restrict__ RU, float *__restrict__ RV)
{
    const int nrows = s.nrows, ncols = s.ncols;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int zb = blockIdx.z;                 // batch * nframes + frame
    const int b = zb / nframes, k = zb - b * nframes;
    if (i >= nrows) return;
    const int ic = min(max(i, 1), nrows - 2);
    const int jc = min(max(j, 1), ncols - 2);
    const long long base = (long long)b * s.bstride;
    const long long pos = base + (long long)jc * nrows + ic;
    const long long fo = (long long)k * nrows * ncols;
    const long long posOS = pos + fo;

    const float *X0 = s.x[0], *X1 = s.x[1];
    const float wW = s.w[W_W][pos], wE = s.w[W_E][pos], wN = s.w[W_N][pos], wS = s.w[W_S][pos];
    float nu, nv;
    if (!LATE) {
        nu = ADD(ADD(ADD(MUL(X0[pos - nrows], wW), MUL(X0[pos + nrows], wE)), MUL(X0[pos - 1], wN)), MUL(X0[pos + 1], wS));
        nv = ADD(ADD(ADD(MUL(X1[pos - nrows], wW), MUL(X1[pos + nrows], wE)), MUL(X1[pos - 1], wN)), MUL(X1[pos + 1], wS));
    } else {
        const float *F0 = s.x0[0], *F1 = s.x0[1];
        {
            const float c0 = F0[pos];
            float t0 = SUB(ADD(X0[pos - nrows], F0[pos - nrows]), c0);
            float t1 = SUB(ADD(X0[pos + nrows], F0[pos + nrows]), c0);
            float t2 = SUB(ADD(X0[pos - 1], F0[pos - 1]), c0);
            float t3 = SUB(ADD(X0[pos + 1], F0[pos + 1]), c0);
            t0 = MUL(t0, wW); t1 = MUL(t1, wE); t2 = MUL(t2, wN); t3 = MUL(t3, wS);
            nu = ADD(ADD(t0, t1), ADD(t2, t3));
        }
        {
            const float c0 = F1[pos];
            float t0 = SUB(ADD(X1[pos - nrows], F1[pos - nrows]), c0);
            float t1 = SUB(ADD(X1[pos + nrows], F1[pos + nrows]), c0);
            float t2 = SUB(ADD(X1[pos - 1], F1[pos - 1]), c0);
            float t3 = SUB(ADD(X1[pos + 1], F1[pos + 1]), c0);
            t0 = MUL(t0, wW); t1 = MUL(t1, wE); t2 = MUL(t2, wN); t3 = MUL(t3, wS);
            nv = ADD(ADD(t0, t1), ADD(t2, t3));
        }
    }
    const float sw = ADD(ADD(wW, wE), ADD(wN, wS));
    const float xu = X0[pos], xv = X1[pos];
    const float M = s.m[posOS];
    float ru, rv;
    if (!LHS) {
        const float Cu = s.c[0][posOS], Cv = s.c[1][posOS];
        if (!is_nan(Cu)) ru = SUB(ADD(SUB(Cu, MUL(M, xv)), nu), MUL(ADD(s.d[0][posOS], sw), xu));
        else             ru = SUB(nu, MUL(sw, xu));
        if (!is_nan(Cv)) rv = SUB(ADD(SUB(Cv, MUL(M, xu)), nv), MUL(ADD(s.d[1][posOS], sw), xv));
        else             rv = SUB(nv, MUL(sw, xv));
    } else {
        const float Du = s.d[0][posOS], Dv = s.d[1][posOS];
        if (!is_nan(Du)) ru = ADD(SUB(MUL(M, xv), nu), MUL(ADD(Du, sw), xu));
        else             ru = ADD(-nu, MUL(sw, xu));
        if (!is_nan(Dv)) rv = ADD(SUB(MUL(M, xu), nv), MUL(ADD(Dv, sw), xv));
        else             rv = ADD(-nv, MUL(sw, xv));
    }
    const long long o = base + fo + (long long)j * nrows + i;   // nframes > 1 only with batch == 1 (checked by the launcher)
    RU[o] = ru;
    RV[o] = rv;
}

// Border-fill defects of the late-linearisation operators, reproduced on request so
// that the gateway is bit-compatible with the reference:
//  * Residuals_llin4_2d: the column fill of RV reads `RV->data[i+nrows]` without the frame offset
//    (opticalflowSolvers.c:912) -> for frames k>0 the west border column of RV is frame 0's column 1.
//  * LHS_llin4_2d: the row fill of AV reads `AV->data[pos+1]` with a stale `pos`
//    (opticalflowSolvers.c:1056) -> north border row of AV is 0 in frame 0 and equals
//    AV_frame0(nrows-2, ncols-2) in later frames.
__global__ void llin4_quirk_kernel(int nrows, int ncols, int nframes, float *RV, bool lhs)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    const long long fo = (long long)k * nrows * ncols;
    if (!lhs) {
        if (k == 0 || t >= nrows) return;
        const int ic = min(max(t, 1), nrows - 2);
        RV[fo + t] = RV[(long long)nrows + ic];          // frame 0, column 1 (already border-filled there)
    } else {
        if (t >= ncols) return;
        const float v = (k == 0) ? 0.0f : RV[(long long)(ncols - 2) * nrows + (nrows - 2)];
        RV[fo + (long long)t * nrows] = v;
    }
}

int op_residual(pdegpu_ctx *ctx, const pdegpu_system *sys, int nframes, float *RU, float *RV, bool lhs)
{
    if (!sys || !RU || !RV) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "residual/lhs: null pointer");
    if (sys->nrows < 3 || sys->ncols < 3) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "residual/lhs: need nrows,ncols >= 3");
    if (nframes < 1 || sys->batch < 1) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "residual/lhs: nframes/batch < 1");
    if (nframes > 1 && sys->batch != 1) return pdegpu_set_error(ctx, PDEGPU_ERR_UNSUPPORTED, "residual/lhs: nframes > 1 needs batch == 1");
    const bool late = (sys->family == PDEGPU_FLOW_LLIN4);
    if (!late && sys->family != PDEGPU_FLOW_ELIN4) return pdegpu_set_error(ctx, PDEGPU_ERR_UNSUPPORTED, "residual/lhs: family %d", sys->family);
    SysView v = make_view(sys);
    dim3 block(128), grid((sys->nrows + 127) / 128, sys->ncols, sys->batch * nframes);
    if (grid.z > 65535u) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "residual/lhs: batch*nframes > 65535");
    if (late) { if (lhs) flow_operator_kernel<true, true><<<grid, block, 0, ctx->stream>>>(v, nframes, RU, RV);
                else     flow_operator_kernel<true, false><<<grid, block, 0, ctx->stream>>>(v, nframes, RU, RV); }
    else      { if (lhs) flow_operator_kernel<false, true><<<grid, block, 0, ctx->stream>>>(v, nframes, RU, RV);
                else     flow_operator_kernel<false, false><<<grid, block, 0, ctx->stream>>>(v, nframes, RU, RV); }
    PDEGPU_LAUNCH_CHECK(ctx, "flow_operator_kernel");
    return PDEGPU_OK;
}

int op_llin4_quirks(pdegpu_ctx *ctx, const pdegpu_system *sys, int nframes, float *RU, float *RV, bool lhs)
{
    (void)RU;
    if (sys->batch != 1) return pdegpu_set_error(ctx, PDEGPU_ERR_UNSUPPORTED, "llin4 quirks: batch must be 1");
    const int n = lhs ? sys->ncols : sys->nrows;
    dim3 block(128), grid((n + 127) / 128, nframes);
    llin4_quirk_kernel<<<grid, block, 0, ctx->stream>>>(sys->nrows, sys->ncols, nframes, RV, lhs);
    PDEGPU_LAUNCH_CHECK(ctx, "llin4_quirk_kernel");
    return PDEGPU_OK;
}

// =============================================================================================
// Backward bilinear warp: bilinInterp2 (imageInterpolation.c:44-140)
// =============================================================================================
// (unsigned int)floor(v) as gcc/x86-64 evaluates it: cvttsd2si to 64 bit, keep the low 32 bits;
// NaN / out-of-range give the "integer indefinite" 0x8000000000000000 whose low half is 0.
__device__ __forceinline__ unsigned int floor_to_uint_x86(float v)
{
    const float f = floorf(v);
    long long q;
    if (!(fabsf(f) < 9.2233720368547758e18f)) q = (long long)0x8000000000000000ull;
    else q = (long long)f;
    return (unsigned int)(unsigned long long)q;
}

__global__ void __launch_bounds__(256)
bilin_kernel(float *__restrict__ Iout, const float *__restrict__ Iin,
             const float *__restrict__ X, const float *__restrict__ Y,
             int nrows, int ncols, int nframes, float oob)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= nrows) return;
    const long long pos = (long long)j * nrows + i;
    const long long fsz = (long long)nrows * ncols;
    // batch of independent warps (blockIdx.z): each has its own coordinates and its own nframes images
    X += (long long)blockIdx.z * fsz; Y += (long long)blockIdx.z * fsz;
    Iin += (long long)blockIdx.z * fsz * nframes; Iout += (long long)blockIdx.z * fsz * nframes;
    const float Xv = X[pos], Yv = Y[pos];
    const unsigned int x = floor_to_uint_x86(SUB(Xv, 1.0f));
    const unsigned int y = floor_to_uint_x86(SUB(Yv, 1.0f));
    if (x < (unsigned)ncols && y < (unsigned)nrows) {
        const float xf = SUB(SUB(Xv, 1.0f), (float)x);
        const float yf = SUB(SUB(Yv, 1.0f), (float)y);
        const float w00 = MUL(SUB(1.0f, xf), SUB(1.0f, yf));
        const float w10 = MUL(xf, SUB(1.0f, yf));
        const float w01 = MUL(SUB(1.0f, xf), yf);
        const float w11 = MUL(xf, yf);
        const long long p00 = (long long)nrows * x + y;
        const bool xin = x < (unsigned)(ncols - 1), yin = y < (unsigned)(nrows - 1);
        const long long p10 = p00 + (xin ? nrows : 0);
        const long long p01 = p00 + (yin ? 1 : 0);
        const long long p11 = p00 + ((xin && yin) ? nrows + 1 : 0);
        for (int k = 0; k < nframes; k++) {
            const float *I = Iin + k * fsz;
            Iout[k * fsz + pos] = ADD(ADD(ADD(MUL(w00, I[p00]), MUL(w10, I[p10])), MUL(w01, I[p01])), MUL(w11, I[p11]));
        }
    } else {
        for (int k = 0; k < nframes; k++) Iout[k * fsz + pos] = oob;
    }
}

int op_bilin(pdegpu_ctx *ctx, float *Iout, const float *Iin, const float *X, const float *Y,
             int nrows, int ncols, int nframes, float oob)
{
    return op_bilin_batch(ctx, Iout, Iin, X, Y, nrows, ncols, nframes, 1, oob);
}

int op_bilin_batch(pdegpu_ctx *ctx, float *Iout, const float *Iin, const float *X, const float *Y,
                   int nrows, int ncols, int nframes, int batch, float oob)
{
    if (!Iout || !Iin || !X || !Y) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "bilin_interp: null pointer");
    if (nrows < 1 || ncols < 1 || nframes < 1 || batch < 1) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "bilin_interp: empty array");
    dim3 block(128), grid((nrows + 127) / 128, ncols, batch);
    PDEGPU_PROF(ctx, "bilin_kernel", 4.0 * nrows * ncols * batch * (2.0 + 2.0 * nframes));
    bilin_kernel<<<grid, block, 0, ctx->stream>>>(Iout, Iin, X, Y, nrows, ncols, nframes, oob);
    PDEGPU_LAUNCH_CHECK(ctx, "bilin_kernel");
    return PDEGPU_OK;
}

// =============================================================================================
// Simoncelli 5-tap derivatives: fstSimoncelli_c / sndSimoncelli_c (imageDerivatives.c:309-482)
// built from VerticalConvWO5 (:66), HorizontalConvWO5 (:126), TemporalConvWO2 (:44).
// The reference makes 5 (first order) / 16 (second order) full-frame passes through temporaries;
// here one CTA stages an image tile (+2 px replicated halo) in shared memory and produces all
// 3 / 5 outputs from it. Intermediate images are rounded to fp32 exactly like the reference's
// temporaries, each 5-tap sum is accumulated left to right, so outputs are bit-identical.
// "Vertical" = along i (rows, the contiguous axis), "horizontal" = along j (columns).
// =============================================================================================
constexpr int DT_I = 64, DT_J = 8, DHALO = 2;
constexpr int DP_I = DT_I + 2 * DHALO;     // 68
constexpr int DP_J = DT_J + 2 * DHALO;     // 12

__constant__ float c_smooth[5] = {0.037659f, 0.249724f, 0.439911f, 0.249724f, 0.037659f};          // FstDerivatives5.c:60
__constant__ float c_deriv1[5] = {-0.104550f, -0.292315f, 0.0f, 0.292315f, 0.104550f};             // FstDerivatives5.c:61
__constant__ float c_deriv2[5] = {0.232905f, 0.002668f, -0.471147f, 0.002668f, 0.232905f};         // SndDerivatives5.c:67

__device__ __forceinline__ float tap5(const float *p, int stride, const float *op)
{
    float acc = MUL(p[0], op[0]);
    acc = ADD(acc, MUL(p[stride], op[1]));
    acc = ADD(acc, MUL(p[2 * stride], op[2]));
    acc = ADD(acc, MUL(p[3 * stride], op[3]));
    acc = ADD(acc, MUL(p[4 * stride], op[4]));
    return acc;
}

// smem images are [DP_J][DP_I] (i contiguous). conv along i of src into dst for i-range [2, DP_I-2), all j rows;
// conv along j for j-range [2, DP_J-2), all i.
__device__ __forceinline__ void conv_i(float (*dst)[DP_I], const float (*src)[DP_I], const float *op, int tid, int nthreads)
{
    for (int t = tid; t < DP_J * DT_I; t += nthreads) {
        const int jj = t / DT_I, ii = t - jj * DT_I + DHALO;
        dst[jj][ii] = tap5(&src[jj][ii - 2], 1, op);
    }
}
__device__ __forceinline__ void conv_j(float (*dst)[DP_I], const float (*src)[DP_I], const float *op, int tid, int nthreads)
{
    for (int t = tid; t < DT_J * DP_I; t += nthreads) {
        const int jj = t / DP_I + DHALO, ii = t - (jj - DHALO) * DP_I;
        dst[jj][ii] = tap5(&src[jj - 2][ii], DP_I, op);
    }
}

template <bool SND>
__global__ void __launch_bounds__(256)
deriv5_kernel(float *__restrict__ o0, float *__restrict__ o1, float *__restrict__ o2, float *__restrict__ o3, float *__restrict__ o4,
              const float *__restrict__ It0, const float *__restrict__ It1, int nrows, int ncols)
{
    __shared__ float sI0[DP_J][DP_I];
    __shared__ float sI1[DP_J][DP_I];
    __shared__ float sA[DP_J][DP_I];
    __shared__ float sB[DP_J][DP_I];
    __shared__ float sC[SND ? DP_J : 1][DP_I];
    __shared__ float sD[SND ? DP_J : 1][DP_I];

    const int tid = threadIdx.x, nth = blockDim.x;
    const int i0 = blockIdx.x * DT_I, j0 = blockIdx.y * DT_J;
    const long long fo = (long long)blockIdx.z * nrows * ncols;
    It0 += fo; It1 += fo;

    for (int t = tid; t < DP_J * DP_I; t += nth) {
        const int jj = t / DP_I, ii = t - jj * DP_I;
        const int gi = min(max(i0 + ii - DHALO, 0), nrows - 1);
        const int gj = min(max(j0 + jj - DHALO, 0), ncols - 1);
        const long long p = (long long)gj * nrows + gi;
        sI0[jj][ii] = It0[p];
        sI1[jj][ii] = It1[p];
    }
    __syncthreads();

    const int li = tid % DT_I, ljb = tid / DT_I;          // 256 threads: 64 x 4
    const int gi = i0 + li;

    if (!SND) {
        // Idx = Hd(Vs(It1)), Idy = Vd(Hs(It1)), Idt = 0.5*It0 + (-0.5)*It1   (imageDerivatives.c:372-382)
        conv_i(sA, sI1, c_smooth, tid, nth);              // Vs(I1): valid i in [2,66), all j
        conv_j(sB, sI1, c_smooth, tid, nth);              // Hs(I1): valid j in [2,10), all i
        __syncthreads();
        for (int lj = ljb; lj < DT_J; lj += nth / DT_I) {
            const int gj = j0 + lj;
            if (gi < nrows && gj < ncols) {
                const long long p = fo + (long long)gj * nrows + gi;
                const int ii = li + DHALO, jj = lj + DHALO;
                o0[p] = ADD(MUL(sI0[jj][ii], 0.5f), MUL(sI1[jj][ii], -0.5f));
                o1[p] = tap5(&sA[jj - 2][ii], DP_I, c_deriv1);
                o2[p] = tap5(&sB[jj][ii - 2], 1, c_deriv1);
            }
        }
    } else {
        // (imageDerivatives.c:457-479)
        // Idxt = 0.5*Hd(Vs(It0)) + (-0.5)*Hd(Vs(It1));  Idyt = 0.5*Vd(Hs(It0)) + (-0.5)*Vd(Hs(It1))
        // Idxx = Hd2(Vs(It1)); Idyy = Vd2(Hs(It1)); Idxy = Vd(Hd(It1))
        conv_i(sA, sI0, c_smooth, tid, nth);              // Vs(I0)
        conv_i(sB, sI1, c_smooth, tid, nth);              // Vs(I1)
        conv_j(sC, sI0, c_smooth, tid, nth);              // Hs(I0)
        conv_j(sD, sI1, c_smooth, tid, nth);              // Hs(I1)
        __syncthreads();
        float r_xt[DT_J * DT_I / 256], r_yt[DT_J * DT_I / 256];
        int q = 0;
        for (int lj = ljb; lj < DT_J; lj += nth / DT_I, q++) {
            const int gj = j0 + lj;
            const int ii = li + DHALO, jj = lj + DHALO;
            const float a0 = tap5(&sA[jj - 2][ii], DP_I, c_deriv1), a1 = tap5(&sB[jj - 2][ii], DP_I, c_deriv1);
            const float b0 = tap5(&sC[jj][ii - 2], 1, c_deriv1),    b1 = tap5(&sD[jj][ii - 2], 1, c_deriv1);
            r_xt[q] = ADD(MUL(a0, 0.5f), MUL(a1, -0.5f));
            r_yt[q] = ADD(MUL(b0, 0.5f), MUL(b1, -0.5f));
            if (gi < nrows && gj < ncols) {
                const long long p = fo + (long long)gj * nrows + gi;
                o0[p] = r_xt[q];
                o1[p] = r_yt[q];
                o2[p] = tap5(&sB[jj - 2][ii], DP_I, c_deriv2);
                o3[p] = tap5(&sD[jj][ii - 2], 1, c_deriv2);
            }
        }
        __syncthreads();
        conv_j(sA, sI1, c_deriv1, tid, nth);              // Hd(I1)
        __syncthreads();
        for (int lj = ljb; lj < DT_J; lj += nth / DT_I) {
            const int gj = j0 + lj;
            if (gi < nrows && gj < ncols) {
                const long long p = fo + (long long)gj * nrows + gi;
                o4[p] = tap5(&sA[lj + DHALO][li], 1, c_deriv1);
            }
        }
    }
}

static int deriv_check(pdegpu_ctx *ctx, int nrows, int ncols, int nframes)
{
    // the reference's edge handling indexes pos+3 unconditionally (imageDerivatives.c:84-90): it needs >= 5 per side
    if (nrows < 5 || ncols < 5 || nframes < 1) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "derivatives5: need nrows,ncols >= 5");
    if (nframes > 65535) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "derivatives5: nframes > 65535");
    return PDEGPU_OK;
}

int op_fst(pdegpu_ctx *ctx, float *Idt, float *Idx, float *Idy, const float *It0, const float *It1, int nrows, int ncols, int nframes)
{
    if (!Idt || !Idx || !Idy || !It0 || !It1) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "fst_derivatives5: null pointer");
    int rc = deriv_check(ctx, nrows, ncols, nframes);
    if (rc) return rc;
    dim3 grid((nrows + DT_I - 1) / DT_I, (ncols + DT_J - 1) / DT_J, nframes);
    PDEGPU_PROF(ctx, "deriv5_kernel<fst>", 20.0 * nrows * ncols * nframes);
    deriv5_kernel<false><<<grid, 256, 0, ctx->stream>>>(Idt, Idx, Idy, nullptr, nullptr, It0, It1, nrows, ncols);
    PDEGPU_LAUNCH_CHECK(ctx, "deriv5_kernel<fst>");
    return PDEGPU_OK;
}

int op_snd(pdegpu_ctx *ctx, float *Idxt, float *Idyt, float *Idxx, float *Idyy, float *Idxy, const float *It0, const float *It1, int nrows, int ncols, int nframes)
{
    if (!Idxt || !Idyt || !Idxx || !Idyy || !Idxy || !It0 || !It1) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "snd_derivatives5: null pointer");
    int rc = deriv_check(ctx, nrows, ncols, nframes);
    if (rc) return rc;
    dim3 grid((nrows + DT_I - 1) / DT_I, (ncols + DT_J - 1) / DT_J, nframes);
    PDEGPU_PROF(ctx, "deriv5_kernel<snd>", 28.0 * nrows * ncols * nframes);
    deriv5_kernel<true><<<grid, 256, 0, ctx->stream>>>(Idxt, Idyt, Idxx, Idyy, Idxy, It0, It1, nrows, ncols);
    PDEGPU_LAUNCH_CHECK(ctx, "deriv5_kernel<snd>");
    return PDEGPU_OK;
}

// =============================================================================================
// Diffusion weights: diffWeights6_2D_c (imageDiffusionWeights.c:341-378) = Dver (:32), Dhor (:73),
// Calc_wW/N/E/S (:111/165/224/277). One pass, no temporaries: the two central differences are
// recomputed from D (which stays in L1/L2), the per-frame maximum is taken in registers.
// =============================================================================================
__device__ __forceinline__ float dver(const float *D, int i, int j, int nrows)
{
    const long long c = (long long)j * nrows;
    return ADD(MUL(0.25f, D[c + max(i - 1, 0)]), MUL(-0.25f, D[c + min(i + 1, nrows - 1)]));
}
__device__ __forceinline__ float dhor(const float *D, int i, int j, int nrows, int ncols)
{
    return ADD(MUL(0.25f, D[(long long)max(j - 1, 0) * nrows + i]), MUL(-0.25f, D[(long long)min(j + 1, ncols - 1) * nrows + i]));
}
__device__ __forceinline__ float edge_weight(float t, float eps)
{
    return __fdiv_rn(1.0f, (float)sqrt((double)ADD(t, eps)));
}

__global__ void __launch_bounds__(256)
ddiff_kernel(float *__restrict__ wW, float *__restrict__ wN, float *__restrict__ wE, float *__restrict__ wS,
             const float *__restrict__ D, int nrows, int ncols, int nframes, float eps)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= nrows) return;
    const long long fsz = (long long)nrows * ncols;
    const long long pos = (long long)j * nrows + i;
    float tW = 0.f, tN = 0.f, tE = 0.f, tS = 0.f;
    for (int k = 0; k < nframes; k++) {
        const float *Dk = D + k * fsz;
        const float d = Dk[pos];
        const float v = dver(Dk, i, j, nrows), h = dhor(Dk, i, j, nrows, ncols);
        float A, B, t;
        if (j >= 1) {
            A = SUB(d, Dk[pos - nrows]); B = ADD(v, dver(Dk, i, j - 1, nrows));
            t = ADD(MUL(A, A), MUL(B, B));
            if (k == 0) tW = t; else if (t > tW) tW = t;
        }
        if (i >= 1) {
            A = SUB(d, Dk[pos - 1]); B = ADD(h, dhor(Dk, i - 1, j, nrows, ncols));
            t = ADD(MUL(A, A), MUL(B, B));
            if (k == 0) tN = t; else if (t > tN) tN = t;
        }
        if (j <= ncols - 2) {
            A = SUB(d, Dk[pos + nrows]); B = ADD(v, dver(Dk, i, j + 1, nrows));
            t = ADD(MUL(A, A), MUL(B, B));
            if (k == 0) tE = t; else if (t > tE) tE = t;
        }
        if (i <= nrows - 2) {
            A = SUB(d, Dk[pos + 1]); B = ADD(h, dhor(Dk, i + 1, j, nrows, ncols));
            t = ADD(MUL(A, A), MUL(B, B));
            if (k == 0) tS = t; else if (t > tS) tS = t;
        }
    }
    wW[pos] = (j >= 1) ? edge_weight(tW, eps) : 0.0f;
    wN[pos] = (i >= 1) ? edge_weight(tN, eps) : 0.0f;
    wE[pos] = (j <= ncols - 2) ? edge_weight(tE, eps) : 0.0f;
    wS[pos] = (i <= nrows - 2) ? edge_weight(tS, eps) : 0.0f;
}

int op_ddiff(pdegpu_ctx *ctx, float *wW, float *wN, float *wE, float *wS, const float *D, int nrows, int ncols, int nframes, float eps)
{
    if (!wW || !wN || !wE || !wS || !D) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "ddiff_weights: null pointer");
    if (nrows < 2 || ncols < 2 || nframes < 1) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "ddiff_weights: need nrows,ncols >= 2");
    dim3 block(128), grid((nrows + 127) / 128, ncols);
    PDEGPU_PROF(ctx, "ddiff_kernel", 4.0 * nrows * ncols * (nframes + 4.0));
    ddiff_kernel<<<grid, block, 0, ctx->stream>>>(wW, wN, wE, wS, D, nrows, ncols, nframes, eps);
    PDEGPU_LAUNCH_CHECK(ctx, "ddiff_kernel");
    return PDEGPU_OK;
}
