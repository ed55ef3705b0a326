// driver_ops.cu -- the stencil arithmetic the reference's Matlab drivers do BETWEEN MEX calls
// (SURVEY.md section 8a rows 17-21), as device-resident kernels so that a whole pyramid level runs
// without host round trips:
//
//   op_diff_weights   OPdiffWeights            FlowEminND_llin_2D_v10.m:389-433 (= FlowEminNDFASFMG_elin_2D_v10.m:469-514)
//   llin_terms        robust weights + terms   FlowEminND_llin_2D_v10.m:235-258, 289-299, 323-327
//   elin_terms        FMG smoother's gd/terms  FlowEminNDFASFMG_elin_2D_v10.m:375-396, 421-440
//   disp_sym_terms    symmetric stereo terms   DispEminND_llin_sym_2D.m:172-180, 197-210, 222-225
//   fas_rhs           (R + A)./gd              FlowEminNDFASFMG_elin_2D_v10.m:250-251
//   imfilter          small 2-D correlation, replicate border, optional pre-scale and decimation:
//                     Gaussian pre-smoothing (:99,107-127), rgb2grad (:374-384), lpf pyramid
//                     (FlowEminNDFASFMG...:98,107-110), full-weighting restriction (:199,212-217)
//   imresize          imresize(..., 'bilinear'|'triangle') with antialiasing when shrinking
//                     (FlowEminND_llin_2D_v10.m:107-108,365-366; FlowEminNDFASFMG...:256-257)
//   medfilt3          medfilt2(., [3 3], 'symmetric')   (FlowEminND_llin_2D_v10.m:354-355)
//   axpby, warp_coords   U+dU, X+U / Y+V (:202,222,228,354)
//
// Arithmetic follows the .m code: single where Matlab computes in single (separate roundings, no FMA
// contraction: __fmul_rn/__fadd_rn), double where the driver casts to double or the toolbox accumulates in
// double. The toolbox functions themselves (imresize, imfilter, medfilt2) are restated from their
// documented behaviour: parity unpinned, checked against oracle/matlab_steps.py.
// All kernels are streaming, HBM-bound, one thread per output pixel (fast axis = Matlab row index).
#include "pdegpu_internal.cuh"
#include "driver_formulas.cuh"

namespace {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ int wrapi(int v, int n) { return v < 0 ? v + n : (v >= n ? v - n : v); }
__device__ __forceinline__ int mirrori(int v, int n)            // 'symmetric' padding / imresize's index mirror
{
    const int p = 2 * n;
    int m = v % p;
    if (m < 0) m += p;
    return m < n ? m : p - 1 - m;
}

// ---------------------------------------------------------------------------------------------
// OPdiffWeights (double precision, circshift wrap) -> 4 single fields
// ---------------------------------------------------------------------------------------------
// (dU, dV given: the weights of U + dU, V + dV -- the sums are formed on the fly, FlowEminND_llin_2D_v10.m:321)
__global__ void __launch_bounds__(256)
op_diff_weights_kernel(float *__restrict__ wW, float *__restrict__ wN, float *__restrict__ wS, float *__restrict__ wE,
                       const float *__restrict__ U, const float *__restrict__ V, const float *__restrict__ dU, const float *__restrict__ dV,
                       int nr, int nc, long long stride)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= nr) return;
    const long long base = (long long)blockIdx.z * stride;
    const OpdiffSrc src = {U + base, V + base, dU ? dU + base : nullptr, dV ? dV + base : nullptr};
    // every edge once, stored on both of its sides (opdiff_east_south_at): each wW / wN element has exactly one writer
    float e, s;
    opdiff_east_south_at(src, i, j, nr, nc, e, s);
    const long long col = base + (long long)j * nr, ecol = base + (long long)df_wrapi(j + 1, nc) * nr;
    wE[col + i] = e;
    wS[col + i] = s;
    wW[ecol + i] = e;
    wN[col + df_wrapi(i + 1, nr)] = s;
}

// ---------------------------------------------------------------------------------------------
// single-precision helpers that round like Matlab / numpy (one rounding per operation)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float mulf(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float addf(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float subf(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float sqf(float a) { return __fmul_rn(a, a); }

__global__ void __launch_bounds__(256)
llin_terms_kernel(const LlinTermsArgs a)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.npix) return;
    const long long b = blockIdx.y;
    float M, Cu, Cv, Du, Dv;
    llin_terms_at(a, b, t, a.dU[b * a.stride + t], a.dV[b * a.stride + t], M, Cu, Cv, Du, Dv);
    const long long o = b * a.stride + t;
    a.out[0][o] = M; a.out[1][o] = Cu; a.out[2][o] = Cv; a.out[3][o] = Du; a.out[4][o] = Dv;
}

struct ElinTermsArgs {
    const float *der[8];      // Idt, Idx, Idy, Idxt, Idyt, Idxx, Idyy, Idxy   (channels)
    const float *coef[5];     // M, Cu, Cv, Du, Dv                              (channels)
    const float *U, *V;
    float *gd;                // may be null; channels
    float *out[5];            // summed: 1 channel; else channels
    int channels, summed;
    float b1, b2, alpha;
    long long npix;
};

__global__ void __launch_bounds__(256)
elin_terms_kernel(const ElinTermsArgs a)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.npix) return;
    const float u = a.U[t], v = a.V[t];
    const float fac = a.summed ? (float)((double)a.channels * (double)a.alpha) : a.alpha;   // channels*param.alpha is a double product, cast to single when it meets the single array
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < a.channels; c++) {
        const long long p = (long long)c * a.npix + t;
        const float Idt = a.der[0][p], Idx = a.der[1][p], Idy = a.der[2][p], Idxt = a.der[3][p], Idyt = a.der[4][p];
        const float Idxx = a.der[5][p], Idyy = a.der[6][p], Idxy = a.der[7][p];
        const float r0 = subf(subf(Idt, mulf(Idx, u)), mulf(Idy, v));
        const float r1 = subf(subf(Idxt, mulf(Idxx, u)), mulf(Idxy, v));
        const float r2 = subf(subf(Idyt, mulf(Idxy, u)), mulf(Idyy, v));
        const float op = addf(mulf(a.b1, sqf(r0)), mulf(a.b2, addf(sqf(r1), sqf(r2))));
        const float g = __fdiv_rn(1.0f, mulf(fac, __fsqrt_rn(addf(op, 0.00001f))));
        if (a.gd) a.gd[p] = g;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const float tk = mulf(a.coef[k][p], g);
            if (a.summed) acc[k] = c == 0 ? tk : addf(acc[k], tk);
            else a.out[k][p] = tk;
        }
    }
    if (a.summed) {
#pragma unroll
        for (int k = 0; k < 5; k++) a.out[k][t] = acc[k];
    }
}

struct DispSymArgs {
    const float *d[6];        // Idt, Idx, Idxt, Idyt, Idxx, Idxy  (channels)
    const float *dU, *Udt, *Udx;
    float *CuG, *DuG;
    int channels;
    float b1, b2, alpha, gsym_num, srdiff2;    // gsym_num = channels*beta/alpha, srdiff2 = srDiff^2 (both rounded to single like Matlab)
    long long npix;
};

__global__ void __launch_bounds__(256)
disp_sym_terms_kernel(const DispSymArgs a)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.npix) return;
    const float du = a.dU[t], udt = a.Udt[t], udx = a.Udx[t];
    float cu = 0.f, dd = 0.f;
    for (int c = 0; c < a.channels; c++) {
        const long long p = (long long)c * a.npix + t;
        const float Idt = a.d[0][p], Idx = a.d[1][p], Idxt = a.d[2][p], Idyt = a.d[3][p], Idxx = a.d[4][p], Idxy = a.d[5][p];
        const float CuD = addf(mulf(mulf(a.b1, Idt), Idx), mulf(a.b2, addf(mulf(Idxt, Idxx), mulf(Idyt, Idxy))));
        const float DuD = addf(mulf(mulf(a.b1, Idx), Idx), mulf(a.b2, addf(mulf(Idxx, Idxx), mulf(Idxy, Idxy))));
        const float op = addf(mulf(a.b1, sqf(subf(Idt, mulf(Idx, du)))),
                              mulf(a.b2, addf(sqf(subf(Idxt, mulf(Idxx, du))), sqf(subf(Idyt, mulf(Idxy, du))))));
        const float g = __fdiv_rn(1.0f, mulf(a.alpha, __fsqrt_rn(addf(op, 0.00001f))));
        const float tc = mulf(g, CuD), td = mulf(g, DuD);
        cu = c == 0 ? tc : addf(cu, tc);
        dd = c == 0 ? td : addf(dd, td);
    }
    const float CuS = mulf(udt, addf(1.0f, udx));
    const float DuS = addf(addf(addf(1.0f, udx), udx), mulf(udx, udx));
    const float sn = sqf(addf(addf(du, udt), mulf(udx, du)));
    const float gs = __fdiv_rn(a.gsym_num, addf(1.0f, __fdiv_rn(sn, a.srdiff2)));
    a.CuG[t] = addf(cu, mulf(-gs, CuS));
    a.DuG[t] = addf(dd, mulf(gs, DuS));
}

__global__ void __launch_bounds__(256)
fas_rhs_kernel(float *__restrict__ f, const float *__restrict__ R, const float *__restrict__ A, const float *__restrict__ gd, long long n)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) f[t] = __fdiv_rn(addf(R[t], A[t]), gd[t]);
}

// ---------------------------------------------------------------------------------------------
// imfilter: correlation with a small kernel (<= 5x5), replicate border, double accumulation,
// optional single-precision pre-scale of the input and decimation of the output
// ---------------------------------------------------------------------------------------------
struct FilterArgs {
    double h[25];             // kernel, column-major like Matlab: h[b*kr + a]
    int kr, kc;
    int step;                 // output (i,j) = filtered (step*i, step*j)
    float prescale;           // input is multiplied by this in single first (1 = no-op)
    int use_prescale;
    int nr, nc, onr, onc;
    long long istride, ostride;   // per plane (blockIdx.z)
};

__global__ void __launch_bounds__(256)
imfilter_kernel(float *__restrict__ out, const float *__restrict__ in, const FilterArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= a.onr) return;
    in += (long long)blockIdx.z * a.istride;
    const int ci = a.step * i, cj = a.step * j, cr = (a.kr - 1) >> 1, cc = (a.kc - 1) >> 1;
    double acc = 0.0;
    for (int ka = 0; ka < a.kr; ka++) {
        const int ii = clampi(ci + ka - cr, 0, a.nr - 1);
        for (int kb = 0; kb < a.kc; kb++) {
            const double h = a.h[kb * a.kr + ka];
            if (h == 0.0) continue;
            const int jj = clampi(cj + kb - cc, 0, a.nc - 1);
            float v = in[(long long)jj * a.nr + ii];
            if (a.use_prescale) v = mulf(v, a.prescale);
            acc += h * (double)v;
        }
    }
    out[(long long)blockIdx.z * a.ostride + (long long)j * a.onr + i] = (float)acc;
}

// ---------------------------------------------------------------------------------------------
// imresize along one dimension, triangle kernel (see oracle/matlab_steps.py imresize_contributions)
// ---------------------------------------------------------------------------------------------
struct ResizeArgs {
    int dim;                  // 0: along rows (fast axis), 1: along columns
    int in_len, out_len, other;   // `other` = length of the untouched dimension
    double scale;
    int antialias;
    int cubic;                // 0: triangle (bilinear), 1: Matlab's cubic convolution kernel (bicubic)
    long long istride, ostride;
};

// The contributions of an output coordinate (first input index `left`, P normalised weights) depend on that coordinate
// only -- not on the position along the other dimension, the plane or the pair of a batch. A small kernel tabulates them
// (double-precision kernel evaluations and one division per tap: the expensive part), the resize kernel proper is then
// P multiply-adds per pixel. Same expressions, same order of accumulation as the one-kernel form it replaces.
__global__ void __launch_bounds__(128)
imresize_table_kernel(double *__restrict__ wtab, int *__restrict__ ltab, const ResizeArgs a, const int P)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;      // 0-based output coordinate along a.dim
    if (t >= a.out_len) return;
    const bool shrink = a.scale < 1.0 && a.antialias;
    const double kw0 = a.cubic ? 4.0 : 2.0;
    const double kw = shrink ? kw0 / a.scale : kw0;
    const int x = t + 1;                                      // 1-based output coordinate
    const double u = (double)x / a.scale + 0.5 * (1.0 - 1.0 / a.scale);
    const int left = (int)floor(u - kw / 2.0);              // P = ceil(kw) + 2 taps (computed once, on the host)
    auto h = [&](double d) -> double {
        if (shrink) d *= a.scale;
        const double ax = fabs(d);
        double f;
        if (a.cubic) {
            const double ax2 = ax * ax, ax3 = ax2 * ax;
            f = (1.5 * ax3 - 2.5 * ax2 + 1.0) * (ax <= 1.0 ? 1.0 : 0.0)
              + (-0.5 * ax3 + 2.5 * ax2 - 4.0 * ax + 2.0) * ((1.0 < ax && ax <= 2.0) ? 1.0 : 0.0);
        } else f = fmax(0.0, 1.0 - ax);
        return shrink ? a.scale * f : f;
    };
    double wsum = 0.0;
    for (int p = 0; p < P; p++) wsum += h(u - (double)(left + p));
    for (int p = 0; p < P; p++) wtab[(long long)p * a.out_len + t] = h(u - (double)(left + p)) / wsum;   // tap-major: coalesced along i
    ltab[t] = left;
}

__global__ void __launch_bounds__(256)
imresize_kernel(float *__restrict__ out, const float *__restrict__ in, const ResizeArgs a,
                const double *__restrict__ wtab, const int *__restrict__ ltab, int P)
{
    // output pixel (i, j) of an onr x onc plane
    const int onr = a.dim == 0 ? a.out_len : a.other, inr = a.dim == 0 ? a.in_len : a.other;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= onr) return;
    in += (long long)blockIdx.z * a.istride;
    const int t = a.dim == 0 ? i : j;
    const int left = ltab[t];
    double acc = 0.0;
    for (int p = 0; p < P; p++) {
        const int src = mirrori(left + p - 1, a.in_len);      // left + p: 1-based input coordinate
        const float v = a.dim == 0 ? in[(long long)j * inr + src] : in[(long long)src * inr + i];
        acc += wtab[(long long)p * a.out_len + t] * (double)v;
    }
    out[(long long)blockIdx.z * a.ostride + (long long)j * onr + i] = (float)acc;
}

// ---------------------------------------------------------------------------------------------
// medfilt2(., [3 3], 'symmetric')
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cswap(float &a, float &b) { const float lo = fminf(a, b), hi = fmaxf(a, b); a = lo; b = hi; }

__global__ void __launch_bounds__(256)
medfilt3_kernel(float *__restrict__ out, const float *__restrict__ in, int nr, int nc, long long stride)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= nr) return;
    in += (long long)blockIdx.z * stride;
    float v[9];
#pragma unroll
    for (int b = 0; b < 3; b++)
#pragma unroll
        for (int a = 0; a < 3; a++)
            v[b * 3 + a] = in[(long long)mirrori(j + b - 1, nc) * nr + mirrori(i + a - 1, nr)];
    // median-of-9 exchange network
    cswap(v[1], v[2]); cswap(v[4], v[5]); cswap(v[7], v[8]);
    cswap(v[0], v[1]); cswap(v[3], v[4]); cswap(v[6], v[7]);
    cswap(v[1], v[2]); cswap(v[4], v[5]); cswap(v[7], v[8]);
    cswap(v[0], v[3]); cswap(v[5], v[8]); cswap(v[4], v[7]);
    cswap(v[3], v[6]); cswap(v[1], v[4]); cswap(v[2], v[5]);
    cswap(v[4], v[7]); cswap(v[4], v[2]); cswap(v[6], v[4]);
    cswap(v[4], v[2]);
    out[(long long)blockIdx.z * stride + (long long)j * nr + i] = v[4];
}

// out = a*x + b*y (single, one rounding per operation); y may be null (b ignored). The pipelines call it in place
// (out == x or out == y): no __restrict__ on purpose.
__global__ void __launch_bounds__(256)
axpby_kernel(float *out, float a, const float *x, float b, const float *y, long long n)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float ax = a == 1.0f ? x[t] : mulf(a, x[t]);
    out[t] = y ? addf(ax, b == 1.0f ? y[t] : mulf(b, y[t])) : ax;
}

// out = x / d in single (Iin = single(Iin)./255, FlowEminND_llin_2D_v10.m:73)
__global__ void __launch_bounds__(256)
div_kernel(float *__restrict__ out, const float *__restrict__ x, float d, long long n)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[t] = __fdiv_rn(x[t], d);
}

// X = (j+1) + U, Y = (i+1) + V  (meshgrid coordinates plus flow, in single: FlowEminND_llin_2D_v10.m:202,222)
__global__ void __launch_bounds__(256)
warp_coords_kernel(float *__restrict__ X, float *__restrict__ Y, const float *__restrict__ U, const float *__restrict__ V, int nr, int nc, long long stride)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= nr) return;
    const long long p = (long long)blockIdx.z * stride + (long long)j * nr + i;
    X[p] = U ? addf((float)(j + 1), U[p]) : (float)(j + 1);
    Y[p] = V ? addf((float)(i + 1), V[p]) : (float)(i + 1);
}

// out = interp2(X, Y, vals, X + shift, Y) on the grid's own rows (DispEminND_llin_sym_2D.m:144-145): linear interpolation
// along j at the 1-based position (j+1) + shift, NaN outside [1, ncols]; double arithmetic
__global__ void __launch_bounds__(256)
interp_rows_kernel(float *__restrict__ out, const float *__restrict__ vals, const float *__restrict__ shift, int nr, int nc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= nr) return;
    const long long p = (long long)j * nr + i;
    const double xq = (double)(j + 1) + (double)shift[p];
    float r = nanf("");
    if (xq >= 1.0 && xq <= (double)nc) {
        int j0 = (int)floor(xq);
        if (j0 > nc - 1) j0 = nc - 1;
        const double t = xq - (double)j0;
        const double v0 = (double)vals[(long long)(j0 - 1) * nr + i], v1 = (double)vals[(long long)j0 * nr + i];
        r = (float)(v0 * (1.0 - t) + v1 * t);
    }
    out[p] = r;
}

// class uint8 after a toolbox call: round half away from zero, saturate to 0..255
__global__ void __launch_bounds__(256)
round_uint8_kernel(float *__restrict__ x, long long n)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double v = floor((double)x[t] + 0.5);
    x[t] = (float)(v < 0.0 ? 0.0 : v > 255.0 ? 255.0 : v);
}

inline dim3 grid2(int nr, int nc, int planes) { return dim3((nr + 255) / 256, nc, planes); }

}  // namespace

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
int op_opdiff(pdegpu_ctx *ctx, float *wW, float *wN, float *wS, float *wE, const float *U, const float *V,
              int nr, int nc, int batch, long long stride)
{
    return op_opdiff_sum(ctx, wW, wN, wS, wE, U, V, nullptr, nullptr, nr, nc, batch, stride);
}

int op_opdiff_sum(pdegpu_ctx *ctx, float *wW, float *wN, float *wS, float *wE, const float *U, const float *V, const float *dU, const float *dV,
                  int nr, int nc, int batch, long long stride)
{
    PDEGPU_PROF(ctx, "op_diff_weights_kernel", (dU ? 32.0 : 24.0) * nr * nc * batch);
    op_diff_weights_kernel<<<grid2(nr, nc, batch), 256, 0, ctx->stream>>>(wW, wN, wS, wE, U, V, dU, dV, nr, nc, stride);
    PDEGPU_LAUNCH_CHECK(ctx, "op_diff_weights_kernel");
    return PDEGPU_OK;
}

int op_llin_terms(pdegpu_ctx *ctx, const pdegpu_llin_terms *t)
{
    LlinTermsArgs a;
    for (int k = 0; k < 3; k++) a.d1[k] = t->d1[k];
    for (int k = 0; k < 5; k++) { a.d2[k] = t->d2[k]; a.out[k] = t->out[k]; }
    a.dU = t->dU; a.dV = t->dV;
    a.c1 = t->channels1; a.c2 = t->channels2; a.gradmag = t->gradmag;
    a.b1 = t->b1; a.b2 = t->b2; a.alpha = t->alpha;
    a.npix = (long long)t->nrows * t->ncols;
    a.stride1 = t->batch_stride1; a.stride2 = t->batch_stride2; a.stride = t->batch_stride;
    PDEGPU_PROF(ctx, "llin_terms_kernel", 4.0 * a.npix * t->batch * (3.0 * a.c1 + (a.gradmag ? 5.0 : 3.0) * a.c2 + 7.0));
    llin_terms_kernel<<<dim3((unsigned)((a.npix + 255) / 256), t->batch), 256, 0, ctx->stream>>>(a);
    PDEGPU_LAUNCH_CHECK(ctx, "llin_terms_kernel");
    return PDEGPU_OK;
}

int op_elin_terms(pdegpu_ctx *ctx, const pdegpu_elin_terms *t)
{
    ElinTermsArgs a;
    for (int k = 0; k < 8; k++) a.der[k] = t->der[k];
    for (int k = 0; k < 5; k++) { a.coef[k] = t->coef[k]; a.out[k] = t->out[k]; }
    a.U = t->U; a.V = t->V; a.gd = t->gd;
    a.channels = t->channels; a.summed = t->summed;
    a.b1 = t->b1; a.b2 = t->b2; a.alpha = t->alpha;
    a.npix = (long long)t->nrows * t->ncols;
    PDEGPU_PROF(ctx, "elin_terms_kernel", 4.0 * a.npix * (13.0 * a.channels + 2 + (a.summed ? 5.0 : 6.0 * a.channels)));
    elin_terms_kernel<<<(unsigned)((a.npix + 255) / 256), 256, 0, ctx->stream>>>(a);
    PDEGPU_LAUNCH_CHECK(ctx, "elin_terms_kernel");
    return PDEGPU_OK;
}

int op_disp_sym_terms(pdegpu_ctx *ctx, const pdegpu_disp_sym_terms *t)
{
    DispSymArgs a;
    for (int k = 0; k < 6; k++) a.d[k] = t->d[k];
    a.dU = t->dU; a.Udt = t->Udt; a.Udx = t->Udx; a.CuG = t->CuG; a.DuG = t->DuG;
    a.channels = t->channels;
    a.b1 = t->b1; a.b2 = t->b2; a.alpha = t->alpha;
    a.gsym_num = (float)((double)t->channels * (double)t->beta / (double)t->alpha_d);
    a.srdiff2 = (float)(t->srdiff * t->srdiff);
    a.npix = (long long)t->nrows * t->ncols;
    PDEGPU_PROF(ctx, "disp_sym_terms_kernel", 4.0 * a.npix * (6.0 * a.channels + 5));
    disp_sym_terms_kernel<<<(unsigned)((a.npix + 255) / 256), 256, 0, ctx->stream>>>(a);
    PDEGPU_LAUNCH_CHECK(ctx, "disp_sym_terms_kernel");
    return PDEGPU_OK;
}

int op_fas_rhs(pdegpu_ctx *ctx, float *f, const float *R, const float *A, const float *gd, long long n)
{
    PDEGPU_PROF(ctx, "fas_rhs_kernel", 16.0 * n);
    fas_rhs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(f, R, A, gd, n);
    PDEGPU_LAUNCH_CHECK(ctx, "fas_rhs_kernel");
    return PDEGPU_OK;
}

int op_imfilter(pdegpu_ctx *ctx, float *out, const float *in, int nr, int nc, int planes, long long istride, long long ostride,
                const double *h, int kr, int kc, int step, float prescale)
{
    if (kr < 1 || kc < 1 || kr * kc > 25 || !(kr & 1) || !(kc & 1) || step < 1) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "imfilter: kernel must be odd-sized, at most 25 taps");
    FilterArgs a;
    for (int k = 0; k < kr * kc; k++) a.h[k] = h[k];
    a.kr = kr; a.kc = kc; a.step = step; a.prescale = prescale; a.use_prescale = prescale != 1.0f;
    a.nr = nr; a.nc = nc; a.onr = (nr + step - 1) / step; a.onc = (nc + step - 1) / step;
    a.istride = istride; a.ostride = ostride;
    PDEGPU_PROF(ctx, "imfilter_kernel", 4.0 * ((double)nr * nc + (double)a.onr * a.onc) * planes);
    imfilter_kernel<<<grid2(a.onr, a.onc, planes), 256, 0, ctx->stream>>>(out, in, a);
    PDEGPU_LAUNCH_CHECK(ctx, "imfilter_kernel");
    return PDEGPU_OK;
}

int op_imresize_dim(pdegpu_ctx *ctx, float *out, const float *in, int dim, int in_len, int out_len, int other, double scale,
                    int antialias, int planes, long long istride, long long ostride, int cubic)
{
    ResizeArgs a;
    a.cubic = cubic;
    a.dim = dim; a.in_len = in_len; a.out_len = out_len; a.other = other; a.scale = scale; a.antialias = antialias;
    a.istride = istride; a.ostride = ostride;
    const int onr = dim == 0 ? out_len : other, onc = dim == 0 ? other : out_len;
    const double kw0 = cubic ? 4.0 : 2.0;
    const int P = (int)ceil((scale < 1.0 && antialias) ? kw0 / scale : kw0) + 2;
    // table of contributions in the context's scratch (stream-ordered with the sweeps that use the same buffer)
    const size_t wbytes = ((size_t)out_len * P * sizeof(double) + 255) & ~(size_t)255;
    int rc = pdegpu_scratch_reserve(ctx, wbytes + (size_t)out_len * sizeof(int));
    if (rc) return rc;
    double *wtab = (double *)ctx->scratch;
    int *ltab = (int *)(ctx->scratch + wbytes);
    PDEGPU_PROF(ctx, "imresize_table_kernel", 0.0);
    imresize_table_kernel<<<(out_len + 127) / 128, 128, 0, ctx->stream>>>(wtab, ltab, a, P);
    PDEGPU_LAUNCH_CHECK(ctx, "imresize_table_kernel");
    PDEGPU_PROF(ctx, "imresize_kernel", 4.0 * ((double)in_len + out_len) * other * planes);
    imresize_kernel<<<grid2(onr, onc, planes), 256, 0, ctx->stream>>>(out, in, a, wtab, ltab, P);
    PDEGPU_LAUNCH_CHECK(ctx, "imresize_kernel");
    return PDEGPU_OK;
}

int op_medfilt3(pdegpu_ctx *ctx, float *out, const float *in, int nr, int nc, int planes, long long stride)
{
    PDEGPU_PROF(ctx, "medfilt3_kernel", 8.0 * nr * nc * planes);
    medfilt3_kernel<<<grid2(nr, nc, planes), 256, 0, ctx->stream>>>(out, in, nr, nc, stride);
    PDEGPU_LAUNCH_CHECK(ctx, "medfilt3_kernel");
    return PDEGPU_OK;
}

int op_axpby(pdegpu_ctx *ctx, float *out, float a, const float *x, float b, const float *y, long long n)
{
    PDEGPU_PROF(ctx, "axpby_kernel", (y ? 12.0 : 8.0) * n);
    axpby_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(out, a, x, b, y, n);
    PDEGPU_LAUNCH_CHECK(ctx, "axpby_kernel");
    return PDEGPU_OK;
}

int op_warp_coords(pdegpu_ctx *ctx, float *X, float *Y, const float *U, const float *V, int nr, int nc, int batch, long long stride)
{
    PDEGPU_PROF(ctx, "warp_coords_kernel", 16.0 * nr * nc * batch);
    warp_coords_kernel<<<grid2(nr, nc, batch), 256, 0, ctx->stream>>>(X, Y, U, V, nr, nc, stride);
    PDEGPU_LAUNCH_CHECK(ctx, "warp_coords_kernel");
    return PDEGPU_OK;
}

int op_interp_rows(pdegpu_ctx *ctx, float *out, const float *vals, const float *shift, int nr, int nc)
{
    PDEGPU_PROF(ctx, "interp_rows_kernel", 16.0 * nr * nc);
    interp_rows_kernel<<<grid2(nr, nc, 1), 256, 0, ctx->stream>>>(out, vals, shift, nr, nc);
    PDEGPU_LAUNCH_CHECK(ctx, "interp_rows_kernel");
    return PDEGPU_OK;
}

int op_round_uint8(pdegpu_ctx *ctx, float *x, long long n)
{
    PDEGPU_PROF(ctx, "round_uint8_kernel", 8.0 * n);
    round_uint8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(x, n);
    PDEGPU_LAUNCH_CHECK(ctx, "round_uint8_kernel");
    return PDEGPU_OK;
}

int op_axpby_div(pdegpu_ctx *ctx, float *out, const float *x, float d, long long n)
{
    PDEGPU_PROF(ctx, "div_kernel", 8.0 * n);
    div_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(out, x, d, n);
    PDEGPU_LAUNCH_CHECK(ctx, "div_kernel");
    return PDEGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// ADdiffWeights (matlab/denoising/TVdenoise8.m:119-231): anisotropic diffusion tensor -> 8 edge weights,
// with lambda = the `quantile` order statistic of the non-zero squared gradient norms (:196-204), found by
// an 8-pass radix select on the device; optionally the TV data terms PsiData/TRACE/B (:83-85) in the same pass.
// Double precision like the .m code; results cast to single as at the MEX call (:87-100).
// ---------------------------------------------------------------------------------------------
namespace {

struct SelState {
    unsigned long long prefix, k, count;
    double lambda;
    unsigned int hist[256];
};

// Alvarez derivatives (imfilter(D, O, 'replicate', 'conv')) of every frame, keep the frame of largest gradient norm
__global__ void __launch_bounds__(256)
ad_grad_kernel(double *__restrict__ mx, double *__restrict__ my, double *__restrict__ nrm, const float *__restrict__ D,
               int nr, int nc, int frames, SelState *st)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= nr) return;
    const double s2 = sqrt(2.0), den = 4.0 + sqrt(8.0);
    // O_dx(a,b) (row a, col b), O_dy likewise; 'conv' = correlation with the flipped kernel
    const double odx[3][3] = {{1 / den, 0, -1 / den}, {s2 / den, 0, -s2 / den}, {1 / den, 0, -1 / den}};
    const double ody[3][3] = {{1 / den, s2 / den, 1 / den}, {0, 0, 0}, {-1 / den, -s2 / den, -1 / den}};
    double bx = 0, by = 0, bn = -1;
    for (int f = 0; f < frames; f++) {
        const float *F = D + (long long)f * nr * nc;
        double gx = 0, gy = 0;
        for (int a = 0; a < 3; a++) {                         // same accumulation order as oracle imfilter: rows, then columns, of the flipped kernel
            const int ii = clampi(i + a - 1, 0, nr - 1);
            for (int b = 0; b < 3; b++) {
                const int jj = clampi(j + b - 1, 0, nc - 1);
                const double v = (double)F[(long long)jj * nr + ii];
                const double hx = odx[2 - a][2 - b], hy = ody[2 - a][2 - b];
                if (hx != 0.0) gx = __dadd_rn(gx, __dmul_rn(hx, v));      // no FMA: lambda is an order statistic of these values
                if (hy != 0.0) gy = __dadd_rn(gy, __dmul_rn(hy, v));
            }
        }
        const double n2 = __dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy));
        if (n2 > bn) { bn = n2; bx = gx; by = gy; }            // first maximum, like Matlab's max
    }
    const long long p = (long long)j * nr + i;
    const double n2 = __dadd_rn(__dmul_rn(bx, bx), __dmul_rn(by, by));
    mx[p] = bx; my[p] = by; nrm[p] = n2;
    if (n2 != 0.0) atomicAdd(&st->count, 1ull);
}

__global__ void sel_init_kernel(SelState *st, double quantile)
{
    if (st->count == 0) { st->k = 0; st->lambda = 1.0; }
    else {
        double r = floor((double)st->count * quantile + 2.220446049250313e-16 + 0.5);     // Matlab round()
        if (r < 1) r = 1;
        if (r > (double)st->count) r = (double)st->count;
        st->k = (unsigned long long)r;
    }
    st->prefix = 0;
    for (int b = 0; b < 256; b++) st->hist[b] = 0;
}

__global__ void __launch_bounds__(256)
sel_hist_kernel(SelState *st, const double *__restrict__ nrm, long long n, int shift)
{
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long prefix = st->prefix;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const unsigned long long key = (unsigned long long)__double_as_longlong(nrm[t]);
        if (key == 0) continue;
        if (shift < 56 && (key >> (shift + 8)) != (prefix >> (shift + 8))) continue;
        atomicAdd(&h[(key >> shift) & 255], 1u);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], h[threadIdx.x]);
}

__global__ void sel_pick_kernel(SelState *st, int shift)
{
    if (st->k == 0) return;
    unsigned long long cum = 0;
    int b = 0;
    for (; b < 256; b++) {
        if (cum + st->hist[b] >= st->k) break;
        cum += st->hist[b];
    }
    st->k -= cum;
    st->prefix |= (unsigned long long)b << shift;
    for (int q = 0; q < 256; q++) st->hist[q] = 0;
    if (shift == 0) st->lambda = __longlong_as_double((long long)st->prefix);
}

struct AdArgs {
    float *w[8];              // W, NW, N, NE, E, SE, S, SW  (scaled by `scale`)
    float *TRACE, *B;         // optional (frames each)
    const float *Iout, *Iin;  // for TRACE/B
    const double *mx, *my, *nrm;
    const SelState *st;
    int nr, nc, frames, wframes;
    double scale;
};

__global__ void __launch_bounds__(256)
ad_weights_kernel(const AdArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= a.nr) return;
    const int nr = a.nr, nc = a.nc;
    const double lam = a.st->lambda;
    auto tensor = [&](int ii, int jj, double &dyy, double &dxx, double &dxy) {
        const long long p = (long long)wrapi(jj, nc) * nr + wrapi(ii, nr);     // circshift wraps; wrapped values are zeroed below
        const double x = a.mx[p], y = a.my[p];
        const double mul = 1.0 / (a.nrm[p] + 2 * lam);
        dyy = mul * (y * y + lam); dxx = mul * (x * x + lam); dxy = -mul * (x * y);
    };
    double cyy, cxx, cxy, tyy, txx, txy;
    tensor(i, j, cyy, cxx, cxy);
    double w[8];
    tensor(i, j - 1, tyy, txx, txy);      w[0] = j == 0 ? 0.0 : 0.5 * (cyy + tyy);                                // W
    tensor(i - 1, j - 1, tyy, txx, txy);  w[1] = (j == 0 || i == 0) ? 0.0 : 0.25 * (cxy + txy);                   // NW
    tensor(i - 1, j, tyy, txx, txy);      w[2] = i == 0 ? 0.0 : 0.5 * (cxx + txx);                                // N
    tensor(i - 1, j + 1, tyy, txx, txy);  w[3] = (j == nc - 1 || i == 0) ? 0.0 : -0.25 * (cxy + txy);             // NE
    tensor(i, j + 1, tyy, txx, txy);      w[4] = j == nc - 1 ? 0.0 : 0.5 * (cyy + tyy);                           // E
    tensor(i + 1, j + 1, tyy, txx, txy);  w[5] = (j == nc - 1 || i == nr - 1) ? 0.0 : 0.25 * (cxy + txy);         // SE
    tensor(i + 1, j, tyy, txx, txy);      w[6] = i == nr - 1 ? 0.0 : 0.5 * (cxx + txx);                           // S
    tensor(i + 1, j - 1, tyy, txx, txy);  w[7] = (i == nr - 1 || j == 0) ? 0.0 : -0.25 * (cxy + txy);             // SW
    const long long p = (long long)j * nr + i;
    // ADdiffWeights repmat's its weights over the frames (:221-230): one copy per frame when wframes > 1
    for (int f = 0; f < a.wframes; f++) {
#pragma unroll
        for (int k = 0; k < 8; k++) a.w[k][(long long)f * nr * nc + p] = (float)(a.scale * w[k]);
    }
    if (a.TRACE) {
        // TVdenoise8.m:83-85 with single images (runme.m:118,144): PsiData and B in single, TRACE = single + double -> single
        const double sw = ((((((w[0] + w[1]) + w[2]) + w[3]) + w[4]) + w[5]) + w[6]) + w[7];
        const float asw = (float)(a.scale * sw);
        for (int f = 0; f < a.frames; f++) {
            const long long q = (long long)f * nr * nc + p;
            const float io = a.Iout[q], ii = a.Iin[q];
            const float df = subf(io, ii);
            const float psi = __fdiv_rn(1.0f, __fsqrt_rn(addf(mulf(df, df), 2.220446049250313e-16f)));
            a.TRACE[q] = addf(psi, asw);
            a.B[q] = mulf(psi, ii);
        }
    }
}

}  // namespace

int op_ad_diff_weights(pdegpu_ctx *ctx, float *const w[8], float *TRACE, float *B, const float *D, const float *Iin,
                       int nr, int nc, int frames, double quantile, double scale, double *lambda_dev, int wframes)
{
    const long long n = (long long)nr * nc;
    const size_t need = 3 * n * sizeof(double) + sizeof(SelState) + 256;
    int rc = pdegpu_scratch_reserve(ctx, need);
    if (rc) return rc;
    double *mx = (double *)ctx->scratch, *my = mx + n, *nrm = my + n;
    SelState *st = (SelState *)(nrm + n);
    PDEGPU_CUDA_OK(ctx, cudaMemsetAsync(st, 0, sizeof(SelState), ctx->stream));
    PDEGPU_PROF(ctx, "ad_grad_kernel", (4.0 * frames + 24.0) * n);
    ad_grad_kernel<<<grid2(nr, nc, 1), 256, 0, ctx->stream>>>(mx, my, nrm, D, nr, nc, frames, st);
    PDEGPU_LAUNCH_CHECK(ctx, "ad_grad_kernel");
    sel_init_kernel<<<1, 1, 0, ctx->stream>>>(st, quantile);
    PDEGPU_LAUNCH_CHECK(ctx, "sel_init_kernel");
    const int blocks = (int)((n + 255) / 256 < 592 ? (n + 255) / 256 : 592);
    for (int shift = 56; shift >= 0; shift -= 8) {
        PDEGPU_PROF(ctx, "sel_hist_kernel", 8.0 * n);
        sel_hist_kernel<<<blocks, 256, 0, ctx->stream>>>(st, nrm, n, shift);
        PDEGPU_LAUNCH_CHECK(ctx, "sel_hist_kernel");
        sel_pick_kernel<<<1, 1, 0, ctx->stream>>>(st, shift);
        PDEGPU_LAUNCH_CHECK(ctx, "sel_pick_kernel");
    }
    AdArgs a;
    for (int k = 0; k < 8; k++) a.w[k] = w[k];
    a.TRACE = TRACE; a.B = B; a.Iout = D; a.Iin = Iin;
    a.mx = mx; a.my = my; a.nrm = nrm; a.st = st;
    a.nr = nr; a.nc = nc; a.frames = frames; a.wframes = wframes < 1 ? 1 : wframes; a.scale = scale;
    PDEGPU_PROF(ctx, "ad_weights_kernel", (24.0 + 32.0 + (TRACE ? 16.0 * frames : 0.0)) * n);
    ad_weights_kernel<<<grid2(nr, nc, 1), 256, 0, ctx->stream>>>(a);
    PDEGPU_LAUNCH_CHECK(ctx, "ad_weights_kernel");
    if (lambda_dev) PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(lambda_dev, &st->lambda, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return PDEGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// Constant terms of the FMG early-linearisation driver, FlowEminNDFASFMG_elin_2D_v10.m:143-149, per channel:
//   M = b1*Idy.*Idx + b2*Idxy.*(Idxx+Idyy), Cu = b1*Idt.*Idx + b2*(Idxt.*Idxx + Idyt.*Idxy), ... (single, left to right)
// and the pre-scaling of :124-125: Ist = (It0+It1).*0.55/255, Idt = (It0-It1)/255.
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
fmg_terms_kernel(float *__restrict__ M, float *__restrict__ Cu, float *__restrict__ Cv, float *__restrict__ Du, float *__restrict__ Dv,
                 const float *__restrict__ Idt, const float *__restrict__ Idx, const float *__restrict__ Idy,
                 const float *__restrict__ Idxt, const float *__restrict__ Idyt, const float *__restrict__ Idxx,
                 const float *__restrict__ Idyy, const float *__restrict__ Idxy, float b1, float b2, long long n)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float dt = Idt[t], dx = Idx[t], dy = Idy[t], xt = Idxt[t], yt = Idyt[t], xx = Idxx[t], yy = Idyy[t], xy = Idxy[t];
    M[t]  = addf(mulf(mulf(b1, dy), dx), mulf(mulf(b2, xy), addf(xx, yy)));
    Cu[t] = addf(mulf(mulf(b1, dt), dx), mulf(b2, addf(mulf(xt, xx), mulf(yt, xy))));
    Cv[t] = addf(mulf(mulf(b1, dt), dy), mulf(b2, addf(mulf(xt, xy), mulf(yt, yy))));
    Du[t] = addf(mulf(mulf(b1, dx), dx), mulf(b2, addf(mulf(xx, xx), mulf(xy, xy))));
    Dv[t] = addf(mulf(mulf(b1, dy), dy), mulf(b2, addf(mulf(xy, xy), mulf(yy, yy))));
}

__global__ void __launch_bounds__(256)
fmg_prescale_kernel(float *__restrict__ Ist, float *__restrict__ Idt, const float *__restrict__ I0, const float *__restrict__ I1, float div, long long n)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    Ist[t] = __fdiv_rn(mulf(addf(I0[t], I1[t]), 0.55f), div);
    Idt[t] = __fdiv_rn(subf(I0[t], I1[t]), div);
}

// out = sum(in, 3) in single, channel after channel (FlowEminHS_elin_2D_v10.m:168-172)
__global__ void __launch_bounds__(256)
channel_sum_kernel(float *__restrict__ out, const float *__restrict__ in, int channels, long long npix)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= npix) return;
    float acc = in[t];
    for (int c = 1; c < channels; c++) acc = addf(acc, in[(long long)c * npix + t]);
    out[t] = acc;
}

__global__ void __launch_bounds__(256)
fill_kernel(float *__restrict__ out, float v, long long n)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[t] = v;
}
}  // namespace

int op_fmg_terms(pdegpu_ctx *ctx, float *const out[5], const float *const der[8], float b1, float b2, long long n)
{
    PDEGPU_PROF(ctx, "fmg_terms_kernel", 52.0 * n);
    fmg_terms_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(out[0], out[1], out[2], out[3], out[4],
        der[0], der[1], der[2], der[3], der[4], der[5], der[6], der[7], b1, b2, n);
    PDEGPU_LAUNCH_CHECK(ctx, "fmg_terms_kernel");
    return PDEGPU_OK;
}

int op_channel_sum(pdegpu_ctx *ctx, float *out, const float *in, int channels, long long npix)
{
    PDEGPU_PROF(ctx, "channel_sum_kernel", 4.0 * npix * (channels + 1));
    channel_sum_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, ctx->stream>>>(out, in, channels, npix);
    PDEGPU_LAUNCH_CHECK(ctx, "channel_sum_kernel");
    return PDEGPU_OK;
}

int op_fill(pdegpu_ctx *ctx, float *out, float v, long long n)
{
    PDEGPU_PROF(ctx, "fill_kernel", 4.0 * n);
    fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(out, v, n);
    PDEGPU_LAUNCH_CHECK(ctx, "fill_kernel");
    return PDEGPU_OK;
}

// Ist = (I0+I1).*0.55/div, Idt = (I0-I1)/div: div = 255 in the FMG driver (FlowEminNDFASFMG_elin_2D_v10.m:124-125),
// 1 in the Horn-Schunck driver, whose frames are already scaled (FlowEminHS_elin_2D_v10.m:139-140)
int op_fmg_prescale(pdegpu_ctx *ctx, float *Ist, float *Idt, const float *I0, const float *I1, long long n, float div)
{
    PDEGPU_PROF(ctx, "fmg_prescale_kernel", 16.0 * n);
    fmg_prescale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(Ist, Idt, I0, I1, div, n);
    PDEGPU_LAUNCH_CHECK(ctx, "fmg_prescale_kernel");
    return PDEGPU_OK;
}
