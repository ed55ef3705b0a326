// driver_formulas.cuh -- per-pixel formulas of the late-linearisation flow driver, shared by
// the stand-alone kernels (driver_ops.cu: op_diff_weights_kernel, llin_terms_kernel) and the FUSED preparation of the
// line kernels (sweeps_tline.cu: weights and data terms are computed while the packed lines are written, north_star
// subsystem 3). Every operation is an explicit round-to-nearest intrinsic: both users produce the same bits.
#pragma once
#include "pdegpu_internal.cuh"

namespace {

struct LlinTermsArgs {
    const float *d1[3];       // I1dt, I1dx, I1dy                       (c1 channels)
    const float *d2[5];       // I2dt, I2dx, I2dy | I2dxt, I2dyt, I2dxx, I2dyy, I2dxy   (c2 channels)
    const float *dU, *dV;
    float *out[5];            // M, Cu, Cv, Du, Dv
    int c1, c2, gradmag;
    float b1, b2, alpha;
    long long npix;           // pixels per channel
    long long stride1, stride2, stride;   // batch strides of the d1 stacks, the d2 stacks, and of dU/dV/outputs
};

__device__ __forceinline__ int df_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ int df_wrapi(int v, int n) { return v < 0 ? v + n : (v >= n ? v - n : v); }
// single precision, one rounding per operation (Matlab / numpy)
__device__ __forceinline__ float df_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float df_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float df_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float df_sq(float a) { return __fmul_rn(a, a); }
__device__ __forceinline__ float df_nansum(float acc, float t) { return is_nan(t) ? acc : __fadd_rn(acc, t); }

// robust data weights + nansum over the channels, FlowEminND_llin_2D_v10.m:235-258,289-299,323-327, at pixel t of problem b
__device__ __forceinline__ void llin_terms_at(const LlinTermsArgs &a, long long b, long long t, float du, float dv,
                                              float &M, float &Cu, float &Cv, float &Du, float &Dv)
{
    M = 0.f; Cu = 0.f; Cv = 0.f; Du = 0.f; Dv = 0.f;
    for (int c = 0; c < a.c1; c++) {
        const long long p = b * a.stride1 + (long long)c * a.npix + t;
        const float It = a.d1[0][p], Ix = a.d1[1][p], Iy = a.d1[2][p];
        const float r = df_sub(df_sub(It, df_mul(Ix, du)), df_mul(Iy, dv));
        const float g = __fdiv_rn(a.b1, df_mul(a.alpha, __fsqrt_rn(df_add(df_sq(r), 0.00001f))));
        M = df_nansum(M, df_mul(df_mul(Iy, Ix), g));
        Cu = df_nansum(Cu, df_mul(df_mul(It, Ix), g));
        Cv = df_nansum(Cv, df_mul(df_mul(It, Iy), g));
        Du = df_nansum(Du, df_mul(df_mul(Ix, Ix), g));
        Dv = df_nansum(Dv, df_mul(df_mul(Iy, Iy), g));
    }
    // the driver sums cat(3, X1.*gD1, X2.*gD2): all first-term channels, then all second-term channels
    for (int c = 0; c < a.c2; c++) {
        const long long p = b * a.stride2 + (long long)c * a.npix + t;
        float m2, cu2, cv2, du2, dv2, op;
        if (!a.gradmag) {
            const float It = a.d2[0][p], Ix = a.d2[1][p], Iy = a.d2[2][p];
            op = df_sq(df_sub(df_sub(It, df_mul(Ix, du)), df_mul(Iy, dv)));
            m2 = df_mul(Iy, Ix); cu2 = df_mul(It, Ix); cv2 = df_mul(It, Iy); du2 = df_mul(Ix, Ix); dv2 = df_mul(Iy, Iy);
        } else {
            const float xt = a.d2[0][p], yt = a.d2[1][p], xx = a.d2[2][p], yy = a.d2[3][p], xy = a.d2[4][p];
            op = df_add(df_sq(df_sub(df_sub(xt, df_mul(xx, du)), df_mul(xy, dv))), df_sq(df_sub(df_sub(yt, df_mul(xy, du)), df_mul(yy, dv))));
            m2 = df_mul(xy, df_add(xx, yy));
            cu2 = df_add(df_mul(xt, xx), df_mul(yt, xy));
            cv2 = df_add(df_mul(xt, xy), df_mul(yt, yy));
            du2 = df_add(df_mul(xx, xx), df_mul(xy, xy));
            dv2 = df_add(df_mul(xy, xy), df_mul(yy, yy));
        }
        const float g = __fdiv_rn(a.b2, df_mul(a.alpha, __fsqrt_rn(df_add(op, 0.00001f))));
        M = df_nansum(M, df_mul(m2, g)); Cu = df_nansum(Cu, df_mul(cu2, g)); Cv = df_nansum(Cv, df_mul(cv2, g));
        Du = df_nansum(Du, df_mul(du2, g)); Dv = df_nansum(Dv, df_mul(dv2, g));
    }
}

// OPdiffWeights (FlowEminND_llin_2D_v10.m:389-433; double precision, circshift wraps, imfilter replicates) at pixel (i, j).
// The fields are U + dU, V + dV (single additions, as the driver forms single(U+dU) before the call) when dU / dV are
// given, else U, V themselves. All pointers at the problem's first pixel.
struct OpdiffSrc {
    const float *U, *V, *dU, *dV;
    __device__ __forceinline__ double u(long long p) const { return (double)(dU ? __fadd_rn(U[p], dU[p]) : U[p]); }
    __device__ __forceinline__ double v(long long p) const { return (double)(dV ? __fadd_rn(V[p], dV[p]) : V[p]); }
};

__device__ __forceinline__ void opdiff_at(const OpdiffSrc &s, int i, int j, int nr, int nc, float &wW, float &wN, float &wS, float &wE)
{
    auto P = [&](int ii, int jj) -> long long { return (long long)jj * nr + ii; };
    auto qd = [](double a, double b) { return __dadd_rn(__dmul_rn(0.25, a), __dmul_rn(-0.25, b)); };
    // vertical / horizontal central differences [0.25 0 -0.25] (correlation, replicate border) at (ii, jj)
    auto uver = [&](int ii, int jj) { return qd(s.u(P(df_clampi(ii - 1, 0, nr - 1), jj)), s.u(P(df_clampi(ii + 1, 0, nr - 1), jj))); };
    auto vver = [&](int ii, int jj) { return qd(s.v(P(df_clampi(ii - 1, 0, nr - 1), jj)), s.v(P(df_clampi(ii + 1, 0, nr - 1), jj))); };
    auto uhor = [&](int ii, int jj) { return qd(s.u(P(ii, df_clampi(jj - 1, 0, nc - 1))), s.u(P(ii, df_clampi(jj + 1, 0, nc - 1)))); };
    auto vhor = [&](int ii, int jj) { return qd(s.v(P(ii, df_clampi(jj - 1, 0, nc - 1))), s.v(P(ii, df_clampi(jj + 1, 0, nc - 1)))); };
    auto sq = [](double x) { return __dmul_rn(x, x); };
    auto edge = [&](double du_, double gu, double dv_, double gv) {
        return __dadd_rn(__dadd_rn(__dadd_rn(sq(du_), sq(gu)), sq(dv_)), sq(gv));
    };
    const int jw = df_wrapi(j - 1, nc), je = df_wrapi(j + 1, nc), in_ = df_wrapi(i - 1, nr), is = df_wrapi(i + 1, nr);
    const double u0 = s.u(P(i, j)), v0 = s.v(P(i, j));
    const double uv = uver(i, j), vv = vver(i, j), uh = uhor(i, j), vh = vhor(i, j);
    const double sW = edge(__dsub_rn(s.u(P(i, jw)), u0), __dadd_rn(uv, uver(i, jw)), __dsub_rn(s.v(P(i, jw)), v0), __dadd_rn(vv, vver(i, jw)));
    const double sE = edge(__dsub_rn(s.u(P(i, je)), u0), __dadd_rn(uv, uver(i, je)), __dsub_rn(s.v(P(i, je)), v0), __dadd_rn(vv, vver(i, je)));
    const double sN = edge(__dsub_rn(s.u(P(in_, j)), u0), __dadd_rn(uh, uhor(in_, j)), __dsub_rn(s.v(P(in_, j)), v0), __dadd_rn(vh, vhor(in_, j)));
    const double sS = edge(__dsub_rn(s.u(P(is, j)), u0), __dadd_rn(uh, uhor(is, j)), __dsub_rn(s.v(P(is, j)), v0), __dadd_rn(vh, vhor(is, j)));
    wW = (float)__ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(sW, 0.00001)));
    wE = (float)__ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(sE, 0.00001)));
    wN = (float)__ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(sN, 0.00001)));
    wS = (float)__ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(sS, 0.00001)));
}

// The east and south weights of pixel (i, j) only. The weight of an edge does not depend on the side it is seen from:
// sE(i, j) and sW(i, j+1) are sums of the same four squares (a difference and its negative, a sum in either order), so
// wW(i, j+1) = wE(i, j) and wN(i+1, j) = wS(i, j) bit for bit (indices wrap like circshift). The stand-alone kernel
// computes each edge once and stores it on both sides: half the double-precision square roots and divisions.
__device__ __forceinline__ void opdiff_east_south_at(const OpdiffSrc &s, int i, int j, int nr, int nc, float &wE, float &wS)
{
    auto P = [&](int ii, int jj) -> long long { return (long long)jj * nr + ii; };
    auto qd = [](double a, double b) { return __dadd_rn(__dmul_rn(0.25, a), __dmul_rn(-0.25, b)); };
    auto uver = [&](int ii, int jj) { return qd(s.u(P(df_clampi(ii - 1, 0, nr - 1), jj)), s.u(P(df_clampi(ii + 1, 0, nr - 1), jj))); };
    auto vver = [&](int ii, int jj) { return qd(s.v(P(df_clampi(ii - 1, 0, nr - 1), jj)), s.v(P(df_clampi(ii + 1, 0, nr - 1), jj))); };
    auto uhor = [&](int ii, int jj) { return qd(s.u(P(ii, df_clampi(jj - 1, 0, nc - 1))), s.u(P(ii, df_clampi(jj + 1, 0, nc - 1)))); };
    auto vhor = [&](int ii, int jj) { return qd(s.v(P(ii, df_clampi(jj - 1, 0, nc - 1))), s.v(P(ii, df_clampi(jj + 1, 0, nc - 1)))); };
    auto sq = [](double x) { return __dmul_rn(x, x); };
    auto edge = [&](double du_, double gu, double dv_, double gv) {
        return __dadd_rn(__dadd_rn(__dadd_rn(sq(du_), sq(gu)), sq(dv_)), sq(gv));
    };
    const int je = df_wrapi(j + 1, nc), is = df_wrapi(i + 1, nr);
    const double u0 = s.u(P(i, j)), v0 = s.v(P(i, j));
    const double sE = edge(__dsub_rn(s.u(P(i, je)), u0), __dadd_rn(uver(i, j), uver(i, je)), __dsub_rn(s.v(P(i, je)), v0), __dadd_rn(vver(i, j), vver(i, je)));
    const double sS = edge(__dsub_rn(s.u(P(is, j)), u0), __dadd_rn(uhor(i, j), uhor(is, j)), __dsub_rn(s.v(P(is, j)), v0), __dadd_rn(vhor(i, j), vhor(is, j)));
    wE = (float)__ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(sE, 0.00001)));
    wS = (float)__ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(sS, 0.00001)));
}

}  // namespace
