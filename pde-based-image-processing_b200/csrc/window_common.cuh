// window_common.cuh -- pieces shared by the sliding-window line-relaxation kernels
// (sweeps_window.cu: one warp per line; sweeps_window2.cu: assembler warps + solver warps).
#pragma once
#include "line_rows.cuh"

namespace {

struct WinParams {
    SysView s;             // problem with contiguous lines (lines = columns of s); s.x = X_in
    float *xout[2];        // X_out, transposed layout: element i of line j of problem b at b*ostride + i*nlines + j
    long long ostride;
    int n, nlines;         // line length (= s.nrows), lines per problem (= s.ncols)
    int NB, TB;            // 8-line blocks per problem, in total
    int R, D;              // ring size in lines (multiple of 8), how many pairs the even lines run ahead
    int NA, NS, NBUF;      // two-stage kernel: assembler warps, solver warps, row buffers
    int vec_ok;            // float4 stores of X_out are aligned
    int aligned;           // float4 loads of the inputs are aligned (else 4 scalar loads per vector)
    float omega;
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned *p, unsigned v)
{
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
// Whole warp waits until *p >= want. EVERY lane polls (one broadcast shared-memory read) and the exit test is a
// warp vote: a lane-0-only spin leaves lane 0 and lanes 1..31 on separate control-flow paths, and the compiler may
// keep them apart for the rest of the loop body, issuing every instruction twice (seen in ncu: avg 16 threads per
// instruction, 2x instruction count). A wait that never ends is a scheduling bug: trap instead of hanging the GPU.
__device__ __forceinline__ void warp_wait_ge(const unsigned *p, unsigned want, int lane)
{
    (void)lane;
    unsigned spins = 0;
    while (!__all_sync(0xffffffffu, ld_acquire(p) >= want)) {
        __nanosleep(100);
        if (++spins > (1u << 23)) __trap();
    }
}

// Partitioned Thomas solve of one line spread over the warp: lane L holds rows L*M .. L*M+M-1 of
//   a_k x_{k-1} + b_k x_k + c_k x_{k+1} = d_k
// (rows past the end of the line are identity rows). On return d[] holds x. a, b are clobbered.
template <int M>
__device__ __forceinline__ void chunk_solve(float (&a)[M], const float (&c)[M], float (&b)[M], float (&d)[M], int lane)
{
    const unsigned FULL = 0xffffffffu;
    // local forward elimination; spike a[] multiplies x_left = last unknown of the previous lane.
    // afterwards row k reads  x_k + b[k]*x_{k+1} + a[k]*x_left = d[k]
    {
        const float inv = fast_rcp(b[0]);
        b[0] = c[0] * inv; d[0] *= inv; a[0] *= inv;
    }
#pragma unroll
    for (int k = 1; k < M; k++) {
        const float ak = a[k];
        const float inv = fast_rcp(b[k] - ak * b[k - 1]);
        b[k] = c[k] * inv;
        d[k] = (d[k] - ak * d[k - 1]) * inv;
        a[k] = (-ak * a[k - 1]) * inv;
    }
    // first unknown of the chunk in terms of the last one and x_left:  x_first = Af - Bf*x_last - Gf*x_left
    float Af, Bf, Gf;
    if (M == 1) { Af = 0.f; Bf = -1.0f; Gf = 0.f; }
    else {
        Af = d[M - 2]; Bf = b[M - 2]; Gf = a[M - 2];
#pragma unroll
        for (int r = M - 3; r >= 0; r--) {
            Af = d[r] - b[r] * Af;
            Bf = -b[r] * Bf;
            Gf = a[r] - b[r] * Gf;
        }
    }
    // interface system in the lanes' last unknowns l:  al*l[-1] + be*l + ga*l[+1] = de
    float An = __shfl_down_sync(FULL, Af, 1), Bn = __shfl_down_sync(FULL, Bf, 1), Gn = __shfl_down_sync(FULL, Gf, 1);
    if (lane == 31) { An = 0.f; Bn = 0.f; Gn = 0.f; }
    const float cp = b[M - 1];
    float al = a[M - 1], be = 1.0f - cp * Gn, ga = -cp * Bn, de = d[M - 1] - cp * An;
#pragma unroll
    for (int st = 1; st < 32; st <<= 1) {
        float alm = __shfl_up_sync(FULL, al, st), bem = __shfl_up_sync(FULL, be, st);
        float gam = __shfl_up_sync(FULL, ga, st), dem = __shfl_up_sync(FULL, de, st);
        float alp = __shfl_down_sync(FULL, al, st), bep = __shfl_down_sync(FULL, be, st);
        float gap = __shfl_down_sync(FULL, ga, st), dep = __shfl_down_sync(FULL, de, st);
        if (lane < st)      { alm = 0.f; bem = 1.0f; gam = 0.f; dem = 0.f; }
        if (lane + st > 31) { alp = 0.f; bep = 1.0f; gap = 0.f; dep = 0.f; }
        const float k1 = al * fast_rcp(bem), k2 = ga * fast_rcp(bep);
        be = be - gam * k1 - alp * k2;
        de = de - dem * k1 - dep * k2;
        al = -alm * k1;
        ga = -gap * k2;
    }
    const float l = de * fast_rcp(be);
    float L = __shfl_up_sync(FULL, l, 1);
    if (lane == 0) L = 0.f;
    // local back substitution
    float x = l;
    d[M - 1] = x;
#pragma unroll
    for (int r = M - 2; r >= 0; r--) {
        x = d[r] - b[r] * x - a[r] * L;
        d[r] = x;
    }
}

template <int NUNK> struct RowF;
template <> struct RowF<2> { enum { A = 0, C = 1, B1 = 2, D1 = 3, B2 = 4, D2 = 5, MM = 6, N = 7 }; };
template <> struct RowF<1> { enum { A = 0, C = 1, B1 = 2, D1 = 3, B2 = 2, D2 = 3, MM = 0, N = 4 }; };

constexpr int kWinMaxWarps = 8;

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
// 4 consecutive elements starting at p; unaligned lines: 4 scalar loads, clamped to the `room` elements left in the line
__device__ __forceinline__ float4 ldv(const float *p, bool aligned, int room)
{
    if (aligned) return ld4(p);
    float4 v;
    v.x = p[0]; v.y = p[min(1, room)]; v.z = p[min(2, room)]; v.w = p[min(3, room)];
    return v;
}
__device__ __forceinline__ void st4(float *p, const float (&v)[4]) { *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]); }
#define V4(v, kk) ((kk) == 0 ? (v).x : (kk) == 1 ? (v).y : (kk) == 2 ? (v).z : (v).w)

// one line of a CTA's schedule
struct WinTask {
    int l, img, jb, j, ibase, dW, dE;
    bool odd, eW, eE, owned;
};

// Raw operands of 4 consecutive pixels of a line (one float4 per field), loaded with no use of the values,
// so that a lane can have the loads of two batches in flight.
template <int FAM>
struct RawBatch {
    using F = Fam<FAM>;
    static constexpr int NUNK = F::NUNK, NN = F::EIGHT ? 8 : 4, NL = F::LATE ? F::NUNK : 1;
    float4 w4[NN], C4[NUNK], D4[NUNK], XO4[NUNK], XW4[NUNK], XE4[NUNK], M4;
    float4 X0C4[NL], X0W4[NL], X0E4[NL];
    float x0l[NUNK], x0r[NUNK];                                    // in-line neighbours beyond the vector
    float xWl[NUNK], xWr[NUNK], xEl[NUNK], xEr[NUNK];              // 8-neighbour: diagonal neighbours beyond the vector
    float x0Wl[NUNK], x0Wr[NUNK], x0El[NUNK], x0Er[NUNK];

    // ec = first element (clamped into the line), T = the line
    __device__ __forceinline__ void issue(const SysView &s, const WinTask &T, int ec, int n, bool al)
    {
        const int ip = T.ibase + ec;
        const int room = n - 1 - ec;                          // elements after ec that are still inside the line
#define ld4(ptr) ldv((ptr), al, room)
        const int ipl = T.ibase + max(ec - 1, 0), ipr = T.ibase + min(ec + 4, n - 1);
        const int dW = T.dW, dE = T.dE;
#pragma unroll
        for (int nn = 0; nn < NN; nn++) w4[nn] = ld4(s.w[nn] + ip);
#pragma unroll
        for (int qq = 0; qq < NUNK; qq++) {
            C4[qq] = ld4(s.c[qq] + ip); D4[qq] = ld4(s.d[qq] + ip); XO4[qq] = ld4(s.x[qq] + ip);
            if (F::LATE) {
                X0C4[qq] = ld4(s.x0[qq] + ip); X0W4[qq] = ld4(s.x0[qq] + ip + dW); X0E4[qq] = ld4(s.x0[qq] + ip + dE);
                x0l[qq] = s.x0[qq][ipl]; x0r[qq] = s.x0[qq][ipr];
                if (F::EIGHT) {
                    x0Wl[qq] = s.x0[qq][ipl + dW]; x0Wr[qq] = s.x0[qq][ipr + dW];
                    x0El[qq] = s.x0[qq][ipl + dE]; x0Er[qq] = s.x0[qq][ipr + dE];
                }
            }
            if (!T.odd) {
                XW4[qq] = ld4(s.x[qq] + ip + dW); XE4[qq] = ld4(s.x[qq] + ip + dE);
                if (F::EIGHT) {
                    xWl[qq] = s.x[qq][ipl + dW]; xWr[qq] = s.x[qq][ipr + dW];
                    xEl[qq] = s.x[qq][ipl + dE]; xEr[qq] = s.x[qq][ipr + dE];
                }
            }
        }
        if (NUNK == 2) M4 = ld4(s.m + ip);
#undef ld4
    }

    // odd lines: the unknowns at the (even) neighbour lines come from the ring of solved lines
    __device__ __forceinline__ void neighbours_from_ring(const float *rsW, const float *rsE, int P, int ec, int n)
    {
#pragma unroll
        for (int qq = 0; qq < NUNK; qq++) {
            XW4[qq] = ld4(rsW + qq * P + ec); XE4[qq] = ld4(rsE + qq * P + ec);
            if (F::EIGHT) {
                xWl[qq] = rsW[qq * P + max(ec - 1, 0)]; xWr[qq] = rsW[qq * P + min(ec + 4, n - 1)];
                xEl[qq] = rsE[qq * P + max(ec - 1, 0)]; xEr[qq] = rsE[qq * P + min(ec + 4, n - 1)];
            }
        }
    }

    // operands of pixel k (0..3) of the vector, in the form the row formulas take
    template <int DIR>
    __device__ __forceinline__ void pixel(int k, int i, int n, bool eW, bool eE, PixelRaw<FAM, DIR> &r) const
    {
        const bool eN = i > 0, eS = i < n - 1;
        r.exmask = (eW ? 1u << W_W : 0u) | (eN ? 1u << W_N : 0u) | (eE ? 1u << W_E : 0u) | (eS ? 1u << W_S : 0u)
                 | (eN && eW ? 1u << W_NW : 0u) | (eN && eE ? 1u << W_NE : 0u) | (eS && eE ? 1u << W_SE : 0u) | (eS && eW ? 1u << W_SW : 0u);
#pragma unroll
        for (int nn = 0; nn < NN; nn++) r.w[nn] = V4(w4[nn], k);
#pragma unroll
        for (int qq = 0; qq < NUNK; qq++) {
            r.C[qq] = V4(C4[qq], k); r.D[qq] = V4(D4[qq], k); r.xo[qq] = V4(XO4[qq], k);
            r.xn[qq][W_W] = V4(XW4[qq], k); r.xn[qq][W_E] = V4(XE4[qq], k);
            if (F::EIGHT) {
                r.xn[qq][W_NW % NN] = k > 0 ? V4(XW4[qq], k - 1) : xWl[qq];
                r.xn[qq][W_SW % NN] = k < 3 ? V4(XW4[qq], k + 1) : xWr[qq];
                r.xn[qq][W_NE % NN] = k > 0 ? V4(XE4[qq], k - 1) : xEl[qq];
                r.xn[qq][W_SE % NN] = k < 3 ? V4(XE4[qq], k + 1) : xEr[qq];
            }
            if (F::LATE) {
                r.x0c[qq] = V4(X0C4[qq], k);
                r.x0n[qq][W_W] = V4(X0W4[qq], k); r.x0n[qq][W_E] = V4(X0E4[qq], k);
                r.x0n[qq][W_N] = k > 0 ? V4(X0C4[qq], k - 1) : x0l[qq];
                r.x0n[qq][W_S] = k < 3 ? V4(X0C4[qq], k + 1) : x0r[qq];
                if (F::EIGHT) {
                    r.x0n[qq][W_NW % NN] = k > 0 ? V4(X0W4[qq], k - 1) : x0Wl[qq];
                    r.x0n[qq][W_SW % NN] = k < 3 ? V4(X0W4[qq], k + 1) : x0Wr[qq];
                    r.x0n[qq][W_NE % NN] = k > 0 ? V4(X0E4[qq], k - 1) : x0El[qq];
                    r.x0n[qq][W_SE % NN] = k < 3 ? V4(X0E4[qq], k + 1) : x0Er[qq];
                }
            }
        }
        r.M = NUNK == 2 ? V4(M4, k) : 0.f;
    }
};


#undef V4

}  // namespace
