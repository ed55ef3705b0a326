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
    int vec_ok;            // stores of X_out: 2 = 16-byte aligned, 1 = 8-byte aligned, 0 = scalar
    int aligned;           // vector loads of the inputs: 2 = 16-byte aligned, 1 = 8-byte aligned, 0 = scalar loads
    int G;                 // two-stage kernel: assembler warps per line
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
#ifndef PDEGPU_POLL_NS
#define PDEGPU_POLL_NS 100
#endif
__device__ __forceinline__ void warp_wait_ge(const unsigned *p, unsigned want, int lane)
{
    (void)lane;
    unsigned spins = 0;
    while (!__all_sync(0xffffffffu, ld_acquire(p) >= want)) {
        __nanosleep(PDEGPU_POLL_NS);
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
__device__ __forceinline__ void st4(float *p, const float (&v)[4]) { *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void stv(float *p, const float (&v)[4]) { st4(p, v); }
__device__ __forceinline__ void stv(float *p, const float (&v)[2]) { *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]); }

// VW consecutive elements starting at p, as ONE vector load (VW = 4: 16 B, VW = 2: 8 B). Unaligned lines: VW scalar
// loads, clamped to the `room` elements left in the line.
// (ALIGNED is a compile-time flag on purpose: with a run-time flag ptxas predicates both paths into one instruction
// stream, and a predicated-off scalar load that names a register of an in-flight vector load still waits for it --
// measured: one full memory latency in the middle of the load-issue block, profiles/r01_alr_window2_sass_regions.txt)
template <int VW> struct Vec { float v[VW]; };
template <int VW, bool ALIGNED>
__device__ __forceinline__ Vec<VW> ldv(const float *p, int room)
{
    Vec<VW> r;
    if (ALIGNED) {
        if (VW == 4) { const float4 t = *reinterpret_cast<const float4 *>(p); r.v[0] = t.x; r.v[1] = t.y; r.v[VW - 2] = t.z; r.v[VW - 1] = t.w; }
        else         { const float2 t = *reinterpret_cast<const float2 *>(p); r.v[0] = t.x; r.v[VW - 1] = t.y; }
    } else {
#pragma unroll
        for (int k = 0; k < VW; k++) r.v[k] = p[min(k, room)];
    }
    return r;
}
// same from shared memory (always aligned)
template <int VW>
__device__ __forceinline__ Vec<VW> ldsv(const float *p) { return ldv<VW, true>(p, VW); }

// one line of a CTA's schedule
struct WinTask {
    int l, img, jb, j, ibase, dW, dE;
    bool odd, eW, eE, owned;
};

// Raw operands of VW consecutive pixels of a line (one vector per field), loaded with no use of the values,
// so that a lane can have the loads of two batches in flight.
template <int FAM, int VW = 4>
struct RawBatch {
    using F = Fam<FAM>;
    using V = Vec<VW>;
    static constexpr int NUNK = F::NUNK, NN = F::EIGHT ? 8 : 4, NL = F::LATE ? F::NUNK : 1;
    V w4[NN], C4[NUNK], D4[NUNK], XO4[NUNK], XW4[NUNK], XE4[NUNK], M4;
    V X0C4[NL], X0W4[NL], X0E4[NL];
    float x0l[NUNK], x0r[NUNK];                                    // in-line neighbours beyond the vector
    float xWl[NUNK], xWr[NUNK], xEl[NUNK], xEr[NUNK];              // 8-neighbour: diagonal neighbours beyond the vector
    float x0Wl[NUNK], x0Wr[NUNK], x0El[NUNK], x0Er[NUNK];

    // ec = first element (clamped into the line), T = the line
    template <bool AL>
    __device__ __forceinline__ void issue(const SysView &s, const WinTask &T, int ec, int n)
    {
        const int ip = T.ibase + ec;
        const int room = n - 1 - ec;                          // elements after ec that are still inside the line
#define LDV(ptr) ldv<VW, AL>((ptr), room)
        const int ipl = T.ibase + max(ec - 1, 0), ipr = T.ibase + min(ec + VW, n - 1);
        const int dW = T.dW, dE = T.dE;
#pragma unroll
        for (int nn = 0; nn < NN; nn++) w4[nn] = LDV(s.w[nn] + ip);
#pragma unroll
        for (int qq = 0; qq < NUNK; qq++) {
            C4[qq] = LDV(s.c[qq] + ip); D4[qq] = LDV(s.d[qq] + ip); XO4[qq] = LDV(s.x[qq] + ip);
            if (F::LATE) {
                X0C4[qq] = LDV(s.x0[qq] + ip); X0W4[qq] = LDV(s.x0[qq] + ip + dW); X0E4[qq] = LDV(s.x0[qq] + ip + dE);
                x0l[qq] = s.x0[qq][ipl]; x0r[qq] = s.x0[qq][ipr];
                if (F::EIGHT) {
                    x0Wl[qq] = s.x0[qq][ipl + dW]; x0Wr[qq] = s.x0[qq][ipr + dW];
                    x0El[qq] = s.x0[qq][ipl + dE]; x0Er[qq] = s.x0[qq][ipr + dE];
                }
            }
            if (!T.odd) {
                XW4[qq] = LDV(s.x[qq] + ip + dW); XE4[qq] = LDV(s.x[qq] + ip + dE);
                if (F::EIGHT) {
                    xWl[qq] = s.x[qq][ipl + dW]; xWr[qq] = s.x[qq][ipr + dW];
                    xEl[qq] = s.x[qq][ipl + dE]; xEr[qq] = s.x[qq][ipr + dE];
                }
            }
        }
        if (NUNK == 2) M4 = LDV(s.m + ip);
#undef LDV
    }

    // odd lines: the unknowns at the (even) neighbour lines come from the ring of solved lines
    __device__ __forceinline__ void neighbours_from_ring(const float *rsW, const float *rsE, int P, int ec, int n)
    {
#pragma unroll
        for (int qq = 0; qq < NUNK; qq++) {
            XW4[qq] = ldsv<VW>(rsW + qq * P + ec); XE4[qq] = ldsv<VW>(rsE + qq * P + ec);
            if (F::EIGHT) {
                xWl[qq] = rsW[qq * P + max(ec - 1, 0)]; xWr[qq] = rsW[qq * P + min(ec + VW, n - 1)];
                xEl[qq] = rsE[qq * P + max(ec - 1, 0)]; xEr[qq] = rsE[qq * P + min(ec + VW, n - 1)];
            }
        }
    }

    // operands of pixel k (0..VW-1) of the vector, in the form the row formulas take
    template <int DIR>
    __device__ __forceinline__ void pixel(int k, int i, int n, bool eW, bool eE, PixelRaw<FAM, DIR> &r) const
    {
        const bool eN = i > 0, eS = i < n - 1;
        const int km = k > 0 ? k - 1 : 0, kp = k < VW - 1 ? k + 1 : VW - 1;
        r.exmask = (eW ? 1u << W_W : 0u) | (eN ? 1u << W_N : 0u) | (eE ? 1u << W_E : 0u) | (eS ? 1u << W_S : 0u)
                 | (eN && eW ? 1u << W_NW : 0u) | (eN && eE ? 1u << W_NE : 0u) | (eS && eE ? 1u << W_SE : 0u) | (eS && eW ? 1u << W_SW : 0u);
#pragma unroll
        for (int nn = 0; nn < NN; nn++) r.w[nn] = w4[nn].v[k];
#pragma unroll
        for (int qq = 0; qq < NUNK; qq++) {
            r.C[qq] = C4[qq].v[k]; r.D[qq] = D4[qq].v[k]; r.xo[qq] = XO4[qq].v[k];
            r.xn[qq][W_W] = XW4[qq].v[k]; r.xn[qq][W_E] = XE4[qq].v[k];
            if (F::EIGHT) {
                r.xn[qq][W_NW % NN] = k > 0 ? XW4[qq].v[km] : xWl[qq];
                r.xn[qq][W_SW % NN] = k < VW - 1 ? XW4[qq].v[kp] : xWr[qq];
                r.xn[qq][W_NE % NN] = k > 0 ? XE4[qq].v[km] : xEl[qq];
                r.xn[qq][W_SE % NN] = k < VW - 1 ? XE4[qq].v[kp] : xEr[qq];
            }
            if (F::LATE) {
                r.x0c[qq] = X0C4[qq].v[k];
                r.x0n[qq][W_W] = X0W4[qq].v[k]; r.x0n[qq][W_E] = X0E4[qq].v[k];
                r.x0n[qq][W_N] = k > 0 ? X0C4[qq].v[km] : x0l[qq];
                r.x0n[qq][W_S] = k < VW - 1 ? X0C4[qq].v[kp] : x0r[qq];
                if (F::EIGHT) {
                    r.x0n[qq][W_NW % NN] = k > 0 ? X0W4[qq].v[km] : x0Wl[qq];
                    r.x0n[qq][W_SW % NN] = k < VW - 1 ? X0W4[qq].v[kp] : x0Wr[qq];
                    r.x0n[qq][W_NE % NN] = k > 0 ? X0E4[qq].v[km] : x0El[qq];
                    r.x0n[qq][W_SE % NN] = k < VW - 1 ? X0E4[qq].v[kp] : x0Er[qq];
                }
            }
        }
        r.M = NUNK == 2 ? M4.v[k] : 0.f;
    }
};

}  // namespace
