// tline_common.cuh -- kernel generation 3 of solver 2 (zebra line relaxation): declarations shared by the
// pass kernel (sweeps_tline_impl.cuh), its preparation kernels and the host-side sequencing (sweeps_tline.cu).
//
// TOTAL-FIELD FORM. The late-linearisation rows of the reference relax an increment dX with fixed X0
// (e.g. middleColumn_llin4 opticalflowSolvers.c:2548-2639):
//     (sum w + D) dX_p - sum_n w_n dX_n = C - M dY_p + sum_n w_n (X0_n - X0_p)
// With T = X0 + dX this is the SAME system written as an early-linearisation one,
//     (sum w + D) T_p  - sum_n w_n T_n  = C' - M TY_p,      C' = C + D X0_p + M Y0_p,
// and SOR commutes with the shift (T_new = X0 + w dX_gs + (1-w) dX_old = w T_gs + (1-w) T_old). So a relax
// call converts once (T, C'), every pass reads 11 instead of 13 fields, and the fixed fields are never read
// again until the increment is recovered at the end (dX = T - X0). A row then needs only its own pixel's
// coefficients plus T at the two perpendicular neighbours: no in-line neighbour VALUES at all, which is what
// lets a line be cut into chunks, lanes and (long lines) segments freely.
//
// LINE LAYOUT. A pass sees a problem as `nlines` contiguous lines of `n` elements, stored with a pitch of n rounded
// up to 4 floats (every line starts 16-byte aligned, the requirement of cp.async.bulk). The column pass reads pitched
// copies of the caller's arrays, the row pass transposed pitched copies, both made once per call by the preparation
// kernel; a pass writes T transposed, i.e. in the layout the next pass reads.
#pragma once
#include "window_common.cuh"
#include <stdint.h>

namespace {

// PACKED LINES. The preparation kernel stores, for every line of a direction, its NC coefficient lines next to each
// other ([problem][line][field][pitch]) in the order below, and the NUNK unknown lines likewise ([problem][line][U,V][pitch]):
// a task's coefficients and a ring entry are then ONE bulk copy each. Fields are named by their role in the pass
// (P/N: previous/next element of the line, L/H: the lines j-1 / j+1; C0/D0: the unknown the pass solves first).
enum { TL_WP = 0, TL_WN = 1, TL_WL = 2, TL_WH = 3, TL_C0 = 4, TL_D0 = 5,
       TL_MM = 6, TL_C1 = 7, TL_D1 = 8 };                  // flow families; then the four diagonal weights LP, LN, HP, HN

struct TLParams {
    const float *coef;             // packed coefficient lines of this direction
    const float *tin;              // packed T lines (physical order U, V)
    float       *tout;             // T out = packed T lines of the OTHER direction (element k of line j -> element j of line k)
    int pitch, opitch, SP;         // floats between fields of a packed line (in / out); floats between ring entries
    int q0;                        // flow families: physical index (0 = U, 1 = V) of the unknown this pass solves first
    int n, nlines;                 // (longest) segment length, lines per problem
    // LONG LINES are cut into S segments of ns elements (the last one nlast): segment s of every line of problem b is
    // "problem" b*S + s of the pass, stored as such by the preparation kernel, and exact within itself. Across a cut the
    // neighbour's value is taken from the START of the pass (block Jacobi between segments: same fixed point). oS / ons:
    // the same for the other direction, whose layout this pass writes.
    int S, ns, nlast, nfull, oS, ons;
    int NB, TB;                    // blocks of BL lines per problem / in total
    int BL;                        // lines per output block (4 or 8)
    int R, D, K, NBR, NCW;         // ring lines, lead of the even lines in pairs, coefficient slabs, block slots, consumer warps
    int skip_border;               // 8-neighbour PDE: the first and the last line of a problem are not relaxed (pdeSolvers.c:1155,1290)
    float omega;
};

__device__ __forceinline__ uint32_t tl_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tl_smem(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tl_smem(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tl_smem(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *b, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(tl_smem(b)), "r"(parity) : "memory");
    return ok != 0;
}
// A wait that never ends is a scheduling bug: trap instead of hanging the GPU (try_wait itself sleeps in hardware).
__device__ __forceinline__ void mbar_wait(uint64_t *b, unsigned parity)
{
    unsigned spins = 0;
    while (!mbar_try(b, parity)) if (++spins > (1u << 22)) __trap();
}
// The same for a thread that has nothing else to do (a producer far ahead of its consumers): it sleeps between polls
// instead of re-issuing try_wait back to back, which costs its scheduler issue slots the solver warps could use.
__device__ __forceinline__ void mbar_wait_lazy(uint64_t *b, unsigned parity)
{
    unsigned spins = 0;
    while (!mbar_try(b, parity)) {
        __nanosleep(256);
        if (++spins > (1u << 22)) __trap();
    }
}
// 1-D bulk copy global -> shared through the TMA unit; completion is counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tl_smem(dst)), "l"(src), "r"(bytes), "r"(tl_smem(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void thread_wait_ge(const unsigned *p, unsigned want)
{
    unsigned spins = 0;
    while (ld_acquire(p) < want) {
        __nanosleep(64);
        if (++spins > (1u << 24)) __trap();
    }
}

}  // namespace
