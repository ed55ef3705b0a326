// sweeps_line.cu -- kernel generation 1, solver 2: zebra alternating line relaxation.
//
// One launch relaxes every line of one colour along one direction, for BOTH unknowns of a flow
// family (first unknown, SOR, then the second unknown with the first one's new values -- the order
// the reference uses inside a line pass, opticalflowSolvers.c:238-257).
//
// A CTA owns G lines (one warp each) and works in three phases:
//   A  assembly   all threads walk the CTA's pixels in memory order (coalesced), evaluate the
//                 reference's tridiagonal row (a,b,c,d) of each pixel (stencil_math.cuh) and park it
//                 in shared memory;
//   B  solve      one warp per line, one CHUNK of consecutive unknowns per lane: partitioned Thomas
//                 (Wang): local forward elimination with a left spike, local backward pass to get the
//                 chunk's interface relation, a 32-unknown interface system solved inside the warp by
//                 parallel cyclic reduction on shuffles, local back substitution. Same solution as a
//                 serial Thomas solve to rounding; 32-way parallel along the line;
//   C  write-back all threads store the relaxed line values in memory order.
// Shared memory holds 9 floats per pixel (flow) / 6 (scalar families); chunks are padded to an odd
// pitch so that lanes (= chunks) hit distinct banks.
//
// HBM traffic per launch = every coefficient of the active lines once + the unknowns of the active
// lines and of their perpendicular neighbours once + one write of the active lines.
#include "line_rows.cuh"
#include <stdlib.h>

namespace {

constexpr int kMaxSmem = 220 * 1024;

template <int NUNK> struct Slots;
template <> struct Slots<2> { enum { A = 0, C = 1, B1 = 2, D1 = 3, B2 = 4, D2 = 5, M = 6, GP = 7, XO = 8, N = 9 }; };
template <> struct Slots<1> { enum { A = 0, C = 1, B1 = 2, D1 = 3, GP = 4, XO = 5, B2 = 2, D2 = 3, M = 0, N = 6 }; };

// Solves a_k x_{k-1} + b_k x_k + c_k x_{k+1} = d_k (- mm_k * y_k when COUPLED), k = 0..n-1, in place:
// on return D[k] holds omega*x_k + (1-omega)*XO[k] (RELAX) or x_k. B and GP are clobbered.
// Every loop loads row r+1 before it works on row r (the recurrences are latency chains).
template <bool COUPLED, bool RELAX>
__device__ __forceinline__ void warp_line_solve(const float *__restrict__ A, const float *__restrict__ Cc,
                                                float *__restrict__ B, float *__restrict__ D, float *__restrict__ GP,
                                                const float *__restrict__ MM, const float *__restrict__ Y,
                                                const float *__restrict__ XO,
                                                int n, int Lc, int Lp, float omega, int lane)
{
    const unsigned FULL = 0xffffffffu;
    const int s0 = lane * Lc;
    const int len = max(0, min(n, s0 + Lc) - s0);
    const int o = lane * Lp;                      // padded smem offset of this lane's chunk
    float cp = 0.f, dp = 0.f, gp = 0.f;
    // pass 1: local forward elimination, spike column gp multiplies x_{s0-1}
    if (len > 0) {
        float an = A[o], bn = B[o], cn = Cc[o], dn = D[o];
        if (COUPLED) dn -= MM[o] * Y[o];
        for (int r = 0; r < len; r++) {
            const float a = an, b = bn, c = cn, d = dn;
            if (r + 1 < len) {
                an = A[o + r + 1]; bn = B[o + r + 1]; cn = Cc[o + r + 1]; dn = D[o + r + 1];
                if (COUPLED) dn -= MM[o + r + 1] * Y[o + r + 1];
            }
            const float inv = fast_rcp(r == 0 ? b : b - a * cp);
            dp = (r == 0 ? d : d - a * dp) * inv;
            gp = (r == 0 ? a : -a * gp) * inv;
            cp = c * inv;
            B[o + r] = cp; D[o + r] = dp; GP[o + r] = gp;
        }
    }
    // last-row relation of the chunk:  x_t + cp*x_{t+1} + gp*x_{s0-1} = dp
    // pass 2: first-row relation  x_s = Af - Bf*x_t - Gf*x_{s0-1}
    float Af = 0.f, Bf = 0.f, Gf = 0.f;
    if (len == 1) Bf = -1.0f;
    else if (len >= 2) {
        Af = D[o + len - 2]; Bf = B[o + len - 2]; Gf = GP[o + len - 2];
        if (len >= 3) {
            float cn = B[o + len - 3], dn = D[o + len - 3], gn = GP[o + len - 3];
            for (int r = len - 3; r >= 0; r--) {
                const float cpr = cn, dpr = dn, gpr = gn;
                if (r > 0) { cn = B[o + r - 1]; dn = D[o + r - 1]; gn = GP[o + r - 1]; }
                Af = dpr - cpr * Af;
                Bf = -cpr * Bf;
                Gf = gpr - cpr * Gf;
            }
        }
    }
    // interface system in the chunks' last unknowns l: al*l[-1] + be*l + ga*l[+1] = de
    float An = __shfl_down_sync(FULL, Af, 1), Bn = __shfl_down_sync(FULL, Bf, 1), Gn = __shfl_down_sync(FULL, Gf, 1);
    if (lane == 31) { An = 0.f; Bn = 0.f; Gn = 0.f; }
    float al, be, ga, de;
    if (len > 0) { al = gp; be = 1.0f - cp * Gn; ga = -cp * Bn; de = dp - cp * An; }
    else         { al = 0.f; be = 1.0f; ga = 0.f; de = 0.f; }
    // parallel cyclic reduction over the 32 lanes
#pragma unroll
    for (int st = 1; st < 32; st <<= 1) {
        float alm = __shfl_up_sync(FULL, al, st), bem = __shfl_up_sync(FULL, be, st);
        float gam = __shfl_up_sync(FULL, ga, st), dem = __shfl_up_sync(FULL, de, st);
        float alp = __shfl_down_sync(FULL, al, st), bep = __shfl_down_sync(FULL, be, st);
        float gap = __shfl_down_sync(FULL, ga, st), dep = __shfl_down_sync(FULL, de, st);
        if (lane < st)       { alm = 0.f; bem = 1.0f; gam = 0.f; dem = 0.f; }
        if (lane + st > 31)  { alp = 0.f; bep = 1.0f; gap = 0.f; dep = 0.f; }
        const float k1 = al * fast_rcp(bem), k2 = ga * fast_rcp(bep);
        be = be - gam * k1 - alp * k2;
        de = de - dem * k1 - dep * k2;
        al = -alm * k1;
        ga = -gap * k2;
    }
    const float l = de * fast_rcp(be);
    float L = __shfl_up_sync(FULL, l, 1);
    if (lane == 0) L = 0.f;
    // pass 3: local back substitution
    if (len > 0) {
        float x = l;
        float xo = RELAX ? XO[o + len - 1] : 0.f;
        float cn = 0.f, dn = 0.f, gn = 0.f, xon = 0.f;
        if (len >= 2) { cn = B[o + len - 2]; dn = D[o + len - 2]; gn = GP[o + len - 2]; if (RELAX) xon = XO[o + len - 2]; }
        D[o + len - 1] = RELAX ? omega * x + (1.0f - omega) * xo : x;
        for (int r = len - 2; r >= 0; r--) {
            const float cpr = cn, dpr = dn, gpr = gn, xor_ = xon;
            if (r > 0) { cn = B[o + r - 1]; dn = D[o + r - 1]; gn = GP[o + r - 1]; if (RELAX) xon = XO[o + r - 1]; }
            x = dpr - cpr * x - gpr * L;
            D[o + r] = RELAX ? omega * x + (1.0f - omega) * xor_ : x;
        }
    }
}

template <int FAM, int DIR, int G, int TPL>      // G lines per CTA, TPL threads per line (>= 32)
__global__ void __launch_bounds__(TPL * G)
alr_kernel(SysView s, int colour, float omega, int first_line, int nslots, int Lc, int Lp, int dbg)
{
    using F = Fam<FAM>;
    using S = Slots<F::NUNK>;
    extern __shared__ float smem[];
    constexpr int qa = (F::NUNK == 2 && DIR != 0) ? 1 : 0, qb = 1 - qa;
    constexpr int NT = TPL * G;
    const int nr = s.nrows, nc = s.ncols;
    const int n = (DIR & 1) == 0 ? nr : nc;
    const int NP = 32 * Lp;                                  // padded line length
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot0 = blockIdx.x * G;                        // first line slot of this CTA
    const int nl = min(G, nslots - slot0);                   // lines really present
    // move every pointer to this CTA's problem once; pixel indices are plain ints from here on
    {
        const long long base = (long long)blockIdx.y * s.bstride;
#pragma unroll
        for (int q = 0; q < F::NUNK; q++) { s.x[q] += base; s.c[q] += base; s.d[q] += base; if (F::LATE) s.x0[q] += base; }
        if (F::NUNK == 2) s.m += base;
#pragma unroll
        for (int q = 0; q < (F::EIGHT ? 8 : 4); q++) s.w[q] += base;
    }
    // thread -> (line g, first position k0, stride) in memory-friendly order:
    //   DIR 0: consecutive threads walk consecutive i of one column (128-B rows);
    //   DIR 1: consecutive threads take the CTA's G neighbouring lines at one j (one sector), then the next j.
    const int g = (DIR & 1) == 0 ? tid / TPL : tid % G;
    const int kfirst = (DIR & 1) == 0 ? tid % TPL : tid / G;
    constexpr int KSTEP = TPL;
    const int line = first_line + colour + 2 * (slot0 + g);
    const bool active = g < nl;
    float *Ls = smem + (size_t)g * S::N * NP;
    const unsigned magic = (unsigned)((0x100000000ull + (unsigned)Lc - 1) / (unsigned)Lc);   // k / Lc == umulhi(k, magic) for k, Lc < 65536

    // ---- phase A ----
#ifndef PDEGPU_ALR_UNR
#define PDEGPU_ALR_UNR 2
#endif
    constexpr int UNR = PDEGPU_ALR_UNR;                      // pixels in flight per thread
    if (active && !(dbg & 2)) {
        for (int k0 = kfirst; k0 < n; k0 += UNR * KSTEP) {
            PixelRaw<FAM, DIR> raw[UNR];
#pragma unroll
            for (int u = 0; u < UNR; u++) {
                const int k = min(k0 + u * KSTEP, n - 1);    // tail: recompute the last pixel, harmless
                const int i = (DIR & 1) == 0 ? k : line, j = (DIR & 1) == 0 ? line : k;
                raw[u].load(s, j * nr + i, i, j);
            }
#pragma unroll
            for (int u = 0; u < UNR; u++) {
                const int k = min(k0 + u * KSTEP, n - 1);
                const int ch = (int)__umulhi((unsigned)k, magic);
                const int ad = ch * Lp + (k - ch * Lc);
                float a, c, b[2], d[2], m;
                raw[u].rows(a, c, b, d, m);
                Ls[S::A * NP + ad] = a;
                Ls[S::C * NP + ad] = c;
                Ls[S::B1 * NP + ad] = b[qa];
                Ls[S::D1 * NP + ad] = d[qa];
                Ls[S::XO * NP + ad] = raw[u].xo[qa];
                if (F::NUNK == 2) {
                    Ls[S::B2 * NP + ad] = b[qb];
                    Ls[S::D2 * NP + ad] = d[qb];
                    Ls[S::M * NP + ad] = m;
                }
            }
        }
    }
    if (dbg & 2) {                                           // probe: benign constant systems instead of phase A
        for (int t = tid; t < G * S::N * NP; t += NT) smem[t] = 1.0f;
        __syncthreads();
        for (int t = tid; t < G * NP; t += NT) {
            const int gq = t / NP, ad = t - gq * NP;
            float *Lq = smem + (size_t)gq * S::N * NP;
            Lq[S::A * NP + ad] = -1.0f; Lq[S::C * NP + ad] = -1.0f; Lq[S::B1 * NP + ad] = 4.0f;
            if (F::NUNK == 2) { Lq[S::B2 * NP + ad] = 4.0f; Lq[S::M * NP + ad] = 0.1f; }
        }
    }
    __syncthreads();

    // ---- phase B ----
    if (warp < nl && !(dbg & 1)) {
        float *Lw = smem + (size_t)warp * S::N * NP;
        warp_line_solve<false, true>(Lw + S::A * NP, Lw + S::C * NP, Lw + S::B1 * NP, Lw + S::D1 * NP, Lw + S::GP * NP,
                                     nullptr, nullptr, Lw + S::XO * NP, n, Lc, Lp, omega, lane);
        if (F::NUNK == 2) {
            __syncwarp();
            warp_line_solve<true, false>(Lw + S::A * NP, Lw + S::C * NP, Lw + S::B2 * NP, Lw + S::D2 * NP, Lw + S::GP * NP,
                                         Lw + S::M * NP, Lw + S::D1 * NP, nullptr, n, Lc, Lp, omega, lane);
        }
    }
    __syncthreads();

    // ---- phase C ----
    if (active && !(dbg & 4)) {
        for (int k = kfirst; k < n; k += KSTEP) {
            const int i = (DIR & 1) == 0 ? k : line, j = (DIR & 1) == 0 ? line : k;
            const int ip = j * nr + i;
            const int ch = (int)__umulhi((unsigned)k, magic);
            const int ad = ch * Lp + (k - ch * Lc);
            s.x[qa][ip] = Ls[S::D1 * NP + ad];
            if (F::NUNK == 2) {
                float *X2 = s.x[qb];
                X2[ip] = omega * Ls[S::D2 * NP + ad] + (1.0f - omega) * X2[ip];
            }
        }
    }
}

#ifndef PDEGPU_ALR_TPL
#define PDEGPU_ALR_TPL 64
#endif
template <int FAM, int DIR, int G, int TPL = (G >= 8 ? 32 : PDEGPU_ALR_TPL)>
int launch_alr(pdegpu_ctx *ctx, const SysView &v, const pdegpu_system *sys, int colour, float omega,
               int first, int nslots, int n, int Lc, int Lp, size_t smem)
{
    {   // every launch: the attribute is per device and the call is cheap (no static per-ordinal bookkeeping)
    cudaError_t e = cudaFuncSetAttribute(alr_kernel<FAM, DIR, G, TPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e != cudaSuccess) return pdegpu_check_cuda(ctx, e, "cudaFuncSetAttribute(alr_kernel)");
    }
    dim3 grid((nslots + G - 1) / G, sys->batch);
    PDEGPU_PROF(ctx, DIR == 0 ? "alr_kernel<dir0>" : DIR == 1 ? "alr_kernel<dir1>" : "alr_kernel<dir1,transposed>", sweep_bytes<FAM>() * (double)nslots * n * sys->batch);
    alr_kernel<FAM, DIR, G, TPL><<<grid, TPL * G, smem, ctx->stream>>>(v, colour, omega, first, nslots, Lc, Lp, getenv("PDEGPU_DBG") ? atoi(getenv("PDEGPU_DBG")) : 0);
    PDEGPU_LAUNCH_CHECK(ctx, "alr_kernel");
    return PDEGPU_OK;
}

// =============================================================================================
// Pipelined variant: one persistent CTA per SM, warp-specialised.
//   loader warps : phase A of line group t into stage t % S of a shared-memory ring, then phase C of
//                  group t-(S-1) (which the solver warps have finished meanwhile);
//   solver warps : phase B, one warp per line of the stage.
// Stages hand over through mbarriers (full -> solved -> freed), so global loads, line solves and
// stores of different line groups overlap and the HBM stream of an SM never pauses for a solve.
// =============================================================================================
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, unsigned parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "MBAR_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra MBAR_DONE;\n"
                 "bra MBAR_WAIT;\n"
                 "MBAR_DONE:\n"
                 "}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}

template <int FAM, int DIR, int G, int NLW, int S>
__global__ void __launch_bounds__((NLW + G) * 32, 1)
alr_pipe_kernel(SysView s, int colour, float omega, int first_line, int nslots, int Lc, int Lp,
                int groups_per_image, int total_groups)
{
    using F = Fam<FAM>;
    using SL = Slots<F::NUNK>;
    extern __shared__ float smem[];
    constexpr int qa = (F::NUNK == 2 && DIR != 0) ? 1 : 0, qb = 1 - qa;
    constexpr int NT = NLW * 32;                             // loader threads
    const int nr = s.nrows, nc = s.ncols;
    const int n = (DIR & 1) == 0 ? nr : nc;
    const int NP = 32 * Lp;
    const size_t stage_floats = (size_t)G * SL::N * NP;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + (size_t)S * stage_floats);
    unsigned long long *full = bars, *solved = bars + S, *freed = bars + 2 * S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int k = 0; k < S; k++) { mbar_init(&full[k], NT); mbar_init(&solved[k], G * 32); mbar_init(&freed[k], NT); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned magic = (unsigned)((0x100000000ull + (unsigned)Lc - 1) / (unsigned)Lc);
    const int nit = (total_groups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // groups of this CTA

    if (warp < NLW) {
        // ------------------------------- loader / writer warps -------------------------------
        constexpr int TPL = NT / G;
        const int g = (DIR & 1) == 0 ? tid / TPL : tid % G;
        const int kfirst = (DIR & 1) == 0 ? tid % TPL : tid / G;
        constexpr int KSTEP = TPL;
        constexpr int UNR = PDEGPU_ALR_UNR;
        auto write_back = [&](int jt) {
            const int gi = blockIdx.x + jt * gridDim.x;
            const int b = gi / groups_per_image, slot0 = (gi - b * groups_per_image) * G;
            const int line = first_line + colour + 2 * (slot0 + g);
            const int st = jt % S;
            mbar_wait(&solved[st], (jt / S) & 1);
            if (slot0 + g < nslots) {
                const float *Ls = smem + st * stage_floats + (size_t)g * SL::N * NP;
                const int ibase = b * (int)s.bstride;
                for (int k = kfirst; k < n; k += KSTEP) {
                    const int i = (DIR & 1) == 0 ? k : line, j = (DIR & 1) == 0 ? line : k;
                    const int ip = ibase + j * nr + i;
                    const int ch = (int)__umulhi((unsigned)k, magic);
                    const int ad = ch * Lp + (k - ch * Lc);
                    s.x[qa][ip] = Ls[SL::D1 * NP + ad];
                    if (F::NUNK == 2) {
                        float *X2 = s.x[qb];
                        X2[ip] = omega * Ls[SL::D2 * NP + ad] + (1.0f - omega) * X2[ip];
                    }
                }
            }
            mbar_arrive(&freed[st]);
        };
        for (int it = 0; it < nit; it++) {
            const int gi = blockIdx.x + it * gridDim.x;
            const int b = gi / groups_per_image, slot0 = (gi - b * groups_per_image) * G;
            const int line = first_line + colour + 2 * (slot0 + g);
            const int st = it % S;
            if (it >= S) mbar_wait(&freed[st], ((it / S) - 1) & 1);
            if (slot0 + g < nslots) {
                float *Ls = smem + st * stage_floats + (size_t)g * SL::N * NP;
                const int ibase = b * (int)s.bstride;
                for (int k0 = kfirst; k0 < n; k0 += UNR * KSTEP) {
                    PixelRaw<FAM, DIR> raw[UNR];
#pragma unroll
                    for (int u = 0; u < UNR; u++) {
                        const int k = min(k0 + u * KSTEP, n - 1);
                        const int i = (DIR & 1) == 0 ? k : line, j = (DIR & 1) == 0 ? line : k;
                        raw[u].load(s, ibase + j * nr + i, i, j);
                    }
#pragma unroll
                    for (int u = 0; u < UNR; u++) {
                        const int k = min(k0 + u * KSTEP, n - 1);
                        const int ch = (int)__umulhi((unsigned)k, magic);
                        const int ad = ch * Lp + (k - ch * Lc);
                        float a, c, bb[2], d[2], m;
                        raw[u].rows(a, c, bb, d, m);
                        Ls[SL::A * NP + ad] = a;
                        Ls[SL::C * NP + ad] = c;
                        Ls[SL::B1 * NP + ad] = bb[qa];
                        Ls[SL::D1 * NP + ad] = d[qa];
                        Ls[SL::XO * NP + ad] = raw[u].xo[qa];
                        if (F::NUNK == 2) {
                            Ls[SL::B2 * NP + ad] = bb[qb];
                            Ls[SL::D2 * NP + ad] = d[qb];
                            Ls[SL::M * NP + ad] = m;
                        }
                    }
                }
            }
            mbar_arrive(&full[st]);
            if (it >= S - 1) write_back(it - (S - 1));
        }
        for (int jt = max(0, nit - (S - 1)); jt < nit; jt++) write_back(jt);
    } else {
        // ------------------------------------ solver warps ------------------------------------
        const int sw = warp - NLW;
        for (int it = 0; it < nit; it++) {
            const int gi = blockIdx.x + it * gridDim.x;
            const int b = gi / groups_per_image, slot0 = (gi - b * groups_per_image) * G;
            const int st = it % S;
            mbar_wait(&full[st], (it / S) & 1);
            if (slot0 + sw < nslots) {
                float *Lw = smem + st * stage_floats + (size_t)sw * SL::N * NP;
                warp_line_solve<false, true>(Lw + SL::A * NP, Lw + SL::C * NP, Lw + SL::B1 * NP, Lw + SL::D1 * NP, Lw + SL::GP * NP,
                                             nullptr, nullptr, Lw + SL::XO * NP, n, Lc, Lp, omega, lane);
                if (F::NUNK == 2) {
                    __syncwarp();
                    warp_line_solve<true, false>(Lw + SL::A * NP, Lw + SL::C * NP, Lw + SL::B2 * NP, Lw + SL::D2 * NP, Lw + SL::GP * NP,
                                                 Lw + SL::M * NP, Lw + SL::D1 * NP, nullptr, n, Lc, Lp, omega, lane);
                }
            }
            __syncwarp();
            mbar_arrive(&solved[st]);
        }
    }
}

template <int FAM, int DIR, int G, int S>
int launch_alr_pipe(pdegpu_ctx *ctx, const SysView &v, const pdegpu_system *sys, int colour, float omega,
                    int first, int nslots, int n, int Lc, int Lp, size_t smem)
{
    constexpr int NLW = 16;
    {   // every launch: the attribute is per device and the call is cheap (no static per-ordinal bookkeeping)
    cudaError_t e = cudaFuncSetAttribute(alr_pipe_kernel<FAM, DIR, G, NLW, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return pdegpu_check_cuda(ctx, e, "cudaFuncSetAttribute(alr_pipe_kernel)");
    }
    const int gpi = (nslots + G - 1) / G;
    const int total = gpi * sys->batch;
    const int grid = total < ctx->sm_count ? total : ctx->sm_count;
    PDEGPU_PROF(ctx, DIR == 0 ? "alr_pipe_kernel<dir0>" : DIR == 1 ? "alr_pipe_kernel<dir1>" : "alr_pipe_kernel<dir1,transposed>", sweep_bytes<FAM>() * (double)nslots * n * sys->batch);
    alr_pipe_kernel<FAM, DIR, G, NLW, S><<<grid, (NLW + G) * 32, smem, ctx->stream>>>(v, colour, omega, first, nslots, Lc, Lp, gpi, total);
    PDEGPU_LAUNCH_CHECK(ctx, "alr_pipe_kernel");
    return PDEGPU_OK;
}

template <int FAM, int DIR>
int alr_pass(pdegpu_ctx *ctx, const SysView &v, const pdegpu_system *sys, int colour, float omega)
{
    using F = Fam<FAM>;
    constexpr bool INTERIOR_ONLY = F::PDE && F::EIGHT;
    const int nlines = (DIR & 1) == 0 ? sys->ncols : sys->nrows;
    const int n = (DIR & 1) == 0 ? sys->nrows : sys->ncols;
    const int first = INTERIOR_ONLY ? 1 : 0, last = INTERIOR_ONLY ? nlines - 2 : nlines - 1;
    if (first + colour > last) return PDEGPU_OK;
    const int nslots = (last - (first + colour)) / 2 + 1;
    const int Lc = (n + 31) / 32, Lp = Lc | 1;
    const size_t line_bytes = (size_t)Slots<F::NUNK>::N * 32 * Lp * sizeof(float);
    // pipelined persistent kernel when a ring of >= 2 stages x 4 lines fits and there is enough work to stream
    {
        static const int use_pipe = getenv("PDEGPU_ALR_PIPE") ? atoi(getenv("PDEGPU_ALR_PIPE")) : 1;
        const size_t room = 227 * 1024 - 256;
        const long long groups = (long long)((nslots + 3) / 4) * sys->batch;
        const bool idx_ok = (long long)sys->batch * sys->batch_stride < (1ll << 31);
        if (use_pipe && idx_ok && groups >= 2 * ctx->sm_count && 2 * 4 * line_bytes <= room) {
            if (3 * 4 * line_bytes <= room)
                return launch_alr_pipe<FAM, DIR, 4, 3>(ctx, v, sys, colour, omega, first, nslots, n, Lc, Lp, 3 * 4 * line_bytes + 256);
            return launch_alr_pipe<FAM, DIR, 4, 2>(ctx, v, sys, colour, omega, first, nslots, n, Lc, Lp, 2 * 4 * line_bytes + 256);
        }
    }
    // G lines per CTA: as many as keep >= 2 CTAs per SM resident; lines along j (DIR 1) are strided in
    // memory, so they want >= 4 neighbouring lines per CTA to use the sectors they fetch.
    int G = 8;
    while (G > 1 && G * line_bytes > (size_t)kMaxSmem / 2) G >>= 1;
    if ((DIR & 1) == 1 && G < 4) { G = 4; while (G > 1 && G * line_bytes > (size_t)kMaxSmem) G >>= 1; }
    if (G * line_bytes > (size_t)kMaxSmem) return PDEGPU_ERR_UNSUPPORTED;   // line too long for shared memory
    const size_t smem = G * line_bytes;
    switch (G) {
    case 8: return launch_alr<FAM, DIR, 8>(ctx, v, sys, colour, omega, first, nslots, n, Lc, Lp, smem);
    case 4: return launch_alr<FAM, DIR, 4>(ctx, v, sys, colour, omega, first, nslots, n, Lc, Lp, smem);
    case 2: return launch_alr<FAM, DIR, 2>(ctx, v, sys, colour, omega, first, nslots, n, Lc, Lp, smem);
    default: return launch_alr<FAM, DIR, 1>(ctx, v, sys, colour, omega, first, nslots, n, Lc, Lp, smem);
    }
}

}  // namespace

// Batched transpose of dense column-major fields: dst[b][i*ncols + j] = src[b*sstride + j*nrows + i].
static __global__ void __launch_bounds__(256)
transpose_kernel(float *__restrict__ dst, const float *__restrict__ src, int nrows, int ncols, long long sstride, long long dstride)
{
    __shared__ float tile[32][33];
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    src += (long long)blockIdx.z * sstride;
    dst += (long long)blockIdx.z * dstride;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + tx, j = j0 + r;
        if (i < nrows && j < ncols) tile[r][tx] = src[(long long)j * nrows + i];
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int j = j0 + tx, i = i0 + r;
        if (i < nrows && j < ncols) dst[(long long)i * ncols + j] = tile[tx][r];
    }
}

int transpose_fields(pdegpu_ctx *ctx, float *dst, const float *src, int nrows, int ncols, int batch, long long sstride, long long dstride)
{
    dim3 grid((nrows + 31) / 32, (ncols + 31) / 32, batch);
    PDEGPU_PROF(ctx, "transpose_kernel", 0);
    transpose_kernel<<<grid, 256, 0, ctx->stream>>>(dst, src, nrows, ncols, sstride, dstride);
    PDEGPU_LAUNCH_CHECK(ctx, "transpose_kernel");
    return PDEGPU_OK;
}

namespace {

// One ALR iteration = lines along i (both colours), then lines along j (both colours). Lines along j are
// strided in memory: a CTA that owns a few of them touches one 32-B sector per element and uses half of
// it (zebra), and the sweep becomes L1-wavefront-bound (measured: 2x the DRAM bytes and 2x the time of the
// contiguous direction). So the row pass runs on a TRANSPOSED copy of the problem, where its lines are
// contiguous too: coefficients are transposed once per call, the unknowns twice per iteration (16 B/px
// each way, against ~150 B/px for a sweep).
template <int FAM>
int alr_run(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega)
{
    using F = Fam<FAM>;
    SysView v = make_view(sys);
    if (F::PDE && F::EIGHT) iter = 1;            // pdeSolvers.c:362
    // refuse up front (before touching the unknowns) if either direction does not fit
    if ((long long)sys->nrows * sys->ncols >= (1ll << 31) || sys->nrows >= 65536 || sys->ncols >= 65536) return PDEGPU_ERR_UNSUPPORTED;
    {
        const size_t per = (size_t)Slots<F::NUNK>::N * 32 * sizeof(float);
        const int L0 = ((sys->nrows + 31) / 32) | 1, L1 = ((sys->ncols + 31) / 32) | 1;
        if (per * L0 > (size_t)kMaxSmem || per * L1 > (size_t)kMaxSmem) return PDEGPU_ERR_UNSUPPORTED;
    }
    constexpr int NN = F::EIGHT ? 8 : 4;
    const long long npix = (long long)sys->nrows * sys->ncols;
    static const int use_xpose = getenv("PDEGPU_ALR_XPOSE") ? atoi(getenv("PDEGPU_ALR_XPOSE")) : 1;
    const int nfields = NN + 3 * F::NUNK + (F::NUNK == 2 ? 1 : 0) + (F::LATE ? F::NUNK : 0);
    bool xpose = use_xpose && sys->nrows >= 8 && sys->ncols >= 8;
    pdegpu_system tsys = *sys;
    float *xT[2] = {nullptr, nullptr};
    int rc;
    if (xpose) {
        if (pdegpu_scratch_reserve(ctx, (size_t)nfields * npix * sys->batch * sizeof(float)) != PDEGPU_OK) xpose = false;
    }
    if (xpose) {
        float *p = (float *)ctx->scratch;
        auto take = [&]() { float *r = p; p += npix * sys->batch; return r; };
        auto tr = [&](const float *src) -> float * {
            float *d = take();
            rc = transpose_fields(ctx, d, src, sys->nrows, sys->ncols, sys->batch, sys->batch_stride, npix);
            return d;
        };
        // neighbour roles swap under transposition: N<->W, S<->E, NE<->SW
        static const int perm[8] = {W_N, W_W, W_S, W_E, W_NW, W_SW, W_SE, W_NE};
        rc = PDEGPU_OK;
        tsys.nrows = sys->ncols; tsys.ncols = sys->nrows; tsys.batch_stride = npix;
        for (int n = 0; n < NN && !rc; n++) tsys.w[n] = tr(sys->w[perm[n]]);
        for (int q = 0; q < F::NUNK && !rc; q++) {
            tsys.c[q] = tr(sys->c[q]);
            if (!rc) tsys.d[q] = tr(sys->d[q]);
            if (F::LATE && !rc) tsys.x0[q] = tr(sys->x0[q]);
            xT[q] = take();
            tsys.x[q] = xT[q];
        }
        if (F::NUNK == 2 && !rc) tsys.m = tr(sys->m);
        if (rc) return rc;
    }
    SysView vt = make_view(&tsys);
    for (int it = 0; it < iter; it++) {
        // interior-only lines (8-neighbour PDE) start at line 1: colour 1 there = the even lines, relaxed first as everywhere
        constexpr int cflip = (F::PDE && F::EIGHT) ? 1 : 0;
        for (int cc = 0; cc < 2; cc++) if ((rc = alr_pass<FAM, 0>(ctx, v, sys, cc ^ cflip, omega))) return rc;
        if (xpose) {
            for (int q = 0; q < F::NUNK; q++)
                if ((rc = transpose_fields(ctx, xT[q], sys->x[q], sys->nrows, sys->ncols, sys->batch, sys->batch_stride, npix))) return rc;
            for (int cc = 0; cc < 2; cc++) if ((rc = alr_pass<FAM, 2>(ctx, vt, &tsys, cc ^ cflip, omega))) return rc;
            for (int q = 0; q < F::NUNK; q++)
                if ((rc = transpose_fields(ctx, sys->x[q], xT[q], sys->ncols, sys->nrows, sys->batch, npix, sys->batch_stride))) return rc;
        } else {
            for (int cc = 0; cc < 2; cc++) if ((rc = alr_pass<FAM, 1>(ctx, v, sys, cc ^ cflip, omega))) return rc;
        }
    }
    return PDEGPU_OK;
}

}  // namespace

int relax_window_line(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega);
int relax_tline(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega);        // sweeps_tline.cu

int relax_stream_line(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega)
{
    // generation 2 (sliding window, sweeps_window.cu) where it has a kernel; PDEGPU_ALR_WINDOW=0 disables it
    static const int use_window = getenv("PDEGPU_ALR_WINDOW") ? atoi(getenv("PDEGPU_ALR_WINDOW")) : 1;
    // generation 3 (TMA-fed, total-field form, sweeps_tline.cu) first; PDEGPU_ALR_GEN=2 keeps generation 2
    static const int gen = getenv("PDEGPU_ALR_GEN") ? atoi(getenv("PDEGPU_ALR_GEN")) : 3;
    if (use_window && gen >= 3) {
        const int rc = relax_tline(ctx, sys, iter, omega);
        if (rc != PDEGPU_ERR_UNSUPPORTED) return rc;
    }
    if (use_window) {
        const int rc = relax_window_line(ctx, sys, iter, omega);
        if (rc != PDEGPU_ERR_UNSUPPORTED) return rc;
    }
    switch (sys->family) {
    case PDEGPU_FLOW_ELIN4: return alr_run<PDEGPU_FLOW_ELIN4>(ctx, sys, iter, omega);
    case PDEGPU_FLOW_LLIN4: return alr_run<PDEGPU_FLOW_LLIN4>(ctx, sys, iter, omega);
    case PDEGPU_FLOW_LLIN8: return alr_run<PDEGPU_FLOW_LLIN8>(ctx, sys, iter, omega);
    case PDEGPU_DISP_LLIN4: return alr_run<PDEGPU_DISP_LLIN4>(ctx, sys, iter, omega);
    case PDEGPU_PDE4:       return alr_run<PDEGPU_PDE4>(ctx, sys, iter, omega);
    case PDEGPU_PDE8:       return alr_run<PDEGPU_PDE8>(ctx, sys, iter, omega);
    default: return PDEGPU_ERR_UNSUPPORTED;
    }
}
