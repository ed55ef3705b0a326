// sweeps_lex_impl.cuh -- solver 2 in the REFERENCE'S ORDER (pdegpu_set_sweep_order(ctx, PDEGPU_ORDER_REFERENCE)).
//
// The reference relaxes the lines of a direction one after the other, each seeing the NEW values of the line before it
// and the OLD values of the line after it (e.g. westColumn/middleColumn/eastColumn_llin4, opticalflowSolvers.c:2415-2760,
// called from GS_ALR_SOR_llin4_2d :733-754): a chain with no parallelism across lines. One such sweep carries
// information across the whole image, which the drivers that do not re-warp (Horn-Schunck, FMG) rely on at their
// default iteration counts. This kernel runs exactly that chain, so its iterates agree with the
// reference's sweep by sweep, not only at convergence:
//
//   * one CTA per problem; the parallelism is the batch (one problem per CTA, several CTAs per SM) and, inside a line,
//     the 32 lanes of the solver warp (partitioned Thomas: lane = chunk of Mr consecutive unknowns, interface system by
//     parallel cyclic reduction -- chunk_solve of window_common.cuh with the rows in shared memory instead of registers,
//     so that Mr is a run-time number and lines of any length the shared memory holds are solved exactly, uncut);
//   * operands reach shared memory as in generation 3 (tline_common.cuh): packed coefficient lines, ONE bulk copy
//     (cp.async.bulk + mbarrier) per line issued by a producer thread one or two lines ahead of the solver;
//   * T lines live in a ring of RL entries: line j-1 (new), j (old, overwritten in place), j+1 (old), j+2 (in flight);
//     the solved line goes back to global memory in the SAME layout (in place); a transposition kernel turns the packed T
//     lines of one direction into those of the other between the passes;
//   * unknown order inside a pass: the reference runs "all lines of U, then all lines of V" (column pass) and "V, then U"
//     (row pass). A row of V couples to U only at its own pixel, so "line j of U, then line j of V" visits the same
//     values in the same state: the two passes are fused line by line.
#pragma once
#include "tline_common.cuh"

namespace {

struct LexParams {
    const float *coef;             // packed coefficient lines of this direction (S = 1: whole lines)
    float       *t;                // packed T lines of this direction, relaxed in place
    int pitch;                     // floats between the fields of a packed line
    int q0;                        // flow families: physical index of the unknown the pass solves first; scalar: 1 = row pass
    int n, nlines;
    int Mr;                        // elements per lane (odd)
    int KS, RL;                    // coefficient slabs, ring lines
    int inslab;                    // long lines: the eliminated rows live in the slab rows only their solver warp reads (see the kernel)
    int skip_border;               // 8-neighbour PDE: first and last line of a problem are not relaxed
    float omega;
};

// (the chain of dependent operations per line is what bounds this path: the reciprocal is the approximate one of the
// zebra kernels unless built with -DPDEGPU_LEX_EXACT_RCP)
__device__ __forceinline__ float lex_rcp(float x)
{
#ifdef PDEGPU_LEX_EXACT_RCP
    return __frcp_rn(x);
#else
    return fast_rcp(x);                                       // MUFU.RCP, 1 ulp: the sweep-by-sweep bar (1e-5) is measured with it
#endif
}

// Threads: warp 0 solves the unknown the pass takes first (all lines), warp 1 the second one (flow families; it runs
// one line behind warp 0: a row of the second unknown needs the NEW first unknown of its own line only), the last warp
// is the producer. Two chains of dependent line solves overlap instead of one chain of twice the length.
template <int NUNK> struct LexThreads { static constexpr int value = 32 * (NUNK + 1); };

template <int NUNK, int NN, int MODE, bool INSLAB>
__global__ void __launch_bounds__(LexThreads<NUNK>::value)
lex_pass_kernel(const LexParams p)
{
    constexpr int NC = 6 + (NUNK == 2 ? 3 : 0) + (NN == 8 ? 4 : 0);
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int rWP = 0, rWN = 1, rWL = 2, rWH = 3, rC = 4, rD = 5, rMM = 6;
    constexpr int rDG = NUNK == 2 ? 9 : 6;
    extern __shared__ __align__(128) float smem[];
    const int P = p.pitch, n = p.n, nlines = p.nlines, Mr = p.Mr, KS = p.KS, RL = p.RL;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int TP = NUNK * P;                                  // floats per ring entry
    float *ring = smem;
    float *slabs = ring + (size_t)RL * TP;
    // per solver warp: the eliminated rows (spike, super-diagonal, right-hand side), one float4 per element. LONG LINES
    // (inslab): two slabs + ring + 16 bytes per element and unknown do not fit 227 KB at 1920 elements, and with ONE slab
    // the second solver warp cannot overlap with the first (it works one line = one slab behind). There the super-diagonal
    // and the right-hand side overwrite the slab rows C_q, D_q -- read by solver warp q only, once per element, just
    // before -- and only the spike keeps a 4-byte scratch: two slabs fit again.
    constexpr bool inslab = INSLAB;                           // (compile time: a run-time flag in the three row loops cost 45 %)
    float4 *rowsAll = reinterpret_cast<float4 *>(slabs + (size_t)KS * NC * P);
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<float *>(rowsAll) + (size_t)NUNK * 32 * Mr * (inslab ? 1 : 4));
    uint64_t *sfull = bars, *sempty = bars + KS, *rfull = bars + 2 * KS, *rempty = bars + 2 * KS + RL;
    unsigned *first_done = reinterpret_cast<unsigned *>(bars + 2 * KS + 2 * RL);   // lines the first solver warp has finished

    if (threadIdx.x == 0) {
        for (int k = 0; k < KS; k++) { mbar_init(&sfull[k], 1); mbar_init(&sempty[k], NUNK); }
        for (int k = 0; k < RL; k++) { mbar_init(&rfull[k], 1); mbar_init(&rempty[k], NUNK); }
        *first_done = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const long long pb = (long long)blockIdx.x * nlines;     // first packed line of this problem
    const float *coef = p.coef + pb * NC * P;
    float *tg = p.t + pb * TP;

    if (warp == NUNK) {
        // ======================================= producer =======================================
        if (lane == 0) {
            const unsigned sbytes = (unsigned)(NC * P) * 4u, tbytes = (unsigned)TP * 4u;
            // ring line g lives in slot g % RL; line 0 first, then (slab j, line j + 1) in turn
            fence_proxy_async();
            mbar_expect_tx(&rfull[0], tbytes);
            bulk_g2s(ring, tg, tbytes, &rfull[0]);
            for (int j = 0; j < nlines; j++) {
                const int ks = j % KS;
                if (j >= KS) mbar_wait_lazy(&sempty[ks], (unsigned)(j / KS - 1) & 1u);
                fence_proxy_async();
                mbar_expect_tx(&sfull[ks], sbytes);
                bulk_g2s(slabs + (size_t)ks * NC * P, coef + (long long)j * NC * P, sbytes, &sfull[ks]);
                const int g = j + 1;
                if (g < nlines) {
                    const int rs = g % RL;
                    if (g >= RL) mbar_wait_lazy(&rempty[rs], (unsigned)(g / RL - 1) & 1u);
                    fence_proxy_async();
                    mbar_expect_tx(&rfull[rs], tbytes);
                    bulk_g2s(ring + (size_t)rs * TP, tg + (long long)g * TP, tbytes, &rfull[rs]);
                }
            }
        }
        return;
    }

    // ========================================= solver warps =========================================
    const int q = warp;                                       // 0: the unknown solved first, 1: the second
    const bool last_solver = q == NUNK - 1;                   // writes the finished line back
    float4 *rows = rowsAll + (size_t)q * 32 * Mr;                                   // (not inslab)
    float *spike = reinterpret_cast<float *>(rowsAll) + (size_t)q * 32 * Mr;        // (inslab)
    const float omega = p.omega, om1 = 1.0f - p.omega;
    const int o = lane * Mr;
    const int t0 = (NUNK == 2 ? p.q0 : 0) * P, t1 = (NUNK == 2 ? (p.q0 ^ 1) : 0) * P;
    const int tq = q == 0 ? t0 : t1, tother = q == 0 ? t1 : t0;
    const int cq = rC + 3 * q, dq = rD + 3 * q;
    for (int j = 0; j < nlines; j++) {
        const int ks = j % KS, rs = j % RL;
        const bool eLo = j > 0, eHi = j + 1 < nlines;
        const bool relaxed = !(p.skip_border && (!eLo || !eHi));
        float *own = ring + (size_t)rs * TP;
        const float *lo = eLo ? ring + (size_t)((j - 1) % RL) * TP : own;
        const float *hi = eHi ? ring + (size_t)((j + 1) % RL) * TP : own;
        const float *sl = slabs + (size_t)ks * NC * P;
        float *slw = slabs + (size_t)ks * NC * P;             // (inslab: rows cq, dq are this warp's to overwrite)
        auto row_put = [&](int e, float a_, float b_, float d_) {
            if (!inslab) rows[e] = make_float4(a_, b_, d_, 0.f);
            else if (e < n) { spike[e] = a_; slw[cq * P + e] = b_; slw[dq * P + e] = d_; }
        };
        auto row_get = [&](int e) -> float4 {                 // (.x spike, .y super-diagonal, .z right-hand side)
            if (!inslab) return rows[e];
            if (e < n) return make_float4(spike[e], slw[cq * P + e], slw[dq * P + e], 0.f);
            return make_float4(0.f, 0.f, 0.f, 0.f);           // identity rows past the end of the line
        };
        mbar_wait(&sfull[ks], (unsigned)(j / KS) & 1u);
        mbar_wait(&rfull[rs], (unsigned)(j / RL) & 1u);
        if (eHi) mbar_wait(&rfull[(j + 1) % RL], (unsigned)((j + 1) / RL) & 1u);
        if (NUNK == 2 && q == 1) warp_wait_ge(first_done, (unsigned)j + 1, lane);     // the new first unknown of this line
        if (relaxed) {
            // ---- rows + local forward elimination: afterwards row k reads  x_k + .y*x_{k+1} + .x*x_left = .z
            float ap = 0.f, bp = 0.f, dp = 0.f;
#pragma unroll 2
            for (int k = 0; k < Mr; k++) {
                const int e = o + k;
                const bool ok = e < n;
                const int ec = ok ? e : n - 1;
                const bool eP = ec > 0, eN = ec < n - 1;
                const float wpr = sl[rWP * P + ec], wnr = sl[rWN * P + ec];
                const float wp = eP ? wpr : 0.f, wn = eN ? wnr : 0.f;
                const float wl = eLo ? sl[rWL * P + ec] : 0.f, wh = eHi ? sl[rWH * P + ec] : 0.f;
                float sw = (wl + wh) + (wp + wn);
                float cr = wl * lo[tq + ec] + wh * hi[tq + ec];
                float quirk = 0.f;
                if (NN == 8) {
                    const int em = eP ? ec - 1 : ec, ep = eN ? ec + 1 : ec;
                    const float rlp = sl[(rDG + 0) * P + ec], rln = sl[(rDG + 1) * P + ec];
                    const float rhp = sl[(rDG + 2) * P + ec], rhn = sl[(rDG + 3) * P + ec];
                    const float wlp = (eLo && eP) ? rlp : 0.f, wln = (eLo && eN) ? rln : 0.f;
                    const float whp = (eHi && eP) ? rhp : 0.f, whn = (eHi && eN) ? rhn : 0.f;
                    sw += (wlp + wln) + (whp + whn);
                    cr += (wlp * lo[tq + em] + wln * lo[tq + ep]) + (whp * hi[tq + em] + whn * hi[tq + ep]);
                    // NaN-TRACE diagonal of pdeSolvers.c:1179, see sweeps_tline_impl.cuh
                    if (MODE == 1) quirk = ((wpr + wnr) + (sl[rWL * P + ec] + sl[rWH * P + ec])) + ((rlp + rlp) + ((p.q0 ? rhp : rln) + rhn));
                }
                const float C = sl[cq * P + ec], Dd = sl[dq * P + ec];
                float bb, dd;
                if (MODE == 1) {
                    const bool has = !is_nan(Dd);
                    bb = has ? Dd : (NN == 8 ? quirk : sw);
                    dd = has ? cr + C : cr;
                } else {
                    const bool has = !is_nan(C);
                    bb = has ? sw + Dd : sw;
                    float t = C;
                    if (NUNK == 2) t -= sl[rMM * P + ec] * own[tother + ec];     // the other unknown as it is NOW: old for the
                    dd = has ? cr + t : cr;                                     // first solver, new for the second
                }
                const float a = ok ? -wp : 0.f, c = ok ? -wn : 0.f;
                const float b = ok ? bb : 1.0f, d = ok ? dd : 0.f;
                const float inv = lex_rcp(k == 0 ? b : b - a * bp);
                const float an = k == 0 ? a * inv : (-a * ap) * inv;
                const float dn = k == 0 ? d * inv : (d - a * dp) * inv;
                bp = c * inv; ap = an; dp = dn;
                row_put(e, ap, bp, dp);
            }
            // ---- first unknown of the chunk in terms of the last one and x_left
            float Af, Bf, Gf;
            if (Mr == 1) { Af = 0.f; Bf = -1.0f; Gf = 0.f; }
            else {
                const float4 r0 = row_get(o + Mr - 2);
                Af = r0.z; Bf = r0.y; Gf = r0.x;
                for (int r = Mr - 3; r >= 0; r--) {
                    const float4 rr = row_get(o + r);
                    Af = rr.z - rr.y * Af;
                    Bf = -rr.y * Bf;
                    Gf = rr.x - rr.y * Gf;
                }
            }
            // ---- interface system in the lanes' last unknowns, parallel cyclic reduction
            float An = __shfl_down_sync(FULL, Af, 1), Bn = __shfl_down_sync(FULL, Bf, 1), Gn = __shfl_down_sync(FULL, Gf, 1);
            if (lane == 31) { An = 0.f; Bn = 0.f; Gn = 0.f; }
            float al = ap, be = 1.0f - bp * Gn, ga = -bp * Bn, de = dp - bp * An;
#pragma unroll
            for (int st = 1; st < 32; st <<= 1) {
                float alm = __shfl_up_sync(FULL, al, st), bem = __shfl_up_sync(FULL, be, st);
                float gam = __shfl_up_sync(FULL, ga, st), dem = __shfl_up_sync(FULL, de, st);
                float alp = __shfl_down_sync(FULL, al, st), bep = __shfl_down_sync(FULL, be, st);
                float gap = __shfl_down_sync(FULL, ga, st), dep = __shfl_down_sync(FULL, de, st);
                if (lane < st)      { alm = 0.f; bem = 1.0f; gam = 0.f; dem = 0.f; }
                if (lane + st > 31) { alp = 0.f; bep = 1.0f; gap = 0.f; dep = 0.f; }
                const float k1 = al * lex_rcp(bem), k2 = ga * lex_rcp(bep);
                be = be - gam * k1 - alp * k2;
                de = de - dem * k1 - dep * k2;
                al = -alm * k1;
                ga = -gap * k2;
            }
            const float l = de * lex_rcp(be);
            float L = __shfl_up_sync(FULL, l, 1);
            if (lane == 0) L = 0.f;
            // ---- local back substitution + SOR, in place in the ring
            float x = l;
            for (int r = Mr - 1; r >= 0; r--) {
                const int e = o + r;
                if (r < Mr - 1) { const float4 rr = row_get(e); x = rr.z - rr.y * x - rr.x * L; }
                if (e < n) own[tq + e] = omega * x + om1 * own[tq + e];
            }
        }
        __syncwarp();
        if (NUNK == 2 && q == 0) {
            if (lane == 0) { __threadfence_block(); st_release(first_done, (unsigned)j + 1); }
        }
        if (last_solver && relaxed) {
            // the relaxed line back to global memory (pads are never touched: they stay as the preparation wrote them)
            for (int qq = 0; qq < NUNK; qq++) {
                const float4 *s4 = reinterpret_cast<const float4 *>(own + qq * P);
                float4 *g4 = reinterpret_cast<float4 *>(tg + (long long)j * TP + qq * P);
                for (int e = lane; e < (P >> 2); e += 32) g4[e] = s4[e];
            }
            __syncwarp();
        }
        if (lane == 0) {
            mbar_arrive(&sempty[ks]);
            if (eLo) mbar_arrive(&rempty[(j - 1) % RL]);     // line j-1 has served this warp's last reader
        }
    }
}

// packed T lines of one direction -> packed T lines of the other: element i of line j (unknown q) -> element j of line i.
// in: [problem][nl_in lines][NUNK][pitch_in], out: [problem][n_in lines][NUNK][pitch_out]; pads of `out` are written as zeros.
static __global__ void __launch_bounds__(256)
lex_transpose_kernel(float *__restrict__ out, const float *__restrict__ in, int n_in, int nl_in, int pitch_in, int pitch_out, int nunk)
{
    __shared__ float tile[32][33];
    const int b = blockIdx.z / nunk, q = blockIdx.z - b * nunk;
    const int e0 = blockIdx.x * 32, l0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float *src = in + ((long long)b * nl_in * nunk + q) * pitch_in;
    float *dst = out + ((long long)b * n_in * nunk + q) * pitch_out;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int l = l0 + ty + 8 * r, e = e0 + tx;
        tile[ty + 8 * r][tx] = (l < nl_in && e < n_in) ? src[(long long)l * nunk * pitch_in + e] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int e = e0 + ty + 8 * r, l = l0 + tx;           // output line e, element l
        if (e < n_in && l < pitch_out) dst[(long long)e * nunk * pitch_out + l] = tile[tx][ty + 8 * r];
    }
}

}  // namespace
