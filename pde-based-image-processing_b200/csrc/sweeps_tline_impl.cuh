// sweeps_tline_impl.cuh -- the pass kernel of generation 3 (see tline_common.cuh for the formulation).
//
// One launch = one complete direction pass (both zebra colours, all unknowns, every line of every problem),
// one persistent CTA per SM owning a contiguous range of blocks of BL lines, as in generation 2. New here:
//
//   * every operand reaches shared memory through the TMA unit (cp.async.bulk, one elected thread, mbarrier
//     transaction counts): no register is held across the memory latency, the bytes in flight are bounded by the
//     shared memory set aside for them instead of by the register file;
//   * PACKED LINES: the preparation kernel stores the NC coefficient lines of a line next to each other, so a task's
//     coefficients are ONE bulk copy (17 KB at 480 elements) into one of K slabs handed out in task order
//     (full/empty mbarriers). A consumer warp turns a slab into tridiagonal rows directly in registers
//     (lane = chunk of M consecutive elements, M odd: conflict-free), so there is no assembler stage and no row
//     buffer. (First version: one copy per field -- the elected thread needed 3400 cycles per task to issue them,
//     twice the time the memory system needs for the line: profiles/r02_tline_probes.txt);
//   * T RING: the unknowns of R consecutive lines (one copy per line), loaded in line order by a second elected
//     thread. A line's entry first holds T_old (what the even lines need from their odd neighbours), is overwritten
//     IN PLACE by the warp that solves the line, then serves the odd neighbours (T_new of even lines) and finally
//     the transposed write of its block of BL lines. One ring instead of generation 2's ring + neighbour loads;
//   * tasks are handed out dynamically (a shared counter) in the schedule's order; the producer writes a
//     task descriptor next to each slab, so consumers never decode the schedule; producer and loader walk the
//     schedule incrementally (no division on their critical path).
//
// Schedule (as generation 2): e0 e2 .. e(2D-2) | e(2D) o1 e(2D+2) o3 ...: the even lines run D pairs ahead of the
// odd ones; the first even line of the next CTA's range is solved redundantly. Every wait is on a task earlier in
// this order or on a producer that itself only waits on earlier tasks; R >= 2D + BL keeps the ring loader from
// waiting on a block whose lines it has not loaded yet (tools/tline_schedule_sim.py replays the protocol).
#pragma once
#include "tline_common.cuh"
#include <stdlib.h>

namespace {

// Phase probes (build with -DTL_PROBE; never in the shipped library): cycles per role and phase, summed over all CTAs.
// consumers [0..15]: 0 wait slab, 1 wait ring, 2 wait even neighbours, 3 rows, 4 solve, 5 relax + store,
//   7 publish, 8 block write, 14 tasks, 15 total;  producer [16..23]: 16 wait empty, 17 issue, 22 tasks, 23 total;
// loader [24..31]: 24 wait free, 25 issue, 30 lines, 31 total
#ifdef TL_PROBE
__device__ unsigned long long g_tl_probe[32];
#define TLP_DECL unsigned long long pr_[16] = {0}; long long pt_ = clock64(); const long long pt0_ = pt_
#define TLP(k) do { const long long n_ = clock64(); pr_[k] += (unsigned long long)(n_ - pt_); pt_ = n_; } while (0)
#define TLP_COUNT(k) do { pr_[k] += 1; } while (0)
#define TLP_FLUSH(base, nk) do { pr_[(nk) - 1] = (unsigned long long)(clock64() - pt0_); for (int k_ = 0; k_ < (nk); k_++) atomicAdd(&g_tl_probe[(base) + k_], pr_[k_]); } while (0)
#else
#define TLP_DECL
#define TLP(k)
#define TLP_COUNT(k)
#define TLP_FLUSH(base, nk)
#endif

// rows of both unknowns live in registers (8 arrays of M): fewer, fatter warps
// (ptxas sizes the register budget for the thread count rounded up to 128: 384 threads -> 168 registers, 256 -> 255)
template <int M> struct TLThreads { static constexpr int value = M <= 9 ? 512 : M <= 17 ? 384 : 256; };

// position of a local line l (0-based inside the CTA's range of blocks) in the batch, advanced without divisions
struct LinePos {
    int l, r, lb, jb, img;             // local line, line inside its block, local block, block of the problem, problem
    __device__ __forceinline__ void init(int l0, int B0, int NB, int BL)
    {
        l = l0; lb = l0 / BL; r = l0 - lb * BL;
        const int gb = B0 + lb;
        img = gb / NB; jb = gb - img * NB;
    }
    __device__ __forceinline__ void advance(int step, int NB, int BL)     // step <= BL
    {
        l += step; r += step;
        if (r >= BL) { r -= BL; lb++; if (++jb == NB) { jb = 0; img++; } }
    }
    __device__ __forceinline__ int j(int BL) const { return BL * jb + r; }
};

// MODE 0: diagonal = sum of the existing weights (+ D where the constant term is a number): flow, disparity.
// MODE 1: diagonal = D (TRACE) where it is a number, else the sum of the weights: pdeSolvers.c.
template <int NUNK, int NN, int MODE, int M>
__global__ void __launch_bounds__(TLThreads<M>::value, 1)
tline_pass_kernel(const TLParams p)
{
    constexpr int NC = 6 + (NUNK == 2 ? 3 : 0) + (NN == 8 ? 4 : 0);
    constexpr unsigned FULLMASK = 0xffffffffu;
    // slab rows in the order the preparation kernel packs them (tline_common.cuh)
    constexpr int rWP = 0, rWN = 1, rWL = 2, rWH = 3, rC = 4, rD = 5;   // C, D of the unknown solved first; second: +3 (6 = M)
    constexpr int rMM = 6;
    constexpr int rDG = NUNK == 2 ? 9 : 6;                    // first of the four diagonal weights (LP, LN, HP, HN)
    extern __shared__ __align__(128) float smem[];
    const int R = p.R, D = p.D, K = p.K, BL = p.BL, NBR = p.NBR, NCW = p.NCW;
    const int P = p.pitch;                                    // floats between the fields of a packed line
    const int SP = p.SP;                                      // floats between ring entries (NUNK * P, padded to skew banks)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *ring = smem;
    // 32M - n floats of padding (rounded to 16 bytes) after the ring and after the slabs: lanes past the end of a line read them
    constexpr int PADF = 32 * M;
    const int pad = PADF - P;
    float *slabs = ring + (size_t)R * SP + pad;
    uint64_t *bars = reinterpret_cast<uint64_t *>(slabs + (size_t)K * NC * P + pad);
    uint64_t *full = bars, *empty = bars + K, *rfull = bars + 2 * K, *solved = bars + 2 * K + R;
    int *desc = reinterpret_cast<int *>(bars + 2 * K + 2 * R);          // 8 ints per slab (16-byte aligned: 2K + 2R is even)
    unsigned *written_seq = reinterpret_cast<unsigned *>(desc + 8 * K);
    unsigned *block_cnt = written_seq + NBR;
    unsigned *grab = block_cnt + NBR;
    // monotone counters: tasks whose slab has been armed / ring lines that have been armed. A consumer looks at them
    // before it waits on a phase PARITY, which by itself cannot tell use u from use u-2 of a slot.
    unsigned *issued = grab + 1, *loaded = grab + 2;

    if (threadIdx.x == 0) {
        for (int k = 0; k < K; k++) { mbar_init(&full[k], 1); mbar_init(&empty[k], 1); }
        for (int k = 0; k < R; k++) { mbar_init(&rfull[k], 1); mbar_init(&solved[k], 1); }
        for (int k = 0; k < 2 * NBR; k++) written_seq[k] = 0;
        grab[0] = 0; grab[1] = 0; grab[2] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int nlines = p.nlines;
    const int B0 = (int)((long long)blockIdx.x * p.TB / gridDim.x), B1 = (int)((long long)(blockIdx.x + 1) * p.TB / gridDim.x);
    const int nblk = B1 - B0;
    const bool redundant = B1 < p.TB && (B1 % p.NB) != 0;
    const int Ltot = BL * nblk + (redundant ? 1 : 0);         // task lines are l = 0 .. Ltot-1 (those that exist)

    if (warp < NCW) {
        // =================================== consumer warps ===================================
        const float omega = p.omega, om1 = 1.0f - p.omega;
        const int o = lane * M;
        TLP_DECL;
        for (;;) {
            TLP(8);
            unsigned s = 0;
            if (lane == 0) s = atomicAdd(grab, 1u);
            s = __shfl_sync(FULLMASK, s, 0);
            const int slot = (int)(s % (unsigned)K);
            warp_wait_ge(issued, s + 1, lane);
            mbar_wait(&full[slot], (s / (unsigned)K) & 1u);
            const int4 d0 = *reinterpret_cast<const int4 *>(desc + slot * 8);
            const int4 d1 = *reinterpret_cast<const int4 *>(desc + slot * 8 + 4);
            const int l = d0.x;
            TLP(0);
            if (l < 0) {                                      // terminator: no more tasks
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[slot]);
                break;
            }
            const int flags = d0.y, pi = d0.z;                 // pi = problem of the pass = image * S + segment
            const int seg = flags >> 16;
            const int n = seg == p.S - 1 ? p.nlast : p.ns;
            const bool cutL = seg > 0, cutR = seg < p.S - 1;
            // values across the cuts, from the start of the pass (global memory; used after the row formulas)
            const bool odd = flags & 1, eLo = flags & 2, eHi = flags & 4, owned = flags & 8, relaxed = !(flags & 16);
            float hL0 = 0.f, hL1 = 0.f, hR0 = 0.f, hR1 = 0.f;
            // 8-neighbour stencils: the diagonal neighbours across a cut, [lo/hi line][left/right cut][unknown]
            float hD[2][2][NUNK] = {};
            if (lane == 0 && (cutL || cutR)) {
                const long long ls = (long long)NUNK * P;
                const float *tl = p.tin + ((long long)(pi - 1) * nlines + d0.w) * ls + (p.ns - 1);
                const float *tr = p.tin + ((long long)(pi + 1) * nlines + d0.w) * ls;
                const int u0 = (NUNK == 2 ? p.q0 : 0) * P, u1 = (NUNK == 2 ? (p.q0 ^ 1) : 0) * P;
                if (cutL) { hL0 = tl[u0]; if (NUNK == 2) hL1 = tl[u1]; }
                if (cutR) { hR0 = tr[u0]; if (NUNK == 2) hR1 = tr[u1]; }
                if (NN == 8) {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        if (!(h == 0 ? eLo : eHi)) continue;
                        const long long dl = (h == 0 ? -1 : 1) * ls;
                        if (cutL) { hD[h][0][0] = tl[dl + u0]; if (NUNK == 2) hD[h][0][NUNK - 1] = tl[dl + u1]; }
                        if (cutR) { hD[h][1][0] = tr[dl + u0]; if (NUNK == 2) hD[h][1][NUNK - 1] = tr[dl + u1]; }
                    }
                }
            }
            const int gh = l + BL + 1;
            // ring slots g % R, (g -+ 1) % R and their use counts g / R, (g -+ 1) / R (g = l + BL)
            const int rs = d1.z, rlo = rs == 0 ? R - 1 : rs - 1, rhi = rs == R - 1 ? 0 : rs + 1;
            const unsigned ug = (unsigned)d1.w, ulo = rs == 0 ? ug - 1 : ug, uhi = rs == R - 1 ? ug + 1 : ug;
            float *own = ring + (size_t)rs * SP;
            const float *lo = ring + (size_t)rlo * SP, *hi = ring + (size_t)rhi * SP;   // (own entry instead where the line does not exist)
            warp_wait_ge(loaded, (unsigned)gh + 1, lane);
            mbar_wait(&rfull[rs], ug & 1u);
            if (relaxed) {
                if (eLo) mbar_wait(&rfull[rlo], ulo & 1u);
                if (eHi) mbar_wait(&rfull[rhi], uhi & 1u);
                TLP(1);
                if (odd) {                                    // new values of the even neighbours
                    mbar_wait(&solved[rlo], ulo & 1u);
                    if (eHi) mbar_wait(&solved[rhi], uhi & 1u);
                }
                const float *sl = slabs + (size_t)slot * NC * P;
                TLP(2);
                TLP_COUNT(14);
                // Rows of BOTH unknowns in one pass over the slab, which is then given back at once (a slab is the scarce
                // resource: held through the first solve, as at first, the producer spent 2/3 of its time waiting for a
                // free one). The second unknown's rows wait in registers (as, bs, ds, ms) while the first is solved; both
                // solves run through the same code (one loop iteration each: one copy in the instruction cache).
                const int t0 = (NUNK == 2 ? p.q0 : 0) * P, t1 = (NUNK == 2 ? (p.q0 ^ 1) : 0) * P;   // where the unknowns live in a packed T line
                // Neighbours that do not exist are taken out of the SLAB (this warp owns it until it gives it back) instead
                // of out of every row: the in-line weights at the two ends of the line, and, for the first / last line of a
                // problem, the whole row of weights towards the missing line, whose values are then read from this line's
                // own entry (any finite numbers: a ring slot that was never loaded may hold NaN bits, and 0 * NaN = NaN).
                // The row formulas below are then free of per-neighbour selects, and free of index clamps: shared memory
                // is padded so that lanes past the end of a short line read (and discard) whatever follows.
                {
                    float *sw_ = const_cast<float *>(sl);
                    if (NN == 4 && lane == 0) {               // (at a cut the neighbour exists: see below)
                        if (!cutL) sw_[rWP * P] = 0.f;
                        if (!cutR) sw_[rWN * P + n - 1] = 0.f;
                    }
                    // (8-neighbour families mask the ends in the row formulas: the NaN-TRACE diagonal of pdeSolvers.c:1179
                    // wants the weights as they are, and the diagonal neighbours across a cut are added after the loop)
                    if (!eLo || !eHi) {
                        for (int e = lane; e < n; e += 32) {
                            if (!eLo) { sw_[rWL * P + e] = 0.f; if (NN == 8) { sw_[(rDG + 0) * P + e] = 0.f; sw_[(rDG + 1) * P + e] = 0.f; } }
                            if (!eHi) { sw_[rWH * P + e] = 0.f; if (NN == 8) { sw_[(rDG + 2) * P + e] = 0.f; sw_[(rDG + 3) * P + e] = 0.f; } }
                        }
                        if (!eLo) lo = own;
                        if (!eHi) hi = own;
                    }
                    __syncwarp();
                }
                float a[M], b[M], c[M], d[M];
                float as[NUNK == 2 ? M : 1], bs[NUNK == 2 ? M : 1], ds[NUNK == 2 ? M : 1], ms[NUNK == 2 ? M : 1];
#pragma unroll
                for (int k = 0; k < M; k++) {
                    const int e = o + k;
                    const bool ok = e < n;
                    float wp = sl[rWP * P + e], wn = sl[rWN * P + e];
                    const float wl = sl[rWL * P + e], wh = sl[rWH * P + e];
                    const float wp_raw = wp, wn_raw = wn;
                    if (NN == 8) {                            // ends of the line (not of a segment): no neighbour
                        if (e == 0 && !cutL) wp = 0.f;
                        if (e == n - 1 && !cutR) wn = 0.f;
                    }
                    float sw = (wl + wh) + (wp + wn);
                    float cr0 = wl * lo[t0 + e] + wh * hi[t0 + e];
                    float cr1 = 0.f;
                    if (NUNK == 2) cr1 = wl * lo[t1 + e] + wh * hi[t1 + e];
                    float quirk = 0.f;
                    if (NN == 8) {
                        // diagonal neighbours: elements e-1 / e+1 of the two lines; none at the first / last element of the
                        // SEGMENT here (those across a cut are added after the loop)
                        const bool eP = e > 0, eN = e < n - 1;
                        const int em = eP ? e - 1 : 0, ep = eN ? e + 1 : e;
                        const float rlp = sl[(rDG + 0) * P + e], rln = sl[(rDG + 1) * P + e];
                        const float rhp = sl[(rDG + 2) * P + e], rhn = sl[(rDG + 3) * P + e];
                        const float wlp = eP ? rlp : 0.f, wln = eN ? rln : 0.f, whp = eP ? rhp : 0.f, whn = eN ? rhn : 0.f;
                        sw += (wlp + wln) + (whp + whn);
                        cr0 += (wlp * lo[t0 + em] + wln * lo[t0 + ep]) + (whp * hi[t0 + em] + whn * hi[t0 + ep]);
                        if (NUNK == 2) cr1 += (wlp * lo[t1 + em] + wln * lo[t1 + ep]) + (whp * hi[t1 + em] + whn * hi[t1 + ep]);
                        // pdeSolvers.c:1179,1209,1238 / :1314,1344,1373: where TRACE is NaN the diagonal is the four
                        // axial weights + wNW twice + wSW + wSE, whatever the position (wNE never). In line roles: column
                        // pass NW SW SE = LP LN HN, row pass = LP HP HN
                        if (MODE == 1) quirk = ((wp_raw + wn_raw) + (wl + wh)) + ((rlp + rlp) + ((p.q0 ? rhp : rln) + rhn));
                    }
                    const float C = sl[rC * P + e], Dd = sl[rD * P + e];
                    float bb, dd;
                    if (MODE == 1) {
                        const bool has = !is_nan(Dd);
                        bb = has ? Dd : (NN == 8 ? quirk : sw);
                        dd = has ? cr0 + C : cr0;
                    } else {
                        const bool has = !is_nan(C);
                        bb = has ? sw + Dd : sw;
                        float t = C;
                        if (NUNK == 2) t -= sl[rMM * P + e] * own[t1 + e];        // coupling to the OLD second unknown
                        dd = has ? cr0 + t : cr0;
                    }
                    a[k] = ok ? -wp : 0.f; c[k] = ok ? -wn : 0.f;
                    b[k] = ok ? bb : 1.0f; d[k] = ok ? dd : 0.f;
                    if (NUNK == 2) {
                        const float C1 = sl[(rC + 3) * P + e], D1 = sl[(rD + 3) * P + e];
                        const bool has = ok && !is_nan(C1);
                        as[k] = a[k];
                        bs[k] = ok ? (has ? sw + D1 : sw) : 1.0f;
                        ds[k] = ok ? (has ? cr1 + C1 : cr1) : 0.f;
                        ms[k] = has ? sl[rMM * P + e] : 0.f;                       // coupling to the NEW first unknown, added after its solve
                    }
                }
                if (cutL || cutR) {
                    // ends of a segment that are not ends of the line: the weight stays in the diagonal, the neighbour's
                    // value (start of the pass) goes to the right-hand side, the row is cut off from the recurrence
                    hR0 = __shfl_sync(FULLMASK, hR0, 0); hR1 = __shfl_sync(FULLMASK, hR1, 0);
                    if (cutL && lane == 0) {
                        d[0] -= a[0] * hL0; a[0] = 0.f;
                        if (NUNK == 2) { ds[0] -= as[0] * hL1; as[0] = 0.f; }
                        if (NN == 8) {                        // its diagonal neighbours on the lines j-1 / j+1
                            const float rlp = sl[(rDG + 0) * P], rhp = sl[(rDG + 2) * P];
                            d[0] += rlp * hD[0][0][0] + rhp * hD[1][0][0];
                            if (NUNK == 2) ds[0] += rlp * hD[0][0][NUNK - 1] + rhp * hD[1][0][NUNK - 1];
                            if (MODE == 0) { b[0] += rlp + rhp; if (NUNK == 2) bs[0] += rlp + rhp; }
                        }
                    }
                    if (cutR) {
                        float dl0 = 0.f, dl1 = 0.f, dh0 = 0.f, dh1 = 0.f, rln = 0.f, rhn = 0.f;
                        if (NN == 8) {
                            dl0 = __shfl_sync(FULLMASK, hD[0][1][0], 0); dh0 = __shfl_sync(FULLMASK, hD[1][1][0], 0);
                            if (NUNK == 2) { dl1 = __shfl_sync(FULLMASK, hD[0][1][NUNK - 1], 0); dh1 = __shfl_sync(FULLMASK, hD[1][1][NUNK - 1], 0); }
                            rln = sl[(rDG + 1) * P + n - 1]; rhn = sl[(rDG + 3) * P + n - 1];
                        }
#pragma unroll
                        for (int k = 0; k < M; k++)
                            if (o + k == n - 1) {
                                d[k] -= c[k] * hR0;
                                if (NUNK == 2) ds[k] -= c[k] * hR1;
                                c[k] = 0.f;
                                if (NN == 8) {
                                    d[k] += rln * dl0 + rhn * dh0;
                                    if (NUNK == 2) ds[k] += rln * dl1 + rhn * dh1;
                                    if (MODE == 0) { b[k] += rln + rhn; if (NUNK == 2) bs[k] += rln + rhn; }
                                }
                            }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[slot]);     // rows are in registers: the slab can be refilled
                TLP(3);
#pragma unroll 1
                for (int q = 0; q < NUNK; q++) {
                    const int tq = q == 0 ? t0 : t1;
                    chunk_solve<M>(a, c, b, d, lane);
                    TLP(4);
#pragma unroll
                    for (int k = 0; k < M; k++) {
                        const int e = o + k;
                        const float t = omega * d[k] + om1 * own[tq + e];
                        if (e < n) own[tq + e] = t;
                        // (lanes past the end of the line: t is made of whatever they read, and 0 * NaN = NaN)
                        if (NUNK == 2) { a[k] = as[k]; b[k] = bs[k]; d[k] = e < n ? ds[k] - ms[k] * t : 0.f; }
                    }
                    TLP(5);
                }
            } else {                                          // a line that is not relaxed (8-neighbour PDE border lines): T_out = T_in
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[slot]);
            }
            __syncwarp();
            const int lb = d1.y;
            unsigned done = 0;
            if (lane == 0) {
                __threadfence_block();
                mbar_arrive(&solved[rs]);
                if (owned) done = atomicAdd(&block_cnt[lb % NBR], 1u) + 1;
            }
            done = __shfl_sync(FULLMASK, done, 0);
            TLP(7);
            const int cnt = (flags >> 8) & 0xff;
            if (owned && (int)done == cnt) {
                // this warp completed block lb: write its lines to T_out (transposed: the packed T lines of the next pass)
                __threadfence_block();
                const int j0 = d1.x;
                const float *rblk = ring + (size_t)(((lb + 1) * BL) % R) * SP;
                const long long ostep = (long long)NUNK * p.opitch;               // floats between consecutive elements of a line
                // element k of this segment is line seg*ns + k of the other direction; lines j0.. are elements of its
                // segment j0 / ons (a block never straddles two: ons is a multiple of BL)
                const int img = pi / p.S, oseg = j0 / p.ons;
                float *obase = p.tout + (((long long)img * p.oS + oseg) * p.nfull + (long long)seg * p.ns) * ostep + (j0 - oseg * p.ons);
#pragma unroll 1
                for (int qq = 0; qq < NUNK; qq++) {
                    float *out = obase + (long long)qq * p.opitch;
                    const float *rq = rblk + qq * P;
                    if (cnt == BL && BL == 8) {
                        // half-warp h writes lines 4h..4h+3 of 16 consecutive elements: full 32-byte sectors
                        const int h = lane >> 4;
                        const float *rh = rq + (size_t)(4 * h) * SP;
#pragma unroll 4
                        for (int i = lane & 15; i < n; i += 16) {
                            float4 v;
                            v.x = rh[i]; v.y = rh[SP + i]; v.z = rh[2 * SP + i]; v.w = rh[3 * SP + i];
                            *reinterpret_cast<float4 *>(out + i * ostep + 4 * h) = v;
                        }
                    } else if (cnt == BL && BL == 4) {
#pragma unroll 4
                        for (int i = lane; i < n; i += 32) {
                            float4 v;
                            v.x = rq[i]; v.y = rq[SP + i]; v.z = rq[2 * SP + i]; v.w = rq[3 * SP + i];
                            *reinterpret_cast<float4 *>(out + i * ostep) = v;
                        }
                    } else {
                        const int k = lane & 7;
                        for (int i = lane >> 3; i < n; i += 4)
                            if (k < cnt) out[i * ostep + k] = rq[(size_t)k * SP + i];
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    block_cnt[lb % NBR] = 0;
                    st_release(&written_seq[lb % NBR], (unsigned)lb + 1);
                }
                __syncwarp();
            }
        }
#ifdef TL_PROBE
        if (lane == 0) TLP_FLUSH(0, 16);
#endif
    } else if (warp == NCW) {
        // ============================ producer of the coefficient slabs ============================
        if (lane == 0) {
            const unsigned bytes = (unsigned)(NC * P) * 4u;
            const long long line_stride = (long long)NC * P;                      // floats between consecutive packed lines
            TLP_DECL;
            // the even stream (it runs D pairs ahead) and the odd stream, each advanced by 2 lines per task
            LinePos pe, po;
            pe.init(0, B0, p.NB, BL);
            po.init(1, B0, p.NB, BL);
            unsigned s = 0;
            int slot = 0;
            unsigned sphase = 0;                              // parity of the slot's use count
            // ring slot / use count of a line follow from g = l + BL; both streams keep theirs incrementally
            int rse = BL % R, rso = (BL + 1) % R;
            unsigned use_e = (unsigned)(BL / R), use_o = (unsigned)((BL + 1) / R);
            const int Q = D + 2 * ((Ltot + 1) >> 1);
            for (int q = 0; q < Q; q++) {
                const bool odd = q >= D && ((q - D) & 1);
                LinePos &lp = odd ? po : pe;
                int &rsl = odd ? rso : rse;
                unsigned &usl = odd ? use_o : use_e;
                const int l = lp.l, j = lp.j(BL);
                if (l < Ltot && j < nlines) {
                    const bool noupdate = p.skip_border && (j == 0 || j == nlines - 1);
                    const int j0 = j - lp.r;
                    const int cnt = min(BL, nlines - j0);
                    TLP(1);
                    mbar_wait(&empty[slot], sphase ^ 1u);
                    TLP(0);
                    TLP_COUNT(6);
                    int *dsc = desc + slot * 8;
                    *reinterpret_cast<int4 *>(dsc) = make_int4(l, (odd ? 1 : 0) | (j > 0 ? 2 : 0) | (j + 1 < nlines ? 4 : 0) | (lp.lb < nblk ? 8 : 0)
                                                                  | (noupdate ? 16 : 0) | (cnt << 8) | ((lp.img % p.S) << 16), lp.img, j);
                    *reinterpret_cast<int4 *>(dsc + 4) = make_int4(j0, lp.lb, rsl, (int)usl);
                    if (noupdate) mbar_arrive(&full[slot]);
                    else {
                        fence_proxy_async();
                        mbar_expect_tx(&full[slot], bytes);
                        bulk_g2s(slabs + (size_t)slot * NC * P, p.coef + ((long long)lp.img * nlines + j) * line_stride, bytes, &full[slot]);
                    }
                    s++;
                    st_release(issued, s);
                    if (++slot == K) { slot = 0; sphase ^= 1u; }
                }
                lp.advance(2, p.NB, BL);
                rsl += 2;
                if (rsl >= R) { rsl -= R; usl++; }
            }
            for (int t = 0; t < NCW; t++) {                   // one terminator per consumer warp
                mbar_wait(&empty[slot], sphase ^ 1u);
                desc[slot * 8] = -1;
                mbar_arrive(&full[slot]);
                s++;
                st_release(issued, s);
                if (++slot == K) { slot = 0; sphase ^= 1u; }
            }
            TLP(1);
            TLP_FLUSH(16, 8);
        }
    } else if (warp == NCW + 1) {
        // =================================== loader of the T ring ===================================
        if (lane == 0) {
            const unsigned bytes = (unsigned)(NUNK * P) * 4u;
            const long long line_stride = (long long)NUNK * P;
            const int Gmax = Ltot + BL;                       // ring lines g = l + BL for the local lines l = -1 .. Ltot
            TLP_DECL;
            {   // g < BL - 1: no such lines; g = BL - 1: the line before the range, if it belongs to the same problem
                const int gb = B0 - 1;
                bool exists = false;
                int img = 0, j = 0;
                if (gb >= 0) {
                    img = gb / p.NB;
                    const int jb = gb - img * p.NB;
                    if (jb != p.NB - 1) { exists = true; j = BL * jb + BL - 1; }
                }
                for (int g = 0; g < BL - 1; g++) { mbar_arrive(&rfull[g % R]); mbar_arrive(&solved[g % R]); }
                const int slot = (BL - 1) % R;
                if (exists) {
                    fence_proxy_async();
                    mbar_expect_tx(&rfull[slot], bytes);
                    bulk_g2s(ring + (size_t)slot * SP, p.tin + ((long long)img * nlines + j) * line_stride, bytes, &rfull[slot]);
                } else mbar_arrive(&rfull[slot]);
                mbar_arrive(&solved[slot]);
                st_release(loaded, (unsigned)BL);
            }
            LinePos lp;
            lp.init(0, B0, p.NB, BL);
            int slot = BL % R;
            int lbp = 0, rp = 0;                              // block / position in it of the slot's previous line l - R (once l >= R)
            for (int g = BL; g <= Gmax; g++) {
                const int l = g - BL;
                TLP(1);
                if (l - R == -1) mbar_wait(&solved[BL % R], (unsigned)(BL / R) & 1u);   // the line before the range: read by task 0 only
                if (l >= R) {                                 // the slot's previous line must be written out and unused
                    if (lbp < nblk) {
                        thread_wait_ge(&written_seq[lbp % NBR], (unsigned)lbp + 1);
                        if (lbp > 0) thread_wait_ge(&written_seq[(lbp - 1) % NBR], (unsigned)lbp);
                    }
                    if (++rp == BL) { rp = 0; lbp++; }
                }
                TLP(0);
                TLP_COUNT(6);
                const int j = lp.j(BL);
                bool exists = l <= Ltot && j < nlines && (B0 + lp.lb) < p.TB;
                if (exists && l >= BL * nblk && !redundant) exists = false;
                const bool istask = exists && l < Ltot;
                if (exists) {
                    fence_proxy_async();
                    mbar_expect_tx(&rfull[slot], bytes);
                    bulk_g2s(ring + (size_t)slot * SP, p.tin + ((long long)lp.img * nlines + j) * line_stride, bytes, &rfull[slot]);
                } else mbar_arrive(&rfull[slot]);
                if (!istask) mbar_arrive(&solved[slot]);
                st_release(loaded, (unsigned)g + 1);
                lp.advance(1, p.NB, BL);
                if (++slot == R) slot = 0;
            }
            TLP(1);
            TLP_FLUSH(24, 8);
        }
    }
}

}  // namespace
