// sweeps_window.cu -- kernel generation 2, solver 2: zebra line relaxation as a SLIDING WINDOW.
//
// One launch = one complete direction pass (BOTH colours, both unknowns) over a batch of problems,
// reading every field exactly once and writing every unknown exactly once:
//
//   * The lines of the pass are contiguous in memory (the row pass runs on a transposed copy of the
//     problem, see alr_window_run). A persistent CTA (one per SM) owns a contiguous range of 8-line
//     blocks of the batch and walks through it line by line.
//   * Zebra order "all even lines, then all odd lines" only constrains neighbours: an odd line needs the
//     NEW values of its two even neighbours, an even line the OLD values of its two odd neighbours. So the
//     CTA relaxes  e0 e2 .. e(2D-2) | e(2D) o1 e(2D+2) o3 ...  -- mathematically the same sweep -- and each
//     line's coefficients are touched once. The pass is out of place (reads X_in, writes X_out), so old
//     values are always available from global memory; new even values are handed to the odd lines through
//     a ring of solved lines in shared memory. The last odd line of a CTA's range needs the first even line
//     of the next range: that one line is solved redundantly (same inputs, same code => same bits).
//   * One WARP per line does everything for it: coalesced global loads (element e = 32t + lane), the
//     reference's tridiagonal row per pixel (line_rows.cuh), a transposition through a private
//     shared-memory scratch to "lane = chunk of M consecutive unknowns", then a partitioned Thomas solve
//     held entirely in registers (local elimination with a left spike, a 32-unknown interface system
//     solved by parallel cyclic reduction on shuffles, local back substitution), SOR, and the result into
//     the ring. 5-8 warps of a CTA work on different lines at any time; warps synchronise only through
//     per-line "solved" and per-block "written" sequence numbers in shared memory.
//   * X_out is written in the TRANSPOSED layout, 8 lines at a time (one full 32-B sector per element), by
//     the warp that completes a block. The next direction pass wants exactly that layout, so alternating
//     directions costs no transposition of the unknowns at all.
//
// HBM traffic per pass = algorithmic: (#fields read + #unknowns written) * 4 B per pixel (+ 1 line per
// CTA range). Same fixed point and same ordering semantics as generation 0/1 (even lines first).
#include "window_common.cuh"
#include <stdlib.h>

namespace {

// M is odd: lane L's chunk [L*M, L*M+M) is read with stride M (conflict-free), and a line sits in shared
// memory in natural order, so 4 consecutive elements are one aligned float4.
template <int FAM, int DIR, int M>
__global__ void __launch_bounds__(kWinMaxWarps * 32, 1)
alr_window_kernel(const WinParams p)
{
    using F = Fam<FAM>;
    using RF = RowF<F::NUNK>;
    static_assert(M & 1, "chunk length must be odd");
    constexpr int NUNK = F::NUNK;
    constexpr int qa = (NUNK == 2 && DIR != 0) ? 1 : 0, qb = 1 - qa;
    constexpr int LS = 32 * M;                   // floats per line (padded with identity rows)
    constexpr int P = LS;                        // ring pitch per unknown
    constexpr int SP = NUNK * P + 4;             // ring pitch per line: = 4 mod 8, so the two half-warps of a block write hit disjoint banks
    constexpr int NT = ((LS + 127) / 128 + 1) & ~1;   // batches of 128 elements (4 per lane), made even for the double buffer
    extern __shared__ float smem[];
    const int R = p.R, D = p.D, NBR = R >> 3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    float *ring = smem;
    float *scratch = ring + (size_t)R * SP + (size_t)warp * RF::N * LS;
    unsigned *flags = reinterpret_cast<unsigned *>(ring + (size_t)R * SP + (size_t)nwarps * RF::N * LS);
    unsigned *solved_seq = flags, *written_seq = flags + R, *block_cnt = flags + R + NBR;
    for (int t = threadIdx.x; t < R + 2 * NBR; t += blockDim.x) flags[t] = 0;
    __syncthreads();

    const int n = p.n, nlines = p.nlines;
    const SysView &s = p.s;
    // this CTA's range of blocks
    const int B0 = (int)((long long)blockIdx.x * p.TB / gridDim.x), B1 = (int)((long long)(blockIdx.x + 1) * p.TB / gridDim.x);
    const int nblk = B1 - B0;
    const bool redundant = B1 < p.TB && (B1 % p.NB) != 0;    // first even line of the next range, solved here too
    const int Ltot = 8 * nblk + (redundant ? 1 : 0);
    const int Q = D + 2 * ((Ltot + 1) >> 1);
    const float omega = p.omega, om1 = 1.0f - p.omega;

    // schedule entry q -> line (evens run D pairs ahead of the odds); false if q names no line
    auto decode = [&](int q, WinTask &T) -> bool {
        if (q < D) { T.l = 2 * q; T.odd = false; }
        else {
            const int r = q - D;
            if (r & 1) { T.l = r; T.odd = true; } else { T.l = 2 * D + r; T.odd = false; }
        }
        if (T.l >= Ltot) return false;
        const int lb = T.l >> 3, gb = B0 + lb;
        T.img = gb / p.NB; T.jb = gb - T.img * p.NB;
        T.j = 8 * T.jb + (T.l & 7);
        if (T.j >= nlines) return false;
        T.owned = lb < nblk;
        T.ibase = T.img * (int)s.bstride + T.j * n;
        T.eW = T.j > 0; T.eE = T.j + 1 < nlines;
        T.dW = T.eW ? -n : 0; T.dE = T.eE ? n : 0;
        return true;
    };
    const int ec_last = (n - 1) & ~3;                        // first element of the last (possibly partial) vector of a line
    const bool al = p.aligned == 2;
    auto first_element = [&](int t, int &e0, int &ec) { e0 = 128 * t + 4 * lane; ec = min(e0, ec_last); };

    RawBatch<FAM> raw[2];
    WinTask cur, nxt;
    int q = warp;
    while (q < Q && !decode(q, cur)) q += nwarps;
    if (q < Q) { int e0, ec; first_element(0, e0, ec); if (al) raw[0].template issue<true>(s, cur, ec, n); else raw[0].template issue<false>(s, cur, ec, n); }

    while (q < Q) {
        int q2 = q + nwarps;
        while (q2 < Q && !decode(q2, nxt)) q2 += nwarps;
        const int l = cur.l, lb = l >> 3;
        // the ring slot of this line must be free: its previous occupant (line l-R) has been written out and
        // consumed by its odd neighbours, i.e. its block and the block before are written
        if (l >= R) {
            const int lbp = (l - R) >> 3;
            warp_wait_ge(&written_seq[lbp % NBR], (unsigned)lbp + 1, lane);
            if (lbp > 0) warp_wait_ge(&written_seq[(lbp - 1) % NBR], (unsigned)lbp, lane);
        }
        float *rs = ring + (size_t)(l % R) * SP;
        const float *rsW = ring + (size_t)((l + R - 1) % R) * SP, *rsE = ring + (size_t)((l + 1) % R) * SP;

        // ---- load + row assembly: lane holds elements 128*t + 4*lane .. +3; the loads of the next batch (or of
        //      the first batch of this warp's next line) are issued before this batch is touched ----
#pragma unroll
        for (int t = 0; t < NT; t++) {
            {
                int e0n, ecn;
                if (t + 1 < NT) { first_element(t + 1, e0n, ecn); if (e0n < LS) { if (al) raw[(t + 1) & 1].template issue<true>(s, cur, ecn, n); else raw[(t + 1) & 1].template issue<false>(s, cur, ecn, n); } }
                else if (q2 < Q) { first_element(0, e0n, ecn); if (al) raw[0].template issue<true>(s, nxt, ecn, n); else raw[0].template issue<false>(s, nxt, ecn, n); }
            }
            int e0, ec;
            first_element(t, e0, ec);
            if (t == 0 && cur.odd) {
                // new values of the even neighbours
                warp_wait_ge(&solved_seq[(l - 1) % R], (unsigned)l, lane);
                if (cur.eE) warp_wait_ge(&solved_seq[(l + 1) % R], (unsigned)l + 2, lane);
            }
            if (e0 < LS) {
                RawBatch<FAM> &rb = raw[t & 1];
                if (cur.odd) rb.neighbours_from_ring(rsW, rsE, P, ec, n);
                float ra[4], rc[4], rb1[4], rd1[4], rb2[4], rd2[4], rm[4], xo0[4], xo1[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    PixelRaw<FAM, DIR> r;
                    const bool valid = e0 + k < n;
                    rb.template pixel<DIR>(k, ec + k, n, cur.eW, cur.eE, r);
                    float a, c, b[2], d[2], m;
                    r.rows(a, c, b, d, m);
                    ra[k] = valid ? a : 0.f; rc[k] = valid ? c : 0.f;
                    rb1[k] = valid ? b[qa] : 1.0f; rd1[k] = valid ? d[qa] : 0.f;
                    rb2[k] = valid ? b[qb] : 1.0f; rd2[k] = valid ? d[qb] : 0.f;
                    rm[k] = valid ? m : 0.f;
                    xo0[k] = valid ? r.xo[0] : 0.f; xo1[k] = (valid && NUNK == 2) ? r.xo[NUNK - 1] : 0.f;
                }
                st4(scratch + RF::A * LS + e0, ra);
                st4(scratch + RF::C * LS + e0, rc);
                st4(scratch + RF::B1 * LS + e0, rb1);
                st4(scratch + RF::D1 * LS + e0, rd1);
                st4(rs + e0, xo0);
                if (NUNK == 2) {
                    st4(scratch + RF::B2 * LS + e0, rb2);
                    st4(scratch + RF::D2 * LS + e0, rd2);
                    st4(scratch + RF::MM * LS + e0, rm);
                    st4(rs + P + e0, xo1);
                }
            }
        }
        __syncwarp();

        // ---- solve, lane = chunk ----
        {
            const int o = lane * M;
            float a[M], c[M], b[M], d[M];
#pragma unroll
            for (int k = 0; k < M; k++) {
                a[k] = scratch[RF::A * LS + o + k]; c[k] = scratch[RF::C * LS + o + k];
                b[k] = scratch[RF::B1 * LS + o + k]; d[k] = scratch[RF::D1 * LS + o + k];
            }
            chunk_solve<M>(a, c, b, d, lane);
#pragma unroll
            for (int k = 0; k < M; k++) {
                d[k] = omega * d[k] + om1 * rs[qa * P + o + k];
                rs[qa * P + o + k] = d[k];
            }
            if (NUNK == 2) {
#pragma unroll
                for (int k = 0; k < M; k++) {
                    a[k] = scratch[RF::A * LS + o + k];
                    b[k] = scratch[RF::B2 * LS + o + k];
                    d[k] = scratch[RF::D2 * LS + o + k] - scratch[RF::MM * LS + o + k] * d[k];
                }
                chunk_solve<M>(a, c, b, d, lane);
#pragma unroll
                for (int k = 0; k < M; k++) rs[qb * P + o + k] = omega * d[k] + om1 * rs[qb * P + o + k];
            }
        }
        __syncwarp();
        unsigned done = 0;
        if (lane == 0) {
            st_release(&solved_seq[l % R], (unsigned)l + 1);
            __threadfence_block();
            if (cur.owned) done = atomicAdd(&block_cnt[lb % NBR], 1u) + 1;
        }
        done = __shfl_sync(0xffffffffu, done, 0);
        const int j0 = 8 * cur.jb, cnt = min(8, nlines - j0);
        if (cur.owned && (int)done == cnt) {
            // ---- this warp completed block lb: write its lines to X_out (transposed layout) ----
            __threadfence_block();
            const float *rblk = ring + (size_t)((8 * lb) % R) * SP;
#pragma unroll
            for (int qq = 0; qq < NUNK; qq++) {
                float *o = p.xout[qq] + (long long)cur.img * p.ostride + j0;
                const float *rq = rblk + qq * P;
                if (cnt == 8 && p.vec_ok == 2) {
                    // half-warp h writes lines 4h..4h+3 of 16 consecutive elements: full 32-B sectors
                    const int h = lane >> 4;
                    const float *rh = rq + (size_t)(4 * h) * SP;
#pragma unroll 2
                    for (int i = lane & 15; i < n; i += 16) {
                        float4 v;
                        v.x = rh[i]; v.y = rh[SP + i]; v.z = rh[2 * SP + i]; v.w = rh[3 * SP + i];
                        *reinterpret_cast<float4 *>(o + (long long)i * nlines + 4 * h) = v;
                    }
                } else {
                    const int k = lane & 7;
                    for (int i = lane >> 3; i < n; i += 4)
                        if (k < cnt) o[(long long)i * nlines + k] = rq[(size_t)k * SP + i];
                }
            }
            __syncwarp();
            if (lane == 0) {
                block_cnt[lb % NBR] = 0;
                st_release(&written_seq[lb % NBR], (unsigned)lb + 1);
            }
            __syncwarp();
        }
        q = q2;
        cur = nxt;
    }
}

struct WinGeom { int M, LS, SP, R, D, NW; size_t smem; };

// Ring + per-warp scratch must fit in one SM's shared memory. With NW lines in flight, the evens must run
// D pairs ahead of the odds with 2D-1 >= NW (an odd line then finds its even neighbours solved), and the
// ring must span the evens' lead, the lines in flight and a block waiting to be written: R >= 2D + NW + 9
// (which also gives R > 7 + 2D, the condition for the schedule to be free of deadlock, tools/window_schedule_sim.py).
static bool win_geometry(int n, int nunk, WinGeom &g)
{
    static const int Ms[] = {5, 9, 15, 21, 25};
    g.M = 0;
    for (int m : Ms) if (32 * m >= n) { g.M = m; break; }
    if (!g.M || n < 8) return false;
    g.LS = 32 * g.M; g.SP = nunk * g.LS + 4;
    const int rowf = nunk == 2 ? 7 : 4;
    const size_t room = 227 * 1024;
    for (int nw = kWinMaxWarps; nw >= 3; nw--) {
        const int D = (nw + 2) / 2, R = (2 * D + nw + 9 + 7) & ~7;
        const size_t bytes = ((size_t)R * g.SP + (size_t)nw * rowf * g.LS + R + 2 * (R / 8)) * sizeof(float);
        if (bytes <= room) { g.NW = nw; g.R = R; g.D = D; g.smem = bytes; return true; }
    }
    return false;
}

template <int FAM, int DIR, int M>
int launch_window(pdegpu_ctx *ctx, const WinParams &p, const WinGeom &g, int batch)
{
    {   // every launch: the attribute is per device and the call is cheap (no static per-ordinal bookkeeping)
    cudaError_t e = cudaFuncSetAttribute(alr_window_kernel<FAM, DIR, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return pdegpu_check_cuda(ctx, e, "cudaFuncSetAttribute(alr_window_kernel)");
    }
    const int grid = p.TB < ctx->sm_count ? p.TB : ctx->sm_count;
    PDEGPU_PROF(ctx, DIR == 0 ? "alr_window_kernel<dir0>" : "alr_window_kernel<dir1,transposed>",
                sweep_bytes<FAM>() * (double)p.n * p.nlines * batch);
    alr_window_kernel<FAM, DIR, M><<<grid, g.NW * 32, g.smem, ctx->stream>>>(p);
    PDEGPU_LAUNCH_CHECK(ctx, "alr_window_kernel");
    return PDEGPU_OK;
}

}  // namespace
int window2_pass(pdegpu_ctx *ctx, int family, int dir, void *params, int M, int batch);   // sweeps_window2.cu
namespace {

template <int FAM, int DIR>
int window_pass(pdegpu_ctx *ctx, const pdegpu_system *sys, float *const xout[2], long long ostride, float omega)
{
    using F = Fam<FAM>;
    WinGeom g;
    if (!win_geometry(sys->nrows, F::NUNK, g)) return PDEGPU_ERR_UNSUPPORTED;
    WinParams p;
    p.s = make_view(sys);
    p.xout[0] = xout[0]; p.xout[1] = xout[1];
    p.ostride = ostride;
    p.n = sys->nrows; p.nlines = sys->ncols;
    p.NB = (sys->ncols + 7) / 8; p.TB = p.NB * sys->batch;
    p.R = g.R; p.D = g.D;
    p.omega = omega;
    p.G = 1;
    auto oal = [&](int vw) {                                 // X_out stores: 2 = 16-byte, 1 = 8-byte (two-stage kernel), 0 = scalar
        bool a = (sys->ncols % vw == 0) && (ostride % vw == 0);
        for (int q = 0; q < F::NUNK; q++) a = a && ((uintptr_t)xout[q] % (4 * vw) == 0);
        return a;
    };
    p.vec_ok = oal(4) ? 2 : oal(2) ? 1 : 0;
    {   // vector loads: 16-byte ones need every field 16-B aligned and lines / problems a multiple of 4 floats long
        // (p.aligned = 2), 8-byte ones the same with 8 and 2 (p.aligned = 1, two-stage kernel only)
        constexpr int NN = F::EIGHT ? 8 : 4;
        auto al = [&](int vw) {
            bool ia = (sys->nrows % vw == 0) && (sys->batch_stride % vw == 0);
            auto ok = [&](const void *q) { return ((uintptr_t)q & (uintptr_t)(4 * vw - 1)) == 0; };
            for (int n = 0; n < NN; n++) ia = ia && ok(sys->w[n]);
            for (int q = 0; q < F::NUNK; q++) ia = ia && ok(sys->x[q]) && ok(sys->c[q]) && ok(sys->d[q]) && (!F::LATE || ok(sys->x0[q]));
            if (F::NUNK == 2) ia = ia && ok(sys->m);
            return ia;
        };
        p.aligned = al(4) ? 2 : al(2) ? 1 : 0;
    }
    {
        static const int gen = getenv("PDEGPU_ALR_WINDOW") ? atoi(getenv("PDEGPU_ALR_WINDOW")) : 2;
        if (gen >= 2) {
            const int rc2 = window2_pass(ctx, FAM, DIR, &p, g.M, sys->batch);
            if (rc2 != PDEGPU_ERR_UNSUPPORTED) return rc2;
        }
    }
    switch (g.M) {
    case 5:  return launch_window<FAM, DIR, 5>(ctx, p, g, sys->batch);
    case 9:  return launch_window<FAM, DIR, 9>(ctx, p, g, sys->batch);
    case 15: return launch_window<FAM, DIR, 15>(ctx, p, g, sys->batch);
    case 21: return launch_window<FAM, DIR, 21>(ctx, p, g, sys->batch);
    case 25: return launch_window<FAM, DIR, 25>(ctx, p, g, sys->batch);
    default: return PDEGPU_ERR_UNSUPPORTED;
    }
}

}  // namespace

// Transposes up to 16 dense column-major fields of a batch in ONE launch: dst[f][b][i*ncols + j] = src[f][b][j*nrows + i].
struct TransposeMany { const float *src[16]; float *dst[16]; int nf; };

static __global__ void __launch_bounds__(256)
transpose_many_kernel(const TransposeMany t, int nrows, int ncols, long long sstride, long long dstride)
{
    __shared__ float tile[32][33];
    const int f = blockIdx.z % t.nf, b = blockIdx.z / t.nf;
    const float *__restrict__ src = t.src[f] + (long long)b * sstride;
    float *__restrict__ dst = t.dst[f] + (long long)b * dstride;
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
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

// One ALR iteration = column pass on the problem as given (X -> XT, written transposed), then the row pass
// as a column pass of the transposed problem (XT -> X, written transposed back). Coefficients are
// transposed once per call.
template <int FAM>
static int alr_window_run(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega)
{
    using F = Fam<FAM>;
    if (F::PDE && F::EIGHT) return PDEGPU_ERR_UNSUPPORTED;        // interior lines only + untouched border lines: generation 1
    WinGeom g0, g1;
    if (!win_geometry(sys->nrows, F::NUNK, g0) || !win_geometry(sys->ncols, F::NUNK, g1)) return PDEGPU_ERR_UNSUPPORTED;
    if (sys->nrows < 8 || sys->ncols < 8) return PDEGPU_ERR_UNSUPPORTED;
    // Lines that are not a multiple of 4 floats long take 4 scalar loads per vector: measured slower than
    // generation 1 once the problem is large enough to be bandwidth- rather than launch-bound.
    // (even sizes take 8-byte vector loads in the two-stage kernel and stay here)
    if (((sys->nrows | sys->ncols) & 1) && (long long)sys->nrows * sys->ncols * sys->batch > (1ll << 20)) return PDEGPU_ERR_UNSUPPORTED;
    if ((long long)sys->batch * sys->batch_stride >= (1ll << 31)) return PDEGPU_ERR_UNSUPPORTED;
    constexpr int NN = F::EIGHT ? 8 : 4;
    const long long npix = (long long)sys->nrows * sys->ncols;
    const int nfields = NN + 3 * F::NUNK + (F::NUNK == 2 ? 1 : 0) + (F::LATE ? F::NUNK : 0);
    const long long fld = (npix * sys->batch + 3) & ~3ll;      // keep every transposed field 16-B aligned
    int rc = pdegpu_scratch_reserve(ctx, (size_t)nfields * fld * sizeof(float));
    if (rc) return rc;
    pdegpu_system tsys = *sys;
    float *p = (float *)ctx->scratch;
    auto take = [&]() { float *r = p; p += fld; return r; };
    TransposeMany tm;
    tm.nf = 0;
    auto tr = [&](const float *src) -> float * {
        float *d = take();
        tm.src[tm.nf] = src; tm.dst[tm.nf] = d; tm.nf++;
        return d;
    };
    // neighbour roles swap under transposition: N<->W, S<->E, NE<->SW
    static const int perm[8] = {W_N, W_W, W_S, W_E, W_NW, W_SW, W_SE, W_NE};
    tsys.nrows = sys->ncols; tsys.ncols = sys->nrows; tsys.batch_stride = npix;
    for (int n = 0; n < NN; n++) tsys.w[n] = tr(sys->w[perm[n]]);
    float *xT[2] = {nullptr, nullptr};
    for (int q = 0; q < F::NUNK; q++) {
        tsys.c[q] = tr(sys->c[q]);
        tsys.d[q] = tr(sys->d[q]);
        if (F::LATE) tsys.x0[q] = tr(sys->x0[q]);
        xT[q] = take();
        tsys.x[q] = xT[q];
    }
    if (F::NUNK == 2) tsys.m = tr(sys->m);
    {
        dim3 grid((sys->nrows + 31) / 32, (sys->ncols + 31) / 32, sys->batch * tm.nf);
        if (grid.z > 65535) return PDEGPU_ERR_UNSUPPORTED;
        PDEGPU_PROF(ctx, "transpose_many_kernel", 8.0 * npix * sys->batch * tm.nf);
        transpose_many_kernel<<<grid, 256, 0, ctx->stream>>>(tm, sys->nrows, sys->ncols, sys->batch_stride, npix);
        PDEGPU_LAUNCH_CHECK(ctx, "transpose_many_kernel");
    }
    if (rc) return rc;
    float *const xN[2] = {sys->x[0], sys->x[1]};
    for (int it = 0; it < iter; it++) {
        if ((rc = window_pass<FAM, 0>(ctx, sys, xT, npix, omega))) return rc;
        if ((rc = window_pass<FAM, 2>(ctx, &tsys, xN, sys->batch_stride, omega))) return rc;
    }
    return PDEGPU_OK;
}

int relax_window_line(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega)
{
    switch (sys->family) {
    case PDEGPU_FLOW_ELIN4: return alr_window_run<PDEGPU_FLOW_ELIN4>(ctx, sys, iter, omega);
    case PDEGPU_FLOW_LLIN4: return alr_window_run<PDEGPU_FLOW_LLIN4>(ctx, sys, iter, omega);
    case PDEGPU_FLOW_LLIN8: return alr_window_run<PDEGPU_FLOW_LLIN8>(ctx, sys, iter, omega);
    case PDEGPU_DISP_LLIN4: return alr_window_run<PDEGPU_DISP_LLIN4>(ctx, sys, iter, omega);
    case PDEGPU_PDE4:       return alr_window_run<PDEGPU_PDE4>(ctx, sys, iter, omega);
    default: return PDEGPU_ERR_UNSUPPORTED;
    }
}
