// sweeps_window2_impl.cuh -- kernel generation 2b, solver 2: the sliding-window zebra line relaxation of
// sweeps_window.cu with the two halves of a line's work on DIFFERENT warps.
//
// In the one-warp-per-line kernel a warp spends ~2/3 of a line's time waiting for global loads while it
// assembles the tridiagonal rows and ~1/3 in the latency chain of the solve, and its 13 KB row scratch is
// occupied the whole time: 7 lines in flight per SM, 12 % occupancy, 25 % issue utilisation (profiles/).
// Here
//   * ASSEMBLER warps (NA = 8) do the coalesced loads and the row formulas of one line each and leave the
//     rows in one of NBUF shared row buffers;
//   * SOLVER warps (NS = 4) pick the buffers up in order, pull the rows into registers (lane = chunk), give
//     the buffer back, solve, relax and publish the line in the ring, exactly as before.
// A buffer is held for the assembly only, so the same shared memory keeps 12 warps busy instead of 7 and the
// load latency of eight lines overlaps the solves of four. Schedule, ring, redundancy rule, arithmetic and
// output layout are those of sweeps_window.cu; every schedule entry q (valid or not) passes through
// assembler q % NA, buffer q % NBUF and solver q % NS, so both sides agree on a buffer's use count without
// communicating. All waits are on smaller q or on the earlier stage of the same q: no deadlock.
#pragma once
#include "window_common.cuh"
#include <stdlib.h>

namespace {

constexpr int kW2Threads = 384;
constexpr int kW2ThreadsV2 = 512;                             // the 8-byte-vector variant needs fewer registers: up to 16 warps
constexpr int kW2MaxG = 4;                                   // at most 4 assembler warps per line

// Phase probes (build with -DW2_PROBE; never in the shipped library): cycles per warp role, summed over all warps.
#ifdef W2_PROBE
__device__ unsigned long long g_w2_probe[16];
#define PROBE_DECL unsigned long long pr_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pt_ = clock64(); const long long pt0_ = pt_
#define PROBE(k) do { const long long n_ = clock64(); pr_[k] += (unsigned long long)(n_ - pt_); pt_ = n_; } while (0)
#define PROBE_FLUSH(base) do { if (lane == 0) { pr_[7] = (unsigned long long)(clock64() - pt0_); for (int k_ = 0; k_ < 8; k_++) atomicAdd(&g_w2_probe[(base) + k_], pr_[k_]); } } while (0)
#define PROBE_USE(v) do { if (__float_as_uint(v) == 0x7fc12345u) __trap(); } while (0)
#else
#define PROBE_DECL
#define PROBE(k)
#define PROBE_FLUSH(base)
#define PROBE_USE(v)
#endif

// VW: pixels per lane and batch = width of the vector loads (4: 16-byte, needs lines of a multiple of 4 floats; 2: 8-byte,
// for the even-sized pyramid levels); AL = false: VW scalar loads per vector (odd line lengths).
template <int FAM, int DIR, int M, int VW, bool AL>
__global__ void __launch_bounds__((VW == 2 && AL) ? kW2ThreadsV2 : kW2Threads, 1)
alr_window2_kernel(const WinParams p)
{
    using F = Fam<FAM>;
    using RF = RowF<F::NUNK>;
    static_assert(M & 1, "chunk length must be odd");
    constexpr int NUNK = F::NUNK;
    constexpr int qa = (NUNK == 2 && DIR != 0) ? 1 : 0, qb = 1 - qa;
    constexpr int LS = 32 * M;
    constexpr int P = LS;
    constexpr int SP = NUNK * P + 4;
    // pixels per lane and batch, batches in flight per lane. Measured (B200, 64 x 480x640, us per pass, lines of 480 / 640):
    // VW 4 single-buffered 395 / 448, VW 2 double-buffered 411 / 542 (same bytes in flight per lane -- the register file
    // is the limit -- and twice the load instructions).
    constexpr int NT = (LS + 32 * VW - 1) / (32 * VW);
    constexpr int BUF = RF::N * LS;                           // floats per row buffer
    extern __shared__ float smem[];
    const int R = p.R, D = p.D, NBR = R >> 3, NA = p.NA, NS = p.NS, NBUF = p.NBUF;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *ring = smem;
    float *bufs = ring + (size_t)R * SP;
    unsigned *flags = reinterpret_cast<unsigned *>(bufs + (size_t)NBUF * BUF);
    unsigned *solved_seq = flags, *written_seq = flags + R, *block_cnt = flags + R + NBR;
    unsigned *filled_seq = flags + R + 2 * NBR, *freed_seq = filled_seq + kW2MaxG * NBUF;
    for (int t = threadIdx.x; t < R + 2 * NBR + (kW2MaxG + 1) * NBUF; t += blockDim.x) flags[t] = 0;
    __syncthreads();

    const int n = p.n, nlines = p.nlines;
    const SysView &s = p.s;
    const int B0 = (int)((long long)blockIdx.x * p.TB / gridDim.x), B1 = (int)((long long)(blockIdx.x + 1) * p.TB / gridDim.x);
    const int nblk = B1 - B0;
    const bool redundant = B1 < p.TB && (B1 % p.NB) != 0;
    const int Ltot = 8 * nblk + (redundant ? 1 : 0);
    const int Q = D + 2 * ((Ltot + 1) >> 1);
    const float omega = p.omega, om1 = 1.0f - p.omega;
    const int ec_last = (n - 1) & ~(VW - 1);

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

    if (warp < NA) {
        // =============================== assembler warps ===============================
        // G warps share one line: warp h of a group takes the batches t = h, h+G, ... (a lane owns VW consecutive pixels
        // of a batch, batch t = elements 32*VW*t ..). A row buffer is then held for 1/G of the time during assembly, and
        // only NA/G lines are in assembly at once, so the other buffers decouple the assemblers from the solvers.
        // The loads of a warp's next batch (or of its first batch of the group's next line) are issued as soon as the
        // registers of the current one are free.
        PROBE_DECL;
        RawBatch<FAM, VW> rawA;
        WinTask T, Tn;
        const int G = p.G, NG = NA / G, h = warp % G;
        int q = warp / G;
        bool valid = q < Q && decode(q, T);
        auto first_el = [&](int t, int &e0, int &ec) { e0 = 32 * VW * t + VW * lane; ec = min(e0, ec_last); };
        auto issue_to = [&](RawBatch<FAM, VW> &rb, const WinTask &TT, int t) {
            int e0, ec;
            first_el(t, e0, ec);
            if (t < NT && e0 < LS) rb.template issue<AL>(s, TT, ec, n);
        };
        if (valid) issue_to(rawA, T, h);
        for (; q < Q; q += NG) {
            const int bi = q % NBUF;
            const unsigned use = (unsigned)(q / NBUF);
            const bool validn = q + NG < Q && decode(q + NG, Tn);
            PROBE(0);
            warp_wait_ge(&freed_seq[bi], use, lane);          // the buffer's previous rows have been picked up
            PROBE(1);
            if (valid) {
                const int l = T.l;
                if (l >= R) {                                 // ring slot free (see sweeps_window.cu)
                    const int lbp = (l - R) >> 3;
                    warp_wait_ge(&written_seq[lbp % NBR], (unsigned)lbp + 1, lane);
                    if (lbp > 0) warp_wait_ge(&written_seq[(lbp - 1) % NBR], (unsigned)lbp, lane);
                }
                PROBE(2);
                float *buf = bufs + (size_t)bi * BUF;
                float *rs = ring + (size_t)(l % R) * SP;
                const float *rsW = ring + (size_t)((l + R - 1) % R) * SP, *rsE = ring + (size_t)((l + 1) % R) * SP;
                auto rows_of = [&](RawBatch<FAM, VW> &rb, int t) {
                    int e0, ec;
                    first_el(t, e0, ec);
                    if (e0 >= LS) return;
                    if (T.odd) rb.neighbours_from_ring(rsW, rsE, P, ec, n);
                    float ra[VW], rc[VW], rb1[VW], rd1[VW], rb2[VW], rd2[VW], rm[VW], xo0[VW], xo1[VW];
#pragma unroll
                    for (int k = 0; k < VW; k++) {
                        PixelRaw<FAM, DIR> r;
                        const bool ok = e0 + k < n;
                        rb.template pixel<DIR>(k, ec + k, n, T.eW, T.eE, r);
                        float a, c, b[2], d[2], m;
                        r.rows(a, c, b, d, m);
                        ra[k] = ok ? a : 0.f; rc[k] = ok ? c : 0.f;
                        rb1[k] = ok ? b[qa] : 1.0f; rd1[k] = ok ? d[qa] : 0.f;
                        rb2[k] = ok ? b[qb] : 1.0f; rd2[k] = ok ? d[qb] : 0.f;
                        rm[k] = ok ? m : 0.f;
                        xo0[k] = ok ? r.xo[0] : 0.f; xo1[k] = (ok && NUNK == 2) ? r.xo[NUNK - 1] : 0.f;
                    }
                    stv(buf + RF::A * LS + e0, ra);
                    stv(buf + RF::C * LS + e0, rc);
                    stv(buf + RF::B1 * LS + e0, rb1);
                    stv(buf + RF::D1 * LS + e0, rd1);
                    stv(rs + e0, xo0);
                    if (NUNK == 2) {
                        stv(buf + RF::B2 * LS + e0, rb2);
                        stv(buf + RF::D2 * LS + e0, rd2);
                        stv(buf + RF::MM * LS + e0, rm);
                        stv(rs + P + e0, xo1);
                    }
                };
#pragma unroll 1
                for (int t = h; t < NT; t += G) {
                    PROBE(0);
                    if (t == h && T.odd) {                    // new values of the even neighbours
                        warp_wait_ge(&solved_seq[(l - 1) % R], (unsigned)l, lane);
                        if (T.eE) warp_wait_ge(&solved_seq[(l + 1) % R], (unsigned)l + 2, lane);
                    }
                    PROBE(4);
                    PROBE_USE(rawA.w4[0].v[0]); PROBE_USE(rawA.XO4[0].v[0]);
                    PROBE(5);
                    rows_of(rawA, t);
                    PROBE(6);
                    // the batch is consumed: its registers take the next one (of this line, or of the group's next line)
                    if (t + G < NT) issue_to(rawA, T, t + G);
                    else if (validn) issue_to(rawA, Tn, h);
                    PROBE(3);
                }
                if (h >= NT && validn) issue_to(rawA, Tn, h);
            } else if (validn) issue_to(rawA, Tn, h);
            __syncwarp();
            if (lane == 0) st_release(&filled_seq[bi * kW2MaxG + h], use + 1);
            __syncwarp();
            T = Tn; valid = validn;
        }
        PROBE_FLUSH(0);
    } else {
        // ================================= solver warps =================================
        PROBE_DECL;
        for (int q = warp - NA; q < Q; q += NS) {
            const int bi = q % NBUF;
            const unsigned use = (unsigned)(q / NBUF);
            WinTask T;
            const bool valid = decode(q, T);
            PROBE(0);
            for (int g = 0; g < p.G; g++) warp_wait_ge(&filled_seq[bi * kW2MaxG + g], use + 1, lane);
            PROBE(1);
            if (!valid) {
                if (lane == 0) st_release(&freed_seq[bi], use + 1);
                __syncwarp();
                continue;
            }
            const int l = T.l, lb = l >> 3;
            const float *buf = bufs + (size_t)bi * BUF;
            float *rs = ring + (size_t)(l % R) * SP;
            {
                const int o = lane * M;
                float a[M], c[M], b[M], d[M];
#pragma unroll
                for (int k = 0; k < M; k++) {
                    a[k] = buf[RF::A * LS + o + k]; c[k] = buf[RF::C * LS + o + k];
                    b[k] = buf[RF::B1 * LS + o + k]; d[k] = buf[RF::D1 * LS + o + k];
                }
                if (NUNK == 1) { __syncwarp(); if (lane == 0) st_release(&freed_seq[bi], use + 1); __syncwarp(); }
                chunk_solve<M>(a, c, b, d, lane);
#pragma unroll
                for (int k = 0; k < M; k++) {
                    d[k] = omega * d[k] + om1 * rs[qa * P + o + k];
                    rs[qa * P + o + k] = d[k];
                }
                if (NUNK == 2) {
#pragma unroll
                    for (int k = 0; k < M; k++) {
                        a[k] = buf[RF::A * LS + o + k];
                        b[k] = buf[RF::B2 * LS + o + k];
                        d[k] = buf[RF::D2 * LS + o + k] - buf[RF::MM * LS + o + k] * d[k];
                    }
                    __syncwarp();
                    if (lane == 0) st_release(&freed_seq[bi], use + 1);      // rows are in registers: the buffer can be refilled
                    __syncwarp();
                    chunk_solve<M>(a, c, b, d, lane);
#pragma unroll
                    for (int k = 0; k < M; k++) rs[qb * P + o + k] = omega * d[k] + om1 * rs[qb * P + o + k];
                }
            }
            __syncwarp();
            PROBE(2);
            unsigned done = 0;
            if (lane == 0) {
                st_release(&solved_seq[l % R], (unsigned)l + 1);
                __threadfence_block();
                if (T.owned) done = atomicAdd(&block_cnt[lb % NBR], 1u) + 1;
            }
            done = __shfl_sync(0xffffffffu, done, 0);
            const int j0 = 8 * T.jb, cnt = min(8, nlines - j0);
            if (T.owned && (int)done == cnt) {
                // this warp completed block lb: write its lines to X_out (transposed layout)
                __threadfence_block();
                const float *rblk = ring + (size_t)((8 * lb) % R) * SP;
#pragma unroll
                for (int qq = 0; qq < NUNK; qq++) {
                    float *o = p.xout[qq] + (long long)T.img * p.ostride + j0;
                    const float *rq = rblk + qq * P;
                    if (cnt == 8 && p.vec_ok == 2) {
                        const int h = lane >> 4;
                        const float *rh = rq + (size_t)(4 * h) * SP;
#pragma unroll 2
                        for (int i = lane & 15; i < n; i += 16) {
                            float4 v;
                            v.x = rh[i]; v.y = rh[SP + i]; v.z = rh[2 * SP + i]; v.w = rh[3 * SP + i];
                            *reinterpret_cast<float4 *>(o + (long long)i * nlines + 4 * h) = v;
                        }
                    } else if (cnt == 8 && p.vec_ok == 1) {
                        // X_out only 8-byte aligned (nlines even, not a multiple of 4): two lines per lane, still one
                        // full 32-B sector per element and instruction
                        const int h = lane >> 3;
                        const float *rh = rq + (size_t)(2 * h) * SP;
#pragma unroll 2
                        for (int i = lane & 7; i < n; i += 8)
                            *reinterpret_cast<float2 *>(o + (long long)i * nlines + 2 * h) = make_float2(rh[i], rh[SP + i]);
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
                PROBE(3);
            }
        }
        PROBE_FLUSH(8);
    }
}

template <int FAM, int DIR, int M, int VW, bool AL>
int launch_window2(pdegpu_ctx *ctx, const WinParams &p, size_t smem, int batch)
{
    {   // every launch: the attribute is per device and the call is cheap (no static per-ordinal bookkeeping)
    cudaError_t e = cudaFuncSetAttribute(alr_window2_kernel<FAM, DIR, M, VW, AL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return pdegpu_check_cuda(ctx, e, "cudaFuncSetAttribute(alr_window2_kernel)");
    }
    const int grid = p.TB < ctx->sm_count ? p.TB : ctx->sm_count;
    PDEGPU_PROF(ctx, DIR == 0 ? "alr_window2_kernel<dir0>" : "alr_window2_kernel<dir1,transposed>",
                sweep_bytes<FAM>() * (double)p.n * p.nlines * batch);
    alr_window2_kernel<FAM, DIR, M, VW, AL><<<grid, (p.NA + p.NS) * 32, smem, ctx->stream>>>(p);
    PDEGPU_LAUNCH_CHECK(ctx, "alr_window2_kernel");
    return PDEGPU_OK;
}

template <int FAM, int DIR>
int window2_dispatch(pdegpu_ctx *ctx, WinParams &p, int M, int nunk, int batch)
{
    // geometry: 8 assemblers + 4 solvers; the ring as large as leaves room for >= 4 row buffers
    const int LS = 32 * M, SP = nunk * LS + 4, rowf = nunk == 2 ? 7 : 4;
    const size_t room = 227 * 1024;
    static const int envNA = getenv("PDEGPU_W2_NA") ? atoi(getenv("PDEGPU_W2_NA")) : 8;       // tuning overrides (NA + NS <= 12)
    static const int envNS = getenv("PDEGPU_W2_NS") ? atoi(getenv("PDEGPU_W2_NS")) : 4;
    static const int envNA2 = getenv("PDEGPU_W2_NA2") ? atoi(getenv("PDEGPU_W2_NA2")) : 12;   // assembler warps of the 8-byte-vector variant (fewer registers: 16 warps fit)
    static const int envVW = getenv("PDEGPU_W2_VW") ? atoi(getenv("PDEGPU_W2_VW")) : 0;        // 2: 8-byte vectors even where 16-byte ones are possible
    if (envVW == 2 && p.aligned == 2) p.aligned = 1;
    p.NA = p.aligned == 1 ? envNA2 : envNA; p.NS = envNS;
    static const int envG = getenv("PDEGPU_W2_G") ? atoi(getenv("PDEGPU_W2_G")) : 2;          // assembler warps per line (measured, us per pass 480 / 640: G=1 413 / 432, G=2 411 / 406, G=4 442 / 439)
    p.G = envG;
    if (p.G < 1 || p.G > kW2MaxG || p.NA % p.G) return PDEGPU_ERR_UNSUPPORTED;
    if (p.NA < 1 || p.NS < 1 || p.NA + p.NS > (p.aligned == 1 ? kW2ThreadsV2 : kW2Threads) / 32) return PDEGPU_ERR_UNSUPPORTED;
    static const int envR = getenv("PDEGPU_W2_R") ? atoi(getenv("PDEGPU_W2_R")) : 0;          // ring lines (multiple of 8) / even lead
    static const int envD = getenv("PDEGPU_W2_D") ? atoi(getenv("PDEGPU_W2_D")) : 0;
    static const int envNBUF = getenv("PDEGPU_W2_NBUF") ? atoi(getenv("PDEGPU_W2_NBUF")) : 8;
    const int RD[][2] = {{envR ? envR : 32, envD ? envD : 5}, {24, 4}, {16, 3}};
    for (auto &rd : RD) {
        const size_t fixed = ((size_t)rd[0] * SP + rd[0] + 2 * (rd[0] / 8) + (kW2MaxG + 1) * 16) * sizeof(float);
        if (fixed >= room) continue;
        int nbuf = (int)((room - fixed) / ((size_t)rowf * LS * sizeof(float)));
        if (nbuf > envNBUF) nbuf = envNBUF;
        if (nbuf > 16) nbuf = 16;
        if (nbuf < 4) continue;
        p.R = rd[0]; p.D = rd[1]; p.NBUF = nbuf;
        const size_t smem = fixed + (size_t)nbuf * rowf * LS * sizeof(float);
        switch (M) {
        case 5: return p.aligned == 2 ? launch_window2<FAM, DIR, 5, 4, true>(ctx, p, smem, batch) : p.aligned == 1 ? launch_window2<FAM, DIR, 5, 2, true>(ctx, p, smem, batch) : launch_window2<FAM, DIR, 5, 4, false>(ctx, p, smem, batch);
        case 9: return p.aligned == 2 ? launch_window2<FAM, DIR, 9, 4, true>(ctx, p, smem, batch) : p.aligned == 1 ? launch_window2<FAM, DIR, 9, 2, true>(ctx, p, smem, batch) : launch_window2<FAM, DIR, 9, 4, false>(ctx, p, smem, batch);
        case 15: return p.aligned == 2 ? launch_window2<FAM, DIR, 15, 4, true>(ctx, p, smem, batch) : p.aligned == 1 ? launch_window2<FAM, DIR, 15, 2, true>(ctx, p, smem, batch) : launch_window2<FAM, DIR, 15, 4, false>(ctx, p, smem, batch);
        case 21: return p.aligned == 2 ? launch_window2<FAM, DIR, 21, 4, true>(ctx, p, smem, batch) : p.aligned == 1 ? launch_window2<FAM, DIR, 21, 2, true>(ctx, p, smem, batch) : launch_window2<FAM, DIR, 21, 4, false>(ctx, p, smem, batch);
        case 25: return p.aligned == 2 ? launch_window2<FAM, DIR, 25, 4, true>(ctx, p, smem, batch) : p.aligned == 1 ? launch_window2<FAM, DIR, 25, 2, true>(ctx, p, smem, batch) : launch_window2<FAM, DIR, 25, 4, false>(ctx, p, smem, batch);
        default: return PDEGPU_ERR_UNSUPPORTED;
        }
    }
    return PDEGPU_ERR_UNSUPPORTED;
}

}  // namespace

// One translation unit per group of families (sweeps_window2*.cu): the 150 instantiations of the kernel compile in
// parallel instead of one after the other.
#define PDEGPU_W2_FAMILY(FAMID)                                                                                        \
    int window2_family_##FAMID(pdegpu_ctx *ctx, int dir, void *params, int M, int batch)                               \
    {                                                                                                                  \
        WinParams &p = *static_cast<WinParams *>(params);                                                              \
        return dir == 0 ? window2_dispatch<FAMID, 0>(ctx, p, M, Fam<FAMID>::NUNK, batch)                               \
                        : window2_dispatch<FAMID, 2>(ctx, p, M, Fam<FAMID>::NUNK, batch);                              \
    }
