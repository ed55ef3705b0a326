// sweeps_window2.cu -- kernel generation 2b, solver 2: the sliding-window zebra line relaxation of
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
#include "window_common.cuh"
#include <stdlib.h>

namespace {

constexpr int kW2Threads = 384;

template <int FAM, int DIR, int M>
__global__ void __launch_bounds__(kW2Threads, 1)
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
    constexpr int NT = (LS + 127) / 128;
    constexpr int BUF = RF::N * LS;                           // floats per row buffer
    extern __shared__ float smem[];
    const int R = p.R, D = p.D, NBR = R >> 3, NA = p.NA, NS = p.NS, NBUF = p.NBUF;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *ring = smem;
    float *bufs = ring + (size_t)R * SP;
    unsigned *flags = reinterpret_cast<unsigned *>(bufs + (size_t)NBUF * BUF);
    unsigned *solved_seq = flags, *written_seq = flags + R, *block_cnt = flags + R + NBR;
    unsigned *filled_seq = flags + R + 2 * NBR, *freed_seq = filled_seq + NBUF;
    for (int t = threadIdx.x; t < R + 2 * NBR + 2 * NBUF; t += blockDim.x) flags[t] = 0;
    __syncthreads();

    const int n = p.n, nlines = p.nlines;
    const SysView &s = p.s;
    const int B0 = (int)((long long)blockIdx.x * p.TB / gridDim.x), B1 = (int)((long long)(blockIdx.x + 1) * p.TB / gridDim.x);
    const int nblk = B1 - B0;
    const bool redundant = B1 < p.TB && (B1 % p.NB) != 0;
    const int Ltot = 8 * nblk + (redundant ? 1 : 0);
    const int Q = D + 2 * ((Ltot + 1) >> 1);
    const float omega = p.omega, om1 = 1.0f - p.omega;
    const int ec_last = (n - 1) & ~3;
    const bool al = p.aligned != 0;

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
        for (int q = warp; q < Q; q += NA) {
            const int bi = q % NBUF;
            const unsigned use = (unsigned)(q / NBUF);
            WinTask T;
            const bool valid = decode(q, T);
            warp_wait_ge(&freed_seq[bi], use, lane);          // the buffer's previous rows have been picked up
            if (valid) {
                const int l = T.l;
                if (l >= R) {                                 // ring slot free (see sweeps_window.cu)
                    const int lbp = (l - R) >> 3;
                    warp_wait_ge(&written_seq[lbp % NBR], (unsigned)lbp + 1, lane);
                    if (lbp > 0) warp_wait_ge(&written_seq[(lbp - 1) % NBR], (unsigned)lbp, lane);
                }
                float *buf = bufs + (size_t)bi * BUF;
                float *rs = ring + (size_t)(l % R) * SP;
                const float *rsW = ring + (size_t)((l + R - 1) % R) * SP, *rsE = ring + (size_t)((l + 1) % R) * SP;
#pragma unroll 1
                for (int t = 0; t < NT; t++) {
                    const int e0 = 128 * t + 4 * lane;
                    const int ec = min(e0, ec_last);
                    RawBatch<FAM> rb;
                    if (e0 < LS) rb.issue(s, T, ec, n, al);
                    if (t == 0 && T.odd) {
                        warp_wait_ge(&solved_seq[(l - 1) % R], (unsigned)l, lane);
                        if (T.eE) warp_wait_ge(&solved_seq[(l + 1) % R], (unsigned)l + 2, lane);
                    }
                    if (e0 < LS) {
                        if (T.odd) rb.neighbours_from_ring(rsW, rsE, P, ec, n);
                        float ra[4], rc[4], rb1[4], rd1[4], rb2[4], rd2[4], rm[4], xo0[4], xo1[4];
#pragma unroll
                        for (int k = 0; k < 4; k++) {
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
                        st4(buf + RF::A * LS + e0, ra);
                        st4(buf + RF::C * LS + e0, rc);
                        st4(buf + RF::B1 * LS + e0, rb1);
                        st4(buf + RF::D1 * LS + e0, rd1);
                        st4(rs + e0, xo0);
                        if (NUNK == 2) {
                            st4(buf + RF::B2 * LS + e0, rb2);
                            st4(buf + RF::D2 * LS + e0, rd2);
                            st4(buf + RF::MM * LS + e0, rm);
                            st4(rs + P + e0, xo1);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) st_release(&filled_seq[bi], use + 1);
            __syncwarp();
        }
    } else {
        // ================================= solver warps =================================
        for (int q = warp - NA; q < Q; q += NS) {
            const int bi = q % NBUF;
            const unsigned use = (unsigned)(q / NBUF);
            WinTask T;
            const bool valid = decode(q, T);
            warp_wait_ge(&filled_seq[bi], use + 1, lane);
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
                    if (cnt == 8 && p.vec_ok) {
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
        }
    }
}

template <int FAM, int DIR, int M>
int launch_window2(pdegpu_ctx *ctx, const WinParams &p, size_t smem, int batch)
{
    static bool attr_set[16] = {false};
    if (!attr_set[ctx->device & 15]) {
        cudaError_t e = cudaFuncSetAttribute(alr_window2_kernel<FAM, DIR, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return pdegpu_check_cuda(ctx, e, "cudaFuncSetAttribute(alr_window2_kernel)");
        attr_set[ctx->device & 15] = true;
    }
    const int grid = p.TB < ctx->sm_count ? p.TB : ctx->sm_count;
    PDEGPU_PROF(ctx, DIR == 0 ? "alr_window2_kernel<dir0>" : "alr_window2_kernel<dir1,transposed>",
                sweep_bytes<FAM>() * (double)p.n * p.nlines * batch);
    alr_window2_kernel<FAM, DIR, M><<<grid, (p.NA + p.NS) * 32, smem, ctx->stream>>>(p);
    PDEGPU_LAUNCH_CHECK(ctx, "alr_window2_kernel");
    return PDEGPU_OK;
}

template <int FAM, int DIR>
int window2_dispatch(pdegpu_ctx *ctx, WinParams &p, int M, int nunk, int batch)
{
    // geometry: 8 assemblers + 4 solvers; the ring as large as leaves room for >= 4 row buffers
    const int LS = 32 * M, SP = nunk * LS + 4, rowf = nunk == 2 ? 7 : 4;
    const size_t room = 227 * 1024;
    p.NA = 8; p.NS = 4;
    static const int RD[][2] = {{32, 5}, {24, 4}, {16, 3}};
    for (auto &rd : RD) {
        const size_t fixed = ((size_t)rd[0] * SP + rd[0] + 2 * (rd[0] / 8) + 32) * sizeof(float);
        if (fixed >= room) continue;
        int nbuf = (int)((room - fixed) / ((size_t)rowf * LS * sizeof(float)));
        if (nbuf > 8) nbuf = 8;
        if (nbuf < 4) continue;
        p.R = rd[0]; p.D = rd[1]; p.NBUF = nbuf;
        const size_t smem = fixed + (size_t)nbuf * rowf * LS * sizeof(float);
        switch (M) {
        case 5:  return launch_window2<FAM, DIR, 5>(ctx, p, smem, batch);
        case 9:  return launch_window2<FAM, DIR, 9>(ctx, p, smem, batch);
        case 15: return launch_window2<FAM, DIR, 15>(ctx, p, smem, batch);
        case 21: return launch_window2<FAM, DIR, 21>(ctx, p, smem, batch);
        case 25: return launch_window2<FAM, DIR, 25>(ctx, p, smem, batch);
        default: return PDEGPU_ERR_UNSUPPORTED;
        }
    }
    return PDEGPU_ERR_UNSUPPORTED;
}

}  // namespace

// entry used by sweeps_window.cu: p is complete except for the geometry
int window2_pass(pdegpu_ctx *ctx, int family, int dir, void *params, int M, int batch)
{
    WinParams &p = *static_cast<WinParams *>(params);
#define W2(FAMID) (dir == 0 ? window2_dispatch<FAMID, 0>(ctx, p, M, Fam<FAMID>::NUNK, batch) : window2_dispatch<FAMID, 2>(ctx, p, M, Fam<FAMID>::NUNK, batch))
    switch (family) {
    case PDEGPU_FLOW_ELIN4: return W2(PDEGPU_FLOW_ELIN4);
    case PDEGPU_FLOW_LLIN4: return W2(PDEGPU_FLOW_LLIN4);
    case PDEGPU_FLOW_LLIN8: return W2(PDEGPU_FLOW_LLIN8);
    case PDEGPU_DISP_LLIN4: return W2(PDEGPU_DISP_LLIN4);
    case PDEGPU_PDE4:       return W2(PDEGPU_PDE4);
    default: return PDEGPU_ERR_UNSUPPORTED;
    }
#undef W2
}
