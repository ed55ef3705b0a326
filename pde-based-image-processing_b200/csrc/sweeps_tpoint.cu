// sweeps_tpoint.cu -- solver 1 (red-black point SOR) with TEMPORAL BLOCKING: two complete sweeps per pass over HBM
// (north_star subsystem 1). Same ordering, same arithmetic (point_formula) and the same bits as generations 0 / 1.
//
// A persistent CTA owns a strip of rows of one problem and slides a window of RC columns along j (the slow axis; a
// column of the strip is contiguous in memory). Every field of a column -- unknowns, fixed fields, all coefficients --
// enters shared memory ONCE (cp.async, two columns ahead) and stays while the four half-sweeps
//     stage 0 = red of sweep 1, stage 1 = black of sweep 1, stage 2 = red of sweep 2, stage 3 = black of sweep 2
// pass over it, two columns apart: in step f stage s relaxes column f - 2s - 1. A stage needs its predecessor finished
// on the columns c-1, c, c+1 and its successor not yet on c-1: with a distance of two columns both hold for everything
// done in EARLIER steps, so the four stages of a step are independent of each other and a step costs one barrier.
// Column f - 9 is final and goes to X_out (out of place: the neighbouring strip still reads old halo rows from X_in).
//
//   * strips: a strip loads rows [a, b) and may only trust rows a+4 .. b-5 after four half-sweeps (each half-sweep
//     spoils one more row from an edge that is not the image border), so it writes those; 8 of 256 rows are redundant;
//   * the reference's border fill (opticalflowSolvers.c:161-179: border pixel := nearest interior pixel, after every
//     sweep) is done eagerly by the thread that relaxes the interior pixel: a border pixel is read by that one interior
//     pixel only, so copying at update time or after the sweep is the same;
//   * HBM traffic per sweep: (fields read once + unknowns written once) / 2 -- 30 B per pixel for the late-linearisation
//     flow system instead of 60.
#include "stencil_math.cuh"
#include <stdlib.h>
#include <stdint.h>

namespace {

constexpr int TP_STAGES = 4;                 // two sweeps = four half-sweeps
// Two columns in flight, a window of 12 (columns f-9 .. f+2): 70 KB per CTA, three CTAs per SM. (Six columns in flight
// and a window of 16 -- two CTAs per SM -- measured no faster: the step time is not the copies' latency, see below.)
constexpr int TP_AHEAD = 2;
constexpr int TP_RC = 12;
constexpr int TP_RMAX = 112;                 // tallest strip
constexpr int TP_PITCH = 112;                // floats per field and column in shared memory
constexpr int TP_PER_STAGE = 64;
constexpr int TP_THREADS = TP_PER_STAGE * TP_STAGES;
constexpr int TP_SLOTS = 3;                  // 16-byte copies per thread and column (>= NFT * (TP_RMAX / 4) / TP_THREADS)
constexpr int TP_HALO = 4;                   // rows a strip cannot trust per inner edge after four half-sweeps

struct TPParams {
    SysView s;
    float *xo[2];
    int S;                                   // strips per problem
    int ntasks;
    int pitch;                               // floats per field and column in shared memory (>= tallest strip, multiple of 4)
    float omega;
};

__device__ __forceinline__ void tp_cp16(void *sm, const void *g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(sm)), "l"(g) : "memory");
}

// strip s of S: rows it writes [oa, ob) and rows it loads [a, b); all multiples of 4 (nrows is)
__host__ __device__ inline void tp_strip(int nr, int S, int s, int &a, int &b, int &oa, int &ob)
{
    oa = s == 0 ? 0 : (int)(((long long)nr * s / S) & ~3ll);
    ob = s == S - 1 ? nr : (int)(((long long)nr * (s + 1) / S) & ~3ll);
    a = s == 0 ? 0 : oa - TP_HALO;
    b = s == S - 1 ? nr : ob + TP_HALO;
}

template <int FAM>
__global__ void __launch_bounds__(TP_THREADS, 2)
rb_window_kernel(const TPParams p)
{
    using F = Fam<FAM>;
    constexpr int NUNK = F::NUNK;
    constexpr int NFX = NUNK * (F::LATE ? 2 : 1);             // x[q], then x0[q]
    constexpr int NCF = 4 + 2 * NUNK + (NUNK == 2 ? 1 : 0);   // w[0..3], C[q], D[q], M
    constexpr int NFT = NFX + NCF;
    constexpr int CI_W = NFX, CI_C = NFX + 4, CI_D = NFX + 4 + NUNK, CI_M = NFX + 4 + 2 * NUNK;
    extern __shared__ __align__(16) float ring[];             // [TP_RC][NFT][pitch]
    constexpr int P = TP_PITCH, COL = NFT * TP_PITCH;          // compile-time: field offsets fold into the load instructions
    const int nr = p.s.nrows, nc = p.s.ncols;
    const int tid = threadIdx.x;
    const int stage = tid / TP_PER_STAGE, u = tid % TP_PER_STAGE;
    const int colour = stage & 1;
    const float omega = p.omega;

    for (int task = blockIdx.x; task < p.ntasks; task += gridDim.x) {
        const int prob = task / p.S, strip = task - prob * p.S;
        int a, b, oa, ob;
        tp_strip(nr, p.S, strip, a, b, oa, ob);
        const int rin = b - a, n4 = rin >> 2;
        const long long base = (long long)prob * p.s.bstride;
        __shared__ const float *src[NFT];                     // the problem's fields in ring order
        if (tid == 0) {
#pragma unroll
            for (int q = 0; q < NUNK; q++) {
                src[q] = p.s.x[q] + base;
                if (F::LATE) src[NUNK + q] = p.s.x0[q] + base;
                src[CI_C + q] = p.s.c[q] + base; src[CI_D + q] = p.s.d[q] + base;
            }
#pragma unroll
            for (int n = 0; n < 4; n++) src[CI_W + n] = p.s.w[n] + base;
            if (NUNK == 2) src[CI_M] = p.s.m + base;
        }
        __syncthreads();

        // this thread's share of a column's copies: the same (field, 16-byte chunk) pairs for every column
        const float *cp_src[TP_SLOTS];
        int cp_dst[TP_SLOTS];
#pragma unroll
        for (int k = 0; k < TP_SLOTS; k++) {
            const int t = tid + k * TP_THREADS;
            const bool on = t < NFT * n4;
            const int f = on ? t / n4 : 0, v = on ? t - f * n4 : 0;
            cp_src[k] = on ? src[f] + a + 4 * v : nullptr;
            cp_dst[k] = f * P + 4 * v;
        }
        // ring slots are advanced by one per step (no modulo in the loop): of the column being loaded, of this thread's
        // stage column, of the column being written out
        int slot_in = 0;
        auto issue = [&](int c) {                             // all fields of column c, rows [a, b), into its ring slot
            if (c < nc) {
                float *dst = ring + slot_in * COL;
                const long long g = (long long)c * nr;
#pragma unroll
                for (int k = 0; k < TP_SLOTS; k++)
                    if (cp_src[k]) tp_cp16(dst + cp_dst[k], cp_src[k] + g);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            slot_in = slot_in == TP_RC - 1 ? 0 : slot_in + 1;
        };
        for (int c = 0; c < TP_AHEAD; c++) issue(c);

        // this thread's share of the write-out of a finished column
        const int o4 = (ob - oa) >> 2;
        const bool wr_on = tid < NUNK * o4;
        const int wr_q = wr_on ? tid / o4 : 0, wr_v = wr_on ? tid - wr_q * o4 : 0;
        float *wr_dst = p.xo[wr_q] + base + oa + 4 * wr_v;
        const int wr_src = wr_q * P + (oa - a) + 4 * wr_v;

        // rows this thread's stage relaxes in a column: a+1 .. b-2, those of the stage's colour
        const int ilo = a + 1, ihi = b - 2;
        int c = -2 * stage - 1;                               // the stage's column in step f: f - 2 stage - 1
        int slot_m = (c % TP_RC + TP_RC) % TP_RC;             // (column c lives in slot c mod RC)
        int co = -9, slot_o = TP_RC - 9;                      // the column that leaves in step f: f - 9
        for (int f = 0; f <= nc - 1 + 9; f++) {
            asm volatile("cp.async.wait_group %0;" ::"n"(TP_AHEAD - 1) : "memory");
            __syncthreads();
            issue(f + TP_AHEAD);
            if (c >= 1 && c <= nc - 2) {
                const int i = ilo + ((ilo + c + colour) & 1) + 2 * u;
                if (i <= ihi) {
                    const int r = i - a;
                    float *cm = ring + slot_m * COL + r;
                    float *cw = ring + (slot_m == 0 ? TP_RC - 1 : slot_m - 1) * COL + r;
                    float *ce = ring + (slot_m == TP_RC - 1 ? 0 : slot_m + 1) * COL + r;
                    float xn[2][4], xc[2], x0n[2][4], x0c[2], out[2], w[4], C[2] = {0.f, 0.f}, D[2] = {0.f, 0.f};
#pragma unroll
                    for (int n = 0; n < 4; n++) w[n] = cm[(CI_W + n) * P];
#pragma unroll
                    for (int q = 0; q < NUNK; q++) {
                        C[q] = cm[(CI_C + q) * P]; D[q] = cm[(CI_D + q) * P];
                        xc[q] = cm[q * P];
                        xn[q][W_W] = cw[q * P]; xn[q][W_E] = ce[q * P];
                        xn[q][W_N] = cm[q * P - 1]; xn[q][W_S] = cm[q * P + 1];
                        if (F::LATE) {
                            x0c[q] = cm[(NUNK + q) * P];
                            x0n[q][W_W] = cw[(NUNK + q) * P]; x0n[q][W_E] = ce[(NUNK + q) * P];
                            x0n[q][W_N] = cm[(NUNK + q) * P - 1]; x0n[q][W_S] = cm[(NUNK + q) * P + 1];
                        }
                    }
                    point_formula<FAM>(w, xn, xc, x0n, x0c, C, D, NUNK == 2 ? cm[CI_M * P] : 0.f, omega, out);
#pragma unroll
                    for (int q = 0; q < NUNK; q++) cm[q * P] = out[q];
                    // the border pixels whose nearest interior pixel this is (rare)
                    const bool top = i == 1, bot = i == nr - 2, lft = c == 1, rgt = c == nc - 2;
                    if (top || bot || lft || rgt) {
#pragma unroll
                        for (int q = 0; q < NUNK; q++) {
                            const float v = out[q];
                            if (top) cm[q * P - 1] = v;
                            if (bot) cm[q * P + 1] = v;
                            if (lft) {
                                cw[q * P] = v;
                                if (top) cw[q * P - 1] = v;
                                if (bot) cw[q * P + 1] = v;
                            }
                            if (rgt) {
                                ce[q * P] = v;
                                if (top) ce[q * P - 1] = v;
                                if (bot) ce[q * P + 1] = v;
                            }
                        }
                    }
                }
            }
            // column f - 9 has seen all four half-sweeps (and the fill from column f - 8): out
            if (co >= 0 && wr_on)
                *reinterpret_cast<float4 *>(wr_dst + (long long)co * nr) = *reinterpret_cast<const float4 *>(ring + slot_o * COL + wr_src);
            c++; co++;
            slot_m = slot_m == TP_RC - 1 ? 0 : slot_m + 1;
            slot_o = slot_o == TP_RC - 1 ? 0 : slot_o + 1;
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
    }
}

}  // namespace

// `pairs` x two sweeps, ping-pong between cur and alt (both hold NUNK fields of batch * batch_stride floats); on return
// *result_in_alt says where the last pass wrote. PDEGPU_ERR_UNSUPPORTED: not built for this case (the caller sweeps one
// by one).
template <int FAM>
static int tpoint_run(pdegpu_ctx *ctx, const pdegpu_system *sys, float *const cur_in[2], float *const alt_in[2], int pairs, float omega, bool *result_in_alt)
{
    using F = Fam<FAM>;
    constexpr int NUNK = F::NUNK;
    constexpr int NFT = NUNK * (F::LATE ? 2 : 1) + 4 + 2 * NUNK + (NUNK == 2 ? 1 : 0);
    const int nr = sys->nrows, nc = sys->ncols;
    if (nr < 8 || nc < 3 || (nr & 3)) return PDEGPU_ERR_UNSUPPORTED;
    TPParams p;
    memset(&p, 0, sizeof p);
    p.S = nr <= TP_RMAX ? 1 : (nr + TP_RMAX - 13) / (TP_RMAX - 12);
    int tall = 0;
    for (int s = 0; s < p.S; s++) {
        int a, b, oa, ob;
        tp_strip(nr, p.S, s, a, b, oa, ob);
        if (ob <= oa) return PDEGPU_ERR_UNSUPPORTED;
        tall = b - a > tall ? b - a : tall;
    }
    if (tall > TP_RMAX) return PDEGPU_ERR_UNSUPPORTED;
    p.pitch = TP_PITCH;
    p.ntasks = sys->batch * p.S;
    p.omega = omega;
    const size_t smem = (size_t)TP_RC * NFT * TP_PITCH * sizeof(float);
    if (smem > 227 * 1024) return PDEGPU_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(rb_window_kernel<FAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return pdegpu_check_cuda(ctx, e, "cudaFuncSetAttribute(rb_window_kernel)");
    static_assert(13 * (TP_PITCH / 4) <= TP_SLOTS * TP_THREADS, "copy slots");
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    static const int force = getenv("PDEGPU_POINT_WINDOW") ? atoi(getenv("PDEGPU_POINT_WINDOW")) : -1;   // 1: always, 0: never
    if (force == 0 || (force < 0 && p.ntasks < 3 * ctx->sm_count)) return PDEGPU_ERR_UNSUPPORTED;         // (see the measurements above)
    const int grid = p.ntasks < ctx->sm_count * per_sm ? p.ntasks : ctx->sm_count * per_sm;
    float *cur[2] = {cur_in[0], cur_in[1]}, *nxt[2] = {alt_in[0], alt_in[1]};
    bool in_alt = false;
    for (int k = 0; k < pairs; k++) {
        p.s = make_view(sys);
        p.s.x[0] = cur[0]; p.s.x[1] = cur[1];
        p.xo[0] = nxt[0]; p.xo[1] = nxt[1];
        PDEGPU_PROF(ctx, "rb_window_kernel<2 sweeps>", 2.0 * sweep_bytes<FAM>() * (double)nr * nc * sys->batch);
        rb_window_kernel<FAM><<<grid, TP_THREADS, smem, ctx->stream>>>(p);
        PDEGPU_LAUNCH_CHECK(ctx, "rb_window_kernel");
        for (int q = 0; q < 2; q++) { float *t = cur[q]; cur[q] = nxt[q]; nxt[q] = t; }
        in_alt = !in_alt;
    }
    *result_in_alt = in_alt;
    return PDEGPU_OK;
}

int relax_tpoint(pdegpu_ctx *ctx, const pdegpu_system *sys, float *const cur[2], float *const alt[2], int pairs, float omega, bool *result_in_alt)
{
    switch (sys->family) {
    case PDEGPU_FLOW_ELIN4: return tpoint_run<PDEGPU_FLOW_ELIN4>(ctx, sys, cur, alt, pairs, omega, result_in_alt);
    case PDEGPU_FLOW_LLIN4:
    case PDEGPU_FLOW_LLIN8: return tpoint_run<PDEGPU_FLOW_LLIN4>(ctx, sys, cur, alt, pairs, omega, result_in_alt);
    case PDEGPU_DISP_LLIN4: return tpoint_run<PDEGPU_DISP_LLIN4>(ctx, sys, cur, alt, pairs, omega, result_in_alt);
    case PDEGPU_PDE4:       return tpoint_run<PDEGPU_PDE4>(ctx, sys, cur, alt, pairs, omega, result_in_alt);
    default: return PDEGPU_ERR_UNSUPPORTED;
    }
}
