// sweeps_tline.cu -- kernel generation 3 of solver 2 (zebra line relaxation): preparation, sequencing of the
// passes and recovery of the increment. Formulation and layouts: tline_common.cuh; pass kernel:
// sweeps_tline_impl.cuh. Replaces, where it applies, generation 2 (sweeps_window*.cu) for
// GS_ALR_SOR_elin4_2d / _llin4_2d / _llin8_2d (opticalflowSolvers.c:196,690,1677), the disparity
// GS_ALR_SOR_llin4_2d (disparitySolvers.c:154) and GS_ALR_SOR_4_2d (pdeSolvers.c:277).
#include "sweeps_tline_impl.cuh"
#include "sweeps_lex_impl.cuh"
#include "driver_formulas.cuh"
#include <cstdlib>

namespace {

// ------------------------------------------------------------------------------------------------------------------
// Preparation: the caller's dense column-major arrays -> packed lines of both passes (tline_common.cuh) + packed T lines.
// One CTA per 32 x 32 tile and problem: every input of the tile is loaded up front (13 independent loads per pixel
// in flight), the derived fields (T, C') are formed in registers, the column-pass lines are stored directly, the
// row-pass lines through one shared tile per field (a single barrier). Pad elements of the pitched lines are written
// as zeros: bulk copies never move uninitialised data.
// ------------------------------------------------------------------------------------------------------------------
struct PrepParams {
    const float *w[8], *m, *c[2], *d[2], *x0[2], *x[2];
    int nrows, ncols, pitch0, pitch1;
    int S0, ns0, S1, ns1;                    // segments per line and their length: column pass (cuts along i), row pass (along j)
    long long ibs;                           // floats between problems of the inputs
    float *pn, *pt, *tn;                     // packed coefficient lines: column pass / row pass; packed T lines (column pass)
    // FUSED preparation (late-linearisation flow driver, north_star subsystem 3): the diffusion weights (OPdiffWeights of
    // x0 + x) and the robust data terms M, C, D are computed here, per pixel, instead of being read: w, m, c, d unused
    LlinTermsArgs terms;
};

// C' = C + D x0 + M y0. Where the reference's row would produce NaN (a number for C next to a NaN D or M) C' stays a
// number and the NaN reaches the row through D / M, as it does in the reference.
__device__ __forceinline__ float tl_cprime(float C, float D, float M, float x0, float y0, bool late, bool flow)
{
    if (!late || is_nan(C) || is_nan(D) || (flow && is_nan(M))) return C;
    float v = C + D * x0;
    if (flow) v += M * y0;
    return v;
}

template <int FAM, bool FUSED = false>
__global__ void __launch_bounds__(256)
tline_prep_kernel(const PrepParams p)
{
    using F = Fam<FAM>;
    constexpr int NUNK = F::NUNK, NN = F::EIGHT ? 8 : 4;
    constexpr int NC = 6 + (NUNK == 2 ? 3 : 0) + (NN == 8 ? 4 : 0);
    constexpr int rDG = NUNK == 2 ? 9 : 6;
    extern __shared__ float tiles[];                         // [NC][32][33], indexed by the row of the ROW-pass line
    const int b = blockIdx.z;
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i = i0 + tx, ic = min(i, p.nrows - 1);
    const int seg0 = min(i / p.ns0, p.S0 - 1), io = i - seg0 * p.ns0;          // segment of the column-pass line, element in it
    // rows of a packed line. Column pass (lines = Matlab columns): previous/next element = N/S, lines j-1 / j+1 = W/E,
    // diagonals (LP, LN, HP, HN) = NW SW NE SE, first unknown first. Row pass (lines = Matlab rows): previous/next = W/E,
    // lines i-1 / i+1 = N/S, diagonals NW NE SW SE, SECOND unknown first (the reference's pass order,
    // opticalflowSolvers.c:735-754).
    constexpr int wsrc[8] = {W_N, W_S, W_W, W_E, W_NW, W_SW, W_NE, W_SE};
    constexpr int wrow0[8] = {TL_WP, TL_WN, TL_WL, TL_WH, rDG + 0, rDG + 1, rDG + 2, rDG + 3};
    constexpr int wrow1[8] = {TL_WL, TL_WH, TL_WP, TL_WN, rDG + 0, rDG + 2, rDG + 1, rDG + 3};
    float vw[4][NN], vm[4], vc[4][NUNK], vd[4][NUNK], vx0[4][NUNK], vx[4][NUNK];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int jc = min(j0 + ty + 8 * r, p.ncols - 1);
        const long long src = (long long)b * p.ibs + (long long)jc * p.nrows + ic;
#pragma unroll
        for (int q = 0; q < NUNK; q++) {
            vx[r][q] = p.x[q][src];
            vx0[r][q] = F::LATE ? p.x0[q][src] : 0.f;
        }
        if (FUSED) {
            // FlowEminND_llin_2D_v10.m:321 (OPdiffWeights(U+dU, V+dV)) and :289-327 (robust weights, channel sums)
            const long long pb = (long long)b * p.ibs;
            const OpdiffSrc os = {p.x0[0] + pb, p.x0[NUNK - 1] + pb, p.x[0] + pb, p.x[NUNK - 1] + pb};
            float wW, wN, wS, wE;
            opdiff_at(os, ic, jc, p.nrows, p.ncols, wW, wN, wS, wE);
            vw[r][0] = wN; vw[r][1] = wS; vw[r][2] = wW; vw[r][3] = wE;       // wsrc order: N S W E
            float M, Cu, Cv, Du, Dv;
            llin_terms_at(p.terms, b, (long long)jc * p.nrows + ic, vx[r][0], vx[r][NUNK - 1], M, Cu, Cv, Du, Dv);
            vm[r] = M; vc[r][0] = Cu; vc[r][NUNK - 1] = Cv; vd[r][0] = Du; vd[r][NUNK - 1] = Dv;
        } else {
#pragma unroll
            for (int k = 0; k < NN; k++) vw[r][k] = p.w[wsrc[k] % NN == wsrc[k] ? wsrc[k] : 0][src];
            vm[r] = NUNK == 2 ? p.m[src] : 0.f;
#pragma unroll
            for (int q = 0; q < NUNK; q++) { vc[r][q] = p.c[q][src]; vd[r][q] = p.d[q][src]; }
        }
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int j = j0 + ty + 8 * r;
        const bool st = io < p.pitch0 && j < p.ncols, in = i < p.nrows;
        const long long line0 = ((long long)b * p.S0 + seg0) * p.ncols + min(j, p.ncols - 1);
        float *ln = p.pn + line0 * NC * p.pitch0 + io;
        float *lt = p.tn + line0 * NUNK * p.pitch0 + io;
        auto put = [&](int row0, int row1, float v) {
            if (st) ln[(long long)row0 * p.pitch0] = in ? v : 0.f;
            tiles[(row1 * 32 + ty + 8 * r) * 33 + tx] = v;
        };
#pragma unroll
        for (int k = 0; k < NN; k++) put(wrow0[k], wrow1[k], vw[r][k]);
        if (NUNK == 2) put(TL_MM, TL_MM, vm[r]);
#pragma unroll
        for (int q = 0; q < NUNK; q++) {
            const int r0 = 3 * q, r1 = NUNK == 2 ? 3 * (1 - q) : 0;              // unknown q is solved q-th / (1-q)-th
            put(TL_D0 + r0, TL_D0 + r1, vd[r][q]);
            put(TL_C0 + r0, TL_C0 + r1, tl_cprime(vc[r][q], vd[r][q], vm[r], vx0[r][q], vx0[r][NUNK - 1 - q], F::LATE, NUNK == 2));
            if (st) lt[(long long)q * p.pitch0] = in ? (F::LATE ? vx0[r][q] + vx[r][q] : vx[r][q]) : 0.f;
        }
    }
    __syncthreads();
    const int jj = j0 + tx;
    const int seg1 = min(jj / p.ns1, p.S1 - 1), jo = jj - seg1 * p.ns1;
    if (jo < p.pitch1) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int ii = i0 + ty + 8 * r;
            if (ii < p.nrows) {
                float *l1 = p.pt + (((long long)b * p.S1 + seg1) * p.nrows + ii) * NC * p.pitch1 + jo;
#pragma unroll
                for (int f = 0; f < NC; f++) l1[(long long)f * p.pitch1] = jj < p.ncols ? tiles[(f * 32 + tx) * 33 + ty + 8 * r] : 0.f;
            }
        }
    }
}

// x = T - x0 (late linearisation) or x = T, from the packed T lines back into the caller's dense arrays
struct FinalParams { float *x[2]; const float *x0[2]; const float *tn; int nunk; int nrows, ncols, pitch0, vec, S0, ns0; long long xbs; };

static __global__ void __launch_bounds__(256)
tline_final_kernel(const FinalParams p)
{
    // one thread per 4 consecutive rows of one column (every thread of a block has work, whatever nrows is)
    const int b = blockIdx.y;
    const int n4 = (p.nrows + 3) >> 2;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n4 * p.ncols) return;
    const int j = (int)(t / n4), i = (int)(t - (long long)j * n4) * 4;
    const long long d = (long long)b * p.xbs + (long long)j * p.nrows + i;
    const int seg0 = i / p.ns0, io = i - seg0 * p.ns0;
    for (int q = 0; q < p.nunk; q++) {
        const float *t_ = p.tn + ((((long long)b * p.S0 + seg0) * p.ncols + j) * p.nunk + q) * p.pitch0 + io;
        if (p.vec) {                                          // 16-byte aligned caller arrays, lines a multiple of 4 long
            float4 v = *reinterpret_cast<const float4 *>(t_);
            if (p.x0[q]) {
                const float4 u = *reinterpret_cast<const float4 *>(p.x0[q] + d);
                v.x -= u.x; v.y -= u.y; v.z -= u.z; v.w -= u.w;
            }
            *reinterpret_cast<float4 *>(p.x[q] + d) = v;
        } else {
            for (int k = 0; k < 4 && i + k < p.nrows; k++) p.x[q][d + k] = t_[k] - (p.x0[q] ? p.x0[q][d + k] : 0.f);
        }
    }
}

struct TLGeom { int M, R, D, K, BL, NBR, NCW, SP; size_t smem; };

static int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

// Shared memory of one SM: the T ring (R lines), K coefficient slabs, barriers. Preference: the roomiest ring that
// still leaves 5 slabs, else the geometry with most slabs.
static bool tline_geometry(int n, int nunk, int nc, TLGeom &g)
{
    static const int Ms[] = {5, 9, 15, 17, 21, 25};
    g.M = 0;
    for (int m : Ms) if (32 * m >= n) { g.M = m; break; }
    if (!g.M || n < 8) return false;
    const int P = (n + 3) & ~3;                                // n: (longest) segment length
    const size_t room = 227 * 1024;
    static const int eBL = env_int("PDEGPU_TL_BL", 0), eR = env_int("PDEGPU_TL_R", 0), eD = env_int("PDEGPU_TL_D", 0), eK = env_int("PDEGPU_TL_K", 0);
    static const int eNCW = env_int("PDEGPU_TL_NCW", 0);
    const int total_warps = g.M <= 9 ? 16 : g.M <= 17 ? 12 : 8;
    g.NCW = eNCW > 0 && eNCW <= total_warps - 2 ? eNCW : total_warps - 2;
    // ring entries 4 floats apart from a multiple of 8: the two half-warps of a block write (lines 4h..4h+3) then read
    // different banks
    g.SP = nunk * P + ((nunk * P) % 8 == 0 ? 4 : 0);
    const size_t line = (size_t)g.SP * 4, slab = (size_t)nc * P * 4;
    // Candidates (BL, R, D), best first; the first that leaves 3 slabs wins. Measured (B200, 64 x 480x640, us per pass at
    // 480 / 640 elements): blocks of 8 lines (full 32-byte sectors in the transposed write) 223 / 229 against 268 / 285 for
    // blocks of 4; 4 to 8 slabs and rings of 24 to 32 lines within 3 % of each other once the slab is given back right
    // after the row formulas (profiles/r02_tline_sweep.txt).
    const int cand[][3] = {{8, 24, 4}, {8, 16, 4}, {4, 16, 3}, {4, 12, 2}, {4, 8, 2}};
    int bestK = 0;
    for (auto &c : cand) {
        const int BL = eBL ? eBL : c[0], R = eR ? eR : c[1], D = eD ? eD : c[2];
        if (R % BL || R < 2 * D + BL || (BL != 4 && BL != 8)) continue;
        const int NBR = R / BL + 2;
        const size_t fixed = (size_t)R * line + 2 * (size_t)(32 * g.M - P) * 4 + (size_t)(2 * R) * 8 + (size_t)(2 * NBR + 4) * 4 + 64;
        if (fixed >= room) continue;
        int K = (int)((room - fixed) / (slab + 16 + 32));
        if (K > 8) K = 8;
        if (eK && eK < K) K = eK;
        if (K < 2 || K <= bestK) continue;
        bestK = K;
        g.BL = BL; g.R = R; g.D = D; g.K = K; g.NBR = NBR;
        g.smem = fixed + (size_t)K * (slab + 16 + 32);
        if (K >= 3) break;
    }
    return bestK > 0;
}

template <int NUNK, int NN, int MODE, int M>
int tline_launch(pdegpu_ctx *ctx, const TLParams &p, const TLGeom &g, double bytes, const char *name)
{
    cudaError_t e = cudaFuncSetAttribute(tline_pass_kernel<NUNK, NN, MODE, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return pdegpu_check_cuda(ctx, e, "cudaFuncSetAttribute(tline_pass_kernel)");
    const int grid = p.TB < ctx->sm_count ? p.TB : ctx->sm_count;
    PDEGPU_PROF(ctx, name, bytes);
    tline_pass_kernel<NUNK, NN, MODE, M><<<grid, (g.NCW + 2) * 32, g.smem, ctx->stream>>>(p);
    PDEGPU_LAUNCH_CHECK(ctx, "tline_pass_kernel");
    return PDEGPU_OK;
}

template <int NUNK, int NN, int MODE>
int tline_pass(pdegpu_ctx *ctx, TLParams &p, int batch, double bytes, const char *name)
{
    constexpr int NC = 6 + (NUNK == 2 ? 3 : 0) + (NN == 8 ? 4 : 0);
    TLGeom g;
    if (!tline_geometry(p.n, NUNK, NC, g)) return PDEGPU_ERR_UNSUPPORTED;
    p.BL = g.BL; p.R = g.R; p.D = g.D; p.K = g.K; p.NBR = g.NBR; p.NCW = g.NCW; p.SP = g.SP;
    p.NB = (p.nlines + g.BL - 1) / g.BL; p.TB = p.NB * batch;
    switch (g.M) {
    case 5:  return tline_launch<NUNK, NN, MODE, 5>(ctx, p, g, bytes, name);
    case 9:  return tline_launch<NUNK, NN, MODE, 9>(ctx, p, g, bytes, name);
    case 15: return tline_launch<NUNK, NN, MODE, 15>(ctx, p, g, bytes, name);
    case 17: return tline_launch<NUNK, NN, MODE, 17>(ctx, p, g, bytes, name);
    case 21: return tline_launch<NUNK, NN, MODE, 21>(ctx, p, g, bytes, name);
    case 25: return tline_launch<NUNK, NN, MODE, 25>(ctx, p, g, bytes, name);
    default: return PDEGPU_ERR_UNSUPPORTED;
    }
}

// Lines of more than 800 elements are cut into segments of at most 544 (a multiple of 8 long, so that a block of lines
// of the other pass never straddles a cut; 544 = 32 lanes x 17: the chunk length with the best measured throughput).
struct SegPlan { int S, ns, nlast, P; };
static SegPlan seg_plan(int n)
{
    SegPlan s;
    if (n <= 800) { s.S = 1; s.ns = n; s.nlast = n; }
    else {
        s.S = (n + 543) / 544;
        s.ns = (((n + s.S - 1) / s.S) + 7) & ~7;
        if ((s.S - 1) * s.ns >= n) s.S = (n + s.ns - 1) / s.ns;
        s.nlast = n - (s.S - 1) * s.ns;
    }
    s.P = (s.ns + 3) & ~3;
    return s;
}

// launch of the preparation kernel, plain or fused (`terms` given: FLOW_LLIN4 only)
template <int FAM>
int tline_prep_launch(pdegpu_ctx *ctx, const PrepParams &pp, dim3 grid, size_t smem, const LlinTermsArgs *terms)
{
    if (terms && FAM == PDEGPU_FLOW_LLIN4) {
        PrepParams q = pp;
        q.terms = *terms;
        cudaError_t e = cudaFuncSetAttribute(tline_prep_kernel<FAM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return pdegpu_check_cuda(ctx, e, "cudaFuncSetAttribute(tline_prep_kernel)");
        tline_prep_kernel<FAM, true><<<grid, 256, smem, ctx->stream>>>(q);
    } else {
        cudaError_t e = cudaFuncSetAttribute(tline_prep_kernel<FAM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return pdegpu_check_cuda(ctx, e, "cudaFuncSetAttribute(tline_prep_kernel)");
        tline_prep_kernel<FAM, false><<<grid, 256, smem, ctx->stream>>>(pp);
    }
    PDEGPU_LAUNCH_CHECK(ctx, "tline_prep_kernel");
    return PDEGPU_OK;
}

template <int FAM>
int tline_run(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega, const LlinTermsArgs *terms = nullptr)
{
    using F = Fam<FAM>;
    constexpr int NUNK = F::NUNK, NN = F::EIGHT ? 8 : 4, MODE = F::PDE ? 1 : 0;
    constexpr int NC = 6 + (NUNK == 2 ? 3 : 0) + (NN == 8 ? 4 : 0);
    const int nr = sys->nrows, nc = sys->ncols, batch = sys->batch;
    if (F::PDE && F::EIGHT) iter = 1;                         // pdeSolvers.c:362: whatever the caller asks for
    if (iter <= 0) return PDEGPU_OK;
    if (nr < 8 || nc < 8) return PDEGPU_ERR_UNSUPPORTED;
    const SegPlan s0 = seg_plan(nr), s1 = seg_plan(nc);
    if (s0.nlast < 8 || s1.nlast < 8 || (long long)batch * (s0.S > s1.S ? s0.S : s1.S) > 65535) return PDEGPU_ERR_UNSUPPORTED;
    TLGeom g0, g1;
    if (!tline_geometry(s0.ns, NUNK, NC, g0) || !tline_geometry(s1.ns, NUNK, NC, g1)) return PDEGPU_ERR_UNSUPPORTED;
    const int pitch0 = s0.P, pitch1 = s1.P;
    // floats per field and problem: column-pass layout (S0 segments x ncols lines x pitch0), row-pass layout
    const long long F0 = (long long)s0.S * nc * pitch0, F1 = (long long)s1.S * nr * pitch1;
    if ((long long)batch * NC * (F0 > F1 ? F0 : F1) >= (1ll << 40)) return PDEGPU_ERR_UNSUPPORTED;

    // scratch: packed coefficient lines of both passes, packed T lines of both passes
    const size_t bytes_pn = (size_t)NC * F0 * batch * 4, bytes_pt = (size_t)NC * F1 * batch * 4;
    const size_t bytes_tn = (size_t)NUNK * F0 * batch * 4, bytes_tt = (size_t)NUNK * F1 * batch * 4;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    int rc = pdegpu_scratch_reserve(ctx, up(bytes_pn) + up(bytes_pt) + up(bytes_tn) + up(bytes_tt));
    if (rc) return rc;
    char *sp = ctx->scratch;
    float *PN = (float *)sp; sp += up(bytes_pn);
    float *PT = (float *)sp; sp += up(bytes_pt);
    float *TN = (float *)sp; sp += up(bytes_tn);
    float *TT = (float *)sp;

    {
        PrepParams pp;
        memset(&pp, 0, sizeof(pp));
        for (int k = 0; k < NN; k++) pp.w[k] = sys->w[k];
        pp.m = sys->m;
        for (int q = 0; q < NUNK; q++) { pp.c[q] = sys->c[q]; pp.d[q] = sys->d[q]; pp.x0[q] = sys->x0[q]; pp.x[q] = sys->x[q]; }
        pp.nrows = nr; pp.ncols = nc; pp.pitch0 = pitch0; pp.pitch1 = pitch1;
        pp.S0 = s0.S; pp.ns0 = s0.ns; pp.S1 = s1.S; pp.ns1 = s1.ns;
        pp.ibs = sys->batch_stride; pp.pn = PN; pp.pt = PT; pp.tn = TN;
        const size_t smem = (size_t)NC * 32 * 33 * sizeof(float);
        const int ci = nr > s0.S * pitch0 ? nr : s0.S * pitch0, cj = nc > s1.S * pitch1 ? nc : s1.S * pitch1;   // incl. the pads
        dim3 grid((ci + 31) / 32, (cj + 31) / 32, batch);
        const double fields_in = terms ? 4.0 + 3.0 * terms->c1 + (terms->gradmag ? 5.0 : 3.0) * terms->c2
                                       : NN + (NUNK == 2 ? 1 : 0) + NUNK * (F::LATE ? 4 : 3), fields_out = 2 * NC + NUNK;
        PDEGPU_PROF(ctx, terms ? "tline_prep_kernel<fused weights+terms>" : "tline_prep_kernel", 4.0 * (fields_in + fields_out) * nr * nc * batch);
        if ((rc = tline_prep_launch<FAM>(ctx, pp, grid, smem, terms))) return rc;
    }

    TLParams p0, p1;
    memset(&p0, 0, sizeof(p0));
    p0.coef = PN; p0.tin = TN; p0.tout = TT;
    p0.pitch = pitch0; p0.opitch = pitch1; p0.q0 = 0; p0.n = s0.ns; p0.nlines = nc; p0.omega = omega;
    p0.skip_border = (F::PDE && F::EIGHT) ? 1 : 0;               // pdeSolvers.c:1155,1290: the outermost lines are not relaxed
    p0.S = s0.S; p0.ns = s0.ns; p0.nlast = s0.nlast; p0.nfull = nr; p0.oS = s1.S; p0.ons = s1.ns;
    p1 = p0;
    p1.coef = PT; p1.tin = TT; p1.tout = TN;
    p1.pitch = pitch1; p1.opitch = pitch0; p1.q0 = 1; p1.n = s1.ns; p1.nlines = nr;   // (q0: scalar families only read it as "row pass")
    p1.S = s1.S; p1.ns = s1.ns; p1.nlast = s1.nlast; p1.nfull = nc; p1.oS = s0.S; p1.ons = s0.ns;

    const double pass_bytes = sweep_bytes<FAM>() * (double)nr * nc * batch;
    for (int it = 0; it < iter; it++) {
        if ((rc = tline_pass<NUNK, NN, MODE>(ctx, p0, batch * s0.S, pass_bytes, "tline_pass_kernel<dir0>"))) return rc;
        if ((rc = tline_pass<NUNK, NN, MODE>(ctx, p1, batch * s1.S, pass_bytes, "tline_pass_kernel<dir1,transposed>"))) return rc;
    }
    {
        FinalParams fp;
        memset(&fp, 0, sizeof(fp));
        for (int q = 0; q < NUNK; q++) { fp.x[q] = sys->x[q]; fp.x0[q] = F::LATE ? sys->x0[q] : nullptr; }
        fp.tn = TN; fp.nunk = NUNK; fp.nrows = nr; fp.ncols = nc; fp.pitch0 = pitch0; fp.xbs = sys->batch_stride;
        fp.S0 = s0.S; fp.ns0 = s0.ns;
        bool vec = (nr % 4 == 0) && (sys->batch_stride % 4 == 0);
        for (int q = 0; q < NUNK; q++) vec = vec && (((uintptr_t)sys->x[q] | (uintptr_t)(F::LATE ? sys->x0[q] : nullptr)) & 15) == 0;
        fp.vec = vec ? 1 : 0;
        dim3 grid((unsigned)(((long long)((nr + 3) / 4) * nc + 255) / 256), batch);
        PDEGPU_PROF(ctx, "tline_final_kernel", 4.0 * NUNK * (F::LATE ? 3 : 2) * nr * nc * batch);
        tline_final_kernel<<<grid, 256, 0, ctx->stream>>>(fp);
        PDEGPU_LAUNCH_CHECK(ctx, "tline_final_kernel");
    }
    return PDEGPU_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// SMALL PROBLEMS (the coarse levels of every pyramid / multigrid cycle): the WHOLE relax call in one launch, one CTA per
// problem, every field resident in shared memory from the first load to the last store (north_star subsystem 4).
// The persistent pass kernel above needs 10 launches per call (preparation, 2 x iter passes, recovery of the increment),
// each of which is pure launch + drain latency on a 37 x 49 level; here the total-field conversion happens on the way
// into shared memory, the sweeps (same zebra order, same row formulas, same partitioned Thomas solve: one warp per line)
// run between __syncthreads, and the increment is formed on the way out.
// Lines along j are strided in shared memory: an odd pitch keeps the lanes' chunks (M odd) on different banks.
// ------------------------------------------------------------------------------------------------------------------
struct SmallParams {
    float *x[2];
    const float *x0[2], *m, *c[2], *d[2], *w[4];
    int nr, nc, pitch, iter;
    long long bs;
    float omega;
};

constexpr int kSmallThreads = 512;

template <int FAM, int M>
__global__ void __launch_bounds__(kSmallThreads, 1)
tline_small_kernel(const SmallParams p)
{
    using F = Fam<FAM>;
    constexpr int NUNK = F::NUNK;
    constexpr bool PDE = F::PDE;
    extern __shared__ float sm[];
    const int nr = p.nr, nc = p.nc, P = p.pitch;
    const int FS = P * nc;                                    // floats per field
    float *W = sm;                                            // wW, wN, wE, wS
    float *MM = W + 4 * FS;                                   // (flow families)
    float *C = MM + (NUNK == 2 ? FS : 0), *D = C + NUNK * FS, *T = D + NUNK * FS;
    const long long base = (long long)blockIdx.x * p.bs;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = kSmallThreads / 32;
    for (int idx = threadIdx.x; idx < nr * nc; idx += kSmallThreads) {
        const int j = idx / nr, i = idx - j * nr, dst = j * P + i;
        const long long src = base + idx;
#pragma unroll
        for (int k = 0; k < 4; k++) W[k * FS + dst] = p.w[k][src];
        const float mm = NUNK == 2 ? p.m[src] : 0.f;
        if (NUNK == 2) MM[dst] = mm;
        float x0[NUNK];
#pragma unroll
        for (int q = 0; q < NUNK; q++) x0[q] = F::LATE ? p.x0[q][src] : 0.f;
#pragma unroll
        for (int q = 0; q < NUNK; q++) {
            const float dd = p.d[q][src];
            D[q * FS + dst] = dd;
            C[q * FS + dst] = tl_cprime(p.c[q][src], dd, mm, x0[q], x0[NUNK - 1 - q], F::LATE, NUNK == 2);
            T[q * FS + dst] = x0[q] + p.x[q][src];
        }
    }
    __syncthreads();
    const float omega = p.omega, om1 = 1.0f - p.omega;
    const int o = lane * M;
    for (int it = 0; it < p.iter; it++) {
#pragma unroll 1
        for (int dir = 0; dir < 2; dir++) {
            const int n = dir == 0 ? nr : nc, nl = dir == 0 ? nc : nr;
            const int es = dir == 0 ? 1 : P, ls = dir == 0 ? P : 1;      // element / line strides
            // roles (tline_common.cuh): column pass previous/next = N/S, lines -/+ = W/E; row pass W/E and N/S
            const float *Wp = W + (dir == 0 ? W_N : W_W) * FS, *Wn = W + (dir == 0 ? W_S : W_E) * FS;
            const float *Wl = W + (dir == 0 ? W_W : W_N) * FS, *Wh = W + (dir == 0 ? W_E : W_S) * FS;
#pragma unroll 1
            for (int colour = 0; colour < 2; colour++) {
#pragma unroll 1
                for (int l = colour + 2 * warp; l < nl; l += 2 * nwarps) {
                    const bool eLo = l > 0, eHi = l < nl - 1;
#pragma unroll 1
                    for (int qq = 0; qq < NUNK; qq++) {
                        const int q = (NUNK == 2 && dir == 1) ? 1 - qq : qq;           // row pass: second unknown first
                        float *Tq = T + q * FS;
                        const float *To = T + (NUNK - 1 - q) * FS, *Cq = C + q * FS, *Dq = D + q * FS;
                        float a[M], b[M], c[M], d[M];
#pragma unroll
                        for (int k = 0; k < M; k++) {
                            const int e = o + k, ec = min(e, n - 1);
                            const bool ok = e < n;
                            const int ix = l * ls + ec * es;
                            const float wp = e > 0 ? Wp[ix] : 0.f, wn = e < n - 1 ? Wn[ix] : 0.f;
                            const float wl = eLo ? Wl[ix] : 0.f, wh = eHi ? Wh[ix] : 0.f;
                            const float sw = (wl + wh) + (wp + wn);
                            const float cr = (eLo ? wl * Tq[ix - ls] : 0.f) + (eHi ? wh * Tq[ix + ls] : 0.f);
                            const float Cc = Cq[ix], Dd = Dq[ix];
                            float bb, dd;
                            if (PDE) {
                                const bool has = !is_nan(Dd);
                                bb = has ? Dd : sw; dd = has ? cr + Cc : cr;
                            } else {
                                const bool has = !is_nan(Cc);
                                bb = has ? sw + Dd : sw;
                                float t = Cc;
                                if (NUNK == 2) t -= MM[ix] * To[ix];
                                dd = has ? cr + t : cr;
                            }
                            a[k] = ok ? -wp : 0.f; c[k] = ok ? -wn : 0.f; b[k] = ok ? bb : 1.0f; d[k] = ok ? dd : 0.f;
                        }
                        chunk_solve<M>(a, c, b, d, lane);
#pragma unroll
                        for (int k = 0; k < M; k++) {
                            const int e = o + k;
                            if (e < n) { const int ix = l * ls + e * es; Tq[ix] = omega * d[k] + om1 * Tq[ix]; }
                        }
                        __syncwarp();
                    }
                }
                __syncthreads();
            }
        }
    }
    for (int idx = threadIdx.x; idx < nr * nc; idx += kSmallThreads) {
        const int j = idx / nr, i = idx - j * nr, src = j * P + i;
        const long long dst = base + idx;
#pragma unroll
        for (int q = 0; q < NUNK; q++) p.x[q][dst] = F::LATE ? T[q * FS + src] - p.x0[q][dst] : T[q * FS + src];
    }
}

template <int FAM>
int tline_small_run(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega)
{
    using F = Fam<FAM>;
    if (F::EIGHT) return PDEGPU_ERR_UNSUPPORTED;
    static const int enabled = env_int("PDEGPU_TL_SMALL", 1);
    if (!enabled || iter <= 0) return PDEGPU_ERR_UNSUPPORTED;
    const int nr = sys->nrows, nc = sys->ncols, nmax = nr > nc ? nr : nc;
    if (nr < 3 || nc < 3 || nmax > 288) return PDEGPU_ERR_UNSUPPORTED;
    const int pitch = nr | 1;
    const int nf = 4 + (F::NUNK == 2 ? 1 : 0) + 3 * F::NUNK;
    const size_t smem = (size_t)nf * pitch * nc * sizeof(float);
    if (smem > 227 * 1024) return PDEGPU_ERR_UNSUPPORTED;
    SmallParams sp;
    memset(&sp, 0, sizeof(sp));
    for (int q = 0; q < F::NUNK; q++) { sp.x[q] = sys->x[q]; sp.x0[q] = sys->x0[q]; sp.c[q] = sys->c[q]; sp.d[q] = sys->d[q]; }
    sp.m = sys->m;
    for (int k = 0; k < 4; k++) sp.w[k] = sys->w[k];
    sp.nr = nr; sp.nc = nc; sp.pitch = pitch; sp.iter = iter; sp.bs = sys->batch_stride; sp.omega = omega;
    const double bytes = sweep_bytes<FAM>() * 2.0 * iter * nr * nc * sys->batch;
    cudaError_t e;
    if (nmax <= 160) {
        e = cudaFuncSetAttribute(tline_small_kernel<FAM, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return pdegpu_check_cuda(ctx, e, "cudaFuncSetAttribute(tline_small_kernel)");
        PDEGPU_PROF(ctx, "tline_small_kernel", bytes);
        tline_small_kernel<FAM, 5><<<sys->batch, kSmallThreads, smem, ctx->stream>>>(sp);
    } else {
        e = cudaFuncSetAttribute(tline_small_kernel<FAM, 9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return pdegpu_check_cuda(ctx, e, "cudaFuncSetAttribute(tline_small_kernel)");
        PDEGPU_PROF(ctx, "tline_small_kernel", bytes);
        tline_small_kernel<FAM, 9><<<sys->batch, kSmallThreads, smem, ctx->stream>>>(sp);
    }
    PDEGPU_LAUNCH_CHECK(ctx, "tline_small_kernel");
    return PDEGPU_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// REFERENCE ORDER (sweeps_lex_impl.cuh): the same total-field formulation and packed lines, whole lines (no segments),
// relaxed one after the other as the reference does. Sequence per call: preparation, then per iteration the column
// pass in place on the column-layout T lines, their transposition, the row pass in place, the transposition back; the
// increment is recovered as in the zebra path.
// ------------------------------------------------------------------------------------------------------------------
struct LexGeom { int Mr, KS, RL, inslab; size_t smem; };

static bool lex_geometry(int n, int nunk, int nc, LexGeom &g)
{
    const int P = (n + 3) & ~3;
    g.Mr = ((n + 31) / 32) | 1;
    const size_t room = 227 * 1024;
    const size_t slab = (size_t)nc * P * 4, line = (size_t)nunk * P * 4;
    // (slabs, ring lines, eliminated rows inside the slab): two slabs first -- with one the two solver warps of a flow
    // family cannot overlap -- then the roomier ring
    const int cand[][3] = {{2, 5, 0}, {2, 4, 0}, {2, 3, 0}, {2, 4, 1}, {2, 3, 1}, {1, 4, 0}, {1, 3, 0}, {1, 3, 1}};
    static const int eKS = env_int("PDEGPU_LEX_KS", 0), eRL = env_int("PDEGPU_LEX_RL", 0);
    if (eKS > 0 && eRL >= 3) {                                 // (tuning override)
        const size_t scratch = (size_t)nunk * 32 * g.Mr * 16;
        const size_t need = eKS * slab + eRL * line + scratch + (size_t)(2 * eKS + 2 * eRL) * 8 + 128;
        if (need <= room) { g.KS = eKS; g.RL = eRL; g.inslab = 0; g.smem = need; return true; }
    }
    for (auto &c : cand) {
        const size_t scratch = (size_t)nunk * 32 * g.Mr * (c[2] ? 4 : 16);
        const size_t need = c[0] * slab + c[1] * line + scratch + (size_t)(2 * c[0] + 2 * c[1]) * 8 + 128;
        if (need <= room) { g.KS = c[0]; g.RL = c[1]; g.inslab = c[2]; g.smem = need; return true; }
    }
    return false;
}

template <int NUNK, int NN, int MODE>
int lex_pass(pdegpu_ctx *ctx, LexParams &p, int batch, double bytes, const char *name)
{
    constexpr int NC = 6 + (NUNK == 2 ? 3 : 0) + (NN == 8 ? 4 : 0);
    LexGeom g;
    if (!lex_geometry(p.n, NUNK, NC, g)) return PDEGPU_ERR_UNSUPPORTED;
    p.Mr = g.Mr; p.KS = g.KS; p.RL = g.RL; p.inslab = g.inslab;
    cudaError_t e = g.inslab ? cudaFuncSetAttribute(lex_pass_kernel<NUNK, NN, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                             : cudaFuncSetAttribute(lex_pass_kernel<NUNK, NN, MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return pdegpu_check_cuda(ctx, e, "cudaFuncSetAttribute(lex_pass_kernel)");
    PDEGPU_PROF(ctx, name, bytes);
    if (g.inslab) lex_pass_kernel<NUNK, NN, MODE, true><<<batch, LexThreads<NUNK>::value, g.smem, ctx->stream>>>(p);
    else          lex_pass_kernel<NUNK, NN, MODE, false><<<batch, LexThreads<NUNK>::value, g.smem, ctx->stream>>>(p);
    PDEGPU_LAUNCH_CHECK(ctx, "lex_pass_kernel");
    return PDEGPU_OK;
}

template <int FAM>
int lex_run(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega, const LlinTermsArgs *terms = nullptr)
{
    using F = Fam<FAM>;
    constexpr int NUNK = F::NUNK, NN = F::EIGHT ? 8 : 4, MODE = F::PDE ? 1 : 0;
    constexpr int NC = 6 + (NUNK == 2 ? 3 : 0) + (NN == 8 ? 4 : 0);
    const int nr = sys->nrows, nc = sys->ncols, batch = sys->batch;
    if (F::PDE && F::EIGHT) iter = 1;                         // pdeSolvers.c:362
    if (iter <= 0) return PDEGPU_OK;
    LexGeom g0, g1;
    if (!lex_geometry(nr, NUNK, NC, g0) || !lex_geometry(nc, NUNK, NC, g1))
        return pdegpu_set_error(ctx, PDEGPU_ERR_UNSUPPORTED, "reference-order line relaxation: lines of %d / %d elements do not fit the shared memory of an SM", nr, nc);
    const int pitch0 = (nr + 3) & ~3, pitch1 = (nc + 3) & ~3;
    const long long F0 = (long long)nc * pitch0, F1 = (long long)nr * pitch1;
    const size_t bytes_pn = (size_t)NC * F0 * batch * 4, bytes_pt = (size_t)NC * F1 * batch * 4;
    const size_t bytes_tn = (size_t)NUNK * F0 * batch * 4, bytes_tt = (size_t)NUNK * F1 * batch * 4;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    int rc = pdegpu_scratch_reserve(ctx, up(bytes_pn) + up(bytes_pt) + up(bytes_tn) + up(bytes_tt));
    if (rc) return rc;
    char *sp = ctx->scratch;
    float *PN = (float *)sp; sp += up(bytes_pn);
    float *PT = (float *)sp; sp += up(bytes_pt);
    float *TN = (float *)sp; sp += up(bytes_tn);
    float *TT = (float *)sp;
    {
        PrepParams pp;
        memset(&pp, 0, sizeof(pp));
        for (int k = 0; k < NN; k++) pp.w[k] = sys->w[k];
        pp.m = sys->m;
        for (int q = 0; q < NUNK; q++) { pp.c[q] = sys->c[q]; pp.d[q] = sys->d[q]; pp.x0[q] = sys->x0[q]; pp.x[q] = sys->x[q]; }
        pp.nrows = nr; pp.ncols = nc; pp.pitch0 = pitch0; pp.pitch1 = pitch1;
        pp.S0 = 1; pp.ns0 = pitch0 > nr ? pitch0 : nr; pp.S1 = 1; pp.ns1 = pitch1 > nc ? pitch1 : nc;
        pp.ibs = sys->batch_stride; pp.pn = PN; pp.pt = PT; pp.tn = TN;
        const size_t smem = (size_t)NC * 32 * 33 * sizeof(float);
        dim3 grid((pitch0 + 31) / 32, (pitch1 + 31) / 32, batch);
        PDEGPU_PROF(ctx, terms ? "tline_prep_kernel<fused weights+terms>" : "tline_prep_kernel", 0);
        if ((rc = tline_prep_launch<FAM>(ctx, pp, grid, smem, terms))) return rc;
    }
    LexParams p0, p1;
    memset(&p0, 0, sizeof(p0));
    p0.coef = PN; p0.t = TN; p0.pitch = pitch0; p0.q0 = 0; p0.n = nr; p0.nlines = nc; p0.omega = omega;
    p0.skip_border = (F::PDE && F::EIGHT) ? 1 : 0;
    p1 = p0;
    p1.coef = PT; p1.t = TT; p1.pitch = pitch1; p1.q0 = 1; p1.n = nc; p1.nlines = nr;
    const double pass_bytes = sweep_bytes<FAM>() * (double)nr * nc * batch;
    for (int it = 0; it < iter; it++) {
        if ((rc = lex_pass<NUNK, NN, MODE>(ctx, p0, batch, pass_bytes, "lex_pass_kernel<dir0>"))) return rc;
        {
            dim3 grid((nr + 31) / 32, (nc + 31) / 32, batch * NUNK);
            PDEGPU_PROF(ctx, "lex_transpose_kernel", 0);
            lex_transpose_kernel<<<grid, 256, 0, ctx->stream>>>(TT, TN, nr, nc, pitch0, pitch1, NUNK);
            PDEGPU_LAUNCH_CHECK(ctx, "lex_transpose_kernel");
        }
        if ((rc = lex_pass<NUNK, NN, MODE>(ctx, p1, batch, pass_bytes, "lex_pass_kernel<dir1>"))) return rc;
        {
            dim3 grid((nc + 31) / 32, (nr + 31) / 32, batch * NUNK);
            PDEGPU_PROF(ctx, "lex_transpose_kernel", 0);
            lex_transpose_kernel<<<grid, 256, 0, ctx->stream>>>(TN, TT, nc, nr, pitch1, pitch0, NUNK);
            PDEGPU_LAUNCH_CHECK(ctx, "lex_transpose_kernel");
        }
    }
    {
        FinalParams fp;
        memset(&fp, 0, sizeof(fp));
        for (int q = 0; q < NUNK; q++) { fp.x[q] = sys->x[q]; fp.x0[q] = F::LATE ? sys->x0[q] : nullptr; }
        fp.tn = TN; fp.nunk = NUNK; fp.nrows = nr; fp.ncols = nc; fp.pitch0 = pitch0; fp.xbs = sys->batch_stride;
        fp.S0 = 1; fp.ns0 = pitch0 > nr ? pitch0 : nr;
        bool vec = (nr % 4 == 0) && (sys->batch_stride % 4 == 0);
        for (int q = 0; q < NUNK; q++) vec = vec && (((uintptr_t)sys->x[q] | (uintptr_t)(F::LATE ? sys->x0[q] : nullptr)) & 15) == 0;
        fp.vec = vec ? 1 : 0;
        dim3 grid((unsigned)(((long long)((nr + 3) / 4) * nc + 255) / 256), batch);
        PDEGPU_PROF(ctx, "tline_final_kernel", 0);
        tline_final_kernel<<<grid, 256, 0, ctx->stream>>>(fp);
        PDEGPU_LAUNCH_CHECK(ctx, "tline_final_kernel");
    }
    return PDEGPU_OK;
}

}  // namespace

#ifdef TL_PROBE
extern "C" int pdegpu_debug_tl_probe(unsigned long long *out32)
{
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out32, g_tl_probe, sizeof(g_tl_probe)) != cudaSuccess) return -1;
    unsigned long long z[32] = {0};
    return cudaMemcpyToSymbol(g_tl_probe, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
#endif

int relax_tline(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega)
{
    {   // small problems: the whole call in one launch, resident in shared memory
        int rc = PDEGPU_ERR_UNSUPPORTED;
        switch (sys->family) {
        case PDEGPU_FLOW_ELIN4: rc = tline_small_run<PDEGPU_FLOW_ELIN4>(ctx, sys, iter, omega); break;
        case PDEGPU_FLOW_LLIN4: rc = tline_small_run<PDEGPU_FLOW_LLIN4>(ctx, sys, iter, omega); break;
        case PDEGPU_DISP_LLIN4: rc = tline_small_run<PDEGPU_DISP_LLIN4>(ctx, sys, iter, omega); break;
        case PDEGPU_PDE4:       rc = tline_small_run<PDEGPU_PDE4>(ctx, sys, iter, omega); break;
        default: break;
        }
        if (rc != PDEGPU_ERR_UNSUPPORTED) return rc;
    }
    switch (sys->family) {
    case PDEGPU_FLOW_ELIN4: return tline_run<PDEGPU_FLOW_ELIN4>(ctx, sys, iter, omega);
    case PDEGPU_FLOW_LLIN4: return tline_run<PDEGPU_FLOW_LLIN4>(ctx, sys, iter, omega);
    case PDEGPU_FLOW_LLIN8: return tline_run<PDEGPU_FLOW_LLIN8>(ctx, sys, iter, omega);
    case PDEGPU_DISP_LLIN4: return tline_run<PDEGPU_DISP_LLIN4>(ctx, sys, iter, omega);
    case PDEGPU_PDE4:       return tline_run<PDEGPU_PDE4>(ctx, sys, iter, omega);
    case PDEGPU_PDE8:       return tline_run<PDEGPU_PDE8>(ctx, sys, iter, omega);
    default: return PDEGPU_ERR_UNSUPPORTED;
    }
}

// solver 2 in the reference's (lexicographic) line order: pdegpu_set_sweep_order(ctx, PDEGPU_ORDER_REFERENCE)
int relax_lexline(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega)
{
    if (sys->nrows < 3 || sys->ncols < 3) return PDEGPU_ERR_UNSUPPORTED;
    switch (sys->family) {
    case PDEGPU_FLOW_ELIN4: return lex_run<PDEGPU_FLOW_ELIN4>(ctx, sys, iter, omega);
    case PDEGPU_FLOW_LLIN4: return lex_run<PDEGPU_FLOW_LLIN4>(ctx, sys, iter, omega);
    case PDEGPU_FLOW_LLIN8: return lex_run<PDEGPU_FLOW_LLIN8>(ctx, sys, iter, omega);
    case PDEGPU_DISP_LLIN4: return lex_run<PDEGPU_DISP_LLIN4>(ctx, sys, iter, omega);
    case PDEGPU_PDE4:       return lex_run<PDEGPU_PDE4>(ctx, sys, iter, omega);
    case PDEGPU_PDE8:       return lex_run<PDEGPU_PDE8>(ctx, sys, iter, omega);
    default: return PDEGPU_ERR_UNSUPPORTED;
    }
}

// The late-linearisation flow driver's inner solve with the diffusion weights and the data terms computed INSIDE the
// preparation of the packed lines (pdegpu_dev_llin_solve, api.cu). sys: family FLOW_LLIN4 with x, x0 set; w, m, c, d are
// not read. PDEGPU_ERR_UNSUPPORTED: this call takes another path (small problems resident in shared memory, short
// lines), the caller runs the three steps one after the other.
int relax_llin_fused(pdegpu_ctx *ctx, const pdegpu_system *sys, const pdegpu_llin_terms *t, int iter, float omega, bool reference_order)
{
    if (iter <= 0) return PDEGPU_ERR_UNSUPPORTED;
    // OFF unless PDEGPU_FUSE=1. Measured (B200, 16 pairs of 640x480, whole llin driver): the fused preparation takes
    // 31.7 ms per batch against 21.8 ms for the four kernels it replaces (two axpby, op_diff_weights, llin_terms) plus the
    // plain preparation: it holds 124 registers for 4 pixels per thread (16 warps per SM) and evaluates the double-
    // precision weight stencil serially per pixel, where the stand-alone kernels run one pixel per thread at full
    // occupancy and already stream at 3-5 TB/s. 512 launches fewer per batch, 6 % fewer flows/s.
    static const int enabled = env_int("PDEGPU_FUSE", 0);
    if (!enabled) return PDEGPU_ERR_UNSUPPORTED;
    LlinTermsArgs a;
    memset(&a, 0, sizeof a);
    for (int k = 0; k < 3; k++) a.d1[k] = t->d1[k];
    for (int k = 0; k < 5; k++) a.d2[k] = t->d2[k];
    a.c1 = t->channels1; a.c2 = t->channels2; a.gradmag = t->gradmag;
    a.b1 = t->b1; a.b2 = t->b2; a.alpha = t->alpha;
    a.npix = (long long)t->nrows * t->ncols;
    a.stride1 = t->batch_stride1; a.stride2 = t->batch_stride2; a.stride = t->batch_stride;
    if (reference_order) return lex_run<PDEGPU_FLOW_LLIN4>(ctx, sys, iter, omega, &a);
    {   // the shared-memory resident kernel takes dense arrays: leave small problems to the unfused sequence
        const int nr = sys->nrows, nc = sys->ncols, nmax = nr > nc ? nr : nc;
        const size_t smem = (size_t)11 * (nr | 1) * nc * sizeof(float);
        if (env_int("PDEGPU_TL_SMALL", 1) && nmax <= 288 && smem <= 227 * 1024) return PDEGPU_ERR_UNSUPPORTED;
    }
    return tline_run<PDEGPU_FLOW_LLIN4>(ctx, sys, iter, omega, &a);
}
