// sweeps_simple.cu -- kernel generation 0: straightforward global-memory relaxation kernels.
//
// They define the SEMANTICS of libpdegpu's two parallel orderings and stay in the library as an
// independent cross-check of the streaming kernels (pdegpu_set_kernel_path(ctx, 0) or
// PDEGPU_KERNELS=simple):
//
//   solver 1  point SOR in RED-BLACK order (colour = (i+j)&1; 8-neighbour stencils use the 4
//             colours (i&1)+2(j&1)), interior pixels only, then the reference's border fill
//             (opticalflowSolvers.c:161-179): every border pixel := nearest interior pixel.
//   solver 2  alternating line relaxation in ZEBRA order. One iteration =
//               lines along i (Matlab columns):  even lines {unknown 0, unknown 1}, odd lines {0, 1}
//               lines along j (Matlab rows):     even lines {unknown 1, unknown 0}, odd lines {1, 0}
//             (the reference relaxes U then V vertically and V then U horizontally,
//              opticalflowSolvers.c:238-257). Every pixel of a line is an unknown, border rows are
//             one-sided, SOR is applied to the whole line after the Thomas solve (:1849-1862).
//             The 8-neighbour PDE solver relaxes interior lines only and runs exactly one
//             iteration (pdeSolvers.c:362,1155,1290).
#include "stencil_math.cuh"

// ---------------------------------------------------------------------------------------------
// solver 1
// ---------------------------------------------------------------------------------------------
template <int FAM>
__global__ void __launch_bounds__(128)
rb_point_kernel(SysView s, int colour, float omega)
{
    constexpr bool FOUR = Fam<FAM>::PDE && Fam<FAM>::EIGHT;   // 4-colour ordering
    const int j = 1 + blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    int i;
    if (FOUR) {
        if ((j & 1) != (colour >> 1)) return;
        i = 1 + ((1 ^ colour) & 1) + 2 * t;                   // first interior i with (i&1)==(colour&1)
    } else {
        i = 1 + ((1 + j + colour) & 1) + 2 * t;               // first interior i with ((i+j)&1)==colour
    }
    if (i > s.nrows - 2) return;
    const long long pos = (long long)blockIdx.z * s.bstride + (long long)j * s.nrows + i;
    point_update<FAM>(s, pos, omega);
}

template <int NUNK>
__global__ void border_fill_kernel(SysView s)
{
    // perimeter index -> (i,j); value := nearest interior pixel (read interior, write border: no race)
    const int nr = s.nrows, nc = s.ncols;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int per = 2 * nr + 2 * nc;
    if (t >= per) return;
    int i, j;
    if (t < nr)               { i = t;            j = 0; }
    else if (t < 2 * nr)      { i = t - nr;       j = nc - 1; }
    else if (t < 2 * nr + nc) { i = 0;            j = t - 2 * nr; }
    else                      { i = nr - 1;       j = t - 2 * nr - nc; }
    const int ic = min(max(i, 1), nr - 2), jc = min(max(j, 1), nc - 2);
    const long long base = (long long)blockIdx.y * s.bstride;
#pragma unroll
    for (int q = 0; q < NUNK; q++)
        s.x[q][base + (long long)j * nr + i] = s.x[q][base + (long long)jc * nr + ic];
}

template <int FAM>
static int run_point(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega)
{
    SysView v = make_view(sys);
    constexpr bool FOUR = Fam<FAM>::PDE && Fam<FAM>::EIGHT;
    const int ni = sys->nrows - 2, nj = sys->ncols - 2;
    dim3 block(128), grid(((ni + 1) / 2 + 127) / 128, nj, sys->batch);
    const int per = 2 * sys->nrows + 2 * sys->ncols;
    dim3 bgrid((per + 127) / 128, sys->batch);
    for (int it = 0; it < iter; it++) {
        for (int colour = 0; colour < (FOUR ? 4 : 2); colour++) {
            PDEGPU_PROF(ctx, "rb_point_kernel", sweep_bytes<FAM>() * (double)ni * nj * sys->batch / (FOUR ? 4 : 2));
            rb_point_kernel<FAM><<<grid, block, 0, ctx->stream>>>(v, colour, omega);
            PDEGPU_LAUNCH_CHECK(ctx, "rb_point_kernel");
        }
        PDEGPU_PROF(ctx, "border_fill_kernel", 0);
        border_fill_kernel<Fam<FAM>::NUNK><<<bgrid, 128, 0, ctx->stream>>>(v);
        PDEGPU_LAUNCH_CHECK(ctx, "border_fill_kernel");
    }
    return PDEGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// solver 2: one thread per line, classic Thomas with (cp,dp) in global scratch.
// scratch layout: [k][line slot] so that neighbouring threads touch neighbouring words.
// ---------------------------------------------------------------------------------------------
template <int FAM, int DIR>
__global__ void __launch_bounds__(64)
zebra_line_kernel(SysView s, int colour, int q, float omega, float *__restrict__ cp, float *__restrict__ dp, int first_line, int nslots)
{
    const int nr = s.nrows, nc = s.ncols;
    const int n = DIR == 0 ? nr : nc;                 // line length
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= nslots) return;
    const int line = first_line + colour + 2 * slot;  // caller guarantees line <= last line
    const long long base = (long long)blockIdx.y * s.bstride;
    const long long sbase = (long long)blockIdx.y * nslots * n;
    const long long step = DIR == 0 ? 1 : nr;
    const long long p0 = base + (DIR == 0 ? (long long)line * nr : (long long)line);
    float *X = s.x[q];

    float a, b, c, d, cprev, dprev;
    {
        line_eq<FAM, DIR>(s, p0, q, DIR == 0 ? 0 : line, DIR == 0 ? line : 0, a, b, c, d);
        cprev = c / b;
        dprev = d / b;
        cp[sbase + slot] = cprev;
        dp[sbase + slot] = dprev;
    }
    for (int k = 1; k < n; k++) {
        const long long p = p0 + k * step;
        line_eq<FAM, DIR>(s, p, q, DIR == 0 ? k : line, DIR == 0 ? line : k, a, b, c, d);
        const float div = 1.0f / (b - cprev * a);
        cprev = c * div;
        dprev = (d - dprev * a) * div;
        cp[sbase + (long long)k * nslots + slot] = cprev;
        dp[sbase + (long long)k * nslots + slot] = dprev;
    }
    // back substitution with the un-relaxed solution, SOR applied per element
    float xn = dprev;                                  // x_{n-1}
    {
        const long long p = p0 + (long long)(n - 1) * step;
        X[p] = omega * xn + (1.0f - omega) * X[p];
    }
    for (int k = n - 2; k >= 0; k--) {
        const long long p = p0 + k * step;
        xn = dp[sbase + (long long)k * nslots + slot] - cp[sbase + (long long)k * nslots + slot] * xn;
        X[p] = omega * xn + (1.0f - omega) * X[p];
    }
}

template <int FAM, int DIR>
static int run_line_pass(pdegpu_ctx *ctx, const SysView &v, const pdegpu_system *sys, int colour, int q, float omega)
{
    constexpr bool INTERIOR_ONLY = Fam<FAM>::PDE && Fam<FAM>::EIGHT;
    const int nlines = DIR == 0 ? sys->ncols : sys->nrows;
    const int n = DIR == 0 ? sys->nrows : sys->ncols;
    const int first = INTERIOR_ONLY ? 1 : 0, last = INTERIOR_ONLY ? nlines - 2 : nlines - 1;
    if (first + colour > last) return PDEGPU_OK;
    const int nslots = (last - (first + colour)) / 2 + 1;
    const size_t need = 2ull * sizeof(float) * (size_t)nslots * n * sys->batch;
    int rc = pdegpu_scratch_reserve(ctx, need);
    if (rc) return rc;
    float *cp = (float *)ctx->scratch, *dp = cp + (size_t)nslots * n * sys->batch;
    dim3 block(64), grid((nslots + 63) / 64, sys->batch);
    PDEGPU_PROF(ctx, DIR == 0 ? "zebra_line_kernel<dir0>" : "zebra_line_kernel<dir1>",
                sweep_bytes<FAM>() * (double)nslots * n * sys->batch / Fam<FAM>::NUNK);
    zebra_line_kernel<FAM, DIR><<<grid, block, 0, ctx->stream>>>(v, colour, q, omega, cp, dp, first, nslots);
    PDEGPU_LAUNCH_CHECK(ctx, "zebra_line_kernel");
    return PDEGPU_OK;
}

template <int FAM>
static int run_line(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega)
{
    SysView v = make_view(sys);
    constexpr int NUNK = Fam<FAM>::NUNK;
    if (Fam<FAM>::PDE && Fam<FAM>::EIGHT) iter = 1;            // pdeSolvers.c:362
    int rc;
    // interior-only lines (8-neighbour PDE) start at line 1: colour 1 there = the even lines, relaxed first as everywhere
    constexpr int cflip = (Fam<FAM>::PDE && Fam<FAM>::EIGHT) ? 1 : 0;
    for (int it = 0; it < iter; it++) {
        for (int cc = 0; cc < 2; cc++)
            for (int q = 0; q < NUNK; q++)
                if ((rc = run_line_pass<FAM, 0>(ctx, v, sys, cc ^ cflip, q, omega))) return rc;
        for (int cc = 0; cc < 2; cc++)
            for (int q = NUNK - 1; q >= 0; q--)
                if ((rc = run_line_pass<FAM, 1>(ctx, v, sys, cc ^ cflip, q, omega))) return rc;
    }
    return PDEGPU_OK;
}

template <int FAM>
static int run_family(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega, int solver)
{
    return solver == 1 ? run_point<FAM>(ctx, sys, iter, omega) : run_line<FAM>(ctx, sys, iter, omega);
}

// ---------------------------------------------------------------------------------------------
// solver 1 in the REFERENCE'S ORDER (pdegpu_set_sweep_order(ctx, PDEGPU_ORDER_REFERENCE)): lexicographic point
// Gauss-Seidel -- for j, for i, in place (GS_SOR_elin4_2d opticalflowSolvers.c:89-158, GS_SOR_llin4_2d :563-655,
// disparitySolvers.c:89-125, pdeSolvers.c:94-125, :208-247), border fill after every sweep. Pixel (i, j) sees the new
// (i-1, j) and (i, j-1) [8-neighbour: also the new (i-1, j-1) and (i+1, j-1)] and old values elsewhere: all pixels with
// the same i + j (8-neighbour: i + 2j) are independent, so a CTA walks the anti-diagonals of one problem with a barrier
// between them. The parallelism is the batch and the length of a diagonal; same point_update as generation 0.
// ---------------------------------------------------------------------------------------------
template <int FAM>
__global__ void __launch_bounds__(512)
lex_point_kernel(SysView s, int iter, float omega)
{
    constexpr int S = (Fam<FAM>::PDE && Fam<FAM>::EIGHT) ? 2 : 1;
    constexpr int NUNK = Fam<FAM>::NUNK;
    const int nr = s.nrows, nc = s.ncols;
    const long long base = (long long)blockIdx.x * s.bstride;
    const int per = 2 * nr + 2 * nc;
    for (int it = 0; it < iter; it++) {
        for (int d = 1 + S; d <= (nr - 2) + S * (nc - 2); d++) {
            // i + S j = d with 1 <= i <= nr-2, 1 <= j <= nc-2
            const int jlo = max(1, (d - (nr - 2) + S - 1) / S), jhi = min(nc - 2, (d - 1) / S);
            for (int j = jlo + threadIdx.x; j <= jhi; j += blockDim.x)
                point_update<FAM>(s, base + (long long)j * nr + (d - S * j), omega);
            __syncthreads();
        }
        for (int t = threadIdx.x; t < per; t += blockDim.x) {         // border pixel := nearest interior pixel
            int i, j;
            if (t < nr)               { i = t;            j = 0; }
            else if (t < 2 * nr)      { i = t - nr;       j = nc - 1; }
            else if (t < 2 * nr + nc) { i = 0;            j = t - 2 * nr; }
            else                      { i = nr - 1;       j = t - 2 * nr - nc; }
            const int ic = min(max(i, 1), nr - 2), jc = min(max(j, 1), nc - 2);
#pragma unroll
            for (int q = 0; q < NUNK; q++) s.x[q][base + (long long)j * nr + i] = s.x[q][base + (long long)jc * nr + ic];
        }
        __syncthreads();
    }
}

template <int FAM>
static int run_lex_point(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega)
{
    if (iter <= 0) return PDEGPU_OK;
    SysView v = make_view(sys);
    PDEGPU_PROF(ctx, "lex_point_kernel", sweep_bytes<FAM>() * (double)sys->nrows * sys->ncols * sys->batch * iter);
    lex_point_kernel<FAM><<<sys->batch, 512, 0, ctx->stream>>>(v, iter, omega);
    PDEGPU_LAUNCH_CHECK(ctx, "lex_point_kernel");
    return PDEGPU_OK;
}

int relax_lexpoint(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega)
{
    switch (sys->family) {
    case PDEGPU_FLOW_ELIN4: return run_lex_point<PDEGPU_FLOW_ELIN4>(ctx, sys, iter, omega);
    case PDEGPU_FLOW_LLIN4:
    case PDEGPU_FLOW_LLIN8: return run_lex_point<PDEGPU_FLOW_LLIN4>(ctx, sys, iter, omega);
    case PDEGPU_DISP_LLIN4: return run_lex_point<PDEGPU_DISP_LLIN4>(ctx, sys, iter, omega);
    case PDEGPU_PDE4:       return run_lex_point<PDEGPU_PDE4>(ctx, sys, iter, omega);
    case PDEGPU_PDE8:       return run_lex_point<PDEGPU_PDE8>(ctx, sys, iter, omega);
    default: return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "relax: unknown family %d", sys->family);
    }
}

int relax_simple(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega, int solver)
{
    switch (sys->family) {
    case PDEGPU_FLOW_ELIN4: return run_family<PDEGPU_FLOW_ELIN4>(ctx, sys, iter, omega, solver);
    case PDEGPU_FLOW_LLIN4: return run_family<PDEGPU_FLOW_LLIN4>(ctx, sys, iter, omega, solver);
    case PDEGPU_FLOW_LLIN8:
        // point solver of the 8-neighbour flow family ignores the diagonals
        return solver == 1 ? run_point<PDEGPU_FLOW_LLIN4>(ctx, sys, iter, omega)
                           : run_line<PDEGPU_FLOW_LLIN8>(ctx, sys, iter, omega);
    case PDEGPU_DISP_LLIN4: return run_family<PDEGPU_DISP_LLIN4>(ctx, sys, iter, omega, solver);
    case PDEGPU_PDE4:       return run_family<PDEGPU_PDE4>(ctx, sys, iter, omega, solver);
    case PDEGPU_PDE8:       return run_family<PDEGPU_PDE8>(ctx, sys, iter, omega, solver);
    default: return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "relax: unknown family %d", sys->family);
    }
}
