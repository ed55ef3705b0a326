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
#include "stencil_math.cuh"

namespace {

constexpr int kMaxSmem = 220 * 1024;

template <int NUNK> struct Slots;
template <> struct Slots<2> { enum { A = 0, C = 1, B1 = 2, D1 = 3, B2 = 4, D2 = 5, M = 6, GP = 7, XO = 8, N = 9 }; };
template <> struct Slots<1> { enum { A = 0, C = 1, B1 = 2, D1 = 3, GP = 4, XO = 5, B2 = 2, D2 = 3, M = 0, N = 6 }; };

// Tridiagonal rows of all unknowns at pixel (i,j) for a line along DIR; `qa` is the unknown solved
// first. d[qa] is complete (coupling taken with the other unknown's current value); d[qb] lacks the
// coupling term, which is m * x_qa(new) and is added by the solver.
template <int FAM, int DIR>
__device__ __forceinline__ void assemble(const SysView &s, long long pos, int i, int j,
                                         float &a, float &c, float (&b)[2], float (&d)[2], float &m)
{
    using F = Fam<FAM>;
    constexpr int NN = F::EIGHT ? 8 : 4;
    constexpr int prev = DIR == 0 ? W_N : W_W, next = DIR == 0 ? W_S : W_E;
    constexpr int qa = (F::NUNK == 2 && DIR == 1) ? 1 : 0, qb = 1 - qa;
    const int nr = s.nrows, nc = s.ncols;
    const bool eN = i > 0, eS = i < nr - 1, eW = j > 0, eE = j < nc - 1;
    const bool ex[8] = {eW, eN, eE, eS, eN && eW, eN && eE, eS && eE, eS && eW};
    const long long off[8] = {-(long long)nr, -1, (long long)nr, 1, -(long long)nr - 1, (long long)nr - 1, (long long)nr + 1, -(long long)nr + 1};
    float w[NN];
#pragma unroll
    for (int n = 0; n < NN; n++) w[n] = s.w[n][pos];
    a = ex[prev] ? -w[prev] : 0.0f;
    c = ex[next] ? -w[next] : 0.0f;
    float bsum = 0.0f, dsum[2] = {0.0f, 0.0f};
    float x0c[2] = {0.0f, 0.0f};
    if (F::LATE) {
#pragma unroll
        for (int q = 0; q < F::NUNK; q++) x0c[q] = s.x0[q][pos];
    }
#pragma unroll
    for (int n = 0; n < NN; n++) {
        if (!ex[n]) continue;
        bsum += w[n];
        const bool inl = (n == prev) || (n == next);
#pragma unroll
        for (int q = 0; q < F::NUNK; q++) {
            if (F::LATE) {
                float t = s.x0[q][pos + off[n]] - x0c[q];
                if (!inl) t += s.x[q][pos + off[n]];
                dsum[q] += w[n] * t;
            } else if (!inl) {
                dsum[q] += w[n] * s.x[q][pos + off[n]];
            }
        }
    }
    m = 0.0f;
    b[1] = 1.0f; d[1] = 0.0f;
    if (F::PDE) {
        const float tr = s.d[0][pos];
        if (!is_nan(tr)) { b[0] = tr; d[0] = dsum[0] + s.c[0][pos]; }
        else {
            if (F::EIGHT)   // pdeSolvers.c:1179: wNW twice, wNE never
                b[0] = (s.w[W_N][pos] + s.w[W_S][pos] + s.w[W_W][pos] + s.w[W_E][pos])
                     + (s.w[W_NW][pos] + s.w[W_NW][pos] + s.w[W_SW][pos] + s.w[W_SE][pos]);
            else b[0] = bsum;
            d[0] = dsum[0];
        }
        return;
    }
#pragma unroll
    for (int q = 0; q < F::NUNK; q++) {
        const float C = s.c[q][pos];
        b[q] = bsum; d[q] = dsum[q];
        if (!is_nan(C)) {
            b[q] += s.d[q][pos];
            d[q] += C;
            if (F::NUNK == 2) {
                const float M = s.m[pos];
                if (q == qa) d[q] -= M * s.x[qb][pos];
                else m = M;
            }
        }
    }
}

// ---- phase B: one warp solves one line held in shared memory -------------------------------
// Solves a_k x_{k-1} + b_k x_k + c_k x_{k+1} = d_k (- mm_k * y_k when COUPLED), k = 0..n-1, in place:
// on return D[k] holds omega*x_k + (1-omega)*XO[k] (RELAX) or x_k. B and GP are clobbered.
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
    for (int r = 0; r < len; r++) {
        const float a = A[o + r], b = B[o + r], c = Cc[o + r];
        float d = D[o + r];
        if (COUPLED) d -= MM[o + r] * Y[o + r];
        float inv, ng;
        if (r == 0) { inv = 1.0f / b; dp = d * inv; ng = a * inv; }
        else        { inv = 1.0f / (b - a * cp); dp = (d - a * dp) * inv; ng = -a * gp * inv; }
        cp = c * inv; gp = ng;
        B[o + r] = cp; D[o + r] = dp; GP[o + r] = gp;
    }
    // last-row relation of the chunk:  x_t + cp*x_{t+1} + gp*x_{s0-1} = dp
    // pass 2: first-row relation  x_s = Af - Bf*x_t - Gf*x_{s0-1}
    float Af = 0.f, Bf = 0.f, Gf = 0.f;
    if (len == 1) Bf = -1.0f;
    else if (len >= 2) {
        Af = D[o + len - 2]; Bf = B[o + len - 2]; Gf = GP[o + len - 2];
        for (int r = len - 3; r >= 0; r--) {
            const float cpr = B[o + r];
            Af = D[o + r] - cpr * Af;
            Bf = -cpr * Bf;
            Gf = GP[o + r] - cpr * Gf;
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
        const float k1 = al / bem, k2 = ga / bep;
        be = be - gam * k1 - alp * k2;
        de = de - dem * k1 - dep * k2;
        al = -alm * k1;
        ga = -gap * k2;
    }
    const float l = de / be;
    float L = __shfl_up_sync(FULL, l, 1);
    if (lane == 0) L = 0.f;
    // pass 3: local back substitution
    float x = l;
    for (int r = len - 1; r >= 0; r--) {
        if (r < len - 1) x = D[o + r] - B[o + r] * x - GP[o + r] * L;
        D[o + r] = RELAX ? omega * x + (1.0f - omega) * XO[o + r] : x;
    }
}

template <int FAM, int DIR, int G>
__global__ void __launch_bounds__(32 * G)
alr_kernel(SysView s, int colour, float omega, int first_line, int nslots, int Lc, int Lp)
{
    using F = Fam<FAM>;
    using S = Slots<F::NUNK>;
    extern __shared__ float smem[];
    constexpr int qa = (F::NUNK == 2 && DIR == 1) ? 1 : 0, qb = 1 - qa;
    const int nr = s.nrows, nc = s.ncols;
    const int n = DIR == 0 ? nr : nc;
    const int NP = 32 * Lp;                                  // padded line length
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot0 = blockIdx.x * G;                        // first line slot of this CTA
    const long long base = (long long)blockIdx.y * s.bstride;
    const int nl = min(G, nslots - slot0);                   // lines really present

    // ---- phase A ----
    const int total = nl * n;
    for (int t = tid; t < total; t += 32 * G) {
        int g, k;
        if (DIR == 0) { g = t / n; k = t - g * n; }          // k (= i) fastest: 128-B rows of a column
        else          { k = t / nl; g = t - k * nl; }        // line (= i) fastest: one sector per j
        const int line = first_line + colour + 2 * (slot0 + g);
        const int i = DIR == 0 ? k : line, j = DIR == 0 ? line : k;
        const long long pos = base + (long long)j * nr + i;
        float a, c, b[2], d[2], m;
        assemble<FAM, DIR>(s, pos, i, j, a, c, b, d, m);
        const int ad = (k / Lc) * Lp + (k % Lc);
        float *Ls = smem + (size_t)g * S::N * NP;
        Ls[S::A * NP + ad] = a;
        Ls[S::C * NP + ad] = c;
        Ls[S::B1 * NP + ad] = b[qa];
        Ls[S::D1 * NP + ad] = d[qa];
        Ls[S::XO * NP + ad] = s.x[qa][pos];
        if (F::NUNK == 2) {
            Ls[S::B2 * NP + ad] = b[qb];
            Ls[S::D2 * NP + ad] = d[qb];
            Ls[S::M * NP + ad] = m;
        }
    }
    __syncthreads();

    // ---- phase B ----
    if (warp < nl) {
        float *Ls = smem + (size_t)warp * S::N * NP;
        warp_line_solve<false, true>(Ls + S::A * NP, Ls + S::C * NP, Ls + S::B1 * NP, Ls + S::D1 * NP, Ls + S::GP * NP,
                                     nullptr, nullptr, Ls + S::XO * NP, n, Lc, Lp, omega, lane);
        if (F::NUNK == 2) {
            __syncwarp();
            warp_line_solve<true, false>(Ls + S::A * NP, Ls + S::C * NP, Ls + S::B2 * NP, Ls + S::D2 * NP, Ls + S::GP * NP,
                                         Ls + S::M * NP, Ls + S::D1 * NP, nullptr, n, Lc, Lp, omega, lane);
        }
    }
    __syncthreads();

    // ---- phase C ----
    for (int t = tid; t < total; t += 32 * G) {
        int g, k;
        if (DIR == 0) { g = t / n; k = t - g * n; }
        else          { k = t / nl; g = t - k * nl; }
        const int line = first_line + colour + 2 * (slot0 + g);
        const int i = DIR == 0 ? k : line, j = DIR == 0 ? line : k;
        const long long pos = base + (long long)j * nr + i;
        const int ad = (k / Lc) * Lp + (k % Lc);
        const float *Ls = smem + (size_t)g * S::N * NP;
        s.x[qa][pos] = Ls[S::D1 * NP + ad];
        if (F::NUNK == 2) {
            float *X2 = s.x[qb];
            X2[pos] = omega * Ls[S::D2 * NP + ad] + (1.0f - omega) * X2[pos];
        }
    }
}

template <int FAM, int DIR, int G>
int launch_alr(pdegpu_ctx *ctx, const SysView &v, const pdegpu_system *sys, int colour, float omega,
               int first, int nslots, int n, int Lc, int Lp, size_t smem)
{
    static bool attr_set[16] = {false};
    if (!attr_set[ctx->device & 15]) {
        cudaError_t e = cudaFuncSetAttribute(alr_kernel<FAM, DIR, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
        if (e != cudaSuccess) return pdegpu_check_cuda(ctx, e, "cudaFuncSetAttribute(alr_kernel)");
        attr_set[ctx->device & 15] = true;
    }
    dim3 grid((nslots + G - 1) / G, sys->batch);
    PDEGPU_PROF(ctx, DIR == 0 ? "alr_kernel<dir0>" : "alr_kernel<dir1>", sweep_bytes<FAM>() * (double)nslots * n * sys->batch);
    alr_kernel<FAM, DIR, G><<<grid, 32 * G, smem, ctx->stream>>>(v, colour, omega, first, nslots, Lc, Lp);
    PDEGPU_LAUNCH_CHECK(ctx, "alr_kernel");
    return PDEGPU_OK;
}

template <int FAM, int DIR>
int alr_pass(pdegpu_ctx *ctx, const SysView &v, const pdegpu_system *sys, int colour, float omega)
{
    using F = Fam<FAM>;
    constexpr bool INTERIOR_ONLY = F::PDE && F::EIGHT;
    const int nlines = DIR == 0 ? sys->ncols : sys->nrows;
    const int n = DIR == 0 ? sys->nrows : sys->ncols;
    const int first = INTERIOR_ONLY ? 1 : 0, last = INTERIOR_ONLY ? nlines - 2 : nlines - 1;
    if (first + colour > last) return PDEGPU_OK;
    const int nslots = (last - (first + colour)) / 2 + 1;
    const int Lc = (n + 31) / 32, Lp = Lc | 1;
    const size_t line_bytes = (size_t)Slots<F::NUNK>::N * 32 * Lp * sizeof(float);
    // G lines per CTA: as many as keep >= 2 CTAs per SM resident; lines along j (DIR 1) are strided in
    // memory, so they want >= 4 neighbouring lines per CTA to use the sectors they fetch.
    int G = 8;
    while (G > 1 && G * line_bytes > (size_t)kMaxSmem / 2) G >>= 1;
    if (DIR == 1 && G < 4) { G = 4; while (G > 1 && G * line_bytes > (size_t)kMaxSmem) G >>= 1; }
    if (G * line_bytes > (size_t)kMaxSmem) return PDEGPU_ERR_UNSUPPORTED;   // line too long for shared memory
    const size_t smem = G * line_bytes;
    switch (G) {
    case 8: return launch_alr<FAM, DIR, 8>(ctx, v, sys, colour, omega, first, nslots, n, Lc, Lp, smem);
    case 4: return launch_alr<FAM, DIR, 4>(ctx, v, sys, colour, omega, first, nslots, n, Lc, Lp, smem);
    case 2: return launch_alr<FAM, DIR, 2>(ctx, v, sys, colour, omega, first, nslots, n, Lc, Lp, smem);
    default: return launch_alr<FAM, DIR, 1>(ctx, v, sys, colour, omega, first, nslots, n, Lc, Lp, smem);
    }
}

template <int FAM>
int alr_run(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega)
{
    SysView v = make_view(sys);
    if (Fam<FAM>::PDE && Fam<FAM>::EIGHT) iter = 1;            // pdeSolvers.c:362
    // refuse up front (before touching the unknowns) if either direction does not fit
    {
        const size_t per = (size_t)Slots<Fam<FAM>::NUNK>::N * 32 * sizeof(float);
        const int L0 = ((sys->nrows + 31) / 32) | 1, L1 = ((sys->ncols + 31) / 32) | 1;
        if (per * L0 > (size_t)kMaxSmem || per * L1 > (size_t)kMaxSmem) return PDEGPU_ERR_UNSUPPORTED;
    }
    int rc;
    for (int it = 0; it < iter; it++) {
        for (int colour = 0; colour < 2; colour++) if ((rc = alr_pass<FAM, 0>(ctx, v, sys, colour, omega))) return rc;
        for (int colour = 0; colour < 2; colour++) if ((rc = alr_pass<FAM, 1>(ctx, v, sys, colour, omega))) return rc;
    }
    return PDEGPU_OK;
}

}  // namespace

int relax_stream_line(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega)
{
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
