// pipeline.cu -- device-resident restatement of the reference's late-linearisation warping flow driver
// (BASELINE.json configs[1], the "640x480 flows/s" metric): matlab/optical_flow/FlowEminND_llin_2D_v10.m.
//
// The driver's structure is kept step for step (pyramid :99-127, feature channels :135-166, coarse-to-fine
// loop :195-367); every step is one libpdegpu kernel on the context's stream, a whole BATCH of image pairs
// moves through each launch, and nothing returns to the host between the upload of the frames and the
// download of the flow. Steps that are MEX calls in the reference (BilinInterp_2d, Fst/SndDerivatives5,
// Oflow_sor_llin4_2d) use the same kernels as the gateways; steps that are Matlab code use driver_ops.cu.
// Not restated: the optional spatial a-priori terms (param.Us/Vs, default empty, :175-193,259-268,301-316).
#include "pdegpu_internal.cuh"
#include <math.h>
#include <vector>

namespace {

struct Bump {
    char *base; size_t cap, used; bool dry;
    float *take(size_t nfloats)
    {
        const size_t bytes = (nfloats * sizeof(float) + 255) & ~(size_t)255;
        float *p = dry ? nullptr : (float *)(base + used);
        used += bytes;
        return p;
    }
};

struct Level { int nr, nc; };

// fspecial('gaussian', [5 5], sigma), column-major
void gaussian5(double sigma, double *h)
{
    double mx = 0, sum = 0;
    for (int b = 0; b < 5; b++) for (int a = 0; a < 5; a++) {
        const double x = b - 2, y = a - 2;
        h[b * 5 + a] = exp(-(x * x + y * y) / (2.0 * sigma * sigma));
        mx = fmax(mx, h[b * 5 + a]);
    }
    for (int k = 0; k < 25; k++) { if (h[k] < 2.220446049250313e-16 * mx) h[k] = 0; sum += h[k]; }
    for (int k = 0; k < 25; k++) h[k] /= sum;
}

int fill_zero(pdegpu_ctx *ctx, float *p, size_t n)
{
    PDEGPU_CUDA_OK(ctx, cudaMemsetAsync(p, 0, n * sizeof(float), ctx->stream));
    return PDEGPU_OK;
}

#define RC(call) do { int rc__ = (call); if (rc__) return rc__; } while (0)

// One pass over the algorithm; with b.dry it only measures the workspace.
int flow_llin_run(pdegpu_ctx *ctx, Bump &b, float *Uout, float *Vout, const float *I0in, const float *I1in,
                  int nrows, int ncols, int C, int B, const pdegpu_flow_llin_params &P)
{
    const bool dry = b.dry;
    // ---- pyramid sizes (FlowEminND_llin_2D_v10.m:105-127): imresize(.., scl_factor) -> ceil(size*scl) ----
    std::vector<Level> L;
    L.push_back({nrows, ncols});
    const int max_scales = P.max_scales > 0 ? P.max_scales : (1 << 30);
    while ((int)L.size() < max_scales) {
        const Level &p = L.back();
        Level n = {(int)ceil(p.nr * P.scl_factor), (int)ceil(p.nc * P.scl_factor)};
        L.push_back(n);
        if (n.nr <= 20 || n.nc <= 20) break;
    }
    const int S = (int)L.size();
    if (L.back().nr < 5 || L.back().nc < 5) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "flow_llin: coarsest level smaller than 5 pixels");
    const int c1 = P.fst_grad ? 2 * C : C;                    // channels of the first constancy term
    const int c2 = P.snd_term ? C : 0;                        // second term works on the plain channels
    const bool gradmag = P.snd_term == 2;
    const size_t np0 = (size_t)nrows * ncols;

    // ---- pyramid of both frames; level k of frame f: It[f][k], C*B planes ----
    std::vector<float *> It[2], F1[2];
    float *tmp = b.take(np0 * C * B);                         // scratch for imresize / smoothing / medfilt
    float *tmp2 = b.take(np0 * C * B);
    for (int f = 0; f < 2; f++) {
        It[f].resize(S); F1[f].resize(S);
        for (int k = 0; k < S; k++) {
            It[f][k] = b.take((size_t)L[k].nr * L[k].nc * C * B);
            F1[f][k] = P.fst_grad ? b.take((size_t)L[k].nr * L[k].nc * c1 * B) : It[f][k];
        }
    }
    double G[25];
    gaussian5(1.25, G);
    if (!dry) {
        for (int f = 0; f < 2; f++) {
            // Iin = single(Iin)./255   (:73)
            RC(op_axpby_div(ctx, It[f][0], f == 0 ? I0in : I1in, 255.0f, np0 * C * B));
            for (int k = 1; k < S; k++) {
                // next level from the UNSMOOTHED current one, then smooth the current one (:107-114)
                RC(pdegpu_dev_imresize_bilinear(ctx, It[f][k], tmp, It[f][k - 1], L[k - 1].nr, L[k - 1].nc, L[k].nr, L[k].nc,
                                                P.scl_factor, P.scl_factor, 1, C * B));
                const size_t n = (size_t)L[k - 1].nr * L[k - 1].nc;
                RC(op_imfilter(ctx, tmp2, It[f][k - 1], L[k - 1].nr, L[k - 1].nc, C * B, n, n, G, 5, 5, 1, 1.0f));
                PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(It[f][k - 1], tmp2, n * C * B * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
            }
            {   // smooth the last scale (:121-122)
                const size_t n = (size_t)L[S - 1].nr * L[S - 1].nc;
                RC(op_imfilter(ctx, tmp2, It[f][S - 1], L[S - 1].nr, L[S - 1].nc, C * B, n, n, G, 5, 5, 1, 1.0f));
                PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(It[f][S - 1], tmp2, n * C * B * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
            }
            if (P.fst_grad) {
                // rgb2grad (:374-384): channel i -> channels 2i (d/dx, [1 0 -1] along columns) and 2i+1 (d/dy)
                const double odx[3] = {1.0, 0.0, -1.0};
                for (int k = 0; k < S; k++) {
                    const size_t n = (size_t)L[k].nr * L[k].nc;
                    RC(op_imfilter(ctx, F1[f][k], It[f][k], L[k].nr, L[k].nc, C * B, n, 2 * n, odx, 1, 3, 1, 1.0f));
                    RC(op_imfilter(ctx, F1[f][k] + n, It[f][k], L[k].nr, L[k].nc, C * B, n, 2 * n, odx, 3, 1, 1, 1.0f));
                }
            }
        }
    }

    // ---- per-level work arrays, sized for the finest level ----
    float *U = b.take(np0 * B), *V = b.take(np0 * B), *dU = b.take(np0 * B), *dV = b.take(np0 * B);
    float *Us = b.take(np0 * B), *Vs = b.take(np0 * B);       // U+dU, V+dV
    float *X = b.take(np0 * B), *Y = b.take(np0 * B);
    float *W1 = b.take(np0 * c1 * B), *W2 = c2 ? b.take(np0 * c2 * B) : nullptr;
    float *D1[3], *D2[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < 3; k++) D1[k] = b.take(np0 * c1 * B);
    for (int k = 0; k < (c2 ? (gradmag ? 5 : 3) : 0); k++) D2[k] = b.take(np0 * c2 * B);
    float *lwork = b.take(np0 * B * 11);                      // pdegpu_dev_llin_solve's scratch (only its unfused path uses it)
    if (dry) return PDEGPU_OK;

    const float up = (float)(1.0 / P.scl_factor);
    for (int s = S - 1; s >= 0; s--) {
        const int nr = L[s].nr, nc = L[s].nc;
        const size_t n = (size_t)nr * nc;
        if (s == S - 1) { RC(fill_zero(ctx, U, n * B)); RC(fill_zero(ctx, V, n * B)); }   // zero flow at the coarsest level (:208-211)
        for (int fl = 0; fl < P.firstLoop; fl++) {
            // warp the second frame's features by the current flow (:222,228)
            RC(op_warp_coords(ctx, X, Y, U, V, nr, nc, B, n));
            RC(op_bilin_batch(ctx, W1, F1[1][s], X, Y, nr, nc, c1, B, P.oob_value));
            RC(op_fst(ctx, D1[0], D1[1], D1[2], F1[0][s], W1, nr, nc, c1 * B));
            if (c2) {
                RC(op_bilin_batch(ctx, W2, It[1][s], X, Y, nr, nc, c2, B, P.oob_value));
                if (gradmag) RC(op_snd(ctx, D2[0], D2[1], D2[2], D2[3], D2[4], It[0][s], W2, nr, nc, c2 * B));
                else         RC(op_fst(ctx, D2[0], D2[1], D2[2], It[0][s], W2, nr, nc, c2 * B));
            }
            RC(fill_zero(ctx, dU, n * B)); RC(fill_zero(ctx, dV, n * B));
            for (int sl = 0; sl < P.secondLoop; sl++) {
                // OPdiffWeights (:321), robust weights + channel sums (:289-327), Oflow_sor_llin4_2d (:332-348): one call,
                // fused into the line kernels' preparation where they run from packed lines (pdegpu_dev_llin_solve)
                pdegpu_llin_terms t;
                memset(&t, 0, sizeof t);
                t.nrows = nr; t.ncols = nc; t.batch = B; t.channels1 = c1; t.channels2 = c2; t.gradmag = gradmag;
                t.b1 = (float)P.b1; t.b2 = (float)P.b2; t.alpha = (float)P.alpha;
                for (int k = 0; k < 3; k++) t.d1[k] = D1[k];
                for (int k = 0; k < 5; k++) t.d2[k] = D2[k];
                t.batch_stride1 = n * c1; t.batch_stride2 = n * (c2 ? c2 : 1); t.batch_stride = n;
                RC(pdegpu_dev_llin_solve(ctx, &t, U, V, dU, dV, lwork, P.iter, (float)P.omega, P.solver));
            }
            // U = medfilt2(U + dU, [3 3], 'symmetric')  (:354-355)
            RC(op_axpby(ctx, Us, 1.0f, U, 1.0f, dU, n * B));
            RC(op_axpby(ctx, Vs, 1.0f, V, 1.0f, dV, n * B));
            RC(op_medfilt3(ctx, U, Us, nr, nc, B, n));
            RC(op_medfilt3(ctx, V, Vs, nr, nc, B, n));
        }
        if (s > 0) {
            // U = imresize(U.*(1/scl_factor), 'OutputSize', isizes{scl-1}, 'Method', 'triangle')  (:365-366)
            const int onr = L[s - 1].nr, onc = L[s - 1].nc;
            RC(op_axpby(ctx, Us, up, U, 0.0f, nullptr, n * B));
            RC(op_axpby(ctx, Vs, up, V, 0.0f, nullptr, n * B));
            RC(pdegpu_dev_imresize_bilinear(ctx, U, tmp, Us, nr, nc, onr, onc, (double)onr / nr, (double)onc / nc, 1, B));
            RC(pdegpu_dev_imresize_bilinear(ctx, V, tmp, Vs, nr, nc, onr, onc, (double)onr / nr, (double)onc / nc, 1, B));
        }
    }
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(Uout, U, np0 * B * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(Vout, V, np0 * B * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    return PDEGPU_OK;
}

}  // namespace

extern "C" void pdegpu_flow_llin_default_params(pdegpu_flow_llin_params *p)
{
    // defaults of FlowEminND_llin_2D_v10.m:52-68
    p->alpha = 0.0420; p->omega = 1.9; p->b1 = 1.4843; p->b2 = 0.2915; p->scl_factor = 0.75;
    p->firstLoop = 4; p->secondLoop = 4; p->iter = 4; p->solver = 2;
    p->fst_grad = 1; p->snd_term = 2;          // runme.m: FlowEminND_llin_2D_v10(Iin, 3, 'grad', 'gradmag')
    p->max_scales = 0;
    p->oob_value = nanf("");
}

extern "C" int pdegpu_dev_flow_llin_2d(pdegpu_ctx *ctx, float *U, float *V, const float *I0, const float *I1,
        int nrows, int ncols, int channels, int batch, const pdegpu_flow_llin_params *params)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    if (!U || !V || !I0 || !I1 || !params) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_flow_llin_2d: null pointer");
    if (nrows < 5 || ncols < 5 || channels < 1 || batch < 1) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "pdegpu_dev_flow_llin_2d: bad shape");
    if (!(params->scl_factor > 0.1 && params->scl_factor < 1.0) || params->firstLoop < 1 || params->secondLoop < 1)
        return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_flow_llin_2d: bad parameters");
    PDEGPU_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    Bump b = {nullptr, 0, 0, true};
    int rc = flow_llin_run(ctx, b, U, V, I0, I1, nrows, ncols, channels, batch, *params);
    if (rc) return rc;
    if ((rc = pdegpu_work_reserve(ctx, b.used, "pdegpu_dev_flow_llin_2d"))) return rc;
    struct Args { pdegpu_ctx *ctx; float *U, *V; const float *I0, *I1; int nrows, ncols, channels, batch; pdegpu_flow_llin_params P; char *work; size_t wb; int id; };
    Args a;
    memset(&a, 0, sizeof a);                                   // (padding is part of the graph key)
    a.ctx = ctx; a.U = U; a.V = V; a.I0 = I0; a.I1 = I1; a.nrows = nrows; a.ncols = ncols; a.channels = channels; a.batch = batch;
    a.P = *params; a.work = ctx->work; a.wb = ctx->work_bytes; a.id = 1;
    pdegpu_graph_body body = {[](void *p) -> int {
        Args &a = *static_cast<Args *>(p);
        Bump w = {a.work, a.wb, 0, false};
        return flow_llin_run(a.ctx, w, a.U, a.V, a.I0, a.I1, a.nrows, a.ncols, a.channels, a.batch, a.P);
    }, &a};
    return pdegpu_graph_run(ctx, &a, sizeof a, body);
}

extern "C" int pdegpu_flow_llin_2d(pdegpu_ctx *ctx, float *U, float *V, const float *I0, const float *I1,
        int nrows, int ncols, int channels, int batch, const pdegpu_flow_llin_params *params)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    if (!U || !V || !I0 || !I1 || !params) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_flow_llin_2d: null pointer");
    PDEGPU_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const size_t nimg = (size_t)nrows * ncols * channels * batch * sizeof(float), nflow = (size_t)nrows * ncols * batch * sizeof(float);
    pdegpu_arena_reset(ctx);
    int rc = pdegpu_arena_reserve(ctx, 2 * nimg + 2 * nflow + 4096);
    if (rc) return rc;
    float *d0 = (float *)pdegpu_arena_alloc(ctx, nimg), *d1 = (float *)pdegpu_arena_alloc(ctx, nimg);
    float *dU = (float *)pdegpu_arena_alloc(ctx, nflow), *dV = (float *)pdegpu_arena_alloc(ctx, nflow);
    if (!d0 || !d1 || !dU || !dV) return pdegpu_set_error(ctx, PDEGPU_ERR_NOMEM, "pdegpu_flow_llin_2d: arena exhausted");
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(d0, I0, nimg, cudaMemcpyHostToDevice, ctx->stream));
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(d1, I1, nimg, cudaMemcpyHostToDevice, ctx->stream));
    rc = pdegpu_dev_flow_llin_2d(ctx, dU, dV, d0, d1, nrows, ncols, channels, batch, params);
    if (rc) return rc;
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(U, dU, nflow, cudaMemcpyDeviceToHost, ctx->stream));
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(V, dV, nflow, cudaMemcpyDeviceToHost, ctx->stream));
    PDEGPU_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return PDEGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// TVdenoise8 (matlab/denoising/TVdenoise8.m, BASELINE configs[3] "8-neighbour anisotropic TV denoising"):
// two-level pyramid (:59-75), per level outer_iter+1 lagged-diffusivity steps of
// { ADdiffWeights (:81), PsiData/TRACE/B (:83-85), PDEsolver8 (:87-100) }, bilinear up-sampling (:112).
// The driver's typo at :72 (`Itin{scl} = imfilter(...)`: the coarsest level is NOT smoothed) is kept.
// ---------------------------------------------------------------------------------------------
namespace {

int tv_run(pdegpu_ctx *ctx, Bump &b, float *Iout, const float *Iin, int nrows, int ncols, int F, const pdegpu_tvdenoise8_params &P)
{
    const bool dry = b.dry;
    const int r1 = (int)ceil(nrows * P.scl_factor), c1 = (int)ceil(ncols * P.scl_factor);
    const size_t n0 = (size_t)nrows * ncols, n1 = (size_t)r1 * c1;
    float *in0 = b.take(n0 * F), *in1 = b.take(n1 * F), *X = b.take(n0 * F), *X2 = b.take(n0 * F), *tmp = b.take(n0 * F);
    float *TR = b.take(n0 * F), *BB = b.take(n0 * F), *w[8];
    for (int k = 0; k < 8; k++) w[k] = b.take(n0 * F);
    if (dry) return PDEGPU_OK;
    double G[25];
    gaussian5(1.25, G);
    // Iin{2} from the unsmoothed Iin{1}, then Iin{1} smoothed (:60-66); the size test of :68 is always met at scl = 2
    RC(pdegpu_dev_imresize_bilinear(ctx, in1, tmp, Iin, nrows, ncols, r1, c1, P.scl_factor, P.scl_factor, 1, F));
    RC(op_imfilter(ctx, in0, Iin, nrows, ncols, F, n0, n0, G, 5, 5, 1, 1.0f));
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(X, in1, n1 * F * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));     // Iout = Iin{scales}
    for (int s = 1; s >= 0; s--) {
        const int nr = s ? r1 : nrows, nc = s ? c1 : ncols;
        const size_t n = (size_t)nr * nc;
        const float *in = s ? in1 : in0;
        for (int it = 0; it <= P.outer_iter; it++) {
            RC(op_ad_diff_weights(ctx, w, TR, BB, X, in, nr, nc, F, 0.5, P.alpha, nullptr, F));
            pdegpu_system sys;
            memset(&sys, 0, sizeof sys);
            sys.family = PDEGPU_PDE8; sys.nrows = nr; sys.ncols = nc; sys.batch = F; sys.batch_stride = n;
            sys.x[0] = X; sys.c[0] = BB; sys.d[0] = TR;
            // ADdiffWeights returns [W NW N NE E SE S SW]
            sys.w[W_W] = w[0]; sys.w[W_NW] = w[1]; sys.w[W_N] = w[2]; sys.w[W_NE] = w[3];
            sys.w[W_E] = w[4]; sys.w[W_SE] = w[5]; sys.w[W_S] = w[6]; sys.w[W_SW] = w[7];
            RC(pdegpu_dev_relax(ctx, &sys, P.inner_iter, (float)P.omega, P.solver));
        }
        if (s > 0) {
            RC(pdegpu_dev_imresize_bilinear(ctx, X2, tmp, X, nr, nc, nrows, ncols, (double)nrows / nr, (double)ncols / nc, 1, F));
            PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(X, X2, n0 * F * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        }
    }
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(Iout, X, n0 * F * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    return PDEGPU_OK;
}

}  // namespace

extern "C" void pdegpu_tvdenoise8_default_params(pdegpu_tvdenoise8_params *p)
{
    p->alpha = 500; p->omega = 1.75; p->outer_iter = 20; p->inner_iter = 4; p->solver = 2; p->scl_factor = 0.75;   // TVdenoise8.m:36-44
}

extern "C" int pdegpu_dev_tvdenoise8(pdegpu_ctx *ctx, float *Iout, const float *Iin, int nrows, int ncols, int nframes,
        const pdegpu_tvdenoise8_params *params)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    if (!Iout || !Iin || !params) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_tvdenoise8: null pointer");
    if (nrows < 8 || ncols < 8 || nframes < 1) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "pdegpu_dev_tvdenoise8: bad shape");
    PDEGPU_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    Bump b = {nullptr, 0, 0, true};
    int rc = tv_run(ctx, b, Iout, Iin, nrows, ncols, nframes, *params);
    if (rc) return rc;
    if ((rc = pdegpu_work_reserve(ctx, b.used, "pdegpu_dev_tvdenoise8"))) return rc;
    Bump w = {ctx->work, ctx->work_bytes, 0, false};
    return tv_run(ctx, w, Iout, Iin, nrows, ncols, nframes, *params);
}

extern "C" int pdegpu_tvdenoise8(pdegpu_ctx *ctx, float *Iout, const float *Iin, int nrows, int ncols, int nframes,
        const pdegpu_tvdenoise8_params *params)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    if (!Iout || !Iin || !params) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_tvdenoise8: null pointer");
    PDEGPU_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const size_t nb = (size_t)nrows * ncols * nframes * sizeof(float);
    pdegpu_arena_reset(ctx);
    int rc = pdegpu_arena_reserve(ctx, 2 * nb + 1024);
    if (rc) return rc;
    float *di = (float *)pdegpu_arena_alloc(ctx, nb), *dout = (float *)pdegpu_arena_alloc(ctx, nb);
    if (!di || !dout) return pdegpu_set_error(ctx, PDEGPU_ERR_NOMEM, "pdegpu_tvdenoise8: arena exhausted");
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(di, Iin, nb, cudaMemcpyHostToDevice, ctx->stream));
    rc = pdegpu_dev_tvdenoise8(ctx, dout, di, nrows, ncols, nframes, params);
    if (rc) return rc;
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(Iout, dout, nb, cudaMemcpyDeviceToHost, ctx->stream));
    PDEGPU_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return PDEGPU_OK;
}
