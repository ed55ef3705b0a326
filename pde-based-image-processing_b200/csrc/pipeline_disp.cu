// pipeline_disp.cu -- device-resident restatement of the reference's symmetric stereo driver (BASELINE.json
// configs[3], "symmetric-constraint stereo disparity"): matlab/disparity/DispEminND_llin_sym_2D.m.
//
// Two disparity fields, U(:,:,1) (left -> right) and U(:,:,2) (right -> left), each with its own late-linearisation
// data term and coupled by a Lorentzian symmetry term. The driver's structure is kept step for step: bilinear x0.75
// pyramid with 3x3 Gaussian smoothing (:86-103), coarse to fine (:111-268); per outer iteration the cross warps of
// the images (BilinInterp_2d) and of the disparities (interp2), derivatives (Fst/SndDerivatives5), symmetry terms
// (:159-180); per inner iteration robust weights + term assembly (:188-225), DdiffWeights, Disp_sor_llin_sym4_2d
// (:227-247); median filter (:255-256), bilinear up-sampling (:265-267).
// Both views live side by side as a "batch" of two (view v at offset v*npix), so the inner solve is ONE relax call
// on two independent scalar systems, exactly what the reference's symmetric solver is.
#include "pdegpu_internal.cuh"
#include <math.h>
#include <vector>

namespace {

struct Bump2 {
    char *base; size_t used; bool dry;
    float *take(size_t nfloats)
    {
        const size_t bytes = (nfloats * sizeof(float) + 255) & ~(size_t)255;
        float *p = dry ? nullptr : (float *)(base + used);
        used += bytes;
        return p;
    }
};

#define RC(call) do { int rc__ = (call); if (rc__) return rc__; } while (0)

// fspecial('gaussian', [3 3], sigma), column-major
void gaussian3(double sigma, double *h)
{
    double sum = 0;
    for (int b = 0; b < 3; b++) for (int a = 0; a < 3; a++) {
        const double x = b - 1, y = a - 1;
        h[b * 3 + a] = exp(-(x * x + y * y) / (2.0 * sigma * sigma));
        sum += h[b * 3 + a];
    }
    for (int k = 0; k < 9; k++) h[k] /= sum;
}

int disp_run(pdegpu_ctx *ctx, Bump2 &b, float *Uout, const float *Il, const float *Ir, int nrows, int ncols, int C,
             const pdegpu_disp_sym_params &P)
{
    const bool dry = b.dry;
    struct Sz { int nr, nc; size_t n; };
    std::vector<Sz> L;
    L.push_back({nrows, ncols, (size_t)nrows * ncols});
    const int max_scales = P.max_scales > 0 ? P.max_scales : (1 << 30);
    while ((int)L.size() < max_scales) {
        Sz n = {(int)ceil(L.back().nr * P.scl_factor), (int)ceil(L.back().nc * P.scl_factor), 0};
        n.n = (size_t)n.nr * n.nc;
        L.push_back(n);
        if (n.nr <= 10 || n.nc <= 10) break;
    }
    const int S = (int)L.size();
    if (L.back().nr < 5 || L.back().nc < 5) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "disp_sym: coarsest level smaller than 5 pixels");
    const size_t n0 = L[0].n;
    // level k: It[k] = [left (C planes), right (C planes)]
    std::vector<float *> It(S);
    for (int k = 0; k < S; k++) It[k] = b.take(2 * L[k].n * C);
    float *tmp = b.take(2 * n0 * C), *tmp2 = b.take(2 * n0 * C);
    float *U = b.take(2 * n0), *dU = b.take(2 * n0), *Us = b.take(2 * n0);
    float *X = b.take(2 * n0), *Y = b.take(2 * n0);
    float *Ww = b.take(2 * n0 * C);                 // [right warped by U0 (C), left warped by U1 (C)]
    float *Sw = b.take(2 * n0);                     // [U1 sampled at X+U0, U0 sampled at X+U1]
    float *Udt = b.take(2 * n0), *Udx = b.take(2 * n0);
    float *D1[3], *D2[5];                           // Fst: Idt, Idx, Idy; Snd: Idxt, Idyt, Idxx, Idyy, Idxy  (2C planes: view 0, view 1)
    for (int k = 0; k < 3; k++) D1[k] = b.take(2 * n0 * C);
    for (int k = 0; k < 5; k++) D2[k] = b.take(2 * n0 * C);
    float *CuG = b.take(2 * n0), *DuG = b.take(2 * n0), *w[4];
    for (int k = 0; k < 4; k++) w[k] = b.take(2 * n0);          // wW, wN, wE, wS (DdiffWeights' order), both views
    if (dry) return PDEGPU_OK;

    double G[9];
    gaussian3(1.0, G);                                                                              // :81
    static const double pre[5] = {0.037659, 0.249724, 0.439911, 0.249724, 0.037659};
    static const double odx_f[5] = {-0.104550, -0.292315, 0.0, 0.292315, 0.104550};                 // O_dx flipped ('conv')
    // ---- pyramid (:86-103): next level from the UNSMOOTHED current one, then smooth the current one; the last level stays unsmoothed ----
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(It[0], Il, n0 * C * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(It[0] + n0 * C, Ir, n0 * C * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    if (P.uint8_input) RC(op_round_uint8(ctx, It[0], (long long)(2 * n0 * C)));
    for (int k = 1; k < S; k++) {
        const Sz &p = L[k - 1], &l = L[k];
        RC(imresize_2d(ctx, It[k], tmp, It[k - 1], p.nr, p.nc, l.nr, l.nc, P.scl_factor, P.scl_factor, 1, 2 * C, 0));
        if (P.uint8_input) RC(op_round_uint8(ctx, It[k], (long long)(2 * l.n * C)));
        RC(op_imfilter(ctx, tmp2, It[k - 1], p.nr, p.nc, 2 * C, (long long)p.n, (long long)p.n, G, 3, 3, 1, 1.0f));
        PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(It[k - 1], tmp2, 2 * p.n * C * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        if (P.uint8_input) RC(op_round_uint8(ctx, It[k - 1], (long long)(2 * p.n * C)));
    }

    const float up = (float)(1.0 / P.scl_factor);
    for (int s = S - 1; s >= 0; s--) {                                                              // :111
        const int nr = L[s].nr, nc = L[s].nc;
        const size_t n = L[s].n;
        const float *left = It[s], *right = It[s] + n * C;
        if (s == S - 1) PDEGPU_CUDA_OK(ctx, cudaMemsetAsync(U, 0, 2 * n * sizeof(float), ctx->stream));
        const double srDiff = 2.0 * pow(1.0 / P.scl_factor, -(double)s);                            // :127
        for (int fl = 0; fl < P.firstLoop; fl++) {                                                  // :133
            // X[v] = grid + U[v]; images: right warped by U0, left warped by U1 (:138-139)
            RC(op_warp_coords(ctx, X, Y, U, nullptr, nr, nc, 2, (long long)n));
            RC(op_bilin(ctx, Ww, right, X, Y, nr, nc, C, P.oob_value));
            RC(op_bilin(ctx, Ww + n * C, left, X + n, Y + n, nr, nc, C, P.oob_value));
            // disparities: Sw[0] = U1 at X+U0, Sw[1] = U0 at X+U1 (:144-145; U1w, U0w)
            RC(op_interp_rows(ctx, Sw, U + n, U, nr, nc));
            RC(op_interp_rows(ctx, Sw + n, U, U + n, nr, nc));
            // derivatives of (left, right warped) and (right, left warped) in one call each: 2C planes (:150-154)
            RC(op_fst(ctx, D1[0], D1[1], D1[2], It[s], Ww, nr, nc, 2 * C));
            RC(op_snd(ctx, D2[0], D2[1], D2[2], D2[3], D2[4], It[s], Ww, nr, nc, 2 * C));
            // Udt[v] = (U[v] + Sw[v])*0.5, Udx[v] = d/dx of Sw[v] (:159-165; Udy is computed by the driver and never used)
            RC(op_axpby(ctx, Udt, 1.0f, U, 1.0f, Sw, (long long)(2 * n)));
            RC(op_axpby(ctx, Udt, 0.5f, Udt, 0.0f, nullptr, (long long)(2 * n)));
            RC(op_imfilter(ctx, tmp, Sw, nr, nc, 2, (long long)n, (long long)n, pre, 5, 1, 1, 1.0f));
            RC(op_imfilter(ctx, Udx, tmp, nr, nc, 2, (long long)n, (long long)n, odx_f, 1, 5, 1, 1.0f));
            PDEGPU_CUDA_OK(ctx, cudaMemsetAsync(dU, 0, 2 * n * sizeof(float), ctx->stream));
            for (int sl = 0; sl < P.secondLoop; sl++) {                                             // :188
                for (int v = 0; v < 2; v++) {
                    pdegpu_disp_sym_terms t;
                    memset(&t, 0, sizeof t);
                    t.nrows = nr; t.ncols = nc; t.channels = C;
                    t.b1 = (float)P.b1; t.b2 = (float)P.b2; t.alpha = (float)P.alpha;
                    t.alpha_d = P.alpha; t.beta = P.beta; t.srdiff = srDiff;
                    const size_t o = (size_t)v * n * C;
                    t.d[0] = D1[0] + o; t.d[1] = D1[1] + o; t.d[2] = D2[0] + o; t.d[3] = D2[1] + o; t.d[4] = D2[2] + o; t.d[5] = D2[4] + o;
                    t.dU = dU + v * n; t.Udt = Udt + v * n; t.Udx = Udx + v * n;
                    t.CuG = CuG + v * n; t.DuG = DuG + v * n;
                    RC(op_disp_sym_terms(ctx, &t));                                                 // :172-180, 197-225
                    RC(op_axpby(ctx, Us + v * n, 1.0f, U + v * n, 1.0f, dU + v * n, (long long)n));
                    RC(op_ddiff(ctx, w[0] + v * n, w[1] + v * n, w[2] + v * n, w[3] + v * n, Us + v * n, nr, nc, 1, 0.00001f));   // :219-220
                }
                pdegpu_system sys;
                memset(&sys, 0, sizeof sys);
                sys.family = PDEGPU_DISP_LLIN4; sys.nrows = nr; sys.ncols = nc; sys.batch = 2; sys.batch_stride = (long long)n;
                sys.x[0] = dU; sys.x0[0] = U; sys.c[0] = CuG; sys.d[0] = DuG;
                sys.w[W_W] = w[0]; sys.w[W_N] = w[1]; sys.w[W_E] = w[2]; sys.w[W_S] = w[3];
                RC(pdegpu_dev_relax(ctx, &sys, P.iter, (float)P.omega, P.solver));                  // :227-247
            }
            RC(op_axpby(ctx, Us, 1.0f, U, 1.0f, dU, (long long)(2 * n)));                           // :255-256
            RC(op_medfilt3(ctx, U, Us, nr, nc, 2, (long long)n));
        }
        if (s > 0) {                                                                                // :265-267
            const int onr = L[s - 1].nr, onc = L[s - 1].nc;
            RC(op_axpby(ctx, Us, up, U, 0.0f, nullptr, (long long)(2 * n)));
            RC(imresize_2d(ctx, U, tmp, Us, nr, nc, onr, onc, (double)onr / nr, (double)onc / nc, 1, 2, 0));
        }
    }
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(Uout, U, 2 * n0 * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    return PDEGPU_OK;
}

}  // namespace

extern "C" void pdegpu_disp_sym_default_params(pdegpu_disp_sym_params *p)
{
    // defaults of DispEminND_llin_sym_2D.m:50-63
    p->alpha = 0.035; p->beta = 0.4; p->omega = 1.9; p->b1 = 0.25; p->b2 = 0.72; p->scl_factor = 0.75;
    p->firstLoop = 3; p->secondLoop = 4; p->iter = 4; p->solver = 2; p->max_scales = 0;
    p->uint8_input = 1;
    p->oob_value = nanf("");
}

extern "C" int pdegpu_dev_disp_sym_2d(pdegpu_ctx *ctx, float *U, const float *Il, const float *Ir,
        int nrows, int ncols, int channels, int batch, const pdegpu_disp_sym_params *params)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    if (!U || !Il || !Ir || !params) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_disp_sym_2d: null pointer");
    if (nrows < 8 || ncols < 8 || channels < 1 || batch < 1) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "pdegpu_dev_disp_sym_2d: bad shape");
    if (!(params->scl_factor > 0.1 && params->scl_factor < 1.0) || params->firstLoop < 1 || params->secondLoop < 1)
        return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_disp_sym_2d: bad parameters");
    PDEGPU_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    Bump2 dry = {nullptr, 0, true};
    int rc = disp_run(ctx, dry, U, Il, Ir, nrows, ncols, channels, *params);
    if (rc) return rc;
    // pairs of a batch run side by side on the context's lanes (one workspace each); a single pair on the context itself
    const int K = pdegpu_lane_count(ctx, batch);
    if ((rc = K > 1 ? pdegpu_lanes_prepare(ctx, K, dry.used, "pdegpu_dev_disp_sym_2d") : pdegpu_work_reserve(ctx, dry.used, "pdegpu_dev_disp_sym_2d"))) return rc;
    struct Args { pdegpu_ctx *ctx; float *U; const float *Il, *Ir; int nrows, ncols, channels, batch; pdegpu_disp_sym_params P; char *work; int id, K; };
    Args a;
    memset(&a, 0, sizeof a);                                   // (padding is part of the graph key)
    a.ctx = ctx; a.U = U; a.Il = Il; a.Ir = Ir; a.nrows = nrows; a.ncols = ncols; a.channels = channels; a.batch = batch;
    a.P = *params; a.work = K > 1 ? ctx->lanes[0]->work : ctx->work; a.id = 4; a.K = K;
    pdegpu_graph_body body = {[](void *p) -> int {
        Args &a = *static_cast<Args *>(p);
        const size_t np = (size_t)a.nrows * a.ncols;
        int rc;
        if (a.K > 1 && (rc = pdegpu_lanes_fork(a.ctx, a.K))) return rc;
        for (int bi = 0; bi < a.batch; bi++) {
            pdegpu_ctx *c = a.K > 1 ? a.ctx->lanes[bi % a.K] : a.ctx;
            Bump2 w = {c->work, 0, false};
            rc = disp_run(c, w, a.U + 2 * bi * np, a.Il + bi * np * a.channels, a.Ir + bi * np * a.channels, a.nrows, a.ncols, a.channels, a.P);
            if (rc) { if (c != a.ctx) memcpy(a.ctx->err, c->err, sizeof c->err); return rc; }
        }
        if (a.K > 1 && (rc = pdegpu_lanes_join(a.ctx, a.K))) return rc;
        return PDEGPU_OK;
    }, &a};
    return pdegpu_graph_run(ctx, &a, sizeof a, body);
}

extern "C" int pdegpu_disp_sym_2d(pdegpu_ctx *ctx, float *U, const float *Il, const float *Ir,
        int nrows, int ncols, int channels, int batch, const pdegpu_disp_sym_params *params)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    if (!U || !Il || !Ir || !params) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_disp_sym_2d: null pointer");
    PDEGPU_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const size_t nimg = (size_t)nrows * ncols * channels * batch * sizeof(float), nflow = (size_t)2 * nrows * ncols * batch * sizeof(float);
    pdegpu_arena_reset(ctx);
    int rc = pdegpu_arena_reserve(ctx, 2 * nimg + nflow + 4096);
    if (rc) return rc;
    float *d0 = (float *)pdegpu_arena_alloc(ctx, nimg), *d1 = (float *)pdegpu_arena_alloc(ctx, nimg);
    float *dU = (float *)pdegpu_arena_alloc(ctx, nflow);
    if (!d0 || !d1 || !dU) return pdegpu_set_error(ctx, PDEGPU_ERR_NOMEM, "pdegpu_disp_sym_2d: arena exhausted");
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(d0, Il, nimg, cudaMemcpyHostToDevice, ctx->stream));
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(d1, Ir, nimg, cudaMemcpyHostToDevice, ctx->stream));
    rc = pdegpu_dev_disp_sym_2d(ctx, dU, d0, d1, nrows, ncols, channels, batch, params);
    if (rc) return rc;
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(U, dU, nflow, cudaMemcpyDeviceToHost, ctx->stream));
    PDEGPU_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return PDEGPU_OK;
}
