// pipeline_fmg.cu -- device-resident restatement of the reference's early-linearisation full-multigrid flow driver
// (BASELINE.json configs[2]): matlab/optical_flow/FlowEminNDFASFMG_elin_2D_v10.m.
//
// Structure kept step for step: Gaussian + lpf pyramid (:97-118), derivative stacks and constant terms per level
// (:123-149), FMG loop coarse to fine (:158-183) with one FAS V- or W-cycle per level (FAS_CYCLE :193-273) around the
// smoother (:367-464: robust weight gd, OPdiffWeights, Oflow_sor_elin4_2d; residuals through the iter = 0 call).
// Every step is one libpdegpu kernel on the context's stream; nothing returns to the host between the upload of the
// frames and the download of the flow. MEX calls of the reference (Oflow_sor_elin4_2d, Oflow_lhs_elin4_2d) run the
// same kernels as the gateways, Matlab steps use driver_ops.cu. Pairs of a batch run one after the other on the
// stream (the residual / LHS kernels take channel stacks for one problem).
#include "pdegpu_internal.cuh"
#include <math.h>
#include <vector>

namespace {

// bump allocator with stack discipline (the FAS recursion releases what a level took) and a high-water mark
struct Stack {
    char *base; size_t used, peak; bool dry;
    float *take(size_t nfloats)
    {
        const size_t bytes = (nfloats * sizeof(float) + 255) & ~(size_t)255;
        float *p = dry ? nullptr : (float *)(base + used);
        used += bytes;
        if (used > peak) peak = used;
        return p;
    }
};

struct Lvl {
    int nr, nc;
    size_t n;
    float *It[2];
    float *der[8];    // Idt, Idx, Idy, Idxt, Idyt, Idxx, Idyy, Idxy   (C planes)
    float *coef[5];   // M, Cu, Cv, Du, Dv                             (C planes)
};

#define RC(call) do { int rc__ = (call); if (rc__) return rc__; } while (0)

struct Fmg {
    pdegpu_ctx *ctx;
    Stack &b;
    const pdegpu_flow_fmg_params &P;
    int C, S;
    std::vector<Lvl> L;
    float *T[5], *TC[5], *w[4], *tmp, *tmp2;       // summed terms, per-channel terms, weights, scratch (finest-level size)

    int relax_system(pdegpu_system &sys, const Lvl &l, float *U, float *V, float *const t[5])
    {
        memset(&sys, 0, sizeof sys);
        sys.family = PDEGPU_FLOW_ELIN4; sys.nrows = l.nr; sys.ncols = l.nc; sys.batch = 1; sys.batch_stride = (long long)l.n;
        sys.x[0] = U; sys.x[1] = V;
        sys.m = t[0]; sys.c[0] = t[1]; sys.c[1] = t[2]; sys.d[0] = t[3]; sys.d[1] = t[4];
        sys.w[W_W] = w[0]; sys.w[W_N] = w[1]; sys.w[W_S] = w[2]; sys.w[W_E] = w[3];      // OPdiffWeights returns [wW wN wS wE]
        return PDEGPU_OK;
    }

    int terms(const Lvl &l, const float *const coef[5], const float *U, const float *V, int summed, float *gd, float *const out[5])
    {
        pdegpu_elin_terms t;
        memset(&t, 0, sizeof t);
        t.nrows = l.nr; t.ncols = l.nc; t.channels = C; t.summed = summed;
        t.b1 = (float)P.b1; t.b2 = (float)P.b2; t.alpha = (float)P.alpha;
        for (int k = 0; k < 8; k++) t.der[k] = l.der[k];
        for (int k = 0; k < 5; k++) { t.coef[k] = coef[k]; t.out[k] = out[k]; }
        t.U = U; t.V = V; t.gd = gd;
        return op_elin_terms(ctx, &t);
    }

    // smooth(), FlowEminNDFASFMG_elin_2D_v10.m:367-464. Cu, Cv: C planes (the level's own terms or the FAS right-hand side).
    int smooth(int s, float *U, float *V, const float *Cu, const float *Cv, float *RU, float *RV)
    {
        const Lvl &l = L[s];
        const float *coef[5] = {l.coef[0], Cu, Cv, l.coef[3], l.coef[4]};
        pdegpu_system sys;
        for (int fl = 0; fl < P.firstLoop; fl++) {
            RC(terms(l, coef, U, V, 1, nullptr, T));                                           // :375-396
            RC(op_opdiff(ctx, w[0], w[1], w[2], w[3], U, V, l.nr, l.nc, 1, (long long)l.n));    // :391
            relax_system(sys, l, U, V, T);
            RC(pdegpu_dev_relax(ctx, &sys, P.iter, (float)P.omega, P.solver));                  // :399-413
        }
        if (RU) {                                                                              // :421-460: per-channel residuals
            RC(terms(l, coef, U, V, 0, nullptr, TC));
            RC(op_opdiff(ctx, w[0], w[1], w[2], w[3], U, V, l.nr, l.nc, 1, (long long)l.n));
            relax_system(sys, l, U, V, TC);
            RC(op_residual(ctx, &sys, C, RU, RV, false));
        }
        return PDEGPU_OK;
    }

    // FAS_CYCLE, :193-273
    int cycle(int s, float *U, float *V, const float *Cu, const float *Cv)
    {
        const Lvl &l = L[s];
        const size_t mark = b.used;
        const bool dry = b.dry;
        bool coarse = false;
        if (s < S - 1) {
            const Lvl &c = L[s + 1];
            const float scl = (float)P.scl_factor, up = (float)(1.0 / P.scl_factor);
            static const double fw[9] = {1 / 16.0, 2 / 16.0, 1 / 16.0, 2 / 16.0, 4 / 16.0, 2 / 16.0, 1 / 16.0, 2 / 16.0, 1 / 16.0};
            float *RU = b.take(l.n * C), *RV = b.take(l.n * C);
            float *RUr = b.take(c.n * C), *RVr = b.take(c.n * C), *Ur = b.take(c.n), *Vr = b.take(c.n);
            float *gd = b.take(c.n * C), *Au = b.take(c.n * C), *Av = b.take(c.n * C), *fu = b.take(c.n * C), *fv = b.take(c.n * C);
            float *Uc = b.take(c.n), *Vc = b.take(c.n), *dlt = b.take(c.n), *res = b.take(l.n);
            for (int ci = 0; ci < P.cycle_index; ci++) {
                if (!dry) {
                    RC(smooth(s, U, V, Cu, Cv, RU, RV));                                                        // :207 pre-smoothing
                    RC(op_imfilter(ctx, RUr, RU, l.nr, l.nc, C, (long long)l.n, (long long)c.n, fw, 3, 3, 2, scl));   // :212-213
                    RC(op_imfilter(ctx, RVr, RV, l.nr, l.nc, C, (long long)l.n, (long long)c.n, fw, 3, 3, 2, scl));
                    RC(op_imfilter(ctx, Ur, U, l.nr, l.nc, 1, (long long)l.n, (long long)c.n, fw, 3, 3, 2, scl));      // :216-217
                    RC(op_imfilter(ctx, Vr, V, l.nr, l.nc, 1, (long long)l.n, (long long)c.n, fw, 3, 3, 2, scl));
                    const float *cc[5] = {c.coef[0], c.coef[0], c.coef[0], c.coef[3], c.coef[4]};
                    RC(terms(c, cc, Ur, Vr, 0, gd, TC));                                                        // :225-237 (TC[1], TC[2] unused)
                    RC(op_opdiff(ctx, w[0], w[1], w[2], w[3], Ur, Vr, c.nr, c.nc, 1, (long long)c.n));           // :234
                    pdegpu_system sys;
                    relax_system(sys, c, Ur, Vr, TC);
                    RC(op_residual(ctx, &sys, C, Au, Av, true));                                                // :239-248 Oflow_lhs_elin4_2d
                    RC(op_fas_rhs(ctx, fu, RUr, Au, gd, (long long)(c.n * C)));                                 // :250-251
                    RC(op_fas_rhs(ctx, fv, RVr, Av, gd, (long long)(c.n * C)));
                    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(Uc, Ur, c.n * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
                    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(Vc, Vr, c.n * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
                }
                RC(cycle(s + 1, Uc, Vc, fu, fv));                                                               // :253
                if (!dry) {
                    // U = U + imresize((Uc-Ures)*(1/scl_factor), size(U), 'bilinear')   :256-257
                    float *xs[2] = {U, V}, *cs[2] = {Uc, Vc}, *rs[2] = {Ur, Vr};
                    for (int q = 0; q < 2; q++) {
                        RC(op_axpby(ctx, dlt, 1.0f, cs[q], -1.0f, rs[q], (long long)c.n));
                        RC(op_axpby(ctx, dlt, up, dlt, 0.0f, nullptr, (long long)c.n));
                        RC(imresize_2d(ctx, res, tmp, dlt, c.nr, c.nc, l.nr, l.nc, (double)l.nr / c.nr, (double)l.nc / c.nc, 1, 1, 0));
                        RC(op_axpby(ctx, xs[q], 1.0f, xs[q], 1.0f, res, (long long)l.n));
                    }
                }
                coarse = true;
            }
        } else if (!dry) {
            RC(smooth(s, U, V, Cu, Cv, nullptr, nullptr));                                                      // :261-263
        }
        if (coarse && !dry) RC(smooth(s, U, V, Cu, Cv, nullptr, nullptr));                                      // :269 post-smoothing
        b.used = mark;
        return PDEGPU_OK;
    }
};

// Derivative stacks of one level, shared by the FMG driver (FlowEminNDFASFMG_elin_2D_v10.m:123-141: frames 0..255,
// Ist and Idt divided by 255, temporal-derivative kernels O_dx/255) and the Horn-Schunck driver
// (FlowEminHS_elin_2D_v10.m:139-154: frames already scaled, div = 1). der = Idt, Idx, Idy, Idxt, Idyt, Idxx, Idyy, Idxy.
int derivative_stack(pdegpu_ctx *ctx, const Lvl &l, int C, float div, float *tmp, float *tmp2, float *Ist)
{
    static const double pre[5] = {0.037659, 0.249724, 0.439911, 0.249724, 0.037659};
    // 'conv' = correlation with the flipped kernel
    static const double odx_f[5] = {-0.104550, -0.292315, 0.0, 0.292315, 0.104550};                 // O_dx flipped
    static const double oxx[5] = {0.232905, 0.002668, -0.471147, 0.002668, 0.232905};
    double odxs_f[5];
    for (int k = 0; k < 5; k++) odxs_f[k] = odx_f[k] / (double)div;                                 // O_dx_scl flipped
    const long long n = (long long)l.n;
    auto filt = [&](float *out, const float *in, const double *h, int kr, int kc) {
        return op_imfilter(ctx, out, in, l.nr, l.nc, C, n, n, h, kr, kc, 1, 1.0f);
    };
    RC(op_fmg_prescale(ctx, Ist, l.der[0], l.It[0], l.It[1], n * C, div));                          // Ist, Idt
    RC(filt(tmp, Ist, pre, 5, 1));                                                                  // prefilter_spa'
    RC(filt(l.der[1], tmp, odx_f, 1, 5));                                                           // Idx
    RC(filt(l.der[5], tmp, oxx, 1, 5));                                                             // Idxx
    RC(filt(tmp, Ist, pre, 1, 5));                                                                  // prefilter_spa
    RC(filt(l.der[2], tmp, odx_f, 5, 1));                                                           // Idy
    RC(filt(l.der[6], tmp, oxx, 5, 1));                                                             // Idyy
    RC(filt(tmp, Ist, odx_f, 1, 5));
    RC(filt(l.der[7], tmp, odx_f, 5, 1));                                                           // Idxy
    // Idxt = Idxt0 - Idxt1, Idyt = Idyt0 - Idyt1
    RC(filt(tmp, l.It[0], pre, 5, 1)); RC(filt(l.der[3], tmp, odxs_f, 1, 5));
    RC(filt(tmp, l.It[1], pre, 5, 1)); RC(filt(tmp2, tmp, odxs_f, 1, 5));
    RC(op_axpby(ctx, l.der[3], 1.0f, l.der[3], -1.0f, tmp2, n * C));
    RC(filt(tmp, l.It[0], pre, 1, 5)); RC(filt(l.der[4], tmp, odxs_f, 5, 1));
    RC(filt(tmp, l.It[1], pre, 1, 5)); RC(filt(tmp2, tmp, odxs_f, 5, 1));
    RC(op_axpby(ctx, l.der[4], 1.0f, l.der[4], -1.0f, tmp2, n * C));
    return PDEGPU_OK;
}

// fspecial('gaussian', [5 5], sigma), column-major
void gaussian5x5(double sigma, double *h)
{
    double mx = 0, sum = 0;
    for (int bb = 0; bb < 5; bb++) for (int a = 0; a < 5; a++) {
        const double x = bb - 2, y = a - 2;
        h[bb * 5 + a] = exp(-(x * x + y * y) / (2.0 * sigma * sigma));
        mx = fmax(mx, h[bb * 5 + a]);
    }
    for (int k = 0; k < 25; k++) { if (h[k] < 2.220446049250313e-16 * mx) h[k] = 0; sum += h[k]; }
    for (int k = 0; k < 25; k++) h[k] /= sum;
}

// one pair; with b.dry only the workspace is measured
int fmg_run(pdegpu_ctx *ctx, Stack &b, float *Uout, float *Vout, const float *I0, const float *I1, int nrows, int ncols, int C,
            const pdegpu_flow_fmg_params &P)
{
    const bool dry = b.dry;
    Fmg f = {ctx, b, P, C, 0, {}, {}, {}, {}, nullptr, nullptr};
    // ---- pyramid sizes (:106-118): (1:2:end) decimation until a dimension is <= 10 ----
    {
        Lvl l; memset(&l, 0, sizeof l);
        l.nr = nrows; l.nc = ncols; l.n = (size_t)nrows * ncols;
        f.L.push_back(l);
        const int max_scales = P.max_scales > 0 ? P.max_scales : (1 << 30);
        while ((int)f.L.size() < max_scales) {
            Lvl n; memset(&n, 0, sizeof n);
            n.nr = (f.L.back().nr + 1) / 2; n.nc = (f.L.back().nc + 1) / 2; n.n = (size_t)n.nr * n.nc;
            f.L.push_back(n);
            if (n.nr <= 10 || n.nc <= 10) break;
        }
        f.S = (int)f.L.size();
        // the 5-tap derivative stack of every level needs 5 pixels per dimension (imageDerivatives.c:66-211)
        if (f.L.back().nr < 5 || f.L.back().nc < 5) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "flow_fmg: coarsest level smaller than 5 pixels (lower max_scales)");
    }
    const int S = f.S;
    const size_t n0 = f.L[0].n;
    for (int s = 0; s < S; s++) {
        Lvl &l = f.L[s];
        for (int k = 0; k < 2; k++) l.It[k] = b.take(l.n * C);
        for (int k = 0; k < 8; k++) l.der[k] = b.take(l.n * C);
        for (int k = 0; k < 5; k++) l.coef[k] = b.take(l.n * C);
    }
    for (int k = 0; k < 5; k++) { f.T[k] = b.take(n0); f.TC[k] = b.take(n0 * C); }
    for (int k = 0; k < 4; k++) f.w[k] = b.take(n0);
    f.tmp = b.take(n0 * C); f.tmp2 = b.take(n0 * C);
    float *Ist = b.take(n0 * C), *U = b.take(n0), *V = b.take(n0), *Us = b.take(n0);

    if (!dry) {
        double G[25];
        gaussian5x5(1.0, G);                                                                        // :97
        static const double lpf[5] = {1 / 16.0, 4 / 16.0, 6 / 16.0, 4 / 16.0, 1 / 16.0};            // :98
        const float *Iin[2] = {I0, I1};
        for (int q = 0; q < 2; q++) {
            RC(op_imfilter(ctx, f.L[0].It[q], Iin[q], nrows, ncols, C, (long long)n0, (long long)n0, G, 5, 5, 1, 1.0f));   // :103-104
            for (int s = 1; s < S; s++) {                                                           // :107-110
                const Lvl &p = f.L[s - 1], &l = f.L[s];
                RC(op_imfilter(ctx, f.tmp, p.It[q], p.nr, p.nc, C, (long long)p.n, (long long)p.n, lpf, 1, 5, 1, 1.0f));
                RC(op_imfilter(ctx, l.It[q], f.tmp, p.nr, p.nc, C, (long long)p.n, (long long)l.n, lpf, 5, 1, 2, 1.0f));
            }
        }
        for (int s = 0; s < S; s++) {                                                               // :123-149
            const Lvl &l = f.L[s];
            RC(derivative_stack(ctx, l, C, 255.0f, f.tmp, f.tmp2, Ist));
            RC(op_fmg_terms(ctx, l.coef, l.der, (float)P.b1, (float)P.b2, (long long)l.n * C));     // :143-149
        }
    }

    // ---- full multigrid: coarse to fine, one FAS cycle per level (:158-183) ----
    const float up = (float)(1.0 / P.scl_factor);
    for (int s = S - 1; s >= 0; s--) {
        const Lvl &l = f.L[s];
        if (s == S - 1 && !dry) {
            PDEGPU_CUDA_OK(ctx, cudaMemsetAsync(U, 0, l.n * sizeof(float), ctx->stream));
            PDEGPU_CUDA_OK(ctx, cudaMemsetAsync(V, 0, l.n * sizeof(float), ctx->stream));
        }
        RC(f.cycle(s, U, V, l.coef[1], l.coef[2]));                                                 // :174
        if (s > 0 && !dry) {
            // U = imresize(U.*(1/scl_factor), isizes{scl-1}(1:2))   default method: bicubic (:180-181)
            const Lvl &o = f.L[s - 1];
            float *xs[2] = {U, V};
            for (int q = 0; q < 2; q++) {
                RC(op_axpby(ctx, Us, up, xs[q], 0.0f, nullptr, (long long)l.n));
                RC(imresize_2d(ctx, xs[q], f.tmp, Us, l.nr, l.nc, o.nr, o.nc, (double)o.nr / l.nr, (double)o.nc / l.nc, 1, 1, 1));
            }
        }
    }
    if (!dry) {
        PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(Uout, U, n0 * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(Vout, V, n0 * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return PDEGPU_OK;
}

}  // namespace

extern "C" void pdegpu_flow_fmg_default_params(pdegpu_flow_fmg_params *p)
{
    // defaults of FlowEminNDFASFMG_elin_2D_v10.m:53-66
    p->alpha = 0.035; p->omega = 1.9; p->b1 = 0.03; p->b2 = 0.97; p->scl_factor = 0.5;
    p->firstLoop = 4; p->iter = 4; p->solver = 2; p->cycle_index = 1; p->max_scales = 0;
}

extern "C" int pdegpu_dev_flow_fmg_2d(pdegpu_ctx *ctx, float *U, float *V, const float *I0, const float *I1,
        int nrows, int ncols, int channels, int batch, const pdegpu_flow_fmg_params *params)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    if (!U || !V || !I0 || !I1 || !params) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_flow_fmg_2d: null pointer");
    if (nrows < 8 || ncols < 8 || channels < 1 || batch < 1) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "pdegpu_dev_flow_fmg_2d: bad shape");
    if (!(params->scl_factor > 0.0) || params->firstLoop < 1 || params->cycle_index < 1 || params->cycle_index > 2)
        return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_flow_fmg_2d: bad parameters");
    // The pyramid, the restriction and the level sizes are the driver's 1:2:end decimation (FlowEminNDFASFMG_elin_2D_v10.m:106-118,
    // 212-217); scl_factor only scales the restricted quantities and the prolongated correction, so any other value would
    // describe an inconsistent multigrid (the reference fails on a size mismatch there).
    if (fabs(params->scl_factor - 0.5) > 1e-12)
        return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_flow_fmg_2d: scl_factor must be 0.5 (the pyramid is decimated 1:2)");
    PDEGPU_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    Stack dry = {nullptr, 0, 0, true};
    int rc = fmg_run(ctx, dry, U, V, I0, I1, nrows, ncols, channels, *params);
    if (rc) return rc;
    // pairs of a batch run side by side on the context's lanes (one workspace each); a single pair on the context itself
    const int K = pdegpu_lane_count(ctx, batch, ctx->sweep_order == PDEGPU_ORDER_REFERENCE || (ctx->sweep_order == PDEGPU_ORDER_AUTO && params->solver == 2));
    if ((rc = K > 1 ? pdegpu_lanes_prepare(ctx, K, dry.peak, "pdegpu_dev_flow_fmg_2d") : pdegpu_work_reserve(ctx, dry.peak, "pdegpu_dev_flow_fmg_2d"))) return rc;
    struct Args { pdegpu_ctx *ctx; float *U, *V; const float *I0, *I1; int nrows, ncols, channels, batch; pdegpu_flow_fmg_params P; char *work; int id, K; };
    Args a;
    memset(&a, 0, sizeof a);                                   // (padding is part of the graph key)
    a.ctx = ctx; a.U = U; a.V = V; a.I0 = I0; a.I1 = I1; a.nrows = nrows; a.ncols = ncols; a.channels = channels; a.batch = batch;
    a.P = *params; a.work = K > 1 ? ctx->lanes[0]->work : ctx->work; a.id = 2; a.K = K;
    pdegpu_graph_body body = {[](void *p) -> int {
        Args &a = *static_cast<Args *>(p);
        const size_t np = (size_t)a.nrows * a.ncols;
        if (a.K > 1) RC(pdegpu_lanes_fork(a.ctx, a.K));
        for (int bi = 0; bi < a.batch; bi++) {
            pdegpu_ctx *c = a.K > 1 ? a.ctx->lanes[bi % a.K] : a.ctx;
            Stack w = {c->work, 0, 0, false};
            const int rc = fmg_run(c, w, a.U + bi * np, a.V + bi * np, a.I0 + bi * np * a.channels, a.I1 + bi * np * a.channels,
                                   a.nrows, a.ncols, a.channels, a.P);
            if (rc) { if (c != a.ctx) memcpy(a.ctx->err, c->err, sizeof c->err); return rc; }
        }
        if (a.K > 1) RC(pdegpu_lanes_join(a.ctx, a.K));
        return PDEGPU_OK;
    }, &a};
    return pdegpu_graph_run(ctx, &a, sizeof a, body);
}

extern "C" int pdegpu_flow_fmg_2d(pdegpu_ctx *ctx, float *U, float *V, const float *I0, const float *I1,
        int nrows, int ncols, int channels, int batch, const pdegpu_flow_fmg_params *params)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    if (!U || !V || !I0 || !I1 || !params) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_flow_fmg_2d: null pointer");
    PDEGPU_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const size_t nimg = (size_t)nrows * ncols * channels * batch * sizeof(float), nflow = (size_t)nrows * ncols * batch * sizeof(float);
    pdegpu_arena_reset(ctx);
    int rc = pdegpu_arena_reserve(ctx, 2 * nimg + 2 * nflow + 4096);
    if (rc) return rc;
    float *d0 = (float *)pdegpu_arena_alloc(ctx, nimg), *d1 = (float *)pdegpu_arena_alloc(ctx, nimg);
    float *dU = (float *)pdegpu_arena_alloc(ctx, nflow), *dV = (float *)pdegpu_arena_alloc(ctx, nflow);
    if (!d0 || !d1 || !dU || !dV) return pdegpu_set_error(ctx, PDEGPU_ERR_NOMEM, "pdegpu_flow_fmg_2d: arena exhausted");
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(d0, I0, nimg, cudaMemcpyHostToDevice, ctx->stream));
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(d1, I1, nimg, cudaMemcpyHostToDevice, ctx->stream));
    rc = pdegpu_dev_flow_fmg_2d(ctx, dU, dV, d0, d1, nrows, ncols, channels, batch, params);
    if (rc) return rc;
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(U, dU, nflow, cudaMemcpyDeviceToHost, ctx->stream));
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(V, dV, nflow, cudaMemcpyDeviceToHost, ctx->stream));
    PDEGPU_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return PDEGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// Horn-Schunck, matlab/optical_flow/FlowEminHS_elin_2D_v10.m (BASELINE configs[0]): bilinear x0.75 pyramid with 5x5
// Gaussian smoothing (:96-115), per level the derivative stacks and quadratic terms (:139-172), ONE linear solve
// Oflow_sor_elin4_2d with constant weights alpha*channels (:127,174-188), median + bicubic up-sampling (:193-196).
// ---------------------------------------------------------------------------------------------
namespace {

int hs_run(pdegpu_ctx *ctx, Stack &b, float *Uout, float *Vout, const float *I0, const float *I1, int nrows, int ncols, int C,
           const pdegpu_flow_hs_params &P)
{
    const bool dry = b.dry;
    std::vector<Lvl> L;
    {
        Lvl l; memset(&l, 0, sizeof l);
        l.nr = nrows; l.nc = ncols; l.n = (size_t)nrows * ncols;
        L.push_back(l);
        const int max_scales = P.max_scales > 0 ? P.max_scales : (1 << 30);
        while ((int)L.size() < max_scales) {
            Lvl n; memset(&n, 0, sizeof n);
            n.nr = (int)ceil(L.back().nr * P.scl_factor); n.nc = (int)ceil(L.back().nc * P.scl_factor); n.n = (size_t)n.nr * n.nc;
            L.push_back(n);
            if (n.nr <= 20 || n.nc <= 20) break;
        }
    }
    const int S = (int)L.size();
    const bool last_smoothed = L[S - 1].nr <= 20 || L[S - 1].nc <= 20;       // the size test smooths the last level (:108-113)
    if (L.back().nr < 5 || L.back().nc < 5) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "flow_hs: coarsest level smaller than 5 pixels");
    const size_t n0 = L[0].n;
    for (int s = 0; s < S; s++) for (int k = 0; k < 2; k++) L[s].It[k] = b.take(L[s].n * C);
    Lvl w; memset(&w, 0, sizeof w);                                          // work arrays, finest-level size, reused per level
    for (int k = 0; k < 8; k++) w.der[k] = b.take(n0 * C);
    for (int k = 0; k < 5; k++) w.coef[k] = b.take(n0 * C);
    float *T[5], *W = b.take(n0), *tmp = b.take(n0 * C), *tmp2 = b.take(n0 * C), *Ist = b.take(n0 * C);
    for (int k = 0; k < 5; k++) T[k] = b.take(n0);
    float *U = b.take(n0), *V = b.take(n0), *Us = b.take(n0), *Um = b.take(n0);
    if (dry) return PDEGPU_OK;

    double G[25];
    gaussian5x5(1.25, G);                                                                           // :88
    const float *Iin[2] = {I0, I1};
    for (int q = 0; q < 2; q++) {
        RC(op_axpby_div(ctx, L[0].It[q], Iin[q], 255.0f, (long long)(n0 * C)));                    // Iin = single(Iin)./255  (:66)
        for (int s = 1; s < S; s++) {                                                               // next level from the UNSMOOTHED one (:97-104)
            const Lvl &p = L[s - 1], &l = L[s];
            RC(imresize_2d(ctx, l.It[q], tmp, p.It[q], p.nr, p.nc, l.nr, l.nc, P.scl_factor, P.scl_factor, 1, C, 0));
            RC(op_imfilter(ctx, tmp2, p.It[q], p.nr, p.nc, C, (long long)p.n, (long long)p.n, G, 5, 5, 1, 1.0f));
            PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(p.It[q], tmp2, p.n * C * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        }
        if (last_smoothed && S > 1) {
            const Lvl &l = L[S - 1];
            RC(op_imfilter(ctx, tmp2, l.It[q], l.nr, l.nc, C, (long long)l.n, (long long)l.n, G, 5, 5, 1, 1.0f));
            PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(l.It[q], tmp2, l.n * C * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        }
    }
    const float up = (float)(1.0 / P.scl_factor);
    for (int s = S - 1; s >= 0; s--) {                                                              // :123
        Lvl l = w;
        l.nr = L[s].nr; l.nc = L[s].nc; l.n = L[s].n; l.It[0] = L[s].It[0]; l.It[1] = L[s].It[1];
        const long long n = (long long)l.n;
        if (s == S - 1) {
            PDEGPU_CUDA_OK(ctx, cudaMemsetAsync(U, 0, l.n * sizeof(float), ctx->stream));
            PDEGPU_CUDA_OK(ctx, cudaMemsetAsync(V, 0, l.n * sizeof(float), ctx->stream));
        }
        RC(op_fill(ctx, W, (float)(P.alpha * C), n));                                               // W = alpha*channels*ones (:127)
        RC(derivative_stack(ctx, l, C, 1.0f, tmp, tmp2, Ist));                                      // :139-154
        RC(op_fmg_terms(ctx, l.coef, l.der, (float)P.b1, (float)P.b2, n * C));                      // :159-163
        for (int k = 0; k < 5; k++) RC(op_channel_sum(ctx, T[k], l.coef[k], C, n));                 // :168-172
        pdegpu_system sys;
        memset(&sys, 0, sizeof sys);
        sys.family = PDEGPU_FLOW_ELIN4; sys.nrows = l.nr; sys.ncols = l.nc; sys.batch = 1; sys.batch_stride = n;
        sys.x[0] = U; sys.x[1] = V;
        sys.m = T[0]; sys.c[0] = T[1]; sys.c[1] = T[2]; sys.d[0] = T[3]; sys.d[1] = T[4];
        for (int k = 0; k < 4; k++) sys.w[k] = W;
        RC(pdegpu_dev_relax(ctx, &sys, P.iter, (float)P.omega, P.solver));                          // :174-188
        if (s > 0) {                                                                                // :193-196
            const Lvl &o = L[s - 1];
            float *xs[2] = {U, V};
            for (int q = 0; q < 2; q++) {
                RC(op_axpby(ctx, Us, up, xs[q], 0.0f, nullptr, n));
                RC(op_medfilt3(ctx, Um, Us, l.nr, l.nc, 1, n));
                RC(imresize_2d(ctx, xs[q], tmp, Um, l.nr, l.nc, o.nr, o.nc, (double)o.nr / l.nr, (double)o.nc / l.nc, 1, 1, 1));
            }
        }
    }
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(Uout, U, n0 * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(Vout, V, n0 * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    return PDEGPU_OK;
}

}  // namespace

extern "C" void pdegpu_flow_hs_default_params(pdegpu_flow_hs_params *p)
{
    // defaults of FlowEminHS_elin_2D_v10.m:48-58
    p->alpha = 0.2; p->omega = 1.9; p->b1 = 0.25; p->b2 = 0.75; p->scl_factor = 0.75;
    p->iter = 20; p->solver = 2; p->max_scales = 0;
}

extern "C" int pdegpu_dev_flow_hs_2d(pdegpu_ctx *ctx, float *U, float *V, const float *I0, const float *I1,
        int nrows, int ncols, int channels, int batch, const pdegpu_flow_hs_params *params)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    if (!U || !V || !I0 || !I1 || !params) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_flow_hs_2d: null pointer");
    if (nrows < 8 || ncols < 8 || channels < 1 || batch < 1) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "pdegpu_dev_flow_hs_2d: bad shape");
    if (!(params->scl_factor > 0.1 && params->scl_factor < 1.0)) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_flow_hs_2d: bad parameters");
    PDEGPU_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    Stack dry = {nullptr, 0, 0, true};
    int rc = hs_run(ctx, dry, U, V, I0, I1, nrows, ncols, channels, *params);
    if (rc) return rc;
    // pairs of a batch run side by side on the context's lanes (one workspace each); a single pair on the context itself
    const int K = pdegpu_lane_count(ctx, batch, ctx->sweep_order == PDEGPU_ORDER_REFERENCE || (ctx->sweep_order == PDEGPU_ORDER_AUTO && params->solver == 2));
    if ((rc = K > 1 ? pdegpu_lanes_prepare(ctx, K, dry.peak, "pdegpu_dev_flow_hs_2d") : pdegpu_work_reserve(ctx, dry.peak, "pdegpu_dev_flow_hs_2d"))) return rc;
    struct Args { pdegpu_ctx *ctx; float *U, *V; const float *I0, *I1; int nrows, ncols, channels, batch; pdegpu_flow_hs_params P; char *work; int id, K; };
    Args a;
    memset(&a, 0, sizeof a);                                   // (padding is part of the graph key)
    a.ctx = ctx; a.U = U; a.V = V; a.I0 = I0; a.I1 = I1; a.nrows = nrows; a.ncols = ncols; a.channels = channels; a.batch = batch;
    a.P = *params; a.work = K > 1 ? ctx->lanes[0]->work : ctx->work; a.id = 3; a.K = K;
    pdegpu_graph_body body = {[](void *p) -> int {
        Args &a = *static_cast<Args *>(p);
        const size_t np = (size_t)a.nrows * a.ncols;
        if (a.K > 1) RC(pdegpu_lanes_fork(a.ctx, a.K));
        for (int bi = 0; bi < a.batch; bi++) {
            pdegpu_ctx *c = a.K > 1 ? a.ctx->lanes[bi % a.K] : a.ctx;
            Stack w = {c->work, 0, 0, false};
            const int rc = hs_run(c, w, a.U + bi * np, a.V + bi * np, a.I0 + bi * np * a.channels, a.I1 + bi * np * a.channels,
                                   a.nrows, a.ncols, a.channels, a.P);
            if (rc) { if (c != a.ctx) memcpy(a.ctx->err, c->err, sizeof c->err); return rc; }
        }
        if (a.K > 1) RC(pdegpu_lanes_join(a.ctx, a.K));
        return PDEGPU_OK;
    }, &a};
    return pdegpu_graph_run(ctx, &a, sizeof a, body);
}

extern "C" int pdegpu_flow_hs_2d(pdegpu_ctx *ctx, float *U, float *V, const float *I0, const float *I1,
        int nrows, int ncols, int channels, int batch, const pdegpu_flow_hs_params *params)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    if (!U || !V || !I0 || !I1 || !params) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_flow_hs_2d: null pointer");
    PDEGPU_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const size_t nimg = (size_t)nrows * ncols * channels * batch * sizeof(float), nflow = (size_t)nrows * ncols * batch * sizeof(float);
    pdegpu_arena_reset(ctx);
    int rc = pdegpu_arena_reserve(ctx, 2 * nimg + 2 * nflow + 4096);
    if (rc) return rc;
    float *d0 = (float *)pdegpu_arena_alloc(ctx, nimg), *d1 = (float *)pdegpu_arena_alloc(ctx, nimg);
    float *dU = (float *)pdegpu_arena_alloc(ctx, nflow), *dV = (float *)pdegpu_arena_alloc(ctx, nflow);
    if (!d0 || !d1 || !dU || !dV) return pdegpu_set_error(ctx, PDEGPU_ERR_NOMEM, "pdegpu_flow_hs_2d: arena exhausted");
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(d0, I0, nimg, cudaMemcpyHostToDevice, ctx->stream));
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(d1, I1, nimg, cudaMemcpyHostToDevice, ctx->stream));
    rc = pdegpu_dev_flow_hs_2d(ctx, dU, dV, d0, d1, nrows, ncols, channels, batch, params);
    if (rc) return rc;
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(U, dU, nflow, cudaMemcpyDeviceToHost, ctx->stream));
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(V, dV, nflow, cudaMemcpyDeviceToHost, ctx->stream));
    PDEGPU_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return PDEGPU_OK;
}
