// api.cu -- the extern "C" surface of libpdegpu (see include/pdegpu.h).
//   pdegpu_dev_*  : validation + dispatch to the kernel launchers (asynchronous).
//   pdegpu_<name> : one MEX gateway call: stage host arrays into the context's device arena,
//                   run the device path, copy the outputs back, synchronise.
#include "pdegpu_internal.cuh"
#include <initializer_list>

// ---------------------------------------------------------------------------------------------
// device-pointer API
// ---------------------------------------------------------------------------------------------
static int check_system(pdegpu_ctx *ctx, const pdegpu_system *s, const char *who)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    if (!s) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "%s: system is NULL", who);
    if (s->family < PDEGPU_FLOW_ELIN4 || s->family > PDEGPU_PDE8) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "%s: unknown family %d", who, s->family);
    if (s->nrows < 3 || s->ncols < 3) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "%s: need nrows,ncols >= 3 (got %d x %d)", who, s->nrows, s->ncols);
    if (s->batch < 1 || s->batch > 65535) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "%s: batch must be in [1,65535]", who);
    if (s->batch > 1 && s->batch_stride < (long long)s->nrows * s->ncols) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "%s: batch_stride smaller than one field", who);
    const bool two = s->family <= PDEGPU_FLOW_LLIN8;
    const bool late = s->family >= PDEGPU_FLOW_LLIN4 && s->family <= PDEGPU_DISP_LLIN4;
    const bool eight = s->family == PDEGPU_FLOW_LLIN8 || s->family == PDEGPU_PDE8;
    const int nunk = two ? 2 : 1;
    for (int q = 0; q < nunk; q++) {
        if (!s->x[q] || !s->c[q] || !s->d[q]) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "%s: x/c/d[%d] is NULL", who, q);
        if (late && !s->x0[q]) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "%s: x0[%d] is NULL", who, q);
    }
    if (two && !s->m) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "%s: m is NULL", who);
    for (int n = 0; n < (eight ? 8 : 4); n++)
        if (!s->w[n]) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "%s: w[%d] is NULL", who, n);
    return PDEGPU_OK;
}

extern "C" int pdegpu_dev_relax(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega, int solver)
{
    int rc = check_system(ctx, sys, "pdegpu_dev_relax");
    if (rc) return rc;
    if (solver != 1 && solver != 2) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_relax: no such solver %d", solver);
    PDEGPU_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    if (iter <= 0 && !(sys->family == PDEGPU_PDE8 && solver == 2)) return PDEGPU_OK;
    if (solver == 1 && ctx->sweep_order == PDEGPU_ORDER_REFERENCE) return relax_lexpoint(ctx, sys, iter, omega);
    if (solver == 2 && (ctx->sweep_order == PDEGPU_ORDER_REFERENCE ||
                        (ctx->sweep_order == PDEGPU_ORDER_AUTO && sys->family == PDEGPU_FLOW_ELIN4)))
        return relax_lexline(ctx, sys, iter, omega);
    if (ctx->kernel_path == 1) {
        rc = relax_stream(ctx, sys, iter, omega, solver);
        if (rc != PDEGPU_ERR_UNSUPPORTED) return rc;
    }
    return relax_simple(ctx, sys, iter, omega, solver);
}

extern "C" int pdegpu_dev_residual(pdegpu_ctx *ctx, const pdegpu_system *sys, int nframes, float *RU, float *RV)
{
    int rc = check_system(ctx, sys, "pdegpu_dev_residual");
    if (rc) return rc;
    PDEGPU_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    return op_residual(ctx, sys, nframes, RU, RV, false);
}

extern "C" int pdegpu_dev_lhs(pdegpu_ctx *ctx, const pdegpu_system *sys, int nframes, float *AU, float *AV)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    if (!sys) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_lhs: system is NULL");
    // LHS needs no c[] terms
    pdegpu_system tmp = *sys;
    if (!tmp.c[0]) tmp.c[0] = tmp.d[0];
    if (!tmp.c[1]) tmp.c[1] = tmp.d[1];
    int rc = check_system(ctx, &tmp, "pdegpu_dev_lhs");
    if (rc) return rc;
    PDEGPU_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    return op_residual(ctx, &tmp, nframes, AU, AV, true);
}

extern "C" int pdegpu_dev_bilin_interp_2d(pdegpu_ctx *ctx, float *Iout, const float *Iin, const float *X, const float *Y,
                                          int nrows, int ncols, int nframes, float oob_value)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    PDEGPU_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    return op_bilin(ctx, Iout, Iin, X, Y, nrows, ncols, nframes, oob_value);
}

extern "C" int pdegpu_dev_fst_derivatives5(pdegpu_ctx *ctx, float *Idt, float *Idx, float *Idy,
                                           const float *It0, const float *It1, int nrows, int ncols, int nframes)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    PDEGPU_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    return op_fst(ctx, Idt, Idx, Idy, It0, It1, nrows, ncols, nframes);
}

extern "C" int pdegpu_dev_snd_derivatives5(pdegpu_ctx *ctx, float *Idxt, float *Idyt, float *Idxx, float *Idyy, float *Idxy,
                                           const float *It0, const float *It1, int nrows, int ncols, int nframes)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    PDEGPU_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    return op_snd(ctx, Idxt, Idyt, Idxx, Idyy, Idxy, It0, It1, nrows, ncols, nframes);
}

extern "C" int pdegpu_dev_ddiff_weights(pdegpu_ctx *ctx, float *wW, float *wN, float *wE, float *wS,
                                        const float *D, int nrows, int ncols, int nframes, float eps)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    PDEGPU_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    return op_ddiff(ctx, wW, wN, wE, wS, D, nrows, ncols, nframes, eps);
}

// ---------------------------------------------------------------------------------------------
// host-pointer API
// ---------------------------------------------------------------------------------------------
namespace {

// Stages one MEX call's arrays through the context's arena.
struct HostCall {
    pdegpu_ctx *ctx;
    int rc;
    HostCall(pdegpu_ctx *c, size_t total_floats, int narrays) : ctx(c), rc(PDEGPU_OK)
    {
        cudaError_t e = cudaSetDevice(ctx->device);
        if (e != cudaSuccess) { rc = pdegpu_check_cuda(ctx, e, "cudaSetDevice"); return; }
        pdegpu_arena_reset(ctx);
        rc = pdegpu_arena_reserve(ctx, total_floats * sizeof(float) + 256 * (size_t)(narrays + 1));
    }
    float *in(const float *h, size_t n)
    {
        if (rc) return nullptr;
        float *d = (float *)pdegpu_arena_alloc(ctx, n * sizeof(float));
        if (!d) { rc = pdegpu_set_error(ctx, PDEGPU_ERR_NOMEM, "device arena exhausted"); return nullptr; }
        cudaError_t e = cudaMemcpyAsync(d, h, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) { rc = pdegpu_check_cuda(ctx, e, "cudaMemcpyAsync(H2D)"); return nullptr; }
        return d;
    }
    float *out(size_t n, bool zero)
    {
        if (rc) return nullptr;
        float *d = (float *)pdegpu_arena_alloc(ctx, n * sizeof(float));
        if (!d) { rc = pdegpu_set_error(ctx, PDEGPU_ERR_NOMEM, "device arena exhausted"); return nullptr; }
        if (zero) {
            cudaError_t e = cudaMemsetAsync(d, 0, n * sizeof(float), ctx->stream);
            if (e != cudaSuccess) { rc = pdegpu_check_cuda(ctx, e, "cudaMemsetAsync"); return nullptr; }
        }
        return d;
    }
    void fetch(float *h, const float *d, size_t n)
    {
        if (rc) return;
        cudaError_t e = cudaMemcpyAsync(h, d, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) rc = pdegpu_check_cuda(ctx, e, "cudaMemcpyAsync(D2H)");
    }
    bool defer = false;              // batched entry points: the caller synchronises the lanes once at the end
    int finish()
    {
        if (rc) { cudaStreamSynchronize(ctx->stream); return rc; }
        if (defer) return PDEGPU_OK;
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) return pdegpu_check_cuda(ctx, e, "cudaStreamSynchronize");
        return PDEGPU_OK;
    }
};

int check_host(pdegpu_ctx *ctx, const char *who, int nrows, int ncols, int nframes, int solver, bool has_solver)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    if (nrows < 3 || ncols < 3) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "%s: need nrows,ncols >= 3 (got %d x %d)", who, nrows, ncols);
    if (nframes < 1) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "%s: nframes < 1", who);
    if (has_solver && solver != 1 && solver != 2) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "%s: no such solver", who);
    return PDEGPU_OK;
}

bool any_null(std::initializer_list<const void *> ps)
{
    for (const void *p : ps) if (!p) return true;
    return false;
}

}  // namespace

#define NULLCHECK(who, ...) \
    if (any_null({__VA_ARGS__})) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, who ": null pointer argument")

// common body of the two 4-neighbour flow gateways
static int flow_sor4(pdegpu_ctx *ctx, const char *who, int family,
                     float *o0, float *o1, float *RU, float *RV,
                     const float *U, const float *V, const float *dU, const float *dV, const float *M,
                     const float *Cu, const float *Cv, const float *Du, const float *Dv,
                     const float *wW, const float *wN, const float *wE, const float *wS,
                     int nrows, int ncols, int nframes, float iter, float omega, int solver,
                     int batch = 1, bool defer = false)
{
    int rc = check_host(ctx, who, nrows, ncols, nframes, solver, true);
    if (rc) return rc;
    const bool late = family == PDEGPU_FLOW_LLIN4;
    const bool want_res = RU && RV;
    if (want_res && batch != 1) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "%s: residuals of a batch are not supported", who);
    // `batch` systems one after the other in every array (n = all of them)
    const size_t n = (size_t)nrows * ncols * batch, nf = want_res ? n * nframes : n;   // data terms: all frames only for residuals
    const int it = (int)iter;
    HostCall hc(ctx, (late ? 4 : 2) * n + 5 * nf + 4 * n + 2 * n + (want_res ? 2 * nf : 0), 20);
    hc.defer = defer;
    pdegpu_system s;
    memset(&s, 0, sizeof(s));
    s.family = family; s.nrows = nrows; s.ncols = ncols; s.batch = batch; s.batch_stride = (long long)nrows * ncols;
    // the unknowns of the residual are the INPUTS (Oflow_sor_elin4_2d.c:350); the sweep works on a copy
    float *xin0 = hc.in(late ? dU : U, n), *xin1 = hc.in(late ? dV : V, n);
    if (late) { s.x0[0] = hc.in(U, n); s.x0[1] = hc.in(V, n); }
    s.m = hc.in(M, nf);
    s.c[0] = hc.in(Cu, nf); s.c[1] = hc.in(Cv, nf);
    s.d[0] = hc.in(Du, nf); s.d[1] = hc.in(Dv, nf);
    s.w[W_W] = hc.in(wW, n); s.w[W_N] = hc.in(wN, n); s.w[W_E] = hc.in(wE, n); s.w[W_S] = hc.in(wS, n);
    float *x0 = hc.out(n, it <= 0), *x1 = hc.out(n, it <= 0);
    float *dRU = nullptr, *dRV = nullptr;
    if (want_res) { dRU = hc.out(nf, false); dRV = hc.out(nf, false); }
    if (hc.rc) return hc.finish();
    if (want_res) {
        s.x[0] = xin0; s.x[1] = xin1;
        if ((rc = op_residual(ctx, &s, nframes, dRU, dRV, false))) { hc.rc = rc; return hc.finish(); }
        if (late && nframes > 1 && (rc = op_llin4_quirks(ctx, &s, nframes, dRU, dRV, false))) { hc.rc = rc; return hc.finish(); }
    }
    if (it > 0) {
        // gateway: memcpy(U_out, U_in) then relax in place (Oflow_sor_elin4_2d.c:341-346)
        cudaError_t e = cudaMemcpyAsync(x0, xin0, n * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(x1, xin1, n * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream);
        if (e != cudaSuccess) { hc.rc = pdegpu_check_cuda(ctx, e, "cudaMemcpyAsync(D2D)"); return hc.finish(); }
        s.x[0] = x0; s.x[1] = x1;
        if ((rc = pdegpu_dev_relax(ctx, &s, it, omega, solver))) { hc.rc = rc; return hc.finish(); }
    }
    hc.fetch(o0, x0, n); hc.fetch(o1, x1, n);
    if (want_res) { hc.fetch(RU, dRU, nf); hc.fetch(RV, dRV, nf); }
    return hc.finish();
}

extern "C" int pdegpu_oflow_sor_elin4_2d(pdegpu_ctx *ctx,
        float *U_out, float *V_out, float *RU, float *RV,
        const float *U, const float *V, const float *M,
        const float *Cu, const float *Cv, const float *Du, const float *Dv,
        const float *wW, const float *wN, const float *wE, const float *wS,
        int nrows, int ncols, int nframes, float iter, float omega, int solver)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    NULLCHECK("pdegpu_oflow_sor_elin4_2d", U_out, V_out, U, V, M, Cu, Cv, Du, Dv, wW, wN, wE, wS);
    return flow_sor4(ctx, "pdegpu_oflow_sor_elin4_2d", PDEGPU_FLOW_ELIN4, U_out, V_out, RU, RV,
                     U, V, nullptr, nullptr, M, Cu, Cv, Du, Dv, wW, wN, wE, wS, nrows, ncols, nframes, iter, omega, solver);
}

extern "C" int pdegpu_oflow_sor_llin4_2d(pdegpu_ctx *ctx,
        float *dU_out, float *dV_out, float *RU, float *RV,
        const float *U, const float *V, const float *dU, const float *dV, const float *M,
        const float *Cu, const float *Cv, const float *Du, const float *Dv,
        const float *wW, const float *wN, const float *wE, const float *wS,
        int nrows, int ncols, int nframes, float iter, float omega, int solver)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    NULLCHECK("pdegpu_oflow_sor_llin4_2d", dU_out, dV_out, U, V, dU, dV, M, Cu, Cv, Du, Dv, wW, wN, wE, wS);
    return flow_sor4(ctx, "pdegpu_oflow_sor_llin4_2d", PDEGPU_FLOW_LLIN4, dU_out, dV_out, RU, RV,
                     U, V, dU, dV, M, Cu, Cv, Du, Dv, wW, wN, wE, wS, nrows, ncols, nframes, iter, omega, solver);
}

// Batched, pipelined form of the two 4-neighbour flow gateways: `batch` independent systems per call, every array holding
// them one after the other. The systems go to the device in chunks on the context's lanes (own stream + arena each), so
// the upload of chunk k+1, the sweeps of chunk k and the download of chunk k-1 overlap; one synchronisation at the end.
static int flow_sor4_batch(pdegpu_ctx *ctx, const char *who, int family, float *o0, float *o1,
                           const float *U, const float *V, const float *dU, const float *dV, const float *M,
                           const float *Cu, const float *Cv, const float *Du, const float *Dv,
                           const float *wW, const float *wN, const float *wE, const float *wS,
                           int nrows, int ncols, int batch, float iter, float omega, int solver)
{
    int rc = check_host(ctx, who, nrows, ncols, 1, solver, true);
    if (rc) return rc;
    if (batch < 1) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "%s: batch < 1", who);
    if (ctx->parent) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "%s: not callable on a lane", who);
    PDEGPU_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const char *e_chunk = getenv("PDEGPU_HOST_CHUNK"), *e_lanes = getenv("PDEGPU_HOST_LANES");
    const int env_chunk = e_chunk ? atoi(e_chunk) : 0, env_lanes = e_lanes && atoi(e_lanes) > 0 ? atoi(e_lanes) : 3;
    const size_t n = (size_t)nrows * ncols;
    // chunks of ~64 MB of operands: large enough for full-rate copies and full-width kernels, small enough that the
    // first upload and the last download (the only parts that do not overlap) are a small share of the call
    int chunk = env_chunk > 0 ? env_chunk : (int)((size_t)(4u << 20) / n);
    if (chunk > (batch + 3) / 4) chunk = (batch + 3) / 4;
    if (chunk < 1) chunk = 1;
    const int nchunks = (batch + chunk - 1) / chunk;
    int nl = env_lanes < 1 ? 1 : (env_lanes > 8 ? 8 : env_lanes);
    if (nl > nchunks) nl = nchunks;
    if ((rc = pdegpu_lanes_prepare(ctx, nl, 0, who))) return rc;
    PDEGPU_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));       // earlier work of this context (the lanes do not wait on events)
    for (int c = 0; c < nchunks && !rc; c++) {
        pdegpu_ctx *l = ctx->lanes[c % nl];
        const size_t o = (size_t)c * chunk * n;
        const int nb = (c + 1) * chunk <= batch ? chunk : batch - c * chunk;
        rc = flow_sor4(l, who, family, o0 + o, o1 + o, nullptr, nullptr, U + o, V + o, dU ? dU + o : nullptr, dV ? dV + o : nullptr,
                       M + o, Cu + o, Cv + o, Du + o, Dv + o, wW + o, wN + o, wE + o, wS + o,
                       nrows, ncols, 1, iter, omega, solver, nb, true);
        if (rc) memcpy(ctx->err, l->err, sizeof ctx->err);
    }
    for (int k = 0; k < nl; k++) {
        pdegpu_ctx *l = ctx->lanes[k];
        cudaError_t e = cudaStreamSynchronize(l->stream);
        if (e != cudaSuccess && !rc) rc = pdegpu_check_cuda(ctx, e, "cudaStreamSynchronize(lane)");
        ctx->launches += l->launches; l->launches = 0;
    }
    return rc;
}

extern "C" int pdegpu_oflow_sor_elin4_2d_batch(pdegpu_ctx *ctx, float *U_out, float *V_out,
        const float *U, const float *V, const float *M,
        const float *Cu, const float *Cv, const float *Du, const float *Dv,
        const float *wW, const float *wN, const float *wE, const float *wS,
        int nrows, int ncols, int batch, float iter, float omega, int solver)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    NULLCHECK("pdegpu_oflow_sor_elin4_2d_batch", U_out, V_out, U, V, M, Cu, Cv, Du, Dv, wW, wN, wE, wS);
    return flow_sor4_batch(ctx, "pdegpu_oflow_sor_elin4_2d_batch", PDEGPU_FLOW_ELIN4, U_out, V_out,
                           U, V, nullptr, nullptr, M, Cu, Cv, Du, Dv, wW, wN, wE, wS, nrows, ncols, batch, iter, omega, solver);
}

extern "C" int pdegpu_oflow_sor_llin4_2d_batch(pdegpu_ctx *ctx, float *dU_out, float *dV_out,
        const float *U, const float *V, const float *dU, const float *dV, const float *M,
        const float *Cu, const float *Cv, const float *Du, const float *Dv,
        const float *wW, const float *wN, const float *wE, const float *wS,
        int nrows, int ncols, int batch, float iter, float omega, int solver)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    NULLCHECK("pdegpu_oflow_sor_llin4_2d_batch", dU_out, dV_out, U, V, dU, dV, M, Cu, Cv, Du, Dv, wW, wN, wE, wS);
    return flow_sor4_batch(ctx, "pdegpu_oflow_sor_llin4_2d_batch", PDEGPU_FLOW_LLIN4, dU_out, dV_out,
                           U, V, dU, dV, M, Cu, Cv, Du, Dv, wW, wN, wE, wS, nrows, ncols, batch, iter, omega, solver);
}

extern "C" int pdegpu_oflow_sor_llin8_2d(pdegpu_ctx *ctx,
        float *dU_out, float *dV_out,
        const float *U, const float *V, const float *dU, const float *dV, const float *M,
        const float *Cu, const float *Cv, const float *Du, const float *Dv,
        const float *wW, const float *wNW, const float *wN, const float *wNE,
        const float *wE, const float *wSE, const float *wS, const float *wSW,
        int nrows, int ncols, float iter, float omega, int solver)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    NULLCHECK("pdegpu_oflow_sor_llin8_2d", dU_out, dV_out, U, V, dU, dV, M, Cu, Cv, Du, Dv, wW, wNW, wN, wNE, wE, wSE, wS, wSW);
    int rc = check_host(ctx, "pdegpu_oflow_sor_llin8_2d", nrows, ncols, 1, solver, true);
    if (rc) return rc;
    const size_t n = (size_t)nrows * ncols;
    const int it = (int)iter;
    HostCall hc(ctx, 19 * n, 20);
    pdegpu_system s;
    memset(&s, 0, sizeof(s));
    s.family = PDEGPU_FLOW_LLIN8; s.nrows = nrows; s.ncols = ncols; s.batch = 1; s.batch_stride = (long long)n;
    float *x0 = hc.out(n, it <= 0), *x1 = hc.out(n, it <= 0);
    if (it > 0 && !hc.rc) {
        cudaError_t e = cudaMemcpyAsync(x0, dU, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(x1, dV, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) hc.rc = pdegpu_check_cuda(ctx, e, "cudaMemcpyAsync(H2D)");
        s.x[0] = x0; s.x[1] = x1;
        s.x0[0] = hc.in(U, n); s.x0[1] = hc.in(V, n);
        s.m = hc.in(M, n);
        s.c[0] = hc.in(Cu, n); s.c[1] = hc.in(Cv, n); s.d[0] = hc.in(Du, n); s.d[1] = hc.in(Dv, n);
        s.w[W_W] = hc.in(wW, n); s.w[W_N] = hc.in(wN, n); s.w[W_E] = hc.in(wE, n); s.w[W_S] = hc.in(wS, n);
        s.w[W_NW] = hc.in(wNW, n); s.w[W_NE] = hc.in(wNE, n); s.w[W_SE] = hc.in(wSE, n); s.w[W_SW] = hc.in(wSW, n);
        if (!hc.rc && (rc = pdegpu_dev_relax(ctx, &s, it, omega, solver))) hc.rc = rc;
    }
    hc.fetch(dU_out, x0, n); hc.fetch(dV_out, x1, n);
    return hc.finish();
}

static int flow_lhs(pdegpu_ctx *ctx, const char *who, int family, float *AU, float *AV,
                    const float *U, const float *V, const float *dU, const float *dV,
                    const float *M, const float *Du, const float *Dv,
                    const float *wW, const float *wN, const float *wE, const float *wS,
                    int nrows, int ncols, int nframes)
{
    int rc = check_host(ctx, who, nrows, ncols, nframes, 0, false);
    if (rc) return rc;
    const bool late = family == PDEGPU_FLOW_LLIN4;
    const size_t n = (size_t)nrows * ncols, nf = n * nframes;
    HostCall hc(ctx, (late ? 4 : 2) * n + 3 * nf + 4 * n + 2 * nf, 16);
    pdegpu_system s;
    memset(&s, 0, sizeof(s));
    s.family = family; s.nrows = nrows; s.ncols = ncols; s.batch = 1; s.batch_stride = (long long)n;
    s.x[0] = hc.in(late ? dU : U, n); s.x[1] = hc.in(late ? dV : V, n);
    if (late) { s.x0[0] = hc.in(U, n); s.x0[1] = hc.in(V, n); }
    s.m = hc.in(M, nf);
    s.d[0] = hc.in(Du, nf); s.d[1] = hc.in(Dv, nf);
    s.c[0] = s.d[0]; s.c[1] = s.d[1];
    s.w[W_W] = hc.in(wW, n); s.w[W_N] = hc.in(wN, n); s.w[W_E] = hc.in(wE, n); s.w[W_S] = hc.in(wS, n);
    float *dAU = hc.out(nf, false), *dAV = hc.out(nf, false);
    if (hc.rc) return hc.finish();
    if ((rc = op_residual(ctx, &s, nframes, dAU, dAV, true))) { hc.rc = rc; return hc.finish(); }
    if (late && (rc = op_llin4_quirks(ctx, &s, nframes, dAU, dAV, true))) { hc.rc = rc; return hc.finish(); }
    hc.fetch(AU, dAU, nf); hc.fetch(AV, dAV, nf);
    return hc.finish();
}

extern "C" int pdegpu_oflow_lhs_elin4_2d(pdegpu_ctx *ctx, float *AU, float *AV,
        const float *U, const float *V, const float *M, const float *Du, const float *Dv,
        const float *wW, const float *wN, const float *wE, const float *wS,
        int nrows, int ncols, int nframes)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    NULLCHECK("pdegpu_oflow_lhs_elin4_2d", AU, AV, U, V, M, Du, Dv, wW, wN, wE, wS);
    return flow_lhs(ctx, "pdegpu_oflow_lhs_elin4_2d", PDEGPU_FLOW_ELIN4, AU, AV, U, V, nullptr, nullptr, M, Du, Dv, wW, wN, wE, wS, nrows, ncols, nframes);
}

extern "C" int pdegpu_oflow_lhs_llin4_2d(pdegpu_ctx *ctx, float *AU, float *AV,
        const float *U, const float *V, const float *dU, const float *dV,
        const float *M, const float *Du, const float *Dv,
        const float *wW, const float *wN, const float *wE, const float *wS,
        int nrows, int ncols, int nframes)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    NULLCHECK("pdegpu_oflow_lhs_llin4_2d", AU, AV, U, V, dU, dV, M, Du, Dv, wW, wN, wE, wS);
    return flow_lhs(ctx, "pdegpu_oflow_lhs_llin4_2d", PDEGPU_FLOW_LLIN4, AU, AV, U, V, dU, dV, M, Du, Dv, wW, wN, wE, wS, nrows, ncols, nframes);
}

// scalar late-linearisation system (disparity). `always_copy`: the symmetric gateway copies the
// initial guess whatever `iter` is (Disp_sor_llin_sym4_2d.c:418-419), the plain one only for iter>0
// (Disp_sor_llin4_2d.c:276-279).
static int disp_stage(HostCall &hc, pdegpu_system &s, float *&xout,
                      const float *U, const float *dU, const float *Cu, const float *Du,
                      const float *wW, const float *wN, const float *wE, const float *wS,
                      int nrows, int ncols, bool copy_guess)
{
    const size_t n = (size_t)nrows * ncols;
    memset(&s, 0, sizeof(s));
    s.family = PDEGPU_DISP_LLIN4; s.nrows = nrows; s.ncols = ncols; s.batch = 1; s.batch_stride = (long long)n;
    xout = hc.out(n, !copy_guess);
    if (copy_guess && !hc.rc) {
        cudaError_t e = cudaMemcpyAsync(xout, dU, n * sizeof(float), cudaMemcpyHostToDevice, hc.ctx->stream);
        if (e != cudaSuccess) hc.rc = pdegpu_check_cuda(hc.ctx, e, "cudaMemcpyAsync(H2D)");
    }
    s.x[0] = xout;
    s.x0[0] = hc.in(U, n);
    s.c[0] = hc.in(Cu, n); s.d[0] = hc.in(Du, n);
    s.w[W_W] = hc.in(wW, n); s.w[W_N] = hc.in(wN, n); s.w[W_E] = hc.in(wE, n); s.w[W_S] = hc.in(wS, n);
    return hc.rc;
}

extern "C" int pdegpu_disp_sor_llin4_2d(pdegpu_ctx *ctx, float *dU_out, float *RU,
        const float *U, const float *dU, const float *Cu, const float *Du,
        const float *wW, const float *wN, const float *wE, const float *wS,
        int nrows, int ncols, float iter, float omega, int solver)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    NULLCHECK("pdegpu_disp_sor_llin4_2d", dU_out, U, dU, Cu, Du, wW, wN, wE, wS);
    int rc = check_host(ctx, "pdegpu_disp_sor_llin4_2d", nrows, ncols, 1, solver, true);
    if (rc) return rc;
    const size_t n = (size_t)nrows * ncols;
    const int it = (int)iter;
    HostCall hc(ctx, 9 * n, 10);
    pdegpu_system s;
    float *xout = nullptr;
    disp_stage(hc, s, xout, U, dU, Cu, Du, wW, wN, wE, wS, nrows, ncols, it > 0);
    if (!hc.rc && it > 0 && (rc = pdegpu_dev_relax(ctx, &s, it, omega, solver))) hc.rc = rc;
    hc.fetch(dU_out, xout, n);
    rc = hc.finish();
    // the reference allocates RU and never fills it (Disp_sor_llin4_2d.c:251-269)
    if (!rc && RU) memset(RU, 0, n * sizeof(float));
    return rc;
}

extern "C" int pdegpu_disp_sor_llin_sym4_2d(pdegpu_ctx *ctx, float *dU0_out, float *dU1_out,
        const float *U0, const float *dU0, const float *Cu0, const float *Du0,
        const float *wW0, const float *wN0, const float *wE0, const float *wS0,
        const float *U1, const float *dU1, const float *Cu1, const float *Du1,
        const float *wW1, const float *wN1, const float *wE1, const float *wS1,
        int nrows, int ncols, float iter, float omega, int solver)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    NULLCHECK("pdegpu_disp_sor_llin_sym4_2d", dU0_out, dU1_out, U0, dU0, Cu0, Du0, wW0, wN0, wE0, wS0, U1, dU1, Cu1, Du1, wW1, wN1, wE1, wS1);
    int rc = check_host(ctx, "pdegpu_disp_sor_llin_sym4_2d", nrows, ncols, 1, solver, true);
    if (rc) return rc;
    const size_t n = (size_t)nrows * ncols;
    const int it = (int)iter;
    HostCall hc(ctx, 18 * n, 20);
    pdegpu_system s0, s1;
    float *x0 = nullptr, *x1 = nullptr;
    disp_stage(hc, s0, x0, U0, dU0, Cu0, Du0, wW0, wN0, wE0, wS0, nrows, ncols, true);
    disp_stage(hc, s1, x1, U1, dU1, Cu1, Du1, wW1, wN1, wE1, wS1, nrows, ncols, true);
    // the two systems are independent (disparitySolvers.c:373-422): relax one after the other
    if (!hc.rc && it > 0 && (rc = pdegpu_dev_relax(ctx, &s0, it, omega, solver))) hc.rc = rc;
    if (!hc.rc && it > 0 && (rc = pdegpu_dev_relax(ctx, &s1, it, omega, solver))) hc.rc = rc;
    hc.fetch(dU0_out, x0, n); hc.fetch(dU1_out, x1, n);
    return hc.finish();
}

static int pde_solve(pdegpu_ctx *ctx, const char *who, int family, float *X_out,
                     const float *X, const float *TRACE, const float *B, const float *const w[8],
                     int nrows, int ncols, int nframes, float iter, float omega, int solver)
{
    int rc = check_host(ctx, who, nrows, ncols, nframes, solver, true);
    if (rc) return rc;
    const int nw = family == PDEGPU_PDE8 ? 8 : 4;
    const size_t n = (size_t)nrows * ncols * nframes;
    HostCall hc(ctx, (3 + nw) * n, 12);
    pdegpu_system s;
    memset(&s, 0, sizeof(s));
    s.family = family; s.nrows = nrows; s.ncols = ncols; s.batch = nframes; s.batch_stride = (long long)nrows * ncols;
    float *x = hc.in(X, n);                                // memcpy(Xnew, Xold) then in place (PDEsolver4.c:239)
    s.x[0] = x;
    s.d[0] = hc.in(TRACE, n); s.c[0] = hc.in(B, n);
    for (int k = 0; k < nw; k++) s.w[k] = hc.in(w[k], n);
    if (!hc.rc && (rc = pdegpu_dev_relax(ctx, &s, (int)iter, omega, solver))) hc.rc = rc;
    hc.fetch(X_out, x, n);
    return hc.finish();
}

extern "C" int pdegpu_pdesolver4(pdegpu_ctx *ctx, float *X_out,
        const float *X, const float *TRACE, const float *B,
        const float *wW, const float *wN, const float *wE, const float *wS,
        int nrows, int ncols, int nframes, float iter, float omega, int solver)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    NULLCHECK("pdegpu_pdesolver4", X_out, X, TRACE, B, wW, wN, wE, wS);
    const float *w[8] = {wW, wN, wE, wS, nullptr, nullptr, nullptr, nullptr};
    return pde_solve(ctx, "pdegpu_pdesolver4", PDEGPU_PDE4, X_out, X, TRACE, B, w, nrows, ncols, nframes, iter, omega, solver);
}

extern "C" int pdegpu_pdesolver8(pdegpu_ctx *ctx, float *X_out,
        const float *X, const float *TRACE, const float *B,
        const float *wW, const float *wNW, const float *wN, const float *wNE,
        const float *wE, const float *wSE, const float *wS, const float *wSW,
        int nrows, int ncols, int nframes, float iter, float omega, int solver)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    NULLCHECK("pdegpu_pdesolver8", X_out, X, TRACE, B, wW, wNW, wN, wNE, wE, wSE, wS, wSW);
    const float *w[8] = {wW, wN, wE, wS, wNW, wNE, wSE, wSW};
    return pde_solve(ctx, "pdegpu_pdesolver8", PDEGPU_PDE8, X_out, X, TRACE, B, w, nrows, ncols, nframes, iter, omega, solver);
}

extern "C" int pdegpu_bilin_interp_2d(pdegpu_ctx *ctx, float *Iout,
        const float *Iin, const float *X, const float *Y,
        int nrows, int ncols, int nframes, float oob_value)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    NULLCHECK("pdegpu_bilin_interp_2d", Iout, Iin, X, Y);
    if (nrows < 1 || ncols < 1 || nframes < 1) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "pdegpu_bilin_interp_2d: empty array");
    const size_t n = (size_t)nrows * ncols;
    HostCall hc(ctx, 2 * n * nframes + 2 * n, 5);
    float *dI = hc.in(Iin, n * nframes), *dX = hc.in(X, n), *dY = hc.in(Y, n), *dO = hc.out(n * nframes, false);
    int rc;
    if (!hc.rc && (rc = op_bilin(ctx, dO, dI, dX, dY, nrows, ncols, nframes, oob_value))) hc.rc = rc;
    hc.fetch(Iout, dO, n * nframes);
    return hc.finish();
}

extern "C" int pdegpu_fst_derivatives5(pdegpu_ctx *ctx, float *Idt, float *Idx, float *Idy,
        const float *It0, const float *It1, int nrows, int ncols, int nframes)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    NULLCHECK("pdegpu_fst_derivatives5", Idt, Idx, Idy, It0, It1);
    if (nrows < 5 || ncols < 5 || nframes < 1) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "pdegpu_fst_derivatives5: need nrows,ncols >= 5");
    const size_t n = (size_t)nrows * ncols * nframes;
    HostCall hc(ctx, 5 * n, 6);
    float *d0 = hc.in(It0, n), *d1 = hc.in(It1, n), *o0 = hc.out(n, false), *o1 = hc.out(n, false), *o2 = hc.out(n, false);
    int rc;
    if (!hc.rc && (rc = op_fst(ctx, o0, o1, o2, d0, d1, nrows, ncols, nframes))) hc.rc = rc;
    hc.fetch(Idt, o0, n); hc.fetch(Idx, o1, n); hc.fetch(Idy, o2, n);
    return hc.finish();
}

extern "C" int pdegpu_snd_derivatives5(pdegpu_ctx *ctx, float *Idxt, float *Idyt, float *Idxx, float *Idyy, float *Idxy,
        const float *It0, const float *It1, int nrows, int ncols, int nframes)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    NULLCHECK("pdegpu_snd_derivatives5", Idxt, Idyt, Idxx, Idyy, Idxy, It0, It1);
    if (nrows < 5 || ncols < 5 || nframes < 1) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "pdegpu_snd_derivatives5: need nrows,ncols >= 5");
    const size_t n = (size_t)nrows * ncols * nframes;
    HostCall hc(ctx, 7 * n, 8);
    float *d0 = hc.in(It0, n), *d1 = hc.in(It1, n);
    float *o[5];
    for (int k = 0; k < 5; k++) o[k] = hc.out(n, false);
    int rc;
    if (!hc.rc && (rc = op_snd(ctx, o[0], o[1], o[2], o[3], o[4], d0, d1, nrows, ncols, nframes))) hc.rc = rc;
    hc.fetch(Idxt, o[0], n); hc.fetch(Idyt, o[1], n); hc.fetch(Idxx, o[2], n); hc.fetch(Idyy, o[3], n); hc.fetch(Idxy, o[4], n);
    return hc.finish();
}

extern "C" int pdegpu_ddiff_weights(pdegpu_ctx *ctx, float *wW, float *wN, float *wE, float *wS,
        const float *D, int nrows, int ncols, int nframes, float eps)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    NULLCHECK("pdegpu_ddiff_weights", wW, wN, wE, wS, D);
    if (nrows < 2 || ncols < 2 || nframes < 1) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "pdegpu_ddiff_weights: need nrows,ncols >= 2");
    const size_t n = (size_t)nrows * ncols;
    HostCall hc(ctx, n * nframes + 4 * n, 6);
    float *dD = hc.in(D, n * nframes);
    float *o[4];
    for (int k = 0; k < 4; k++) o[k] = hc.out(n, false);
    int rc;
    if (!hc.rc && (rc = op_ddiff(ctx, o[0], o[1], o[2], o[3], dD, nrows, ncols, nframes, eps))) hc.rc = rc;
    hc.fetch(wW, o[0], n); hc.fetch(wN, o[1], n); hc.fetch(wE, o[2], n); hc.fetch(wS, o[3], n);
    return hc.finish();
}

// ---------------------------------------------------------------------------------------------
// driver-side stencils
// ---------------------------------------------------------------------------------------------
#define PDEGPU_ENTER(ctx)                                          \
    if (!(ctx)) return PDEGPU_ERR_ARG;                             \
    PDEGPU_CUDA_OK(ctx, cudaSetDevice((ctx)->device))

extern "C" int pdegpu_dev_op_diff_weights(pdegpu_ctx *ctx, float *wW, float *wN, float *wS, float *wE,
        const float *U, const float *V, int nrows, int ncols, int batch, long long batch_stride)
{
    PDEGPU_ENTER(ctx);
    if (!wW || !wN || !wS || !wE || !U || !V || nrows < 1 || ncols < 1 || batch < 1) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_op_diff_weights: bad argument");
    return op_opdiff(ctx, wW, wN, wS, wE, U, V, nrows, ncols, batch, batch_stride);
}

extern "C" int pdegpu_dev_llin_terms(pdegpu_ctx *ctx, const pdegpu_llin_terms *t)
{
    PDEGPU_ENTER(ctx);
    if (!t || t->nrows < 1 || t->ncols < 1 || t->batch < 1 || t->channels1 < 1 || t->channels2 < 0) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_llin_terms: bad argument");
    for (int k = 0; k < 3; k++) if (!t->d1[k]) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_llin_terms: d1[%d] is NULL", k);
    for (int k = 0; k < (t->channels2 ? (t->gradmag ? 5 : 3) : 0); k++) if (!t->d2[k]) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_llin_terms: d2[%d] is NULL", k);
    for (int k = 0; k < 5; k++) if (!t->out[k]) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_llin_terms: out[%d] is NULL", k);
    if (!t->dU || !t->dV) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_llin_terms: dU/dV is NULL");
    return op_llin_terms(ctx, t);
}

// The inner solve of the late-linearisation flow driver, FlowEminND_llin_2D_v10.m:278-348, as ONE call: diffusion
// weights of (U+dU, V+dV), robust data weights and channel sums, Oflow_sor_llin4_2d. Where the line kernels run from
// packed lines (generation 3 and the reference order) weights and terms are computed inside their preparation kernel
// and never exist as arrays (north_star subsystem 3); elsewhere the three steps run one after the other through `work`.
extern "C" int pdegpu_dev_llin_solve(pdegpu_ctx *ctx, const pdegpu_llin_terms *t, const float *U, const float *V, float *dU, float *dV,
                                     float *work, int iter, float omega, int solver)
{
    PDEGPU_ENTER(ctx);
    if (!t || !U || !V || !dU || !dV || !work || t->nrows < 3 || t->ncols < 3 || t->batch < 1 || t->channels1 < 1 || t->channels2 < 0)
        return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_llin_solve: bad argument");
    for (int k = 0; k < 3; k++) if (!t->d1[k]) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_llin_solve: d1[%d] is NULL", k);
    for (int k = 0; k < (t->channels2 ? (t->gradmag ? 5 : 3) : 0); k++) if (!t->d2[k]) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_llin_solve: d2[%d] is NULL", k);
    if (solver != 1 && solver != 2) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_llin_solve: no such solver %d", solver);
    const long long n = t->batch_stride, all = n * t->batch;
    pdegpu_llin_terms tt = *t;
    tt.dU = dU; tt.dV = dV;
    pdegpu_system sys;
    memset(&sys, 0, sizeof sys);
    sys.family = PDEGPU_FLOW_LLIN4; sys.nrows = t->nrows; sys.ncols = t->ncols; sys.batch = t->batch; sys.batch_stride = n;
    sys.x[0] = dU; sys.x[1] = dV; sys.x0[0] = U; sys.x0[1] = V;
    if (solver == 2 && ctx->kernel_path == 1 && iter > 0) {
        const int rc = relax_llin_fused(ctx, &sys, &tt, iter, omega, ctx->sweep_order == PDEGPU_ORDER_REFERENCE);
        if (rc != PDEGPU_ERR_UNSUPPORTED) return rc;
    }
    float *w[4], *T[5];
    for (int k = 0; k < 4; k++) w[k] = work + (2 + k) * all;
    for (int k = 0; k < 5; k++) { T[k] = work + (6 + k) * all; tt.out[k] = T[k]; }
    int rc;
    // OPdiffWeights(U+dU, V+dV): the sums are formed inside the weight kernel (no U+dU, V+dV arrays)    [wW wN wS wE]
    if ((rc = op_opdiff_sum(ctx, w[0], w[1], w[2], w[3], U, V, dU, dV, t->nrows, t->ncols, t->batch, n))) return rc;
    if ((rc = op_llin_terms(ctx, &tt))) return rc;
    sys.m = T[0]; sys.c[0] = T[1]; sys.c[1] = T[2]; sys.d[0] = T[3]; sys.d[1] = T[4];
    sys.w[W_W] = w[0]; sys.w[W_N] = w[1]; sys.w[W_S] = w[2]; sys.w[W_E] = w[3];
    return pdegpu_dev_relax(ctx, &sys, iter, omega, solver);
}

extern "C" int pdegpu_dev_elin_terms(pdegpu_ctx *ctx, const pdegpu_elin_terms *t)
{
    PDEGPU_ENTER(ctx);
    if (!t || t->nrows < 1 || t->ncols < 1 || t->channels < 1 || !t->U || !t->V) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_elin_terms: bad argument");
    for (int k = 0; k < 8; k++) if (!t->der[k]) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_elin_terms: der[%d] is NULL", k);
    for (int k = 0; k < 5; k++) if (!t->coef[k] || !t->out[k]) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_elin_terms: coef/out[%d] is NULL", k);
    return op_elin_terms(ctx, t);
}

extern "C" int pdegpu_dev_disp_sym_terms(pdegpu_ctx *ctx, const pdegpu_disp_sym_terms *t)
{
    PDEGPU_ENTER(ctx);
    if (!t || t->nrows < 1 || t->ncols < 1 || t->channels < 1 || !t->dU || !t->Udt || !t->Udx || !t->CuG || !t->DuG) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_disp_sym_terms: bad argument");
    for (int k = 0; k < 6; k++) if (!t->d[k]) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_disp_sym_terms: d[%d] is NULL", k);
    return op_disp_sym_terms(ctx, t);
}

extern "C" int pdegpu_dev_fas_rhs(pdegpu_ctx *ctx, float *f, const float *R, const float *A, const float *gd, long long n)
{
    PDEGPU_ENTER(ctx);
    if (!f || !R || !A || !gd || n < 1) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_fas_rhs: bad argument");
    return op_fas_rhs(ctx, f, R, A, gd, n);
}

extern "C" int pdegpu_dev_imfilter(pdegpu_ctx *ctx, float *out, const float *in, int nrows, int ncols, int planes,
        long long in_stride, long long out_stride, const double *h, int kr, int kc, int step, float prescale)
{
    PDEGPU_ENTER(ctx);
    if (!out || !in || !h || nrows < 1 || ncols < 1 || planes < 1) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_imfilter: bad argument");
    return op_imfilter(ctx, out, in, nrows, ncols, planes, in_stride, out_stride, h, kr, kc, step, prescale);
}

int imresize_2d(pdegpu_ctx *ctx, float *out, float *tmp, const float *in, int in_rows, int in_cols, int out_rows, int out_cols,
                double scale_rows, double scale_cols, int antialias, int planes, int cubic)
{
    // imresize resizes the dimension with the smaller scale factor first (rows on a tie)
    int rc;
    if (scale_rows <= scale_cols) {
        rc = op_imresize_dim(ctx, tmp, in, 0, in_rows, out_rows, in_cols, scale_rows, antialias, planes, (long long)in_rows * in_cols, (long long)out_rows * in_cols, cubic);
        if (rc) return rc;
        return op_imresize_dim(ctx, out, tmp, 1, in_cols, out_cols, out_rows, scale_cols, antialias, planes, (long long)out_rows * in_cols, (long long)out_rows * out_cols, cubic);
    }
    rc = op_imresize_dim(ctx, tmp, in, 1, in_cols, out_cols, in_rows, scale_cols, antialias, planes, (long long)in_rows * in_cols, (long long)in_rows * out_cols, cubic);
    if (rc) return rc;
    return op_imresize_dim(ctx, out, tmp, 0, in_rows, out_rows, out_cols, scale_rows, antialias, planes, (long long)in_rows * out_cols, (long long)out_rows * out_cols, cubic);
}

static int imresize_check(pdegpu_ctx *ctx, const void *out, const void *tmp, const void *in, int in_rows, int in_cols, int out_rows, int out_cols,
                          double scale_rows, double scale_cols, int planes)
{
    if (!out || !tmp || !in || in_rows < 1 || in_cols < 1 || out_rows < 1 || out_cols < 1 || planes < 1 || !(scale_rows > 0) || !(scale_cols > 0))
        return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_imresize: bad argument");
    return PDEGPU_OK;
}

extern "C" int pdegpu_dev_imresize_bilinear(pdegpu_ctx *ctx, float *out, float *tmp, const float *in,
        int in_rows, int in_cols, int out_rows, int out_cols, double scale_rows, double scale_cols, int antialias, int planes)
{
    PDEGPU_ENTER(ctx);
    int rc = imresize_check(ctx, out, tmp, in, in_rows, in_cols, out_rows, out_cols, scale_rows, scale_cols, planes);
    if (rc) return rc;
    return imresize_2d(ctx, out, tmp, in, in_rows, in_cols, out_rows, out_cols, scale_rows, scale_cols, antialias, planes, 0);
}

extern "C" int pdegpu_dev_imresize_bicubic(pdegpu_ctx *ctx, float *out, float *tmp, const float *in,
        int in_rows, int in_cols, int out_rows, int out_cols, double scale_rows, double scale_cols, int antialias, int planes)
{
    PDEGPU_ENTER(ctx);
    int rc = imresize_check(ctx, out, tmp, in, in_rows, in_cols, out_rows, out_cols, scale_rows, scale_cols, planes);
    if (rc) return rc;
    return imresize_2d(ctx, out, tmp, in, in_rows, in_cols, out_rows, out_cols, scale_rows, scale_cols, antialias, planes, 1);
}

extern "C" int pdegpu_dev_medfilt3(pdegpu_ctx *ctx, float *out, const float *in, int nrows, int ncols, int planes, long long stride)
{
    PDEGPU_ENTER(ctx);
    if (!out || !in || out == in || nrows < 1 || ncols < 1 || planes < 1) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_medfilt3: bad argument (out must differ from in)");
    return op_medfilt3(ctx, out, in, nrows, ncols, planes, stride);
}

extern "C" int pdegpu_dev_axpby(pdegpu_ctx *ctx, float *out, float a, const float *x, float b, const float *y, long long n)
{
    PDEGPU_ENTER(ctx);
    if (!out || !x || n < 1) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_axpby: bad argument");
    return op_axpby(ctx, out, a, x, b, y, n);
}

extern "C" int pdegpu_dev_warp_coords(pdegpu_ctx *ctx, float *X, float *Y, const float *U, const float *V,
        int nrows, int ncols, int batch, long long batch_stride)
{
    PDEGPU_ENTER(ctx);
    if (!X || !Y || nrows < 1 || ncols < 1 || batch < 1) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_warp_coords: bad argument");
    return op_warp_coords(ctx, X, Y, U, V, nrows, ncols, batch, batch_stride);
}

extern "C" int pdegpu_dev_ad_diff_weights(pdegpu_ctx *ctx, float *const w[8], float *TRACE, float *B,
        const float *D, const float *Iin, int nrows, int ncols, int nframes, double quantile, double scale, double *lambda_dev)
{
    PDEGPU_ENTER(ctx);
    if (!w || !D || nrows < 1 || ncols < 1 || nframes < 1 || !(quantile >= 0.0 && quantile <= 1.0)) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_ad_diff_weights: bad argument");
    for (int k = 0; k < 8; k++) if (!w[k]) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_ad_diff_weights: w[%d] is NULL", k);
    if ((TRACE || B) && !(TRACE && B && Iin)) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_dev_ad_diff_weights: TRACE, B and Iin go together");
    return op_ad_diff_weights(ctx, w, TRACE, B, D, Iin, nrows, ncols, nframes, quantile, scale, lambda_dev);
}
