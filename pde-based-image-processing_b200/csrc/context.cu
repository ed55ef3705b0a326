// context.cu -- context lifetime, error reporting, device arena / scratch, memory helpers.
#include "pdegpu_internal.cuh"
#include <stdarg.h>
#include <stdlib.h>

static thread_local char g_init_err[512] = "";

int pdegpu_set_error(pdegpu_ctx *ctx, int status, const char *fmt, ...)
{
    char *dst = ctx ? ctx->err : g_init_err;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return status;
}

int pdegpu_check_cuda(pdegpu_ctx *ctx, cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return PDEGPU_OK;
    int st = (e == cudaErrorMemoryAllocation) ? PDEGPU_ERR_NOMEM : PDEGPU_ERR_CUDA;
    return pdegpu_set_error(ctx, st, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
}

extern "C" const char *pdegpu_version(void) { return "libpdegpu 0.1 (sm_100a)"; }

extern "C" int pdegpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" const char *pdegpu_last_error(const pdegpu_ctx *ctx) { return ctx ? ctx->err : g_init_err; }

extern "C" int pdegpu_init(int device, pdegpu_ctx **out)
{
    if (!out) return pdegpu_set_error(nullptr, PDEGPU_ERR_ARG, "pdegpu_init: ctx is NULL");
    *out = nullptr;
    int n = pdegpu_device_count();
    if (n <= 0) return pdegpu_set_error(nullptr, PDEGPU_ERR_NODEVICE, "pdegpu_init: no CUDA device visible (libpdegpu has no CPU fallback)");
    if (device < 0 || device >= n) return pdegpu_set_error(nullptr, PDEGPU_ERR_ARG, "pdegpu_init: device %d out of range [0,%d)", device, n);
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return pdegpu_check_cuda(nullptr, e, "cudaSetDevice");
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return pdegpu_check_cuda(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major != 10)
        return pdegpu_set_error(nullptr, PDEGPU_ERR_NODEVICE,
                                "pdegpu_init: device %d is sm_%d%d; libpdegpu is built for sm_100a only", device, prop.major, prop.minor);
    pdegpu_ctx *ctx = (pdegpu_ctx *)calloc(1, sizeof(pdegpu_ctx));
    if (!ctx) return pdegpu_set_error(nullptr, PDEGPU_ERR_NOMEM, "pdegpu_init: out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->kernel_path = 1;
    const char *env = getenv("PDEGPU_KERNELS");
    if (env && strcmp(env, "simple") == 0) ctx->kernel_path = 0;
    ctx->sweep_order = pdegpu_order_from_env();
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { free(ctx); return pdegpu_check_cuda(nullptr, e, "cudaStreamCreate"); }
    *out = ctx;
    return PDEGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// CUDA graphs of whole pipelines
// ---------------------------------------------------------------------------------------------
constexpr int kGraphSlots = 8, kGraphKeyMax = 256;
struct pdegpu_graph_entry {
    unsigned char key[kGraphKeyMax];
    size_t key_len;
    unsigned epoch;
    int state;                       // 0 empty, 1 seen once (ran directly), 2 graph ready, 3 capture failed: always direct
    cudaGraphExec_t exec;
    unsigned long long launches;     // launches one replay stands for
    unsigned long long last_use;
};

static void pdegpu_graphs_drop(pdegpu_ctx *ctx)
{
    if (!ctx->graphs) return;
    for (int k = 0; k < kGraphSlots; k++) {
        if (ctx->graphs[k].state == 2 && ctx->graphs[k].exec) cudaGraphExecDestroy(ctx->graphs[k].exec);
        ctx->graphs[k].state = 0; ctx->graphs[k].exec = nullptr;
    }
}

int pdegpu_graph_run(pdegpu_ctx *ctx, const void *key, size_t key_len, pdegpu_graph_body body)
{
    static const int enabled = getenv("PDEGPU_GRAPHS") ? atoi(getenv("PDEGPU_GRAPHS")) : 1;
    unsigned long long &tick = ctx->graph_tick;             // least-recently-used clock of this context's cache
    if (!enabled || ctx->prof_on || ctx->capturing || key_len > (size_t)kGraphKeyMax) return body.fn(body.arg);
    if (!ctx->graphs) {
        ctx->graphs = (pdegpu_graph_entry *)calloc(kGraphSlots, sizeof(pdegpu_graph_entry));
        if (!ctx->graphs) return body.fn(body.arg);
    }
    pdegpu_graph_entry *e = nullptr, *victim = nullptr;
    for (int k = 0; k < kGraphSlots; k++) {
        pdegpu_graph_entry &g = ctx->graphs[k];
        if (g.state && g.epoch != ctx->graph_epoch) {                         // its device pointers are gone
            if (g.state == 2 && g.exec) cudaGraphExecDestroy(g.exec);
            g.state = 0; g.exec = nullptr;
        }
        if (g.state && g.key_len == key_len && memcmp(g.key, key, key_len) == 0) e = &g;
        // replacement: an empty slot if there is one, else the least recently used
        if (!victim || (victim->state != 0 && (g.state == 0 || g.last_use < victim->last_use))) victim = &g;
    }
    if (!e) {                                                                 // first sight of this call: run it as it is
        if (victim->state == 2 && victim->exec) cudaGraphExecDestroy(victim->exec);
        memset(victim, 0, sizeof *victim);
        memcpy(victim->key, key, key_len);
        victim->key_len = key_len; victim->state = 1; victim->last_use = ++tick;
        const int rc = body.fn(body.arg);
        victim->epoch = ctx->graph_epoch;                                     // after the run: the run itself may have grown a buffer
        return rc;
    }
    e->last_use = ++tick;
    if (e->state == 3) return body.fn(body.arg);
    if (e->state == 1) {                                                      // second call: capture
        const unsigned long long l0 = ctx->launches;
        if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError(); e->state = 3; return body.fn(body.arg);
        }
        ctx->capturing = 1;
        const int rc = body.fn(body.arg);
        ctx->capturing = 0;
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
        if (rc != PDEGPU_OK || ce != cudaSuccess || !graph || e->epoch != ctx->graph_epoch) {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            e->state = 3; e->epoch = ctx->graph_epoch;
            ctx->launches = l0;
            return body.fn(body.arg);                                         // nothing ran during the capture
        }
        cudaGraphExec_t exec = nullptr;
        const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess || !exec) { cudaGetLastError(); e->state = 3; ctx->launches = l0; return body.fn(body.arg); }
        e->exec = exec; e->state = 2; e->launches = ctx->launches - l0;
        ctx->launches = l0;
    }
    ctx->launches += e->launches;
    PDEGPU_CUDA_OK(ctx, cudaGraphLaunch(e->exec, ctx->stream));
    return PDEGPU_OK;
}

extern "C" void pdegpu_free(pdegpu_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->prof) {
        for (int k = 0; k < ctx->prof_cap; k++) if (ctx->prof[k].e0) { cudaEventDestroy(ctx->prof[k].e0); cudaEventDestroy(ctx->prof[k].e1); }
        free(ctx->prof);
    }
    pdegpu_graphs_drop(ctx);
    free(ctx->graphs);
    for (int k = 0; k < ctx->nlanes; k++) pdegpu_free(ctx->lanes[k]);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->arena) cudaFree(ctx->arena);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->work) cudaFree(ctx->work);
    cudaStreamDestroy(ctx->stream);
    free(ctx);
}

// ---------------------------------------------------------------------------------------------
// lanes (pdegpu_internal.cuh)
// ---------------------------------------------------------------------------------------------
int pdegpu_lane_count(pdegpu_ctx *ctx, int batch, bool serial_order)
{
    // zebra / red-black kernels fill the GPU from one problem: 32 lanes only hide the launch-bound coarse levels. The
    // reference-order kernels run ONE CTA per problem -- there the lanes ARE the parallelism (FMG 1080p: 4.3 / 18.0 / 34.5
    // flows/s with 8 / 32 / 64 pairs side by side), at one workspace per lane (0.67 GB per 1080p pair)
    static const int env_lanes = getenv("PDEGPU_LANES") ? atoi(getenv("PDEGPU_LANES")) : 0;
    const int max_lanes = env_lanes > 0 ? env_lanes : (serial_order ? 64 : 32);
    if (ctx->prof_on || ctx->parent || batch < 2 || max_lanes < 2) return 1;
    const int cap = (int)(sizeof(ctx->lanes) / sizeof(ctx->lanes[0]));
    int n = batch < max_lanes ? batch : max_lanes;
    return n < cap ? n : cap;
}

int pdegpu_lanes_prepare(pdegpu_ctx *ctx, int n, size_t work_bytes, const char *who)
{
    if (!ctx->ev_fork) {
        PDEGPU_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    }
    while (ctx->nlanes < n) {
        pdegpu_ctx *c = (pdegpu_ctx *)calloc(1, sizeof(pdegpu_ctx));
        if (!c) return pdegpu_set_error(ctx, PDEGPU_ERR_NOMEM, "%s: out of host memory", who);
        c->device = ctx->device; c->sm_count = ctx->sm_count; c->parent = ctx;
        cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
        if (e != cudaSuccess) { free(c); return pdegpu_check_cuda(ctx, e, "lane stream"); }
        ctx->lanes[ctx->nlanes++] = c;
    }
    for (int k = 0; k < n; k++) {
        pdegpu_ctx *c = ctx->lanes[k];
        c->kernel_path = ctx->kernel_path; c->sweep_order = ctx->sweep_order; c->capturing = 1;   // (a lane never builds graphs of its own)
        const int rc = pdegpu_work_reserve(c, work_bytes, who);
        if (rc) { memcpy(ctx->err, c->err, sizeof ctx->err); return rc; }
    }
    return PDEGPU_OK;
}

int pdegpu_lanes_fork(pdegpu_ctx *ctx, int n)
{
    PDEGPU_CUDA_OK(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    for (int k = 0; k < n; k++) PDEGPU_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->lanes[k]->stream, ctx->ev_fork, 0));
    return PDEGPU_OK;
}

int pdegpu_lanes_join(pdegpu_ctx *ctx, int n)
{
    for (int k = 0; k < n; k++) {
        pdegpu_ctx *c = ctx->lanes[k];
        PDEGPU_CUDA_OK(ctx, cudaEventRecord(c->ev_join, c->stream));
        PDEGPU_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->stream, c->ev_join, 0));
        ctx->launches += c->launches;
        c->launches = 0;
    }
    return PDEGPU_OK;
}

extern "C" int pdegpu_sync(pdegpu_ctx *ctx)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    PDEGPU_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return PDEGPU_OK;
}

extern "C" void *pdegpu_stream(pdegpu_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" unsigned long long pdegpu_launch_count(const pdegpu_ctx *ctx) { return ctx ? ctx->launches : 0ull; }

// ---------------------------------------------------------------------------------------------
// per-launch profile: event pair around every kernel launch made through PDEGPU_PROF / LAUNCH_CHECK
// ---------------------------------------------------------------------------------------------
void pdegpu_prof_begin(pdegpu_ctx *ctx, const char *name, double bytes)
{
    if (ctx->prof_n >= ctx->prof_cap) return;
    pdegpu_prof_rec *r = &ctx->prof[ctx->prof_n];
    r->name = name;
    r->bytes = bytes;
    if (!r->e0) { cudaEventCreate(&r->e0); cudaEventCreate(&r->e1); }
    cudaEventRecord(r->e0, ctx->stream);
    ctx->prof_on = 2;      // a record is open
}

void pdegpu_prof_end(pdegpu_ctx *ctx)
{
    if (ctx->prof_on != 2) return;
    cudaEventRecord(ctx->prof[ctx->prof_n].e1, ctx->stream);
    ctx->prof_n++;
    ctx->prof_on = 1;
}

extern "C" int pdegpu_profile_enable(pdegpu_ctx *ctx, int on)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    if (on && !ctx->prof) {
        ctx->prof_cap = 1 << 15;
        ctx->prof = (pdegpu_prof_rec *)calloc(ctx->prof_cap, sizeof(pdegpu_prof_rec));
        if (!ctx->prof) { ctx->prof_cap = 0; return pdegpu_set_error(ctx, PDEGPU_ERR_NOMEM, "profile buffer"); }
    }
    ctx->prof_n = 0;
    ctx->prof_on = on ? 1 : 0;
    return PDEGPU_OK;
}

// Aggregates the recorded launches by kernel name into a JSON array:
// [{"kernel": "...", "launches": n, "ms_total": t, "bytes_total": b}, ...]. Synchronises the stream.
extern "C" int pdegpu_profile_report(pdegpu_ctx *ctx, char *buf, size_t buflen)
{
    if (!ctx || !buf || buflen < 3) return PDEGPU_ERR_ARG;
    PDEGPU_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    struct agg { const char *name; int n; double ms, bytes; } a[64];
    int na = 0;
    for (int k = 0; k < ctx->prof_n; k++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->prof[k].e0, ctx->prof[k].e1) != cudaSuccess) { cudaGetLastError(); continue; }
        int q = 0;
        for (; q < na; q++) if (strcmp(a[q].name, ctx->prof[k].name) == 0) break;
        if (q == na) { if (na == 64) continue; a[na].name = ctx->prof[k].name; a[na].n = 0; a[na].ms = 0; a[na].bytes = 0; na++; }
        a[q].n++; a[q].ms += ms; a[q].bytes += ctx->prof[k].bytes;
    }
    size_t off = 0;
    off += snprintf(buf + off, buflen - off, "[");
    for (int q = 0; q < na && off + 200 < buflen; q++)
        off += snprintf(buf + off, buflen - off, "%s{\"kernel\": \"%s\", \"launches\": %d, \"ms_total\": %.6f, \"bytes_total\": %.0f}",
                        q ? ", " : "", a[q].name, a[q].n, a[q].ms, a[q].bytes);
    snprintf(buf + off, buflen - off, "]");
    return PDEGPU_OK;
}

extern "C" int pdegpu_set_kernel_path(pdegpu_ctx *ctx, int path)
{
    if (!ctx || path < 0 || path > 1) return PDEGPU_ERR_ARG;
    if (ctx->kernel_path != path) ctx->graph_epoch++;      // captured graphs hold the other path's kernels
    ctx->kernel_path = path;
    return PDEGPU_OK;
}

extern "C" int pdegpu_order_from_env(void)
{
    const char *env = getenv("PDEGPU_ORDER");
    if (env && strcmp(env, "reference") == 0) return PDEGPU_ORDER_REFERENCE;
    if (env && strcmp(env, "fast") == 0) return PDEGPU_ORDER_FAST;
    return PDEGPU_ORDER_AUTO;
}

extern "C" int pdegpu_set_sweep_order(pdegpu_ctx *ctx, int order)
{
    if (!ctx || order < PDEGPU_ORDER_FAST || order > PDEGPU_ORDER_AUTO) return PDEGPU_ERR_ARG;
    if (ctx->sweep_order != order) ctx->graph_epoch++;     // captured graphs hold the other order's kernels
    ctx->sweep_order = order;
    return PDEGPU_OK;
}

extern "C" int pdegpu_get_sweep_order(const pdegpu_ctx *ctx) { return ctx ? ctx->sweep_order : PDEGPU_ERR_ARG; }

// ---------------------------------------------------------------------------------------------
int pdegpu_arena_reserve(pdegpu_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->arena_bytes) return PDEGPU_OK;
    // growing invalidates earlier blocks: only legal while the arena is empty
    if (ctx->arena_used != 0) return pdegpu_set_error(ctx, PDEGPU_ERR_NOMEM, "arena grow while in use");
    PDEGPU_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->arena) { cudaFree(ctx->arena); ctx->arena = nullptr; ctx->arena_bytes = 0; }
    ctx->graph_epoch++;
    size_t want = bytes + (bytes >> 3) + (1u << 20);
    cudaError_t e = cudaMalloc((void **)&ctx->arena, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        e = cudaMalloc((void **)&ctx->arena, want);
        if (e != cudaSuccess) return pdegpu_check_cuda(ctx, e, "cudaMalloc(arena)");
    }
    ctx->arena_bytes = want;
    return PDEGPU_OK;
}

void pdegpu_arena_reset(pdegpu_ctx *ctx) { ctx->arena_used = 0; }

void *pdegpu_arena_alloc(pdegpu_ctx *ctx, size_t bytes)
{
    size_t off = (ctx->arena_used + 255) & ~(size_t)255;
    if (off + bytes > ctx->arena_bytes) return nullptr;
    ctx->arena_used = off + bytes;
    return ctx->arena + off;
}

int pdegpu_scratch_reserve(pdegpu_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->scratch_bytes) return PDEGPU_OK;
    PDEGPU_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->scratch) { cudaFree(ctx->scratch); ctx->scratch = nullptr; ctx->scratch_bytes = 0; }
    ctx->graph_epoch++;
    if (ctx->parent) ctx->parent->graph_epoch++;
    size_t want = bytes + (bytes >> 3) + (1u << 20);
    cudaError_t e = cudaMalloc((void **)&ctx->scratch, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        e = cudaMalloc((void **)&ctx->scratch, want);
        if (e != cudaSuccess) return pdegpu_check_cuda(ctx, e, "cudaMalloc(scratch)");
    }
    ctx->scratch_bytes = want;
    return PDEGPU_OK;
}

// workspace of the device-resident pipelines
int pdegpu_work_reserve(pdegpu_ctx *ctx, size_t bytes, const char *who)
{
    if (bytes <= ctx->work_bytes) return PDEGPU_OK;
    PDEGPU_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->work) cudaFree(ctx->work);
    ctx->work = nullptr; ctx->work_bytes = 0;
    ctx->graph_epoch++;
    if (ctx->parent) ctx->parent->graph_epoch++;
    if (cudaMalloc((void **)&ctx->work, bytes) != cudaSuccess) {
        cudaGetLastError();
        return pdegpu_set_error(ctx, PDEGPU_ERR_NOMEM, "%s: cannot allocate %zu bytes of workspace", who, bytes);
    }
    ctx->work_bytes = bytes;
    return PDEGPU_OK;
}

// ---------------------------------------------------------------------------------------------
extern "C" int pdegpu_malloc(pdegpu_ctx *ctx, void **dptr, size_t bytes)
{
    if (!ctx || !dptr) return PDEGPU_ERR_ARG;
    PDEGPU_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    PDEGPU_CUDA_OK(ctx, cudaMalloc(dptr, bytes ? bytes : 1));
    return PDEGPU_OK;
}

extern "C" int pdegpu_free_mem(pdegpu_ctx *ctx, void *dptr)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    PDEGPU_CUDA_OK(ctx, cudaFree(dptr));
    return PDEGPU_OK;
}

extern "C" int pdegpu_host_alloc(pdegpu_ctx *ctx, void **hptr, size_t bytes)
{
    if (!ctx || !hptr) return PDEGPU_ERR_ARG;
    PDEGPU_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    PDEGPU_CUDA_OK(ctx, cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return PDEGPU_OK;
}

extern "C" int pdegpu_host_free(pdegpu_ctx *ctx, void *hptr)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    PDEGPU_CUDA_OK(ctx, cudaFreeHost(hptr));
    return PDEGPU_OK;
}

extern "C" int pdegpu_upload(pdegpu_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes)
{
    if (!ctx || (!dst_dev && bytes) || (!src_host && bytes)) return PDEGPU_ERR_ARG;
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return PDEGPU_OK;
}

extern "C" int pdegpu_download(pdegpu_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes)
{
    if (!ctx || (!dst_host && bytes) || (!src_dev && bytes)) return PDEGPU_ERR_ARG;
    PDEGPU_CUDA_OK(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return PDEGPU_OK;
}
