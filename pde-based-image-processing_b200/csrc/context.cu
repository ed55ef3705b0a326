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
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { free(ctx); return pdegpu_check_cuda(nullptr, e, "cudaStreamCreate"); }
    *out = ctx;
    return PDEGPU_OK;
}

extern "C" void pdegpu_free(pdegpu_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->arena) cudaFree(ctx->arena);
    if (ctx->scratch) cudaFree(ctx->scratch);
    cudaStreamDestroy(ctx->stream);
    free(ctx);
}

extern "C" int pdegpu_sync(pdegpu_ctx *ctx)
{
    if (!ctx) return PDEGPU_ERR_ARG;
    PDEGPU_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return PDEGPU_OK;
}

extern "C" void *pdegpu_stream(pdegpu_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" unsigned long long pdegpu_launch_count(const pdegpu_ctx *ctx) { return ctx ? ctx->launches : 0ull; }

extern "C" int pdegpu_set_kernel_path(pdegpu_ctx *ctx, int path)
{
    if (!ctx || path < 0 || path > 1) return PDEGPU_ERR_ARG;
    ctx->kernel_path = path;
    return PDEGPU_OK;
}

// ---------------------------------------------------------------------------------------------
int pdegpu_arena_reserve(pdegpu_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->arena_bytes) return PDEGPU_OK;
    // growing invalidates earlier blocks: only legal while the arena is empty
    if (ctx->arena_used != 0) return pdegpu_set_error(ctx, PDEGPU_ERR_NOMEM, "arena grow while in use");
    PDEGPU_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->arena) { cudaFree(ctx->arena); ctx->arena = nullptr; ctx->arena_bytes = 0; }
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
