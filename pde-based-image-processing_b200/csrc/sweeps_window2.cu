// sweeps_window2.cu -- entry of kernel generation 2b (sweeps_window2_impl.cuh) + the late-linearisation flow family,
// the one the benchmark runs. The other families are instantiated in sweeps_window2_b/c/d.cu.
#include "sweeps_window2_impl.cuh"

int window2_family_0(pdegpu_ctx *ctx, int dir, void *params, int M, int batch);
int window2_family_1(pdegpu_ctx *ctx, int dir, void *params, int M, int batch);
int window2_family_2(pdegpu_ctx *ctx, int dir, void *params, int M, int batch);
int window2_family_3(pdegpu_ctx *ctx, int dir, void *params, int M, int batch);
int window2_family_4(pdegpu_ctx *ctx, int dir, void *params, int M, int batch);

static_assert(PDEGPU_FLOW_ELIN4 == 0 && PDEGPU_FLOW_LLIN4 == 1 && PDEGPU_FLOW_LLIN8 == 2 && PDEGPU_DISP_LLIN4 == 3 && PDEGPU_PDE4 == 4,
              "family ids are spelled out in the names of the per-family entry points");

PDEGPU_W2_FAMILY(1)
#ifdef W2_PROBE          // the probe counters live in this translation unit: a probe build keeps every family here
PDEGPU_W2_FAMILY(0)
PDEGPU_W2_FAMILY(2)
PDEGPU_W2_FAMILY(3)
PDEGPU_W2_FAMILY(4)
#endif

#ifdef W2_PROBE
extern "C" int pdegpu_debug_w2_probe(unsigned long long *out16)
{
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out16, g_w2_probe, sizeof(g_w2_probe)) != cudaSuccess) return -1;
    unsigned long long z[16] = {0};
    return cudaMemcpyToSymbol(g_w2_probe, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
#endif

// entry used by sweeps_window.cu: p is complete except for the geometry
int window2_pass(pdegpu_ctx *ctx, int family, int dir, void *params, int M, int batch)
{
    switch (family) {
    case PDEGPU_FLOW_ELIN4: return window2_family_0(ctx, dir, params, M, batch);
    case PDEGPU_FLOW_LLIN4: return window2_family_1(ctx, dir, params, M, batch);
    case PDEGPU_FLOW_LLIN8: return window2_family_2(ctx, dir, params, M, batch);
    case PDEGPU_DISP_LLIN4: return window2_family_3(ctx, dir, params, M, batch);
    case PDEGPU_PDE4:       return window2_family_4(ctx, dir, params, M, batch);
    default: return PDEGPU_ERR_UNSUPPORTED;
    }
}
