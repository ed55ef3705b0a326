// pdegpu_internal.cuh -- shared declarations of libpdegpu's translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#include "../../include/pdegpu.h"

// ------------------------------------------------------------------------------------------
// Context
// ------------------------------------------------------------------------------------------
struct pdegpu_ctx {
    int           device;
    int           sm_count;
    cudaStream_t  stream;
    // growable device arena used by the host-pointer entry points (one bump allocation per call)
    char         *arena;
    size_t        arena_bytes;
    size_t        arena_used;
    // growable scratch for the kernels themselves (Thomas coefficients, ping-pong fields)
    char         *scratch;
    size_t        scratch_bytes;
    unsigned long long launches;
    int           kernel_path;     // 0 simple, 1 streaming
    int           sweep_order;     // PDEGPU_ORDER_FAST / PDEGPU_ORDER_REFERENCE (solver 2)
    // optional per-launch timing (pdegpu_profile_*): one CUDA event pair per kernel launch
    int           prof_on;
    int           prof_n, prof_cap;
    struct pdegpu_prof_rec *prof;
    // workspace of the device-resident pipelines (pipeline.cu), grown on demand
    char         *work;
    size_t        work_bytes;
    // CUDA graphs of whole driver pipelines (pdegpu_graph_run): device pointers are baked in, so every reallocation of
    // arena / scratch / work bumps the epoch and the cached graphs die
    unsigned      graph_epoch;
    unsigned long long graph_tick;
    int           capturing;
    struct pdegpu_graph_entry *graphs;
    // LANES: child contexts (own stream, scratch, workspace) on which the pairs of a batch run side by side where a
    // driver pipeline is written for one pair (FMG, Horn-Schunck, symmetric stereo): see pdegpu_lanes_* below
    struct pdegpu_ctx *parent;
    struct pdegpu_ctx *lanes[128];  // 128 = the device's limit of concurrently resident grids
    int           nlanes;
    cudaEvent_t   ev_fork, ev_join;
    char          err[512];
};

// Pairs of a batch on parallel streams. pdegpu_lane_count: how many lanes a batch of `batch` pairs gets (1 = run on
// the context itself: while profiling, or with PDEGPU_LANES=1). prepare: the lanes exist, carry the context's current
// settings and own at least `work_bytes` of workspace each. fork / join: the lanes' streams wait for everything
// enqueued on the context's stream / the context's stream waits for the lanes (both legal inside a stream capture:
// the captured graph then has one branch per lane).
int pdegpu_lane_count(pdegpu_ctx *ctx, int batch, bool serial_order = false);   // serial_order: the line sweeps run in the reference's order
int pdegpu_lanes_prepare(pdegpu_ctx *ctx, int n, size_t work_bytes, const char *who);
int pdegpu_lanes_fork(pdegpu_ctx *ctx, int n);
int pdegpu_lanes_join(pdegpu_ctx *ctx, int n);

// Runs `body` (a sequence of launches on ctx->stream with no host synchronisation) and, from the second call with the
// same `key` on, replays it as a CUDA graph: the pipelines issue thousands of small dependent launches per call (4400
// per 1080p FMG pair), whose launch latency a graph removes. Falls back to running `body` directly when profiling is
// on, when PDEGPU_GRAPHS=0, or when the capture fails. `key` must cover every argument the launches depend on.
struct pdegpu_graph_body { int (*fn)(void *); void *arg; };
int pdegpu_graph_run(pdegpu_ctx *ctx, const void *key, size_t key_len, pdegpu_graph_body body);

struct pdegpu_prof_rec {
    const char  *name;
    double       bytes;            // algorithmic bytes of this launch, 0 if not a roofline kernel
    cudaEvent_t  e0, e1;
};

void pdegpu_prof_begin(pdegpu_ctx *ctx, const char *name, double bytes);
void pdegpu_prof_end(pdegpu_ctx *ctx);

int  pdegpu_set_error(pdegpu_ctx *ctx, int status, const char *fmt, ...);
int  pdegpu_check_cuda(pdegpu_ctx *ctx, cudaError_t e, const char *what);
// arena: reset at the start of a host-pointer call, bump-allocate 256-B aligned blocks
int  pdegpu_arena_reserve(pdegpu_ctx *ctx, size_t bytes);     // make sure capacity >= bytes (may realloc; only when empty)
void pdegpu_arena_reset(pdegpu_ctx *ctx);
void *pdegpu_arena_alloc(pdegpu_ctx *ctx, size_t bytes);      // nullptr if exhausted
int  pdegpu_scratch_reserve(pdegpu_ctx *ctx, size_t bytes);   // ctx->scratch valid for >= bytes afterwards
int  pdegpu_work_reserve(pdegpu_ctx *ctx, size_t bytes, const char *who);   // ctx->work valid for >= bytes afterwards

#define PDEGPU_CUDA_OK(ctx, call)                                                 \
    do {                                                                          \
        cudaError_t e__ = (call);                                                 \
        if (e__ != cudaSuccess) return pdegpu_check_cuda((ctx), e__, #call);      \
    } while (0)

#define PDEGPU_LAUNCH_CHECK(ctx, name)                                            \
    do {                                                                          \
        (ctx)->launches++;                                                        \
        if ((ctx)->prof_on) pdegpu_prof_end(ctx);                                 \
        cudaError_t e__ = cudaGetLastError();                                     \
        if (e__ != cudaSuccess) return pdegpu_check_cuda((ctx), e__, name);       \
    } while (0)

// put in front of a kernel launch: names it and states its algorithmic bytes for the profile
#define PDEGPU_PROF(ctx, name, bytes) do { if ((ctx)->prof_on) pdegpu_prof_begin((ctx), (name), (double)(bytes)); } while (0)

// ------------------------------------------------------------------------------------------
// Device-side view of one system (see pdegpu_system in pdegpu.h)
// ------------------------------------------------------------------------------------------
enum { W_W = 0, W_N = 1, W_E = 2, W_S = 3, W_NW = 4, W_NE = 5, W_SE = 6, W_SW = 7 };

struct SysView {
    float       *x[2];
    const float *x0[2];
    const float *m;
    const float *c[2];
    const float *d[2];
    const float *w[8];
    int          nrows, ncols;
    long long    bstride;
};

static inline SysView make_view(const pdegpu_system *s)
{
    SysView v;
    for (int k = 0; k < 2; k++) { v.x[k] = s->x[k]; v.x0[k] = s->x0[k]; v.c[k] = s->c[k]; v.d[k] = s->d[k]; }
    v.m = s->m;
    for (int k = 0; k < 8; k++) v.w[k] = s->w[k];
    v.nrows = s->nrows; v.ncols = s->ncols; v.bstride = s->batch_stride;
    return v;
}

__device__ __forceinline__ bool is_nan(float v) { return v != v; }

// family traits
template <int FAM> struct Fam;
template <> struct Fam<PDEGPU_FLOW_ELIN4> { static constexpr int NUNK = 2; static constexpr bool LATE = false, EIGHT = false, PDE = false; };
template <> struct Fam<PDEGPU_FLOW_LLIN4> { static constexpr int NUNK = 2; static constexpr bool LATE = true,  EIGHT = false, PDE = false; };
template <> struct Fam<PDEGPU_FLOW_LLIN8> { static constexpr int NUNK = 2; static constexpr bool LATE = true,  EIGHT = true,  PDE = false; };
template <> struct Fam<PDEGPU_DISP_LLIN4> { static constexpr int NUNK = 1; static constexpr bool LATE = true,  EIGHT = false, PDE = false; };
template <> struct Fam<PDEGPU_PDE4>       { static constexpr int NUNK = 1; static constexpr bool LATE = false, EIGHT = false, PDE = true;  };
template <> struct Fam<PDEGPU_PDE8>       { static constexpr int NUNK = 1; static constexpr bool LATE = false, EIGHT = true,  PDE = true;  };

// algorithmic HBM bytes per pixel per sweep, B1 of SURVEY.md section 8(d): 4 B x (streams read once + written once)
template <int FAM> constexpr double sweep_bytes()
{
    return FAM == PDEGPU_FLOW_ELIN4 ? 52.0 : FAM == PDEGPU_FLOW_LLIN4 ? 60.0 : FAM == PDEGPU_FLOW_LLIN8 ? 76.0
         : FAM == PDEGPU_DISP_LLIN4 ? 36.0 : FAM == PDEGPU_PDE4 ? 32.0 : 48.0;
}

// ------------------------------------------------------------------------------------------
// Kernel launchers implemented in the .cu files (all asynchronous on ctx->stream)
// ------------------------------------------------------------------------------------------
int relax_simple(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega, int solver);
int relax_llin_fused(pdegpu_ctx *ctx, const pdegpu_system *sys, const struct pdegpu_llin_terms *t, int iter, float omega, bool reference_order);   // sweeps_tline.cu
int relax_lexpoint(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega);               // solver 1 in the reference's lexicographic order (sweeps_simple.cu)
int relax_lexline(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega);                // solver 2 in the reference's line order (sweeps_tline.cu)
int relax_stream(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega, int solver);   // returns PDEGPU_ERR_UNSUPPORTED when it has no kernel for the case

int op_residual(pdegpu_ctx *ctx, const pdegpu_system *sys, int nframes, float *RU, float *RV, bool lhs);
int op_llin4_quirks(pdegpu_ctx *ctx, const pdegpu_system *sys, int nframes, float *RU, float *RV, bool lhs);
int op_bilin(pdegpu_ctx *ctx, float *Iout, const float *Iin, const float *X, const float *Y, int nrows, int ncols, int nframes, float oob);
int op_bilin_batch(pdegpu_ctx *ctx, float *Iout, const float *Iin, const float *X, const float *Y, int nrows, int ncols, int nframes, int batch, float oob);
int op_fst(pdegpu_ctx *ctx, float *Idt, float *Idx, float *Idy, const float *It0, const float *It1, int nrows, int ncols, int nframes);
int op_snd(pdegpu_ctx *ctx, float *Idxt, float *Idyt, float *Idxx, float *Idyy, float *Idxy, const float *It0, const float *It1, int nrows, int ncols, int nframes);
int op_ddiff(pdegpu_ctx *ctx, float *wW, float *wN, float *wE, float *wS, const float *D, int nrows, int ncols, int nframes, float eps);

// driver-side stencils (driver_ops.cu)
int op_opdiff(pdegpu_ctx *ctx, float *wW, float *wN, float *wS, float *wE, const float *U, const float *V, int nr, int nc, int batch, long long stride);
int op_opdiff_sum(pdegpu_ctx *ctx, float *wW, float *wN, float *wS, float *wE, const float *U, const float *V, const float *dU, const float *dV,
                  int nr, int nc, int batch, long long stride);   // weights of (U + dU, V + dV), sums formed on the fly
int op_llin_terms(pdegpu_ctx *ctx, const pdegpu_llin_terms *t);
int op_elin_terms(pdegpu_ctx *ctx, const pdegpu_elin_terms *t);
int op_disp_sym_terms(pdegpu_ctx *ctx, const pdegpu_disp_sym_terms *t);
int op_fas_rhs(pdegpu_ctx *ctx, float *f, const float *R, const float *A, const float *gd, long long n);
int op_imfilter(pdegpu_ctx *ctx, float *out, const float *in, int nr, int nc, int planes, long long istride, long long ostride,
                const double *h, int kr, int kc, int step, float prescale);
int op_imresize_dim(pdegpu_ctx *ctx, float *out, const float *in, int dim, int in_len, int out_len, int other, double scale,
                    int antialias, int planes, long long istride, long long ostride, int cubic = 0);
int op_medfilt3(pdegpu_ctx *ctx, float *out, const float *in, int nr, int nc, int planes, long long stride);
int op_axpby(pdegpu_ctx *ctx, float *out, float a, const float *x, float b, const float *y, long long n);
int op_warp_coords(pdegpu_ctx *ctx, float *X, float *Y, const float *U, const float *V, int nr, int nc, int batch, long long stride);
int op_axpby_div(pdegpu_ctx *ctx, float *out, const float *x, float d, long long n);
int op_ad_diff_weights(pdegpu_ctx *ctx, float *const w[8], float *TRACE, float *B, const float *D, const float *Iin, int nr, int nc, int frames, double quantile, double scale, double *lambda_dev, int wframes = 1);
int op_fmg_terms(pdegpu_ctx *ctx, float *const out[5], const float *const der[8], float b1, float b2, long long n);
int op_fmg_prescale(pdegpu_ctx *ctx, float *Ist, float *Idt, const float *I0, const float *I1, long long n, float div = 255.0f);
int op_channel_sum(pdegpu_ctx *ctx, float *out, const float *in, int channels, long long npix);
int op_fill(pdegpu_ctx *ctx, float *out, float v, long long n);
int op_interp_rows(pdegpu_ctx *ctx, float *out, const float *vals, const float *shift, int nr, int nc);
int op_round_uint8(pdegpu_ctx *ctx, float *x, long long n);
int imresize_2d(pdegpu_ctx *ctx, float *out, float *tmp, const float *in, int in_rows, int in_cols, int out_rows, int out_cols, double scale_rows, double scale_cols, int antialias, int planes, int cubic);
