/*
 * pdegpu.h -- C ABI of libpdegpu, the B200-native (sm_100a) replacement for the iterative
 * variational solver core of JediZ/PDE-based-image-processing (reference: mex/source/library/).
 *
 * Scope = SURVEY.md section 8: relaxation sweeps (point red-black SOR / zebra line SOR),
 * residual / LHS operators, bilinear warp, Simoncelli derivatives, diffusion weights.
 * Every entry point names the reference interface it replaces (file:line under /root/reference).
 *
 * Conventions (identical to the reference's MEX gateways):
 *   - all arrays are fp32, dense, COLUMN-MAJOR (Matlab): element (i,j,k) of an
 *     nrows x ncols x nframes array lives at  k*nrows*ncols + j*nrows + i
 *     (reference opticalflowSolvers.c:81).
 *   - `iter`, `omega` are passed as floats exactly as the gateways receive them; `iter`
 *     is truncated to int like the reference does ((int)Mparam.iter, opticalflowSolvers.c:74).
 *   - `solver`: 1 = point-wise Gauss-Seidel SOR, 2 = alternating line relaxation (ALR).
 *     The reference runs both in lexicographic order, which is inherently serial; libpdegpu
 *     runs the SAME fixed-point iteration in parallel orderings: solver 1 -> red-black
 *     (4-colour for 8-neighbour stencils), solver 2 -> zebra (even lines, then odd lines).
 *     Same linear system, same boundary model, same relaxation => same converged solution;
 *     iterates differ at finite `iter` (BASELINE.json north_star: "agree at convergence").
 *   - NaN in a data term (Cu/Du/TRACE) means "no data term at this pixel" exactly as in the
 *     reference (opticalflowSolvers.c:118-149, disparitySolvers.c:96-112, pdeSolvers.c:101-114).
 *   - No CPU fallback exists. Every function returns PDEGPU_OK (0) or a negative pdegpu_status;
 *     pdegpu_last_error() gives the message. Nothing throws, nothing prints.
 *
 * Two families of entry points:
 *   pdegpu_<name>(ctx, host pointers...)      drop-in for one MEX call: H2D, kernels, D2H.
 *   pdegpu_dev_<name>(ctx, device pointers..) device-resident, asynchronous on the context's
 *                                             stream, with a `batch` of independent problems;
 *                                             used by fused pipelines, benchmarks, multi-GPU.
 *
 * Threading: one context per (host thread, GPU); a context must not be shared between threads.
 */
#ifndef PDEGPU_H
#define PDEGPU_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDEGPU_VERSION_MAJOR 0
#define PDEGPU_VERSION_MINOR 1

typedef enum pdegpu_status {
    PDEGPU_OK              =  0,
    PDEGPU_ERR_ARG         = -1,   /* null pointer, bad solver id, ...                       */
    PDEGPU_ERR_SHAPE       = -2,   /* unsupported shape (e.g. nrows < 3 for a solver)         */
    PDEGPU_ERR_CUDA        = -3,   /* CUDA runtime error, see pdegpu_last_error()             */
    PDEGPU_ERR_NOMEM       = -4,   /* device or pinned-host allocation failed                 */
    PDEGPU_ERR_NODEVICE    = -5,   /* no usable sm_100 GPU                                    */
    PDEGPU_ERR_UNSUPPORTED = -6
} pdegpu_status;

typedef struct pdegpu_ctx pdegpu_ctx;

/* ------------------------------------------------------------------------------------------
 * Context
 * ------------------------------------------------------------------------------------------ */
int         pdegpu_device_count(void);                       /* number of visible CUDA devices (0 if none) */
int         pdegpu_init(int device, pdegpu_ctx **ctx);       /* create a context + stream on `device`    */
void        pdegpu_free(pdegpu_ctx *ctx);
const char *pdegpu_last_error(const pdegpu_ctx *ctx);        /* ctx may be NULL: last error of pdegpu_init */
const char *pdegpu_version(void);
int         pdegpu_sync(pdegpu_ctx *ctx);                    /* wait for the context's stream           */
void       *pdegpu_stream(pdegpu_ctx *ctx);                  /* the cudaStream_t, as void*              */
/* Number of libpdegpu kernels launched through this context since creation (bench evidence). */
unsigned long long pdegpu_launch_count(const pdegpu_ctx *ctx);
/* Per-launch timing for benchmarks: while enabled, every kernel launch is bracketed by a CUDA event
 * pair on the context's stream. pdegpu_profile_report() synchronises and writes a JSON array
 * [{"kernel","launches","ms_total","bytes_total"}] aggregated by kernel name; bytes_total is the
 * ALGORITHMIC byte count of those launches, so bytes_total/ms_total is the roofline
 * figure of that kernel. */
int         pdegpu_profile_enable(pdegpu_ctx *ctx, int on);
int         pdegpu_profile_report(pdegpu_ctx *ctx, char *buf, size_t buflen);
/* Kernel generation: 0 = "simple" global-memory kernels (kept as an in-library cross-check),
 * 1 = streaming register/shared-memory kernels (default). */
int         pdegpu_set_kernel_path(pdegpu_ctx *ctx, int path);
/* Order in which the sweeps visit the unknowns: solver 2 (alternating line relaxation) the lines of a direction, solver 1
 * (point relaxation) the pixels.
 *   PDEGPU_ORDER_FAST: zebra -- even lines, then odd lines; every line of a colour in parallel. Same fixed point as the
 *     reference, the HBM-bound kernel of the throughput numbers; its iterate after a FEW sweeps differs from the
 *     reference's (information travels two lines per sweep instead of across the image).
 *   PDEGPU_ORDER_REFERENCE: the reference's own order -- line j sees the new line j-1 and the old line j+1
 *     (GS_ALR_SOR_*: opticalflowSolvers.c:196,690,1677; disparitySolvers.c:154; pdeSolvers.c:277,344). Iterates agree
 *     with the reference sweep by sweep (1e-5 of the field's range, tests/test_gpu_reference_order.py), so the unchanged
 *     .m drivers give the reference's flow at the reference's iteration counts. Serial from line to line: the
 *     parallelism is the batch (one CTA per problem, several per SM) and the lanes inside a line. Solver 1 in this order
 *     is the reference's lexicographic point Gauss-Seidel (for j, for i, in place), run as a wavefront over the
 *     anti-diagonals of a problem.
 *   PDEGPU_ORDER_AUTO (default): REFERENCE for the early-linearisation flow family (PDEGPU_FLOW_ELIN4, i.e. the
 *     Horn-Schunck and FMG drivers, which solve each level ONCE and never re-warp: the iterate after `iter` sweeps is
 *     their result, and at their default iter = 20 / 4 the zebra iterate is a measurably different flow -- Yosemite
 *     FMG: 0.69 px against the reference's 0.21 px average end-point error); FAST for every other family (their
 *     drivers re-warp and re-linearise, both orders recover the ground truth equally well at the default settings).
 * The environment variable PDEGPU_ORDER = reference | fast | auto sets the initial order of a context; the MEX gateways
 * (which own their context) follow it at every call. FAST for solver 1 is the red-black (8-neighbour: four-colour)
 * order; AUTO never changes solver 1 (no driver uses it by default). Applies to
 * pdegpu_dev_relax and everything built on it (host-pointer entry points, driver pipelines). */
#define PDEGPU_ORDER_FAST      0
#define PDEGPU_ORDER_REFERENCE 1
#define PDEGPU_ORDER_AUTO      2
int         pdegpu_set_sweep_order(pdegpu_ctx *ctx, int order);
int         pdegpu_get_sweep_order(const pdegpu_ctx *ctx);
int         pdegpu_order_from_env(void);                    /* what PDEGPU_ORDER says right now (AUTO if unset) */

/* Device / pinned memory owned by the library (for the pdegpu_dev_* entry points). */
int         pdegpu_malloc(pdegpu_ctx *ctx, void **dptr, size_t bytes);
int         pdegpu_free_mem(pdegpu_ctx *ctx, void *dptr);
int         pdegpu_host_alloc(pdegpu_ctx *ctx, void **hptr, size_t bytes);   /* pinned */
int         pdegpu_host_free(pdegpu_ctx *ctx, void *hptr);
int         pdegpu_upload(pdegpu_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);   /* async */
int         pdegpu_download(pdegpu_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes); /* async */

/* ------------------------------------------------------------------------------------------
 * Host-pointer, drop-in entry points: one per hot-path MEX gateway.
 * Outputs are written completely (same contents the gateway's plhs[] would hold).
 * ------------------------------------------------------------------------------------------ */

/* [U V (RU RV)] = Oflow_sor_elin4_2d(U,V,M,Cu,Cv,Du,Dv,wW,wN,wE,wS,iter,omega,solver)
 * replaces mexFunction in mex/source/Oflow_sor_elin4_2d.c:64-352, i.e.
 * GS_SOR_elin4_2d (opticalflowSolvers.c:41) / GS_ALR_SOR_elin4_2d (:196) and
 * Residuals_elin4_2d (:269).
 * U_out,V_out: nrows x ncols. iter<=0 -> U_out,V_out are ZERO (gateway never copies, :341-346).
 * RU,RV (may be NULL): nrows x ncols x nframes residuals of the INPUT U,V (:350), where
 * M,Cu,Cv,Du,Dv may carry nframes>1 channels; the sweep itself uses channel 0 only. */
int pdegpu_oflow_sor_elin4_2d(pdegpu_ctx *ctx,
        float *U_out, float *V_out, float *RU, float *RV,
        const float *U, const float *V, const float *M,
        const float *Cu, const float *Cv, const float *Du, const float *Dv,
        const float *wW, const float *wN, const float *wE, const float *wS,
        int nrows, int ncols, int nframes, float iter, float omega, int solver);

/* [dU dV (RU RV)] = Oflow_sor_llin4_2d(U,V,dU,dV,M,Cu,Cv,Du,Dv,wW,wN,wE,wS,iter,omega,solver)
 * replaces mex/source/Oflow_sor_llin4_2d.c (GS_SOR_llin4_2d opticalflowSolvers.c:504,
 * GS_ALR_SOR_llin4_2d :690, Residuals_llin4_2d :766). Residuals are of the INPUT dU,dV. */
int pdegpu_oflow_sor_llin4_2d(pdegpu_ctx *ctx,
        float *dU_out, float *dV_out, float *RU, float *RV,
        const float *U, const float *V, const float *dU, const float *dV, const float *M,
        const float *Cu, const float *Cv, const float *Du, const float *Dv,
        const float *wW, const float *wN, const float *wE, const float *wS,
        int nrows, int ncols, int nframes, float iter, float omega, int solver);

/* Batched, pipelined form of the two gateways above (no reference counterpart: the reference solves one system per MEX
 * call, Oflow_sor_llin4_2d.c:355-361; a caller with many frame pairs -- BASELINE configs[4] -- loops over them). `batch`
 * independent systems per call, system b of EVERY array at element offset b*nrows*ncols, channel 0 only, no residual
 * outputs. The systems travel in chunks on parallel streams, so the upload of one chunk, the sweeps of the previous
 * one and the download of the one before overlap: the call runs at PCIe speed instead of copy + sweep + copy.
 * Results are bitwise those of `batch` single calls. PDEGPU_HOST_CHUNK / PDEGPU_HOST_LANES override the chunking. */
int pdegpu_oflow_sor_elin4_2d_batch(pdegpu_ctx *ctx, float *U_out, float *V_out,
        const float *U, const float *V, const float *M,
        const float *Cu, const float *Cv, const float *Du, const float *Dv,
        const float *wW, const float *wN, const float *wE, const float *wS,
        int nrows, int ncols, int batch, float iter, float omega, int solver);
int pdegpu_oflow_sor_llin4_2d_batch(pdegpu_ctx *ctx, float *dU_out, float *dV_out,
        const float *U, const float *V, const float *dU, const float *dV, const float *M,
        const float *Cu, const float *Cv, const float *Du, const float *Dv,
        const float *wW, const float *wN, const float *wE, const float *wS,
        int nrows, int ncols, int batch, float iter, float omega, int solver);

/* [dU dV] = Oflow_sor_llin8_2d(U,V,dU,dV,M,Cu,Cv,Du,Dv,wW,wNW,wN,wNE,wE,wSE,wS,wSW,iter,omega,solver)
 * replaces mex/source/Oflow_sor_llin8_2d.c (GS_SOR_llin8_2d opticalflowSolvers.c:1487 -- which
 * ignores the diagonal weights, SURVEY Q6 -- and GS_ALR_SOR_llin8_2d :1677). */
int pdegpu_oflow_sor_llin8_2d(pdegpu_ctx *ctx,
        float *dU_out, float *dV_out,
        const float *U, const float *V, const float *dU, const float *dV, const float *M,
        const float *Cu, const float *Cv, const float *Du, const float *Dv,
        const float *wW, const float *wNW, const float *wN, const float *wNE,
        const float *wE, const float *wSE, const float *wS, const float *wSW,
        int nrows, int ncols, float iter, float omega, int solver);

/* [AU AV] = Oflow_lhs_elin4_2d(U,V,M,Du,Dv,wW,wN,wE,wS)
 * replaces mex/source/Oflow_lhs_elin4_2d.c (LHS_elin4_2d opticalflowSolvers.c:387).
 * M,Du,Dv: nrows x ncols x nframes; AU,AV the same. */
int pdegpu_oflow_lhs_elin4_2d(pdegpu_ctx *ctx, float *AU, float *AV,
        const float *U, const float *V, const float *M, const float *Du, const float *Dv,
        const float *wW, const float *wN, const float *wE, const float *wS,
        int nrows, int ncols, int nframes);

/* [AU AV] = Oflow_lhs_llin4_2d(U,V,dU,dV,M,Du,Dv,wW,wN,wE,wS)
 * replaces mex/source/Oflow_lhs_llin4_2d.c (LHS_llin4_2d opticalflowSolvers.c:923), including its
 * border-fill defect (north border row of AV is left 0 for columns 1..ncols-2). */
int pdegpu_oflow_lhs_llin4_2d(pdegpu_ctx *ctx, float *AU, float *AV,
        const float *U, const float *V, const float *dU, const float *dV,
        const float *M, const float *Du, const float *Dv,
        const float *wW, const float *wN, const float *wE, const float *wS,
        int nrows, int ncols, int nframes);

/* [dU (RU)] = Disp_sor_llin4_2d(U,dU,Cu,Du,wW,wN,wE,wS,iter,omega,solver)
 * replaces mex/source/Disp_sor_llin4_2d.c (GS_SOR_llin4_2d disparitySolvers.c:41,
 * GS_ALR_SOR_llin4_2d :154). The reference allocates RU but never fills it
 * (Disp_sor_llin4_2d.c:251-269): RU, when requested, is all zeros. */
int pdegpu_disp_sor_llin4_2d(pdegpu_ctx *ctx, float *dU_out, float *RU,
        const float *U, const float *dU, const float *Cu, const float *Du,
        const float *wW, const float *wN, const float *wE, const float *wS,
        int nrows, int ncols, float iter, float omega, int solver);

/* [dU0 dU1] = Disp_sor_llin_sym4_2d(U0,dU0,Cu0,Du0,wW0,wN0,wE0,wS0,U1,dU1,...,iter,omega,solver)
 * replaces mex/source/Disp_sor_llin_sym4_2d.c (GS_SOR_llinsym4_2d disparitySolvers.c:301,
 * GS_ALR_SOR_llinsym4_2d :462): two independent scalar systems relaxed side by side. */
int pdegpu_disp_sor_llin_sym4_2d(pdegpu_ctx *ctx, float *dU0_out, float *dU1_out,
        const float *U0, const float *dU0, const float *Cu0, const float *Du0,
        const float *wW0, const float *wN0, const float *wE0, const float *wS0,
        const float *U1, const float *dU1, const float *Cu1, const float *Du1,
        const float *wW1, const float *wN1, const float *wE1, const float *wS1,
        int nrows, int ncols, float iter, float omega, int solver);

/* X = PDEsolver4(X,TRACE,B,wW,wN,wE,wS,iter,omega,solver)
 * replaces mex/source/PDEsolver4.c (GS_SOR_4_2d pdeSolvers.c:44, GS_ALR_SOR_4_2d :277).
 * All arrays nrows x ncols x nframes; frames are independent systems. */
int pdegpu_pdesolver4(pdegpu_ctx *ctx, float *X_out,
        const float *X, const float *TRACE, const float *B,
        const float *wW, const float *wN, const float *wE, const float *wS,
        int nrows, int ncols, int nframes, float iter, float omega, int solver);

/* X = PDEsolver8(X,TRACE,B,wW,wNW,wN,wNE,wE,wSE,wS,wSW,iter,omega,solver)
 * replaces mex/source/PDEsolver8.c (GS_SOR_8_2d pdeSolvers.c:153, GS_ALR_SOR_8_2d :344).
 * Reference quirks kept: solver 2 performs ONE iteration whatever `iter` is (pdeSolvers.c:362,
 * SURVEY Q4), relaxes interior lines only (:1155,:1290) and its NaN-TRACE diagonal counts wNW
 * twice and wNE never (:1179). */
int pdegpu_pdesolver8(pdegpu_ctx *ctx, float *X_out,
        const float *X, const float *TRACE, const float *B,
        const float *wW, const float *wNW, const float *wN, const float *wNE,
        const float *wE, const float *wSE, const float *wS, const float *wSW,
        int nrows, int ncols, int nframes, float iter, float omega, int solver);

/* Iout = BilinInterp_2d(Iin,X,Y)
 * replaces mex/source/BilinInterp_2d.c (bilinInterp2 imageInterpolation.c:44).
 * (X,Y) are 1-based Matlab coordinates, X along columns, Y along rows. `oob_value` is written
 * where the look-up falls outside the image: the reference's 5th parameter, which its own
 * gateway forgets to pass -- the libpdegpu gateway passes NaN, the author's intent. */
int pdegpu_bilin_interp_2d(pdegpu_ctx *ctx, float *Iout,
        const float *Iin, const float *X, const float *Y,
        int nrows, int ncols, int nframes, float oob_value);

/* [Idt Idx Idy] = FstDerivatives5(It0,It1)
 * replaces mex/source/FstDerivatives5.c (fstSimoncelli_c imageDerivatives.c:309, 5-tap path). */
int pdegpu_fst_derivatives5(pdegpu_ctx *ctx, float *Idt, float *Idx, float *Idy,
        const float *It0, const float *It1, int nrows, int ncols, int nframes);

/* [Idxt Idyt Idxx Idyy Idxy] = SndDerivatives5(It0,It1)
 * replaces mex/source/SndDerivatives5.c (sndSimoncelli_c imageDerivatives.c:391). */
int pdegpu_snd_derivatives5(pdegpu_ctx *ctx, float *Idxt, float *Idyt, float *Idxx, float *Idyy, float *Idxy,
        const float *It0, const float *It1, int nrows, int ncols, int nframes);

/* [wW wN wE wS] = DdiffWeights(D,eps)
 * replaces mex/source/DdiffWeights.c (diffWeights6_2D_c imageDiffusionWeights.c:341).
 * D: nrows x ncols x nframes; outputs: nrows x ncols (max over frames), outward edges 0. */
int pdegpu_ddiff_weights(pdegpu_ctx *ctx, float *wW, float *wN, float *wE, float *wS,
        const float *D, int nrows, int ncols, int nframes, float eps);

/* ------------------------------------------------------------------------------------------
 * Device-pointer entry points (asynchronous on pdegpu_stream(ctx)).
 * `batch` independent problems of identical shape; problem b of every array starts at
 * ptr + b*batch_stride (elements). For multi-frame arrays of ONE problem use the host API or
 * set batch = nframes where frames are independent (PDE solvers).
 * ------------------------------------------------------------------------------------------ */

typedef enum pdegpu_family {
    PDEGPU_FLOW_ELIN4 = 0,   /* unknowns U,V                       (opticalflowSolvers.c:41,196)   */
    PDEGPU_FLOW_LLIN4 = 1,   /* unknowns dU,dV; fixed U,V          (:504,:690)                      */
    PDEGPU_FLOW_LLIN8 = 2,   /* 8-neighbour late-lin               (:1487,:1677)                    */
    PDEGPU_DISP_LLIN4 = 3,   /* unknown dU; fixed U                (disparitySolvers.c:41,154)      */
    PDEGPU_PDE4       = 4,   /* unknown X; TRACE,B                 (pdeSolvers.c:44,277)            */
    PDEGPU_PDE8       = 5    /* 8-neighbour                        (pdeSolvers.c:153,344)           */
} pdegpu_family;

/* One linear system family instance in device memory. Unused pointers are NULL.
 *   x[0],x[1] : unknowns, relaxed IN PLACE (U,V | dU,dV | dU | X)
 *   x0[0],x0[1]: fixed fields of the late-linearisation families (U,V | U)
 *   m          : coupling term M (flow families)
 *   c[0],c[1]  : Cu,Cv | Cu | B
 *   d[0],d[1]  : Du,Dv | Du | TRACE
 *   w[8]       : wW,wN,wE,wS,wNW,wNE,wSE,wSW  */
typedef struct pdegpu_system {
    int family;
    int nrows, ncols, batch;
    long long batch_stride;
    float *x[2];
    const float *x0[2];
    const float *m;
    const float *c[2];
    const float *d[2];
    const float *w[8];
} pdegpu_system;

/* Run `iter` relaxation iterations of `solver` (1 point red-black, 2 zebra line) in place. */
int pdegpu_dev_relax(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega, int solver);

/* r = b - A x (elin4 / llin4 flow families), Residuals_* semantics incl. border replication.
 * c,d,m may have `nframes` channels per problem (channel stride nrows*ncols); x,x0,w have one. */
int pdegpu_dev_residual(pdegpu_ctx *ctx, const pdegpu_system *sys, int nframes, float *RU, float *RV);
/* A x (LHS_* semantics). */
int pdegpu_dev_lhs(pdegpu_ctx *ctx, const pdegpu_system *sys, int nframes, float *AU, float *AV);

int pdegpu_dev_bilin_interp_2d(pdegpu_ctx *ctx, float *Iout, const float *Iin, const float *X, const float *Y,
        int nrows, int ncols, int nframes, float oob_value);
int pdegpu_dev_fst_derivatives5(pdegpu_ctx *ctx, float *Idt, float *Idx, float *Idy,
        const float *It0, const float *It1, int nrows, int ncols, int nframes);
int pdegpu_dev_snd_derivatives5(pdegpu_ctx *ctx, float *Idxt, float *Idyt, float *Idxx, float *Idyy, float *Idxy,
        const float *It0, const float *It1, int nrows, int ncols, int nframes);
int pdegpu_dev_ddiff_weights(pdegpu_ctx *ctx, float *wW, float *wN, float *wE, float *wS,
        const float *D, int nrows, int ncols, int nframes, float eps);

/* ------------------------------------------------------------------------------------------
 * Driver-side stencils: what the reference's Matlab drivers compute between
 * MEX calls, as device-resident entry points so that a pyramid level needs no host round trip.
 * Device pointers, asynchronous on the context's stream. Arrays are column-major nrows x ncols
 * (x channels); channel c of a stack starts at c*nrows*ncols.
 * ------------------------------------------------------------------------------------------ */

/* [wW wN wS wE] = OPdiffWeights(U, V)   matlab/optical_flow/FlowEminND_llin_2D_v10.m:389-433
 * (same function in FlowEminNDFASFMG_elin_2D_v10.m:469-514). Computed in double with the reference's
 * circshift wrap-around, stored as single (the cast the drivers apply at the MEX call). */
int pdegpu_dev_op_diff_weights(pdegpu_ctx *ctx, float *wW, float *wN, float *wS, float *wE,
        const float *U, const float *V, int nrows, int ncols, int batch, long long batch_stride);

/* Robust data weights + data-term assembly of the late-linearisation flow driver:
 * products FlowEminND_llin_2D_v10.m:235-258, gD1/gD2 :289-299, nansum over channels :323-327.
 * d1 = {I1dt, I1dx, I1dy} (channels1); d2 = {I2dt, I2dx, I2dy} or, with gradmag != 0,
 * {I2dxt, I2dyt, I2dxx, I2dyy, I2dxy} (channels2, 0 = no second term). out = {M, Cu, Cv, Du, Dv}. */
typedef struct pdegpu_llin_terms {
    int nrows, ncols, batch;
    int channels1, channels2, gradmag;
    float b1, b2, alpha;
    const float *d1[3];
    const float *d2[5];
    const float *dU, *dV;
    float *out[5];
    long long batch_stride1, batch_stride2, batch_stride;   /* of the d1 stacks, the d2 stacks, dU/dV/out */
} pdegpu_llin_terms;
int pdegpu_dev_llin_terms(pdegpu_ctx *ctx, const pdegpu_llin_terms *t);
/* One inner solve of the late-linearisation flow driver, FlowEminND_llin_2D_v10.m:278-348, in ONE call:
 * OPdiffWeights(U+dU, V+dV) (:321), the robust data weights and channel sums (:289-327, `t`; t->dU, t->dV and t->out are
 * ignored) and Oflow_sor_llin4_2d (:332-348) relaxing dU, dV in place. With solver 2 on the packed-line kernels the
 * weights and terms are computed inside the kernel that prepares the lines and are never written as arrays (north_star
 * subsystem 3: weight stencils fused into the sweep's producer); otherwise the steps run one after the other through
 * `work` (11 * batch * batch_stride floats of device memory). Results are identical either way. */
int pdegpu_dev_llin_solve(pdegpu_ctx *ctx, const pdegpu_llin_terms *t, const float *U, const float *V, float *dU, float *dV,
                          float *work, int iter, float omega, int solver);

/* gd and gd-weighted terms of the FMG early-linearisation smoother:
 * summed != 0: FlowEminNDFASFMG_elin_2D_v10.m:375-396 (gd = 1/(channels*alpha*sqrt(.)), terms summed over
 * channels -> out[k] has one channel); summed == 0: :421-440 (gd = 1/(alpha*sqrt(.)), per-channel terms).
 * der = {Idt, Idx, Idy, Idxt, Idyt, Idxx, Idyy, Idxy}, coef = {M, Cu, Cv, Du, Dv}; gd may be NULL. */
typedef struct pdegpu_elin_terms {
    int nrows, ncols, channels, summed;
    float b1, b2, alpha;
    const float *der[8];
    const float *coef[5];
    const float *U, *V;
    float *gd;
    float *out[5];
} pdegpu_elin_terms;
int pdegpu_dev_elin_terms(pdegpu_ctx *ctx, const pdegpu_elin_terms *t);

/* Data + symmetry terms of ONE view of the symmetric stereo driver:
 * matlab/disparity/DispEminND_llin_sym_2D.m:172-180, 197-210, 222-225.
 * d = {Idt, Idx, Idxt, Idyt, Idxx, Idxy} (channels). alpha_d/beta/srdiff are the driver's doubles. */
typedef struct pdegpu_disp_sym_terms {
    int nrows, ncols, channels;
    float b1, b2, alpha;
    double alpha_d, beta, srdiff;
    const float *d[6];
    const float *dU, *Udt, *Udx;
    float *CuG, *DuG;
} pdegpu_disp_sym_terms;
int pdegpu_dev_disp_sym_terms(pdegpu_ctx *ctx, const pdegpu_disp_sym_terms *t);

/* f = (R + A)./gd   FAS coarse right-hand side, FlowEminNDFASFMG_elin_2D_v10.m:250-251 */
int pdegpu_dev_fas_rhs(pdegpu_ctx *ctx, float *f, const float *R, const float *A, const float *gd, long long n);

/* out = imfilter(single(in)*prescale, h, 'replicate') (correlation; pass a flipped kernel for 'conv'),
 * then out(1:step:end, 1:step:end). h is kr x kc, column-major, odd sizes, at most 25 taps; double
 * accumulation. Covers the drivers' Gaussian smoothing, rgb2grad, the lpf pyramid
 * (FlowEminNDFASFMG_elin_2D_v10.m:98,107-110) and the full-weighting restriction (:199,212-217).
 * `planes` images, plane p at in + p*in_stride / out + p*out_stride. */
int pdegpu_dev_imfilter(pdegpu_ctx *ctx, float *out, const float *in, int nrows, int ncols, int planes,
        long long in_stride, long long out_stride, const double *h, int kr, int kc, int step, float prescale);

/* out = imresize(in, 'OutputSize', [out_rows out_cols], 'Method', 'bilinear') with Matlab's
 * antialiasing when shrinking (FlowEminND_llin_2D_v10.m:107-108,365-366). `tmp` holds
 * planes*max(in,out) pixels. scale_rows/scale_cols: the scale imresize would use (out/in, or the scalar
 * scale argument when the driver passes one). */
int pdegpu_dev_imresize_bilinear(pdegpu_ctx *ctx, float *out, float *tmp, const float *in,
        int in_rows, int in_cols, int out_rows, int out_cols, double scale_rows, double scale_cols,
        int antialias, int planes);

/* Same with imresize's default 'bicubic' kernel (the FMG and Horn-Schunck drivers up-sample the flow with it:
 * FlowEminNDFASFMG_elin_2D_v10.m:180-181, FlowEminHS_elin_2D_v10.m:189-190). */
int pdegpu_dev_imresize_bicubic(pdegpu_ctx *ctx, float *out, float *tmp, const float *in,
        int in_rows, int in_cols, int out_rows, int out_cols, double scale_rows, double scale_cols,
        int antialias, int planes);

/* out = medfilt2(in, [3 3], 'symmetric')  (FlowEminND_llin_2D_v10.m:354-355) */
int pdegpu_dev_medfilt3(pdegpu_ctx *ctx, float *out, const float *in, int nrows, int ncols, int planes, long long stride);

/* out = a*x + b*y in single (y may be NULL: out = a*x) */
int pdegpu_dev_axpby(pdegpu_ctx *ctx, float *out, float a, const float *x, float b, const float *y, long long n);

/* X = meshgrid columns + U, Y = meshgrid rows + V (1-based, single): the warp coordinates the drivers
 * hand to BilinInterp_2d (FlowEminND_llin_2D_v10.m:202,222). U/V may be NULL. */
int pdegpu_dev_warp_coords(pdegpu_ctx *ctx, float *X, float *Y, const float *U, const float *V,
        int nrows, int ncols, int batch, long long batch_stride);

/* [W NW N NE E SE S SW] = ADdiffWeights(D)   matlab/denoising/TVdenoise8.m:119-231 (Alvarez derivatives, the frame of
 * largest gradient per pixel, lambda = the `quantile` order statistic of the non-zero squared gradient norms,
 * found on the device). w[k] = single(scale * weight) -- the driver passes single(param.alpha*wX) to PDEsolver8
 * (:87-100); scale = 1 gives the plain weights. If TRACE and B are not NULL the TV data terms of :83-85 are
 * produced in the same pass, D being the current estimate Iout and Iin the noisy input (nframes each):
 *   PsiData = 1./sqrt((Iout-Iin).^2 + eps), TRACE = PsiData + scale*(sum of the 8 weights), B = PsiData.*Iin.
 * lambda_dev (optional): device double receiving lambda. All in double like the .m code. */
int pdegpu_dev_ad_diff_weights(pdegpu_ctx *ctx, float *const w[8], float *TRACE, float *B,
        const float *D, const float *Iin, int nrows, int ncols, int nframes, double quantile, double scale, double *lambda_dev);

/* ------------------------------------------------------------------------------------------
 * Whole driver, device resident: [U V] = FlowEminND_llin_2D_v10(Iin, channels, fstTerm, sndTerm, ...)
 * (matlab/optical_flow/FlowEminND_llin_2D_v10.m; BASELINE.json configs[1], the "640x480 flows/s" metric)
 * for a BATCH of image pairs. I0, I1: nrows x ncols x channels per pair, values 0..255 as the driver
 * expects (it divides by 255); pair b starts at b*nrows*ncols*channels. U, V: nrows x ncols per pair.
 * The toolbox steps (imresize, imfilter, medfilt2) follow their documented behaviour (parity unpinned,
 * DESIGN.md); the optional spatial a-priori inputs (param.Us/Vs) are not restated.
 * pdegpu_dev_*: device pointers, asynchronous; pdegpu_flow_llin_2d: host pointers, H2D + D2H + sync.
 * ------------------------------------------------------------------------------------------ */
typedef struct pdegpu_flow_llin_params {
    double alpha, omega, b1, b2, scl_factor;   /* 0.0420, 1.9, 1.4843, 0.2915, 0.75 (:52-62) */
    int firstLoop, secondLoop, iter, solver;   /* 4, 4, 4, 2 */
    int fst_grad;                              /* fstTerm: 0 'rgb', 1 'grad' */
    int snd_term;                              /* sndTerm: 0 'none', 1 'rgb', 2 'gradmag' */
    int max_scales;                            /* 0 = until a side is <= 20 pixels (:116) */
    float oob_value;                           /* value of out-of-image warps (NaN) */
} pdegpu_flow_llin_params;
void pdegpu_flow_llin_default_params(pdegpu_flow_llin_params *p);
int pdegpu_dev_flow_llin_2d(pdegpu_ctx *ctx, float *U, float *V, const float *I0, const float *I1,
        int nrows, int ncols, int channels, int batch, const pdegpu_flow_llin_params *params);
int pdegpu_flow_llin_2d(pdegpu_ctx *ctx, float *U, float *V, const float *I0, const float *I1,
        int nrows, int ncols, int channels, int batch, const pdegpu_flow_llin_params *params);

/* [U V] = FlowEminNDFASFMG_elin_2D_v10(cat(3, I0, I1), channels)   matlab/optical_flow/FlowEminNDFASFMG_elin_2D_v10.m
 * (BASELINE configs[2]: early-linearisation flow with full multigrid): Gaussian + lpf pyramid (:97-118), derivative
 * stacks and constant terms per level (:123-149), FMG coarse to fine with one FAS V-cycle (cycle_index 1) or W-cycle (2)
 * per level (:158-273) around the smoother (:367-464), bicubic up-sampling of the flow between levels (:180-181).
 * I0, I1: nrows x ncols x channels per pair, values 0..255 (the driver does not rescale, :72); pair b starts at
 * b*nrows*ncols*channels; U, V: nrows x ncols per pair. Pairs of a batch run one after the other on the stream. */
typedef struct pdegpu_flow_fmg_params {
    double alpha, omega, b1, b2, scl_factor;   /* 0.035, 1.9, 0.03, 0.97, 0.5 (:53-59) */
    int firstLoop, iter, solver;               /* 4, 4, 2 */
    int cycle_index;                           /* 1 = V-cycle, 2 = W-cycle (:63-65) */
    int max_scales;                            /* 0 = until a side is <= 10 pixels (:114) */
} pdegpu_flow_fmg_params;
void pdegpu_flow_fmg_default_params(pdegpu_flow_fmg_params *p);
int pdegpu_dev_flow_fmg_2d(pdegpu_ctx *ctx, float *U, float *V, const float *I0, const float *I1,
        int nrows, int ncols, int channels, int batch, const pdegpu_flow_fmg_params *params);
int pdegpu_flow_fmg_2d(pdegpu_ctx *ctx, float *U, float *V, const float *I0, const float *I1,
        int nrows, int ncols, int channels, int batch, const pdegpu_flow_fmg_params *params);

/* [U V] = FlowEminHS_elin_2D_v10(cat(3, I0, I1), channels)   matlab/optical_flow/FlowEminHS_elin_2D_v10.m
 * (BASELINE configs[0]: Horn-Schunck): bilinear x scl_factor pyramid with 5x5 Gaussian smoothing (:96-115), per level
 * derivative stacks and quadratic terms summed over channels (:139-172), ONE Oflow_sor_elin4_2d solve with constant
 * weights alpha*channels (:127,174-188), 3x3 median + bicubic up-sampling of the flow (:193-196).
 * Same array conventions as pdegpu_flow_fmg_2d; values 0..255 (the driver divides by 255, :66). */
typedef struct pdegpu_flow_hs_params {
    double alpha, omega, b1, b2, scl_factor;   /* 0.2, 1.9, 0.25, 0.75, 0.75 (:48-53) */
    int iter, solver;                          /* 20, 2 */
    int max_scales;                            /* 0 = until a side is <= 20 pixels (:108) */
} pdegpu_flow_hs_params;
void pdegpu_flow_hs_default_params(pdegpu_flow_hs_params *p);
int pdegpu_dev_flow_hs_2d(pdegpu_ctx *ctx, float *U, float *V, const float *I0, const float *I1,
        int nrows, int ncols, int channels, int batch, const pdegpu_flow_hs_params *params);
int pdegpu_flow_hs_2d(pdegpu_ctx *ctx, float *U, float *V, const float *I0, const float *I1,
        int nrows, int ncols, int channels, int batch, const pdegpu_flow_hs_params *params);

/* U = DispEminND_llin_sym_2D(Il, Ir)   matlab/disparity/DispEminND_llin_sym_2D.m (BASELINE configs[3]: symmetric-
 * constraint stereo): two disparity fields coupled by a symmetry term; pyramid, cross warps of images and disparities,
 * derivatives, robust data + symmetry weights, DdiffWeights, Disp_sor_llin_sym4_2d, median, bilinear up-sampling.
 * Il, Ir: nrows x ncols x channels per pair, values 0..255; U: nrows x ncols x 2 per pair (U(:,:,1) left -> right,
 * U(:,:,2) right -> left). uint8_input != 0 keeps the pyramid in class uint8 (toolbox results rounded and saturated) as
 * for the imread images runme.m:17-28 passes. Pairs of a batch run one after the other on the stream. */
typedef struct pdegpu_disp_sym_params {
    double alpha, beta, omega, b1, b2, scl_factor;   /* 0.035, 0.4, 1.9, 0.25, 0.72, 0.75 (:50-60) */
    int firstLoop, secondLoop, iter, solver;         /* 3, 4, 4, 2 */
    int max_scales;                                  /* 0 = until a side is <= 10 pixels (:99) */
    int uint8_input;                                 /* 1 */
    float oob_value;                                 /* value of out-of-image warps (NaN) */
} pdegpu_disp_sym_params;
void pdegpu_disp_sym_default_params(pdegpu_disp_sym_params *p);
int pdegpu_dev_disp_sym_2d(pdegpu_ctx *ctx, float *U, const float *Il, const float *Ir,
        int nrows, int ncols, int channels, int batch, const pdegpu_disp_sym_params *params);
int pdegpu_disp_sym_2d(pdegpu_ctx *ctx, float *U, const float *Il, const float *Ir,
        int nrows, int ncols, int channels, int batch, const pdegpu_disp_sym_params *params);

/* Iout = TVdenoise8(I_in)   matlab/denoising/TVdenoise8.m (BASELINE configs[3]: 8-neighbour anisotropic TV denoising):
 * two-level pyramid, outer_iter+1 lagged-diffusivity steps per level of ADdiffWeights -> TRACE/B -> PDEsolver8,
 * bilinear up-sampling. I_in: nrows x ncols x nframes single (as runme.m:118,144 passes it). */
typedef struct pdegpu_tvdenoise8_params {
    double alpha, omega, scl_factor;      /* 500, 1.75, 0.75 (:36-44) */
    int outer_iter, inner_iter, solver;   /* 20, 4, 2 */
} pdegpu_tvdenoise8_params;
void pdegpu_tvdenoise8_default_params(pdegpu_tvdenoise8_params *p);
int pdegpu_dev_tvdenoise8(pdegpu_ctx *ctx, float *Iout, const float *Iin, int nrows, int ncols, int nframes, const pdegpu_tvdenoise8_params *params);
int pdegpu_tvdenoise8(pdegpu_ctx *ctx, float *Iout, const float *Iin, int nrows, int ncols, int nframes, const pdegpu_tvdenoise8_params *params);

/* ------------------------------------------------------------------------------------------
 * One very large image on several GPUs: column bands with halo exchange (BASELINE configs[4]).
 * The caller cuts the image along the slow axis (Matlab column index j) into contiguous bands with EVEN first columns,
 * one per GPU, keeps every field of a band with H = 2T halo columns per inner side, and alternates
 *     pdegpu_band_exchange(...);  pdegpu_dev_relax(ctx, &local_system, T, omega, 1);
 * A red-black sweep spoils two columns from each cut, so after T sweeps exactly the halo is stale and the owned
 * columns equal the single-GPU result bit for bit (pinned in tests/test_bands.py, tests/test_gpu_bands.py).
 * The exchange is two kernels on the context's stream -- peer stores into the neighbour's mailbox over NVLink, step
 * flags with system-scope release / acquire -- and no host synchronisation, NCCL call or event per step (csrc/band.cu).
 * Neighbours are reached through CUDA IPC handles (one process per GPU: export, pass the 64 bytes to the neighbours by
 * any means, connect) or directly (several contexts in one process: connect_local; peer access is enabled on demand).
 * A process that drives SEVERAL bands must enqueue the exchange of step s for all of them before it does anything that
 * may synchronise with the host on one of them (the first pdegpu_dev_relax of a context allocates its scratch and
 * synchronises): exchange all, then relax all. With one process per GPU there is no such constraint.
 * Only solver 1 splits this way: lines of the line solver along j cross the cuts.
 * ------------------------------------------------------------------------------------------ */
typedef struct pdegpu_band pdegpu_band;
int  pdegpu_band_create(pdegpu_ctx *ctx, int nrows, int halo_cols, int nunknowns, int has_left, int has_right, pdegpu_band **band);
int  pdegpu_band_export(pdegpu_band *band, void *handle64);                                   /* cudaIpcMemHandle_t, 64 bytes */
int  pdegpu_band_connect(pdegpu_band *band, const void *left_handle64, const void *right_handle64);   /* NULL where there is no neighbour */
int  pdegpu_band_connect_local(pdegpu_band *band, pdegpu_band *left, pdegpu_band *right);
/* unknowns[q]: the band's local array of unknown q (column-major, nrows rows, halo columns included, 16-byte aligned);
 * [own0, own1): the owned columns in local indices. Asynchronous on the context's stream. */
int  pdegpu_band_exchange(pdegpu_band *band, float *const unknowns[], int own0, int own1);
unsigned long long pdegpu_band_bytes_sent(const pdegpu_band *band);
void pdegpu_band_free(pdegpu_band *band);

#ifdef __cplusplus
}
#endif
#endif /* PDEGPU_H */
