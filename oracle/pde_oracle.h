/*
 * pde_oracle.h -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call this. It is the checker, never the product: libpdegpu has no CPU path.
 *
 * Plain C89-style loops over column-major fp32 arrays, one function per reference entry point,
 * written to follow the reference's evaluation order so that results are bit-identical to the
 * reference compiled with `gcc -O2` (verified by tests/test_oracle_vs_reference.py wherever
 * oracle/_ref is available, and against tests/golden/ fixtures everywhere else).
 * 8-neighbour line solvers are restated in natural summation order (agreement ~1e-6 relative).
 */
#ifndef PDE_ORACLE_H
#define PDE_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- flow (2 unknowns). late=0: early linearisation (unknowns X0,X1 = U,V; F0,F1 unused),
 *      late=1: late linearisation (unknowns X0,X1 = dU,dV; F0,F1 = U,V fixed). In place. ---- */
void orc_flow_gs(int late, float *X0, float *X1, const float *F0, const float *F1,
                 const float *M, const float *Cu, const float *Cv, const float *Du, const float *Dv,
                 const float *wW, const float *wN, const float *wE, const float *wS,
                 int nrows, int ncols, int iter, float omega);
void orc_flow_alr(int late, float *X0, float *X1, const float *F0, const float *F1,
                  const float *M, const float *Cu, const float *Cv, const float *Du, const float *Dv,
                  const float *wW, const float *wN, const float *wE, const float *wS,
                  int nrows, int ncols, int iter, float omega);
/* 8-neighbour late-linearisation ALR; w8 = {wW,wN,wE,wS,wNW,wNE,wSE,wSW} */
void orc_flow_alr8(float *dU, float *dV, const float *U, const float *V,
                   const float *M, const float *Cu, const float *Cv, const float *Du, const float *Dv,
                   const float *const w8[8], int nrows, int ncols, int iter, float omega);
/* residual (lhs=0) or A*x (lhs=1); M,C*,D* have nframes channels; quirks=1 reproduces the
 * border-fill defects of the late-linearisation versions */
void orc_flow_operator(int late, int lhs, int quirks, float *RU, float *RV,
                       const float *X0, const float *X1, const float *F0, const float *F1,
                       const float *M, const float *Cu, const float *Cv, const float *Du, const float *Dv,
                       const float *wW, const float *wN, const float *wE, const float *wS,
                       int nrows, int ncols, int nframes);

/* ---- disparity (1 unknown dU, fixed U) ---- */
void orc_disp_gs(float *dU, const float *U, const float *Cu, const float *Du,
                 const float *wW, const float *wN, const float *wE, const float *wS,
                 int nrows, int ncols, int iter, float omega);
void orc_disp_gs_sym(float *dU0, const float *U0, const float *Cu0, const float *Du0,
                     const float *wW0, const float *wN0, const float *wE0, const float *wS0,
                     float *dU1, const float *U1, const float *Cu1, const float *Du1,
                     const float *wW1, const float *wN1, const float *wE1, const float *wS1,
                     int nrows, int ncols, int iter, float omega);
void orc_disp_alr(float *dU, const float *U, const float *Cu, const float *Du,
                  const float *wW, const float *wN, const float *wE, const float *wS,
                  int nrows, int ncols, int iter, float omega);

/* ---- generic PDE: TRACE*x - sum w x_n = B; frames independent; w = {wW,wN,wE,wS[,wNW,wNE,wSE,wSW]} ---- */
void orc_pde_gs(int eight, float *X, const float *TRACE, const float *B, const float *const w[8],
                int nrows, int ncols, int nframes, int iter, float omega);
void orc_pde_alr(int eight, float *X, const float *TRACE, const float *B, const float *const w[8],
                 int nrows, int ncols, int nframes, int iter, float omega);

/* ---- streaming kernels ---- */
void orc_bilin(float *Iout, const float *Iin, const float *X, const float *Y,
               int nrows, int ncols, int nframes, float oob);
void orc_fst(float *Idt, float *Idx, float *Idy, const float *It0, const float *It1,
             int nrows, int ncols, int nframes);
void orc_snd(float *Idxt, float *Idyt, float *Idxx, float *Idyy, float *Idxy,
             const float *It0, const float *It1, int nrows, int ncols, int nframes);
void orc_ddiff(float *wW, float *wN, float *wE, float *wS, const float *D,
               int nrows, int ncols, int nframes, float eps);

#ifdef __cplusplus
}
#endif
#endif
