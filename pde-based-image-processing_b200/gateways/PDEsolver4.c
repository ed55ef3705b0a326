/*
 * X = PDEsolver4(X,TRACE,B,wW,wN,wE,wS,iter,omega,solver)
 *
 * libpdegpu gateway with the Matlab-visible signature of the reference's mex/source/PDEsolver4.c
 * (10 inputs :85; X may be 3-D, frames are independent :234-239). Solver ids other than 1 and 2
 * are rejected (the reference's `case 3` calls through an uninitialised pointer, :228).
 */
#include "gw_common.h"
#define GW "PDEsolver4"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    static const char *names[7] = {"Xold", "TRACE", "B", "wW", "wN", "wE", "wS"};
    gw_arr a[7];
    float iter, omega, *Xn;
    int k, solver;
    pdegpu_ctx *ctx;

    if (nrhs != 10) gw_fail(GW, "error: wrong number of input parameters!");
    for (k = 0; k < 7; k++) a[k] = gw_in(prhs[k], GW, names[k]);
    iter = gw_scalar(prhs[7], GW, "iter");
    omega = gw_scalar(prhs[8], GW, "omega");
    solver = (int)gw_scalar(prhs[9], GW, "solver");
    if (nlhs < 1) gw_fail(GW, "error insufficient number of outputs.");
    for (k = 1; k < 7; k++) gw_need(&a[k], a[0].nrows * a[0].ncols * a[0].nframes, GW, names[k]);
    Xn = gw_out_like(&plhs[0], prhs[0], GW, "Xnew");
    if (solver != 1 && solver != 2) gw_fail(GW, "error: no such solver");
    ctx = gw_ctx(GW);
    gw_check(ctx, pdegpu_pdesolver4(ctx, Xn, a[0].p, a[1].p, a[2].p, a[3].p, a[4].p, a[5].p, a[6].p,
                                    (int)a[0].nrows, (int)a[0].ncols, (int)a[0].nframes, iter, omega, solver), GW);
}
