/*
 * X = PDEsolver8(X,TRACE,B,wW,wNW,wN,wNE,wE,wSE,wS,wSW,iter,omega,solver)
 *
 * libpdegpu gateway with the Matlab-visible signature of the reference's mex/source/PDEsolver8.c
 * (14 inputs :93). solver 2 performs exactly one line-relaxation iteration whatever `iter` is,
 * like GS_ALR_SOR_8_2d (pdeSolvers.c:362).
 */
#include "gw_common.h"
#define GW "PDEsolver8"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    static const char *names[11] = {"Xold", "TRACE", "B", "wW", "wNW", "wN", "wNE", "wE", "wSE", "wS", "wSW"};
    gw_arr a[11];
    float iter, omega, *Xn;
    int k, solver;
    pdegpu_ctx *ctx;

    if (nrhs != 14) gw_fail(GW, "error: wrong number of input parameters!");
    for (k = 0; k < 11; k++) a[k] = gw_in(prhs[k], GW, names[k]);
    iter = gw_scalar(prhs[11], GW, "iter");
    omega = gw_scalar(prhs[12], GW, "omega");
    solver = (int)gw_scalar(prhs[13], GW, "solver");
    if (nlhs < 1) gw_fail(GW, "error insufficient number of outputs.");
    for (k = 1; k < 11; k++) gw_need(&a[k], a[0].nrows * a[0].ncols * a[0].nframes, GW, names[k]);
    Xn = gw_out_like(&plhs[0], prhs[0], GW, "Xnew");
    if (solver != 1 && solver != 2) gw_fail(GW, "error: no such solver");
    ctx = gw_ctx(GW);
    gw_check(ctx, pdegpu_pdesolver8(ctx, Xn, a[0].p, a[1].p, a[2].p, a[3].p, a[4].p, a[5].p, a[6].p, a[7].p, a[8].p, a[9].p, a[10].p,
                                    (int)a[0].nrows, (int)a[0].ncols, (int)a[0].nframes, iter, omega, solver), GW);
}
