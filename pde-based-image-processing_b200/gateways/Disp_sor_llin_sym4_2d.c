/*
 * [dU0 dU1] = Disp_sor_llin_sym4_2d(U0,dU0,Cu0,Du0,wW0,wN0,wE0,wS0, U1,dU1,Cu1,Du1,wW1,wN1,wE1,wS1, iter,omega,solver)
 *
 * libpdegpu gateway with the Matlab-visible signature of the reference's
 * mex/source/Disp_sor_llin_sym4_2d.c (19 inputs :136; the initial guess is copied whatever `iter`
 * is :418-419).
 */
#include "gw_common.h"
#define GW "Disp_sor_llin_sym4_2d"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    static const char *names[16] = {"U_in0", "dU_in0", "Cu0", "Du0", "wW0", "wN0", "wE0", "wS0",
                                    "U_in1", "dU_in1", "Cu1", "Du1", "wW1", "wN1", "wE1", "wS1"};
    gw_arr a[16];
    float iter, omega, *o0, *o1;
    int k, solver;
    size_t n;
    pdegpu_ctx *ctx;

    if (nrhs != 19) gw_fail(GW, "parameter error: wrong number of input parameters!");
    for (k = 0; k < 16; k++) a[k] = gw_in(prhs[k], GW, names[k]);
    iter = gw_scalar(prhs[16], GW, "iter");
    omega = gw_scalar(prhs[17], GW, "omega");
    solver = (int)gw_scalar(prhs[18], GW, "solver");
    if (nlhs < 2) gw_fail(GW, "insufficient number of outputs. Outputs from this function are 'dU0' and 'dU1'");
    n = a[2].nrows * a[2].ncols;                            /* size comes from Cu0 (disparitySolvers.c:323-324) */
    for (k = 0; k < 16; k++) gw_need(&a[k], n, GW, names[k]);
    o0 = gw_out_like(&plhs[0], prhs[1], GW, "dU_out0");
    o1 = gw_out_like(&plhs[1], prhs[9], GW, "dU_out1");
    if (solver != 1 && solver != 2) gw_fail(GW, "no such solver");
    ctx = gw_ctx(GW);
    gw_check(ctx, pdegpu_disp_sor_llin_sym4_2d(ctx, o0, o1,
                                               a[0].p, a[1].p, a[2].p, a[3].p, a[4].p, a[5].p, a[6].p, a[7].p,
                                               a[8].p, a[9].p, a[10].p, a[11].p, a[12].p, a[13].p, a[14].p, a[15].p,
                                               (int)a[2].nrows, (int)a[2].ncols, iter, omega, solver), GW);
}
