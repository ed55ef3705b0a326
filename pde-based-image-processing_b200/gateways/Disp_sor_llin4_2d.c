/*
 * [dU (RU)] = Disp_sor_llin4_2d(U,dU,Cu,Du,wW,wN,wE,wS,iter,omega,solver)
 *
 * libpdegpu gateway with the Matlab-visible signature of the reference's
 * mex/source/Disp_sor_llin4_2d.c (11 inputs :94; RU created but never filled :251-269;
 * iter<=0 leaves dU zero :276-280).
 */
#include "gw_common.h"
#define GW "Disp_sor_llin4_2d"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    static const char *names[8] = {"U_in", "dU_in", "Cu", "Du", "wW", "wN", "wE", "wS"};
    gw_arr a[8];
    float iter, omega, *o0, *RU = NULL;
    int k, solver;
    size_t n;
    pdegpu_ctx *ctx;

    if (nrhs != 11) gw_fail(GW, "parameter error: wrong number of input parameters!");
    for (k = 0; k < 8; k++) a[k] = gw_in(prhs[k], GW, names[k]);
    iter = gw_scalar(prhs[8], GW, "iter");
    omega = gw_scalar(prhs[9], GW, "omega");
    solver = (int)gw_scalar(prhs[10], GW, "solver");
    if (nlhs < 1) gw_fail(GW, "insufficient number of outputs. Outputs from this function is 'dU'");
    n = a[2].nrows * a[2].ncols;                            /* size comes from Cu (disparitySolvers.c:54-55) */
    for (k = 0; k < 8; k++) gw_need(&a[k], n, GW, names[k]);
    o0 = gw_out_like(&plhs[0], prhs[1], GW, "dU_out");
    if (nlhs >= 2) RU = gw_out_like(&plhs[1], prhs[0], GW, "RU");
    if (solver != 1 && solver != 2) gw_fail(GW, "no such solver");
    ctx = gw_ctx(GW);
    gw_check(ctx, pdegpu_disp_sor_llin4_2d(ctx, o0, NULL, a[0].p, a[1].p, a[2].p, a[3].p, a[4].p, a[5].p, a[6].p, a[7].p,
                                           (int)a[2].nrows, (int)a[2].ncols, iter, omega, solver), GW);
    (void)RU;                                               /* stays zero, as in the reference */
}
