/*
 * [dU dV (RU RV)] = Oflow_sor_llin8_2d(U,V,dU,dV,M,Cu,Cv,Du,Dv,wW,wNW,wN,wNE,wE,wSE,wS,wSW,iter,omega,solver)
 *
 * libpdegpu gateway with the Matlab-visible signature of the reference's
 * mex/source/Oflow_sor_llin8_2d.c (20 inputs :128). With nlhs>=4 the reference creates RU,RV but
 * its residual call is commented out (:465-488): they stay zero, and so they do here.
 */
#include "gw_common.h"
#define GW "Oflow_sor_llin8_2d"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    static const char *names[17] = {"U_in", "V_in", "dU_in", "dV_in", "M", "Cu", "Cv", "Du", "Dv",
                                    "wW", "wNW", "wN", "wNE", "wE", "wSE", "wS", "wSW"};
    gw_arr a[17];
    float iter, omega, *o0, *o1;
    int k, solver;
    size_t n;
    pdegpu_ctx *ctx;

    if (nrhs != 20) gw_fail(GW, "parameter error: wrong number of input parameters!");
    for (k = 0; k < 17; k++) a[k] = gw_in(prhs[k], GW, names[k]);
    iter = gw_scalar(prhs[17], GW, "iter");
    omega = gw_scalar(prhs[18], GW, "omega");
    solver = (int)gw_scalar(prhs[19], GW, "solver");
    if (nlhs < 2) gw_fail(GW, "insufficient number of outputs. Outputs from this function are 'dU' and 'dV'");
    n = a[4].nrows * a[4].ncols;
    for (k = 0; k < 17; k++) gw_need(&a[k], n, GW, names[k]);
    o0 = gw_out_like(&plhs[0], prhs[2], GW, "dU_out");
    o1 = gw_out_like(&plhs[1], prhs[3], GW, "dV_out");
    if (nlhs >= 4) {
        gw_out_like(&plhs[2], prhs[4], GW, "RU");
        gw_out_like(&plhs[3], prhs[4], GW, "RV");
    }
    if (solver != 1 && solver != 2) gw_fail(GW, "no such solver");
    ctx = gw_ctx(GW);
    gw_check(ctx, pdegpu_oflow_sor_llin8_2d(ctx, o0, o1, a[0].p, a[1].p, a[2].p, a[3].p, a[4].p, a[5].p, a[6].p, a[7].p, a[8].p,
                                            a[9].p, a[10].p, a[11].p, a[12].p, a[13].p, a[14].p, a[15].p, a[16].p,
                                            (int)a[4].nrows, (int)a[4].ncols, iter, omega, solver), GW);
}
