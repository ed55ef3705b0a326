/*
 * [dU dV (RU RV)] = Oflow_sor_llin4_2d(U,V,dU,dV,M,Cu,Cv,Du,Dv,wW,wN,wE,wS,iter,omega,solver)
 *
 * libpdegpu gateway with the Matlab-visible signature of the reference's
 * mex/source/Oflow_sor_llin4_2d.c (16 inputs, nrhs check :115; outputs :328-358; residuals of the
 * INPUT dU,dV :384-385; iter<=0 leaves dU,dV zero :376-381).
 */
#include "gw_common.h"
#define GW "Oflow_sor_llin4_2d"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    static const char *names[13] = {"U_in", "V_in", "dU_in", "dV_in", "M", "Cu", "Cv", "Du", "Dv", "wW", "wN", "wE", "wS"};
    gw_arr a[13];
    float iter, omega, *o0, *o1, *RU = NULL, *RV = NULL;
    int k, solver;
    size_t n;
    pdegpu_ctx *ctx;

    if (nrhs != 16) gw_fail(GW, "parameter error: wrong number of input parameters!");
    for (k = 0; k < 13; k++) a[k] = gw_in(prhs[k], GW, names[k]);
    iter = gw_scalar(prhs[13], GW, "iter");
    omega = gw_scalar(prhs[14], GW, "omega");
    solver = (int)gw_scalar(prhs[15], GW, "solver");
    if (nlhs < 2) gw_fail(GW, "insufficient number of outputs. Outputs from this function are 'dU' and 'dV'");
    n = a[4].nrows * a[4].ncols;                            /* the library takes the size from M (:523-524) */
    for (k = 0; k < 13; k++) gw_need(&a[k], n, GW, names[k]);
    if (nlhs >= 4) for (k = 5; k < 9; k++) gw_need(&a[k], a[4].numel, GW, names[k]);
    o0 = gw_out_like(&plhs[0], prhs[2], GW, "dU_out");
    o1 = gw_out_like(&plhs[1], prhs[3], GW, "dV_out");
    if (nlhs >= 4) {
        RU = gw_out_like(&plhs[2], prhs[4], GW, "RU");
        RV = gw_out_like(&plhs[3], prhs[4], GW, "RV");
    }
    if (solver != 1 && solver != 2) gw_fail(GW, "no such solver");
    ctx = gw_ctx(GW);
    gw_check(ctx, pdegpu_oflow_sor_llin4_2d(ctx, o0, o1, RU, RV, a[0].p, a[1].p, a[2].p, a[3].p, a[4].p, a[5].p, a[6].p,
                                            a[7].p, a[8].p, a[9].p, a[10].p, a[11].p, a[12].p,
                                            (int)a[4].nrows, (int)a[4].ncols, nlhs >= 4 ? (int)a[4].nframes : 1,
                                            iter, omega, solver), GW);
}
