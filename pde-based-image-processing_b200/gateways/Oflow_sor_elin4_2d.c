/*
 * [U V (RU RV)] = Oflow_sor_elin4_2d(U,V,M,Cu,Cv,Du,Dv,wW,wN,wE,wS,iter,omega,solver)
 *
 * libpdegpu gateway with the Matlab-visible signature of the reference's
 * mex/source/Oflow_sor_elin4_2d.c:64-352 (14 inputs, all single; >=2 outputs; RU,RV when nlhs>=4,
 * shaped like M; residuals are those of the INPUT U,V; iter<=0 leaves U,V zero).
 * solver 1 -> red-black point SOR, solver 2 -> zebra line SOR (see include/pdegpu.h).
 */
#include "gw_common.h"
#define GW "Oflow_sor_elin4_2d"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    static const char *names[11] = {"U_in", "V_in", "M", "Cu", "Cv", "Du", "Dv", "wW", "wN", "wE", "wS"};
    gw_arr a[11];
    float iter, omega, *Uo, *Vo, *RU = NULL, *RV = NULL;
    int k, solver;
    size_t n;
    pdegpu_ctx *ctx;

    if (nrhs != 14) gw_fail(GW, "parameter error: wrong number of input parameters!");
    for (k = 0; k < 11; k++) a[k] = gw_in(prhs[k], GW, names[k]);
    iter = gw_scalar(prhs[11], GW, "iter");
    omega = gw_scalar(prhs[12], GW, "omega");
    solver = (int)gw_scalar(prhs[13], GW, "solver");
    if (nlhs < 2) gw_fail(GW, "insufficient number of outputs. Outputs from this function are 'U' and 'V'");
    n = a[0].nrows * a[0].ncols;
    for (k = 1; k < 11; k++) gw_need(&a[k], n, GW, names[k]);
    if (nlhs >= 4) for (k = 3; k < 7; k++) gw_need(&a[k], a[2].numel, GW, names[k]);
    Uo = gw_out_like(&plhs[0], prhs[1], GW, "U_out");      /* the reference shapes both like V_in (:299,:304) */
    Vo = gw_out_like(&plhs[1], prhs[1], GW, "V_out");
    if (nlhs >= 4) {
        RU = gw_out_like(&plhs[2], prhs[2], GW, "RU");
        RV = gw_out_like(&plhs[3], prhs[2], GW, "RV");
    }
    if (solver != 1 && solver != 2) gw_fail(GW, "no such solver");
    ctx = gw_ctx(GW);
    gw_check(ctx, pdegpu_oflow_sor_elin4_2d(ctx, Uo, Vo, RU, RV, a[0].p, a[1].p, a[2].p, a[3].p, a[4].p, a[5].p, a[6].p,
                                            a[7].p, a[8].p, a[9].p, a[10].p, (int)a[0].nrows, (int)a[0].ncols,
                                            nlhs >= 4 ? (int)a[2].nframes : 1, iter, omega, solver), GW);
}
