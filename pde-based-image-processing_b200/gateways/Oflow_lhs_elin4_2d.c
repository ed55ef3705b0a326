/*
 * [AU AV] = Oflow_lhs_elin4_2d(U,V,M,Du,Dv,wW,wN,wE,wS)
 *
 * libpdegpu gateway with the Matlab-visible signature of the reference's
 * mex/source/Oflow_lhs_elin4_2d.c (9 inputs :82; AU,AV shaped like M :205-227; LHS_elin4_2d :229).
 */
#include "gw_common.h"
#define GW "Oflow_lhs_elin4_2d"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    static const char *names[9] = {"U_in", "V_in", "M", "Du", "Dv", "wW", "wN", "wE", "wS"};
    gw_arr a[9];
    float *AU, *AV;
    int k;
    size_t n;
    pdegpu_ctx *ctx;

    if (nrhs != 9) gw_fail(GW, "parameter error: wrong number of input parameters!");
    for (k = 0; k < 9; k++) a[k] = gw_in(prhs[k], GW, names[k]);
    if (nlhs < 2) gw_fail(GW, "insufficient number of outputs. Outputs from this function are 'AU' and 'AV'");
    n = a[2].nrows * a[2].ncols;
    for (k = 0; k < 9; k++) gw_need(&a[k], n, GW, names[k]);
    gw_need(&a[3], a[2].numel, GW, "Du");
    gw_need(&a[4], a[2].numel, GW, "Dv");
    AU = gw_out_like(&plhs[0], prhs[2], GW, "AU");
    AV = gw_out_like(&plhs[1], prhs[2], GW, "AV");
    ctx = gw_ctx(GW);
    gw_check(ctx, pdegpu_oflow_lhs_elin4_2d(ctx, AU, AV, a[0].p, a[1].p, a[2].p, a[3].p, a[4].p, a[5].p, a[6].p, a[7].p, a[8].p,
                                            (int)a[2].nrows, (int)a[2].ncols, (int)a[2].nframes), GW);
}
