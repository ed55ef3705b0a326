/*
 * [AU AV] = Oflow_lhs_llin4_2d(U,V,dU,dV,M,Du,Dv,wW,wN,wE,wS)
 *
 * libpdegpu gateway with the Matlab-visible signature of the reference's
 * mex/source/Oflow_lhs_llin4_2d.c (11 inputs :87; LHS_llin4_2d :259, border-fill defect kept).
 */
#include "gw_common.h"
#define GW "Oflow_lhs_llin4_2d"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    static const char *names[11] = {"U_in", "V_in", "dU_in", "dV_in", "M", "Du", "Dv", "wW", "wN", "wE", "wS"};
    gw_arr a[11];
    float *AU, *AV;
    int k;
    size_t n;
    pdegpu_ctx *ctx;

    if (nrhs != 11) gw_fail(GW, "parameter error: wrong number of input parameters!");
    for (k = 0; k < 11; k++) a[k] = gw_in(prhs[k], GW, names[k]);
    if (nlhs < 2) gw_fail(GW, "insufficient number of outputs. Outputs from this function are 'AU' and 'AV'");
    n = a[4].nrows * a[4].ncols;
    for (k = 0; k < 11; k++) gw_need(&a[k], n, GW, names[k]);
    gw_need(&a[5], a[4].numel, GW, "Du");
    gw_need(&a[6], a[4].numel, GW, "Dv");
    AU = gw_out_like(&plhs[0], prhs[4], GW, "AU");
    AV = gw_out_like(&plhs[1], prhs[4], GW, "AV");
    ctx = gw_ctx(GW);
    gw_check(ctx, pdegpu_oflow_lhs_llin4_2d(ctx, AU, AV, a[0].p, a[1].p, a[2].p, a[3].p, a[4].p, a[5].p, a[6].p,
                                            a[7].p, a[8].p, a[9].p, a[10].p,
                                            (int)a[4].nrows, (int)a[4].ncols, (int)a[4].nframes), GW);
}
