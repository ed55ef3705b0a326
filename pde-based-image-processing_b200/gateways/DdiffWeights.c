/*
 * [wW wN wE wS] = DdiffWeights(D,eps)
 *
 * libpdegpu gateway with the Matlab-visible signature of the reference's mex/source/DdiffWeights.c
 * (2 inputs :69; four outputs shaped like D :99-135). For a multi-frame D only the first frame of
 * each output is written (the maximum over frames, imageDiffusionWeights.c:141-143); the other
 * frames stay zero, as in the reference.
 */
#include "gw_common.h"
#define GW "DdiffWeights"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    static const char *onames[4] = {"wW", "wN", "wE", "wS"};
    gw_arr D;
    float eps, *o[4];
    int k;
    pdegpu_ctx *ctx;

    if (nrhs != 2) gw_fail(GW, "parameter error: wrong number of input parameters!");
    D = gw_in(prhs[0], GW, "D");
    eps = gw_scalar(prhs[1], GW, "eps");
    if (nlhs < 4) gw_fail(GW, "error insufficient number of outputs. Outputs from this function are 'wW', 'wN', 'wE' and 'wS'");
    for (k = 0; k < 4; k++) o[k] = gw_out_like(&plhs[k], prhs[0], GW, onames[k]);
    ctx = gw_ctx(GW);
    gw_check(ctx, pdegpu_ddiff_weights(ctx, o[0], o[1], o[2], o[3], D.p, (int)D.nrows, (int)D.ncols, (int)D.nframes, eps), GW);
}
