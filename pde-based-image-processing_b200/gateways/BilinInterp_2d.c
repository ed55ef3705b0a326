/*
 * Iout = BilinInterp_2d(Iin,X,Y)
 *
 * libpdegpu gateway with the Matlab-visible signature of the reference's mex/source/BilinInterp_2d.c
 * (3 inputs :54; Iout shaped like Iin :103-118). Out-of-image look-ups yield NaN: the value the
 * reference's library function is written for (imageInterpolation.c:44-48,133) but that its gateway
 * forgets to pass (:120-123). Set PDEGPU_WARP_OOB=<float> to choose another value.
 */
#include "gw_common.h"
#include <math.h>
#define GW "BilinInterp_2d"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    gw_arr I, X, Y;
    float *out, oob = NAN;
    const char *env;
    pdegpu_ctx *ctx;

    if (nrhs != 3 || nlhs > 3) gw_fail(GW, "proper function call is 'bilinInterp2( Iin, X, Y)'");
    I = gw_in(prhs[0], GW, "Iin");
    X = gw_in(prhs[1], GW, "X");
    Y = gw_in(prhs[2], GW, "Y");
    if (nlhs < 1) gw_fail(GW, "insufficient number of outputs. Outputs from this function is 'Iout'");
    gw_need(&X, I.nrows * I.ncols, GW, "X");
    gw_need(&Y, I.nrows * I.ncols, GW, "Y");
    out = gw_out_like(&plhs[0], prhs[0], GW, "Iout");
    if (I.numel == 0) return;
    env = getenv("PDEGPU_WARP_OOB");
    if (env) oob = (float)atof(env);
    ctx = gw_ctx(GW);
    gw_check(ctx, pdegpu_bilin_interp_2d(ctx, out, I.p, X.p, Y.p, (int)I.nrows, (int)I.ncols, (int)I.nframes, oob), GW);
}
