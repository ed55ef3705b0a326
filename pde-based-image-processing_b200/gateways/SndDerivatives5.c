/*
 * [Idxt Idyt Idxx Idyy Idxy] = SndDerivatives5(It0,It1)
 *
 * libpdegpu gateway with the Matlab-visible signature of the reference's mex/source/SndDerivatives5.c
 * (2 inputs :76; five outputs shaped like It0; sndSimoncelli_c :168).
 */
#include "gw_common.h"
#define GW "SndDerivatives5"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    static const char *onames[5] = {"Idxt", "Idyt", "Idxx", "Idyy", "Idxy"};
    gw_arr I0, I1;
    float *o[5];
    int k;
    pdegpu_ctx *ctx;

    if (nrhs != 2) gw_fail(GW, "wrong number of input parameters!");
    I0 = gw_in(prhs[0], GW, "It0");
    I1 = gw_in(prhs[1], GW, "It1");
    if (nlhs < 5) gw_fail(GW, "insufficient number of outputs...outputs from this function are 'Idxt', 'Idyt', 'Idxx', 'Idyy' and 'Idxy'.");
    gw_need(&I1, I0.nrows * I0.ncols * I0.nframes, GW, "It1");
    for (k = 0; k < 5; k++) o[k] = gw_out_like(&plhs[k], prhs[0], GW, onames[k]);
    ctx = gw_ctx(GW);
    gw_check(ctx, pdegpu_snd_derivatives5(ctx, o[0], o[1], o[2], o[3], o[4], I0.p, I1.p,
                                          (int)I0.nrows, (int)I0.ncols, (int)I0.nframes), GW);
}
