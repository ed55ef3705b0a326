/*
 * [Idt Idx Idy] = FstDerivatives5(It0,It1)
 *
 * libpdegpu gateway with the Matlab-visible signature of the reference's mex/source/FstDerivatives5.c
 * (2 inputs :70; three outputs shaped like It0 :100-133; fstSimoncelli_c with the 5-tap filters :141).
 */
#include "gw_common.h"
#define GW "FstDerivatives5"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    gw_arr I0, I1;
    float *o[3];
    pdegpu_ctx *ctx;

    if (nrhs != 2) gw_fail(GW, "wrong number of input parameters!");
    I0 = gw_in(prhs[0], GW, "It0");
    I1 = gw_in(prhs[1], GW, "It1");
    if (nlhs < 3) gw_fail(GW, "insufficient number of outputs...outputs from this function are 'Idt', 'Idx' and 'Idy'.");
    gw_need(&I1, I0.nrows * I0.ncols * I0.nframes, GW, "It1");
    o[0] = gw_out_like(&plhs[0], prhs[0], GW, "Idt");
    o[1] = gw_out_like(&plhs[1], prhs[0], GW, "Idx");
    o[2] = gw_out_like(&plhs[2], prhs[0], GW, "Idy");
    ctx = gw_ctx(GW);
    gw_check(ctx, pdegpu_fst_derivatives5(ctx, o[0], o[1], o[2], I0.p, I1.p, (int)I0.nrows, (int)I0.ncols, (int)I0.nframes), GW);
}
