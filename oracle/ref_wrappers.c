/*
 * ref_wrappers.c -- ORACLE / TEST INFRASTRUCTURE ONLY (linked into oracle/_ref/ref_interp.so).
 *
 * The reference defines bilinInterp2() with FIVE parameters (the last one is the value written
 * for out-of-image look-ups, reference mex/source/library/imageInterpolation.c:44-48) but its
 * gateway calls it with FOUR through an implicit declaration
 * (reference mex/source/BilinInterp_2d.c:120-123; no prototype in imageInterpolation.h:28-59),
 * so the out-of-image value of the shipped MEX is whatever xmm0 happens to hold.
 * This wrapper calls the unmodified library function through an explicit 5-argument prototype
 * so that the oracle can be driven with a defined out-of-image value (NaN = the author's intent,
 * 0.0f = what gcc 13 -O2 produces for the 4-argument call).
 */
#include "imageInterpolation.h"

void bilinInterp2(struct matrixM *Iout, struct matrixM *Iin, struct matrixM *X, struct matrixM *Y, float NaN);

void ref_bilinInterp2(float *out, float *in, float *X, float *Y,
                      unsigned int nrows, unsigned int ncols, unsigned int nframes, float oob)
{
    unsigned int dims[3];
    struct matrixM mo = {0, NULL, NULL, 0, 0, 0, 0};
    struct matrixM mi = {0, NULL, NULL, 0, 0, 0, 0};
    struct matrixM mx = {0, NULL, NULL, 0, 0, 0, 0};
    struct matrixM my = {0, NULL, NULL, 0, 0, 0, 0};
    dims[0] = nrows; dims[1] = ncols; dims[2] = nframes;
    mo.ndims = mi.ndims = (nframes > 1) ? 3 : 2;
    mx.ndims = my.ndims = 2;
    mo.dimElems = mi.dimElems = mx.dimElems = my.dimElems = dims;
    mo.data = out; mi.data = in; mx.data = X; my.data = Y;
    bilinInterp2(&mo, &mi, &mx, &my, oob);
}
