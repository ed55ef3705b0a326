/*
 * gw_common.h -- shared plumbing of libpdegpu's MEX gateways.
 *
 * Each gateway (one .c file per MEX function, same file names as the reference's mex/source/*.c)
 * keeps the reference's Matlab-visible signature: same nrhs/nlhs, every numeric argument must be
 * real `single` (scalars included), outputs are freshly created single arrays shaped like the
 * reference shapes them. The body is marshalling only: the arithmetic lives in libpdegpu.
 *
 * Differences from the reference gateways, all on the safe side:
 *   - mwSize is read as mwSize (the reference stores it in `unsigned int*`);
 *   - arrays that must share a shape are checked (the reference reads out of bounds instead);
 *   - library errors (no GPU, out of memory, ...) surface as mexErrMsgTxt.
 */
#ifndef PDEGPU_GW_COMMON_H
#define PDEGPU_GW_COMMON_H

#include "mex.h"
#include "pdegpu.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct gw_arr {
    const float *p;
    size_t nrows, ncols, nframes, numel;
    const mxArray *mx;
} gw_arr;

/* process-wide context, created on first use on device $PDEGPU_DEVICE (default 0) */
pdegpu_ctx *gw_ctx(const char *gw);

static void gw_fail(const char *gw, const char *what)
{
    char msg[400];
    snprintf(msg, sizeof(msg), "%s: %s", gw, what);
    mexErrMsgTxt(msg);
}

static gw_arr gw_in(const mxArray *a, const char *gw, const char *name)
{
    gw_arr r;
    char msg[200];
    const mwSize *d;
    mwSize nd, k;
    if (!mxIsSingle(a)) {
        snprintf(msg, sizeof(msg), "'%s' must be a noncomplex single-valued matrix.", name);
        gw_fail(gw, msg);
    }
    nd = mxGetNumberOfDimensions(a);
    d = mxGetDimensions(a);
    r.mx = a;
    r.p = (const float *)mxGetData(a);
    r.nrows = nd > 0 ? (size_t)d[0] : 1;
    r.ncols = nd > 1 ? (size_t)d[1] : 1;
    r.nframes = nd > 2 ? (size_t)d[2] : 1;
    r.numel = 1;
    for (k = 0; k < nd; k++) r.numel *= (size_t)d[k];
    return r;
}

static float gw_scalar(const mxArray *a, const char *gw, const char *name)
{
    char msg[200];
    if (!mxIsSingle(a)) {
        snprintf(msg, sizeof(msg), "'%s' must be a noncomplex, single-type scalar", name);
        gw_fail(gw, msg);
    }
    if (mxGetNumberOfElements(a) < 1) {
        snprintf(msg, sizeof(msg), "'%s' is empty", name);
        gw_fail(gw, msg);
    }
    return *(const float *)mxGetData(a);
}

/* every field array of a solver call must hold at least nrows*ncols*frames elements */
static void gw_need(const gw_arr *a, size_t n, const char *gw, const char *name)
{
    char msg[200];
    if (a->numel < n) {
        snprintf(msg, sizeof(msg), "'%s' has fewer elements than the solution arrays", name);
        gw_fail(gw, msg);
    }
}

/* new single array shaped like `like` (what mxCreateNumericArray(like.ndims, like.dims) gives) */
static float *gw_out_like(mxArray **slot, const mxArray *like, const char *gw, const char *name)
{
    char msg[200];
    *slot = mxCreateNumericArray(mxGetNumberOfDimensions(like), mxGetDimensions(like), mxSINGLE_CLASS, mxREAL);
    if (*slot == NULL || mxGetData(*slot) == NULL) {
        snprintf(msg, sizeof(msg), "error reserving space for output variable '%s'", name);
        gw_fail(gw, msg);
    }
    return (float *)mxGetData(*slot);
}

static void gw_check(pdegpu_ctx *ctx, int rc, const char *gw)
{
    if (rc != PDEGPU_OK) gw_fail(gw, pdegpu_last_error(ctx));
}

#endif
