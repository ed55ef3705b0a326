/*
 * mex.h -- minimal stand-in for Matlab's MEX C API, for building MEX gateways
 * as plain shared objects that a C / Python(ctypes) harness can drive.
 *
 * It implements exactly the subset of the API that the 13 hot-path gateways of
 * JediZ/PDE-based-image-processing use:
 *   mxIsSingle mxIsDouble mxGetNumberOfDimensions mxGetDimensions mxGetPr mxGetData
 *   mxGetNumberOfElements mxCreateNumericArray mxCalloc mxMalloc mxFree
 *   mexErrMsgTxt mexPrintf mxDestroyArray
 *
 * Two builds of the same shim exist:
 *   - default            : mwSize is size_t (what Matlab >= R2006b / Octave use);
 *                          libpdegpu's own gateways are built this way.
 *   - PDE_SHIM_MWSIZE32  : mwSize is a 32-bit unsigned int. The reference stores
 *                          mxGetDimensions() in `const unsigned int*`
 *                          (reference mex/source/library/opticalflowSolvers.h:46),
 *                          which is only right with 32-bit mwSize, so
 *                          the oracle build of the reference uses this mode.
 *
 * mexErrMsgTxt() does not return: it longjmp()s back into shim_call().
 */
#ifndef PDE_MEX_SHIM_MEX_H
#define PDE_MEX_SHIM_MEX_H

#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifdef PDE_SHIM_MWSIZE32
typedef unsigned int mwSize;
typedef unsigned int mwIndex;
#else
typedef size_t mwSize;
typedef size_t mwIndex;
#endif

typedef enum {
    mxUNKNOWN_CLASS = 0,
    mxDOUBLE_CLASS  = 6,
    mxSINGLE_CLASS  = 7,
    mxINT32_CLASS   = 12
} mxClassID;

typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;

#define PDE_SHIM_MAXDIMS 8

typedef struct mxArray_tag {
    int     classid;
    int     ndims;
    mwSize  dims[PDE_SHIM_MAXDIMS];
    void   *data;
    int     owns_data;
} mxArray;

bool          mxIsSingle(const mxArray *a);
bool          mxIsDouble(const mxArray *a);
mwSize        mxGetNumberOfDimensions(const mxArray *a);
const mwSize *mxGetDimensions(const mxArray *a);
size_t        mxGetNumberOfElements(const mxArray *a);
double       *mxGetPr(const mxArray *a);
void         *mxGetData(const mxArray *a);
mxArray      *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID classid, mxComplexity flag);
void          mxDestroyArray(mxArray *a);
void         *mxCalloc(size_t n, size_t size);
void         *mxMalloc(size_t n);
void          mxFree(void *p);
void          mexErrMsgTxt(const char *msg);
int           mexPrintf(const char *fmt, ...);
int           mexAtExit(void (*fn)(void));

/* the entry point every gateway defines (renamed per gateway with -DmexFunction=...) */
typedef void (*pde_mex_fn)(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);

/* ---- harness side (not part of Matlab's API) ---- */
/* Wrap caller-owned memory in an mxArray (data is borrowed, never freed by the shim). */
mxArray *shim_wrap(int classid, int ndims, const unsigned long long *dims, void *data);
int      shim_ndims(const mxArray *a);
unsigned long long shim_dim(const mxArray *a, int k);
void    *shim_data(const mxArray *a);
int      shim_classid(const mxArray *a);
int      shim_sizeof_mwsize(void);
/* Call a gateway. Returns 0 on normal return, 1 if it raised mexErrMsgTxt (message
 * copied to errbuf). On error every array/buffer the gateway allocated is released. */
int      shim_call(pde_mex_fn fn, int nlhs, mxArray **plhs, int nrhs, const mxArray **prhs,
                   char *errbuf, int errlen);

#ifdef __cplusplus
}
#endif
#endif
