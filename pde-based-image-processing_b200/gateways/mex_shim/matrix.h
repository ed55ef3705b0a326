/* matrix.h -- the reference includes both "mex.h" and "matrix.h"; everything lives in mex.h. */
#ifndef PDE_MEX_SHIM_MATRIX_H
#define PDE_MEX_SHIM_MATRIX_H
#include "mex.h"
#endif
