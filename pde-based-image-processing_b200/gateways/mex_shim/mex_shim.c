/*
 * mex_shim.c -- implementation of the minimal MEX API declared in mex.h.
 * See mex.h for scope. Single-threaded per shared object (one jmp_buf), like Matlab's
 * interpreter thread that calls mexFunction.
 */
#include "mex.h"
#include <setjmp.h>
#include <stdarg.h>

#define SHIM_MAX_TRACK 256

static jmp_buf  g_jmp;
static int      g_in_call = 0;
static char     g_err[512];
static void    *g_blocks[SHIM_MAX_TRACK];   /* mxCalloc/mxMalloc blocks alive in this call */
static int      g_nblocks = 0;
static mxArray *g_arrays[SHIM_MAX_TRACK];   /* arrays created in this call */
static int      g_narrays = 0;

bool mxIsSingle(const mxArray *a) { return a && a->classid == mxSINGLE_CLASS; }
bool mxIsDouble(const mxArray *a) { return a && a->classid == mxDOUBLE_CLASS; }
mwSize mxGetNumberOfDimensions(const mxArray *a) { return (mwSize)a->ndims; }
const mwSize *mxGetDimensions(const mxArray *a) { return a->dims; }
double *mxGetPr(const mxArray *a) { return (double *)a->data; }
void *mxGetData(const mxArray *a) { return a->data; }

size_t mxGetNumberOfElements(const mxArray *a)
{
    size_t n = 1;
    int k;
    for (k = 0; k < a->ndims; k++) n *= (size_t)a->dims[k];
    return n;
}

static size_t class_size(int classid)
{
    switch (classid) {
    case mxDOUBLE_CLASS: return 8;
    case mxSINGLE_CLASS: return 4;
    case mxINT32_CLASS:  return 4;
    default:             return 0;
    }
}

mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID classid, mxComplexity flag)
{
    mxArray *a;
    size_t n = 1, es = class_size(classid);
    mwSize k;
    (void)flag;
    if (ndim > PDE_SHIM_MAXDIMS || es == 0) return NULL;
    a = (mxArray *)calloc(1, sizeof(mxArray));
    if (!a) return NULL;
    a->classid = classid;
    a->ndims = (int)ndim;
    for (k = 0; k < ndim; k++) { a->dims[k] = dims[k]; n *= (size_t)dims[k]; }
    /* Matlab never reports fewer than 2 dims */
    if (a->ndims < 2) { for (k = ndim; k < 2; k++) a->dims[k] = 1; a->ndims = 2; }
    a->data = calloc(n ? n : 1, es);          /* zero-initialised, like Matlab */
    a->owns_data = 1;
    if (!a->data) { free(a); return NULL; }
    if (g_in_call && g_narrays < SHIM_MAX_TRACK) g_arrays[g_narrays++] = a;
    return a;
}

void mxDestroyArray(mxArray *a)
{
    int k;
    if (!a) return;
    for (k = 0; k < g_narrays; k++) if (g_arrays[k] == a) g_arrays[k] = NULL;
    if (a->owns_data) free(a->data);
    free(a);
}

static void *track(void *p)
{
    if (p && g_in_call && g_nblocks < SHIM_MAX_TRACK) g_blocks[g_nblocks++] = p;
    return p;
}

void *mxCalloc(size_t n, size_t size) { return track(calloc(n ? n : 1, size ? size : 1)); }
void *mxMalloc(size_t n) { return track(malloc(n ? n : 1)); }

void mxFree(void *p)
{
    int k;
    if (!p) return;
    for (k = 0; k < g_nblocks; k++) if (g_blocks[k] == p) g_blocks[k] = NULL;
    free(p);
}

void mexErrMsgTxt(const char *msg)
{
    snprintf(g_err, sizeof(g_err), "%s", msg ? msg : "");
    if (g_in_call) longjmp(g_jmp, 1);
    fprintf(stderr, "mexErrMsgTxt outside shim_call: %s\n", g_err);
    abort();
}

int mexPrintf(const char *fmt, ...)
{
    int r;
    va_list ap;
    va_start(ap, fmt);
    r = vfprintf(stdout, fmt, ap);
    va_end(ap);
    return r;
}

int mexAtExit(void (*fn)(void))
{
    /* Matlab runs fn when the MEX file is cleared; in the harness the process exit does the job */
    return atexit(fn);
}

/* ---------------- harness side ---------------- */

mxArray *shim_wrap(int classid, int ndims, const unsigned long long *dims, void *data)
{
    mxArray *a;
    int k;
    if (ndims > PDE_SHIM_MAXDIMS) return NULL;
    a = (mxArray *)calloc(1, sizeof(mxArray));
    if (!a) return NULL;
    a->classid = classid;
    a->ndims = ndims;
    for (k = 0; k < ndims; k++) a->dims[k] = (mwSize)dims[k];
    if (a->ndims < 2) { for (k = ndims; k < 2; k++) a->dims[k] = 1; a->ndims = 2; }
    a->data = data;
    a->owns_data = 0;
    return a;
}

int shim_ndims(const mxArray *a) { return a->ndims; }
unsigned long long shim_dim(const mxArray *a, int k) { return (unsigned long long)a->dims[k]; }
void *shim_data(const mxArray *a) { return a->data; }
int shim_classid(const mxArray *a) { return a->classid; }
int shim_sizeof_mwsize(void) { return (int)sizeof(mwSize); }

int shim_call(pde_mex_fn fn, int nlhs, mxArray **plhs, int nrhs, const mxArray **prhs,
              char *errbuf, int errlen)
{
    int k;
    g_nblocks = 0;
    g_narrays = 0;
    g_err[0] = 0;
    g_in_call = 1;
    if (setjmp(g_jmp) == 0) {
        fn(nlhs, plhs, nrhs, prhs);
        g_in_call = 0;
        /* a well-behaved gateway freed its mxCalloc blocks; Matlab would free the rest */
        for (k = 0; k < g_nblocks; k++) if (g_blocks[k]) free(g_blocks[k]);
        g_nblocks = 0;
        g_narrays = 0;
        if (errbuf && errlen > 0) errbuf[0] = 0;
        return 0;
    }
    /* error path: release everything the gateway allocated, like Matlab does */
    g_in_call = 0;
    for (k = 0; k < g_nblocks; k++) if (g_blocks[k]) free(g_blocks[k]);
    for (k = 0; k < g_narrays; k++) {
        if (g_arrays[k]) {
            int q;
            for (q = 0; q < nlhs; q++) if (plhs[q] == g_arrays[k]) plhs[q] = NULL;
            if (g_arrays[k]->owns_data) free(g_arrays[k]->data);
            free(g_arrays[k]);
        }
    }
    g_nblocks = 0;
    g_narrays = 0;
    if (errbuf && errlen > 0) snprintf(errbuf, (size_t)errlen, "%s", g_err);
    return 1;
}
