/* mexstub.c - minimal in-memory mxArray behind tests/mexstub/mex.h, plus the hs_* harness that
 * tests/test_mex_gateway.py drives through ctypes.  Test infrastructure only. */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mex.h"

struct mxArray_tag {
    mxClassID cls;
    mwSize ndim, dims[3];
    void *data;                 /* numeric / char payload (char: one byte per character) */
    int sparse;
    mwIndex *ir, *jc;
    int nfields;
    char **names;
    mxArray **fields;
};

static jmp_buf g_jmp;
static int g_in_call = 0;
static char g_err_id[256], g_err_msg[1024];
static int g_lock = 0;
static void (*g_at_exit)(void) = NULL;

static size_t elsize(mxClassID c) {
    switch (c) {
        case mxDOUBLE_CLASS: case mxINT64_CLASS: case mxUINT64_CLASS: return 8;
        case mxINT32_CLASS: case mxUINT32_CLASS: case mxSINGLE_CLASS: return 4;
        case mxINT16_CLASS: case mxUINT16_CLASS: return 2;
        default: return 1;
    }
}
static mxArray *mk(mxClassID cls, mwSize ndim, const mwSize *dims) {
    mxArray *a = (mxArray *)calloc(1, sizeof(mxArray));
    size_t numel = 1;
    a->cls = cls; a->ndim = ndim;
    a->dims[0] = a->dims[1] = a->dims[2] = 1;
    for (mwSize k = 0; k < ndim && k < 3; ++k) { a->dims[k] = dims[k]; numel *= dims[k]; }
    a->data = calloc(numel ? numel : 1, elsize(cls));
    return a;
}

bool mxIsDouble(const mxArray *a) { return a->cls == mxDOUBLE_CLASS; }
bool mxIsInt64(const mxArray *a) { return a->cls == mxINT64_CLASS; }
bool mxIsUint64(const mxArray *a) { return a->cls == mxUINT64_CLASS; }
bool mxIsComplex(const mxArray *a) { (void)a; return false; }
bool mxIsSparse(const mxArray *a) { return a->sparse != 0; }
bool mxIsStruct(const mxArray *a) { return a->cls == mxSTRUCT_CLASS; }
bool mxIsLogicalScalarTrue(const mxArray *a) {
    return a->cls == mxLOGICAL_CLASS && mxGetNumberOfElements(a) == 1 && *(unsigned char *)a->data != 0;
}
size_t mxGetNumberOfElements(const mxArray *a) { return a->dims[0] * a->dims[1] * a->dims[2]; }
size_t mxGetM(const mxArray *a) { return a->dims[0]; }
size_t mxGetN(const mxArray *a) { return a->dims[1] * a->dims[2]; }
void mxSetN(mxArray *a, mwSize n) { a->dims[1] = n; a->dims[2] = 1; if (a->ndim > 2) a->ndim = 2; }
void *mxGetData(const mxArray *a) { return a->data; }
mxDouble *mxGetDoubles(const mxArray *a) { return a->cls == mxDOUBLE_CLASS ? (mxDouble *)a->data : NULL; }
double mxGetScalar(const mxArray *a) {
    switch (a->cls) {
        case mxDOUBLE_CLASS: return *(double *)a->data;
        case mxINT64_CLASS: return (double)*(int64_t *)a->data;
        case mxUINT64_CLASS: return (double)*(uint64_t *)a->data;
        case mxINT32_CLASS: return (double)*(int32_t *)a->data;
        case mxLOGICAL_CLASS: return (double)*(unsigned char *)a->data;
        default: return 0.0;
    }
}
int mxGetString(const mxArray *a, char *buf, mwSize buflen) {
    if (a->cls != mxCHAR_CLASS) return 1;
    size_t n = mxGetNumberOfElements(a);
    if (n + 1 > buflen) return 1;
    memcpy(buf, a->data, n);
    buf[n] = 0;
    return 0;
}
mxArray *mxGetField(const mxArray *a, mwIndex index, const char *name) {
    if (a->cls != mxSTRUCT_CLASS || index != 0) return NULL;
    for (int k = 0; k < a->nfields; ++k) if (!strcmp(a->names[k], name)) return a->fields[k];
    return NULL;
}
void mxSetField(mxArray *a, mwIndex index, const char *name, mxArray *value) {
    if (a->cls != mxSTRUCT_CLASS || index != 0) return;
    for (int k = 0; k < a->nfields; ++k) if (!strcmp(a->names[k], name)) { a->fields[k] = value; return; }
}
mwIndex *mxGetIr(const mxArray *a) { return a->ir; }
mwIndex *mxGetJc(const mxArray *a) { return a->jc; }
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity flag) {
    (void)flag; const mwSize d[2] = {m, n}; return mk(mxDOUBLE_CLASS, 2, d);
}
mxArray *mxCreateDoubleScalar(double v) { mxArray *a = mxCreateDoubleMatrix(1, 1, mxREAL); *(double *)a->data = v; return a; }
mxArray *mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity flag) {
    (void)flag; const mwSize d[2] = {m, n}; return mk(cls, 2, d);
}
mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID cls, mxComplexity flag) {
    (void)flag; return mk(cls, ndim, dims);
}
mxArray *mxCreateStructMatrix(mwSize m, mwSize n, int nfields, const char **names) {
    const mwSize d[2] = {m, n};
    mxArray *a = mk(mxSTRUCT_CLASS, 2, d);
    a->nfields = nfields;
    a->names = (char **)calloc(nfields ? nfields : 1, sizeof(char *));
    a->fields = (mxArray **)calloc(nfields ? nfields : 1, sizeof(mxArray *));
    for (int k = 0; k < nfields; ++k) a->names[k] = strdup(names[k]);
    return a;
}
mxArray *mxCreateSparse(mwSize m, mwSize n, mwSize nzmax, mxComplexity flag) {
    (void)flag;
    const mwSize d[2] = {m, n};
    mxArray *a = (mxArray *)calloc(1, sizeof(mxArray));
    a->cls = mxDOUBLE_CLASS; a->ndim = 2; a->dims[0] = d[0]; a->dims[1] = d[1]; a->dims[2] = 1;
    a->sparse = 1;
    a->data = calloc(nzmax ? nzmax : 1, sizeof(double));
    a->ir = (mwIndex *)calloc(nzmax ? nzmax : 1, sizeof(mwIndex));
    a->jc = (mwIndex *)calloc(n + 1, sizeof(mwIndex));
    return a;
}
void mxDestroyArray(mxArray *a) {
    if (!a) return;
    for (int k = 0; k < a->nfields; ++k) { mxDestroyArray(a->fields[k]); free(a->names[k]); }
    free(a->names); free(a->fields); free(a->ir); free(a->jc); free(a->data); free(a);
}
void *mxMalloc(mwSize n) { return malloc(n ? n : 1); }
void mxFree(void *p) { free(p); }

void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...) {
    va_list ap;
    snprintf(g_err_id, sizeof(g_err_id), "%s", id);
    va_start(ap, fmt);
    vsnprintf(g_err_msg, sizeof(g_err_msg), fmt, ap);
    va_end(ap);
    if (!g_in_call) { fprintf(stderr, "mexErrMsgIdAndTxt outside hs_call: %s: %s\n", g_err_id, g_err_msg); abort(); }
    longjmp(g_jmp, 1);
}
void mexLock(void) { ++g_lock; }
void mexUnlock(void) { --g_lock; }
int mexAtExit(void (*fn)(void)) { g_at_exit = fn; return 0; }

/* ---- harness (ctypes side) ---- */
mxArray *hs_double(size_t m, size_t n, const double *src) {
    mxArray *a = mxCreateDoubleMatrix(m, n, mxREAL);
    if (src && m * n > 0) memcpy(a->data, src, m * n * sizeof(double));
    return a;
}
mxArray *hs_int64(size_t m, size_t n, const int64_t *src) {
    mxArray *a = mxCreateNumericMatrix(m, n, mxINT64_CLASS, mxREAL);
    if (src && m * n > 0) memcpy(a->data, src, m * n * sizeof(int64_t));
    return a;
}
mxArray *hs_uint64(uint64_t v) {
    mxArray *a = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
    *(uint64_t *)a->data = v;
    return a;
}
mxArray *hs_logical(int v) {
    mxArray *a = mxCreateNumericMatrix(1, 1, mxLOGICAL_CLASS, mxREAL);
    *(unsigned char *)a->data = (unsigned char)(v != 0);
    return a;
}
mxArray *hs_string(const char *s) {
    const mwSize d[2] = {1, strlen(s)};
    mxArray *a = mk(mxCHAR_CLASS, 2, d);
    memcpy(a->data, s, d[1]);
    return a;
}
mxArray *hs_struct(int nfields, const char **names) { return mxCreateStructMatrix(1, 1, nfields, names); }
int hs_class(const mxArray *a) { return (int)a->cls; }
size_t hs_dim(const mxArray *a, int k) { return a->dims[k]; }
int hs_lock_count(void) { return g_lock; }
int hs_has_at_exit(void) { return g_at_exit != NULL; }
const char *hs_err_id(void) { return g_err_id; }
const char *hs_err_msg(void) { return g_err_msg; }

/* Calls mexFunction; 0 = returned normally, 1 = left through mexErrMsgIdAndTxt (id / message in hs_err_*). */
int hs_call(int nlhs, mxArray **plhs, int nrhs, mxArray **prhs) {
    g_err_id[0] = g_err_msg[0] = 0;
    g_in_call = 1;
    if (setjmp(g_jmp)) { g_in_call = 0; return 1; }
    mexFunction(nlhs, plhs, nrhs, (const mxArray **)prhs);
    g_in_call = 0;
    return 0;
}
int hs_nfields(const mxArray *a) { return a->nfields; }
const char *hs_field_name(const mxArray *a, int k) { return a->names[k]; }
mxArray *hs_field(const mxArray *a, int k) { return a->fields[k]; }
