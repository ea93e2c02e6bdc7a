/* mex.h - test stand-in for MATLAB's MEX API (the image has no MATLAB).
 *
 * Declares exactly the documented C Matrix / MEX API subset that mex/dbat_mex.c uses (interleaved-complex
 * R2018a names: mxGetDoubles etc.), with the documented signatures, so that the gateway can be compiled with
 * -Wall -Werror, linked against libdbatgpu.so and driven from tests/test_mex_gateway.py.  The implementation
 * (mexstub.c) is a minimal in-memory mxArray; mexErrMsgIdAndTxt leaves mexFunction by longjmp, as MATLAB's
 * does by its own exception mechanism.  Test infrastructure only - never part of the product. */
#ifndef DBAT_TEST_MEX_H
#define DBAT_TEST_MEX_H
#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>

typedef size_t mwSize;
typedef size_t mwIndex;
typedef uint64_t uint64_T;
typedef int64_t int64_T;
typedef double mxDouble;
typedef struct mxArray_tag mxArray;

typedef enum {
    mxUNKNOWN_CLASS = 0, mxCELL_CLASS, mxSTRUCT_CLASS, mxLOGICAL_CLASS, mxCHAR_CLASS, mxVOID_CLASS,
    mxDOUBLE_CLASS, mxSINGLE_CLASS, mxINT8_CLASS, mxUINT8_CLASS, mxINT16_CLASS, mxUINT16_CLASS,
    mxINT32_CLASS, mxUINT32_CLASS, mxINT64_CLASS, mxUINT64_CLASS, mxFUNCTION_CLASS
} mxClassID;
typedef enum { mxREAL = 0, mxCOMPLEX } mxComplexity;

#ifdef __cplusplus
extern "C" {
#endif
bool mxIsDouble(const mxArray *a);
bool mxIsInt64(const mxArray *a);
bool mxIsUint64(const mxArray *a);
bool mxIsComplex(const mxArray *a);
bool mxIsSparse(const mxArray *a);
bool mxIsStruct(const mxArray *a);
bool mxIsLogicalScalarTrue(const mxArray *a);
size_t mxGetNumberOfElements(const mxArray *a);
size_t mxGetM(const mxArray *a);
size_t mxGetN(const mxArray *a);
void mxSetN(mxArray *a, mwSize n);
void *mxGetData(const mxArray *a);
mxDouble *mxGetDoubles(const mxArray *a);
double mxGetScalar(const mxArray *a);
int mxGetString(const mxArray *a, char *buf, mwSize buflen);
mxArray *mxGetField(const mxArray *a, mwIndex index, const char *name);
void mxSetField(mxArray *a, mwIndex index, const char *name, mxArray *value);
mwIndex *mxGetIr(const mxArray *a);
mwIndex *mxGetJc(const mxArray *a);
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity flag);
mxArray *mxCreateDoubleScalar(double v);
mxArray *mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity flag);
mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID cls, mxComplexity flag);
mxArray *mxCreateStructMatrix(mwSize m, mwSize n, int nfields, const char **names);
mxArray *mxCreateSparse(mwSize m, mwSize n, mwSize nzmax, mxComplexity flag);
void mxDestroyArray(mxArray *a);
void *mxMalloc(mwSize n);
void mxFree(void *p);

void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...) __attribute__((noreturn, format(printf, 2, 3)));
void mexLock(void);
void mexUnlock(void);
int mexAtExit(void (*fn)(void));

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);
#ifdef __cplusplus
}
#endif
#endif
