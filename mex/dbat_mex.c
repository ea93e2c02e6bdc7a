/* dbat_mex.c — MEX gateway between MATLAB and libdbatgpu.so (include/dbat_gpu.h).
 *
 * Logic-free marshalling only; follows the conventions of the reference's own MEX files
 * (code/test/postcov/icpc_mex.c:495-611): argument checks raise
 * mexErrMsgIdAndTxt("DBAT:dbat_mex:<id>", ...), every returned array is allocated by the
 * gateway so MATLAB owns it, and a same-named dbat_mex.m stub errors 'Mex file not found.'
 * (icpc_mex.m:14) until this file has been compiled:
 *
 *     mex -R2018a CFLAGS='$CFLAGS -Wall' -I../include dbat_mex.c -L../dbat_b200 -ldbatgpu
 *
 * No MATLAB / mex.h exists in the build image.  tests/test_mex_gateway.py compiles this file (-Wall -Wextra
 * -Werror) against tests/mexstub/mex.h, a stand-in declaring the documented MEX API subset used here, links it
 * to libdbatgpu.so and drives mexFunction: argument validation and error ids on CPU, and on a B200 a whole
 * create / eval / jacobian / solve / cov / covstats / forwintersect / resect3 / destroy session whose results equal the
 * ctypes path bit for bit (profiles/mex_gateway_session_r2g.log).
 *
 *   h        = dbat_mex('create', d)        d: struct with the fields of dbat_problem_desc
 *   [f]      = dbat_mex('eval', h, x, weighted)
 *   J        = dbat_mex('jacobian', h, weighted)             sparse m-by-n
 *   r        = dbat_mex('solve', h, method, opts, x0)        struct with x,code,n,p,T,rr,damping,...
 *   C        = dbat_mex('cov', h, which, s0)
 *              dbat_mex('destroy', h)
 */
#include <string.h>
#include "mex.h"
#include "matrix.h"
#include "dbat_gpu.h"

#define ERR(id, msg) mexErrMsgIdAndTxt("DBAT:dbat_mex:" id, "%s", msg)

static int nHandles = 0;

static dbat_handle *get_handle(const mxArray *a) {
    if (!mxIsUint64(a) || mxGetNumberOfElements(a) != 1) ERR("badHandle", "Handle must be a uint64 scalar.");
    return (dbat_handle *)(uintptr_t)(*(uint64_T *)mxGetData(a));
}
static const double *dfield(const mxArray *s, const char *name, mwSize numel) {
    const mxArray *f = mxGetField(s, 0, name);
    if (!f || !mxIsDouble(f) || mxIsComplex(f) || mxIsSparse(f)) ERR("badField", name);
    if (numel && mxGetNumberOfElements(f) != numel) ERR("badSize", name);
    return mxGetDoubles(f);
}
static const int64_t *ifield(const mxArray *s, const char *name, mwSize *numel) {
    const mxArray *f = mxGetField(s, 0, name);
    if (!f || !mxIsInt64(f)) ERR("badField", name);
    if (numel) *numel = mxGetNumberOfElements(f);
    return (const int64_t *)mxGetData(f);
}
static double sfield(const mxArray *s, const char *name) {
    const mxArray *f = mxGetField(s, 0, name);
    if (!f || mxGetNumberOfElements(f) != 1) ERR("badField", name);
    return mxGetScalar(f);
}
static void check(dbat_handle *h, int rc) {
    if (rc != DBAT_OK) mexErrMsgIdAndTxt("DBAT:dbat_mex:library", "%s (code %d)", dbat_last_error(h), rc);
}
static void at_exit(void) { /* handles are destroyed by their owners (bundle.m onCleanup) */ }

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    char cmd[32];
    if (nrhs < 1 || mxGetString(prhs[0], cmd, sizeof(cmd))) ERR("nrhs", "First argument must be a command string.");
    mexAtExit(at_exit);

    if (!strcmp(cmd, "create")) {
        if (nrhs != 2 || !mxIsStruct(prhs[1])) ERR("nrhs", "create needs one struct argument.");
        const mxArray *d = prhs[1];
        dbat_problem_desc p;
        mwSize k;
        memset(&p, 0, sizeof(p));
        p.nImg = (int64_t)sfield(d, "nImg"); p.nOP = (int64_t)sfield(d, "nOP"); p.nIP = (int64_t)sfield(d, "nIP");
        p.distModel = (int32_t)sfield(d, "distModel"); p.nK = (int32_t)sfield(d, "nK"); p.nP = (int32_t)sfield(d, "nP");
        p.n = (int64_t)sfield(d, "n");
        p.IOval = dfield(d, "IOval", (5 + p.nK + p.nP) * p.nImg);
        p.EOval = dfield(d, "EOval", 6 * p.nImg);
        p.OPval = dfield(d, "OPval", 3 * p.nOP);
        p.IPval = dfield(d, "IPval", 2 * p.nIP);
        p.IPstd = dfield(d, "IPstd", 2 * p.nIP);
        p.pxSize = dfield(d, "pxSize", 2 * p.nImg);
        p.IPimg = ifield(d, "IPimg", &k); if (k != (mwSize)p.nIP) ERR("badSize", "IPimg");
        p.IPop = ifield(d, "IPop", &k);   if (k != (mwSize)p.nIP) ERR("badSize", "IPop");
        p.IOdes_src = ifield(d, "IOdes_src", &k); p.nIOdes = k; p.IOdes_dest = ifield(d, "IOdes_dest", &k);
        p.EOdes_src = ifield(d, "EOdes_src", &k); p.nEOdes = k; p.EOdes_dest = ifield(d, "EOdes_dest", &k);
        p.OPdes_src = ifield(d, "OPdes_src", &k); p.nOPdes = k; p.OPdes_dest = ifield(d, "OPdes_dest", &k);
        p.nPriorIO = (int64_t)sfield(d, "nPriorIO"); p.nPriorEO = (int64_t)sfield(d, "nPriorEO");
        p.nPriorOP = (int64_t)sfield(d, "nPriorOP");
        p.prior_x = ifield(d, "prior_x", NULL);
        p.prior_val = dfield(d, "prior_val", 0); p.prior_std = dfield(d, "prior_std", 0);
        dbat_handle *h = NULL;
        int rc = dbat_create(&p, &h);
        if (rc != DBAT_OK) mexErrMsgIdAndTxt("DBAT:dbat_mex:create", "%s (code %d)", dbat_last_error(NULL), rc);
        plhs[0] = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
        *(uint64_T *)mxGetData(plhs[0]) = (uint64_T)(uintptr_t)h;
        if (nHandles++ == 0) mexLock();
        return;
    }
    if (!strcmp(cmd, "forwintersect")) {
        /* [OP,res]=dbat_mex('forwintersect',IO,EO,pxSize,IPval,im,op,pts,nK,nP)  (forwintersect.m:28-41) */
        if (nrhs != 10) ERR("nrhs", "forwintersect(IO,EO,pxSize,IPval,im,op,pts,nK,nP)");
        for (int a = 5; a <= 7; ++a) if (!mxIsInt64(prhs[a])) ERR("badType", "im, op and pts must be int64.");
        dbat_fwi_desc f;
        memset(&f, 0, sizeof(f));
        f.NC = (int64_t)mxGetM(prhs[1]); f.nImg = (int64_t)mxGetN(prhs[1]);
        f.nObs = (int64_t)mxGetN(prhs[4]); f.nPts = (int64_t)mxGetNumberOfElements(prhs[7]);
        f.nK = (int64_t)mxGetScalar(prhs[8]); f.nP = (int64_t)mxGetScalar(prhs[9]);
        if (mxGetM(prhs[2]) != 6 || (int64_t)mxGetN(prhs[2]) != f.nImg) ERR("badSize", "EO must be 6-by-nImg.");
        if (mxGetNumberOfElements(prhs[3]) != (mwSize)(2 * f.nImg)) ERR("badSize", "pxSize must be 2-by-nImg.");
        if ((int64_t)mxGetNumberOfElements(prhs[5]) != f.nObs || (int64_t)mxGetNumberOfElements(prhs[6]) != f.nObs)
            ERR("badSize", "im and op must have one element per image point.");
        f.IO = mxGetDoubles(prhs[1]); f.EO = mxGetDoubles(prhs[2]); f.pxSize = mxGetDoubles(prhs[3]);
        f.IPval = mxGetDoubles(prhs[4]);
        f.obs_img = (const int64_t *)mxGetData(prhs[5]); f.obs_op = (const int64_t *)mxGetData(prhs[6]);
        f.pts = (const int64_t *)mxGetData(prhs[7]);
        f.nOP = 0;
        for (int64_t k = 0; k < f.nObs; ++k) if (f.obs_op[k] > f.nOP) f.nOP = f.obs_op[k];
        for (int64_t k = 0; k < f.nPts; ++k) if (f.pts[k] > f.nOP) f.nOP = f.pts[k];
        plhs[0] = mxCreateDoubleMatrix(3, (mwSize)f.nPts, mxREAL);
        mxArray *res = mxCreateDoubleMatrix(1, (mwSize)f.nPts, mxREAL);
        int rc = dbat_forwintersect(&f, mxGetDoubles(plhs[0]), mxGetDoubles(res), NULL);
        if (rc != DBAT_OK) mexErrMsgIdAndTxt("DBAT:dbat_mex:forwintersect", "%s (code %d)", dbat_forwintersect_error(), rc);
        if (nlhs > 1) plhs[1] = res; else mxDestroyArray(res);
        return;
    }
    if (!strcmp(cmd, "resect3")) {
        /* [EO,res]=dbat_mex('resect3',X3,x3,testStart,XT,xT,behind)  (resect.m:96-129, pm_resect_3pt.m) */
        if (nrhs != 7) ERR("nrhs", "resect3(X3,x3,testStart,XT,xT,behind)");
        if (!mxIsInt64(prhs[3])) ERR("badType", "testStart must be int64 (0-based offsets).");
        dbat_resect_desc r;
        memset(&r, 0, sizeof(r));
        r.nCam = (int64_t)(mxGetNumberOfElements(prhs[1]) / 9);
        if (mxGetNumberOfElements(prhs[2]) != (mwSize)(6 * r.nCam) || mxGetNumberOfElements(prhs[3]) != (mwSize)(r.nCam + 1))
            ERR("badSize", "X3 is 3x3xN, x3 is 2x3xN, testStart has N+1 elements.");
        r.X3 = mxGetDoubles(prhs[1]); r.x3 = mxGetDoubles(prhs[2]);
        r.test_start = (const int64_t *)mxGetData(prhs[3]);
        r.XT = mxGetDoubles(prhs[4]); r.xT = mxGetDoubles(prhs[5]);
        r.behind = mxIsLogicalScalarTrue(prhs[6]) || mxGetScalar(prhs[6]) != 0;
        if (mxGetNumberOfElements(prhs[4]) != (mwSize)(3 * r.test_start[r.nCam]) ||
            mxGetNumberOfElements(prhs[5]) != (mwSize)(2 * r.test_start[r.nCam])) ERR("badSize", "XT is 3xT, xT is 2xT.");
        plhs[0] = mxCreateDoubleMatrix(6, (mwSize)r.nCam, mxREAL);
        mxArray *res = mxCreateDoubleMatrix(1, (mwSize)r.nCam, mxREAL);
        int rc = dbat_resect3(&r, mxGetDoubles(plhs[0]), mxGetDoubles(res));
        if (rc != DBAT_OK) mexErrMsgIdAndTxt("DBAT:dbat_mex:resect3", "%s (code %d)", dbat_forwintersect_error(), rc);
        if (nlhs > 1) plhs[1] = res; else mxDestroyArray(res);
        return;
    }
    if (nrhs < 2) ERR("nrhs", "Handle required.");
    dbat_handle *h = get_handle(prhs[1]);
    const mwSize n = (mwSize)dbat_num_unknowns(h), m = (mwSize)dbat_num_residuals(h);

    if (!strcmp(cmd, "destroy")) {
        dbat_destroy(h);
        if (--nHandles == 0) mexUnlock();
    } else if (!strcmp(cmd, "devices")) {
        /* dbat_mex('devices', h, [0 1 2 3]): one MATLAB process drives several GPUs (dbat_set_devices) */
        if (nrhs != 3) ERR("nrhs", "devices(h,deviceList)");
        const mwSize nd = mxGetNumberOfElements(prhs[2]);
        int dev[64];
        if (nd < 1 || nd > 64) ERR("badSize", "1 to 64 devices.");
        for (mwSize k = 0; k < nd; ++k) dev[k] = (int)mxGetDoubles(prhs[2])[k];
        check(h, dbat_set_devices(h, dev, (int)nd));
    } else if (!strcmp(cmd, "eval")) {
        if (nrhs != 4 || mxGetNumberOfElements(prhs[2]) != n) ERR("badSize", "x must have n elements.");
        plhs[0] = mxCreateDoubleMatrix(m, 1, mxREAL);
        check(h, dbat_eval(h, mxGetDoubles(prhs[2]), mxGetDoubles(plhs[0]), mxGetScalar(prhs[3]) != 0));
    } else if (!strcmp(cmd, "jacobian")) {
        const int w = nrhs > 2 && mxGetScalar(prhs[2]) != 0;
        int64_t nnz = 0;
        check(h, dbat_jacobian_nnz(h, w, &nnz));
        plhs[0] = mxCreateSparse(m, n, (mwSize)(nnz > 0 ? nnz : 1), mxREAL);
        /* mwIndex is a 64-bit unsigned integer on every platform -R2018a supports */
        check(h, dbat_jacobian_csc(h, w, (int64_t *)mxGetJc(plhs[0]), (int64_t *)mxGetIr(plhs[0]),
                                   mxGetDoubles(plhs[0])));
    } else if (!strcmp(cmd, "solve")) {
        if (nrhs != 5 || !mxIsStruct(prhs[3]) || mxGetNumberOfElements(prhs[4]) != n) ERR("nrhs", "solve(h,method,opts,x0)");
        const int method = (int)mxGetScalar(prhs[2]);
        dbat_opts o;
        dbat_default_opts(method, &o);
        const mxArray *so = prhs[3];
        o.maxIter = (int32_t)sfield(so, "maxIter"); o.convTol = sfield(so, "convTol");
        o.absTerm = (int32_t)sfield(so, "absTerm"); o.singularTest = (int32_t)sfield(so, "singularTest");
        o.doTrace = (int32_t)sfield(so, "doTrace");
        if (mxGetField(so, 0, "lambda0")) { o.lambda0 = sfield(so, "lambda0"); o.lambdaMin = sfield(so, "lambdaMin"); }
        if (mxGetField(so, 0, "delta0")) o.delta0 = sfield(so, "delta0");
        if (mxGetField(so, 0, "mu")) o.mu = sfield(so, "mu");
        if (mxGetField(so, 0, "eta")) o.eta = sfield(so, "eta");
        if (mxGetField(so, 0, "alphaMin")) o.alphaMin = sfield(so, "alphaMin");
        const mwSize cap = (mwSize)o.maxIter + 2;
        const char *fn[] = {"x", "p", "r_w", "r_u", "T", "rr", "damping", "rhos", "steps", "code", "n", "seconds"};
        mxArray *R = mxCreateStructMatrix(1, 1, 12, fn);
        mxArray *x = mxCreateDoubleMatrix(n, 1, mxREAL), *p = mxCreateDoubleMatrix(n, 1, mxREAL);
        mxArray *rw = mxCreateDoubleMatrix(m, 1, mxREAL), *ru = mxCreateDoubleMatrix(m, 1, mxREAL);
        mxArray *T = mxCreateDoubleMatrix(n, cap, mxREAL), *rr = mxCreateDoubleMatrix(1, cap + 1, mxREAL);
        mxArray *dm = mxCreateDoubleMatrix(1, cap + 1, mxREAL), *rh = mxCreateDoubleMatrix(1, cap, mxREAL);
        mxArray *st = mxCreateNumericMatrix(1, cap, mxINT32_CLASS, mxREAL);
        dbat_result res;
        memset(&res, 0, sizeof(res));
        res.x = mxGetDoubles(x); res.p = mxGetDoubles(p); res.r_w = mxGetDoubles(rw); res.r_u = mxGetDoubles(ru);
        res.trace = mxGetDoubles(T); res.rr = mxGetDoubles(rr); res.damping = mxGetDoubles(dm);
        res.rhos = mxGetDoubles(rh); res.steps = (int32_t *)mxGetData(st);
        check(h, dbat_solve(h, method, &o, mxGetDoubles(prhs[4]), &res));
        mxSetN(T, res.nTrace); mxSetN(rr, res.nRr); mxSetN(dm, res.nDamping); mxSetN(rh, res.nRhos);
        mxSetN(st, res.nDamping);
        mxSetField(R, 0, "x", x); mxSetField(R, 0, "p", p); mxSetField(R, 0, "r_w", rw); mxSetField(R, 0, "r_u", ru);
        mxSetField(R, 0, "T", T); mxSetField(R, 0, "rr", rr); mxSetField(R, 0, "damping", dm);
        mxSetField(R, 0, "rhos", rh); mxSetField(R, 0, "steps", st);
        mxSetField(R, 0, "code", mxCreateDoubleScalar(res.code)); mxSetField(R, 0, "n", mxCreateDoubleScalar(res.iters));
        mxSetField(R, 0, "seconds", mxCreateDoubleScalar(res.seconds));
        plhs[0] = R;
    } else if (!strcmp(cmd, "cov")) {
        if (nrhs != 5) ERR("nrhs", "cov(h,which,s0,dims)");
        /* dims = [rows cols pages] of the block array, computed by bundle_cov.m from the struct */
        const double *dims = mxGetDoubles(prhs[4]);
        const mwSize sz[3] = {(mwSize)dims[0], (mwSize)dims[1], (mwSize)dims[2]};
        plhs[0] = mxCreateNumericArray(3, sz, mxDOUBLE_CLASS, mxREAL);
        check(h, dbat_cov(h, (int)mxGetScalar(prhs[2]), mxGetScalar(prhs[3]), mxGetDoubles(plhs[0])));
    } else if (!strcmp(cmd, "covstats")) {
        /* [sd, eo, op, io] = dbat_mex('covstats', h, s0, thres): standard deviations (n x 1) and, per block kind, a
         * k x 4 matrix [block row col rho] (1-based) of the pairs with |correlation| > thres (dbat_cov_stats) */
        if (nrhs != 4) ERR("nrhs", "covstats(h,s0,thres)");
        plhs[0] = mxCreateDoubleMatrix(n, 1, mxREAL);
        dbat_cov_hit_list L[3];
        const mwSize cap = 1 << 20;
        for (int k = 0; k < 3; ++k) {
            L[k].cap = (int64_t)cap; L[k].n = 0;
            L[k].block = (int64_t *)mxMalloc(cap * sizeof(int64_t));
            L[k].row = (int32_t *)mxMalloc(cap * sizeof(int32_t));
            L[k].col = (int32_t *)mxMalloc(cap * sizeof(int32_t));
            L[k].rho = (double *)mxMalloc(cap * sizeof(double));
        }
        check(h, dbat_cov_stats(h, mxGetScalar(prhs[2]), mxGetScalar(prhs[3]), mxGetDoubles(plhs[0]), &L[2], &L[0], &L[1]));
        for (int k = 0; k < 3 && k + 1 < nlhs; ++k) {
            const mwSize m = (mwSize)(L[k].n < L[k].cap ? L[k].n : L[k].cap);
            plhs[k + 1] = mxCreateDoubleMatrix(m, 4, mxREAL);
            double *o = mxGetDoubles(plhs[k + 1]);
            for (mwSize q = 0; q < m; ++q) {
                o[q] = (double)L[k].block[q] + 1; o[m + q] = L[k].row[q] + 1; o[2 * m + q] = L[k].col[q] + 1; o[3 * m + q] = L[k].rho[q];
            }
        }
        for (int k = 0; k < 3; ++k) { mxFree(L[k].block); mxFree(L[k].row); mxFree(L[k].col); mxFree(L[k].rho); }
    } else {
        ERR("badCommand", "Unknown command.");
    }
}
