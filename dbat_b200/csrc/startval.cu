// startval.cu — start-value step that precedes bundle() in every demo:
// forward intersection of object points from known cameras, one thread per point.
// Replaces code/photogrammetry/forwintersect.m:19-46 -> pm_multiforwintersect.m:15-51 (loop over the
// distinct camera combinations) -> pm_forwintersect3.m:11-73 (one dense least-squares solve per
// point) and the lens correction pm_multilenscorr1.m:36-69 / pm_lens1.m:38-72.
//
// Per ray: q = pxSize .* [u; -v] (mm), xy = q - lens(q), direction t ∝ M (x - ppx, y - ppy, -f) with
// M = R1(ω)R2(φ)R3(κ) (= RR' of pm_eulerrotmat; the reference gets the same ray from pinv(P) x).
// The reference minimises || [I t_i][p; α_i] - C_i || over p and the ray parameters; eliminating the
// α_i gives the 3x3 system  Σ(I - t t') p = Σ(I - t t') C_i  solved here, and its residual norm
// sqrt(Σ ||(I - t t')(C_i - p)||²)/n is the reference's `res`.
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/dbat_gpu.h"
#include "launch.h"

namespace {
struct FwiDev {
    const double* IO; const double* EO; const double* px; int NC, nK, nP;
    const double2* uv; const int* img; const int* pt_start;     // point-major observations
    const int* pts; int nPts;
    double* OP; double* res;
};

__device__ __forceinline__ void fwi_ray(const FwiDev& D, int o, double t[3], double C[3]) {
    const int i = D.img[o];
    const double* io = D.IO + (size_t)i * D.NC;
    const double* eo = D.EO + (size_t)i * 6;
    const double2 u = D.uv[o];
    const double qx = D.px[2 * i] * u.x, qy = -D.px[2 * i + 1] * u.y;           // diag([1,-1]) then pixel size
    const double xb = qx - io[1], yb = qy - io[2];
    const double r2 = xb * xb + yb * yb;
    double Kr = 0.0, pw = 1.0;
    for (int k = 0; k < D.nK; ++k) { pw *= r2; Kr += io[5 + k] * pw; }          // pm_lens1.m:44-53
    double dx = xb * Kr, dy = yb * Kr;
    if (D.nP > 0) {                                                            // pm_lens1.m:61-69
        const double P1 = io[5 + D.nK], P2 = D.nP > 1 ? io[6 + D.nK] : 0.0, P3 = D.nP > 2 ? io[7 + D.nK] : 0.0;
        dx += (P1 * (r2 + 2 * xb * xb) + 2 * P2 * xb * yb) * (1 + P3);
        dy += (P2 * (r2 + 2 * yb * yb) + 2 * P1 * xb * yb) * (1 + P3);
    }
    const double x = qx - dx - io[1], y = qy - dy - io[2], z = -io[0];         // K^-1 [x;y;1] up to scale
    double sw, cw, sp, cp, sk, ck;
    sincos(eo[3], &sw, &cw); sincos(eo[4], &sp, &cp); sincos(eo[5], &sk, &ck);
    const double M00 = cp * ck, M01 = -cp * sk, M02 = sp;
    const double M10 = cw * sk + sw * sp * ck, M11 = cw * ck - sw * sp * sk, M12 = -sw * cp;
    const double M20 = sw * sk - cw * sp * ck, M21 = sw * ck + cw * sp * sk, M22 = cw * cp;
    double d0 = M00 * x + M01 * y + M02 * z, d1 = M10 * x + M11 * y + M12 * z, d2 = M20 * x + M21 * y + M22 * z;
    const double inv = 1.0 / sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    t[0] = d0 * inv; t[1] = d1 * inv; t[2] = d2 * inv;
    C[0] = eo[0]; C[1] = eo[1]; C[2] = eo[2];
}

__global__ void __launch_bounds__(128) k_forwintersect(FwiDev D) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= D.nPts) return;
    const int j = D.pts[k];
    const int o0 = D.pt_start[j], o1 = D.pt_start[j + 1], n = o1 - o0;
    const double nanv = __longlong_as_double(0x7ff8000000000000LL);
    if (n < 2) {                                                               // pm_multiforwintersect.m:41
        D.OP[3 * (size_t)k] = nanv; D.OP[3 * (size_t)k + 1] = nanv; D.OP[3 * (size_t)k + 2] = nanv;
        if (D.res) D.res[k] = nanv;
        return;
    }
    double a00 = 0, a01 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0, b0 = 0, b1 = 0, b2 = 0;
    for (int o = o0; o < o1; ++o) {
        double t[3], C[3];
        fwi_ray(D, o, t, C);
        const double tc = t[0] * C[0] + t[1] * C[1] + t[2] * C[2];
        a00 += 1.0 - t[0] * t[0]; a01 -= t[0] * t[1]; a02 -= t[0] * t[2];
        a11 += 1.0 - t[1] * t[1]; a12 -= t[1] * t[2]; a22 += 1.0 - t[2] * t[2];
        b0 += C[0] - t[0] * tc; b1 += C[1] - t[1] * tc; b2 += C[2] - t[2] * tc;
    }
    // 3x3 Cholesky solve
    const double l00 = sqrt(a00), l10 = a01 / l00, l20 = a02 / l00;
    const double l11 = sqrt(a11 - l10 * l10), l21 = (a12 - l20 * l10) / l11;
    const double l22 = sqrt(a22 - l20 * l20 - l21 * l21);
    const double y0 = b0 / l00, y1 = (b1 - l10 * y0) / l11, y2 = (b2 - l20 * y0 - l21 * y1) / l22;
    const double p2 = y2 / l22, p1 = (y1 - l21 * p2) / l11, p0 = (y0 - l10 * p1 - l20 * p2) / l00;
    D.OP[3 * (size_t)k] = p0; D.OP[3 * (size_t)k + 1] = p1; D.OP[3 * (size_t)k + 2] = p2;
    if (D.res) {
        double ss = 0.0;
        for (int o = o0; o < o1; ++o) {
            double t[3], C[3];
            fwi_ray(D, o, t, C);
            const double e0 = C[0] - p0, e1 = C[1] - p1, e2 = C[2] - p2;
            const double te = t[0] * e0 + t[1] * e1 + t[2] * e2;
            const double r0 = e0 - t[0] * te, r1 = e1 - t[1] * te, r2 = e2 - t[2] * te;
            ss += r0 * r0 + r1 * r1 + r2 * r2;
        }
        D.res[k] = sqrt(ss) / n;                                               // pm_forwintersect3.m:71
    }
}
}  // namespace

static std::string g_fwi_err;
extern "C" const char* dbat_forwintersect_error(void) { return g_fwi_err.c_str(); }

extern "C" int dbat_forwintersect(const dbat_fwi_desc* d, double* OP, double* res, double* kernel_ms) {
    if (!d || !OP || d->nImg <= 0 || d->nOP <= 0 || d->nObs < 0 || d->nPts < 0 || d->NC < 5 + d->nK + d->nP ||
        !d->IO || !d->EO || !d->pxSize || (d->nObs > 0 && (!d->IPval || !d->obs_img || !d->obs_op)) || (d->nPts > 0 && !d->pts)) {
        g_fwi_err = "dbat_forwintersect: bad argument"; return DBAT_E_BADARG;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_fwi_err = "dbat_forwintersect: no CUDA device"; return DBAT_E_CUDA; }
    for (int64_t k = 0; k < 6 * d->nImg; ++k) if (!std::isfinite(d->EO[k])) { g_fwi_err = "Bad or uninitialized EO data"; return DBAT_E_BADARG; }
    for (int64_t k = 0; k < d->NC * d->nImg; ++k) if (!std::isfinite(d->IO[k])) { g_fwi_err = "Bad or uninitialized IO data"; return DBAT_E_BADARG; }
    const int nOP = (int)d->nOP, nObs = (int)d->nObs, nImg = (int)d->nImg, nPts = (int)d->nPts;
    // point-major observation list (images in input order inside a point)
    std::vector<int> pt_start(nOP + 1, 0), img(std::max(1, nObs)), pts(std::max(1, nPts));
    std::vector<double2> uv(std::max(1, nObs));
    for (int k = 0; k < nObs; ++k) {
        const int64_t j = d->obs_op[k] - 1, i = d->obs_img[k] - 1;
        if (j < 0 || j >= nOP || i < 0 || i >= nImg) { g_fwi_err = "dbat_forwintersect: observation index out of range"; return DBAT_E_BADARG; }
        pt_start[j + 1]++;
    }
    for (int j = 0; j < nOP; ++j) pt_start[j + 1] += pt_start[j];
    {
        std::vector<int> fill(pt_start.begin(), pt_start.end() - 1);
        for (int k = 0; k < nObs; ++k) {
            const int o = fill[d->obs_op[k] - 1]++;
            img[o] = (int)(d->obs_img[k] - 1);
            uv[o] = make_double2(d->IPval[2 * (size_t)k], d->IPval[2 * (size_t)k + 1]);
        }
    }
    for (int k = 0; k < nPts; ++k) {
        if (d->pts[k] < 1 || d->pts[k] > nOP) { g_fwi_err = "dbat_forwintersect: point number out of range"; return DBAT_E_BADARG; }
        pts[k] = (int)(d->pts[k] - 1);
    }
    if (nPts == 0) return DBAT_OK;
    FwiDev D{};
    std::vector<void*> allocs;
    auto up = [&](const void* h, size_t bytes) -> void* {
        void* p = nullptr;
        if (cudaMalloc(&p, std::max<size_t>(bytes, 16)) != cudaSuccess) return nullptr;
        allocs.push_back(p);
        if (h && bytes) cudaMemcpy(p, h, bytes, cudaMemcpyHostToDevice);
        return p;
    };
    D.IO = (const double*)up(d->IO, sizeof(double) * d->NC * nImg);
    D.EO = (const double*)up(d->EO, sizeof(double) * 6 * nImg);
    D.px = (const double*)up(d->pxSize, sizeof(double) * 2 * nImg);
    D.uv = (const double2*)up(uv.data(), sizeof(double2) * nObs);
    D.img = (const int*)up(img.data(), sizeof(int) * nObs);
    D.pt_start = (const int*)up(pt_start.data(), sizeof(int) * (nOP + 1));
    D.pts = (const int*)up(pts.data(), sizeof(int) * nPts);
    D.OP = (double*)up(nullptr, sizeof(double) * 3 * nPts);
    D.res = res ? (double*)up(nullptr, sizeof(double) * nPts) : nullptr;
    D.NC = (int)d->NC; D.nK = (int)d->nK; D.nP = (int)d->nP; D.nPts = nPts;
    int rc = DBAT_OK;
    if (!D.IO || !D.EO || !D.px || !D.uv || !D.img || !D.pt_start || !D.pts || !D.OP || (res && !D.res)) {
        g_fwi_err = "dbat_forwintersect: out of device memory"; rc = DBAT_E_OOM;
    } else {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_forwintersect<<<(nPts + 127) / 128, 128>>>(D);
        count_launch();
        cudaEventRecord(e1);
        cudaMemcpy(OP, D.OP, sizeof(double) * 3 * nPts, cudaMemcpyDeviceToHost);
        if (res) cudaMemcpy(res, D.res, sizeof(double) * nPts, cudaMemcpyDeviceToHost);
        const cudaError_t e = cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (kernel_ms) *kernel_ms = ms;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (e != cudaSuccess) { g_fwi_err = std::string("dbat_forwintersect: ") + cudaGetErrorString(e); rc = DBAT_E_CUDA; }
    }
    for (void* p : allocs) cudaFree(p);
    return rc;
}


// ---------------------------------------------------------------------------------------------
// Batched 3-point spatial resection with residual test, one thread per camera.
// Replaces the per-camera body of code/photogrammetry/resect.m:96-129 -> pm_resect_3pt.m:38-147
// (Grunert's quartic, absolute orientation of every admissible root, mean reprojection residual over
// the camera's test points, best root) and the conversion of the best camera matrix to EO
// (euclidean(null(P)), derotmat3d.m:19-21).  The choice of the three points (largesttriangle.m) and
// the lens correction of the few control-point measurements stay on the host.
// ---------------------------------------------------------------------------------------------
namespace {
struct Cplx { double re, im; };
__device__ __forceinline__ Cplx cmul(Cplx a, Cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ Cplx csub(Cplx a, Cplx b) { return {a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ Cplx cdiv(Cplx a, Cplx b) {
    const double d = b.re * b.re + b.im * b.im;
    return {(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}
// all four roots of c4 z^4 + ... + c0 (MATLAB `roots`): Durand-Kerner on the monic polynomial
__device__ void quartic_roots(const double c[5], Cplx z[4]) {
    const double a3 = c[3] / c[4], a2 = c[2] / c[4], a1 = c[1] / c[4], a0 = c[0] / c[4];
    const double bound = 1.0 + fmax(fmax(fabs(a3), fabs(a2)), fmax(fabs(a1), fabs(a0)));
    Cplx w = {0.4, 0.9}, pw = {1.0, 0.0};
    for (int i = 0; i < 4; ++i) { z[i] = {pw.re * bound * 0.5, pw.im * bound * 0.5}; pw = cmul(pw, w); }
    for (int it = 0; it < 200; ++it) {
        double moved = 0.0;
        for (int i = 0; i < 4; ++i) {
            // p(z_i) by Horner
            Cplx p = {1.0, 0.0};
            p = cmul(p, z[i]); p.re += a3;
            p = cmul(p, z[i]); p.re += a2;
            p = cmul(p, z[i]); p.re += a1;
            p = cmul(p, z[i]); p.re += a0;
            Cplx q = {1.0, 0.0};
            for (int j = 0; j < 4; ++j) if (j != i) q = cmul(q, csub(z[i], z[j]));
            const Cplx d = cdiv(p, q);
            z[i] = csub(z[i], d);
            moved = fmax(moved, fabs(d.re) + fabs(d.im));
        }
        if (moved < 1e-15 * bound) break;
    }
}
__device__ __forceinline__ void cross3(const double a[3], const double b[3], double c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ void unit3(double a[3]) {
    const double n = 1.0 / sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    a[0] *= n; a[1] *= n; a[2] *= n;
}
// columns r1 = ob/|ob|, r2 = (ob x oc)/|.|, r3 = (ob x (ob x oc))/|.|   (pm_resect_3pt.m:103-118)
__device__ void tri_frame(const double p0[3], const double p1[3], const double p2[3], double R[3][3]) {
    double ob[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
    double oc[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
    double r2[3], r3[3];
    cross3(ob, oc, r2);
    cross3(ob, r2, r3);
    unit3(ob); unit3(r2); unit3(r3);
    for (int i = 0; i < 3; ++i) { R[i][0] = ob[i]; R[i][1] = r2[i]; R[i][2] = r3[i]; }
}

struct ResectDev {
    int nCam, behind;
    const double* X3; const double* x3; const int* tstart; const double* XT; const double* xT;
    double* EO; double* res;
};
__global__ void __launch_bounds__(64) k_resect3(ResectDev D) {
    const int cam = blockIdx.x * blockDim.x + threadIdx.x;
    if (cam >= D.nCam) return;
    const double* X = D.X3 + 9 * (size_t)cam;          // 3 object points, column-major 3x3
    const double* xi = D.x3 + 6 * (size_t)cam;         // 3 normalised image points, 2x3
    double x[3][3];                                    // unit rays
    for (int k = 0; k < 3; ++k) {
        x[k][0] = xi[2 * k]; x[k][1] = xi[2 * k + 1]; x[k][2] = 1.0;
        unit3(x[k]);
    }
    auto dist = [&](int i, int j) {
        const double d0 = X[3 * i] - X[3 * j], d1 = X[3 * i + 1] - X[3 * j + 1], d2 = X[3 * i + 2] - X[3 * j + 2];
        return sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    };
    auto cosang = [&](int i, int j) { return fabs(x[i][0] * x[j][0] + x[i][1] * x[j][1] + x[i][2] * x[j][2]); };   // subspace(): acute
    const double a = dist(1, 2), b = dist(0, 2), c = dist(0, 1);
    const double ca = cosang(1, 2), cb = cosang(0, 2), cg = cosang(0, 1);
    const double b2 = b * b;
    const double q1 = (a * a - c * c) / b2, q2 = (a * a + c * c) / b2, q3 = (b2 - c * c) / b2, q4 = (b2 - a * a) / b2;
    double co[5];
    co[4] = (q1 - 1) * (q1 - 1) - 4 * c * c / b2 * ca * ca;
    co[3] = 4 * (q1 * (1 - q1) * cb + 2 * c * c / b2 * ca * ca * cb - (1 - q2) * ca * cg);
    co[2] = 2 * (q1 * q1 + 2 * q1 * q1 * cb * cb + 2 * q3 * ca * ca + 2 * q4 * cg * cg - 4 * q2 * ca * cb * cg - 1);
    co[1] = 4 * (-q1 * (1 + q1) * cb + 2 * a * a / b2 * cg * cg * cb - (1 - q2) * ca * cg);
    co[0] = (1 + q1) * (1 + q1) - 4 * a * a / b2 * cg * cg;
    Cplx z[4];
    quartic_roots(co, z);
    const int t0 = D.tstart[cam], t1 = D.tstart[cam + 1];
    double best = 1e300, bestR[3][3], bestC[3];
    bool found = false;
    double oR[3][3];
    tri_frame(X, X + 3, X + 6, oR);
    for (int r = 0; r < 4; ++r) {
        const double mag = sqrt(z[r].re * z[r].re + z[r].im * z[r].im);
        if (!(fabs(z[r].im) / mag < 1e-3)) continue;   // pm_resect_3pt.m:74: real roots only
        const double v = z[r].re;
        const double u = ((-1 + q1) * v * v - 2 * q1 * cb * v + 1 + q1) / (2 * (cg - v * ca));
        const double s1 = sqrt(b2 / (1 + v * v - 2 * v * cb));
        const double s3 = v * s1, s2 = u * s1;
        if (!(s1 >= 0 && s2 >= 0 && s3 >= 0)) continue;
        const double sgn = D.behind ? -1.0 : 1.0;
        double cx[3][3];
        const double ss[3] = {s1, s2, s3};
        for (int k = 0; k < 3; ++k) for (int i = 0; i < 3; ++i) cx[k][i] = sgn * ss[k] * x[k][i];
        double cRd[3][3];
        tri_frame(cx[0], cx[1], cx[2], cRd);
        double R[3][3];                                 // cRo = cRdelta * oRdelta'
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R[i][j] = cRd[i][0] * oR[j][0] + cRd[i][1] * oR[j][1] + cRd[i][2] * oR[j][2];
        double C[3];                                    // oxO = X1 - cRo' * cx1
        for (int i = 0; i < 3; ++i) C[i] = X[i] - (R[0][i] * cx[0][0] + R[1][i] * cx[0][1] + R[2][i] * cx[0][2]);
        double ssq = 0.0;
        for (int t = t0; t < t1; ++t) {
            const double d0 = D.XT[3 * (size_t)t] - C[0], d1 = D.XT[3 * (size_t)t + 1] - C[1], d2 = D.XT[3 * (size_t)t + 2] - C[2];
            const double p0 = R[0][0] * d0 + R[0][1] * d1 + R[0][2] * d2;
            const double p1 = R[1][0] * d0 + R[1][1] * d1 + R[1][2] * d2;
            const double p2 = R[2][0] * d0 + R[2][1] * d1 + R[2][2] * d2;
            const double e0 = p0 / p2 - D.xT[2 * (size_t)t], e1 = p1 / p2 - D.xT[2 * (size_t)t + 1];
            ssq += e0 * e0 + e1 * e1;
        }
        const double rr = sqrt(ssq / (t1 - t0));
        if (rr < best) {
            best = rr; found = true;
            for (int i = 0; i < 3; ++i) { bestC[i] = C[i]; for (int j = 0; j < 3; ++j) bestR[i][j] = R[i][j]; }
        }
    }
    double* eo = D.EO + 6 * (size_t)cam;
    const double nanv = __longlong_as_double(0x7ff8000000000000LL);
    if (found) {
        eo[0] = bestC[0]; eo[1] = bestC[1]; eo[2] = bestC[2];
        eo[3] = atan2(-bestR[2][1], bestR[2][2]);       // derotmat3d.m:19-21
        eo[4] = asin(bestR[2][0]);
        eo[5] = atan2(-bestR[1][0], bestR[0][0]);
        D.res[cam] = best;
    } else {
        for (int i = 0; i < 6; ++i) eo[i] = nanv;
        D.res[cam] = nanv;                              // resect.m:120 reports inf/fail; the host mirror maps NaN to failure
    }
}
}  // namespace

extern "C" int dbat_resect3(const dbat_resect_desc* d, double* EO, double* res) {
    if (!d || !EO || !res || d->nCam < 0 || (d->nCam > 0 && (!d->X3 || !d->x3 || !d->test_start))) {
        g_fwi_err = "dbat_resect3: bad argument"; return DBAT_E_BADARG;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_fwi_err = "dbat_resect3: no CUDA device"; return DBAT_E_CUDA; }
    const int nCam = (int)d->nCam;
    if (nCam == 0) return DBAT_OK;
    std::vector<int> ts(nCam + 1);
    for (int k = 0; k <= nCam; ++k) {
        ts[k] = (int)d->test_start[k];
        if (ts[k] < 0 || (k > 0 && ts[k] < ts[k - 1])) { g_fwi_err = "dbat_resect3: test_start must be non-decreasing"; return DBAT_E_BADARG; }
    }
    const int nT = ts[nCam];
    if (nT > 0 && (!d->XT || !d->xT)) { g_fwi_err = "dbat_resect3: bad argument"; return DBAT_E_BADARG; }
    std::vector<void*> allocs;
    auto up = [&](const void* h, size_t bytes) -> void* {
        void* p = nullptr;
        if (cudaMalloc(&p, std::max<size_t>(bytes, 16)) != cudaSuccess) return nullptr;
        allocs.push_back(p);
        if (h && bytes) cudaMemcpy(p, h, bytes, cudaMemcpyHostToDevice);
        return p;
    };
    ResectDev D{};
    D.nCam = nCam; D.behind = d->behind ? 1 : 0;
    D.X3 = (const double*)up(d->X3, sizeof(double) * 9 * nCam);
    D.x3 = (const double*)up(d->x3, sizeof(double) * 6 * nCam);
    D.tstart = (const int*)up(ts.data(), sizeof(int) * (nCam + 1));
    D.XT = (const double*)up(d->XT, sizeof(double) * 3 * nT);
    D.xT = (const double*)up(d->xT, sizeof(double) * 2 * nT);
    D.EO = (double*)up(nullptr, sizeof(double) * 6 * nCam);
    D.res = (double*)up(nullptr, sizeof(double) * nCam);
    int rc = DBAT_OK;
    if (!D.X3 || !D.x3 || !D.tstart || !D.XT || !D.xT || !D.EO || !D.res) { g_fwi_err = "dbat_resect3: out of device memory"; rc = DBAT_E_OOM; }
    else {
        k_resect3<<<(nCam + 63) / 64, 64>>>(D);
        count_launch();
        cudaMemcpy(EO, D.EO, sizeof(double) * 6 * nCam, cudaMemcpyDeviceToHost);
        cudaMemcpy(res, D.res, sizeof(double) * nCam, cudaMemcpyDeviceToHost);
        const cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { g_fwi_err = std::string("dbat_resect3: ") + cudaGetErrorString(e); rc = DBAT_E_CUDA; }
    }
    for (void* p : allocs) cudaFree(p);
    return rc;
}
