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
