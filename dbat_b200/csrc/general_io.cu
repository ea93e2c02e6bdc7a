// general_io.cu — normal-equation assembly, Schur reduction and back-substitution for a GENERAL IO block
// structure: image-variant parameters (code/demo/romabundledemo_imagevariant.m:44: one principal point per
// image), several cameras (code/script/setdbatcamsandimages.m:28: one block per <camera>), any mixture that
// IO.struct.block / buildserialindices.m:162-221 can express.  multi_res.m:92-111 takes, per image, the IO
// column of that image; here every image carries a map of its NSLOT + 6 camera-side parameters to x columns
// (DevProblem::cam_colx) and to positions in the reduced system (cam_s), and every observation keeps the full
// (NSLOT + 6) x 3 cross block, so nothing assumes that two images share a column.
//
// This path trades speed for generality (one warp per object point, FP64 atomics into the tiles of S); the
// single-shared-block case - every BASELINE config - keeps the specialised kernels of eval.cu / schur.cu.
#include <cstdlib>
#include "kernels.cuh"
#include "launch.h"

__device__ __forceinline__ double gram_g(const double* __restrict__ G, int R, int C) {
    if (R < C) { const int t = R; R = C; C = t; }
    const int p = R >> 3, q = C >> 3, i = R & 7, j = C & 7;
    return G[(p * (p + 1) / 2 + q) * 64 + (i * 4 + (j >> 1)) * 2 + (j & 1)];
}
// same arithmetic as point_inverse in schur.cu
__device__ __forceinline__ void point_inverse_g(const double* __restrict__ rec, const int* __restrict__ opc,
                                                double lambda, double Vi[6]) {
    double a00 = rec[0], a01 = rec[1], a02 = rec[2], a11 = rec[3], a12 = rec[4], a22 = rec[5];
    const bool f0 = opc[0] >= 0, f1 = opc[1] >= 0, f2 = opc[2] >= 0;
    a00 = f0 ? a00 + lambda : 1.0; a11 = f1 ? a11 + lambda : 1.0; a22 = f2 ? a22 + lambda : 1.0;
    const double l00 = sqrt(a00), l10 = a01 / l00, l20 = a02 / l00;
    const double l11 = sqrt(a11 - l10 * l10), l21 = (a12 - l20 * l10) / l11;
    const double l22 = sqrt(a22 - l20 * l20 - l21 * l21);
    const double m00 = 1.0 / l00, m11 = 1.0 / l11, m22 = 1.0 / l22;
    const double m10 = -l10 * m00 * m11, m21 = -l21 * m11 * m22, m20 = -(l20 * m00 + l21 * m10) * m22;
    Vi[0] = m00 * m00 + m10 * m10 + m20 * m20; Vi[1] = m10 * m11 + m20 * m21; Vi[2] = m20 * m22;
    Vi[3] = m11 * m11 + m21 * m21; Vi[4] = m21 * m22; Vi[5] = m22 * m22;
    if (!f0) { Vi[0] = 0.0; Vi[1] = 0.0; Vi[2] = 0.0; }
    if (!f1) { Vi[1] = 0.0; Vi[3] = 0.0; Vi[4] = 0.0; }
    if (!f2) { Vi[2] = 0.0; Vi[4] = 0.0; Vi[5] = 0.0; }
}

// ---------------------------------------------------------------------------------------------
// point side: one thread per observation (the blocks of k_point_side_obs); V_j and g_j summed per point in image
// order, the full cross block of every observation written out
// ---------------------------------------------------------------------------------------------
template <int MODEL>
__global__ void __launch_bounds__(DBAT_PSB) k_point_side_gen(DevProblem P, const int* __restrict__ list, int nList) {
    __shared__ double rows[DBAT_PSB * 9];
    const int tid = threadIdx.x;
    int p0, p1;
    if (list) { p0 = list[blockIdx.x]; p1 = p0 + 1; } else { p0 = P.psb_pt[2 * blockIdx.x]; p1 = P.psb_pt[2 * blockIdx.x + 1]; }
    (void)nList;
    const int ob0 = P.pt_start[p0], nob = P.pt_start[p1] - ob0;
    double V[6] = {0, 0, 0, 0, 0, 0}, gq[3] = {0, 0, 0};
    for (int base = 0; base < nob; base += DBAT_PSB) {          // one pass unless the block is one over-long point
        const int r = base + tid;
        if (r < nob) {
            const int ob = ob0 + r;
            const int j = P.pt_pm[ob];
            const double2 uv = P.uv_pm[ob];
            const double2 is = P.isig_pm[ob];
            const ImgRec g = P.img[P.img_pm[ob]];
            const IORec io = P.io[g.io];
            const double Q[3] = {P.OPval[3 * (size_t)j], P.OPval[3 * (size_t)j + 1], P.OPval[3 * (size_t)j + 2]};
            ObsJac o;
            obs_model<MODEL, true, true>(Q, g, io, P.nK, P.nP, uv.x, uv.y, o);
            double A[2][3];
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const double m = P.op_col[3 * (size_t)j + t] >= 0 ? 1.0 : 0.0;
                A[0][t] = o.dOP[0][t] * is.x * m; A[1][t] = o.dOP[1][t] * is.y * m;
            }
            const double r0 = o.r[0] * is.x, r1 = o.r[1] * is.y;
            double* row = rows + tid * 9;
            row[0] = A[0][0] * A[0][0] + A[1][0] * A[1][0];
            row[1] = A[0][0] * A[0][1] + A[1][0] * A[1][1];
            row[2] = A[0][0] * A[0][2] + A[1][0] * A[1][2];
            row[3] = A[0][1] * A[0][1] + A[1][1] * A[1][1];
            row[4] = A[0][1] * A[0][2] + A[1][1] * A[1][2];
            row[5] = A[0][2] * A[0][2] + A[1][2] * A[1][2];
#pragma unroll
            for (int t = 0; t < 3; ++t) row[6 + t] = A[0][t] * r0 + A[1][t] * r1;
            double* Wf = P.Wfull + (size_t)ob * DBAT_WF_STRIDE;
#pragma unroll
            for (int s = 0; s < DBAT_NSLOT; ++s) {
                const double a0 = o.dIO[s][0] * is.x, a1 = o.dIO[s][1] * is.y;
#pragma unroll
                for (int t = 0; t < 3; ++t) Wf[3 * s + t] = a0 * A[0][t] + a1 * A[1][t];
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double a0 = o.dC[0][c] * is.x, a1 = o.dC[1][c] * is.y;
                const double b0 = o.dA[0][c] * is.x, b1 = o.dA[1][c] * is.y;
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    Wf[3 * (DBAT_NSLOT + c) + t] = a0 * A[0][t] + a1 * A[1][t];
                    Wf[3 * (DBAT_NSLOT + 3 + c) + t] = b0 * A[0][t] + b1 * A[1][t];
                }
            }
        }
        __syncthreads();
        if (list) {                                             // one point: thread 0 adds this pass in order
            if (tid == 0)
                for (int q = 0; q < min(DBAT_PSB, nob - base); ++q) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) V[k] += rows[q * 9 + k];
#pragma unroll
                    for (int t = 0; t < 3; ++t) gq[t] += rows[q * 9 + 6 + t];
                }
            __syncthreads();
        }
    }
    if (list) {
        if (tid == 0) {
            double* rec = P.pt + (size_t)p0 * DBAT_PT_STRIDE;
            for (int k = 0; k < 6; ++k) rec[k] = V[k];
            for (int t = 0; t < 3; ++t) rec[6 + t] = gq[t];
            for (int k = 9; k < DBAT_PT_STRIDE; ++k) rec[k] = 0.0;
        }
        return;
    }
    for (int j = p0 + tid; j < p1; j += DBAT_PSB) {
        double* rec = P.pt + (size_t)j * DBAT_PT_STRIDE;
        double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int r = P.pt_start[j] - ob0; r < P.pt_start[j + 1] - ob0; ++r)
#pragma unroll
            for (int k = 0; k < 9; ++k) acc[k] += rows[r * 9 + k];
#pragma unroll
        for (int k = 0; k < 9; ++k) rec[k] = acc[k];
        for (int k = 9; k < DBAT_PT_STRIDE; ++k) rec[k] = 0.0;    // no shared IO x OP block in this mode
    }
}
void launch_point_side_gen(const DevProblem& P, cudaStream_t st) {
    if (P.nOP <= 0) return;
#define PSG(M) { if (P.nPsb > 0) k_point_side_gen<M><<<P.nPsb, DBAT_PSB, 0, st>>>(P, nullptr, 0); \
                 if (P.nPsbig > 0) k_point_side_gen<M><<<P.nPsbig, DBAT_PSB, 0, st>>>(P, P.psbig, P.nPsbig); }
    switch (P.model) { case 0: PSG(0) break; case 1: PSG(1) break; case 2: PSG(2) break; case 3: PSG(3) break; default: PSG(4) break; }
#undef PSG
    count_launch(2);
}

// ---------------------------------------------------------------------------------------------
// S := N_cc + lambda I, rhs := -g_c from the per-image Grams through the per-image maps.  Entries between two
// parameters of one image that nobody else uses are written; everything that touches a column used by several
// images is accumulated with atomics (S and rhs are zero on entry).
// ---------------------------------------------------------------------------------------------
__global__ void k_build_S_gen(DevProblem P, const double* __restrict__ camDiag, const double* __restrict__ camG,
                              double lambda) {
    const int i = blockIdx.x;
    if (i < P.nImg) {
        const double* G = P.imgG + (size_t)i * DBAT_GSZ;
        const int* cs = P.cam_s + (size_t)DBAT_NCAM * i;
        const int* cx = P.cam_colx + (size_t)DBAT_NCAM * i;
        const int ioS = P.ldS;                                   // S indices >= first global column <=> global
        (void)ioS;
        for (int e = threadIdx.x; e < DBAT_NCAM * (DBAT_NCAM + 1); e += blockDim.x) {
            const int a = e / (DBAT_NCAM + 1), b = e % (DBAT_NCAM + 1);
            const int row = cs[a];
            if (row < 0) continue;
            // Gram column of parameter a: IO slot a, or EO element at DBAT_COL_EO + (a - NSLOT)
            const int ga = a < DBAT_NSLOT ? a : DBAT_COL_EO + a - DBAT_NSLOT;
            if (b < DBAT_NCAM) {
                const int col = cs[b];
                if (col < 0 || col > row || (col == row && b != a)) continue;
                const int gb = b < DBAT_NSLOT ? b : DBAT_COL_EO + b - DBAT_NSLOT;
                atomicAdd(tc_at(P.T, row, col), gram_g(G, ga, gb));
            } else {
                atomicAdd(&P.rhs[row], -gram_g(G, ga, DBAT_COL_R));
            }
        }
        (void)cx;
    } else {
        // once per column: prior terms, damping; padding positions get a unit diagonal
        for (int k = threadIdx.x; k < P.ldS; k += blockDim.x) {
            const int c = P.s2x[k];
            if (c >= 0) {
                atomicAdd(tc_at(P.T, k, k), camDiag[c] + lambda);
                atomicAdd(&P.rhs[k], -camG[c]);
            } else if (k != P.ldS - 1) {
                *tc_at(P.T, k, k) = 1.0;
            }
        }
    }
}
void launch_build_S_gen(const DevProblem& P, const double* camDiag, const double* camG, double lambda, cudaStream_t st) {
    tchol_zero_dev(P.T, st);
    cudaMemsetAsync(P.rhs, 0, sizeof(double) * P.ldS, st);
    k_build_S_gen<<<P.nImg + 1, 128, 0, st>>>(P, camDiag, camG, lambda);
    count_launch();
}

// ---------------------------------------------------------------------------------------------
// Schur update, one warp per object point.  Rows of W~_j: (observation o, parameter a), a over the NSLOT + 6
// camera-side parameters of the observation's image.  For every ordered pair of rows the product
// y_(o,a)' w_(o',b) goes to S[r, c] when r >= c (r, c the S positions): two rows that map to the same
// position (a column shared by two images of the point) then contribute both cross terms to the diagonal,
// and mirrored pairs are counted once.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_schur_gen(DevProblem P, double lambda) {
    const int lane = threadIdx.x & 31;
    const int warpGlobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nWarps = (gridDim.x * blockDim.x) >> 5;
    for (int j = warpGlobal; j < P.nOP; j += nWarps) {
        const int* opc = P.op_col + 3 * (size_t)j;
        if (opc[0] < 0 && opc[1] < 0 && opc[2] < 0) continue;
        const double* rec = P.pt + (size_t)j * DBAT_PT_STRIDE;
        double Vi[6];
        point_inverse_g(rec, opc, lambda, Vi);
        const double gj[3] = {rec[6], rec[7], rec[8]};
        const int o0 = P.pt_start[j], k = P.pt_start[j + 1] - o0;
        const int nRows = DBAT_NCAM * k;
        for (int ra = lane; ra < nRows; ra += 32) {
            const int o = ra / DBAT_NCAM, a = ra - o * DBAT_NCAM;
            const int row = P.cam_s[(size_t)DBAT_NCAM * P.img_pm[o0 + o] + a];
            if (row < 0) continue;
            const double* wa = P.Wfull + (size_t)(o0 + o) * DBAT_WF_STRIDE + 3 * a;
            const double ya0 = Vi[0] * wa[0] + Vi[1] * wa[1] + Vi[2] * wa[2];
            const double ya1 = Vi[1] * wa[0] + Vi[3] * wa[1] + Vi[4] * wa[2];
            const double ya2 = Vi[2] * wa[0] + Vi[4] * wa[1] + Vi[5] * wa[2];
            atomicAdd(&P.rhs[row], ya0 * gj[0] + ya1 * gj[1] + ya2 * gj[2]);
            for (int rb = 0; rb < nRows; ++rb) {
                const int o2 = rb / DBAT_NCAM, b = rb - o2 * DBAT_NCAM;
                const int col = P.cam_s[(size_t)DBAT_NCAM * P.img_pm[o0 + o2] + b];
                if (col < 0 || col > row) continue;
                const double* wb = P.Wfull + (size_t)(o0 + o2) * DBAT_WF_STRIDE + 3 * b;
                const double v = ya0 * wb[0] + ya1 * wb[1] + ya2 * wb[2];
                if (v != 0.0) atomicAdd(tc_at(P.T, row, col), -v);
            }
        }
    }
}
void launch_schur_gen(const DevProblem& P, double lambda, cudaStream_t st) {
    if (P.nOP <= 0) return;
    k_schur_gen<<<std::min((P.nOP + 7) / 8, 148 * 8), 256, 0, st>>>(P, lambda);
    count_launch();
}

// ---------------------------------------------------------------------------------------------
// back-substitution (one thread per point) with the |Jp|^2 / r'Jp by-products of k_backsub
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_backsub_gen(DevProblem P, double lambda, double* __restrict__ p,
                                                      double* __restrict__ stats, int statStride) {
    __shared__ double red[2][4];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    double jp2 = 0.0, rjp = 0.0;
    if (j < P.nOP) {
        const int* opc = P.op_col + 3 * (size_t)j;
        if (!(opc[0] < 0 && opc[1] < 0 && opc[2] < 0)) {
            const double* rec = P.pt + (size_t)j * DBAT_PT_STRIDE;
            double Vi[6];
            point_inverse_g(rec, opc, lambda, Vi);
            double a[3] = {0.0, 0.0, 0.0};
            for (int ob = P.pt_start[j]; ob < P.pt_start[j + 1]; ++ob) {
                const int* cx = P.cam_colx + (size_t)DBAT_NCAM * P.img_pm[ob];
                const double* Wf = P.Wfull + (size_t)ob * DBAT_WF_STRIDE;
                for (int e = 0; e < DBAT_NCAM; ++e) {
                    const int c = cx[e];
                    if (c < 0) continue;
                    const double pv = p[c];
                    a[0] += Wf[3 * e] * pv; a[1] += Wf[3 * e + 1] * pv; a[2] += Wf[3 * e + 2] * pv;
                }
            }
            const double t[3] = {-rec[6] - a[0], -rec[7] - a[1], -rec[8] - a[2]};
            double y[3];
            y[0] = Vi[0] * t[0] + Vi[1] * t[1] + Vi[2] * t[2];
            y[1] = Vi[1] * t[0] + Vi[3] * t[1] + Vi[4] * t[2];
            y[2] = Vi[2] * t[0] + Vi[4] * t[1] + Vi[5] * t[2];
#pragma unroll
            for (int e = 0; e < 3; ++e) if (opc[e] >= 0) p[opc[e]] = y[e]; else y[e] = 0.0;
            const double Vy0 = rec[0] * y[0] + rec[1] * y[1] + rec[2] * y[2];
            const double Vy1 = rec[1] * y[0] + rec[3] * y[1] + rec[4] * y[2];
            const double Vy2 = rec[2] * y[0] + rec[4] * y[1] + rec[5] * y[2];
            jp2 = y[0] * (Vy0 + 2.0 * a[0]) + y[1] * (Vy1 + 2.0 * a[1]) + y[2] * (Vy2 + 2.0 * a[2]);
            rjp = rec[6] * y[0] + rec[7] * y[1] + rec[8] * y[2];
        }
    }
    if (stats) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { jp2 += __shfl_xor_sync(0xffffffffu, jp2, o); rjp += __shfl_xor_sync(0xffffffffu, rjp, o); }
        if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = jp2; red[1][threadIdx.x >> 5] = rjp; }
        __syncthreads();
        if (threadIdx.x == 0) {
            stats[blockIdx.x] = (red[0][0] + red[0][1]) + (red[0][2] + red[0][3]);
            stats[statStride + blockIdx.x] = (red[1][0] + red[1][1]) + (red[1][2] + red[1][3]);
        }
    }
}
int launch_backsub_gen(const DevProblem& P, double lambda, double* p, double* stats, cudaStream_t st) {
    const int nb = (P.nOP + 127) / 128;
    if (P.nOP > 0) { k_backsub_gen<<<nb, 128, 0, st>>>(P, lambda, p, stats, nb); count_launch(); }
    return nb;
}

// ---------------------------------------------------------------------------------------------
// COP: 3 x 3 posterior covariance blocks, one warp per point (the rows of k_cop through the per-image maps)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_cop_gen(DevProblem P, const double* __restrict__ C, int ldc,
                                                  const double* __restrict__ dsc, double s02, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= P.nOP) return;
    const int* opc = P.op_col + 3 * (size_t)j;
    if (opc[0] < 0 && opc[1] < 0 && opc[2] < 0) return;
    const double* rec = P.pt + (size_t)j * DBAT_PT_STRIDE;
    double Vi[6];
    point_inverse_g(rec, opc, 0.0, Vi);
    const int o0 = P.pt_start[j], k = P.pt_start[j + 1] - o0;
    const int nRows = DBAT_NCAM * k;
    auto row_s = [&](int a) { return P.cam_s[(size_t)DBAT_NCAM * P.img_pm[o0 + a / DBAT_NCAM] + a % DBAT_NCAM]; };
    auto row_T = [&](int a, double T[3]) {
        const double* w = P.Wfull + (size_t)(o0 + a / DBAT_NCAM) * DBAT_WF_STRIDE + 3 * (a % DBAT_NCAM);
        T[0] = Vi[0] * w[0] + Vi[1] * w[1] + Vi[2] * w[2];
        T[1] = Vi[1] * w[0] + Vi[3] * w[1] + Vi[4] * w[2];
        T[2] = Vi[2] * w[0] + Vi[4] * w[1] + Vi[5] * w[2];
    };
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int a = lane; a < nRows; a += 32) {
        const int ca = row_s(a);
        if (ca < 0) continue;
        double Ta[3]; row_T(a, Ta);
        double u[3] = {0, 0, 0};
        for (int b = 0; b < nRows; ++b) {
            const int cb = row_s(b);
            if (cb < 0) continue;
            double Tb[3]; row_T(b, Tb);
            const double c = C[(size_t)cb * ldc + ca] * dsc[ca] * dsc[cb];
            u[0] += c * Tb[0]; u[1] += c * Tb[1]; u[2] += c * Tb[2];
        }
        acc[0] += Ta[0] * u[0]; acc[1] += Ta[0] * u[1]; acc[2] += Ta[0] * u[2];
        acc[3] += Ta[1] * u[1]; acc[4] += Ta[1] * u[2]; acc[5] += Ta[2] * u[2];
    }
#pragma unroll
    for (int q = 0; q < 6; ++q)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
    if (lane == 0) {
        double* o9 = out + 9 * (size_t)j;
        const double c00 = s02 * (Vi[0] + acc[0]), c01 = s02 * (Vi[1] + acc[1]), c02 = s02 * (Vi[2] + acc[2]);
        const double c11 = s02 * (Vi[3] + acc[3]), c12 = s02 * (Vi[4] + acc[4]), c22 = s02 * (Vi[5] + acc[5]);
        o9[0] = c00; o9[1] = c01; o9[2] = c02; o9[3] = c01; o9[4] = c11; o9[5] = c12; o9[6] = c02; o9[7] = c12; o9[8] = c22;
    }
}
void launch_cop_gen(const DevProblem& P, const double* C, int ldc, const double* dsc, double s02, double* out, cudaStream_t st) {
    if (P.nOP > 0) { k_cop_gen<<<(P.nOP + 3) / 4, 128, 0, st>>>(P, C, ldc, dsc, s02, out); count_launch(); }
}

// diag(J'J) and gradient entries of the IO columns in general mode: sums over the images that use a column
__global__ void k_io_diag_grad_gen(DevProblem P, const double* __restrict__ camDiag, const double* __restrict__ camG,
                                   double* __restrict__ diagN, double* __restrict__ grad) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.nImg * DBAT_NSLOT) return;
    const int i = t / DBAT_NSLOT, s = t - i * DBAT_NSLOT;
    const int c = P.cam_colx[(size_t)DBAT_NCAM * i + s];
    if (c < 0) return;
    const double* G = P.imgG + (size_t)i * DBAT_GSZ;
    if (diagN) atomicAdd(&diagN[c], gram_g(G, s, s));
    if (grad) atomicAdd(&grad[c], gram_g(G, s, DBAT_COL_R));
    (void)camDiag; (void)camG;
}
void launch_io_diag_grad_gen(const DevProblem& P, const double* camDiag, const double* camG, double* diagN, double* grad,
                             cudaStream_t st) {
    const int n = P.nImg * DBAT_NSLOT;
    k_io_diag_grad_gen<<<(n + 255) / 256, 256, 0, st>>>(P, camDiag, camG, diagN, grad);
    count_launch();
}
