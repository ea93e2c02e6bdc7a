// chol.cu — dense FP64 blocked Cholesky of the reduced camera system on the FP64 tensor
// pipe (DMMA mma.sync.m8n8k4.f64), triangular solves and the explicit inverse used by the
// posterior covariances.  Replaces MATLAB's `\` / chol (CHOLMOD) call sites
// code/bundle/lsa/levenberg_marquardt.m:119, gauss_newton_armijo.m:172,
// code/bundle/bundle_cov.m:87.
//
// Right-looking, block size 128:  for k: L_kk = chol(A_kk) (+ inv(L_kk))  (k_potrf128) ;
//   A_ik <- A_ik inv(L_kk)'  (GEMM) ;  A_ij -= L_ik L_jk'  for i>=j>k (GEMM, lower tiles only).
// All GEMMs are instances of one templated cp.async + DMMA kernel ("NT": both operands
// column-major, C = beta*C + alpha*A*B').
#include <cstdio>
#include <cstdlib>
#include "launch.h"

#define NB 128
#define KS 16                 // k-slice per pipeline stage

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N)); }

// C(tile) = beta*C + alpha * A(rows of tile, 0:K) * B(rows of tile col, 0:K)'
// CTA tile BM x BN, one warp per WTM x WTN sub-tile, STAGES-deep cp.async pipeline over 16-wide
// k-slices.  Smem strides BM+4 / BN+4 (= 4 mod 16): the 8-byte fragment loads of a half-warp
// (addresses (lane&3)*stride + lane/4) fall into 16 distinct banks.
//   <128, 64, 64x32 | 32x32, 3 stages, 2 CTAs/SM>  trailing update (one CTA's C read-modify-write
//              epilogue overlaps the other's main loop)
//   < 64, 64, 32x32, 8 stages>  the look-ahead column update on the critical path: 4x the CTAs of
//              the big tile, and with K = 128 every slice is in flight from the start
//   < 32,128, 32x32, 8 stages>  panel solve on the critical path; the CTA owns whole rows and K
//              spans the whole panel, so C may alias A (every A slice is in shared memory before
//              the epilogue)
// tile list: if tri != 0 the 1-D grid enumerates the tiles (ti, tj) with BN*tj <= BM*ti + BM-1 of
// the lower triangle (BM = 2 BN only); otherwise blockIdx.x = ti, blockIdx.y = tj.
template <int BM, int BN, int WTM, int WTN, int STAGES, int MINB>
__global__ void __launch_bounds__((BM / WTM) * (BN / WTN) * 32, MINB)
k_gemm_nt(const double* A, int lda, const double* __restrict__ B, int ldb,
          double* C, int ldc, int K, double alpha, double beta, int tri) {
    constexpr int NT = (BM / WTM) * (BN / WTN) * 32;   // threads
    constexpr int WM = BM / WTM;                  // warps along M
    constexpr int MI = WTM / 8, NJ = WTN / 8;     // 8x8 DMMA tiles per warp
    constexpr int LA = BM + 4, LB = BN + 4;
    constexpr int STAGE = KS * (LA + LB);
    extern __shared__ __align__(16) double sm[];
    int ti, tj;
    if (tri & 1) {
        const int t = blockIdx.x;
        if (BM == 2 * BN) {                       // rows of 2 (ti + 1) tiles
            ti = (int)((sqrtf(4.0f * t + 1.0f) - 1.0f) * 0.5f);
            while ((ti + 1) * (ti + 2) <= t) ++ti;
            while (ti * (ti + 1) > t) --ti;
            tj = t - ti * (ti + 1);
        } else {                                  // square tiles: rows of ti + 1 tiles
            ti = (int)((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
            while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
            while (ti * (ti + 1) / 2 > t) --ti;
            tj = t - ti * (ti + 1) / 2;
        }
    } else { ti = blockIdx.x; tj = blockIdx.y; }
    const double* Ag = A + (size_t)ti * BM;
    const double* Bg = B + (size_t)tj * BN;
    double* Cg = C + (size_t)tj * BN * ldc + (size_t)ti * BM;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp % WM) * WTM;
    const int wn = (warp / WM) * WTN;
    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    const int nk = K / KS;
    auto load_stage = [&](int stage, int kt) {
        double* As = sm + (size_t)stage * STAGE;
        double* Bs = As + KS * LA;
#pragma unroll
        for (int c = 0; c < KS * BM / 2 / NT; ++c) {       // A: KS x BM doubles in 16-byte chunks
            const int chunk = tid + c * NT;
            const int kk = chunk / (BM / 2), m = (chunk % (BM / 2)) * 2;
            cp_async16(As + kk * LA + m, Ag + (size_t)(kt * KS + kk) * lda + m);
        }
#pragma unroll
        for (int c = 0; c < KS * BN / 2 / NT; ++c) {
            const int chunk = tid + c * NT;
            const int kk = chunk / (BN / 2), m = (chunk % (BN / 2)) * 2;
            cp_async16(Bs + kk * LB + m, Bg + (size_t)(kt * KS + kk) * ldb + m);
        }
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) { if (s < nk) load_stage(s, s); cp_async_commit(); }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (kt + STAGES - 1 < nk) load_stage((kt + STAGES - 1) % STAGES, kt + STAGES - 1);
        cp_async_commit();
        const double* As = sm + (size_t)(kt % STAGES) * STAGE;
        const double* Bs = As + KS * LA;
#pragma unroll
        for (int k0 = 0; k0 < KS; k0 += 4) {
            double af[MI], bf[NJ];
            const double* ap = As + (k0 + (lane & 3)) * LA + wm + (lane >> 2);
            const double* bp = Bs + (k0 + (lane & 3)) * LB + wn + (lane >> 2);
#pragma unroll
            for (int i = 0; i < MI; ++i) af[i] = ap[8 * i];
#pragma unroll
            for (int j = 0; j < NJ; ++j) bf[j] = bp[8 * j];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();
    // epilogue: C fragment (row = lane/4, cols 2*(lane%4)+{0,1}) per 8x8 tile
    if (beta == 0.0) {
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                double* p0 = Cg + (size_t)(wn + 8 * j + 2 * (lane & 3)) * ldc + wm + 8 * i + (lane >> 2);
                p0[0] = alpha * acc[i][j][0]; p0[ldc] = alpha * acc[i][j][1];
            }
    } else {
        constexpr int IC = MI > 4 ? 4 : MI;           // row tiles per read-modify-write batch
#pragma unroll
        for (int i0 = 0; i0 < MI; i0 += IC) {
            double cv[IC][NJ][2];
#pragma unroll
            for (int i = 0; i < IC; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const double* p0 = Cg + (size_t)(wn + 8 * j + 2 * (lane & 3)) * ldc + wm + 8 * (i0 + i) + (lane >> 2);
                    cv[i][j][0] = p0[0]; cv[i][j][1] = p0[ldc];
                }
#pragma unroll
            for (int i = 0; i < IC; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    double* p0 = Cg + (size_t)(wn + 8 * j + 2 * (lane & 3)) * ldc + wm + 8 * (i0 + i) + (lane >> 2);
                    p0[0] = beta * cv[i][j][0] + alpha * acc[i0 + i][j][0];
                    p0[ldc] = beta * cv[i][j][1] + alpha * acc[i0 + i][j][1];
                }
        }
    }
}

template <int BM, int BN, int WTM, int WTN, int STAGES, int MINB>
struct GemmCfg {
    static constexpr int threads = (BM / WTM) * (BN / WTN) * 32;
    static constexpr int smem = STAGES * KS * (BM + 4 + BN + 4) * 8;
    static void launch(dim3 grid, cudaStream_t st, const double* A, int lda, const double* B, int ldb,
                       double* C, int ldc, int K, double alpha, double beta, int tri) {
        static thread_local bool done = false;       // per thread = per device (one host thread drives one device)
        if (!done) {
            cudaFuncSetAttribute(k_gemm_nt<BM, BN, WTM, WTN, STAGES, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            done = true;
        }
        k_gemm_nt<BM, BN, WTM, WTN, STAGES, MINB><<<grid, threads, smem, st>>>(A, lda, B, ldb, C, ldc, K, alpha, beta, tri);
        count_launch();
    }
};
typedef GemmCfg<128, 64, 32, 32, 3, 2> GemmBig32;     // 8 warps, two CTAs per SM (DBAT_GEMM_CFG=0)
typedef GemmCfg<64, 64, 32, 32, 3, 3> GemmSq64;       // 4 warps, three CTAs per SM (default)
// Other shapes were timed and dropped (profiles/README.md): 128x64 with 64x32 warp tiles, 128x128,
// 64x64 at four CTAs per SM or four stages, 64x32, and 64x64 with the C tile prefetched into registers.
typedef GemmCfg<64, 64, 32, 32, 8, 1> GemmCol;
typedef GemmCfg<32, 128, 32, 32, 8, 1> GemmPanel;

// mt = number of 128-row tiles, nt = number of 128-column blocks
static void gemm_nt(const double* A, int lda, const double* B, int ldb, double* C, int ldc,
                    int mt, int nt, int K, double alpha, double beta, bool tri, cudaStream_t st) {
    static int cfg = -1;
    if (cfg < 0) { const char* e = getenv("DBAT_GEMM_CFG"); cfg = e ? atoi(e) : 2; }
    if (mt <= 0 || nt <= 0) return;
    const int f = tri ? 1 : 0;
    if (cfg == 0) GemmBig32::launch(tri ? dim3(mt * (mt + 1)) : dim3(mt, 2 * nt), st, A, lda, B, ldb, C, ldc, K, alpha, beta, f);
    else GemmSq64::launch(tri ? dim3(mt * (2 * mt + 1)) : dim3(2 * mt, 2 * nt), st, A, lda, B, ldb, C, ldc, K, alpha, beta, f);
}

#define PB 16                 // panel width inside the 128x128 diagonal block
// ---------------------------------------------------------------------------------------------
// k_potrf128: Cholesky of one 128x128 diagonal block + inverse of its factor, organised around
// the only inherently serial part, the 128 pivots.
//
//   warp 0 ("pivot warp")   D(p): factors the 16x16 diagonal block of panel p in registers
//                           (lane r < 16 holds row r; lanes 16..31 carry the columns of the
//                           identity through the same column sweep, so inv(L_pp) falls out of the
//                           same instruction stream), then after the panel solve updates the next
//                           diagonal block first (U-critical) and goes straight on to D(p+1).
//   warps 1..7 ("workers")  W(p), running under D(p+1): the rest of the rank-16 trailing update
//                           of panel p, row block p of inv(L) (X_p,0:p = -inv(L_pp) L_p,0:p X_0:p,0:p),
//                           and the write-back of the finished panel / inverse rows to global.
//   all warps               R(p): L_ip = A_ip inv(L_pp)' for the 8-row tiles below the block (DMMA).
//
// Shared tile As is column-major with stride 132 (= 4 mod 16: every DMMA fragment load below is
// bank-conflict free).  L lives in the lower triangle; X = inv(L) is stored transposed in the
// strictly upper triangle (X(r,c), r > c, at As[r*PLD2 + c]), its diagonal blocks also zero-padded
// in XdAll for use as DMMA operands.
// ---------------------------------------------------------------------------------------------
#define PLD2 132
#define XLD 20
#define TLD 12
#define POTRF_SMEM_DOUBLES (NB * PLD2 + 8 * PB * XLD + 3 * PB + 32)

__device__ __forceinline__ double rsqrt_nr(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));      // ~20 bits
    double h = d * y, e = fma(-h, y, 1.0);
    y = fma(0.5 * y, e, y);                                        // ~40 bits
    h = d * y; e = fma(-h, y, 1.0);
    return fma(0.5 * y, e, y);                                     // full precision
}

#ifdef POTRF_PROFILE
#define PTRACE(slot) { if (lane == 0 && blk == 1) ptrace[(slot)] = (double)(clock64() - tstart); }
#else
#define PTRACE(slot)
#endif

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bar_sync_named(int id, int n) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive_named(int id, int n) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bulk_store(double* gdst, const double* ssrc, int bytes) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(s), "r"(bytes) : "memory");
}

// D(p): one warp.  As holds the fully updated diagonal block at (c0, c0).  The pivot chain runs
// in LDL' form: v(:,j) is the unscaled column u(:,j), d_j the pivot.  Between two pivots there are
// only: the hardware reciprocal seed y of d_j, one quartic Newton step folded into the scaled
// column entry  w = u(j+1,j)/d_j = (u y)(1+e)(1+e^2), e = 1 - d y  (relative error e^4), the update
// d_{j+1} = a(j+1,j+1) - w u(j+1,j) and one shuffle.  Square roots (L(:,j) = u(:,j)/sqrt(d_j)) are
// off the chain.  Lanes 16..31 carry the columns of the identity through the same sweep: they end
// up holding inv(L_pp) (b_r -= (b_j/d_j) u(r,j) is the same update, x_j = b_j/sqrt(d_j) the same
// scaling).
__device__ __forceinline__ void potrf_diag16(double* As, double* Xd, double* colb, int c0,
                                             int lane, int gcol0, int nvalid, double& lmin, double& lmax,
                                             int& bad) {
    double* spd = colb + 2 * PB;                 // [16] pivots of this block
    const int r = lane & 15;
    const bool isX = lane >= 16;
    double v[PB];
#pragma unroll
    for (int c = 0; c < PB; ++c) {
        const double a = As[(c0 + c) * PLD2 + c0 + r];
        v[c] = isX ? (c == r ? 1.0 : 0.0) : a;
    }
    double d = __shfl_sync(0xffffffffu, v[0], 0);
#pragma unroll
    for (int j = 0; j < PB; ++j) {
        double* cb = colb + (j & 1) * PB;
        if (!isX) cb[r] = v[j];                  // unscaled column j, final since the end of step j-1
        double y;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
        const double e = fma(-d, y, 1.0);        // 1/d = y (1 + e + e^2 + ...), |e| ~ 2^-20
        const double vy = v[j] * y;
        const double dj = d;
        if (j + 1 < PB) {
            // next pivot (lane j+1): a - u^2/d with 1/d = y (1 + t), t = e + e^2  (error e^3);
            // only e -> t -> cand sit between the reciprocal seed and the shuffle
            const double q = v[j] * vy;
            const double base = v[j + 1] - q;
            const double t = fma(e, e, e);
            const double cand = fma(-q, t, base);
            d = __shfl_sync(0xffffffffu, cand, j + 1);
        }
        const double e2 = e * e;
        const double vy1 = fma(vy, e, vy);
        const double w = fma(vy1, e2, vy1);      // lanes r > j: u(r,j)/d_j ; X lanes: b_j/d_j  (error e^4)
        __syncwarp();
#pragma unroll
        for (int c = j + 1; c < PB; ++c) v[c] = fma(-w, cb[c], v[c]);
        if (lane == 0) spd[j] = dj;              // pivot, for the scaling pass below
    }
    // off the chain: 1/sqrt(d_j) is computed once, by lane j, and broadcast for the column scaling
    __syncwarp();
    {
        const double dr = spd[r];
        const double isdr = rsqrt_nr(dr);
        if (!(dr > 0.0)) bad = 1;
        if (!isX && gcol0 + r < nvalid) { const double l = dr * isdr; lmin = fmin(lmin, l); lmax = fmax(lmax, l); }
#pragma unroll
        for (int j = 0; j < PB; ++j) v[j] *= __shfl_sync(0xffffffffu, isdr, j);
    }
    if (!isX) {
#pragma unroll
        for (int c = 0; c < PB; ++c) if (c <= r) As[(c0 + c) * PLD2 + c0 + r] = v[c];
    } else {
#pragma unroll
        for (int j = 0; j < PB; ++j) {
            Xd[j * XLD + r] = v[j];                                   // zero above the diagonal
            if (j > r) As[(c0 + j) * PLD2 + c0 + r] = v[j];          // X(j, r) transposed into the upper triangle
        }
    }
}

// R(p) for NT (1 or 2) 8-row tiles below the diagonal block: L(i, panel) = A(i, panel) * inv(L_pp)',
// one accumulator per k-step (all DMMAs independent).  The result stays in res[e][tn][0..1]
// (C fragments: row lane/4, panel columns 8tn + 2(lane%4) + {0,1}) and is stored to As.
template <int NT>
__device__ __forceinline__ void potrf_rtiles(double* As, const double* Xd, int c0, int q0, int q1, int lane,
                                             double (&res)[2][2][2]) {
    const int fr = lane >> 2, fk = lane & 3;
    const double* pb = Xd + fr * XLD + fk;                               // B(k,n) = Xd(n,k)
    const double* pa = As + (c0 + fk) * PLD2 + c0 + PB + fr;
    double r[NT][2][4][2];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double b0v = pb[4 * k], b1v = pb[8 * XLD + 4 * k];
#pragma unroll
        for (int e = 0; e < NT; ++e) {
            const double a = pa[4 * k * PLD2 + 8 * (e ? q1 : q0)];
            r[e][0][k][0] = r[e][0][k][1] = r[e][1][k][0] = r[e][1][k][1] = 0.0;
            dmma(r[e][0][k][0], r[e][0][k][1], a, b0v);
            dmma(r[e][1][k][0], r[e][1][k][1], a, b1v);
        }
    }
#pragma unroll
    for (int e = 0; e < NT; ++e) {
        double* pc = As + (c0 + 2 * fk) * PLD2 + c0 + PB + 8 * (e ? q1 : q0) + fr;
#pragma unroll
        for (int tn = 0; tn < 2; ++tn) {
            res[e][tn][0] = (r[e][tn][0][0] + r[e][tn][1][0]) + (r[e][tn][2][0] + r[e][tn][3][0]);
            res[e][tn][1] = (r[e][tn][0][1] + r[e][tn][1][1]) + (r[e][tn][2][1] + r[e][tn][3][1]);
            pc[8 * tn * PLD2] = res[e][tn][0];
            pc[(8 * tn + 1) * PLD2] = res[e][tn][1];
        }
    }
}

// U(p), one group of four 8x8 tiles (list positions t0..t0+3 of the trailing lower triangle; the
// first three positions belong to the next diagonal block and are done by the pivot warp):
// A(i,j) -= L(i,panel p) L(j,panel p)', one accumulator per k-step: 16 independent DMMAs.
__device__ __forceinline__ void potrf_update_group(double* As, const unsigned char* tij, int p, int t0, int ntile,
                                                   int lane) {
    const int fr = lane >> 2, fk = lane & 3;
    const int c0 = PB * p, b0 = c0 + PB;
    const double* pa0 = As + (c0 + fk) * PLD2 + b0 + fr;
    int ti[4], tj[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int t = min(t0 + e, ntile - 1);
        ti[e] = tij[2 * t]; tj[e] = tij[2 * t + 1];
    }
    double acc[4][4][2];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            acc[e][k][0] = acc[e][k][1] = 0.0;
            dmma(acc[e][k][0], acc[e][k][1], pa0[4 * k * PLD2 + 8 * ti[e]], pa0[4 * k * PLD2 + 8 * tj[e]]);
        }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        if (t0 + e < ntile) {
            double* pc = As + (b0 + 8 * tj[e] + 2 * fk) * PLD2 + b0 + 8 * ti[e] + fr;
            pc[0] -= (acc[e][0][0] + acc[e][1][0]) + (acc[e][2][0] + acc[e][3][0]);
            pc[PLD2] -= (acc[e][0][1] + acc[e][1][1]) + (acc[e][2][1] + acc[e][3][1]);
        }
    }
}

// inv(L), right-looking by row blocks.  The slot of X(ib, jb) (ib > jb; transposed in the upper
// triangle: X(r,c) at As[r*PLD2 + c]) first accumulates T(ib,jb) = sum_{jb<=kb<ib} L(ib,kb) X(kb,jb).
// One task = one 8-column tile (jb, tn) of row block p:
//   finalise  X(p, tile) = -inv(L_pp) T(p, tile)            (jb < p; needs D(p))
//   propagate T(ib, tile) += L(ib, p) X(p, tile), ib > p    (needs R(p))
__device__ __forceinline__ void potrf_x_task(double* As, const double* XdAll, int p, int x, int lane,
                                             double* __restrict__ out) {
    const int fr = lane >> 2, fk = lane & 3;
    const int c0 = PB * p;
    const int jb = x >> 1, tn = x & 1;
    const int col0 = PB * jb + 8 * tn;
    double bx[4];
    if (jb < p) {
        const double* pt = As + (c0 + fk) * PLD2 + col0 + fr;                  // T(k, n)
        const double* pxa = XdAll + p * PB * XLD + fr * XLD + fk;             // Xd_p(i, k)
        double x0[4][2], x1[4][2];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double b = pt[4 * k * PLD2];
            x0[k][0] = x0[k][1] = x1[k][0] = x1[k][1] = 0.0;
            dmma(x0[k][0], x0[k][1], pxa[4 * k], b);
            dmma(x1[k][0], x1[k][1], pxa[8 * XLD + 4 * k], b);
        }
        const double xa = -((x0[0][0] + x0[1][0]) + (x0[2][0] + x0[3][0]));
        const double xb = -((x0[0][1] + x0[1][1]) + (x0[2][1] + x0[3][1]));
        const double xc = -((x1[0][0] + x1[1][0]) + (x1[2][0] + x1[3][0]));
        const double xd = -((x1[0][1] + x1[1][1]) + (x1[2][1] + x1[3][1]));
        double* px = As + (c0 + fr) * PLD2 + col0 + 2 * fk;                    // X(c0 + i, col) transposed
        px[0] = xa; px[1] = xb;
        px[8 * PLD2] = xc; px[8 * PLD2 + 1] = xd;
        double* po = out + (size_t)(col0 + 2 * fk) * NB + c0 + fr;
        po[0] = xa; po[NB] = xb;
        po[8] = xc; po[NB + 8] = xd;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; ++k) bx[k] = pt[4 * k * PLD2];                 // X(p)(k, n), now final
    } else {
        const double* pd = XdAll + p * PB * XLD + fk * XLD + 8 * tn + fr;     // Xd_p(k, 8tn + n)
#pragma unroll
        for (int k = 0; k < 4; ++k) bx[k] = pd[4 * k * XLD];
    }
    const double* pa = As + (c0 + fk) * PLD2 + fr;                             // L(row, c0 + k)
#pragma unroll 2
    for (int ib = p + 1; ib < NB / PB; ++ib) {
        double a0[4][2], a1[4][2];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            a0[k][0] = a0[k][1] = a1[k][0] = a1[k][1] = 0.0;
            dmma(a0[k][0], a0[k][1], pa[4 * k * PLD2 + PB * ib], bx[k]);
            dmma(a1[k][0], a1[k][1], pa[4 * k * PLD2 + PB * ib + 8], bx[k]);
        }
        double* pc = As + (PB * ib + fr) * PLD2 + col0 + 2 * fk;
        const double s00 = (a0[0][0] + a0[1][0]) + (a0[2][0] + a0[3][0]), s01 = (a0[0][1] + a0[1][1]) + (a0[2][1] + a0[3][1]);
        const double s10 = (a1[0][0] + a1[1][0]) + (a1[2][0] + a1[3][0]), s11 = (a1[0][1] + a1[1][1]) + (a1[2][1] + a1[3][1]);
        if (jb == p) { pc[0] = s00; pc[1] = s01; pc[8 * PLD2] = s10; pc[8 * PLD2 + 1] = s11; }
        else { pc[0] += s00; pc[1] += s01; pc[8 * PLD2] += s10; pc[8 * PLD2 + 1] += s11; }
    }
}

// W(p): the tasks of one round, dealt round-robin to the nW workers (X tasks first, they are longer;
// the U groups continue the same round-robin so that the load stays even).
__device__ __forceinline__ void potrf_worker(double* As, const double* XdAll, const unsigned char* tij,
                                             int p, int wid, int nW, int lane, double* __restrict__ out) {
    const int m = 14 - 2 * p;
    const int ntile = m * (m + 1) / 2;
    const int nX = p < NB / PB - 1 ? 2 * (p + 1) : 2 * p;       // the last row block has nothing to propagate to
    const int nU = ntile > 3 ? (ntile - 3 + 3) / 4 : 0;
    for (int task = wid; task < nX + nU; task += nW) {
        if (task < nX) potrf_x_task(As, XdAll, p, task, lane, out);
        else potrf_update_group(As, tij, p, 3 + 4 * (task - nX), ntile, lane);
    }
}

// IO warp, after panel p is final: its columns go back to global with one bulk copy each
// (lower triangle; for odd columns the copy starts one row above the diagonal to stay 16-byte
// aligned - the strictly upper triangle of a diagonal block of A is scratch), and the 16x16
// diagonal block of inv(L).  The zeros of inv(L) above the diagonal are never written: the
// buffer is cleared once at allocation.
__device__ __forceinline__ void potrf_io(const double* As, const double* XdAll, int p, int lane,
                                         double* __restrict__ A, int lda, double* __restrict__ out) {
    const int c0 = PB * p;
    if (lane < PB) {
        const int c = c0 + lane, rs = c & ~1;
        bulk_store(A + (size_t)c * lda + rs, As + c * PLD2 + rs, (NB - rs) * 8);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int idx = lane + 32 * k, i = idx & 15, cc = idx >> 4;
        if (cc <= i) out[(size_t)(c0 + cc) * NB + c0 + i] = XdAll[p * PB * XLD + i * XLD + cc];
    }
}

#define POTRF_THREADS 512
__global__ void __launch_bounds__(POTRF_THREADS, 1)
k_potrf128(double* __restrict__ A, int lda, double* __restrict__ invL, int blk, int nvalid,
           int* __restrict__ info, double* __restrict__ minmax) {
    extern __shared__ __align__(16) double sm[];
    double* As = sm;                          // [128 columns][PLD2]
    double* XdAll = As + NB * PLD2;           // [8][16][XLD]
    double* colb = XdAll + 8 * PB * XLD;      // [2][16]
    unsigned char* tij = (unsigned char*)(colb + 3 * PB);   // [105][2] (ti, tj) of the lower-triangular tile list
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    double* out = invL + (size_t)blk * NB * NB;
    double lmin = 1e300, lmax = 0.0;
    int bad = 0;
#ifdef POTRF_PROFILE
    const long long tstart = clock64();
    double* ptrace = minmax + 16 + warp * 40;     // per warp: 4 stamps per panel
#endif
    // Roles (16 warps, 4 per scheduler).  Warp 0: pivots.  Warp 4 only moves data (write-backs).
    // The other 14 warps are the DMMA workers (measured slightly faster than keeping warps 8 and 12,
    // which share the pivot warp's scheduler, idle: -DPOTRF_W12).
#ifndef POTRF_W12
    const bool isPivot = warp == 0, isIO = warp == 4, isWorker = !isPivot && !isIO;
    const int wid = warp < 4 ? warp - 1 : warp - 2;          // worker id 0..13
#define POTRF_NW 14
#else                                            // keep the pivot warp's scheduler free of DMMA work
    const bool isPivot = warp == 0, isIO = warp == 4, isWorker = (warp & 3) != 0;
    const int wid = (warp >> 2) * 3 + (warp & 3) - 1;        // worker id 0..11 (workers only)
#define POTRF_NW 12
#endif
    // ---- load the lower triangle (16-byte chunks).  The pivot warp fetches only the first
    //      diagonal block and starts as soon as that has landed.
    if (isPivot) {
        for (int idx = lane; idx < PB * (PB / 2); idx += 32) {
            const int c = idx >> 3, ci = idx & 7;
            cp_async16(As + c * PLD2 + 2 * ci, A + (size_t)c * lda + 2 * ci);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
    } else {
        for (int idx = t - 32; idx < NB * (NB / 2); idx += POTRF_THREADS - 32) {
            const int c = idx >> 6, ci = idx & 63;
            if (2 * ci + 1 >= c && !(c < PB && ci < PB / 2))
                cp_async16(As + c * PLD2 + 2 * ci, A + (size_t)c * lda + 2 * ci);
        }
        cp_async_commit();
        const int u = t - 32;
        if (u < 105) {
            int i = (int)((sqrtf(8.0f * u + 1.0f) - 1.0f) * 0.5f);
            while ((i + 1) * (i + 2) / 2 <= u) ++i;
            while (i * (i + 1) / 2 > u) --i;
            tij[2 * u] = (unsigned char)i; tij[2 * u + 1] = (unsigned char)(u - i * (i + 1) / 2);
        }
    }
    for (int p = 0; p < NB / PB; ++p) {
        const int c0 = PB * p;
        const int m = 14 - 2 * p;                 // 8-row tiles below the diagonal block
        if (isPivot) {
            potrf_diag16(As, XdAll + p * PB * XLD, colb, c0, lane, blk * NB + c0, nvalid, lmin, lmax, bad);
        } else if (p > 0) {
            if (isIO) potrf_io(As, XdAll, p - 1, lane, A, lda, out);
            else if (isWorker) potrf_worker(As, XdAll, tij, p - 1, wid, POTRF_NW, lane, out);
        } else {
            cp_async_wait<0>();
        }
        PTRACE(4 * p + 0)
        fence_async_smem();
        __syncthreads();                          // S1: D(p) and W(p-1) complete
        if (isPivot) {
            // the two row tiles the next diagonal block depends on, then that block's update
            // straight from the result registers: with the panel columns taken in the order the
            // C fragments hold them, (lane/4, lane%4) has the A and the B operand of every k-step
            if (m >= 2) {
                const int fr = lane >> 2, fk = lane & 3;
                double R[2][2][2];
                potrf_rtiles<2>(As, XdAll + p * PB * XLD, c0, 0, 1, lane, R);
                fence_async_smem();
                bar_arrive_named(1, POTRF_THREADS);
                PTRACE(4 * p + 1)
                double u[3][4][2];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double a0 = R[0][k >> 1][k & 1], a1 = R[1][k >> 1][k & 1];
                    u[0][k][0] = u[0][k][1] = u[1][k][0] = u[1][k][1] = u[2][k][0] = u[2][k][1] = 0.0;
                    dmma(u[0][k][0], u[0][k][1], a0, a0);
                    dmma(u[1][k][0], u[1][k][1], a1, a0);
                    dmma(u[2][k][0], u[2][k][1], a1, a1);
                }
                const int b0 = c0 + PB;
                double* pc = As + (b0 + 2 * fk) * PLD2 + b0 + fr;
                pc[0] -= (u[0][0][0] + u[0][1][0]) + (u[0][2][0] + u[0][3][0]);
                pc[PLD2] -= (u[0][0][1] + u[0][1][1]) + (u[0][2][1] + u[0][3][1]);
                pc[8] -= (u[1][0][0] + u[1][1][0]) + (u[1][2][0] + u[1][3][0]);
                pc[PLD2 + 8] -= (u[1][0][1] + u[1][1][1]) + (u[1][2][1] + u[1][3][1]);
                pc[8 * PLD2 + 8] -= (u[2][0][0] + u[2][1][0]) + (u[2][2][0] + u[2][3][0]);
                pc[9 * PLD2 + 8] -= (u[2][0][1] + u[2][1][1]) + (u[2][2][1] + u[2][3][1]);
                __syncwarp();
            } else {
                bar_arrive_named(1, POTRF_THREADS);
            }
            PTRACE(4 * p + 2)
        } else {
            if (isWorker && 2 + wid < m) {        // the other row tiles, one per worker (at most 12)
                double R[2][2][2];
                potrf_rtiles<1>(As, XdAll + p * PB * XLD, c0, 2 + wid, 2 + wid, lane, R);
            }
            fence_async_smem();
            PTRACE(4 * p + 1)
            bar_sync_named(1, POTRF_THREADS);     // S2: panel p final (the pivot warp only arrives)
        }
    }
    // tail: finalise row block 7 of the inverse, last diagonal block (IO warp)
    if (isIO) {
        potrf_io(As, XdAll, NB / PB - 1, lane, A, lda, out);
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    } else if (isWorker || isPivot) potrf_worker(As, XdAll, tij, NB / PB - 1, isPivot ? POTRF_NW : wid, POTRF_NW + 1, lane, out);
    PTRACE(4 * 8 + 0)
    if (isPivot) {                               // per-lane pivot statistics -> lane 0
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lmin = fmin(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
            lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        }
        bad = __any_sync(0xffffffffu, bad);
    }
    if (t == 0) {
        if (bad) atomicCAS(info, 0, blk + 1);
        double omin = minmax[0], omax = minmax[1];
        if (blk == 0) { omin = 1e300; omax = 0.0; }
        minmax[0] = fmin(omin, lmin); minmax[1] = fmax(omax, lmax);
    }
}

void chol_alloc(CholWork& w, int n, int ld) {
    w.n = n; w.ld = ld; w.nb = ld / NB;
    cudaMalloc(&w.invL, sizeof(double) * (size_t)w.nb * NB * NB);
    cudaMemset(w.invL, 0, sizeof(double) * (size_t)w.nb * NB * NB);   // k_potrf128 never writes the zeros above the diagonal
    cudaMalloc(&w.info, sizeof(int));
    cudaMalloc(&w.minmax, sizeof(double) * (16 + 16 * 40));
}
void chol_free(CholWork& w) {
    if (w.invL) cudaFree(w.invL);
    if (w.info) cudaFree(w.info);
    if (w.minmax) cudaFree(w.minmax);
    if (w.graphExec) cudaGraphExecDestroy((cudaGraphExec_t)w.graphExec);
    w = CholWork();
}

static thread_local cudaStream_t g_aux = nullptr;
static thread_local cudaEvent_t g_evA = nullptr, g_evB = nullptr;

// Right-looking blocked Cholesky with one step of look-ahead: as soon as block column k+1 has
// received the update of step k, its potrf + panel solve run on a second stream while the main
// stream finishes the rest of the trailing update of step k.
static void chol_factor_body(CholWork& w, double* A, cudaStream_t st) {
    static thread_local bool attr = false;
    const int psmem = POTRF_SMEM_DOUBLES * 8;
    if (!attr) {
        cudaFuncSetAttribute(k_potrf128, cudaFuncAttributeMaxDynamicSharedMemorySize, psmem);
        {   // the look-ahead stream carries the critical path: its CTAs must be dispatched ahead of the
            // remaining CTAs of the trailing update running on the main stream
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            cudaStreamCreateWithPriority(&g_aux, cudaStreamNonBlocking, hi);
        }
        cudaEventCreateWithFlags(&g_evA, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&g_evB, cudaEventDisableTiming);
        attr = true;
    }
    cudaMemsetAsync(w.info, 0, sizeof(int), st);
    const int ld = w.ld, nb = w.nb;
    auto diag = [&](int k) { return A + (size_t)k * NB * ld + (size_t)k * NB; };
    auto panel_step = [&](int k, cudaStream_t s) {        // potrf(k) + L_ik = A_ik inv(L_kk)' (in place)
        k_potrf128<<<1, POTRF_THREADS, psmem, s>>>(diag(k), ld, w.invL, k, w.n, w.info, w.minmax);
        count_launch();
        const int rem = nb - k - 1;
        if (rem <= 0) return;
        // in place: each CTA owns 32 whole rows of the panel
        GemmPanel::launch(dim3(rem * NB / 32, 1), s, diag(k) + NB, ld, w.invL + (size_t)k * NB * NB, NB,
                          diag(k) + NB, ld, NB, 1.0, 0.0, 0);
    };
    // every potrf + panel solve runs on the aux stream under the trailing update (K = 128) of the
    // previous step.  (Taking the steps in pairs, with K = 256 trailing updates, halves the C traffic
    // but leaves every second panel step without anything to overlap with: measured slower.)
    panel_step(0, st);
    for (int k = 0; k + 1 < nb; ++k) {
        const int rem = nb - k - 1;
        double* Apanel = diag(k) + NB;
        GemmCol::launch(dim3(rem * NB / 64, NB / 64), st, Apanel, ld, Apanel, ld, diag(k + 1), ld, NB, -1.0, 1.0, 0);
        cudaEventRecord(g_evA, st);
        cudaStreamWaitEvent(g_aux, g_evA, 0);
        panel_step(k + 1, g_aux);
        cudaEventRecord(g_evB, g_aux);
        if (rem > 1) gemm_nt(Apanel + NB, ld, Apanel + NB, ld, diag(k + 2), ld, rem - 1, rem - 1, NB, -1.0, 1.0, true, st);
        cudaStreamWaitEvent(st, g_evB, 0);
    }
}

// The launch sequence of a factorisation is static for a given (matrix, size): it is captured into
// a CUDA graph on its second use and replayed afterwards, which removes most of the host launch
// latency from the 47-step dependency chain.  DBAT_CHOL_GRAPH=0 disables this.
void chol_factor(CholWork& w, double* A, cudaStream_t st) {
    static int use_graph = -1;
    if (use_graph < 0) { const char* e = getenv("DBAT_CHOL_GRAPH"); use_graph = (e && e[0] == '0') ? 0 : 1; }
    if (!use_graph || w.nb < 4) { chol_factor_body(w, A, st); return; }
    if (w.graphExec && w.graphA == A && w.graphStream == st) {
        if (cudaGraphLaunch((cudaGraphExec_t)w.graphExec, st) == cudaSuccess) { count_launch(w.graphLaunches); return; }
        cudaGetLastError();
    }
    if (w.seen != A) {                       // first use with this matrix: plain launch (also warms the statics)
        w.seen = A;
        chol_factor_body(w, A, st);
        return;
    }
    if (w.graphExec) { cudaGraphExecDestroy((cudaGraphExec_t)w.graphExec); w.graphExec = nullptr; }
    const int64_t l0 = g_dbat_launches;
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); chol_factor_body(w, A, st); return; }
    chol_factor_body(w, A, st);
    if (cudaStreamEndCapture(st, &graph) != cudaSuccess || !graph) { cudaGetLastError(); w.seen = nullptr; chol_factor_body(w, A, st); return; }
    cudaGraphExec_t exec = nullptr;
    if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) { cudaGetLastError(); cudaGraphDestroy(graph); w.seen = nullptr; chol_factor_body(w, A, st); return; }
    cudaGraphDestroy(graph);
    w.graphExec = exec; w.graphA = A; w.graphStream = st; w.graphLaunches = (int)(g_dbat_launches - l0);
    g_dbat_launches = l0;                    // the capture itself launched nothing
    cudaGraphLaunch(exec, st);
    count_launch(w.graphLaunches);
}

// ---------------------------------------------------------------------------------------------
// triangular solves.  The forward substitution is folded into the factorisation: the caller
// stores rhs' in the last (padding) row of A with a huge diagonal entry, so after chol_factor
// that row holds y' = (L^-1 rhs)'.  Only the backward substitution L'x = y remains.
// ---------------------------------------------------------------------------------------------
__global__ void k_put_rhs_row(double* __restrict__ A, int ld, const double* __restrict__ rhs, int n) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) A[(size_t)c * ld + ld - 1] = rhs[c];
    if (c == 0) A[(size_t)(ld - 1) * ld + ld - 1] = 1e300;
}
__global__ void k_get_y_row(const double* __restrict__ A, int ld, double* __restrict__ y, int n) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ld) y[c] = (c < n) ? A[(size_t)c * ld + ld - 1] : 0.0;
}
void chol_put_rhs(const CholWork& w, double* A, const double* rhs, cudaStream_t st) {
    k_put_rhs_row<<<(w.n + 255) / 256, 256, 0, st>>>(A, w.ld, rhs, w.n);
    count_launch();
}

// out[c] (c = 0..127) = sum_r M[r + c*ldm] * v[r], M 128x128 column-major.  256 threads.
// Each warp owns 16 columns and processes them 4 at a time (16 independent loads in flight).
__device__ __forceinline__ void tmatvec128(const double* __restrict__ M, size_t ldm, const double* v_sm,
                                           double* out_sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = v_sm[lane + 32 * i];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        double s[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double* col = M + (size_t)(warp * 16 + g * 4 + q) * ldm;
            s[q] = col[lane] * v[0] + col[lane + 32] * v[1] + col[lane + 64] * v[2] + col[lane + 96] * v[3];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int q = 0; q < 4; ++q) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) out_sm[warp * 16 + g * 4 + q] = s[q];
        }
    }
}
// x_last = invL_last' * y_last
__global__ void __launch_bounds__(256) k_bwd_first(const double* __restrict__ invL, int k,
                                                    const double* __restrict__ y, double* __restrict__ x) {
    __shared__ double v[NB], o[NB];
    if (threadIdx.x < NB) v[threadIdx.x] = y[(size_t)k * NB + threadIdx.x];
    __syncthreads();
    tmatvec128(invL + (size_t)k * NB * NB, NB, v, o);
    __syncthreads();
    if (threadIdx.x < NB) x[(size_t)k * NB + threadIdx.x] = o[threadIdx.x];
}
// step k: CTA j<k: y_j -= L_kj' x_k ; the CTA of block k-1 then forms x_{k-1} = invL_{k-1}' y_{k-1}
__global__ void __launch_bounds__(256) k_bwd_step(const double* __restrict__ A, int ld,
                                                   const double* __restrict__ invL, int k,
                                                   double* __restrict__ y, double* __restrict__ x) {
    __shared__ double v[NB], o[NB];
    const int t = threadIdx.x, j = blockIdx.x;
    if (t < NB) v[t] = x[(size_t)k * NB + t];
    __syncthreads();
    tmatvec128(A + (size_t)j * NB * ld + (size_t)k * NB, (size_t)ld, v, o);
    __syncthreads();
    double yj = 0.0;
    if (t < NB) { yj = y[(size_t)j * NB + t] - o[t]; y[(size_t)j * NB + t] = yj; }
    if (j != k - 1) return;
    __syncthreads();
    if (t < NB) v[t] = yj;
    __syncthreads();
    tmatvec128(invL + (size_t)j * NB * NB, NB, v, o);
    __syncthreads();
    if (t < NB) x[(size_t)j * NB + t] = o[t];
}

// Whole backward substitution in one cooperative launch: CTA j owns block j of the solution,
// keeps y_j in shared memory, applies y_j -= L_kj' x_k as soon as x_k is published (flag),
// then publishes x_j = invL_j' y_j.  All CTAs are co-resident (cooperative launch), so the
// spin-waits cannot deadlock.
__global__ void __launch_bounds__(256) k_bwd_coop(const double* __restrict__ A, int ld,
                                                   const double* __restrict__ invL, int nb,
                                                   const double* __restrict__ y, double* x, int* flags) {
    __shared__ double v[NB], o[NB], yj[NB];
    const int t = threadIdx.x, j = blockIdx.x;
    if (t < NB) yj[t] = y[(size_t)j * NB + t];
    __syncthreads();
    for (int k = nb - 1; k > j; --k) {
        if (t == 0) { while (atomicAdd(&flags[k], 0) == 0) { } }
        __syncthreads();
        if (t < NB) v[t] = __ldcg(x + (size_t)k * NB + t);
        __syncthreads();
        tmatvec128(A + (size_t)j * NB * ld + (size_t)k * NB, (size_t)ld, v, o);
        __syncthreads();
        if (t < NB) yj[t] -= o[t];
        __syncthreads();
    }
    tmatvec128(invL + (size_t)j * NB * NB, NB, yj, o);
    __syncthreads();
    if (t < NB) x[(size_t)j * NB + t] = o[t];
    __threadfence();
    __syncthreads();
    if (t == 0) atomicExch(&flags[j], 1);
}

// Pipelined backward substitution, one cooperative launch.  CTA j owns block j of the solution.
// Its work list is the block column below the diagonal, bottom-up, then its own inverse block:
//   L(nb-1, j), L(nb-2, j), ..., L(j+1, j), invL_j          (each 128 x 128, column-major)
// streamed through a ring of three 64-row half-block buffers with cp.async, independent of the
// flags - only the vector x_i a half-block is multiplied with has to be waited for.  The chain
// x_i -> x_{i-1} therefore costs: flag + 1 KB vector load + two 64 x 128 shared-memory mat-vecs
// for L(i, i-1)' + two for invL' + publish.
#define BWD_LD 66             // half-block column stride in shared memory (64 rows + 2)
#define BWD_BUF (NB * BWD_LD)
__device__ __forceinline__ void bwd_issue_half(double* buf, const double* __restrict__ src, size_t lds, int half, int t) {
    // 128 columns x 64 rows, 16-byte chunks: 32 chunks per column
    for (int idx = t; idx < NB * 32; idx += 256) {
        const int c = idx >> 5, r2 = (idx & 31) * 2;
        cp_async16(buf + c * BWD_LD + r2, src + (size_t)c * lds + half * 64 + r2);
    }
}
__global__ void __launch_bounds__(256, 1) k_bwd_pipe(const double* __restrict__ A, int ld,
                                                     const double* __restrict__ invL, int nb,
                                                     const double* __restrict__ y, double* x, int* flags) {
    extern __shared__ __align__(16) double sm[];
    double* ring = sm;                       // [3][BWD_BUF]
    double* v = sm + 3 * BWD_BUF;            // [128] vector being multiplied
    double* yj = v + NB;                     // [128] running right-hand side of block j
    double* part = yj + NB;                  // [2][128] partial sums of the two row halves of a thread pair
    const int t = threadIdx.x, j = blockIdx.x;
    const int nItems = 2 * (nb - 1 - j) + 2; // half-blocks in the work list
    auto item_src = [&](int it, const double*& src, size_t& lds, int& half) {
        const int blk = it >> 1; half = it & 1;
        if (blk < nb - 1 - j) { const int i = nb - 1 - blk; src = A + (size_t)j * NB * ld + (size_t)i * NB; lds = (size_t)ld; }
        else { src = invL + (size_t)j * NB * NB; lds = NB; }
    };
    if (t < NB) yj[t] = y[(size_t)j * NB + t];
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        if (it < nItems) { const double* src; size_t lds; int half; item_src(it, src, lds, half); bwd_issue_half(ring + it * BWD_BUF, src, lds, half, t); }
        cp_async_commit();
    }
    const int c = t & 127, hr = t >> 7;      // thread pair (c, hr): column c, rows 32 hr .. 32 hr + 31 of the half
    for (int it = 0; it < nItems; ++it) {
        // keep two half-blocks in flight beyond the current one
        if (it + 2 < nItems) { const double* src; size_t lds; int half; item_src(it + 2, src, lds, half); bwd_issue_half(ring + ((it + 2) % 3) * BWD_BUF, src, lds, half, t); }
        cp_async_commit();
        const int blk = it >> 1, half = it & 1;
        const bool isInv = blk >= nb - 1 - j;
        if (half == 0) {                     // new vector
            if (!isInv) {
                const int i = nb - 1 - blk;
                if (t == 0) { while (atomicAdd(&flags[i], 0) == 0) { } }
                __syncthreads();
                if (t < NB) v[t] = __ldcg(x + (size_t)i * NB + t);
            } else {
                __syncthreads();
                if (t < NB) v[t] = yj[t];
            }
        }
        cp_async_wait<2>();
        __syncthreads();
        const double* M = ring + (it % 3) * BWD_BUF + c * BWD_LD + 32 * hr;
        const double* vv = v + 64 * half + 32 * hr;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int r = 0; r < 32; r += 2) {
            const double2 m2 = *reinterpret_cast<const double2*>(M + r);
            const double2 v2 = *reinterpret_cast<const double2*>(vv + r);
            s0 = fma(m2.x, v2.x, s0); s1 = fma(m2.y, v2.y, s1);
        }
        part[hr * NB + c] = s0 + s1;
        __syncthreads();
        if (t < NB) {
            const double s = part[t] + part[NB + t];
            if (!isInv) yj[t] -= s;
            else if (half == 0) yj[t] = s;   // x_j accumulates in place of yj (v holds the old yj)
            else yj[t] += s;
        }
        // the buffer (it % 3) is refilled by the issue at the top of iteration it + 1: the
        // barrier after its vector load / cp_async_wait orders that after these reads
    }
    __syncthreads();
    if (t < NB) x[(size_t)j * NB + t] = yj[t];
    __threadfence();
    __syncthreads();
    if (t == 0) atomicExch(&flags[j], 1);
}

static thread_local double* g_solve_tmp = nullptr;
static thread_local int* g_solve_flags = nullptr;
static thread_local int g_solve_tmp_n = 0;
static thread_local int g_coop_max = -1;
// After chol_factor on a matrix prepared with chol_put_rhs: x (length ld) = A^-1 rhs.
void chol_solve(const CholWork& w, const double* A, double* x, cudaStream_t st) {
    if (g_solve_tmp_n < w.ld) {
        if (g_solve_tmp) { cudaFree(g_solve_tmp); cudaFree(g_solve_flags); }
        cudaMalloc(&g_solve_tmp, sizeof(double) * w.ld);
        cudaMalloc(&g_solve_flags, sizeof(int) * (w.ld / NB + 1));
        g_solve_tmp_n = w.ld;
    }
    if (g_coop_max < 0) {
        int dev = 0, coop = 0, sms = 0, occ = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_bwd_coop, 256, 0);
        g_coop_max = coop ? sms * occ : 0;
    }
    double* y = g_solve_tmp;
    k_get_y_row<<<(w.ld + 255) / 256, 256, 0, st>>>(A, w.ld, y, w.n);
    count_launch();
    static thread_local int pipe_ok = -1;
    const int pipe_smem = (3 * BWD_BUF + 4 * NB) * 8;
    if (pipe_ok < 0) {
        const char* e = getenv("DBAT_BWD");
        int dev = 0, coop = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        pipe_ok = (coop && !(e && e[0] == '0') &&
                   cudaFuncSetAttribute(k_bwd_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, pipe_smem) == cudaSuccess) ? sms : 0;
    }
    if (w.nb <= pipe_ok) {
        cudaMemsetAsync(g_solve_flags, 0, sizeof(int) * w.nb, st);
        const double* Ac = A; const double* iL = w.invL; int ld = w.ld, nb = w.nb; const double* yc = y;
        int* fl = g_solve_flags;
        void* args[] = {(void*)&Ac, (void*)&ld, (void*)&iL, (void*)&nb, (void*)&yc, (void*)&x, (void*)&fl};
        cudaLaunchCooperativeKernel((void*)k_bwd_pipe, dim3(w.nb), dim3(256), args, pipe_smem, st);
        count_launch();
        return;
    }
    if (w.nb <= g_coop_max) {
        cudaMemsetAsync(g_solve_flags, 0, sizeof(int) * w.nb, st);
        const double* Ac = A; const double* iL = w.invL; int ld = w.ld, nb = w.nb; const double* yc = y;
        int* fl = g_solve_flags;
        void* args[] = {(void*)&Ac, (void*)&ld, (void*)&iL, (void*)&nb, (void*)&yc, (void*)&x, (void*)&fl};
        cudaLaunchCooperativeKernel((void*)k_bwd_coop, dim3(w.nb), dim3(256), args, 0, st);
        count_launch();
        return;
    }
    k_bwd_first<<<1, 256, 0, st>>>(w.invL, w.nb - 1, y, x);
    count_launch();
    for (int k = w.nb - 1; k >= 1; --k) { k_bwd_step<<<k, 256, 0, st>>>(A, w.ld, w.invL, k, y, x); count_launch(); }
}

// ---------------------------------------------------------------------------------------------
// explicit inverse: Z = inv(L) by block forward substitution (GEMMs), then C = Z'Z = inv(A)
// ---------------------------------------------------------------------------------------------
__global__ void k_copy_block(const double* __restrict__ src, int lds, double* __restrict__ dst, int ldd) {
    const int c = blockIdx.x, r = threadIdx.x;
    dst[(size_t)c * ldd + r] = src[(size_t)c * lds + r];
}
__global__ void k_mirror_lower(double* __restrict__ C, int ld) {
    const int c = blockIdx.y * 32 + threadIdx.y, r = blockIdx.x * 32 + threadIdx.x;
    if (r > c) C[(size_t)r * ld + c] = C[(size_t)c * ld + r];   // upper(c,r) := lower(r,c)
}
// Zt(by.., bx..) = Z(bx.., by..)'  on 32x32 tiles of a 128x128 block
__global__ void k_transpose(const double* __restrict__ Z, double* __restrict__ Zt, int ld) {
    __shared__ double tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    tile[threadIdx.y][threadIdx.x] = Z[(size_t)(by + threadIdx.y) * ld + bx + threadIdx.x];
    __syncthreads();
    Zt[(size_t)(bx + threadIdx.y) * ld + by + threadIdx.x] = tile[threadIdx.x][threadIdx.y];
}
// out(:,c) = -M * T(:,c) for a 128x128 lower-triangular M (column-major); also writes the
// transposed result outT(c, r) = out(r, c).  One CTA per 32 columns; in place (out == T) is fine.
__global__ void __launch_bounds__(128) k_left_mul_neg(const double* __restrict__ M, const double* T,
                                                       int ldt, double* out, int ldo,
                                                       double* __restrict__ outT, int ldoT) {
    __shared__ double Ts[NB][33];
    const int c0 = blockIdx.x * 32, t = threadIdx.x;
    for (int c = 0; c < 32; ++c) Ts[t][c] = T[(size_t)(c0 + c) * ldt + t];
    __syncthreads();
    double acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.0;
    for (int k = 0; k <= t; ++k) {
        const double m = M[(size_t)k * NB + t];
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] += m * Ts[k][c];
    }
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        out[(size_t)(c0 + c) * ldo + t] = -acc[c];
        outT[(size_t)t * ldoT + c0 + c] = -acc[c];
    }
}

// Z and C are ld x ld scratch/output buffers.  On return C holds the full symmetric inv(A);
// Z holds inv(L)' (upper triangular).
void chol_inverse(const CholWork& w, const double* A, double* Z, double* C, cudaStream_t st) {
    const int ld = w.ld, nb = w.nb;
    double* Zt = C;                                       // C doubles as storage for Z' until the end
    cudaMemsetAsync(Z, 0, sizeof(double) * (size_t)ld * ld, st);
    cudaMemsetAsync(Zt, 0, sizeof(double) * (size_t)ld * ld, st);
    const dim3 tb(32, 32), tg(NB / 32, NB / 32);
    for (int i = 0; i < nb; ++i) {
        double* Zii = Z + (size_t)i * NB * ld + (size_t)i * NB;
        k_copy_block<<<NB, NB, 0, st>>>(w.invL + (size_t)i * NB * NB, NB, Zii, ld);
        k_transpose<<<tg, tb, 0, st>>>(Zii, Zt + (size_t)i * NB * ld + (size_t)i * NB, ld);
        count_launch(2);
        if (i == 0) continue;
        // T = L(i,0:i) * Z(0:i,0:i)  (NT with B = Z'),  written into Z(i,0:i)
        double* T = Z + (size_t)i * NB;
        gemm_nt(A + (size_t)i * NB, ld, Zt, ld, T, ld, 1, i, NB * i, 1.0, 0.0, false, st);
        // Z(i,0:i) = -inv(L_ii) * T   (+ transposed copy into Zt(0:i, i))
        k_left_mul_neg<<<(NB * i) / 32, 128, 0, st>>>(w.invL + (size_t)i * NB * NB, T, ld, T, ld,
                                                       Zt + (size_t)i * NB * ld, ld);
        count_launch();
    }
    // C = Z'Z = Zt * Zt'.  C aliases Zt, so move Zt into Z first.
    cudaMemcpyAsync(Z, Zt, sizeof(double) * (size_t)ld * ld, cudaMemcpyDeviceToDevice, st);
    gemm_nt(Z, ld, Z, ld, C, ld, nb, nb, ld, 1.0, 0.0, true, st);
    k_mirror_lower<<<dim3(ld / 32, ld / 32), dim3(32, 32), 0, st>>>(C, ld);
    count_launch();
}
