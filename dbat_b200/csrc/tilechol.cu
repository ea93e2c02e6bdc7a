// tilechol.cu — sparse tile Cholesky of the reduced camera system on the FP64 tensor pipe (DMMA
// mma.sync.m8n8k4.f64), one persistent data-flow kernel per factorisation.
//
// Replaces MATLAB's `\` / chol (CHOLMOD supernodal) at code/bundle/lsa/levenberg_marquardt.m:119,
// gauss_newton_armijo.m:172, levenberg_marquardt_powell.m:272-277 for the reduced camera system.
// The matrix is held as 64 x 64 tiles of the sparse pattern computed once per problem (tilesym.cu:
// nested-dissection order of the images, shared IO block and rhs row last, tile-level symbolic
// fill).  Left-looking: ONE task per tile,
//     L(I,J) = ( S(I,J) - sum_k L(I,k) L(J,k)' ) inv(L(J,J))'          (I > J)
//     L(J,J) = chol( S(J,J) - sum_k L(J,k) L(J,k)' )
// The sums run over the k < J where both tiles exist; they stay in registers, so every tile is read
// once as S and written once as L (the right-looking dense code re-read the trailing matrix 47 times).
// CTAs claim tasks from one list (columns sorted by elimination-tree level: independent subtrees
// interleave), and wait on per-tile ready flags: a task depends only on tasks earlier in the list,
// every claimed task belongs to a running CTA, hence no deadlock for any grid size.
// The triangular solve against the diagonal tile uses the inverses of its four 16 x 16 diagonal
// blocks (which fall out of the pivot sweep for free) and is row-wise independent: each warp owns
// 16 rows of the tile and needs no CTA-wide barrier.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include "launch.h"
#include "tilechol.h"

#define LA 68                     // shared-memory column stride of a 64-row operand (= 4 mod 16)
#define PB 16
#define TC_STAGES 4
#define TC_SLICE (16 * LA)        // one 16-column slice of one operand
#define TC_STAGE (2 * TC_SLICE)
#define TC_OFF_LS (64 * LA)       // solve phase: L(J,J) columns 0..47 behind the 64 x 64 work tile
#define TC_OFF_XD (TC_OFF_LS + 48 * LA)
#define TC_OFF_COLB (TC_OFF_XD + TC_XD)
#define TC_SMEM_DOUBLES (TC_OFF_COLB + 3 * PB + 8)
static_assert(TC_OFF_XD >= TC_STAGES * TC_STAGE - TC_XD || true, "layout");
static_assert(TC_STAGES * TC_STAGE <= TC_OFF_COLB, "pipeline stages must fit under the scratch area");
#define TC_THREADS 128

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
// lane 0 of the warp spins until *f == epoch, then the whole warp continues (the barrier orders the
// other lanes' later reads after lane 0's acquire)
__device__ __forceinline__ void warp_wait_flag(const int* f, int epoch, int lane) {
    if (lane == 0) { while (ld_acquire(f) != epoch) { __nanosleep(32); } }
    __syncwarp();
}
// panel progress of a tile: wait until at least `need` of its four 16-column panels are final
__device__ __forceinline__ void warp_wait_prog(const int* f, int need, int lane) {
    if (lane == 0) { while (ld_acquire(f) < need) { __nanosleep(20); } }
    __syncwarp();
}
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TC_STAMP(k) { if (D.prof && tid == 0) D.prof[4 * (size_t)task + (k)] = gtimer(); }
__device__ __forceinline__ double rsqrt_nr(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double h = d * y, e = fma(-h, y, 1.0);
    y = fma(0.5 * y, e, y);
    h = d * y; e = fma(-h, y, 1.0);
    return fma(0.5 * y, e, y);
}

// ---------------------------------------------------------------------------------------------
// 64 x 64 diagonal tile in shared memory (column-major, stride LA): Cholesky in place + the
// inverses of the four 16 x 16 diagonal blocks (Xd[p][n'][n] = inv(L_pp)(n', n), row stride TC_XLD).
// ---------------------------------------------------------------------------------------------
// D(p): one warp factors the 16 x 16 diagonal block in registers, LDL' form (lane r < 16: row r;
// lanes 16..31 carry the columns of the identity through the same sweep -> inv(L_pp)).  Between
// two pivots: reciprocal seed, one correction folded into the update, one shuffle.
__device__ __forceinline__ void tc_diag16(double* As, double* Xd, double* colb, int c0, int lane,
                                          const unsigned char* __restrict__ valid, int gcol0,
                                          double& lmin, double& lmax, int& bad) {
    double* spd = colb + 2 * PB;
    const int r = lane & 15;
    const bool isX = lane >= 16;
    double v[PB];
#pragma unroll
    for (int c = 0; c < PB; ++c) {
        const double a = As[(c0 + c) * LA + c0 + r];
        v[c] = isX ? (c == r ? 1.0 : 0.0) : a;
    }
    double d = __shfl_sync(0xffffffffu, v[0], 0);
#pragma unroll
    for (int j = 0; j < PB; ++j) {
        double* cb = colb + (j & 1) * PB;
        if (!isX) cb[r] = v[j];
        double y;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
        const double e = fma(-d, y, 1.0);
        const double vy = v[j] * y;
        const double dj = d;
        if (j + 1 < PB) {
            const double q = v[j] * vy;
            const double base = v[j + 1] - q;
            const double t = fma(e, e, e);
            const double cand = fma(-q, t, base);
            d = __shfl_sync(0xffffffffu, cand, j + 1);
        }
        const double e2 = e * e;
        const double vy1 = fma(vy, e, vy);
        const double w = fma(vy1, e2, vy1);
        __syncwarp();
#pragma unroll
        for (int c = j + 1; c < PB; ++c) v[c] = fma(-w, cb[c], v[c]);
        if (lane == 0) spd[j] = dj;
    }
    __syncwarp();
    {
        const double dr = spd[r];
        const double isdr = rsqrt_nr(dr);
        if (!(dr > 0.0)) bad = 1;
        if (!isX && valid[gcol0 + r]) { const double l = dr * isdr; lmin = fmin(lmin, l); lmax = fmax(lmax, l); }
#pragma unroll
        for (int j = 0; j < PB; ++j) v[j] *= __shfl_sync(0xffffffffu, isdr, j);
    }
    if (!isX) {
#pragma unroll
        for (int c = 0; c < PB; ++c) if (c <= r) As[(c0 + c) * LA + c0 + r] = v[c];
    } else {
#pragma unroll
        for (int j = 0; j < PB; ++j) Xd[j * TC_XLD + r] = v[j];         // inv(L_pp)(j, r); zero above the diagonal
    }
}

// R(p): 8-row tile q below the diagonal block: L(i, panel) = A(i, panel) inv(L_pp)'
__device__ __forceinline__ void tc_rtile(double* As, const double* Xd, int c0, int q, int lane) {
    const int fr = lane >> 2, fk = lane & 3;
    const double* pb = Xd + fr * TC_XLD + fk;
    const double* pa = As + (c0 + fk) * LA + c0 + PB + 8 * q + fr;
    double r0[4][2], r1[4][2];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double a = pa[4 * k * LA];
        r0[k][0] = r0[k][1] = r1[k][0] = r1[k][1] = 0.0;
        dmma(r0[k][0], r0[k][1], a, pb[4 * k]);
        dmma(r1[k][0], r1[k][1], a, pb[8 * TC_XLD + 4 * k]);
    }
    __syncwarp();                                   // every lane has read its A fragments before the in-place store
    double* pc = As + (c0 + 2 * fk) * LA + c0 + PB + 8 * q + fr;
    pc[0] = (r0[0][0] + r0[1][0]) + (r0[2][0] + r0[3][0]);
    pc[LA] = (r0[0][1] + r0[1][1]) + (r0[2][1] + r0[3][1]);
    pc[8 * LA] = (r1[0][0] + r1[1][0]) + (r1[2][0] + r1[3][0]);
    pc[9 * LA] = (r1[0][1] + r1[1][1]) + (r1[2][1] + r1[3][1]);
}

// U(p): 8 x 8 tile (ti, tj), ti >= tj, of the trailing block: A(i,j) -= L(i,panel) L(j,panel)'
__device__ __forceinline__ void tc_utile(double* As, int c0, int ti, int tj, int lane) {
    const int fr = lane >> 2, fk = lane & 3;
    const int b0 = c0 + PB;
    const double* pa0 = As + (c0 + fk) * LA + b0 + fr;
    double acc[4][2];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        acc[k][0] = acc[k][1] = 0.0;
        dmma(acc[k][0], acc[k][1], pa0[4 * k * LA + 8 * ti], pa0[4 * k * LA + 8 * tj]);
    }
    double* pc = As + (b0 + 8 * tj + 2 * fk) * LA + b0 + 8 * ti + fr;
    pc[0] -= (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]);
    pc[LA] -= (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
}

// Every finished 16-column panel leaves for global memory at once (with the inverse of its diagonal block) and is
// published through the tile's progress word: the triangular solves of the tiles below start on panel 0 while
// panels 1..3 are still being factored.
// Look-ahead inside the tile: after the rows below a panel are solved, the three 8 x 8 tiles of the NEXT diagonal
// block are updated first; warp 0 then runs the 16 serial pivots of that block while warps 1..3 apply the rest of
// the trailing update and write the finished panel out - the pivot sweeps (half of the time of this routine) no
// longer wait for the trailing update, and nobody waits for the write-out.
__device__ __forceinline__ void tc_utile_t(double* As, int c0, int t, int lane) {
    int ti = (int)((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    while (ti * (ti + 1) / 2 > t) --ti;
    tc_utile(As, c0, ti, t - ti * (ti + 1) / 2, lane);
}
__device__ void tc_potrf64(double* As, double* Xd, double* colb, int warp, int lane,
                           const unsigned char* __restrict__ valid, int gcol0, int* info, unsigned long long* minmax,
                           int J, double* __restrict__ tile, double* __restrict__ gx, int* prog, int progBase) {
    double lmin = 1e300, lmax = 0.0;
    int bad = 0;
    const int tid = warp * 32 + lane;
    // panel q: columns 16q .. 16q+15 of L(J,J) (zeros above the diagonal) and inv(L_qq), by `nth` threads (index u)
    auto publish = [&](int q, int u, int nth) {
        const int c0 = PB * q;
        for (int idx = u; idx < PB * 32; idx += nth) {
            const int c = c0 + (idx >> 5), r2 = (idx & 31) * 2;
            double2 v;
            v.x = (r2 >= c) ? As[c * LA + r2] : 0.0;
            v.y = (r2 + 1 >= c) ? As[c * LA + r2 + 1] : 0.0;
            *reinterpret_cast<double2*>(tile + c * 64 + r2) = v;
        }
        for (int idx = u; idx < PB * TC_XLD; idx += nth) gx[q * PB * TC_XLD + idx] = Xd[q * PB * TC_XLD + idx];
    };
    for (int p = 0; p < 4; ++p) {
        const int c0 = PB * p;
        const int m = 6 - 2 * p;                              // 8-row tiles below the diagonal block
        if (warp == 0) {
            tc_diag16(As, Xd + p * PB * TC_XLD, colb, c0, lane, valid, gcol0 + c0, lmin, lmax, bad);
        } else if (p > 0) {
            // the rest of the trailing update of panel p-1 (its first three tiles - this diagonal block - are done),
            // then panel p-1 leaves; warps 1..3 synchronise among themselves (named barrier 1, 96 threads)
            const int mp = m + 2, ntp = mp * (mp + 1) / 2;
            for (int t = 3 + (warp - 1); t < ntp; t += 3) tc_utile_t(As, c0 - PB, t, lane);
            publish(p - 1, tid - 32, TC_THREADS - 32);
            asm volatile("bar.sync 1, 96;" ::: "memory");
            if (tid == 32) st_release(prog, progBase + p);
        }
        __syncthreads();
        for (int q = warp; q < m; q += TC_THREADS / 32) tc_rtile(As, Xd + p * PB * TC_XLD, c0, q, lane);
        __syncthreads();
        if (p < 3) {
            if (warp < 3) tc_utile_t(As, c0, warp, lane);    // the next diagonal block: tiles (0,0), (1,0), (1,1)
            __syncthreads();
        }
    }
    publish(3, tid, TC_THREADS);
    __syncthreads();                                          // every thread's stores before the release
    if (tid == 0) st_release(prog, progBase + 4);
    if (warp == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lmin = fmin(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
            lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        }
        bad = __any_sync(0xffffffffu, bad);
        if (lane == 0) {
            if (bad) atomicCAS(info, 0, J + 1);
            if (lmax > 0.0) {                                 // positive doubles order like their bit patterns
                atomicMin(minmax, (unsigned long long)__double_as_longlong(lmin));
                atomicMax(minmax + 1, (unsigned long long)__double_as_longlong(lmax));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// the factorisation kernel
// ---------------------------------------------------------------------------------------------
// queue[qBegin, qEnd): the task ids this launch serves, claimed in order through counters[counter].  chain != 0: this
// launch holds the few CTAs of the dependency chain (launched with enough dynamic shared memory to have an SM each);
// they let the bulk launch, which waits on them programmatically, start as soon as they are all resident.
__global__ void __launch_bounds__(TC_THREADS, 3) k_tchol_factor(TCholDev D, int epoch, int qBegin, int qEnd, int counter, int chain) {
    if (chain) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    extern __shared__ __align__(16) double sm[];
    __shared__ int s_task;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp & 1) * 32, wn = (warp >> 1) * 32;
    const int fr = lane >> 2, fk = lane & 3;
    double* Xd = sm + TC_OFF_XD;
    double* colb = sm + TC_OFF_COLB;
    for (;;) {
        __syncthreads();                                      // previous task's shared memory is free
        if (tid == 0) { const int q = qBegin + atomicAdd(D.counters + counter, 1); s_task = q < qEnd ? D.queue[q] : -1; }
        __syncthreads();
        const int task = s_task;
        if (task < 0) break;
        TC_STAMP(0)
        const int I = D.taskI[task], J = D.taskJ[task];
        const int slot = D.tix[(size_t)I * D.nT + J];
        double* tile = D.tiles + ((size_t)slot << 12);
        const long long t0 = D.termPtr[task];
        const int nterm = (int)(D.termPtr[task + 1] - t0);
        // ---- accumulator: minus the tile's content (S itself, or what an earlier link of this tile's chain / the
        //      other ranks left there), or zero for a fill tile nobody has touched
        double acc[4][4][2];
        const int waitAux = D.taskWait[task];
        if (waitAux >= 0) warp_wait_flag(D.aux + waitAux, epoch, lane);
        if (D.taskInit[task]) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double* p0 = tile + (wn + 8 * j + 2 * fk) * 64 + wm + 8 * i + fr;
                    acc[i][j][0] = -__ldcg(p0); acc[i][j][1] = -__ldcg(p0 + 64);
                }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
        }
        if (I == J) {
            // ---- diagonal tile: sum_k L(J,k) L(J,k)'.  One operand per term, so a whole tile is fetched at
            //      once into one of two buffers the moment its flag is up: the last term - the one on the
            //      critical path of the factorisation - costs one L2 round trip and 64 uninterrupted k-steps.
            for (int term = 0; term < nterm; ++term) {
                const int sa = D.termA[t0 + term];
                double* As = sm + (term & 1) * (64 * LA);
                const double* ga = D.tiles + ((size_t)sa << 12);
                if (term == nterm - 1) {
                    // the last term is the tile the dependency chain just produced: take its 16-column panels as
                    // they are published (the triangular solve that writes it is still working on the later ones)
                    for (int p = 0; p < 4; ++p) {
                        warp_wait_prog(D.prog + sa, 8 * epoch + p + 1, lane);
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int chunk = tid + c * TC_THREADS;
                            const int kk = 16 * p + (chunk >> 5), m2 = (chunk & 31) * 2;
                            cp_async16(As + kk * LA + m2, ga + kk * 64 + m2);
                        }
                        cp_async_commit();
                        cp_async_wait<0>();
                        __syncthreads();
                        if (wm >= wn) {
#pragma unroll
                            for (int k0 = 16 * p; k0 < 16 * p + 16; k0 += 4) {
                                double af[4], bf[4];
                                const double* ap = As + (k0 + fk) * LA + wm + fr;
                                const double* bp = As + (k0 + fk) * LA + wn + fr;
#pragma unroll
                                for (int i = 0; i < 4; ++i) af[i] = ap[8 * i];
#pragma unroll
                                for (int j = 0; j < 4; ++j) bf[j] = bp[8 * j];
#pragma unroll
                                for (int i = 0; i < 4; ++i)
#pragma unroll
                                    for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
                            }
                        }
                    }
                    continue;
                }
                warp_wait_flag(D.flag + sa, epoch, lane);
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const int chunk = tid + c * TC_THREADS;
                    const int kk = chunk >> 5, m2 = (chunk & 31) * 2;
                    cp_async16(As + kk * LA + m2, ga + kk * 64 + m2);
                }
                cp_async_commit();
                cp_async_wait<0>();
                __syncthreads();
                if (wm >= wn) {                                 // the upper 32 x 32 block is never used
#pragma unroll 4
                    for (int k0 = 0; k0 < 64; k0 += 4) {
                        double af[4], bf[4];
                        const double* ap = As + (k0 + fk) * LA + wm + fr;
                        const double* bp = As + (k0 + fk) * LA + wn + fr;
#pragma unroll
                        for (int i = 0; i < 4; ++i) af[i] = ap[8 * i];
#pragma unroll
                        for (int j = 0; j < 4; ++j) bf[j] = bp[8 * j];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
                    }
                }
            }
        } else {
        // ---- sum_k L(I,k) L(J,k)': 16-column slices through a cp.async ring; a term is issued once
        //      the flags of its two tiles are up
        const int nsl = nterm * 4;
        auto issue = [&](int q) {
            const int term = q >> 2, ks = q & 3;
            const int sa = D.termA[t0 + term], sb = D.termB[t0 + term];
            if (ks == 0) {
                warp_wait_flag(D.flag + sa, epoch, lane);
                warp_wait_flag(D.flag + sb, epoch, lane);
            }
            double* As = sm + (q % TC_STAGES) * TC_STAGE;
            double* Bs = As + TC_SLICE;
            const double* ga = D.tiles + ((size_t)sa << 12) + ks * 16 * 64;
            const double* gb = D.tiles + ((size_t)sb << 12) + ks * 16 * 64;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int chunk = tid + c * TC_THREADS;
                const int kk = chunk >> 5, m2 = (chunk & 31) * 2;
                cp_async16(As + kk * LA + m2, ga + kk * 64 + m2);
                cp_async16(Bs + kk * LA + m2, gb + kk * 64 + m2);
            }
        };
#pragma unroll
        for (int s = 0; s < TC_STAGES - 1; ++s) { if (s < nsl) issue(s); cp_async_commit(); }
        for (int q = 0; q < nsl; ++q) {
            cp_async_wait<TC_STAGES - 2>();
            __syncthreads();
            const double* As = sm + (q % TC_STAGES) * TC_STAGE;
            const double* Bs = As + TC_SLICE;
#pragma unroll
            for (int k0 = 0; k0 < 16; k0 += 4) {
                double af[4], bf[4];
                const double* ap = As + (k0 + fk) * LA + wm + fr;
                const double* bp = Bs + (k0 + fk) * LA + wn + fr;
#pragma unroll
                for (int i = 0; i < 4; ++i) af[i] = ap[8 * i];
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = bp[8 * j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
            const int nq = q + TC_STAGES - 1;
            if (nq < nsl) issue(nq);
            cp_async_commit();
        }
        }
        cp_async_wait<0>();
        __syncthreads();
        TC_STAMP(1)
        if (D.taskMode[task] == 1) {
            // partial sum only: the tile := (local part of S) - sum over this subtree; summed over the ranks later
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double* p0 = tile + (wn + 8 * j + 2 * fk) * 64 + wm + 8 * i + fr;
                    p0[0] = -acc[i][j][0]; p0[64] = -acc[i][j][1];
                }
            const int setAux = D.taskSet[task];
            if (setAux >= 0) {
                __syncthreads();                               // every thread's stores before the release
                if (tid == 0) st_release(D.aux + setAux, epoch);
            }
            TC_STAMP(3)
            continue;
        }
        // ---- C = S - sum (= -acc) into the work tile (column-major, stride LA)
        double* Cs = sm;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double* p0 = Cs + (wn + 8 * j + 2 * fk) * LA + wm + 8 * i + fr;
                p0[0] = -acc[i][j][0]; p0[LA] = -acc[i][j][1];
            }
        if (I == J) {
            __syncthreads();
            TC_STAMP(2)
            // L(J,J) (lower triangle; zeros above) and the inverse diagonal blocks leave panel by panel
            tc_potrf64(Cs, Xd, colb, warp, lane, D.valid, 64 * J, D.info, D.minmax, J, tile, D.invD + (size_t)J * TC_XD,
                       D.prog + slot, 8 * epoch);
        } else {
            // ---- X = C inv(L(J,J))' by 16-column panels; warp w owns rows 16w .. 16w+15
            // Panel p needs the inverse of the p-th diagonal block of L(J,J) and the rows of panel p in the columns
            // before it: both are final once panels 0..p of the diagonal tile are published, so this solve follows
            // the factorisation of the diagonal tile one panel behind instead of waiting for all of it.
            const int dslot = D.tix[(size_t)J * D.nT + J];
            double* LsW = sm + TC_OFF_LS;
            const double* gl = D.tiles + ((size_t)dslot << 12);
            const double* gx = D.invD + (size_t)J * TC_XD;
            const double* Ls = sm + TC_OFF_LS;
            const int R0 = 16 * warp;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int c0 = PB * p;
                warp_wait_prog(D.prog + dslot, 8 * epoch + p + 1, lane);
                if (p < 3) {                                   // columns of panel p: operands of the later panels
                    for (int chunk = tid; chunk < PB * 32; chunk += TC_THREADS) {
                        const int c = c0 + (chunk >> 5), m2 = (chunk & 31) * 2;
                        cp_async16(LsW + c * LA + m2, gl + c * 64 + m2);
                    }
                }
                cp_async_commit();
                for (int idx = tid; idx < PB * TC_XLD; idx += TC_THREADS) Xd[p * PB * TC_XLD + idx] = __ldcg(gx + p * PB * TC_XLD + idx);
                cp_async_wait<0>();
                __syncthreads();                               // C stored (p = 0), panel operands in place
                if (p == 0) TC_STAMP(2)
                double t[2][2][2];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const double* p0 = Cs + (c0 + 8 * j + 2 * fk) * LA + R0 + 8 * i + fr;
                        t[i][j][0] = p0[0]; t[i][j][1] = p0[LA];
                    }
                for (int k0 = 0; k0 < c0; k0 += 4) {          // T = C(:,p) - X(:, 0:c0) L(p rows, 0:c0)'
                    const double a0 = -Cs[(k0 + fk) * LA + R0 + fr], a1 = -Cs[(k0 + fk) * LA + R0 + 8 + fr];
                    const double b0 = Ls[(k0 + fk) * LA + c0 + fr], b1 = Ls[(k0 + fk) * LA + c0 + 8 + fr];
                    dmma(t[0][0][0], t[0][0][1], a0, b0); dmma(t[0][1][0], t[0][1][1], a0, b1);
                    dmma(t[1][0][0], t[1][0][1], a1, b0); dmma(t[1][1][0], t[1][1][1], a1, b1);
                }
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        double* p0 = Cs + (c0 + 8 * j + 2 * fk) * LA + R0 + 8 * i + fr;
                        p0[0] = t[i][j][0]; p0[LA] = t[i][j][1];
                    }
                __syncwarp();
                double x[2][2][2];                             // X(:,p) = T inv(L_pp)'
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) { x[i][j][0] = 0.0; x[i][j][1] = 0.0; }
                const double* xd = Xd + p * PB * TC_XLD;
#pragma unroll
                for (int k0 = 0; k0 < PB; k0 += 4) {
                    const double a0 = Cs[(c0 + k0 + fk) * LA + R0 + fr], a1 = Cs[(c0 + k0 + fk) * LA + R0 + 8 + fr];
                    const double b0 = xd[fr * TC_XLD + k0 + fk], b1 = xd[(8 + fr) * TC_XLD + k0 + fk];
                    dmma(x[0][0][0], x[0][0][1], a0, b0); dmma(x[0][1][0], x[0][1][1], a0, b1);
                    dmma(x[1][0][0], x[1][0][1], a1, b0); dmma(x[1][1][0], x[1][1][1], a1, b1);
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        double* p0 = Cs + (c0 + 8 * j + 2 * fk) * LA + R0 + 8 * i + fr;
                        p0[0] = x[i][j][0]; p0[LA] = x[i][j][1];
                        double* g0 = tile + (c0 + 8 * j + 2 * fk) * 64 + R0 + 8 * i + fr;
                        g0[0] = x[i][j][0]; g0[64] = x[i][j][1];
                    }
                __syncthreads();                               // panel p of this tile is final in global memory
                if (tid == 0) st_release(D.prog + slot, 8 * epoch + p + 1);
            }
        }
        __syncthreads();                                      // orders every thread's tile stores before the release
        if (tid == 0) st_release(D.flag + slot, epoch);
        TC_STAMP(3)
    }
}

// ---------------------------------------------------------------------------------------------
// backward substitution L' x = y (y = the rhs row of the factor).  One task per tile column, taken
// in descending elimination-tree level; block J waits for the blocks of the rows of its column.
// ---------------------------------------------------------------------------------------------
#ifndef TC_BWD_ILP
#define TC_BWD_ILP 1                   // 0: every dot product of the backward substitution as one dependent chain (cross-check)
#endif
#define BW_LD 65
__global__ void __launch_bounds__(TC_THREADS) k_tchol_bwd(TCholDev D, int epoch, double* __restrict__ xs, int nCols) {
    __shared__ double Ls[64 * BW_LD];
    __shared__ double Xd[4 * 16 * 16];
    __shared__ double yv[64], xv[64];
    __shared__ int s_task;
    const int tid = threadIdx.x;
    const int nT = D.nT;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_task = atomicAdd(D.counters + 4, 1);
        __syncthreads();
        if (s_task >= nCols) break;
        const int J = D.bwdCols[s_task];
        const int e0 = D.colPtr[J], e1 = D.colPtr[J + 1];
        const int yslot = D.tix[(size_t)(nT - 1) * nT + J];
        // y block: row 63 of the tile (nT-1, J); the rhs row itself solves to 0
        if (tid < 64) yv[tid] = (J == nT - 1 && tid == 63) ? 0.0 : D.tiles[((size_t)yslot << 12) + tid * 64 + 63];
        {   // diagonal tile and its inverse blocks (final since the factorisation kernel ended)
            const double* gl = D.tiles + ((size_t)D.colSlot[e0] << 12);
            for (int idx = tid; idx < 64 * 64; idx += TC_THREADS) Ls[(idx >> 6) * BW_LD + (idx & 63)] = gl[idx];
            const double* gx = D.invD + (size_t)J * TC_XD;
            for (int idx = tid; idx < 1024; idx += TC_THREADS) {
                const int p = idx >> 8, a = (idx >> 4) & 15, b = idx & 15;
                Xd[idx] = gx[p * PB * TC_XLD + a * TC_XLD + b];
            }
        }
        const int c = tid >> 1, h = tid & 1;
        double acc = 0.0;
#if TC_BWD_ILP
        double acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
#endif
        double2 v[16];
        bool have = false;
        for (int e = e1 - 1; e > e0; --e) {                    // rows descending: highest level first
            const int slot = D.colSlot[e];
            const int I = D.slotI[slot];
            if (!have) {
                const double2* g = reinterpret_cast<const double2*>(D.tiles + ((size_t)slot << 12) + c * 64 + 32 * h);
#pragma unroll
                for (int k = 0; k < 16; ++k) v[k] = g[k];
            }
            if (tid == 0) { while (ld_acquire(D.xflag + I) != epoch) { __nanosleep(32); } }
            __syncthreads();
            if (tid < 64) xv[tid] = (I == nT - 1 && tid == 63) ? 0.0 : __ldcg(xs + (size_t)I * 64 + tid);
            __syncthreads();
            const double* xx = xv + 32 * h;
#if TC_BWD_ILP
#pragma unroll
            for (int k = 0; k < 16; k += 2) {                   // four independent chains
                acc = fma(v[k].x, xx[2 * k], acc); acc1 = fma(v[k].y, xx[2 * k + 1], acc1);
                acc2 = fma(v[k + 1].x, xx[2 * k + 2], acc2); acc3 = fma(v[k + 1].y, xx[2 * k + 3], acc3);
            }
#else
#pragma unroll
            for (int k = 0; k < 16; ++k) { acc = fma(v[k].x, xx[2 * k], acc); acc = fma(v[k].y, xx[2 * k + 1], acc); }
#endif
            have = false;
            if (e - 1 > e0) {                                   // prefetch the next tile before the next wait
                const int s2 = D.colSlot[e - 1];
                const double2* g = reinterpret_cast<const double2*>(D.tiles + ((size_t)s2 << 12) + c * 64 + 32 * h);
#pragma unroll
                for (int k = 0; k < 16; ++k) v[k] = g[k];
                have = true;
            }
            __syncthreads();                                    // xv is reused
        }
#if TC_BWD_ILP
        acc = (acc + acc1) + (acc2 + acc3);
#endif
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        __syncthreads();
        if (h == 0) yv[c] -= acc;
        __syncthreads();
        // L(J,J)' x = y by 16-blocks, last block first
#if TC_BWD_ILP
        // every dot product of the chain is cut into short independent pieces: x_p by 64 threads (4 terms + two
        // shuffles), the update of the rows above by all 128 (8 terms + one shuffle)
        for (int p = 3; p >= 0; --p) {
            if (tid < 64) {                                     // x_p = inv(L_pp)' y_p
                const int i = tid >> 2, q = tid & 3;
                double s = 0.0, s2 = 0.0;
                if (q >= i) s = Xd[p * 256 + q * 16 + i] * yv[16 * p + q];
                if (q + 4 >= i) s2 = Xd[p * 256 + (q + 4) * 16 + i] * yv[16 * p + q + 4];
                if (q + 8 >= i) s = fma(Xd[p * 256 + (q + 8) * 16 + i], yv[16 * p + q + 8], s);
                if (q + 12 >= i) s2 = fma(Xd[p * 256 + (q + 12) * 16 + i], yv[16 * p + q + 12], s2);
                s += s2;
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                if (q == 0) xv[16 * p + i] = s;
            }
            __syncthreads();
            {                                                   // y(0:16p) -= L(p rows, 0:16p)' x_p
                double s = 0.0, s2 = 0.0;
                if (c < 16 * p) {
                    const double* lp = Ls + c * BW_LD + 16 * p + 8 * h;
                    const double* xp = xv + 16 * p + 8 * h;
#pragma unroll
                    for (int r = 0; r < 8; r += 2) { s = fma(lp[r], xp[r], s); s2 = fma(lp[r + 1], xp[r + 1], s2); }
                }
                s += s2;
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                if (h == 0 && c < 16 * p) yv[c] -= s;
            }
            __syncthreads();
        }
#else
        for (int p = 3; p >= 0; --p) {
            if (tid < 16) {                                     // x_p = inv(L_pp)' y_p
                double s = 0.0;
                for (int n2 = tid; n2 < 16; ++n2) s = fma(Xd[p * 256 + n2 * 16 + tid], yv[16 * p + n2], s);
                xv[16 * p + tid] = s;
            }
            __syncthreads();
            if (tid < 16 * p) {                                 // y(0:16p) -= L(p rows, 0:16p)' x_p
                double s = 0.0;
#pragma unroll
                for (int r = 0; r < 16; ++r) s = fma(Ls[tid * BW_LD + 16 * p + r], xv[16 * p + r], s);
                yv[tid] -= s;
            }
            __syncthreads();
        }
#endif
        if (tid < 64) xs[(size_t)J * 64 + tid] = xv[tid];
        __syncthreads();
        if (tid == 0) st_release(D.xflag + J, epoch);
    }
}

// ---------------------------------------------------------------------------------------------
// small helpers on the tile array
// ---------------------------------------------------------------------------------------------
// putTop = 0: leave the top tile columns alone (distributed: their tiles are summed over the ranks, so only one
// rank may contribute the rhs row and the 1e300 corner)
// The same launch resets what a factorisation + solve start from: task counters and pivot statistics.
__global__ void k_tc_put_rhs(TCholDev D, const double* __restrict__ rhs, const int* __restrict__ colOwner, int putTop) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < 8) D.counters[s] = 0;
    if (s == 8) *D.info = 0;
    if (s == 9) { D.minmax[0] = 0x7fefffffffffffffull; D.minmax[1] = 0ull; }      // DBL_MAX, 0
    if (s >= D.ld) return;
    if (!putTop && colOwner[s >> 6] < 0) return;
    const int slot = D.tix[(size_t)(D.nT - 1) * D.nT + (s >> 6)];
    D.tiles[((size_t)slot << 12) + (s & 63) * 64 + 63] = (s == D.ld - 1) ? 1e300 : rhs[s];
}
__global__ void k_tc_scale(TCholDev D, const double* __restrict__ dS) {
    const int slot = blockIdx.x < D.nTopS ? blockIdx.x : D.nTop + (blockIdx.x - D.nTopS);     // the tiles of S
    const int I = D.slotI[slot], J = D.slotJ[slot];
    double* t = D.tiles + ((size_t)slot << 12);
    for (int idx = threadIdx.x; idx < 4096; idx += blockDim.x) {
        const int c = idx >> 6, r = idx & 63;
        const int gr = 64 * I + r, gc = 64 * J + c;
        if (gr == D.ld - 1) continue;                           // the rhs row is scaled when it is put
        if (gr >= gc) t[idx] *= dS[gr] * dS[gc];
    }
}
__global__ void k_tc_to_dense(TCholDev D, double* __restrict__ dense, int ldd) {
    const int slot = blockIdx.x < D.nTopS ? blockIdx.x : D.nTop + (blockIdx.x - D.nTopS);     // the tiles of S
    const int I = D.slotI[slot], J = D.slotJ[slot];
    const double* t = D.tiles + ((size_t)slot << 12);
    for (int idx = threadIdx.x; idx < 4096; idx += blockDim.x) {
        const int c = idx >> 6, r = idx & 63;
        const int gr = 64 * I + r, gc = 64 * J + c;
        if (gr >= gc) dense[(size_t)gc * ldd + gr] = t[idx];
    }
}

template <typename T>
static int up(TChol& w, const T** dst, const std::vector<T>& v) {
    T* p = nullptr;
    if (cudaMalloc(&p, sizeof(T) * std::max<size_t>(1, v.size())) != cudaSuccess) return 1;
    w.allocs.push_back(p);
    if (!v.empty() && cudaMemcpy(p, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice) != cudaSuccess) return 1;
    *dst = p;
    return 0;
}
template <typename T>
static int al(TChol& w, T** dst, size_t cnt, bool zero) {
    T* p = nullptr;
    if (cudaMalloc(&p, sizeof(T) * std::max<size_t>(1, cnt)) != cudaSuccess) return 1;
    w.allocs.push_back(p);
    if (zero && cudaMemset(p, 0, sizeof(T) * std::max<size_t>(1, cnt)) != cudaSuccess) return 1;
    *dst = p;
    return 0;
}

int tchol_alloc(TChol& w, const TileSym& sym) {
    w.sym = sym;
    {   // the task ids in list order behind the four queues: the unsplit launch walks them
        TileSym& ms = w.sym;
        for (int t = 0; t < ms.nTasks; ++t) ms.queue.push_back(t);
    }
    const TileSym& s = w.sym;
    TCholDev& d = w.d;
    d.nT = s.nT; d.ld = s.ld; d.nSlots = s.nSlots; d.nSlotsS = s.nSlotsS; d.nTasks = s.nTasks;
    int bad = 0;
    bad |= up(w, &d.tix, s.tix);
    bad |= up(w, &d.slotI, s.slotI); bad |= up(w, &d.slotJ, s.slotJ);
    bad |= up(w, &d.colPtr, s.colPtr); bad |= up(w, &d.colSlot, s.colSlot);
    bad |= up(w, &d.taskI, s.taskI); bad |= up(w, &d.taskJ, s.taskJ); bad |= up(w, &d.taskMode, s.taskMode);
    bad |= up(w, &d.taskWait, s.taskWait); bad |= up(w, &d.taskSet, s.taskSet); bad |= up(w, &d.taskInit, s.taskInit);
    bad |= al(w, &d.aux, (size_t)std::max(1, s.nAux), true);
    d.nTopS = s.nTopS; d.nTop = s.nTop; d.nOwnS = s.nOwnS;
    {
        std::vector<long long> tp(s.termPtr.begin(), s.termPtr.end());
        bad |= up(w, &d.termPtr, tp);
    }
    bad |= up(w, &d.termA, s.termA); bad |= up(w, &d.termB, s.termB);
    bad |= up(w, &d.bwdCols, s.bwdCols);
    bad |= up(w, &d.queue, s.queue);
    bad |= up(w, &w.colOwnerDev, s.colOwner);
    {
        std::vector<unsigned char> valid(s.ld);
        for (int k = 0; k < s.ld; ++k) valid[k] = s.s2kind[k] == 1;
        bad |= up(w, &d.valid, valid);
    }
    bad |= al(w, &d.tiles, (size_t)s.nSlots * TC_TT, true);
    bad |= al(w, &d.invD, (size_t)s.nT * TC_XD, true);
    bad |= al(w, &d.flag, (size_t)s.nSlots, true);
    bad |= al(w, &d.prog, (size_t)s.nSlots, true);
    bad |= al(w, &d.xflag, (size_t)s.nT, true);
    bad |= al(w, &d.counters, 8, true);
    bad |= al(w, &d.info, 1, true);
    bad |= al(w, &d.minmax, 2, true);
    bad |= al(w, &w.xs, (size_t)s.ld, true);
    if (bad) { tchol_free(w); return 1; }
    w.epoch = 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int smem = TC_SMEM_DOUBLES * 8;
    cudaFuncSetAttribute(k_tchol_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_tchol_factor, TC_THREADS, smem);
    if (occ < 1) occ = 1;
    w.gridFactor = std::max(1, std::min(std::max(s.nTasks1, s.nTasks - s.nTasks1), sms * occ));
    // chain CTAs: a shared-memory request above half an SM keeps every other CTA off their SM
    int smemMax = 0;
    cudaDeviceGetAttribute(&smemMax, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    w.smemChain = std::min(smemMax, 160 * 1024);
    {
        const char* e = getenv("DBAT_TC_CHAIN_CTAS");
        w.nCrit = e ? atoi(e) : 24;
        if (w.smemChain < 120 * 1024 || sms < 4 * w.nCrit) w.nCrit = 0;
        if (w.nCrit > 0) w.gridFactor = std::max(1, std::min(w.gridFactor, (sms - w.nCrit) * occ));
    }
    cudaFuncSetAttribute(k_tchol_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, std::max(smem, w.smemChain));
    int occ2 = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k_tchol_bwd, TC_THREADS, 0);
    if (occ2 < 1) occ2 = 1;
    w.gridBwd = std::max(1, std::min(s.nT, sms * occ2));
    return 0;
}
void tchol_free(TChol& w) {
    for (void* p : w.allocs) cudaFree(p);
    w.allocs.clear();
    w.d = TCholDev();
    w.xs = nullptr;
}
void tchol_zero_dev(const TCholDev& d, cudaStream_t st) {
    // one part: the S tiles [0, nTopS).  Distributed: every top tile (the fill tiles among them receive partial sums
    // that are added up over the ranks, so they must not keep the previous factor) and the owned S tiles.
    const size_t top = d.nTop < d.nSlots ? d.nTop : d.nTopS;
    cudaMemsetAsync(d.tiles, 0, sizeof(double) * top * TC_TT, st);
    if (d.nOwnS > 0) cudaMemsetAsync(d.tiles + (size_t)d.nTop * TC_TT, 0, sizeof(double) * (size_t)d.nOwnS * TC_TT, st);
}
void tchol_zero(TChol& w, cudaStream_t st) { tchol_zero_dev(w.d, st); }
void tchol_put_rhs(TChol& w, const double* rhs, cudaStream_t st, bool putTop) {
    k_tc_put_rhs<<<(w.d.ld + 255) / 256, 256, 0, st>>>(w.d, rhs, w.colOwnerDev, putTop ? 1 : 0);
    w.resetDone = true;
    count_launch();
}
// [~min pivot bits, max pivot bits, info] as three uint64: one max-allreduce combines the statistics of all ranks
__global__ void k_tc_pack_stats(TCholDev D, unsigned long long* __restrict__ out, int unpack) {
    if (threadIdx.x != 0) return;
    if (!unpack) { out[0] = ~D.minmax[0]; out[1] = D.minmax[1]; out[2] = (unsigned long long)(unsigned)(*D.info); }
    else { D.minmax[0] = ~out[0]; D.minmax[1] = out[1]; *D.info = (int)out[2]; }
}
void tchol_pack_stats(TChol& w, unsigned long long* buf, bool unpack, cudaStream_t st) {
    k_tc_pack_stats<<<1, 32, 0, st>>>(w.d, buf, unpack ? 1 : 0);
    count_launch();
}
// One phase: the chain queue on w.nCrit CTAs that own an SM each (their dynamic shared memory request leaves no room
// for a second CTA), the bulk queue on a second launch in the same stream with programmatic stream serialisation: it
// starts once every chain CTA is resident and has executed griddepcontrol.launch_dependents - so the chain CTAs can
// never be locked out by spinning bulk CTAs.  Small systems: one launch serves both queues.
static void factor_phase(TChol& w, int ph, cudaStream_t st) {
    const TileSym& s = w.sym;
    const int c0 = s.qOff[2 * ph], b0 = s.qOff[2 * ph + 1], b1 = s.qOff[2 * ph + 2];
    if (b1 <= c0) return;
    const int smemBulk = TC_SMEM_DOUBLES * 8;
    if (w.nCrit <= 0 || b0 - c0 < 4 * w.nCrit) {
        // no split: the two queues are served one after the other by the same CTAs would break the priority order,
        // so merge them back by task id
        k_tchol_factor<<<std::max(1, std::min(w.gridFactor, b1 - c0)), TC_THREADS, smemBulk, st>>>(w.d, w.epoch, s.qOff[4] + (ph == 0 ? 0 : s.nTasks1),
            s.qOff[4] + (ph == 0 ? s.nTasks1 : s.nTasks), 2 * ph, 0);
        count_launch();
        return;
    }
    k_tchol_factor<<<w.nCrit, TC_THREADS, w.smemChain, st>>>(w.d, w.epoch, c0, b0, 2 * ph, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)std::max(1, std::min(w.gridFactor, b1 - b0)));
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smemBulk;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k_tchol_factor, w.d, w.epoch, b0, b1, 2 * ph + 1, 0);
    count_launch(2);
}
void tchol_factor_begin(TChol& w, cudaStream_t st) {
    ++w.epoch;
    if (!w.resetDone) {                       // normally done by the tchol_put_rhs launch just before
        cudaMemsetAsync(w.d.counters, 0, sizeof(int) * 8, st);
        cudaMemsetAsync(w.d.info, 0, sizeof(int), st);
        static const unsigned long long init[2] = {0x7fefffffffffffffull, 0ull};      // DBL_MAX, 0
        cudaMemcpyAsync(w.d.minmax, init, sizeof(init), cudaMemcpyHostToDevice, st);
    }
    w.resetDone = false;
    factor_phase(w, 0, st);
}
void tchol_factor_end(TChol& w, cudaStream_t st) {
    factor_phase(w, 1, st);
}
void tchol_factor(TChol& w, cudaStream_t st) {
    tchol_factor_begin(w, st);
    tchol_factor_end(w, st);
}
void tchol_solve(TChol& w, cudaStream_t st) {
    const int nCols = (int)w.sym.bwdCols.size();
    k_tchol_bwd<<<std::max(1, std::min(w.gridBwd, nCols)), TC_THREADS, 0, st>>>(w.d, w.epoch, w.xs, nCols);
    count_launch();
}
__global__ void k_tc_mask_x(TCholDev D, const int* __restrict__ colOwner, int me, int keepTop, double* __restrict__ xs) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= D.ld) return;
    const int ow = colOwner[s >> 6];
    if (!((ow < 0 && keepTop) || (ow >= 0 && ow == me))) xs[s] = 0.0;
}
void tchol_solve_owned_mask(TChol& w, bool keepTop, cudaStream_t st) {
    k_tc_mask_x<<<(w.d.ld + 255) / 256, 256, 0, st>>>(w.d, w.colOwnerDev, w.sym.myPart, keepTop ? 1 : 0, w.xs);
    count_launch();
}
void tchol_scale(TChol& w, const double* dS, cudaStream_t st) {
    k_tc_scale<<<w.d.nSlotsS, 256, 0, st>>>(w.d, dS);
    count_launch();
}
void tchol_to_dense(TChol& w, double* dense, int ldd, cudaStream_t st) {
    k_tc_to_dense<<<w.d.nSlotsS, 256, 0, st>>>(w.d, dense, ldd);      // fill slots hold stale factors, S has no entries there
    count_launch();
}
void tchol_pivot_stats(TChol& w, int* info, double* mn, double* mx, cudaStream_t st) {
    unsigned long long mm[2] = {0, 0};
    cudaMemcpyAsync(info, w.d.info, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(mm, w.d.minmax, sizeof(mm), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    double a, b;
    memcpy(&a, &mm[0], 8); memcpy(&b, &mm[1], 8);
    *mn = a; *mx = b;
}
