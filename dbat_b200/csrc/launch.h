// launch.h — host-side launchers (one per kernel family) shared by the translation units.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

struct DevProblem;

// every kernel launch of the library is counted (reported as gpu_launches by bench.py)
extern int64_t g_dbat_launches;
static inline void count_launch(int n = 1) { __atomic_fetch_add(&g_dbat_launches, (int64_t)n, __ATOMIC_RELAXED); }

// eval.cu
void launch_deserialize(const double* x, const int* src, const int* dest, double* arr, int cnt, cudaStream_t st);
void launch_param_setup(const DevProblem& P, const int* rep, int nIO, cudaStream_t st);
struct ScatterList { const int* src; const int* dest; double* arr; int cnt; int nb; };
void launch_set_params(const DevProblem& P, const double* x, ScatterList io, ScatterList eo, ScatterList op,
                       const int* rep, int nIO, cudaStream_t st);
void launch_cam_side(const DevProblem& P, const int* img_chunk_start, double* tmp, cudaStream_t st);
void launch_point_side(const DevProblem& P, cudaStream_t st);
void launch_prior_apply(const DevProblem& P, const double* x, double* camDiag, double* camG,
                        const int* col2pt, bool clear, cudaStream_t st);
void launch_resid(const DevProblem& P, const double* x, double* partial, double* scal, int slot,
                  double* r_out, int weighted, cudaStream_t st);
void launch_prior_rr(const DevProblem& P, const double* x, double* partial, double* scal, int slot, bool clear, cudaStream_t st);
void launch_jp(const DevProblem& P, const double* x, const double* p, double* partial, double* scal,
               int slot2, int slotr, cudaStream_t st);
void launch_export_jac(const DevProblem& P, double* out, int weighted, cudaStream_t st);
void launch_export_mask(const DevProblem& P, unsigned long long* out, cudaStream_t st);

// schur.cu
void launch_diag(const DevProblem& P, const double* camDiag, double* diagN, cudaStream_t st);
void launch_build_S(const DevProblem& P, const double* camDiag, const double* camG, double lambda,
                    cudaStream_t st);
void launch_schur(const DevProblem& P, double lambda, cudaStream_t st);
void launch_point_vinv(const DevProblem& P, double lambda, cudaStream_t st);   // P.vinv = (V_j + lambda I)^-1
void launch_scale_prep(const DevProblem& P, const double* d, double* dS, cudaStream_t st);
void launch_unpermute(const DevProblem& P, const double* xs, const double* d, double* pc, double* pEO, cudaStream_t st);
void launch_backsub(const DevProblem& P, double lambda, const double* pc, const double* pEO, double* p, double* partial,
                    const double* camDiag, const double* camG, double* jpOut, cudaStream_t st,
                    cudaStream_t st2 = nullptr, cudaEvent_t evFork = nullptr, cudaEvent_t evJoin = nullptr);
void launch_vec_ops_sum(const double* a, int n, double* partial, double* scal, int slot, cudaStream_t st);
void launch_dot(const double* a, const double* b, int n, double* partial, double* scal, int slot, cudaStream_t st);
void launch_axpy(double alpha, const double* x, const double* y, double* out, int n, cudaStream_t st);
void launch_inv_sqrt(const double* in, double* out, int n, cudaStream_t st);
void launch_mul(const double* a, const double* b, double* out, int n, cudaStream_t st);

// schur_win.cu (default Schur reduction: window accumulation per point cluster + fixed-order sums, no atomics)
void launch_schur_win(const DevProblem& P, double lambda, const double* shAcc, cudaStream_t st);

// general_io.cu (general IO block structure)
void launch_point_side_gen(const DevProblem& P, cudaStream_t st);
void launch_build_S_gen(const DevProblem& P, const double* camDiag, const double* camG, double lambda, cudaStream_t st);
void launch_schur_gen(const DevProblem& P, double lambda, cudaStream_t st);
int launch_backsub_gen(const DevProblem& P, double lambda, double* p, double* stats, cudaStream_t st);
void launch_cop_gen(const DevProblem& P, const double* C, int ldc, const double* dsc, double s02, double* out, cudaStream_t st);
void launch_io_diag_grad_gen(const DevProblem& P, const double* camDiag, const double* camG, double* diagN, double* grad,
                             cudaStream_t st);

// schur_index.cu
int build_pair_index(const int* d_pt_start, const int* d_img_pm, const long long* h_pair_off, int nOP,
                     int nImg, long long** d_pairs, long long** d_blk_key, long long** d_blk_off, int* nBlk,
                     cudaStream_t st);

// chol.cu
struct CholWork {
    int n = 0, ld = 0, nb = 0;      // order, leading dimension (multiple of 128), # 128-blocks
    double* invL = nullptr;         // nb x 128 x 128 inverses of the diagonal blocks
    int* info = nullptr;            // device flag: 0 ok, k>0 = first non-positive pivot (1-based)
    double* minmax = nullptr;       // device: [0]=min pivot, [1]=max pivot (of L's diagonal)
    void* graphExec = nullptr;      // captured launch sequence of chol_factor (cudaGraphExec_t)
    const double* graphA = nullptr; const double* seen = nullptr; cudaStream_t graphStream = nullptr;
    int graphLaunches = 0;
};
void chol_alloc(CholWork& w, int n, int ld);
void chol_free(CholWork& w);
// In-place lower Cholesky of the ld x ld column-major matrix A (rows/cols >= n are padding and
// must hold the identity).  Asynchronous; read w.info afterwards.
void chol_factor(CholWork& w, double* A, cudaStream_t st);
// Store rhs' in the last padding row of A (before chol_factor): the forward substitution then
// happens inside the factorisation.  Requires ld > n.
void chol_put_rhs(const CholWork& w, double* A, const double* rhs, cudaStream_t st);
// After chol_factor on a matrix prepared with chol_put_rhs: x (length ld) = A^-1 rhs.
void chol_solve(const CholWork& w, const double* A, double* x, cudaStream_t st);
// Z = inv(L) (lower) into Zout (ld x ld); then C = Z'Z (= inv(A)) full symmetric into Cout.
void chol_inverse(const CholWork& w, const double* A, double* Z, double* C, cudaStream_t st);
