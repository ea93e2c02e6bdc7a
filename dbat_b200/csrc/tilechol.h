// tilechol.h — sparse tile Cholesky of the reduced camera system: device view + host API.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "tilesym.h"

#define TC_XLD 20                 // row stride of the 16 x 16 inverse diagonal blocks (= 4 mod 16)
#define TC_XD (4 * 16 * TC_XLD)   // doubles per diagonal tile

// Device view (all pointers are device pointers).  Tile (I, J), I >= J, lives in slot tix[I * nT + J]:
// 64 x 64 doubles, column-major.  S index s = 64 * tile + offset; the last row ld-1 carries the rhs.
struct TCholDev {
    double* tiles = nullptr;
    double* invD = nullptr;            // nT x TC_XD: inverse of the four 16 x 16 diagonal blocks of L(J,J)
    int* flag = nullptr;               // nSlots: == epoch when the tile holds its final L values
    int* prog = nullptr;               // nSlots: 8 * epoch + k once the first k 16-column panels of the tile are final (k = 1..4)
    int* xflag = nullptr;              // nT: == epoch when block J of the backward solution is published
    int* counters = nullptr;           // [0] next factor task, [1] next backward column
    int* info = nullptr;               // != 0: a pivot was not positive
    unsigned long long* minmax = nullptr;   // [0] smallest, [1] largest pivot (bit patterns of positive doubles)
    const int* tix = nullptr;
    const int* slotI = nullptr; const int* slotJ = nullptr;
    const int* colPtr = nullptr; const int* colSlot = nullptr;
    const int* taskI = nullptr; const int* taskJ = nullptr; const unsigned char* taskMode = nullptr;
    const int* taskWait = nullptr; const int* taskSet = nullptr; const unsigned char* taskInit = nullptr;
    const int* queue = nullptr;        // task ids of the chain / bulk queues of both phases
    int* aux = nullptr;                // nAux flags of the partial-sum chains (== epoch when the link is done)
    const long long* termPtr = nullptr; const int* termA = nullptr; const int* termB = nullptr;
    const int* bwdCols = nullptr;
    const unsigned char* valid = nullptr;   // per S index: 1 = real unknown (enters the pivot statistics)
    unsigned long long* prof = nullptr;     // optional: 4 globaltimer stamps per task (claim, terms done, deps done, end)
    int nT = 0, ld = 0, nSlots = 0, nSlotsS = 0, nTasks = 0;
    int nTopS = 0, nTop = 0, nOwnS = 0;     // slot ranges: S tiles are [0, nTopS) and [nTop, nTop + nOwnS)
    __host__ __device__ bool inS(int slot) const { return slot < nTopS || (slot >= nTop && slot < nTop + nOwnS); }
};

struct TChol {
    TileSym sym;
    TCholDev d;
    int epoch = 0;
    bool resetDone = false;            // tchol_put_rhs has reset the task counters / pivot statistics for the next factorisation
    int smemChain = 0;
    int gridFactor = 0, gridBwd = 0, nCrit = 0;    // nCrit > 0: that many CTAs, alone on their SMs, serve the chain queue
    double* xs = nullptr;              // ld: solution in S order
    const int* colOwnerDev = nullptr;  // nT: owning part of every tile column (-1 = top)
    std::vector<void*> allocs;
};

int tchol_alloc(TChol& w, const TileSym& sym);           // uploads the symbolic data, allocates tiles / flags
void tchol_free(TChol& w);
// zero the tiles of the pattern of S (fill tiles are never read before they are written)
void tchol_zero(TChol& w, cudaStream_t st);
void tchol_zero_dev(const TCholDev& d, cudaStream_t st);
// row ld-1 := rhs (S order, length ld), S(ld-1, ld-1) := 1e300: forward substitution rides in the factorisation
void tchol_put_rhs(TChol& w, const double* rhs, cudaStream_t st, bool putTop = true);
void tchol_pack_stats(TChol& w, unsigned long long* buf3, bool unpack, cudaStream_t st);
// in-place factorisation; afterwards w.d.info / w.d.minmax hold the pivot statistics.
// Distributed (w.sym.nParts > 1): tchol_factor_begin runs phase 1 (this part's own columns and its partial sums
// into the top tiles); the caller then sums the top tiles [0, nTop) over the ranks and calls tchol_factor_end.
void tchol_factor(TChol& w, cudaStream_t st);
void tchol_factor_begin(TChol& w, cudaStream_t st);
void tchol_factor_end(TChol& w, cudaStream_t st);
// after tchol_put_rhs + tchol_factor: w.xs = S^-1 rhs (S order).  Distributed: only the top blocks and this part's
// own blocks of w.xs are computed (tchol_solve_owned_mask zeroes the rest so that a sum over the ranks completes it)
void tchol_solve(TChol& w, cudaStream_t st);
void tchol_solve_owned_mask(TChol& w, bool keepTop, cudaStream_t st);
// S := D S D and rhs-row scaling with d given per S index (Jacobi scaling; padding entries must be 1)
void tchol_scale(TChol& w, const double* dS, cudaStream_t st);
// dense copy of S (before any factorisation touched the tiles: the slots of the pattern of S only; column-major,
// ldd >= ld, lower triangle, `dense` zeroed by the caller) for the covariance path
void tchol_to_dense(TChol& w, double* dense, int ldd, cudaStream_t st);
void tchol_pivot_stats(TChol& w, int* info, double* mn, double* mx, cudaStream_t st);

// device-side address of S(r, c), r >= c in S order; the tile must exist in the symbolic pattern
__device__ __forceinline__ double* tc_at(const TCholDev& d, int r, int c) {
    const int slot = d.tix[(size_t)(r >> 6) * d.nT + (c >> 6)];
    return d.tiles + ((size_t)slot << 12) + ((c & 63) << 6) + (r & 63);
}
