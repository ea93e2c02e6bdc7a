// schur_index.cu — one-off construction of the camera-pair contribution index used by the
// deterministic Schur reduction: for every object point, every pair (oA, oB) of its observations
// with image(oA) >= image(oB) contributes Y_oA * W_oB' to the 6x6 block (image(oA), image(oB)) of
// the reduced camera system.  Pairs are sorted by block key (stable radix sort, so inside a block
// they stay in point order) and run-length encoded into block offsets.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_scan.cuh>
#include "kernels.cuh"
#include "launch.h"

__global__ void k_gen_pairs(const int* __restrict__ pt_start, const int* __restrict__ img_pm,
                            const long long* __restrict__ pair_off, int nOP, long long nImg,
                            long long* __restrict__ keys, long long* __restrict__ vals) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nOP) return;
    const int o0 = pt_start[j], o1 = pt_start[j + 1];
    long long w = pair_off[j];
    for (int a = o0; a < o1; ++a)
        for (int b = o0; b <= a; ++b) {            // images ascend inside a point: img(a) >= img(b)
            keys[w] = (long long)img_pm[a] * nImg + img_pm[b];
            vals[w] = ((long long)a << 32) | (unsigned int)b;
            ++w;
        }
}

// Returns 0 on success.  Outputs are device allocations owned by the caller (cudaFree).
int build_pair_index(const int* d_pt_start, const int* d_img_pm, const long long* h_pair_off, int nOP,
                     int nImg, long long** d_pairs, long long** d_blk_key, long long** d_blk_off, int* nBlk,
                     cudaStream_t st) {
    const long long nPairs = h_pair_off[nOP];
    *d_pairs = nullptr; *d_blk_key = nullptr; *d_blk_off = nullptr; *nBlk = 0;
    if (nPairs == 0) {
        cudaMalloc(d_pairs, 8); cudaMalloc(d_blk_key, 8); cudaMalloc(d_blk_off, 16);
        cudaMemset(*d_blk_off, 0, 16);
        return 0;
    }
    long long *pair_off = nullptr, *k0 = nullptr, *k1 = nullptr, *v0 = nullptr, *v1 = nullptr;
    long long *ukeys = nullptr, *counts = nullptr, *offs = nullptr;
    int* dnum = nullptr;
    void* tmp = nullptr;
    size_t tmpBytes = 0, need = 0;
    int rc = 1;
    do {
        if (cudaMalloc(&pair_off, sizeof(long long) * (nOP + 1)) != cudaSuccess) break;
        cudaMemcpyAsync(pair_off, h_pair_off, sizeof(long long) * (nOP + 1), cudaMemcpyHostToDevice, st);
        if (cudaMalloc(&k0, 8 * nPairs) || cudaMalloc(&k1, 8 * nPairs) || cudaMalloc(&v0, 8 * nPairs) ||
            cudaMalloc(&v1, 8 * nPairs)) break;
        k_gen_pairs<<<(nOP + 127) / 128, 128, 0, st>>>(d_pt_start, d_img_pm, pair_off, nOP, nImg, k0, v0);
        int bits = 1;
        while ((1ll << bits) < (long long)nImg * nImg && bits < 62) ++bits;
        cub::DeviceRadixSort::SortPairs(nullptr, need, k0, k1, v0, v1, nPairs, 0, bits, st);
        tmpBytes = need;
        if (cudaMalloc(&tmp, tmpBytes) != cudaSuccess) break;
        cub::DeviceRadixSort::SortPairs(tmp, need, k0, k1, v0, v1, nPairs, 0, bits, st);
        // run-length encode the sorted keys: unique block keys + counts
        if (cudaMalloc(&ukeys, 8 * nPairs) || cudaMalloc(&counts, 8 * (nPairs + 1)) || cudaMalloc(&dnum, sizeof(int))) break;
        size_t need2 = 0;
        cub::DeviceRunLengthEncode::Encode(nullptr, need2, k1, ukeys, counts, dnum, nPairs, st);
        if (need2 > tmpBytes) { cudaFree(tmp); tmp = nullptr; tmpBytes = need2; if (cudaMalloc(&tmp, tmpBytes) != cudaSuccess) break; }
        cub::DeviceRunLengthEncode::Encode(tmp, need2, k1, ukeys, counts, dnum, nPairs, st);
        int hnum = 0;
        cudaMemcpyAsync(&hnum, dnum, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) break;
        if (cudaMalloc(&offs, 8 * (size_t)(hnum + 1)) != cudaSuccess) break;
        size_t need3 = 0;
        cudaMemsetAsync(counts + hnum, 0, 8, st);
        cub::DeviceScan::ExclusiveSum(nullptr, need3, counts, offs, hnum + 1, st);
        if (need3 > tmpBytes) { cudaFree(tmp); tmp = nullptr; tmpBytes = need3; if (cudaMalloc(&tmp, tmpBytes) != cudaSuccess) break; }
        cub::DeviceScan::ExclusiveSum(tmp, need3, counts, offs, hnum + 1, st);
        long long* keysOut = nullptr;
        if (cudaMalloc(&keysOut, 8 * (size_t)hnum) != cudaSuccess) break;
        cudaMemcpyAsync(keysOut, ukeys, 8 * (size_t)hnum, cudaMemcpyDeviceToDevice, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) { cudaFree(keysOut); break; }
        *d_pairs = v1; v1 = nullptr;
        *d_blk_key = keysOut;
        *d_blk_off = offs; offs = nullptr;
        *nBlk = hnum;
        rc = 0;
    } while (0);
    cudaFree(pair_off); cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1);
    cudaFree(ukeys); cudaFree(counts); cudaFree(offs); cudaFree(dnum); cudaFree(tmp);
    return rc;
}
