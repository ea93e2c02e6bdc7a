// Host-side ordering of the reduced camera system (groundwork for a banded factorisation).
//
// S has a non-zero 6x6 block (i, j) exactly when images i and j observe a common object point.  With the
// images in reverse Cuthill-McKee order of that co-visibility graph and the shared IO block last, S is a band
// matrix with a dense border and its Cholesky factor stays inside the band (tests/test_reduced_structure.py).
// No device code here: this runs once per problem on the host.
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <queue>
#include <vector>

#include "../../include/dbat_gpu.h"

extern "C" int dbat_camera_order(int64_t nImg, int64_t nOP, int64_t nObs, const int64_t* obs_img,
                                 const int64_t* obs_op, int64_t* perm, int64_t* bandwidth) {
    if (nImg <= 0 || nOP < 0 || nObs < 0 || !perm || (nObs > 0 && (!obs_img || !obs_op))) return DBAT_E_BADARG;
    // image lists per point (counting sort by point)
    std::vector<int64_t> start((size_t)nOP + 1, 0);
    for (int64_t o = 0; o < nObs; ++o) {
        const int64_t j = obs_op[o] - 1, i = obs_img[o] - 1;
        if (j < 0 || j >= nOP || i < 0 || i >= nImg) return DBAT_E_BADARG;
        ++start[(size_t)j + 1];
    }
    std::partial_sum(start.begin(), start.end(), start.begin());
    std::vector<int32_t> imgs((size_t)nObs);
    {
        std::vector<int64_t> fill(start.begin(), start.end() - 1);
        for (int64_t o = 0; o < nObs; ++o) imgs[(size_t)fill[(size_t)obs_op[o] - 1]++] = (int32_t)(obs_img[o] - 1);
    }
    // undirected edges (a < b) of the co-visibility graph, deduplicated
    std::vector<uint64_t> edges;
    for (int64_t j = 0; j < nOP; ++j) {
        const int64_t a0 = start[(size_t)j], a1 = start[(size_t)j + 1];
        for (int64_t a = a0; a < a1; ++a)
            for (int64_t b = a + 1; b < a1; ++b) {
                const uint32_t x = (uint32_t)imgs[(size_t)a], y = (uint32_t)imgs[(size_t)b];
                if (x != y) edges.push_back(((uint64_t)std::min(x, y) << 32) | std::max(x, y));
            }
        if (edges.size() > ((size_t)1 << 24)) {            // keep the working set bounded
            std::sort(edges.begin(), edges.end());
            edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
        }
    }
    std::sort(edges.begin(), edges.end());
    edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
    std::vector<int64_t> off((size_t)nImg + 1, 0);
    for (uint64_t e : edges) { ++off[(size_t)(e >> 32) + 1]; ++off[(size_t)(e & 0xffffffffu) + 1]; }
    std::partial_sum(off.begin(), off.end(), off.begin());
    std::vector<int32_t> adj((size_t)off[(size_t)nImg]);
    {
        std::vector<int64_t> fill(off.begin(), off.end() - 1);
        for (uint64_t e : edges) {
            const int32_t x = (int32_t)(e >> 32), y = (int32_t)(e & 0xffffffffu);
            adj[(size_t)fill[(size_t)x]++] = y;
            adj[(size_t)fill[(size_t)y]++] = x;
        }
    }
    auto degree = [&](int32_t v) { return off[(size_t)v + 1] - off[(size_t)v]; };
    // Cuthill-McKee: components in order of their lowest-degree node, neighbours by increasing degree
    std::vector<int32_t> byDeg((size_t)nImg);
    std::iota(byDeg.begin(), byDeg.end(), 0);
    std::stable_sort(byDeg.begin(), byDeg.end(), [&](int32_t a, int32_t b) { return degree(a) < degree(b); });
    std::vector<char> seen((size_t)nImg, 0);
    std::vector<int32_t> order;
    order.reserve((size_t)nImg);
    std::vector<int32_t> nb;
    for (int32_t root : byDeg) {
        if (seen[(size_t)root]) continue;
        seen[(size_t)root] = 1;
        size_t head = order.size();
        order.push_back(root);
        while (head < order.size()) {
            const int32_t v = order[head++];
            nb.clear();
            for (int64_t k = off[(size_t)v]; k < off[(size_t)v + 1]; ++k)
                if (!seen[(size_t)adj[(size_t)k]]) { seen[(size_t)adj[(size_t)k]] = 1; nb.push_back(adj[(size_t)k]); }
            std::stable_sort(nb.begin(), nb.end(), [&](int32_t a, int32_t b) { return degree(a) < degree(b); });
            order.insert(order.end(), nb.begin(), nb.end());
        }
    }
    std::reverse(order.begin(), order.end());
    std::vector<int64_t> pos((size_t)nImg);
    for (int64_t k = 0; k < nImg; ++k) { perm[k] = order[(size_t)k] + 1; pos[(size_t)order[(size_t)k]] = k; }
    if (bandwidth) {
        int64_t bw = 0;
        for (uint64_t e : edges) {
            const int64_t d = pos[(size_t)(e >> 32)] - pos[(size_t)(e & 0xffffffffu)];
            bw = std::max<int64_t>(bw, d < 0 ? -d : d);
        }
        *bandwidth = bw;
    }
    return DBAT_OK;
}
