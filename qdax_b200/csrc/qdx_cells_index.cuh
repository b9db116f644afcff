// Exact nearest-centroid search through a uniform bucket index, for low-dimensional CVT tessellations (Dd <= 3).
//
// get_cells_indices (qdax/core/containers/mapelites_repertoire.py:111-137 under /root/reference) is
// argmin_k sum_d (x_d - c_kd)^2 with the first index on ties: O(K) per descriptor.  For K = 10^4 two-dimensional
// centroids that is 10^4 distance evaluations where ~30 decide the answer.  The centroids are bucketed once per
// tessellation into a uniform grid over their bounding box (~2 per bucket), stored sorted by bucket (ascending
// centroid id inside a bucket) so that a row of buckets is one contiguous range.  A query walks Chebyshev rings
// around its own bucket, evaluates the candidates with the reference expression (same float32 operations, same
// order, lexicographic (distance, id) minimum -> identical ties), and stops once the best distance found is
// strictly below a conservative lower bound on the distance to anything outside the searched box.  The bound
// carries margins for every rounding it depends on (bucket boundaries: 1e-3 h; squared comparison: 1e-5 relative,
// against <= 3 ulp error of a float32 distance), so the result equals the brute-force argmin bit for bit.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define QDX_INDEX_MAX_DIM 3

struct QdxCvtIndex {
    int32_t dd;                        // 0 = no index
    int32_t g[QDX_INDEX_MAX_DIM];      // buckets per dimension; bucket id = b0 + g0 * (b1 + g1 * b2)
    float lo[QDX_INDEX_MAX_DIM];       // lower corner of the bounding box
    float h[QDX_INDEX_MAX_DIM];        // bucket width (> 0)
    const int32_t* start;              // device, prod(g) + 1 offsets into ids / pts
    const int32_t* ids;                // device, K centroid ids sorted by (bucket, id)
    const float* pts;                  // device, K x dd centroid coordinates in the same order
};

template <int DD>
__device__ __forceinline__ void qdx_index_scan_range(const float* x, const QdxCvtIndex& ix, int32_t s, int32_t e, float& best, int32_t& bid) {
    for (int32_t k = s; k < e; ++k) {
        float acc;
#pragma unroll
        for (int d = 0; d < DD; ++d) { const float df = x[d] - __ldg(ix.pts + (int64_t)k * DD + d); const float sq = df * df; acc = d ? acc + sq : sq; }
        const int32_t id = __ldg(ix.ids + k);
        if (acc < best || (acc == best && id < bid)) { best = acc; bid = id; }
    }
}

// Cell of descriptor x (finite or not): identical to the brute-force first-index argmin over all K centroids.
template <int DD>
__device__ __forceinline__ int32_t qdx_index_cell(const float* x, const QdxCvtIndex& ix) {
    int32_t b[DD];
#pragma unroll
    for (int d = 0; d < DD; ++d) {
        if (!(fabsf(x[d]) <= 3.40282347e+38f)) return 0;     // NaN / inf: every distance NaN or inf -> first index
        const float t = floorf((x[d] - ix.lo[d]) / ix.h[d]);
        b[d] = t < 0.0f ? 0 : (t >= (float)ix.g[d] ? ix.g[d] - 1 : (int32_t)t);
    }
    float best = INFINITY; int32_t bid = 0x7fffffff;
    int rmax = 0;
#pragma unroll
    for (int d = 0; d < DD; ++d) { const int m = b[d] > ix.g[d] - 1 - b[d] ? b[d] : ix.g[d] - 1 - b[d]; rmax = m > rmax ? m : rmax; }
    for (int r = 0; r <= rmax; ++r) {
        // ---- ring r: buckets at Chebyshev distance exactly r.  Outer dims (1..DD-1) enumerate rows of buckets; a row is
        // contiguous along dim 0, so it is either one full range [b0-r, b0+r] (row on the ring's face) or its two ends.
        const int i0 = b[0] - r < 0 ? 0 : b[0] - r, i1 = b[0] + r > ix.g[0] - 1 ? ix.g[0] - 1 : b[0] + r;
        const int j0 = DD > 1 ? (b[DD > 1 ? 1 : 0] - r < 0 ? 0 : b[DD > 1 ? 1 : 0] - r) : 0;
        const int j1 = DD > 1 ? (b[DD > 1 ? 1 : 0] + r > ix.g[DD > 1 ? 1 : 0] - 1 ? ix.g[DD > 1 ? 1 : 0] - 1 : b[DD > 1 ? 1 : 0] + r) : 0;
        const int k0 = DD > 2 ? (b[DD > 2 ? 2 : 0] - r < 0 ? 0 : b[DD > 2 ? 2 : 0] - r) : 0;
        const int k1 = DD > 2 ? (b[DD > 2 ? 2 : 0] + r > ix.g[DD > 2 ? 2 : 0] - 1 ? ix.g[DD > 2 ? 2 : 0] - 1 : b[DD > 2 ? 2 : 0] + r) : 0;
        for (int k = k0; k <= k1; ++k)
            for (int j = j0; j <= j1; ++j) {
                int face = 0;                                    // is this row at distance r in an outer dimension?
                if (DD > 1) { const int dj = j - b[DD > 1 ? 1 : 0]; face |= (dj == r) | (dj == -r); }
                if (DD > 2) { const int dk = k - b[DD > 2 ? 2 : 0]; face |= (dk == r) | (dk == -r); }
                const int64_t row = (int64_t)ix.g[0] * (j + (int64_t)(DD > 1 ? ix.g[DD > 1 ? 1 : 0] : 1) * k);
                if (face || r == 0) {
                    qdx_index_scan_range<DD>(x, ix, __ldg(ix.start + row + i0), __ldg(ix.start + row + i1 + 1), best, bid);
                } else {
                    if (b[0] - r >= 0) qdx_index_scan_range<DD>(x, ix, __ldg(ix.start + row + b[0] - r), __ldg(ix.start + row + b[0] - r + 1), best, bid);
                    if (b[0] + r <= ix.g[0] - 1) qdx_index_scan_range<DD>(x, ix, __ldg(ix.start + row + b[0] + r), __ldg(ix.start + row + b[0] + r + 1), best, bid);
                }
            }
        // ---- stop when nothing outside the searched box [b-r, b+r] can beat (or tie) the best found
        float m = INFINITY;                                      // lower bound on the distance to any unsearched centroid
#pragma unroll
        for (int d = 0; d < DD; ++d) {
            if (b[d] - r > 0) { const float side = x[d] - (ix.lo[d] + (float)(b[d] - r) * ix.h[d]) - 1e-3f * ix.h[d]; m = fminf(m, side); }
            if (b[d] + r < ix.g[d] - 1) { const float side = (ix.lo[d] + (float)(b[d] + r + 1) * ix.h[d]) - x[d] - 1e-3f * ix.h[d]; m = fminf(m, side); }
        }
        if (m == INFINITY) break;                                // the box covers the whole grid
        if (m > 0.0f && best < m * m * 0.99999f) break;
    }
    return bid == 0x7fffffff ? 0 : bid;                          // all distances inf (huge finite x): first index
}
