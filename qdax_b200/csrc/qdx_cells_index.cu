// Bucket index over low-dimensional CVT centroids: host-side builder + standalone assignment kernel.
// The search itself (qdx_index_cell, qdx_cells_index.cuh) is also fused into the generate kernel.
// Reference semantics: get_cells_indices, qdax/core/containers/mapelites_repertoire.py:111-137 (under /root/reference).
#include <cmath>
#include <cstring>
#include <vector>
#include "qdx_common.cuh"
#include "qdx_cells_index.cuh"
#include "../../include/qdx.h"

template <int DD>
__global__ void __launch_bounds__(256) qdx_cells_index_kernel(const float* __restrict__ desc, int64_t B, const QdxCvtIndex ix, int64_t K,
                                                              int32_t* __restrict__ cells, void* ws, const float* rep_f,
                                                              const float* __restrict__ fit, int32_t offer, uint32_t idx_base,
                                                              int32_t first_wins) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= B) return;
    float x[DD];
#pragma unroll
    for (int d = 0; d < DD; ++d) x[d] = desc[row * DD + d];
    const int32_t cell = qdx_index_cell<DD>(x, ix);
    cells[row] = cell;
    if (offer) qdx_offer(ws, K, rep_f, cell, fit[row], idx_base + (uint32_t)row, first_wins);
}

int qdx_fill_cvt_index(const qdx_cvt_index* in, QdxCvtIndex* out) {
    memset(out, 0, sizeof(*out));
    if (!in || in->dd == 0) return 0;
    if (in->dd < 1 || in->dd > QDX_INDEX_MAX_DIM || !in->start || !in->ids || !in->pts) return QDX_ERR_ARG;
    out->dd = in->dd;
    for (int d = 0; d < in->dd; ++d) {
        if (in->g[d] < 1 || !(in->h[d] > 0.0f)) return QDX_ERR_ARG;
        out->g[d] = in->g[d]; out->lo[d] = in->lo[d]; out->h[d] = in->h[d];
    }
    out->start = in->start; out->ids = in->ids; out->pts = in->pts;
    return 0;
}

extern "C" {

// Plan: bounding box, buckets per dimension (~2 centroids per bucket, bucket width kept >= 2e-3 * max|coordinate| so
// that the search's rounding margins hold), bucket count.  Returns QDX_ERR_UNSUPPORTED when an index does not apply
// (dd > 3, non-finite centroids, too few centroids, degenerate box): callers use the brute-force kernel.
int qdx_cvt_index_plan(const float* centroids_host, int64_t K, int32_t dd, qdx_cvt_index* plan, int64_t* n_buckets) {
    if (!centroids_host || !plan || !n_buckets || K <= 0 || dd < 1) return QDX_ERR_ARG;
    memset(plan, 0, sizeof(*plan));
    if (dd > QDX_INDEX_MAX_DIM || K < 64 || K >= (1ll << 31)) return QDX_ERR_UNSUPPORTED;
    double lo[3], hi[3], maxabs = 0.0;
    for (int d = 0; d < dd; ++d) { lo[d] = INFINITY; hi[d] = -INFINITY; }
    for (int64_t k = 0; k < K; ++k)
        for (int d = 0; d < dd; ++d) {
            const double v = centroids_host[k * dd + d];
            if (!std::isfinite(v)) return QDX_ERR_UNSUPPORTED;
            lo[d] = v < lo[d] ? v : lo[d]; hi[d] = v > hi[d] ? v : hi[d];
            maxabs = std::fabs(v) > maxabs ? std::fabs(v) : maxabs;
        }
    const double per_dim = std::pow((double)K / 2.0, 1.0 / dd);
    const double hmin = 2e-3 * (maxabs > 1e-30 ? maxabs : 1e-30);
    int64_t nb = 1;
    for (int d = 0; d < dd; ++d) {
        const double span = hi[d] - lo[d];
        int64_t g = (int64_t)std::llround(per_dim);
        if (g < 1) g = 1;
        if (!(span > 0.0)) g = 1;
        else if (span / g < hmin) g = (int64_t)std::floor(span / hmin);
        if (g < 1) g = 1;
        if (g > 2048) g = 2048;
        plan->g[d] = (int32_t)g;
        plan->lo[d] = (float)lo[d];
        if ((double)plan->lo[d] > lo[d]) plan->lo[d] = std::nextafterf(plan->lo[d], -INFINITY);
        float h = (float)((hi[d] - (double)plan->lo[d]) / g);
        // the last bucket must reach hi: grow h until lo + g*h >= hi in exact arithmetic on the float32 values
        while ((double)plan->lo[d] + (double)g * (double)h < hi[d]) h = std::nextafterf(h, INFINITY);
        if (!(h > 0.0f)) h = 1.0f;
        plan->h[d] = h;
        nb *= g;
    }
    if (nb < 4) return QDX_ERR_UNSUPPORTED;
    plan->dd = dd;
    *n_buckets = nb;
    return 0;
}

// Build (host): bucket of centroid c = floor((c - lo) / h) per dimension in double on the float32 plan values, clamped;
// counting sort by bucket keeps ascending centroid id inside a bucket.
int qdx_cvt_index_build(const float* centroids_host, int64_t K, const qdx_cvt_index* plan, int32_t* start_host, int32_t* ids_host,
                        float* pts_host) {
    if (!centroids_host || !plan || !start_host || !ids_host || !pts_host || K <= 0) return QDX_ERR_ARG;
    const int dd = plan->dd;
    if (dd < 1 || dd > QDX_INDEX_MAX_DIM) return QDX_ERR_ARG;
    int64_t nb = 1;
    for (int d = 0; d < dd; ++d) nb *= plan->g[d];
    std::vector<int64_t> bucket((size_t)K);
    std::vector<int32_t> count((size_t)nb + 1, 0);
    for (int64_t k = 0; k < K; ++k) {
        int64_t flat = 0, mul = 1;
        for (int d = 0; d < dd; ++d) {
            double t = std::floor(((double)centroids_host[k * dd + d] - (double)plan->lo[d]) / (double)plan->h[d]);
            int64_t b = t < 0.0 ? 0 : (t >= (double)plan->g[d] ? plan->g[d] - 1 : (int64_t)t);
            flat += b * mul; mul *= plan->g[d];
        }
        bucket[(size_t)k] = flat;
        ++count[(size_t)flat + 1];
    }
    for (int64_t b = 0; b < nb; ++b) count[(size_t)b + 1] += count[(size_t)b];
    for (int64_t b = 0; b <= nb; ++b) start_host[b] = count[(size_t)b];
    std::vector<int32_t> cursor(count.begin(), count.end() - 1);
    for (int64_t k = 0; k < K; ++k) {
        const int32_t pos = cursor[(size_t)bucket[(size_t)k]]++;
        ids_host[pos] = (int32_t)k;
        for (int d = 0; d < dd; ++d) pts_host[(int64_t)pos * dd + d] = centroids_host[k * dd + d];
    }
    return 0;
}

int qdx_cells_indexed(const float* desc, int64_t B, const qdx_cvt_index* index, int64_t K, int32_t* out_cells, void* ws,
                      const float* rep_fitness, const float* fitness, int32_t offer, uint32_t idx_base, int32_t first_wins,
                      void* stream) {
    if (!desc || !index || !out_cells || B < 0 || K <= 0) return QDX_ERR_ARG;
    if (offer && (!ws || !rep_fitness || !fitness)) return QDX_ERR_ARG;
    if ((uint64_t)idx_base + (uint64_t)B > 0x7FFFFFFFull) return QDX_ERR_ARG;
    if (B == 0) return 0;
    QdxCvtIndex ix;
    int rc = qdx_fill_cvt_index(index, &ix);
    if (rc) return rc;
    if (ix.dd == 0) return QDX_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 gr((unsigned)((B + 255) / 256));
    switch (ix.dd) {
        case 1: qdx_cells_index_kernel<1><<<gr, 256, 0, st>>>(desc, B, ix, K, out_cells, ws, rep_fitness, fitness, offer, idx_base, first_wins); break;
        case 2: qdx_cells_index_kernel<2><<<gr, 256, 0, st>>>(desc, B, ix, K, out_cells, ws, rep_fitness, fitness, offer, idx_base, first_wins); break;
        default: qdx_cells_index_kernel<3><<<gr, 256, 0, st>>>(desc, B, ix, K, out_cells, ws, rep_fitness, fitness, offer, idx_base, first_wins); break;
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

}  // extern "C"
