// Sibling insertion rules that reuse the cell assignment (c) and the packed-key election + streaming commit (d) of the
// MAP-Elites path.  Reference files under /root/reference:
//   MELSRepertoire.add          qdax/core/containers/mels_repertoire.py:89-230  (_dispersion :26-48, _mode :51-57)
//   compute_cvt_centroids       qdax/core/containers/mapelites_repertoire.py:30-72   (Lloyd iterations on the GPU, rank 4)
#include "qdx_common.cuh"
#include "../../include/qdx.h"

#define QDX_CHECK_LAUNCH_C()                             \
    do {                                                 \
        cudaError_t e_ = cudaGetLastError();             \
        if (e_ != cudaSuccess) return (int)e_;           \
    } while (0)

namespace {

// MAP-Elites Low-Spread: every individual was evaluated S times.  Thread = individual:
//   cell     = most frequent of its S cells, the smallest such cell on ties (jnp.unique is sorted, argmax takes the first)
//   spread   = sum_{i<j} ||d_i - d_j||_2 / (S (S-1) / 2), 0 when S == 1                       (:149-158, :26-48)
//   fitness  = mean of the S fitnesses                                                         (:169)
//   desc     = centroid of `cell`                                                              (:162-164)
// and the candidate is offered to its cell iff fitness > current fitness AND spread <= current spread (:181-187).  The
// reference then scatters EVERY passing candidate (no segment_max): which one survives a collision is unspecified
// (its own test accepts either, tests/core_test/containers_test/mels_repertoire_test.py:103-105); here the first / last
// offspring index wins, elected by the same 64-bit atomicMax with a constant fitness field.
// Sums are sequential, left to right (QDX-F32 spec).
__global__ void __launch_bounds__(128) qdx_mels_kernel(const int32_t* __restrict__ cells_all, const float* __restrict__ desc_all,
                                                       const float* __restrict__ fit_all, int64_t B, int32_t S, int32_t Dd,
                                                       const float* __restrict__ centroids, int64_t K, void* ws,
                                                       const float* __restrict__ rep_f, const float* __restrict__ rep_spread,
                                                       int32_t first_wins, int32_t* __restrict__ out_cells, float* __restrict__ out_f,
                                                       float* __restrict__ out_spread, float* __restrict__ out_desc) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int32_t* c = cells_all + b * S;
    int32_t mode = c[0]; int best = 0;
    for (int i = 0; i < S; ++i) {
        const int32_t v = c[i];
        int n = 0;
        for (int j = 0; j < S; ++j) n += (c[j] == v);
        if (n > best || (n == best && v < mode)) { best = n; mode = v; }
    }
    const float* d = desc_all + b * (int64_t)S * Dd;
    float spread = 0.0f;
    if (S > 1) {
        float sum = 0.0f;
        for (int i = 0; i < S; ++i)
            for (int j = i + 1; j < S; ++j) {
                float acc = 0.0f;
                for (int k = 0; k < Dd; ++k) { const float t = d[i * Dd + k] - d[j * Dd + k]; acc = acc + t * t; }
                sum = sum + __fsqrt_rn(acc);
            }
        spread = __fdiv_rn(sum, (float)((double)S * (double)(S - 1) / 2.0));
    }
    float fs = 0.0f;
    for (int i = 0; i < S; ++i) fs = fs + fit_all[b * S + i];
    const float f = __fdiv_rn(fs, (float)S);
    out_cells[b] = mode; out_f[b] = f; out_spread[b] = spread;
    const bool in_range = mode >= 0 && mode < K;
    for (int k = 0; k < Dd; ++k) out_desc[b * Dd + k] = in_range ? centroids[(int64_t)mode * Dd + k] : 0.0f;
    if (!in_range) { qdx_set_error(ws, QDX_ERR_BAD_CELL); return; }
    if (f > rep_f[mode] && spread <= rep_spread[mode])
        atomicMax(qdx_ws_keytab(ws, K) + mode, qdx_pack_key(0.0f, (uint32_t)b, first_wins));
}

// dst[c, :] = src[source_of_cell[c], :] for every cell a commit changed (source_of_cell = qdx_commit's added_cells, -1 =
// unchanged): the extra per-cell arrays that travel with an insertion (MELS spreads, extra_scores)
__global__ void __launch_bounds__(256) qdx_scatter_by_source_kernel(const int32_t* __restrict__ source_of_cell, const float* __restrict__ src,
                                                                    int64_t K, int64_t W, float* __restrict__ dst) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= K * W) return;
    const int64_t c = e / W, w = e - c * W;
    const int32_t i = source_of_cell[c];
    if (i >= 0) dst[e] = src[(int64_t)i * W + w];
}

// ---- Lloyd iteration of compute_cvt_centroids (qdax/core/containers/mapelites_repertoire.py:30-72; the reference calls
// scikit-learn's KMeans on the host).  Assignment = the cell-assignment kernels (c); the update below is ORDER-FREE and
// therefore bit-reproducible: samples lie in [0, 1) (the reference clusters in the unit cube and rescales afterwards, :55-72),
// each coordinate is quantised to 32 fractional bits and summed per (cluster, dimension) with 64-bit integer atomics; the new
// centroid is (double)sum / count * 2^-32 rounded to float32; an empty cluster keeps its centroid.
__global__ void __launch_bounds__(256) qdx_kmeans_accumulate_kernel(const float* __restrict__ x, const int32_t* __restrict__ cells,
                                                                    const int32_t* __restrict__ prev_cells, int64_t N, int32_t Dd, int64_t K,
                                                                    unsigned long long* __restrict__ acc, int32_t* __restrict__ count,
                                                                    int32_t* __restrict__ changed) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * Dd) return;
    const int64_t i = e / Dd;
    const int32_t d = (int32_t)(e - i * Dd);
    const int32_t c = cells[i];
    if (c < 0 || c >= K) return;
    float v = x[e];
    v = v < 0.0f ? 0.0f : v;                                           // also maps NaN to 0
    double q = (double)v * 4294967296.0;
    if (q > 4294967295.0) q = 4294967295.0;
    atomicAdd(acc + (int64_t)c * Dd + d, (unsigned long long)q);
    if (d == 0) {
        atomicAdd(count + c, 1);
        if (prev_cells && prev_cells[i] != c) atomicAdd(changed, 1);
    }
}
__global__ void __launch_bounds__(256) qdx_kmeans_update_kernel(const unsigned long long* __restrict__ acc, const int32_t* __restrict__ count,
                                                                const float* __restrict__ old_c, int64_t K, int32_t Dd, float* __restrict__ new_c) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= K * Dd) return;
    const int32_t n = count[e / Dd];
    new_c[e] = n > 0 ? (float)(((double)acc[e] / (double)n) * (1.0 / 4294967296.0)) : old_c[e];
}

}  // namespace

extern "C" int qdx_kmeans_accumulate(const float* x, const int32_t* cells, const int32_t* prev_cells, int64_t N, int32_t desc_dim, int64_t K,
                                     unsigned long long* acc, int32_t* count, int32_t* changed, void* stream) {
    if (!x || !cells || !acc || !count || !changed || N < 0 || desc_dim < 1 || K <= 0) return QDX_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(unsigned long long) * (size_t)K * desc_dim, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(count, 0, sizeof(int32_t) * (size_t)K, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(changed, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return (int)e;
    if (N == 0) return 0;
    const int64_t n = N * desc_dim;
    qdx_kmeans_accumulate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, cells, prev_cells, N, desc_dim, K, acc, count, changed);
    QDX_CHECK_LAUNCH_C();
    return 0;
}

extern "C" int qdx_kmeans_update(const unsigned long long* acc, const int32_t* count, const float* old_centroids, int64_t K, int32_t desc_dim,
                                 float* new_centroids, void* stream) {
    if (!acc || !count || !old_centroids || !new_centroids || K <= 0 || desc_dim < 1) return QDX_ERR_ARG;
    const int64_t n = K * desc_dim;
    qdx_kmeans_update_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(acc, count, old_centroids, K, desc_dim, new_centroids);
    QDX_CHECK_LAUNCH_C();
    return 0;
}

extern "C" int qdx_mels_offer(const int32_t* cells_all, const float* desc_all, const float* fit_all, int64_t B, int32_t S, int32_t desc_dim,
                              const float* centroids, int64_t K, void* ws, const float* rep_fitness, const float* rep_spread,
                              int32_t first_wins, int32_t* out_cells, float* out_fitness, float* out_spread, float* out_desc,
                              void* stream) {
    if (!cells_all || !desc_all || !fit_all || !centroids || !ws || !rep_fitness || !rep_spread || !out_cells || !out_fitness ||
        !out_spread || !out_desc)
        return QDX_ERR_ARG;
    if (B < 0 || S < 1 || desc_dim < 1 || K <= 0 || B >= (1ll << 31)) return QDX_ERR_ARG;
    if (B == 0) return 0;
    qdx_mels_kernel<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(cells_all, desc_all, fit_all, B, S, desc_dim, centroids, K, ws,
                                                                                  rep_fitness, rep_spread, first_wins, out_cells, out_fitness,
                                                                                  out_spread, out_desc);
    QDX_CHECK_LAUNCH_C();
    return 0;
}

extern "C" int qdx_scatter_rows_by_source(const int32_t* source_of_cell, const float* src, int64_t K, int64_t W, float* dst, void* stream) {
    if (!source_of_cell || !src || !dst || K < 0 || W < 1) return QDX_ERR_ARG;
    if (K == 0) return 0;
    const int64_t n = K * W;
    qdx_scatter_by_source_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(source_of_cell, src, K, W, dst);
    QDX_CHECK_LAUNCH_C();
    return 0;
}
