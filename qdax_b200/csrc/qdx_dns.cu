// Dominated Novelty Search competition on B200 (sm_100a).
//
// Reference semantics (under /root/reference): qdax/core/containers/dns_repertoire.py:22-76
// (_novelty_and_dominated_novelty) and :94-165 (DominatedNoveltyRepertoire.add).  The reference materialises
// several dense (N, N) arrays; here every query row keeps a k-entry sorted list in registers while candidate
// tiles (fitness, descriptor) stream through shared memory, so the N x N pair space is never stored.
//   dn_i   = mean of the k smallest sqrt(sum_d (x_id - x_jd)^2) over j != i, both valid, f_i <= f_j
//            (fewer than k such j: mean over those; none: 0/0 = NaN)
//   meta_i = valid_i ? dn_i : -inf
//   order  = argsort(meta)[::-1][:P]   (stable ascending, NaN last, reversed: NaN first, then descending,
//            higher index first among equals)  -> realised as a descending rank on the unique 64-bit key
//            (order_key(meta) << 32 | index).
// Ranking on the squared distance and taking sqrt of the k survivors is exact: sqrt is monotone, so the
// multiset of the k smallest distances is unchanged.
#include "qdx_common.cuh"
#include "../../include/qdx.h"

#define QDX_CHECK_LAUNCH() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return (int)e_; } while (0)

struct QdxCat {       // candidates = population rows followed by batch rows (the reference concatenates)
    const float* pf; const float* pd; const float* bf; const float* bd; int64_t P; int64_t N; int32_t Dd;
    __device__ __forceinline__ float fit(int64_t i) const { return i < P ? pf[i] : bf[i - P]; }
    __device__ __forceinline__ float desc(int64_t i, int d) const { return i < P ? pd[i * Dd + d] : bd[(i - P) * Dd + d]; }
};

template <int KMAX, int DD>
__global__ void __launch_bounds__(256) qdx_dns_novelty_kernel(QdxCat c, int32_t k, float* __restrict__ meta) {
    constexpr int TILE = 1024;
    __shared__ float s_f[TILE];
    __shared__ float s_d[TILE * DD];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid_i = i < c.N;
    const float fi = valid_i ? c.fit(i) : -INFINITY;
    float xi[DD];
#pragma unroll
    for (int d = 0; d < DD; ++d) xi[d] = valid_i ? c.desc(i, d) : 0.0f;
    float top[KMAX];
#pragma unroll
    for (int t = 0; t < KMAX; ++t) top[t] = INFINITY;
    int cnt = 0;
    for (int64_t j0 = 0; j0 < c.N; j0 += TILE) {
        const int n = (c.N - j0) < TILE ? (int)(c.N - j0) : TILE;
        __syncthreads();
        for (int t = threadIdx.x; t < n; t += blockDim.x) {
            s_f[t] = c.fit(j0 + t);
#pragma unroll
            for (int d = 0; d < DD; ++d) s_d[t * DD + d] = c.desc(j0 + t, d);
        }
        __syncthreads();
        if (!valid_i || fi == -INFINITY) continue;
        for (int t = 0; t < n; ++t) {
            const float fj = s_f[t];
            if (!(fi <= fj) || fj == -INFINITY || j0 + t == i) continue;       // dns_repertoire.py:44-49
            float acc;
#pragma unroll
            for (int d = 0; d < DD; ++d) { float df = xi[d] - s_d[t * DD + d]; float s = df * df; acc = d ? acc + s : s; }
            if (cnt < k) ++cnt;
            if (acc < top[KMAX - 1]) {          // sorted insertion, ascending, compile-time indices
                float v = acc;
#pragma unroll
                for (int u = 0; u < KMAX; ++u) { const float lo = fminf(top[u], v); v = fmaxf(top[u], v); top[u] = lo; }
            }
        }
    }
    if (!valid_i) return;
    float out;
    if (fi == -INFINITY) out = -INFINITY;                                           // :144-145
    else {
        float tot = 0.0f;
#pragma unroll
        for (int u = 0; u < KMAX; ++u) if (u < cnt) tot = tot + __fsqrt_rn(top[u]);   // :52, :70-74 (top-k order)
        out = __fdiv_rn(tot, (float)cnt);                                           // 0/0 -> NaN
    }
    meta[i] = out;
}

// generic descriptor dimension: descriptor of the query row re-read from global (L1-resident), Dd <= 64
template <int KMAX>
__global__ void __launch_bounds__(128) qdx_dns_novelty_generic_kernel(QdxCat c, int32_t k, float* __restrict__ meta) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N) return;
    const float fi = c.fit(i);
    if (fi == -INFINITY) { meta[i] = -INFINITY; return; }
    float top[KMAX];
#pragma unroll
    for (int t = 0; t < KMAX; ++t) top[t] = INFINITY;
    int cnt = 0;
    for (int64_t j = 0; j < c.N; ++j) {
        const float fj = c.fit(j);
        if (!(fi <= fj) || fj == -INFINITY || j == i) continue;
        float acc = 0.0f;
        for (int d = 0; d < c.Dd; ++d) { float df = c.desc(i, d) - c.desc(j, d); float s = df * df; acc = d ? acc + s : s; }
        if (cnt < k) ++cnt;
        if (acc < top[KMAX - 1]) {
            float v = acc;
#pragma unroll
            for (int u = 0; u < KMAX; ++u) { const float lo = fminf(top[u], v); v = fmaxf(top[u], v); top[u] = lo; }
        }
    }
    float tot = 0.0f;
#pragma unroll
    for (int u = 0; u < KMAX; ++u) if (u < cnt) tot = tot + __fsqrt_rn(top[u]);
    meta[i] = __fdiv_rn(tot, (float)cnt);
}

// descending rank of the unique key (order_key(meta) << 32 | i); survivors[rank] = i for rank < P
__global__ void __launch_bounds__(256) qdx_dns_rank_kernel(const float* __restrict__ meta, int64_t N, int64_t P,
                                                           int32_t* __restrict__ survivors) {
    constexpr int TILE = 2048;
    __shared__ unsigned long long s_key[TILE];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long ki = i < N ? (((unsigned long long)qdx_order_key(meta[i]) << 32) | (unsigned long long)(uint32_t)i) : ~0ull;
    int64_t rank = 0;
    for (int64_t j0 = 0; j0 < N; j0 += TILE) {
        const int n = (N - j0) < TILE ? (int)(N - j0) : TILE;
        __syncthreads();
        for (int t = threadIdx.x; t < n; t += blockDim.x)
            s_key[t] = ((unsigned long long)qdx_order_key(meta[j0 + t]) << 32) | (unsigned long long)(uint32_t)(j0 + t);
        __syncthreads();
        int r = 0;
#pragma unroll 8
        for (int t = 0; t < n; ++t) r += (s_key[t] > ki);
        rank += r;
    }
    if (i < N && rank < P) survivors[rank] = (int32_t)i;
}

// new population rows = candidates[survivors] (dns_repertoire.py:151-158)
__global__ void __launch_bounds__(256) qdx_dns_gather_kernel(const float* __restrict__ pg, const float* __restrict__ bg,
                                                             QdxCat c, int32_t D, const int32_t* __restrict__ survivors,
                                                             float* __restrict__ out_g, float* __restrict__ out_f,
                                                             float* __restrict__ out_d) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= c.P) return;
    const int64_t i = survivors[row];
    const float* s = i < c.P ? pg + i * D : bg + (i - c.P) * D;
    float* o = out_g + row * D;
    if ((D & 3) == 0) {
        for (int q = lane; q < (D >> 2); q += 32) reinterpret_cast<float4*>(o)[q] = __ldg(reinterpret_cast<const float4*>(s) + q);
    } else {
        for (int d = lane; d < D; d += 32) o[d] = s[d];
    }
    for (int d = lane; d < c.Dd; d += 32) out_d[row * c.Dd + d] = c.desc(i, d);
    if (lane == 0) out_f[row] = c.fit(i);
}

extern "C" {

int qdx_dns_add(const float* pop_genotypes, const float* pop_fitness, const float* pop_desc, int64_t P,
                const float* batch_genotypes, const float* batch_fitness, const float* batch_desc, int64_t B, int64_t D,
                int32_t desc_dim, int32_t k, float* out_genotypes, float* out_fitness, float* out_desc, float* meta_scratch,
                int32_t* survivors_scratch, void* stream) {
    if (!pop_genotypes || !pop_fitness || !pop_desc || !out_genotypes || !out_fitness || !out_desc || !meta_scratch || !survivors_scratch)
        return QDX_ERR_ARG;
    if (P <= 0 || B < 0 || D <= 0 || desc_dim < 1 || k < 1 || k > 32) return QDX_ERR_ARG;
    if (B > 0 && (!batch_genotypes || !batch_fitness || !batch_desc)) return QDX_ERR_ARG;
    if (out_genotypes == pop_genotypes || out_fitness == pop_fitness || out_desc == pop_desc) return QDX_ERR_ARG;  // not in place
    cudaStream_t st = (cudaStream_t)stream;
    QdxCat c{pop_fitness, pop_desc, batch_fitness, batch_desc, P, P + B, desc_dim};
    const int64_t N = P + B;
    const unsigned g256 = (unsigned)((N + 255) / 256);
#define QDX_DNS_LAUNCH(KM)                                                                              \
    do {                                                                                                \
        if (desc_dim == 1) qdx_dns_novelty_kernel<KM, 1><<<g256, 256, 0, st>>>(c, k, meta_scratch);       \
        else if (desc_dim == 2) qdx_dns_novelty_kernel<KM, 2><<<g256, 256, 0, st>>>(c, k, meta_scratch);  \
        else if (desc_dim == 3) qdx_dns_novelty_kernel<KM, 3><<<g256, 256, 0, st>>>(c, k, meta_scratch);  \
        else if (desc_dim == 4) qdx_dns_novelty_kernel<KM, 4><<<g256, 256, 0, st>>>(c, k, meta_scratch);  \
        else qdx_dns_novelty_generic_kernel<KM><<<(unsigned)((N + 127) / 128), 128, 0, st>>>(c, k, meta_scratch); \
    } while (0)
    if (k <= 4) QDX_DNS_LAUNCH(4);
    else if (k <= 16) QDX_DNS_LAUNCH(16);
    else QDX_DNS_LAUNCH(32);
#undef QDX_DNS_LAUNCH
    QDX_CHECK_LAUNCH();
    qdx_dns_rank_kernel<<<g256, 256, 0, st>>>(meta_scratch, N, P, survivors_scratch);
    QDX_CHECK_LAUNCH();
    qdx_dns_gather_kernel<<<(unsigned)((P * 32 + 255) / 256), 256, 0, st>>>(pop_genotypes, batch_genotypes, c, (int32_t)D,
                                                                           survivors_scratch, out_genotypes, out_fitness, out_desc);
    QDX_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
