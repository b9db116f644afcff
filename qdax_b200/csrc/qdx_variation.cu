// polynomial_mutation / polynomial_crossover on B200 (sm_100a).
//
// Reference semantics (under /root/reference): qdax/core/emitters/mutation_operators.py:12-117 and :120-172.
// Per row r (key_r = split(key, B)[r]):
//   mutation : key, sub = split(key_r); positions = permutation(sub, D)[:n]   (jax.random.choice(replace=False));
//              key, sub = split(key); rand = uniform(sub, (n,)); polynomial delta on the n selected genes; clip.
//   crossover: indices = randint(key_r, (n,), 0, D) (with replacement); x1[indices] <- x2[indices].
// jax.random.permutation = `rounds` passes of (k, s = split(k); stable sort of the array by random_bits(s, (D,)));
// rounds = ceil(3 ln D / ln(2^32 - 1)) (1 for D <= 1625).  One warp owns a row; the stable sort is realised as a
// rank by counting over the D 32-bit keys held in shared memory (rank_i = #{j : key_j < key_i or (== and j < i)}),
// which needs no data-dependent control flow and is exact for duplicate keys.
#include "qdx_common.cuh"
#include "../../include/qdx.h"

#define QDX_CHECK_LAUNCH() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return (int)e_; } while (0)

__global__ void __launch_bounds__(128) qdx_polymut_kernel(const float* __restrict__ x, int64_t B, int32_t D, QdxKey key, int32_t n,
                                                          int32_t rounds, float ep1, float mutpow, float minv, float maxv,
                                                          float* __restrict__ out) {
    extern __shared__ uint32_t s_mut[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * nwarps + warp;
    if (row >= B) return;
    uint32_t* keys = s_mut + (size_t)warp * 3 * D;
    int32_t* perm = (int32_t*)(keys + D);
    int32_t* perm2 = perm + D;
    const QdxKey kr = qdx_split(key, (uint64_t)row);                 // :107
    QdxKey pk = qdx_split(kr, 1);                                    // :41  key, subkey = split(key) -> choice(subkey)
    const QdxKey k1 = qdx_split(kr, 0);
    for (int i = lane; i < D; i += 32) perm[i] = i;
    for (int r = 0; r < rounds; ++r) {
        const QdxKey sub = qdx_split(pk, 1); pk = qdx_split(pk, 0);
        __syncwarp();
        for (int i = lane; i < D; i += 32) keys[i] = qdx_bits32(sub, (uint64_t)i);
        __syncwarp();
        for (int i = lane; i < D; i += 32) {
            const uint32_t ki = keys[i];
            int rank = 0;
            for (int j = 0; j < D; ++j) { const uint32_t kj = keys[j]; rank += (kj < ki) || (kj == ki && j < i); }
            perm2[rank] = perm[i];
        }
        __syncwarp();
        int32_t* t = perm; perm = perm2; perm2 = t;
    }
    const QdxKey sub2 = qdx_split(k1, 1);                            // :53
    const float rng = maxv - minv;
    const float* xr = x + row * D; float* o = out + row * D;
    for (int d = lane; d < D; d += 32) o[d] = qdx_min_nanprop(qdx_max_nanprop(xr[d], minv), maxv);   // :75 on untouched genes
    __syncwarp();
    for (int j = lane; j < n; j += 32) {
        const int32_t pos = perm[j];                                 // :42-45
        const float mx = xr[pos];
        const float d1 = __fdiv_rn(mx - minv, rng), d2 = __fdiv_rn(maxv - mx, rng);    // :49-50
        const float r = qdx_unit_float(qdx_bits32(sub2, (uint64_t)j));                 // :54-60
        float v1 = 2.0f * r + qdx_powf(d1, ep1) * (1.0f - 2.0f * r);                   // :62
        float v2 = 2.0f * (1.0f - r) + 2.0f * (qdx_powf(d2, ep1) * (r - 0.5f));        // :63
        v1 = qdx_powf(v1, mutpow) - 1.0f;                                              // :64
        v2 = 1.0f - qdx_powf(v2, mutpow);                                              // :65
        const float dq = r < 0.5f ? v1 : v2;                                           // :67-69
        o[pos] = qdx_min_nanprop(qdx_max_nanprop(mx + dq * rng, minv), maxv);          // :72, :75
    }
}

// jax.random.randint(key, (n,), 0, span): two 32-bit draws per element from split(key), combined modulo span in uint32
__device__ __forceinline__ uint32_t qdx_randint(QdxKey k1, QdxKey k2, uint64_t j, uint32_t span) {
    const uint32_t hi = qdx_bits32(k1, j), lo = qdx_bits32(k2, j);
    uint32_t mult = 65536u % span; mult = (mult * mult) % span;
    return ((hi % span) * mult + (lo % span)) % span;
}

__global__ void __launch_bounds__(128) qdx_polycross_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int64_t B,
                                                            int32_t D, QdxKey key, int32_t n, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= B) return;
    const QdxKey kr = qdx_split(key, (uint64_t)row);                 // :164
    const QdxKey r1 = qdx_split(kr, 0), r2 = qdx_split(kr, 1);
    const float* a = x1 + row * D; const float* b = x2 + row * D; float* o = out + row * D;
    for (int d = lane; d < D; d += 32) o[d] = a[d];
    __syncwarp();
    for (int j = lane; j < n; j += 32) {
        const uint32_t idx = qdx_randint(r1, r2, (uint64_t)j, (uint32_t)D);            // :132
        o[idx] = b[idx];                                                               // :133 (duplicates write the same value)
    }
}

extern "C" {

int qdx_polynomial_mutation(const float* x, int64_t B, int64_t D, uint32_t k0, uint32_t k1, int32_t n_mutate, float eta_plus_1,
                            float mutpow, float minval, float maxval, float* out, void* stream) {
    if (!x || !out || B < 0 || D <= 0 || n_mutate < 0 || n_mutate > D || x == out) return QDX_ERR_ARG;
    if (D > 4096) return QDX_ERR_UNSUPPORTED;
    if (B == 0) return 0;
    const int rounds = D <= 1 ? 1 : (int)ceil(3.0 * log((double)D) / log(4294967295.0));
    const int warps = D <= 1024 ? 4 : 1;
    const size_t smem = (size_t)warps * 3 * D * sizeof(uint32_t);
    cudaError_t e = cudaFuncSetAttribute(qdx_polymut_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    qdx_polymut_kernel<<<(unsigned)((B + warps - 1) / warps), warps * 32, smem, (cudaStream_t)stream>>>(
        x, B, (int32_t)D, QdxKey{k0, k1}, n_mutate, rounds, eta_plus_1, mutpow, minval, maxval, out);
    QDX_CHECK_LAUNCH();
    return 0;
}

int qdx_polynomial_crossover(const float* x1, const float* x2, int64_t B, int64_t D, uint32_t k0, uint32_t k1, int32_t n_change,
                             float* out, void* stream) {
    if (!x1 || !x2 || !out || B < 0 || D <= 0 || D > 65535 || n_change < 0 || out == x1 || out == x2) return QDX_ERR_ARG;
    if (B == 0) return 0;
    qdx_polycross_kernel<<<(unsigned)((B * 32 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x1, x2, B, (int32_t)D, QdxKey{k0, k1}, n_change, out);
    QDX_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
