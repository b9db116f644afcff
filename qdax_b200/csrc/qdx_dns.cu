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
#include <cstdlib>
#include "qdx_common.cuh"
#include "../../include/qdx.h"

#define QDX_CHECK_LAUNCH() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return (int)e_; } while (0)

struct QdxCat {       // candidates = population rows followed by batch rows (the reference concatenates)
    const float* pf; const float* pd; const float* bf; const float* bd; int64_t P; int64_t N; int32_t Dd;
    __device__ __forceinline__ float fit(int64_t i) const { return i < P ? pf[i] : bf[i - P]; }
    __device__ __forceinline__ float desc(int64_t i, int d) const { return i < P ? pd[i * Dd + d] : bd[(i - P) * Dd + d]; }
};

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

// ---- descending rank of N unique 64-bit keys (order_key(value) << 32 | index):  out[rank] = index for rank < limit.
// Replaces the argsort of the reference (dns_repertoire.py:148) -- and the N^2 counting pass this file started with -- by a
// bucketed count: keys are spread over NB buckets linearly between the smallest and largest order key (NaN and -inf get
// buckets of their own), rank = (keys in higher buckets) + (keys in my bucket that are greater), the second term by a scan
// of the bucket's members.  Exact for any input; degenerates to the quadratic count only when all values are (nearly) equal.
constexpr int QDX_RANK_NB = 8192;
struct QdxRankWs {
    uint32_t lo, hi;                          // range of the ordinary order keys
    uint32_t hist[QDX_RANK_NB + 2];           // bucket 0 = -inf, 1..NB ordinary, NB+1 = NaN
    uint32_t start[QDX_RANK_NB + 3];          // members of bucket b live at grouped[start[b] .. start[b+1])
    uint32_t cursor[QDX_RANK_NB + 2];
};
QDX_DEV int qdx_rank_bucket(uint32_t ok, uint32_t lo, uint32_t hi) {
    if (ok == 0xFFFFFFFFu) return QDX_RANK_NB + 1;
    if (ok == 0x007FFFFFu) return 0;                                   // order key of -inf
    const unsigned long long span = (unsigned long long)(hi - lo) + 1ull;
    return 1 + (int)(((unsigned long long)(ok - lo) * (unsigned long long)QDX_RANK_NB) / span);
}
__global__ void __launch_bounds__(256) qdx_rank_keys_kernel(const float* __restrict__ val, int64_t N, unsigned long long* __restrict__ keys,
                                                            QdxRankWs* ws) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t ok = 0u; bool ordinary = false;
    if (i < N) {
        ok = qdx_order_key(val[i]);
        keys[i] = ((unsigned long long)ok << 32) | (unsigned long long)(uint32_t)i;
        ordinary = ok != 0xFFFFFFFFu && ok != 0x007FFFFFu;
    }
    uint32_t lo = ordinary ? ok : 0xFFFFFFFFu, hi = ordinary ? ok : 0u;
    for (int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if ((threadIdx.x & 31) == 0 && lo <= hi) { atomicMin(&ws->lo, lo); atomicMax(&ws->hi, hi); }
}
__global__ void __launch_bounds__(256) qdx_rank_hist_kernel(const unsigned long long* __restrict__ keys, int64_t N, QdxRankWs* ws) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const uint32_t lo = ws->lo, hi = ws->hi >= ws->lo ? ws->hi : ws->lo;
    atomicAdd(&ws->hist[qdx_rank_bucket((uint32_t)(keys[i] >> 32), lo, hi)], 1u);
}
__global__ void __launch_bounds__(1024) qdx_rank_scan_kernel(QdxRankWs* ws) {      // one CTA: start[] ascending by bucket
    __shared__ uint32_t s_part[1024];
    constexpr int TOT = QDX_RANK_NB + 2, PER = (TOT + 1023) / 1024;
    uint32_t loc[PER]; uint32_t sum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) { const int b = threadIdx.x * PER + j; loc[j] = b < TOT ? ws->hist[b] : 0u; sum += loc[j]; }
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const uint32_t v = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0u;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = s_part[threadIdx.x] - sum;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const int b = threadIdx.x * PER + j;
        if (b < TOT) { ws->start[b] = run; ws->cursor[b] = run; }
        run += loc[j];
    }
    if (threadIdx.x == 1023) ws->start[TOT] = s_part[1023];
}
__global__ void __launch_bounds__(256) qdx_rank_group_kernel(const unsigned long long* __restrict__ keys, int64_t N, QdxRankWs* ws,
                                                             unsigned long long* __restrict__ grouped) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const uint32_t lo = ws->lo, hi = ws->hi >= ws->lo ? ws->hi : ws->lo;
    const unsigned long long k = keys[i];
    grouped[atomicAdd(&ws->cursor[qdx_rank_bucket((uint32_t)(k >> 32), lo, hi)], 1u)] = k;
}
__global__ void __launch_bounds__(256) qdx_rank_out_kernel(const unsigned long long* __restrict__ keys, int64_t N, int64_t limit,
                                                           const QdxRankWs* __restrict__ ws, const unsigned long long* __restrict__ grouped,
                                                           int32_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const uint32_t lo = ws->lo, hi = ws->hi >= ws->lo ? ws->hi : ws->lo;
    const unsigned long long k = keys[i];
    const int b = qdx_rank_bucket((uint32_t)(k >> 32), lo, hi);
    const uint32_t s = ws->start[b], e = ws->start[b + 1];
    int64_t rank = (int64_t)N - (int64_t)e;                               // every key in a higher bucket is greater
    for (uint32_t t = s; t < e; ++t) rank += (grouped[t] > k);
    if (rank < limit) out[rank] = (int32_t)i;
}

// out[rank] = index, rank = descending position of (value, index); scratch is allocated stream-ordered
static int qdx_rank_desc(const float* val, int64_t N, int64_t limit, int32_t* out, cudaStream_t st) {
    unsigned long long* keys = nullptr; unsigned long long* grouped = nullptr; QdxRankWs* ws = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&keys, sizeof(unsigned long long) * (size_t)N * 2 + sizeof(QdxRankWs), st);
    if (e != cudaSuccess) return (int)e;
    grouped = keys + N;
    ws = (QdxRankWs*)(grouped + N);
    e = cudaMemsetAsync(ws, 0, sizeof(QdxRankWs), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(&ws->lo, 0xFF, sizeof(uint32_t), st);
    const unsigned g = (unsigned)((N + 255) / 256);
    if (e == cudaSuccess) {
        qdx_rank_keys_kernel<<<g, 256, 0, st>>>(val, N, keys, ws);
        qdx_rank_hist_kernel<<<g, 256, 0, st>>>(keys, N, ws);
        qdx_rank_scan_kernel<<<1, 1024, 0, st>>>(ws);
        qdx_rank_group_kernel<<<g, 256, 0, st>>>(keys, N, ws, grouped);
        qdx_rank_out_kernel<<<g, 256, 0, st>>>(keys, N, limit, ws, grouped, out);
        e = cudaGetLastError();
    }
    cudaError_t e2 = cudaFreeAsync(keys, st);
    return (int)(e != cudaSuccess ? e : e2);
}

// ---- k-NN competition over FITNESS-SORTED candidates.  Row i competes only against j with f_i <= f_j: with the
// candidates sorted by descending fitness those are a PREFIX of the array, so the pair space is a triangle -- half the
// work of the dense scan -- and a CTA of 256 consecutive sorted rows shares one bound (the end of the ties of its last
// row).  The order in which candidates are visited does not matter: the result is the multiset of the k smallest
// distances.  Inner loop: one LDS.128 per candidate (fitness + descriptor packed), the reference's subtract / square /
// left-to-right sum, one fused predicate (closer than the current k-th; AND fitter only in the few tiles where the sorted
// order does not already imply it), one rare-path branch per 8 candidates;
// the self pair is excluded only in the one tile that contains the CTA's own rows; the number of fitter neighbours is not
// counted pair by pair but read off the sorted order (end of the row's ties - NaN rows - itself).  Heavy CTAs (long
// prefixes) are scheduled first.
template <int KMAX, int DD, bool SELF, bool CHECK>
__device__ __forceinline__ void qdx_dns_scan_tile(const float4* __restrict__ s_c, const float* __restrict__ s_d3, int n8, int self_t,
                                                  float fi, const float (&xi)[DD], float (&top)[KMAX]) {
    for (int t = 0; t < n8; t += 8) {
        float v[8]; bool hit[8]; bool any = false;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            float4 c;                                                           // (d0, d1, fitness, d2): d0 / d1 are one aligned LDS.64
            if (CHECK || DD >= 3) c = s_c[t + u];
            else { const float2 d01 = *reinterpret_cast<const float2*>(&s_c[t + u]); c = make_float4(d01.x, d01.y, 0.0f, 0.0f); }
            float acc, df;
            if (DD == 2 && QDX_PACKED_F32) {
                // both coordinates at once (FADD2 / FMUL2: two IEEE results per instruction, the kernel is issue-bound); the sum of
                // the two rounded squares stays a scalar add -- ptxas would contract a packed product feeding a packed sum
                const QdxF2 dd = qdx_f2(xi[0], xi[1]) - qdx_f2(c.x, c.y);
                float s0, s1;
                qdx_f2_get(dd * dd, s0, s1);
                acc = s0 + s1;
            } else {
                df = xi[0] - c.x;
                acc = df * df;
                if (DD >= 2) { df = xi[1] - c.y; acc = acc + df * df; }
            }
            if (DD >= 3) { df = xi[2] - c.w; acc = acc + df * df; }
            if (DD >= 4) { df = xi[3] - s_d3[t + u]; acc = acc + df * df; }
            v[u] = acc;
            hit[u] = acc < top[KMAX - 1];
            if (CHECK) hit[u] = hit[u] && (fi <= c.z);                          // dns_repertoire.py:44-49 (NaN never passes)
            if (SELF) hit[u] = hit[u] && (t + u != self_t);
            any = any || hit[u];
        }
        if (any) {                                                              // rare once the list has warmed up
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if ((CHECK || SELF ? hit[u] : true) && v[u] < top[KMAX - 1]) {
                    float w = v[u];
#pragma unroll
                    for (int q = 0; q < KMAX; ++q) { const float lo = fminf(top[q], w); w = fmaxf(top[q], w); top[q] = lo; }
                }
        }
    }
}

#ifndef QDX_DNS_ROWS
#define QDX_DNS_ROWS 256         // query rows (= threads) per CTA
#endif
template <int KMAX, int DD>
__global__ void __launch_bounds__(QDX_DNS_ROWS) qdx_dns_knn_sorted_kernel(const float* __restrict__ sf, const float* __restrict__ sd,
                                                                 const int32_t* __restrict__ perm, int64_t N, int32_t k,
                                                                 float* __restrict__ meta, float* __restrict__ partial,
                                                                 unsigned* __restrict__ tickets) {
    constexpr int ROWS = QDX_DNS_ROWS;
    // gridDim.y CTAs share a block of rows: CTA (x, y) scans the candidate tiles y, y + gridDim.y, ... of the block's prefix (an
    // interleaved split halves every block's serial scan, long or short, and doubles the warps in flight); each keeps its own
    // sorted k-list, the last one to finish (ticket) merges the lists -- the k smallest of a union are the k smallest of the
    // parts' k smallest -- and writes the result.
    const int ns = (int)gridDim.y, part = (int)blockIdx.y;
    __shared__ bool s_merge;
    constexpr int TILE = 1024;
    __shared__ float4 s_c[TILE];                    // (d0, d1, fitness, d2)
    __shared__ float s_d3[DD == 4 ? TILE : 4];
    __shared__ long long s_bound, s_nnan;
    // (Measured, profiles/r2_notes.md section 5: the work of a CTA is proportional to its prefix and with one CTA per block of rows
    //  the whole grid is resident at once, so the kernel lasts as long as the unluckiest SM -- 2.2 ... 3.6 ms from run to run.
    //  Walking odd rounds of CTAs backwards to pair long with short prefixes did not help: the hardware does not deal CTAs
    //  round-robin.  Splitting every block's tiles over 3 CTAs of 256 rows makes 1.33 waves of CTAs, the second of which fills
    //  SMs as they drain: 1.92 ms, stable; sweep of rows 64 / 128 / 256 x split 1 ... 8 in the notes.)
    const int64_t b = (int64_t)blockIdx.x;
    const int64_t cta = (int64_t)gridDim.x - 1 - b;                        // longest prefixes first
    const int64_t r0 = cta * ROWS, r = r0 + threadIdx.x;
    const bool valid = r < N;
    const float fi = valid ? sf[r] : -INFINITY;
    if (sf[r0] == -INFINITY) {                       // sorted descending: the whole CTA is empty slots (:144-145)
        if (valid) meta[perm[r]] = -INFINITY;
        return;
    }
    if (threadIdx.x == 0) {                          // end of the ties of the CTA's last (lowest-fitness) row
        const int64_t last = r0 + ROWS - 1 < N - 1 ? r0 + ROWS - 1 : N - 1;
        const uint32_t ok = qdx_order_key(sf[last]);
        int64_t lo = last + 1, hi = N;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (qdx_order_key(sf[mid]) >= ok) lo = mid + 1; else hi = mid; }
        s_bound = lo;
        lo = 0; hi = N;                              // NaN fitnesses sort first: how many are there
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (qdx_order_key(sf[mid]) == 0xFFFFFFFFu) lo = mid + 1; else hi = mid; }
        s_nnan = lo;
    }
    float xi[DD];
#pragma unroll
    for (int d = 0; d < DD; ++d) xi[d] = valid ? sd[r * DD + d] : 0.0f;
    float top[KMAX];
#pragma unroll
    for (int t = 0; t < KMAX; ++t) top[t] = INFINITY;
    __syncthreads();
    const int64_t bound = s_bound;
    const int64_t self_tile = (r0 / TILE) * TILE;
    const bool active = valid && fi != -INFINITY && fi == fi;
    int cnt = 0;                                     // fitter neighbours: rows up to the end of my ties, minus NaN rows, minus me
    if (active) {
        const uint32_t ok = qdx_order_key(fi);
        int64_t lo = r + 1, hi = bound;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (qdx_order_key(sf[mid]) >= ok) lo = mid + 1; else hi = mid; }
        const int64_t m = lo - s_nnan - 1;
        cnt = m < (int64_t)k ? (int)m : k;
    }
    for (int64_t j0 = (int64_t)part * TILE; j0 < bound; j0 += (int64_t)ns * TILE) {
        const int n = (bound - j0) < TILE ? (int)(bound - j0) : TILE;
        const int n8 = (n + 7) & ~7;
        __syncthreads();
        for (int t = threadIdx.x; t < n8; t += blockDim.x) {
            float4 c = make_float4(0.0f, 0.0f, NAN, 0.0f);                     // padding: never passes f_i <= NaN
            if (t < n) {
                const float* dj = sd + (j0 + t) * DD;
                c.z = sf[j0 + t]; c.x = dj[0];
                if (DD >= 2) c.y = dj[1];
                if (DD >= 3) c.w = dj[2];
                if (DD >= 4) s_d3[t] = dj[3];
            } else if (DD >= 4) s_d3[t] = 0.0f;
            s_c[t] = c;
        }
        __syncthreads();
        if (active) {
            // every candidate in a tile that lies behind the NaN rows and entirely ahead of this CTA's first row is fitter
            // than (or ties with) every row of the CTA: sorted order makes the fitness test redundant there
            if (j0 == self_tile) qdx_dns_scan_tile<KMAX, DD, true, true>(s_c, s_d3, n8, (int)(r - j0), fi, xi, top);
            else if (j0 < s_nnan || j0 + n > r0 || n8 != n) qdx_dns_scan_tile<KMAX, DD, false, true>(s_c, s_d3, n8, -1, fi, xi, top);
            else qdx_dns_scan_tile<KMAX, DD, false, false>(s_c, s_d3, n8, -1, fi, xi, top);
        }
    }
    if (ns > 1) {
        if (valid) {
            float* mine = partial + ((int64_t)part * N + r) * KMAX;
#pragma unroll
            for (int u = 0; u < KMAX; ++u) mine[u] = top[u];
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_merge = atomicAdd(&tickets[cta], 1u) == (unsigned)(ns - 1);
        __syncthreads();
        if (!s_merge) return;
        __threadfence();
        if (threadIdx.x == 0) tickets[cta] = 0u;                                    // re-armed for the next call
        if (valid) {
            for (int h = 0; h < ns; ++h) {
                if (h == part) continue;
                const float* other = partial + ((int64_t)h * N + r) * KMAX;
#pragma unroll
                for (int u = 0; u < KMAX; ++u) {
                    float w = __ldcg(other + u);
                    if (w < top[KMAX - 1]) {
#pragma unroll
                        for (int q = 0; q < KMAX; ++q) { const float lo = fminf(top[q], w); w = fmaxf(top[q], w); top[q] = lo; }
                    }
                }
            }
        }
    }
    if (!valid) return;
    float out;
    if (fi == -INFINITY) out = -INFINITY;                                           // :144-145
    else {
        float tot = 0.0f;
#pragma unroll
        for (int u = 0; u < KMAX; ++u) if (u < cnt) tot = tot + __fsqrt_rn(top[u]);   // :52, :70-74 (top-k order)
        out = __fdiv_rn(tot, (float)cnt);                                           // 0/0 -> NaN
    }
    meta[perm[r]] = out;
}

__global__ void __launch_bounds__(256) qdx_dns_cat_fit_kernel(QdxCat c, float* __restrict__ f_all) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c.N) f_all[i] = c.fit(i);
}
__global__ void __launch_bounds__(256) qdx_dns_sorted_gather_kernel(QdxCat c, const int32_t* __restrict__ perm, float* __restrict__ sf,
                                                                    float* __restrict__ sd) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= c.N) return;
    const int64_t i = perm[r];
    sf[r] = c.fit(i);
    for (int d = 0; d < c.Dd; ++d) sd[r * c.Dd + d] = c.desc(i, d);
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
    // 128-bit copies only for 16-byte aligned rows (D % 4 == 0 and aligned bases: an offset view is not)
    if ((D & 3) == 0 && ((((uintptr_t)pg) | ((uintptr_t)bg) | ((uintptr_t)out_g)) & 15u) == 0) {
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
    const unsigned gknn = (unsigned)((N + QDX_DNS_ROWS - 1) / QDX_DNS_ROWS);
    if (desc_dim <= 4) {
        // candidates sorted by descending fitness (bucketed rank), then the triangular k-NN scan
        float* f_all = nullptr;
        static int ns = -1;          // QDX_DNS_SPLIT=1: one CTA per block of rows (A/B)
        if (ns < 0) { const char* e_ = getenv("QDX_DNS_SPLIT"); ns = e_ ? atoi(e_) : 3; if (ns < 1 || ns > 8) ns = 3; }
        const int kmax = k <= 4 ? 4 : (k <= 16 ? 16 : 32);
        const size_t part_floats = ns > 1 ? (size_t)ns * (size_t)N * (size_t)kmax : 0;
        cudaError_t e = cudaMallocAsync((void**)&f_all, sizeof(float) * ((size_t)N * (2 + desc_dim) + part_floats) + sizeof(int32_t) * (size_t)N + sizeof(unsigned) * (size_t)gknn, st);
        if (e != cudaSuccess) return (int)e;
        float* sf = f_all + N; float* sd = sf + N; int32_t* perm = (int32_t*)(sd + N * desc_dim);
        float* partial = (float*)(perm + N); unsigned* tickets = (unsigned*)(partial + part_floats);
        e = cudaMemsetAsync(tickets, 0, sizeof(unsigned) * (size_t)gknn, st);
        if (e != cudaSuccess) { cudaFreeAsync(f_all, st); return (int)e; }
        qdx_dns_cat_fit_kernel<<<g256, 256, 0, st>>>(c, f_all);
        int rc = qdx_rank_desc(f_all, N, N, perm, st);
        if (rc == 0) {
            qdx_dns_sorted_gather_kernel<<<g256, 256, 0, st>>>(c, perm, sf, sd);
#define QDX_DNS_SORTED(KM)                                                                                             \
    do {                                                                                                               \
        if (desc_dim == 1) qdx_dns_knn_sorted_kernel<KM, 1><<<dim3(gknn, (unsigned)ns), QDX_DNS_ROWS, 0, st>>>(sf, sd, perm, N, k, meta_scratch, partial, tickets);      \
        else if (desc_dim == 2) qdx_dns_knn_sorted_kernel<KM, 2><<<dim3(gknn, (unsigned)ns), QDX_DNS_ROWS, 0, st>>>(sf, sd, perm, N, k, meta_scratch, partial, tickets); \
        else if (desc_dim == 3) qdx_dns_knn_sorted_kernel<KM, 3><<<dim3(gknn, (unsigned)ns), QDX_DNS_ROWS, 0, st>>>(sf, sd, perm, N, k, meta_scratch, partial, tickets); \
        else qdx_dns_knn_sorted_kernel<KM, 4><<<dim3(gknn, (unsigned)ns), QDX_DNS_ROWS, 0, st>>>(sf, sd, perm, N, k, meta_scratch, partial, tickets);                    \
    } while (0)
            if (k <= 4) QDX_DNS_SORTED(4);
            else if (k <= 16) QDX_DNS_SORTED(16);
            else QDX_DNS_SORTED(32);
#undef QDX_DNS_SORTED
            e = cudaGetLastError();
            if (e != cudaSuccess) rc = (int)e;
        }
        e = cudaFreeAsync(f_all, st);
        if (rc) return rc;
        if (e != cudaSuccess) return (int)e;
    } else {
        if (k <= 4) qdx_dns_novelty_generic_kernel<4><<<(unsigned)((N + 127) / 128), 128, 0, st>>>(c, k, meta_scratch);
        else if (k <= 16) qdx_dns_novelty_generic_kernel<16><<<(unsigned)((N + 127) / 128), 128, 0, st>>>(c, k, meta_scratch);
        else qdx_dns_novelty_generic_kernel<32><<<(unsigned)((N + 127) / 128), 128, 0, st>>>(c, k, meta_scratch);
        QDX_CHECK_LAUNCH();
    }
    { const int rc = qdx_rank_desc(meta_scratch, N, P, survivors_scratch, st); if (rc) return rc; }
    qdx_dns_gather_kernel<<<(unsigned)((P * 32 + 255) / 256), 256, 0, st>>>(pop_genotypes, batch_genotypes, c, (int32_t)D,
                                                                           survivors_scratch, out_genotypes, out_fitness, out_desc);
    QDX_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
