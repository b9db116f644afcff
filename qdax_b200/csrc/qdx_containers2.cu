// The two remaining sibling insertion rules of SURVEY.md 8(f) rank 3.  Reference files under /root/reference:
//   MOMERepertoire.add           qdax/core/containers/mome_repertoire.py:211-322 (_update_masked_pareto_front :72-209,
//                                qdax/utils/pareto_front.py:48-95)
//   UnstructuredRepertoire.add   qdax/core/containers/unstructured_repertoire.py:162-337 (get_cells_indices :21-66,
//                                intra_batch_comp :69-129)
// Both are restated line by line (oracle/qdax_containers_numpy.py is the literal NumPy twin and says, in its header, where the
// source does something surprising and how that is read here: PARITY UNPINNED at the jaxlib boundary).
#include "qdx_common.cuh"
#include "../../include/qdx.h"

#define QDX_CHECK_LAUNCH_D()                             \
    do {                                                 \
        cudaError_t e_ = cudaGetLastError();             \
        if (e_ != cudaSuccess) return (int)e_;           \
    } while (0)

namespace {

constexpr int MOME_MAX_L = 256;      // Pareto-front slots per cell
constexpr int MOME_MAX_C = 8;        // objectives

// ---------------------------------------------------------------------------------------------------------------- MOME
// The reference scans the batch with lax.scan, one offspring after the other, each updating the front of its cell
// (mome_repertoire.py:246-322).  Offspring of different cells are independent, those of one cell are not: one WARP per
// cell walks the batch in index order (every lane reads cells[b]: one broadcast load) and applies _add_one to the
// offspring that land in its cell:
//   mask_m   = any_c(f[m, c] == -inf)                                          (:260)
//   front_i  = !mask_i && !exists j unmasked: any_c(f_j - f_i > 0) && all_c(f_j - f_i >= 0)   over the L slots + the new point
//   the front members keep their order and move to the head (indices = sort(i * front + L * !front), :148-152); every slot
//   behind them receives A COPY OF THE NEW POINT's genotype (index L of the concatenation), zero descriptors and -inf
//   fitness; a new point that does not fit (L + 1 front members) is dropped by the truncation to L (:181).
// float * bool products (:179, :184, :193, :288): ieee_literal = 0 follows XLA's compiled Select(mask, x, 0); 1 takes the IEEE
// product of the converted mask (inf * 0 = NaN: every valid fitness of a touched cell becomes NaN).
__global__ void __launch_bounds__(128) qdx_mome_add_kernel(float* __restrict__ rep_f, float* __restrict__ rep_g, float* __restrict__ rep_d,
                                                           int64_t K, int32_t L, int32_t C, int32_t D, int32_t Dd,
                                                           const int32_t* __restrict__ cells, const float* __restrict__ bf,
                                                           const float* __restrict__ bg, const float* __restrict__ bd, int64_t B,
                                                           int32_t ieee_literal) {
    __shared__ float s_f[4][(MOME_MAX_L + 1) * MOME_MAX_C];
    __shared__ unsigned char s_mask[4][MOME_MAX_L + 1];
    __shared__ unsigned char s_front[4][MOME_MAX_L + 1];
    __shared__ int16_t s_src[4][MOME_MAX_L];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t c = (int64_t)blockIdx.x * 4 + w;
    if (c >= K) return;
    float* cf = rep_f + c * (int64_t)L * C;
    float* cg = rep_g + c * (int64_t)L * D;
    float* cd = rep_d + c * (int64_t)L * Dd;
    float* f = s_f[w]; unsigned char* mask = s_mask[w]; unsigned char* front = s_front[w]; int16_t* src = s_src[w];
    const float NANF = __int_as_float(0x7fc00000);
    for (int64_t b = 0; b < B; ++b) {
        if (__ldg(cells + b) != (int32_t)c) continue;                    // warp-uniform
        // ---- concatenation (front, new point) in shared memory
        for (int i = lane; i < L * C; i += 32) f[i] = cf[i];
        for (int i = lane; i < C; i += 32) f[L * C + i] = bf[b * C + i];
        __syncwarp();
        for (int m = lane; m <= L; m += 32) {
            bool mk = false;
            if (m < L) for (int k = 0; k < C; ++k) mk = mk || (f[m * C + k] == -INFINITY);
            mask[m] = mk;
        }
        __syncwarp();
        // ---- masked Pareto front (pareto_front.py:48-95), member i on lane i % 32
        for (int i = lane; i <= L; i += 32) {
            bool dominated = false;
            for (int j = 0; j <= L && !dominated; ++j) {
                bool any_gt = false, all_ge = true;
                for (int k = 0; k < C; ++k) {
                    const float diff = mask[j] ? -1.0f : f[j * C + k] - f[i * C + k];
                    any_gt = any_gt || (diff > 0.0f);
                    all_ge = all_ge && (diff >= 0.0f);
                }
                dominated = any_gt && all_ge;
            }
            front[i] = (!dominated && !mask[i]) ? 1 : 0;
        }
        __syncwarp();
        // ---- ordered compaction: src[pos] = pos-th front member (L = the new point); lane 0 (L <= 256)
        int num = 0;
        if (lane == 0) {
            for (int i = 0; i <= L; ++i) if (front[i]) { if (num < L) src[num] = (int16_t)i; ++num; }
            for (int pos = (num < L ? num : L); pos < L; ++pos) src[pos] = (int16_t)L;
        }
        num = __shfl_sync(0xffffffffu, num, 0);
        __syncwarp();
        const bool any_front = num > 0;                                  // new_mask_indices[0] (:184)
        // ---- rows, ascending positions: a member only ever moves towards the head, so the source of position pos is still intact
        for (int pos = 0; pos < L; ++pos) {
            const int s = src[pos];
            const bool valid = pos < num;
            if (s == pos && valid && !ieee_literal) continue;             // unchanged slot
            const float* gs = s == L ? bg + b * (int64_t)D : cg + (int64_t)s * D;
            const float* ds = s == L ? bd + b * (int64_t)Dd : cd + (int64_t)s * Dd;
            for (int k = lane; k < D; k += 32) {
                const float x = gs[k];
                cg[(int64_t)pos * D + k] = any_front ? x : (ieee_literal ? x * 0.0f : 0.0f);
            }
            for (int k = lane; k < Dd; k += 32) {
                const float x = ds[k];
                cd[(int64_t)pos * Dd + k] = valid ? x : (ieee_literal ? x * 0.0f : 0.0f);
            }
            for (int k = lane; k < C; k += 32) {
                const float x = f[s * C + k];
                float v;
                if (valid) v = ieee_literal ? NANF : x - 0.0f;           // x - inf * 0  |  x - select(mask, inf, 0)
                else v = ieee_literal ? (x * 0.0f) - INFINITY : -INFINITY;
                cf[pos * C + k] = v;
            }
            __syncwarp();
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------- unstructured
struct UnstrHeader { int32_t n_occ; int32_t first_occ; int32_t n_empty; int32_t add_one; int32_t n_new; int32_t pad[3]; };

// One CTA: occupancy of the archive (fitness != -inf), its first occupied slot, the ordered list of the first B slots with
// an infinite fitness (jnp.nonzero(isinf(fitnesses), size=B, fill_value=-1), :223-229), and whether the batch's finite
// fitnesses are all equal (the "virtual fitness" switch of intra_batch_comp, :89-91).
__global__ void __launch_bounds__(1024) qdx_unstr_scan_kernel(const float* __restrict__ rep_f, int64_t N, const float* __restrict__ bf,
                                                              int64_t B, UnstrHeader* hdr, int32_t* __restrict__ empty_idx) {
    __shared__ int s_cnt[32], s_emp[32], s_first[32];
    __shared__ float s_mx[32], s_mn[32];
    __shared__ int s_has[32];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5, nw = blockDim.x >> 5;
    const int64_t chunk = (N + nw - 1) / nw;
    const int64_t lo = (int64_t)w * chunk, hi = lo + chunk < N ? lo + chunk : N;
    int cnt = 0, emp = 0, first = 0x7fffffff;
    for (int64_t c0 = lo; c0 < hi; c0 += 32) {
        const int64_t c = c0 + lane;
        const float v = c < hi ? rep_f[c] : 0.0f;
        const bool occ = c < hi && v != -INFINITY, e = c < hi && isinf(v);
        const unsigned bo = __ballot_sync(0xffffffffu, occ), be = __ballot_sync(0xffffffffu, e);
        if (bo && first == 0x7fffffff) first = (int)(c0 + __ffs(bo) - 1);
        cnt += __popc(bo); emp += __popc(be);
    }
    if (lane == 0) { s_cnt[w] = cnt; s_emp[w] = emp; s_first[w] = first; }
    // batch statistics: nanmax == nanmin over the fitnesses with +-inf -> NaN (:83-91)
    float mx = -INFINITY, mn = INFINITY; int has = 0;
    for (int64_t i = t; i < B; i += blockDim.x) {
        const float v = bf[i];
        if (v == v && !isinf(v)) { mx = fmaxf(mx, v); mn = fminf(mn, v); has = 1; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); has |= __shfl_xor_sync(0xffffffffu, has, o);
    }
    if (lane == 0) { s_mx[w] = mx; s_mn[w] = mn; s_has[w] = has; }
    __syncthreads();
    int before = 0;                                       // empties in the warps ahead of mine
    for (int i = 0; i < w; ++i) before += s_emp[i];
    for (int64_t c0 = lo; c0 < hi && before < B; c0 += 32) {
        const int64_t c = c0 + lane;
        const bool e = c < hi && isinf(rep_f[c]);
        const unsigned be = __ballot_sync(0xffffffffu, e);
        const int pos = before + __popc(be & ((1u << lane) - 1u));
        if (e && pos < B) empty_idx[pos] = (int32_t)c;
        before += __popc(be);
    }
    if (t == 0) {
        int n = 0, ne = 0, f0 = 0x7fffffff, h = 0; float M = -INFINITY, m = INFINITY;
        for (int i = 0; i < nw; ++i) { n += s_cnt[i]; ne += s_emp[i]; if (s_first[i] < f0) f0 = s_first[i]; M = fmaxf(M, s_mx[i]); m = fminf(m, s_mn[i]); h |= s_has[i]; }
        hdr->n_occ = n; hdr->first_occ = n ? f0 : 0; hdr->n_empty = ne < B ? ne : (int32_t)B; hdr->add_one = (h && M == m) ? 1 : 0;
    }
    __syncthreads();
    {   // pad the list with -1
        int ne = 0;
        for (int i = 0; i < nw; ++i) ne += s_emp[i];
        for (int64_t i = ne + t; i < B; i += blockDim.x) empty_idx[i] = -1;
    }
}

// Thread = offspring: its distance to "every occupied slot" as the reference computes it -- the Frobenius norm over ALL
// stored descriptors (see the header of oracle/qdax_containers_numpy.py): per slot left to right over d, slot after slot
// from +0, sqrt -- and the two l-value tests on the nearest / second-nearest distance (:196-234).
__global__ void __launch_bounds__(128) qdx_unstr_near_kernel(const float* __restrict__ bd, int64_t B, int32_t Dd, const float* __restrict__ rep_d,
                                                             int64_t N, const UnstrHeader* hdr, float l_value, unsigned char* __restrict__ near,
                                                             unsigned char* __restrict__ not_novel) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float tot = 0.0f;
    for (int64_t j = 0; j < N; ++j) {
        float s = 0.0f;
        for (int k = 0; k < Dd; ++k) { const float df = bd[b * Dd + k] - __ldg(rep_d + j * Dd + k); const float q = df * df; s = k ? s + q : q; }
        tot = tot + s;
    }
    const float F = __fsqrt_rn(tot);
    const float d0 = hdr->n_occ >= 1 ? F : INFINITY, d1 = hdr->n_occ >= 2 ? F : INFINITY;
    near[b] = d0 <= l_value;
    not_novel[b] = d1 <= l_value;
}

// One CTA: the re-ordering of the batch (:238-263).  Offspring that are not near an occupied slot carry index -1 and come
// first, in batch order; the near ones (all aimed at the same slot, the first occupied one) follow, in batch order --
// jax.lax.top_k(-indices, B) is a stable ascending sort.  Position p gets its target slot: the first occupied slot, or the
// p-th empty slot (-1 = none left, which wraps to the last slot like jnp indexing does).
__global__ void __launch_bounds__(1024) qdx_unstr_order_kernel(const unsigned char* __restrict__ near, int64_t B, int64_t N, UnstrHeader* hdr,
                                                               const int32_t* __restrict__ empty_idx, int32_t* __restrict__ order,
                                                               int32_t* __restrict__ target) {
    __shared__ int s_new[32];
    __shared__ int s_total_new;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5, nw = blockDim.x >> 5;
    const int64_t chunk = ((B + nw - 1) / nw + 31) / 32 * 32;
    const int64_t lo = (int64_t)w * chunk, hi = lo + chunk < B ? lo + chunk : B;
    int cnt = 0;
    for (int64_t i0 = lo; i0 < hi; i0 += 32) { const int64_t i = i0 + lane; cnt += __popc(__ballot_sync(0xffffffffu, i < hi && !near[i])); }
    if (lane == 0) s_new[w] = cnt;
    __syncthreads();
    if (t == 0) { int n = 0; for (int i = 0; i < nw; ++i) n += s_new[i]; s_total_new = n; hdr->n_new = n; }
    __syncthreads();
    const int total_new = s_total_new;
    int new_before = 0;
    for (int i = 0; i < w; ++i) new_before += s_new[i];
    int64_t near_before = lo - new_before;
    for (int64_t i0 = lo; i0 < hi; i0 += 32) {
        const int64_t i = i0 + lane;
        const bool in = i < hi, nr = in && near[i];
        const unsigned bn = __ballot_sync(0xffffffffu, in && !nr), br = __ballot_sync(0xffffffffu, nr);
        const unsigned below = (1u << lane) - 1u;
        if (in) {
            const int64_t pos = nr ? total_new + near_before + __popc(br & below) : new_before + __popc(bn & below);
            order[pos] = (int32_t)i;
            int32_t tg = nr ? hdr->first_occ : empty_idx[pos];
            if (tg < 0) tg += (int32_t)N;
            target[pos] = tg;
        }
        new_before += __popc(bn); near_before += __popc(br);
    }
}

// Thread = position in the re-ordered batch: intra_batch_comp (:69-129) -- discarded when another offspring closer than
// l_value (Euclidean, left-to-right sum, sqrt) has a strictly higher (virtual) fitness, or when its own descriptor holds a NaN.
__global__ void __launch_bounds__(128) qdx_unstr_keep_kernel(const float* __restrict__ d, const float* __restrict__ f, int64_t B, int32_t Dd,
                                                             const UnstrHeader* hdr, float l_value, const unsigned char* __restrict__ not_novel,
                                                             const int32_t* __restrict__ order, unsigned char* __restrict__ keep) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const float add = hdr->add_one ? 1.0f : 0.0f;
    auto virt = [&](int64_t j) {                                     // eval_scores + linspace(0, add, B) (:83-99)
        float v = f[j];
        if (isinf(v) || v != v) v = -INFINITY;
        const float lin = B > 1 ? (j < B - 1 ? 0.0f * (1.0f - __fdiv_rn((float)j, (float)(B - 1))) + add * __fdiv_rn((float)j, (float)(B - 1)) : add) : 0.0f;
        return v + lin;
    };
    bool not_existent = false;
    for (int k = 0; k < Dd; ++k) not_existent = not_existent || (d[i * Dd + k] != d[i * Dd + k]);
    const float mine = virt(i);
    bool discard = not_existent;
    for (int64_t j = 0; j < B && !discard; ++j) {
        if (j == i) continue;
        float s = 0.0f;
        for (int k = 0; k < Dd; ++k) {
            float x = d[i * Dd + k];
            if (x != x) x = INFINITY;                                  // :81
            const float df = x - d[j * Dd + k]; const float q = df * df; s = k ? s + q : q;
        }
        if (__fsqrt_rn(s) < l_value && virt(j) > mine) discard = true;
    }
    keep[i] = !discard && !not_novel[order[i]];
}

// segment_max + tie-break + strict improvement (:288-311) on the packed-key machinery: pass 0 takes the per-slot maximum
// fitness over ALL offspring (kept or not) in a scratch table; pass 1 offers the kept ones that equal it and beat the occupant.
__global__ void __launch_bounds__(256) qdx_unstr_offer_kernel(const int32_t* __restrict__ target, const float* __restrict__ f, const unsigned char* __restrict__ keep,
                                                              int64_t B, int64_t N, void* ws, const float* __restrict__ rep_f,
                                                              unsigned long long* __restrict__ best, int32_t first_wins, int32_t pass) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const int32_t c = target[i];
    if (c < 0 || c >= N) { qdx_set_error(ws, QDX_ERR_BAD_CELL); return; }
    const float v = f[i];
    const unsigned long long fk = (unsigned long long)qdx_order_key(v);
    if (pass == 0) { atomicMax(best + c, fk + 1ull); return; }          // + 1: 0 = no offspring (order_key(-inf) = 0x007FFFFF > 0 anyway)
    const unsigned long long bk = best[c] - 1ull;
    if (bk == 0xFFFFFFFFull) return;                                    // a NaN poisons its slot (segment_max propagates it)
    if (fk != bk || !keep[i]) return;
    qdx_offer(ws, N, rep_f, c, v, (uint32_t)i, first_wins);
}

}  // namespace

extern "C" {

int qdx_mome_add(float* rep_fitness, float* rep_genotypes, float* rep_desc, int64_t K, int32_t front_len, int32_t num_criteria, int64_t D,
                 int32_t desc_dim, const int32_t* cells, const float* fitness, const float* genotypes, const float* desc, int64_t B,
                 int32_t ieee_literal, void* stream) {
    if (!rep_fitness || !rep_genotypes || !rep_desc || K <= 0 || D <= 0 || desc_dim < 1 || B < 0) return QDX_ERR_ARG;
    if (front_len < 1 || front_len > MOME_MAX_L || num_criteria < 1 || num_criteria > MOME_MAX_C) return QDX_ERR_UNSUPPORTED;
    if (B == 0) return 0;
    if (!cells || !fitness || !genotypes || !desc) return QDX_ERR_ARG;
    qdx_mome_add_kernel<<<(unsigned)((K + 3) / 4), 128, 0, (cudaStream_t)stream>>>(rep_fitness, rep_genotypes, rep_desc, K, front_len, num_criteria,
                                                                                     (int32_t)D, desc_dim, cells, fitness, genotypes, desc, B, ieee_literal);
    QDX_CHECK_LAUNCH_D();
    return 0;
}

int qdx_unstructured_scratch(int64_t N, int64_t B, int64_t* bytes) {
    if (N <= 0 || B < 0 || !bytes) return QDX_ERR_ARG;
    // header | empty_idx (B) | order (B) | target (B) | near, not_novel, keep (B each) | best (N u64)
    *bytes = (int64_t)(256 + qdx_align_up(sizeof(int32_t) * (size_t)B, 256) * 3 + qdx_align_up((size_t)B, 256) * 3 + sizeof(unsigned long long) * (size_t)N);
    return 0;
}

int qdx_unstructured_plan(const float* rep_fitness, const float* rep_desc, int64_t N, int32_t desc_dim, const float* fitness, const float* desc,
                          int64_t B, float l_value, void* scratch, int32_t* out_order, void* stream) {
    if (!rep_fitness || !rep_desc || !scratch || !out_order || N <= 0 || desc_dim < 1 || B <= 0 || !fitness || !desc) return QDX_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    char* s = (char*)scratch;
    const size_t bi = qdx_align_up(sizeof(int32_t) * (size_t)B, 256), bb = qdx_align_up((size_t)B, 256);
    UnstrHeader* hdr = (UnstrHeader*)s;
    int32_t* empty_idx = (int32_t*)(s + 256); int32_t* target = (int32_t*)(s + 256 + 2 * bi);
    unsigned char* near = (unsigned char*)(s + 256 + 3 * bi); unsigned char* not_novel = near + bb;
    qdx_unstr_scan_kernel<<<1, 1024, 0, st>>>(rep_fitness, N, fitness, B, hdr, empty_idx);
    qdx_unstr_near_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(desc, B, desc_dim, rep_desc, N, hdr, l_value, near, not_novel);
    qdx_unstr_order_kernel<<<1, 1024, 0, st>>>(near, B, N, hdr, empty_idx, out_order, target);
    QDX_CHECK_LAUNCH_D();
    return 0;
}

int qdx_unstructured_offer(const float* sorted_fitness, const float* sorted_desc, int64_t B, int32_t desc_dim, int64_t N, float l_value,
                           void* scratch, const int32_t* order, void* ws, const float* rep_fitness, int32_t first_wins, void* stream) {
    if (!sorted_fitness || !sorted_desc || !scratch || !order || !ws || !rep_fitness || B <= 0 || N <= 0 || desc_dim < 1) return QDX_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    char* s = (char*)scratch;
    const size_t bi = qdx_align_up(sizeof(int32_t) * (size_t)B, 256), bb = qdx_align_up((size_t)B, 256);
    const UnstrHeader* hdr = (const UnstrHeader*)s;
    const int32_t* target = (const int32_t*)(s + 256 + 2 * bi);
    const unsigned char* not_novel = (const unsigned char*)(s + 256 + 3 * bi) + bb;
    unsigned char* keep = (unsigned char*)(s + 256 + 3 * bi) + 2 * bb;
    unsigned long long* best = (unsigned long long*)(s + 256 + 3 * bi + 3 * bb);
    cudaError_t e = cudaMemsetAsync(best, 0, sizeof(unsigned long long) * (size_t)N, st);
    if (e != cudaSuccess) return (int)e;
    qdx_unstr_keep_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(sorted_desc, sorted_fitness, B, desc_dim, hdr, l_value, not_novel, order, keep);
    const unsigned g = (unsigned)((B + 255) / 256);
    qdx_unstr_offer_kernel<<<g, 256, 0, st>>>(target, sorted_fitness, keep, B, N, ws, rep_fitness, best, first_wins, 0);
    qdx_unstr_offer_kernel<<<g, 256, 0, st>>>(target, sorted_fitness, keep, B, N, ws, rep_fitness, best, first_wins, 1);
    QDX_CHECK_LAUNCH_D();
    return 0;
}

}  // extern "C"
