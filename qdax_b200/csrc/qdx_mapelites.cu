// MAP-Elites generation step on B200 (sm_100a): hand-written kernels behind the C ABI of include/qdx.h.
//
//   prepare   occupancy scan -> occupied-cell list + selection segments; device-side key chain
//   generate  parent sampling + Iso+LineDD variation (Threefry-2x32, counter-based) + fused task scoring
//             + grid cell assignment + packed 64-bit atomicMax offer          (stages a, b, c-grid, d-offer)
//   cells     brute-force nearest centroid, first-index argmin, + offer       (stage c, d-offer)
//   commit    per-cell winner row copy into the HBM-resident repertoire + QD metrics  (stage d)
//
// Layout: everything float32 row-major; a warp owns a tile of 32 consecutive offspring rows staged in shared
// memory as the exact global-memory image of those rows, so the tile leaves the SM with ONE bulk async copy
// (cp.async.bulk shared->global, the TMA engine) while the warp moves on.  Work inside a tile has two shapes:
// gene-parallel (RNG + variation: lanes stride over float4 quads of the tile, parents gathered with 128-bit
// coalesced loads) and row-serial (scoring: lane = row, sequential float32 reductions exactly as the spec).
//
// Reference semantics being reproduced (under /root/reference): qdax/core/map_elites.py:148-225,
// qdax/core/emitters/standard_emitters.py:27-82, .../repertoire_selectors/uniform_selector.py:22-62,
// qdax/core/emitters/mutation_operators.py:175-226, qdax/tasks/arm.py:9-50,
// qdax/tasks/standard_functions.py:9-48, qdax/core/containers/mapelites_repertoire.py:111-266,
// qdax/utils/metrics.py:74-98.
#include <cstdlib>
#include "qdx_common.cuh"
#include "qdx_cells_index.cuh"
#include "../../include/qdx.h"

int qdx_fill_cvt_index(const qdx_cvt_index* in, QdxCvtIndex* out);     // qdx_cells_index.cu

#define QDX_CHECK_LAUNCH() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return (int)e_; } while (0)

// =====================================================================================================
// prepare
// =====================================================================================================
__device__ void qdx_derive_gen_keys(QdxKey emit, QdxGenKeys* out) {
    QdxKey e0 = qdx_split(emit, 0), e1 = qdx_split(emit, 1), kv = qdx_split(emit, 2);   // standard_emitters.py:55
    out->sel1 = qdx_split(e0, 1);                                                        // uniform_selector.py:48
    out->sel2 = qdx_split(e1, 1);
    out->line = qdx_split(kv, 1);                                                        // mutation_operators.py:205
    out->leaf = qdx_split(qdx_split(kv, 0), 0);                                          // :220 (one leaf)
}

// key_mode: 0 keep keys; 1 `key` = key of MAPElites.update; 2 scan step on ws->carry; 3 `key` = key of
// DistributedMAPElites.update; 4 `key` = emit key.
__global__ void __launch_bounds__(1024) qdx_prepare_kernel(const float* __restrict__ rep_f, int64_t K, void* ws_raw,
                                                           QdxKey key, int key_mode, int rank_slot) {
    QdxWorkspace* ws = (QdxWorkspace*)ws_raw;
    __shared__ int32_t s_warp[33];
    const int t = threadIdx.x;

    if (rank_slot >= 0) {     // the other ranks' key slots must be empty before the all-reduce(max) fills them
        unsigned long long* slots = qdx_ws_keytab(ws_raw, K) + K;
        for (int j = t; j < 8 * QDX_MAX_RANKS; j += blockDim.x) if ((j >> 3) != rank_slot) slots[j] = 0ull;
    }
    if (t == 32 && key_mode != 0) {      // key chain on warp 1, overlapping the scan
        if (key_mode == 2) { QdxKey c = ws->carry; key = qdx_split(c, 1); ws->carry = qdx_split(c, 0); }  // map_elites.py:214
        QdxKey emit = key;
        if (key_mode == 1 || key_mode == 2) emit = qdx_split(qdx_split(key, 1), 1);                         // :177, :241
        else if (key_mode == 3) emit = qdx_split(key, 1);                                // distributed_map_elites.py:124
        qdx_derive_gen_keys(emit, &ws->keys);
        if (rank_slot >= 0) {   // publish this rank's generation keys behind the key table: they ride in the same all-reduce
            unsigned long long* slot = qdx_ws_keytab(ws_raw, K) + K + 8 * rank_slot;
            const QdxGenKeys g = ws->keys;
            const uint32_t w[8] = {g.sel1.a, g.sel1.b, g.sel2.a, g.sel2.b, g.line.a, g.line.b, g.leaf.a, g.leaf.b};
            for (int j = 0; j < 8; ++j) slot[j] = (unsigned long long)w[j];
        }
    }

    qdx_cta_occupancy_scan(rep_f, K, ws_raw, s_warp);
    if (t == 0 && (ws->sel.M <= 0 || ws->sel.nseg <= 0)) qdx_set_error(ws, QDX_ERR_EMPTY_REPERTOIRE);
}

// =====================================================================================================
// grid cell assignment (exactly equal to the brute-force argmin)
// =====================================================================================================
template <int DD>
__device__ __forceinline__ int32_t qdx_cell_bruteforce_row(const float* x, const float* __restrict__ cent, int64_t K) {
    float best = INFINITY; int32_t bk = 0;
    for (int64_t k = 0; k < K; ++k) {
        float acc = 0.0f;
#pragma unroll
        for (int d = 0; d < DD; ++d) { float df = x[d] - cent[k * DD + d]; float s = df * df; acc = d ? acc + s : s; }
        if (acc != acc) return (int32_t)k;
        if (acc < best) { best = acc; bk = (int32_t)k; }
    }
    return bk;
}

template <int DD>
__device__ __forceinline__ int32_t qdx_grid_cell(const float* x, const QdxGrid& g, const float* s_axes,
                                                 const float* __restrict__ cent, int64_t K) {
    int32_t ci[DD][3]; float ca[DD][3]; int32_t cn[DD];
#pragma unroll
    for (int d = 0; d < DD; ++d) {
        const float xd = x[d];
        if (!(fabsf(xd) <= 3.40282347e+38f)) return 0;           // NaN / inf: every distance NaN or inf -> index 0
        if (xd < g.lo[d] || xd > g.hi[d]) return qdx_cell_bruteforce_row<DD>(x, cent, K);   // rare, exact
        const float* ax = s_axes + g.off[d];
        int lo = 0, hi = g.n[d];                                 // lower_bound: first ax[p] >= xd
        while (lo < hi) { int mid = (lo + hi) >> 1; if (ax[mid] < xd) lo = mid + 1; else hi = mid; }
        int c0 = lo - 1 < 0 ? 0 : lo - 1;
        int c1 = lo + 1 > g.n[d] - 1 ? g.n[d] - 1 : lo + 1;
        cn[d] = c1 - c0 + 1;
#pragma unroll
        for (int m = 0; m < 3; ++m) { const int c = (c0 + m <= c1) ? c0 + m : c1; float df = xd - ax[c]; ci[d][m] = c; ca[d][m] = df * df; }
    }
    float best = INFINITY; int32_t bflat = 0x7fffffff;
    // all 3^DD candidate cells, compile-time indices (registers only); lowest flat index wins ties
#pragma unroll
    for (int m0 = 0; m0 < 3; ++m0)
#pragma unroll
        for (int m1 = 0; m1 < (DD > 1 ? 3 : 1); ++m1)
#pragma unroll
            for (int m2 = 0; m2 < (DD > 2 ? 3 : 1); ++m2)
#pragma unroll
                for (int m3 = 0; m3 < (DD > 3 ? 3 : 1); ++m3) {
                    bool ok = m0 < cn[0];
                    float acc = ca[0][m0]; int32_t flat = ci[0][m0] * g.stride[0];
                    if (DD > 1) { ok = ok && m1 < cn[DD > 1 ? 1 : 0]; acc = acc + ca[DD > 1 ? 1 : 0][m1]; flat += ci[DD > 1 ? 1 : 0][m1] * g.stride[1]; }
                    if (DD > 2) { ok = ok && m2 < cn[DD > 2 ? 2 : 0]; acc = acc + ca[DD > 2 ? 2 : 0][m2]; flat += ci[DD > 2 ? 2 : 0][m2] * g.stride[2]; }
                    if (DD > 3) { ok = ok && m3 < cn[DD > 3 ? 3 : 0]; acc = acc + ca[DD > 3 ? 3 : 0][m3]; flat += ci[DD > 3 ? 3 : 0][m3] * g.stride[3]; }
                    if (ok && (acc < best || (acc == best && flat < bflat))) { best = acc; bflat = flat; }
                }
    return bflat;
}

// =====================================================================================================
// generate
// =====================================================================================================
struct QdxGenParams {
    const float* rep_g; const float* rep_f; const float* centroids;
    void* ws;
    int64_t B; int64_t K; int32_t D; int32_t DC;     // DC = genes per staged chunk (multiple of 4)
    int32_t DS;                                       // shared-memory row stride in floats (DS/4 odd: conflict-free LDS.128)
    float iso_sigma, line_sigma; int32_t has_min, has_max; float minv, maxv;
    float* out_g; float* out_f; float* out_d; int32_t* out_cell; int32_t* out_p1; int32_t* out_p2;
    int32_t desc_dim;
    QdxGrid grid;
    int32_t offer; uint32_t idx_base; int32_t first_wins;
    int32_t keys_by_value; QdxGenKeys keys;          // generation keys derived on the host (qdx_host_generation_keys)
    int32_t tile_rows;                                // rows per tile handed to a warp (<= 32): the batch is a whole number of tiles per warp
    int32_t store_mode;                               // 0: every offspring row (bulk copy of the tile); 1: only the rows whose offer fired
    int32_t out_xchg;                                 // 1: out_g / out_f / out_d = this rank's offspring block of the exchange buffer (epoch parity)
    QdxCvtIndex cvt;                                  // bucket index over non-grid centroids (GRID_DD < 0)
    QdxLeafTab leaves;                                // pytree genotypes (MULTI instantiation only)
};

QDX_DEV void qdx_bulk_store(void* gptr, const void* sptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
                 :: "l"(gptr), "r"((uint32_t)__cvta_generic_to_shared(sptr)), "r"(bytes) : "memory");
}

// One offspring row, shared memory -> global, by the lane that owns it (store_mode 1: only the rare rows whose offer
// fired are kept -- in steady state ~0.1 % of them -- instead of streaming the whole 400 B x B offspring buffer to HBM every
// generation); `sys`: the row will be read by peer GPUs, fence at system scope.
QDX_DEV void qdx_store_row(float* __restrict__ dst, const float* src_smem, int n, int sys) {
    for (int d = 0; d < n; d += 4) *reinterpret_cast<float4*>(dst + d) = *reinterpret_cast<const float4*>(src_smem + d);
    (void)sys;      // ordered before the arrival flags by the publisher's fence (see qdx_xchg_publish)
}

constexpr int QDX_GEN_WARPS = 4;

#ifndef QDX_GEN_PLAIN_STORE
#define QDX_GEN_PLAIN_STORE 0
#endif
#ifndef QDX_GEN_TRACE
#define QDX_GEN_TRACE 0       // timing experiments only: first / last CTA start and end of the generate kernel (tools/time_generate.py)
#endif
#if QDX_GEN_TRACE
__device__ unsigned long long g_gen_trace[4];     // min start, max start, min end, max end (globaltimer ns)
#define QDX_GEN_STAMP(i, op) do { if (threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); op(&g_gen_trace[i], t_); } } while (0)
#else
#define QDX_GEN_STAMP(i, op)
#endif

// ARM_CLIP: arm.py:27 clips the genotype to [0,1] before scoring; when the variation already clipped to a range
// inside [0,1] that clip is the identity and is compiled out (bit-identical result).
template <int TASK, int GRID_DD, bool ARM_CLIP, bool MULTI>
__device__ __forceinline__ void qdx_generate_body(const QdxGenParams& p);

// MULTI: the packed row is the concatenation of several pytree leaves, each with its own noise key and counter space.
template <int TASK, int GRID_DD, bool ARM_CLIP, bool MULTI = false>
__global__ void __launch_bounds__(QDX_GEN_WARPS * 32, 4) qdx_generate_kernel(const QdxGenParams p) {
    qdx_pdl_enter();         // the selection tables / repertoire rows read below are the previous kernel's (commit) output
    QDX_GEN_STAMP(0, atomicMin); QDX_GEN_STAMP(1, atomicMax);
    qdx_generate_body<TASK, GRID_DD, ARM_CLIP, MULTI>(p);
#if QDX_GEN_TRACE
    __syncthreads();
    QDX_GEN_STAMP(2, atomicMin); QDX_GEN_STAMP(3, atomicMax);
#endif
    // the last CTA of the grid to finish re-arms the row counter for the next launch; multi-GPU peer-memory exchange: this
    // CTA's offers (and their pushes into the peers) are done, the last CTA also publishes this rank's generation keys and
    // raises its arrival flag in every peer
    __syncthreads();
    if (threadIdx.x == 0) {
        QdxWorkspace* ws = (QdxWorkspace*)p.ws;
        const bool publish = GRID_DD != 0 && p.offer && p.keys_by_value && ws->xchg_nranks > 0;
        __threadfence();
        if (atomicAdd(&ws->gen_done, 1u) == gridDim.x - 1u) {
            ws->gen_done = 0u; ws->gen_next_row = 0u;
            if (publish) { __threadfence(); qdx_xchg_publish(ws, p.K, p.keys); }
        }
    }
}

template <int TASK, int GRID_DD, bool ARM_CLIP, bool MULTI>
__device__ __forceinline__ void qdx_generate_body(const QdxGenParams& p) {
    extern __shared__ __align__(128) float s_tiles[];
    __shared__ QdxSeg s_seg[QDX_MAX_SEG];
    __shared__ float s_last[QDX_MAX_SEG];

    const QdxWorkspace* ws = (const QdxWorkspace*)p.ws;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nseg = ws->sel.nseg;
    const int32_t D = p.D, DC = p.DC, DS = p.DS;
    float* s_axes = s_tiles + (size_t)QDX_GEN_WARPS * 32 * DS;      // grid axis values, after the tiles
    for (int i = threadIdx.x; i < nseg; i += blockDim.x) { s_seg[i] = ws->sel.seg[i]; s_last[i] = ws->sel.last[i]; }
    if (GRID_DD > 0) for (int i = threadIdx.x; i < p.grid.total_axes; i += blockDim.x) s_axes[i] = p.grid.axes[i];
    __syncthreads();
    if (nseg <= 0) {           // empty repertoire: p = 0/0 in the reference (uniform_selector.py:45)
        if (blockIdx.x == 0 && threadIdx.x == 0) qdx_set_error(p.ws, QDX_ERR_EMPTY_REPERTOIRE);
        return;
    }

    float* tile = s_tiles + (size_t)warp * 32 * DS;
    // Persistent grid (at most one resident wave of CTAs), rows handed out to the WARPS from a grid-wide counter in tiles of
    // p.tile_rows <= 32 rows.  The tile height is chosen by the launcher so that the batch is a whole number of tiles per warp
    // (131 072 rows over 2368 warps = 55.4 rows each = two tiles of 28): with ceil(B / 128) CTAs of one 32-row tile per warp
    // the second, partial wave of such a shard (8-GPU run) ran with 27 % of the SMs' warp slots empty.  The deal must be
    // dynamic: co-resident warps do not progress at the same rate (the issue arbiter is priority-based), a static split left
    // SMs half empty for the last third of the kernel (first CTA done at 442 us, last at 737 us for 2^20 rows).  A finer
    // guided deal (tiles shrinking to 8 rows) lost more to the per-tile cost of the row-serial scoring phase than it gained.
    // Results do not depend on who processes which row: the random streams are counter-based on the row index.
    QdxWorkspace* wsm = (QdxWorkspace*)p.ws;
    const QdxGenKeys keys = p.keys_by_value ? p.keys : ws->keys;
    const int32_t* __restrict__ occ = qdx_ws_occ(p.ws);
    const float total = ws->sel.total;
    float* out_g = p.out_g; float* const out_f = p.out_f; float* const out_d = p.out_d;
    float* xf = nullptr; float* xd = nullptr;
    if (p.out_xchg) {        // multi-GPU: the rows go where the peers can read them (this rank's block of the current epoch parity),
        const QdxOffBlock ob = qdx_xchg_block(p.ws, p.K, ws->xchg_rank, qdx_xchg_parity(p.ws));      // fitness / descriptors to both places
        out_g = ob.g; xf = ob.f; xd = ob.d;
    }
    const uint32_t take = (uint32_t)p.tile_rows;
    for (;;) {
    uint32_t r0 = 0u;
    if (lane == 0) r0 = atomicAdd(&wsm->gen_next_row, take);
    r0 = __shfl_sync(0xffffffffu, r0, 0);
    if ((int64_t)r0 >= p.B) break;
    const int64_t row0 = (int64_t)r0;
    const int64_t row = row0 + lane;
    const int nrows = (p.B - row0) < (int64_t)take ? (int)(p.B - row0) : (int)take;
    const bool valid = lane < nrows;

    // ---- phase 0: parents + line noise, lane = row --------------------------------------------------
    // (lanes past the end of the tile compute harmless values and never store)
    int32_t p1, p2; float line;
    {
        float u1 = qdx_unit_float(qdx_bits32(keys.sel1, (uint64_t)row));
        float u2 = qdx_unit_float(qdx_bits32(keys.sel2, (uint64_t)row));
        p1 = occ[qdx_sel_rank(s_seg, s_last, nseg, total * (1.0f - u1)) - 1];
        p2 = occ[qdx_sel_rank(s_seg, s_last, nseg, total * (1.0f - u2)) - 1];
        __syncwarp();
        line = qdx_normal_from_bits_t<true>(qdx_bits32(keys.line, (uint64_t)row)) * p.line_sigma;
        if (valid && p.out_p1) p.out_p1[row] = p1;
        if (valid && p.out_p2) p.out_p2[row] = p2;
    }

    float acc0 = 0.0f;        // rastrigin / sphere running sum, carried across chunks
    const int nchunks = (D + DC - 1) / DC;
    for (int ch = 0; ch < nchunks; ++ch) {
        const int d0 = ch * DC;
        const int dc = (D - d0) < DC ? (D - d0) : DC;     // genes in this chunk (multiple of 4)
        const int q = dc >> 2;                            // quads per row in this chunk (<= 32)
        const uint32_t qmagic = (1u << 20) / (uint32_t)q + 1u;   // qi / q == (qi * qmagic) >> 20 for qi * q < 2^20
        // ---- phase 1: gene-parallel variation over the tile ---------------------------------------------
        // All 32 lanes stay converged (lanes past the end recompute the last quad and skip the store), so the
        // normal transform can vote with a full mask.  (Measured, no gain: prefetch.global.L2 of the next iteration's parent quads
        // when the repertoire is larger than L2 (c4: 200 MB) -- 0.460 vs 0.463 ms, and +4 % at c3 when forced (gpurun_out/r3s_*);
        // 2x unrolling (profiles/r1_notes.md); a software
        // pipeline that issues the NEXT quad's Threefry blocks next to the current quad's float transforms, to mix ALU- and
        // FMA-pipe work inside one warp: 0.710 vs 0.680 ms -- ptxas keeps the two streams apart, co-resident warps already
        // mix them (profiles/r2_notes.md).)
        const int total_quads = nrows * q;
        for (int qb = 0; qb < total_quads; qb += 32) {
            const bool act = qb + lane < total_quads;
            const int qi = act ? qb + lane : total_quads - 1;
            const int rr = (int)(((uint32_t)qi * qmagic) >> 20);
            const int dq = qi - rr * q;
            const int32_t pa = __shfl_sync(0xffffffffu, p1, rr);
            const int32_t pb = __shfl_sync(0xffffffffu, p2, rr);
            const float ln = __shfl_sync(0xffffffffu, line, rr);
            const int d = d0 + (dq << 2);
            const float4 a = __ldg(reinterpret_cast<const float4*>(p.rep_g + (int64_t)pa * D + d));
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.rep_g + (int64_t)pb * D + d));
            const uint64_t ctr = (uint64_t)(row0 + rr) * (uint64_t)D + (uint64_t)d;
            float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w}, xv[4], nz[4];
#ifdef QDX_GEN_SERIAL_NORMALS      // A/B switch (profiles/r1_notes.md): one serial chain + one branch per draw
#pragma unroll
            for (int j = 0; j < 4; ++j) nz[j] = qdx_normal_from_bits_t<true>(qdx_bits32(keys.leaf, ctr + j));
#else
            uint32_t bits[4];
            if (MULTI) {                                    // per gene: the key and the counter space of its leaf
                int l = qdx_leaf_of(p.leaves, d);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    while (l + 1 < p.leaves.n && d + j >= p.leaves.off[l + 1]) ++l;
                    const int32_t o = p.leaves.off[l], sz = p.leaves.off[l + 1] - o;
                    bits[j] = qdx_bits32(p.leaves.key[l], (uint64_t)(row0 + rr) * (uint64_t)sz + (uint64_t)(d + j - o));
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) bits[j] = qdx_bits32(keys.leaf, ctr + j);     // four independent Threefry chains
            }
            qdx_normal4_from_bits(bits, nz);
#endif
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float iso = nz[j] * p.iso_sigma;
                float t1 = av[j] + iso;
                float t2 = bv[j] - av[j];
                float t3 = t2 * ln;
                float x = t1 + t3;                                  // mutation_operators.py:211
                if (p.has_min) x = qdx_max_nanprop(x, p.minv);      // :214-215
                if (p.has_max) x = qdx_min_nanprop(x, p.maxv);
                xv[j] = x;
            }
            if (act) *reinterpret_cast<float4*>(tile + rr * DS + (dq << 2)) = make_float4(xv[0], xv[1], xv[2], xv[3]);
        }
        // ---- phase 3 (issued early): tile -> global through the bulk-copy engine ------------------------
        // writers make their generic-proxy stores visible to the async proxy, then the warp syncs, then issue
#if QDX_GEN_PLAIN_STORE      // sanitizer experiment (tools/sanitize.sh): the same rows with ordinary 128-bit stores instead of the bulk-copy engine
        __syncwarp();
        if (out_g && p.store_mode == 0)
            for (int i = lane; i < total_quads; i += 32) {
                const int rr = (int)(((uint32_t)i * qmagic) >> 20), dq = i - rr * q;
                *reinterpret_cast<float4*>(out_g + (row0 + rr) * D + d0 + (dq << 2)) = *reinterpret_cast<const float4*>(tile + rr * DS + (dq << 2));
            }
        if (false) {
#else
        const bool bulk_rows = out_g && p.store_mode == 0;
        if (bulk_rows) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncwarp();
        if (bulk_rows) {
#endif
            if (DS == D) {      // tile is the exact global image of nrows consecutive rows: one bulk copy
                if (lane == 0) qdx_bulk_store(out_g + row0 * D, tile, (uint32_t)(nrows * D * 4));
            } else if (valid) {
                qdx_bulk_store(out_g + row * D + d0, tile + lane * DS, (uint32_t)(dc * 4));
            }
            asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        }
        // ---- phase 2: row-serial scoring, lane = row; sums run left to right from +0 (the spec) ----------
        if (TASK != QDX_TASK_NONE && valid) {
            const float* xr = tile + lane * DS;
            if (TASK == QDX_TASK_ARM) {        // single chunk guaranteed by the launcher
                float sum = 0.0f;
#pragma unroll 5
                for (int d = 0; d < dc; d += 4) {
                    float4 v = *reinterpret_cast<const float4*>(xr + d);
                    float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float x = ARM_CLIP ? qdx_min_nanprop(qdx_max_nanprop(xs[j], 0.0f), 1.0f) : xs[j];
                        sum = sum + x;
                    }
                }
                const float mean = __fdiv_rn(sum, (float)D);
                float sq = 0.0f, th = 0.0f, cs = 0.0f, sn = 0.0f;
                for (int d = 0; d < dc; d += 4) {
                    float4 v = *reinterpret_cast<const float4*>(xr + d);
                    float xs[4] = {v.x, v.y, v.z, v.w};
                    float thj[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {          // the serial chains: variance sum and joint-angle cumsum
                        float x = ARM_CLIP ? qdx_min_nanprop(qdx_max_nanprop(xs[j], 0.0f), 1.0f) : xs[j];
                        float dev = x - mean;
                        sq = sq + dev * dev;
                        th = th + (0x1.921fb6p+2f * x - 0x1.921fb6p+1f);
                        thj[j] = th;
                    }
                    float sj[4], cj[4];
#pragma unroll
                    for (int j = 0; j < 4; j += 2) qdx_sincosf2(thj[j], thj[j + 1], sj[j], cj[j], sj[j + 1], cj[j + 1]);   // two packed pairs (ILP)
#pragma unroll
                    for (int j = 0; j < 4; ++j) { cs = cs + cj[j]; sn = sn + sj[j]; }
                }
                const float fit = -__fsqrt_rn(__fdiv_rn(sq, (float)D));
                const float dx = __fdiv_rn(cs, (float)(2 * D)) + 0.5f;
                const float dy = __fdiv_rn(sn, (float)(2 * D)) + 0.5f;
                out_f[row] = fit;
                reinterpret_cast<float2*>(out_d)[row] = make_float2(dx, dy);
                if (xf) { xf[row] = fit; reinterpret_cast<float2*>(xd)[row] = make_float2(dx, dy); }
                if (GRID_DD != 0) {
                    float xd[QDX_MAX_GRID_DIM] = {dx, dy, 0.0f, 0.0f};
                    const int32_t cell = GRID_DD > 0 ? qdx_grid_cell<(GRID_DD > 0 ? GRID_DD : 1)>(xd, p.grid, s_axes, p.centroids, p.K)
                                                     : qdx_index_cell<(GRID_DD < 0 ? -GRID_DD : 1)>(xd, p.cvt);
                    if (p.out_cell) p.out_cell[row] = cell;
                    if (p.offer && qdx_offer(p.ws, p.K, p.rep_f, cell, fit, p.idx_base + (uint32_t)row, p.first_wins) && p.store_mode == 1)
                        qdx_store_row(out_g + row * D, xr, dc, p.out_xchg);
                }
            } else {
                for (int d = 0; d < dc; d += 4) {
                    float4 v = *reinterpret_cast<const float4*>(xr + d);
                    float xs[4] = {v.x, v.y, v.z, v.w};
                    float term[4], xx[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        xx[j] = xs[j] * 10.0f - 5.0f;
                        term[j] = xx[j] * xx[j];
                    }
                    if (TASK == QDX_TASK_RASTRIGIN) {
#pragma unroll
                        for (int j = 0; j < 4; j += 2) {       // two packed pairs
                            float s0, c0, s1, c1;
                            qdx_sincosf2(0x1.921fb6p+2f * xx[j], 0x1.921fb6p+2f * xx[j + 1], s0, c0, s1, c1);
                            term[j] = term[j] - 10.0f * c0;
                            term[j + 1] = term[j + 1] - 10.0f * c1;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc0 = acc0 + term[j];
                }
                if (ch == 0) {
                    for (int j = 0; j < p.desc_dim; ++j) out_d[row * p.desc_dim + j] = xr[j];   // desc = first genes
                    if (xd) for (int j = 0; j < p.desc_dim; ++j) xd[row * p.desc_dim + j] = xr[j];
                }
                if (ch == nchunks - 1) {
                    float f = acc0;
                    if (TASK == QDX_TASK_RASTRIGIN) f = (float)(10.0 * (double)D) + f;
                    const float fit = -f;
                    out_f[row] = fit;
                    if (xf) xf[row] = fit;
                    if (GRID_DD != 0) {
                        float xd[QDX_MAX_GRID_DIM];
#pragma unroll
                        for (int j = 0; j < (GRID_DD > 0 ? GRID_DD : -GRID_DD); ++j) xd[j] = out_d[row * p.desc_dim + j];
                        const int32_t cell = GRID_DD > 0 ? qdx_grid_cell<(GRID_DD > 0 ? GRID_DD : 1)>(xd, p.grid, s_axes, p.centroids, p.K)
                                                         : qdx_index_cell<(GRID_DD < 0 ? -GRID_DD : 1)>(xd, p.cvt);
                        if (p.out_cell) p.out_cell[row] = cell;
                        if (p.offer && qdx_offer(p.ws, p.K, p.rep_f, cell, fit, p.idx_base + (uint32_t)row, p.first_wins) && p.store_mode == 1)
                            qdx_store_row(out_g + row * D, xr, dc, p.out_xchg);
                    }
                }
            }
        }
        // the tile is rewritten by the next chunk / tile: wait until the bulk engine has finished READING it
        if (out_g && p.store_mode == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
        __syncwarp();
    }
    }   // tiles of this warp
    if (p.out_xchg && out_g && p.store_mode == 0) {
        // peers read these rows straight out of this rank's memory once its arrival flag is up: the bulk copies must have
        // COMPLETED (not merely been read out of shared memory).  Ordinary stores (fitness / descriptors of every row, fired
        // rows) need nothing here: they are ordered before the flags by the CTA barrier, the ticket and the publisher's
        // system-scope fence (a fence per warp at this point cost ~20 us per launch: MEMBAR.SYS serialises within an SM).
        asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
        __threadfence_system();
    }
}

// =====================================================================================================
// standalone scoring:  (B, D) genotypes -> fitness (B,), descriptors (B, Dd)
// =====================================================================================================
// NOISY (arm only): noisy_arm_scoring_function (qdax/tasks/arm.py:53-81) -- normal(p_sub, params.shape) * params_variance is
// added while the rows are staged, normal(f_sub) * fit_variance / normal(d_sub) * desc_variance to the results.
struct QdxNoise { QdxKey kf, kd, kp; float fit_var, desc_var, params_var; };
template <int TASK, bool NOISY = false>
__global__ void __launch_bounds__(128) qdx_score_kernel(const float* __restrict__ g, int64_t B, int32_t D, int32_t desc_dim,
                                                        float* __restrict__ out_f, float* __restrict__ out_d, const QdxNoise nz) {
    // One warp per 32-row tile; genotype chunks are staged row-major in shared memory with a coalesced
    // cooperative copy, then consumed row-serially (lane = row) in the canonical left-to-right order.
    extern __shared__ __align__(128) float s_tiles[];
    constexpr int DC = 64;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* tile = s_tiles + (size_t)warp * 32 * (DC + 1);
    const int64_t row0 = ((int64_t)blockIdx.x * 4 + warp) * 32;
    if (row0 >= B) return;
    const int64_t row = row0 + lane;
    const bool valid = row < B;
    const int nrows = (B - row0) < 32 ? (int)(B - row0) : 32;
    const int npass = (TASK == QDX_TASK_ARM) ? 2 : 1;
    float sum = 0.0f, mean = 0.0f, sq = 0.0f, th = 0.0f, cs = 0.0f, sn = 0.0f, acc = 0.0f;
    for (int pass = 0; pass < npass; ++pass) {
        for (int d0 = 0; d0 < D; d0 += DC) {
            const int dc = (D - d0) < DC ? (D - d0) : DC;
            __syncwarp();
            for (int i = lane; i < nrows * dc; i += 32) {       // coalesced: consecutive lanes -> consecutive genes
                const int r = i / dc, d = i - r * dc;
                float v = g[(row0 + r) * D + d0 + d];
                if (NOISY) {
                    const float t = qdx_normal_from_bits(qdx_bits32(nz.kp, (uint64_t)(row0 + r) * (uint64_t)D + (uint64_t)(d0 + d))) * nz.params_var;
                    v = v + t;
                }
                tile[r * (DC + 1) + d] = v;
            }
            __syncwarp();
            if (!valid) continue;
            const float* xr = tile + lane * (DC + 1);
            if (TASK == QDX_TASK_ARM) {
                if (pass == 0) {
                    for (int d = 0; d < dc; ++d) {
                        float x = qdx_min_nanprop(qdx_max_nanprop(xr[d], 0.0f), 1.0f);
                        sum = sum + x;
                    }
                } else {
                    for (int d = 0; d < dc; ++d) {
                        float x = qdx_min_nanprop(qdx_max_nanprop(xr[d], 0.0f), 1.0f);
                        float dev = x - mean;
                        float dd = dev * dev;
                        float ang = 0x1.921fb6p+2f * x - 0x1.921fb6p+1f;
                        float s, c;
                        sq = sq + dd; th = th + ang; qdx_sincosf(th, s, c); cs = cs + c; sn = sn + s;
                    }
                }
            } else {
                for (int d = 0; d < dc; ++d) {
                    float x = xr[d] * 10.0f - 5.0f;
                    float term = x * x;
                    if (TASK == QDX_TASK_RASTRIGIN) { float s, c; qdx_sincosf(0x1.921fb6p+2f * x, s, c); term = term - 10.0f * c; }
                    acc = acc + term;
                    if (d0 + d < desc_dim) out_d[row * desc_dim + d0 + d] = xr[d];
                }
            }
        }
        if (TASK == QDX_TASK_ARM && pass == 0) mean = __fdiv_rn(sum, (float)D);
    }
    if (!valid) return;
    if (TASK == QDX_TASK_ARM) {
        float f = -__fsqrt_rn(__fdiv_rn(sq, (float)D));
        float dx = __fdiv_rn(cs, (float)(2 * D)) + 0.5f, dy = __fdiv_rn(sn, (float)(2 * D)) + 0.5f;
        if (NOISY) {
            const float nf = qdx_normal_from_bits(qdx_bits32(nz.kf, (uint64_t)row)) * nz.fit_var;
            const float nx = qdx_normal_from_bits(qdx_bits32(nz.kd, 2ull * (uint64_t)row)) * nz.desc_var;
            const float ny = qdx_normal_from_bits(qdx_bits32(nz.kd, 2ull * (uint64_t)row + 1ull)) * nz.desc_var;
            f = f + nf; dx = dx + nx; dy = dy + ny;
        }
        out_f[row] = f;
        out_d[row * 2 + 0] = dx;
        out_d[row * 2 + 1] = dy;
    } else {
        float f = acc;
        if (TASK == QDX_TASK_RASTRIGIN) f = (float)(10.0 * (double)D) + f;
        out_f[row] = -f;
    }
}

// =====================================================================================================
// brute-force cell assignment (+ optional offer)
// =====================================================================================================
// Each thread owns one descriptor row; centroids stream through shared memory in tiles (broadcast LDS).
// Inner loop per 16-centroid group keeps only a running minimum (FMNMX); the group is rescanned for the
// FIRST index only when it improves the row's best, which happens O(log K) times per row.
template <int DD>
__global__ void __launch_bounds__(256) qdx_cells_bf_kernel(const float* __restrict__ desc, int64_t B,
                                                           const float* __restrict__ cent, int64_t K,
                                                           int32_t* __restrict__ cells, void* ws, const float* rep_f,
                                                           const float* __restrict__ fit, int32_t offer,
                                                           uint32_t idx_base, int32_t first_wins) {
    extern __shared__ __align__(16) float s_cent[];
    constexpr int TILE = 2048;                  // centroids per shared-memory tile
    constexpr int GROUP = 16;
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = row < B;
    float x[DD];
    bool finite = true;
#pragma unroll
    for (int d = 0; d < DD; ++d) { x[d] = valid ? desc[row * DD + d] : 0.0f; finite = finite && (fabsf(x[d]) <= 3.40282347e+38f); }
    float best = INFINITY; int32_t bk = 0;
    for (int64_t k0 = 0; k0 < K; k0 += TILE) {
        const int n = (K - k0) < TILE ? (int)(K - k0) : TILE;
        __syncthreads();
        for (int i = threadIdx.x; i < n * DD; i += blockDim.x) s_cent[i] = cent[k0 * DD + i];
        for (int i = n * DD + threadIdx.x; i < ((n + GROUP - 1) / GROUP) * GROUP * DD; i += blockDim.x) s_cent[i] = INFINITY;  // pad
        __syncthreads();
        if (!valid || !finite) continue;
        for (int g0 = 0; g0 < n; g0 += GROUP) {
            float gmin = INFINITY;
#pragma unroll
            for (int t = 0; t < GROUP; ++t) {
                float acc;
#pragma unroll
                for (int d = 0; d < DD; ++d) { float df = x[d] - s_cent[(g0 + t) * DD + d]; float s = df * df; acc = d ? acc + s : s; }
                gmin = fminf(gmin, acc);       // padded centroids give inf/NaN-free inf; fminf drops NaN (none here)
            }
            if (gmin < best) {
                best = gmin;
                for (int t = 0; t < GROUP; ++t) {
                    float acc;
#pragma unroll
                    for (int d = 0; d < DD; ++d) { float df = x[d] - s_cent[(g0 + t) * DD + d]; float s = df * df; acc = d ? acc + s : s; }
                    if (acc == gmin) { bk = (int32_t)(k0 + g0 + t); break; }
                }
            }
        }
    }
    if (!valid) return;
    // non-finite descriptor: every distance is inf or NaN -> argmin = 0 (first inf / first NaN), centroids finite
    cells[row] = bk;
    if (offer) qdx_offer(ws, K, rep_f, bk, fit[row], idx_base + (uint32_t)row, first_wins);
}

// generic descriptor dimension (runtime Dd): same algorithm, descriptor row kept in shared memory
__global__ void __launch_bounds__(128) qdx_cells_bf_generic_kernel(const float* __restrict__ desc, int64_t B, int32_t Dd,
                                                                   const float* __restrict__ cent, int64_t K,
                                                                   int32_t* __restrict__ cells, void* ws, const float* rep_f,
                                                                   const float* __restrict__ fit, int32_t offer,
                                                                   uint32_t idx_base, int32_t first_wins) {
    // one warp per descriptor row: lanes split the centroids, each computes full sequential-over-d distances
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= B) return;
    const float* x = desc + row * Dd;
    float best = INFINITY; int64_t bk = 0x7fffffff; bool any_nan = false; int64_t nan_k = 0x7fffffff;
    for (int64_t k = lane; k < K; k += 32) {
        const float* c = cent + k * Dd;
        float acc = 0.0f;
        for (int d = 0; d < Dd; ++d) { float df = x[d] - c[d]; float s = df * df; acc = d ? acc + s : s; }
        if (acc != acc) { if (!any_nan) { any_nan = true; nan_k = k; } }
        else if (acc < best) { best = acc; bk = k; }
    }
    // lexicographic (dist, k) min across lanes; NaN wins with its smallest index
    for (int o = 16; o > 0; o >>= 1) {
        float ob = __shfl_xor_sync(0xffffffffu, best, o);
        long long ok = __shfl_xor_sync(0xffffffffu, (long long)bk, o);
        long long on = __shfl_xor_sync(0xffffffffu, (long long)nan_k, o);
        if (ob < best || (ob == best && ok < bk)) { best = ob; bk = ok; }
        if (on < nan_k) nan_k = on;
    }
    if (lane == 0) {
        int32_t cell = (nan_k != 0x7fffffff) ? (int32_t)nan_k : (bk == 0x7fffffff ? 0 : (int32_t)bk);
        cells[row] = cell;
        if (offer) qdx_offer(ws, K, rep_f, cell, fit[row], idx_base + (uint32_t)row, first_wins);
    }
}

// standalone grid assignment (get_cells_indices on a grid tessellation) + optional offer
template <int DD>
__global__ void __launch_bounds__(256) qdx_cells_grid_kernel(const float* __restrict__ desc, int64_t B, const QdxGrid grid,
                                                             const float* __restrict__ cent, int64_t K,
                                                             int32_t* __restrict__ cells, void* ws, const float* rep_f,
                                                             const float* __restrict__ fit, int32_t offer,
                                                             uint32_t idx_base, int32_t first_wins) {
    __shared__ float s_axes[QDX_MAX_AXES];
    for (int i = threadIdx.x; i < grid.total_axes; i += blockDim.x) s_axes[i] = grid.axes[i];
    __syncthreads();
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= B) return;
    float x[DD];
#pragma unroll
    for (int d = 0; d < DD; ++d) x[d] = desc[row * DD + d];
    const int32_t cell = qdx_grid_cell<DD>(x, grid, s_axes, cent, K);
    cells[row] = cell;
    if (offer) qdx_offer(ws, K, rep_f, cell, fit[row], idx_base + (uint32_t)row, first_wins);
}

// offer only: cells already known (tell / add with injected cells, or after an all-gather)
__global__ void __launch_bounds__(256) qdx_offer_kernel(const int32_t* __restrict__ cells, const float* __restrict__ fit,
                                                        int64_t B, int64_t K, void* ws, const float* rep_f,
                                                        uint32_t idx_base, int32_t first_wins) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= B) return;
    const int32_t c = cells[row];
    if (c < 0 || c >= K) { qdx_set_error(ws, QDX_ERR_BAD_CELL); return; }
    qdx_offer(ws, K, rep_f, c, fit[row], idx_base + (uint32_t)row, first_wins);
}

// =====================================================================================================
// commit: winners -> repertoire rows, reset the key table, QD metrics by the last CTA
// =====================================================================================================
__global__ void __launch_bounds__(256) qdx_commit_kernel(void* ws_raw, int64_t K, int32_t D, int32_t Dd,
                                                         const float* __restrict__ off_g, const float* __restrict__ off_f,
                                                         const float* __restrict__ off_d, uint32_t idx_base, int64_t B,
                                                         int32_t first_wins, float* __restrict__ rep_g,
                                                         float* __restrict__ rep_f, float* __restrict__ rep_d,
                                                         float qd_offset, float* __restrict__ metrics_out,
                                                         int32_t* __restrict__ added_cells, int32_t mode) {
    // mode 0: commit winners whose offspring rows are in off_* (index = global idx - idx_base), reset keys, metrics
    // mode 1: stage -- copy only the winners owned by [idx_base, idx_base + B) into rep_* (= staging rows by
    //         cell), keep the key table, no metrics                       (multi-GPU winners-only exchange)
    // mode 2: apply -- off_* are staging rows indexed by CELL; reset keys, metrics
    QdxWorkspace* ws = (QdxWorkspace*)ws_raw;
    if (mode == 2 && *(volatile int32_t*)&ws->error == QDX_ERR_PEER_TIMEOUT) {       // see qdx_commit_stream_kernel
        if (metrics_out && blockIdx.x == 0 && threadIdx.x < 4) metrics_out[threadIdx.x] = __int_as_float(0x7fc00000);
        return;
    }
    unsigned long long* keytab = qdx_ws_keytab(ws_raw, K);
    const int lane = threadIdx.x & 31;
    // 128-bit row copies need 16-byte aligned rows: D % 4 == 0 AND 16-byte aligned bases (an offset view of a larger buffer
    // reaches this fallback precisely because it is not)
    const bool vec_rows = (D & 3) == 0 && ((((uintptr_t)off_g) | ((uintptr_t)rep_g)) & 15u) == 0;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int added = 0, newly = 0;
    double sum = 0.0; float mx = -INFINITY; int cnt = 0; int nan = 0;      // metrics of this warp's cells (lane 0)
    // One warp per CELL (grid-stride): every resident warp has a winner row in flight, which is what lets the row
    // traffic approach the HBM roofline when rows are large (K/32-warp parallelism measured 35 % of peak).
    unsigned long long key_next = (lane == 0 && warp_global < K) ? keytab[warp_global] : 0ull;     // lane 0 owns the table entries
    for (int64_t cell = warp_global; cell < K; cell += nwarps) {
        const unsigned long long key = __shfl_sync(0xffffffffu, key_next, 0);
        if (lane == 0 && cell + nwarps < K) key_next = keytab[cell + nwarps];                        // prefetch the next entry
        float fcell;
        const bool win = key != 0ull && !qdx_key_is_nan(key);                // NaN-poisoned cells accept nobody
        int64_t i = -1;
        if (win) {
            i = (int64_t)qdx_key_index(key, first_wins) - (int64_t)idx_base;
            if (mode == 2) i = cell;
            else if (i < 0 || i >= B) { if (mode == 0 && lane == 0) qdx_set_error(ws, QDX_ERR_BAD_INDEX); i = -1; }
        }
        if (i >= 0) {
            const float* srow = off_g + i * D; float* drow = rep_g + cell * D;
            if (vec_rows) {
                const float4* s4 = reinterpret_cast<const float4*>(srow); float4* d4 = reinterpret_cast<float4*>(drow);
                const int nq = D >> 2;
                int q = lane;
                for (; q + 96 < nq; q += 128) {
                    const float4 v0 = __ldcs(s4 + q), v1 = __ldcs(s4 + q + 32), v2 = __ldcs(s4 + q + 64), v3 = __ldcs(s4 + q + 96);
                    d4[q] = v0; d4[q + 32] = v1; d4[q + 64] = v2; d4[q + 96] = v3;
                }
                for (; q < nq; q += 32) d4[q] = __ldg(s4 + q);
            } else {
                for (int d = lane; d < D; d += 32) drow[d] = srow[d];
            }
            for (int d = lane; d < Dd; d += 32) rep_d[cell * Dd + d] = off_d[i * Dd + d];
            fcell = off_f[i];
            if (lane == 0) {
                if (__ldcg(rep_f + cell) == -INFINITY) ++newly;     // the occupied-cell list changes: rescan at the end
                rep_f[cell] = fcell;
                if (added_cells) added_cells[cell] = (int32_t)i;
            }
            ++added;
        } else {
            fcell = (mode == 1) ? -INFINITY : __ldcg(rep_f + cell);
        }
        if (key != 0ull && mode != 1 && lane == 0) keytab[cell] = 0ull;
        if (fcell != -INFINITY) { sum += (double)fcell; ++cnt; }
        if (fcell != fcell) nan = 1; else if (fcell > mx) mx = fcell;
    }
    // ---- metrics: CTAs publish partials of their cells, the last CTA to finish sums them in CTA order (deterministic)
    if (mode == 1) return;
    __shared__ double s_sum[8]; __shared__ float s_max[8]; __shared__ int s_cnt[8]; __shared__ int s_nan[8]; __shared__ int s_add[8];
    __shared__ int s_new[8];
    __shared__ bool s_last;
    const int wid = threadIdx.x >> 5;
    if (lane == 0) { s_sum[wid] = sum; s_max[wid] = mx; s_cnt[wid] = cnt; s_nan[wid] = nan; s_add[wid] = added; s_new[wid] = newly; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0; float m = -INFINITY; int n = 0, nn = 0, a = 0, nw = 0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) { s += s_sum[w]; m = fmaxf(m, s_max[w]); n += s_cnt[w]; nn |= s_nan[w]; a += s_add[w]; nw += s_new[w]; }
        ws->part_sum[blockIdx.x] = s; ws->part_max[blockIdx.x] = m; ws->part_cnt[blockIdx.x] = n; ws->part_nan[blockIdx.x] = nn;
        ws->part_add[blockIdx.x] = a; ws->part_new[blockIdx.x] = nw;
        __threadfence();
        const unsigned t = atomicAdd(&ws->ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // fixed-shape reduction of the per-CTA partials: thread t sums partials t, t+256, ... then warps, then thread 0
    {
        double s = 0.0; float m = -INFINITY; int n = 0, nn = 0, a = 0, nw = 0;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
            s += *(volatile double*)&ws->part_sum[b]; m = fmaxf(m, *(volatile float*)&ws->part_max[b]);
            n += *(volatile int32_t*)&ws->part_cnt[b]; nn |= *(volatile int32_t*)&ws->part_nan[b]; a += *(volatile int32_t*)&ws->part_add[b];
            nw += *(volatile int32_t*)&ws->part_new[b];
        }
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o); m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            n += __shfl_xor_sync(0xffffffffu, n, o); nn |= __shfl_xor_sync(0xffffffffu, nn, o); a += __shfl_xor_sync(0xffffffffu, a, o);
            nw += __shfl_xor_sync(0xffffffffu, nw, o);
        }
        __syncthreads();
        if (lane == 0) { s_sum[wid] = s; s_max[wid] = m; s_cnt[wid] = n; s_nan[wid] = nn; s_add[wid] = a; s_new[wid] = nw; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double s = 0.0; float m = -INFINITY; int n = 0, nn = 0, a = 0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) { s += s_sum[w]; m = fmaxf(m, s_max[w]); n += s_cnt[w]; nn |= s_nan[w]; a += s_add[w]; }
        float out[4];
        out[0] = (float)s + qd_offset * (float)n;                 // qd_score   (metrics.py:92-93)
        out[1] = nn ? NAN : m;                                     // max_fitness (:95)
        out[2] = 100.0f * __fdiv_rn((float)n, (float)K);           // coverage   (:94)
        out[3] = (float)a;                                         // offspring inserted by this call
        for (int j = 0; j < 4; ++j) { ws->metrics[j] = out[j]; if (metrics_out) metrics_out[j] = out[j]; }
        ws->ticket = 0u;
        if (mode == 2 && ws->xchg_nranks > 0) {                              // next generation: the other key table
            uint32_t* ep = (uint32_t*)((char*)ws->xchg_peer[ws->xchg_rank] + QDX_XCHG_EPOCH_OFFSET);
            *ep = *ep + 1u;
        }
    }
    // ---- the repertoire is final: leave the NEXT generation's parent selection ready (no prepare launch).  The
    // occupied-cell list only changes when a cell turned from empty to occupied (never in steady state), or when the
    // workspace has not seen this repertoire yet (M mismatch: e.g. first commit after init).
    __shared__ int32_t s_scan[33];
    int total_new = 0, total_cnt = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) { total_new += s_new[w]; total_cnt += s_cnt[w]; }
#ifndef QDX_COMMIT_NO_TAIL     // A/B switch for timing only (selection tables would go stale)
    if (total_new != 0 || total_cnt != ws->sel.M || ws->sel.nseg <= 0) qdx_cta_occupancy_scan(rep_f, K, ws_raw, s_scan);
#endif
}

// =====================================================================================================
// peer-memory exchange (multi-GPU): push the local per-cell bests into every peer's key table over NVLink
// =====================================================================================================
// DistributedMAPElites (distributed_map_elites.py:133-146) needs, per cell, the best offspring over ALL ranks.  Every
// offer that improves its cell's local best is max-merged into every peer's key table by the offering thread itself
// (qdx_offer: system-scope 64-bit atomicMax straight into the peer's HBM, mapped with cudaIpc, NVLink 5 / NVSwitch; no
// staging, no collective library).  When the last offering warp of a rank is done it publishes "rank `me`, epoch e has
// landed" in every peer's flag array (release, system scope; qdx_xchg_warp_done in the generate kernel, or the
// stand-alone kernel below); consumers acquire-spin on their LOCAL flags (qdx_elect_kernel).
// Stand-alone publication for generations whose offers come from a separate cells kernel (non-fused tessellations):
// stream order guarantees every offering kernel (and the system-scope fences of its pushes) has completed.
__global__ void qdx_publish_kernel(void* ws_raw, int64_t K, const QdxGenKeys keys) {
    if (threadIdx.x == 0 && blockIdx.x == 0) qdx_xchg_publish((QdxWorkspace*)ws_raw, K, keys);
}

// =====================================================================================================
// elect: regenerate + score the elected winners (multi-GPU "regen" / "p2p" exchanges)
// =====================================================================================================
// After the key tables have been max-merged (NCCL all-reduce, or the peer-memory push above) every rank knows, per
// cell, the global index of the winning offspring and (tail slots) every rank's generation keys.  The RNG is
// counter-based and the repertoire replicated, so each rank recomputes the winner's genotype bit for bit as its owner
// produced it, scores it, and leaves (genotype, fitness, descriptor) in per-cell staging rows for qdx_commit(mode 2).
// One warp per elected cell.  Scoring keeps the spec's sequential float32 sums (lane 0) but spreads the independent
// per-gene work (sincos, rastrigin terms) over the 32 lanes.
template <int TASK>
__global__ void __launch_bounds__(256) qdx_elect_kernel(void* ws_raw, int64_t K, int32_t D, int32_t desc_dim, int64_t B_dev,
                                                        int32_t nranks, const float* __restrict__ rep_g, float iso_sigma,
                                                        float line_sigma, int32_t has_min, float minv, int32_t has_max,
                                                        float maxv, int32_t first_wins, float* __restrict__ stage_g,
                                                        float* __restrict__ stage_f, float* __restrict__ stage_d,
                                                        int32_t wait_peers) {
    __shared__ QdxSeg s_seg[QDX_MAX_SEG];
    __shared__ float s_last[QDX_MAX_SEG];
    __shared__ float s_buf[8][3][128];
    QdxWorkspace* ws = (QdxWorkspace*)ws_raw;
#if QDX_XCHG_TRACE
    const unsigned long long tr_t0 = qdx_now();
#endif
    // A peer that never arrives must not be elected around: on timeout the sticky error is raised (host mirror included)
    // and qdx_commit(mode 2) of this generation does nothing -- the repertoire, the key tables and the epoch stay as they
    // are, and the host raises QDX_ERR_PEER_TIMEOUT on its next call instead of letting the replicas diverge silently.
    if (wait_peers && threadIdx.x < ws->xchg_nranks) {     // acquire-spin on the LOCAL arrival flags (bounded: wait_peers ms)
        const unsigned long long* flag = (const unsigned long long*)ws->xchg_peer[ws->xchg_rank] + threadIdx.x;
        const uint32_t want = *(const uint32_t*)((const char*)ws->xchg_peer[ws->xchg_rank] + QDX_XCHG_EPOCH_OFFSET) + 1u;
        unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while ((int32_t)((uint32_t)qdx_ld_acquire_sys(flag) - want) < 0) {
            unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > (unsigned long long)wait_peers * 1000000ull) { qdx_set_error(ws, QDX_ERR_PEER_TIMEOUT); break; }
            __nanosleep(64);
        }
    }
    const int nseg = ws->sel.nseg;
    for (int i = threadIdx.x; i < nseg; i += blockDim.x) { s_seg[i] = ws->sel.seg[i]; s_last[i] = ws->sel.last[i]; }
    __syncthreads();
#if QDX_XCHG_TRACE
    const unsigned long long tr_t1 = qdx_now();
#endif
    if (nseg <= 0) return;
    if (wait_peers && *(volatile int32_t*)&ws->error == QDX_ERR_PEER_TIMEOUT) return;      // this CTA (or an earlier generation) timed out
    const unsigned long long* keytab = qdx_ws_keytab(ws_raw, K);
    const int32_t* __restrict__ occ = qdx_ws_occ(ws_raw);
    const float total = ws->sel.total;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* s_x = s_buf[wid][0]; float* s_a = s_buf[wid][1]; float* s_b = s_buf[wid][2];
    // The CTA owns a contiguous block of cells (the grid is one wave): coalesced key loads, the elected cells of
    // each 256-cell slab are compacted into a shared list, and the 8 warps take them round-robin -- with ~10 % of the
    // cells elected, every warp gets real work instead of one warp per (mostly empty) cell.
    __shared__ int32_t s_list[256];
    __shared__ unsigned long long s_key[256];
    __shared__ int32_t s_n;
    const int64_t per_cta = ((K + gridDim.x - 1) / gridDim.x + 31) / 32 * 32;
    const int64_t c_lo = (int64_t)blockIdx.x * per_cta, c_hi = c_lo + per_cta < K ? c_lo + per_cta : K;
    for (int64_t slab = c_lo; slab < c_hi; slab += 256) {
        __syncthreads();
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        {
            const int64_t c = slab + threadIdx.x;
            const unsigned long long k = c < c_hi ? __ldcg(keytab + c) : 0ull;
            const bool el = k != 0ull && !qdx_key_is_nan(k);
            const unsigned b = __ballot_sync(0xffffffffu, el);
            int base = 0;
            if (lane == 0 && b) base = atomicAdd(&s_n, __popc(b));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (el) { const int pos = base + __popc(b & ((1u << lane) - 1u)); s_list[pos] = (int32_t)(c - slab); s_key[pos] = k; }
        }
        __syncthreads();
        const int n_el = s_n;
    for (int e = wid; e < n_el; e += 8) {
        const int64_t c = slab + s_list[e];
        const unsigned long long key = s_key[e];
        const uint32_t idx = qdx_key_index(key, first_wins);
        const int64_t r = idx / B_dev, i = idx % B_dev;
        if (r >= nranks) continue;
        const unsigned long long* slot = keytab + K + 8 * r;
        const QdxKey sel1{(uint32_t)__ldcg(slot + 0), (uint32_t)__ldcg(slot + 1)}, sel2{(uint32_t)__ldcg(slot + 2), (uint32_t)__ldcg(slot + 3)};
        const QdxKey kline{(uint32_t)__ldcg(slot + 4), (uint32_t)__ldcg(slot + 5)}, kleaf{(uint32_t)__ldcg(slot + 6), (uint32_t)__ldcg(slot + 7)};
        const float u1 = qdx_unit_float(qdx_bits32(sel1, (uint64_t)i)), u2 = qdx_unit_float(qdx_bits32(sel2, (uint64_t)i));
        const int32_t p1 = occ[qdx_sel_rank(s_seg, s_last, nseg, total * (1.0f - u1)) - 1];
        const int32_t p2 = occ[qdx_sel_rank(s_seg, s_last, nseg, total * (1.0f - u2)) - 1];
        const float ln = qdx_normal_from_bits(qdx_bits32(kline, (uint64_t)i)) * line_sigma;
        float acc = 0.0f;                                   // rastrigin / sphere running sum (lane 0)
        for (int d0 = 0; d0 < D; d0 += 128) {
            const int dc = (D - d0) < 128 ? (D - d0) : 128;
            __syncwarp();
            for (int d = lane; d < dc; d += 32) {
                const int dg = d0 + d;
                const float a = __ldg(rep_g + (int64_t)p1 * D + dg), b = __ldg(rep_g + (int64_t)p2 * D + dg);
                const float iso = qdx_normal_from_bits(qdx_bits32(kleaf, (uint64_t)i * (uint64_t)D + (uint64_t)dg)) * iso_sigma;
                float t1 = a + iso, t2 = b - a, t3 = t2 * ln;
                float x = t1 + t3;                                  // mutation_operators.py:211
                if (has_min) x = qdx_max_nanprop(x, minv);
                if (has_max) x = qdx_min_nanprop(x, maxv);
                stage_g[c * D + dg] = x;
                s_x[d] = x;
            }
            __syncwarp();
            if (TASK == QDX_TASK_ARM) {                 // arm.py:27-44; D <= 128 guaranteed by the launcher
                float mean = 0.0f;
                if (lane == 0) {
                    float sum = 0.0f;
                    for (int d = 0; d < dc; ++d) sum = sum + qdx_min_nanprop(qdx_max_nanprop(s_x[d], 0.0f), 1.0f);
                    mean = __fdiv_rn(sum, (float)D);
                    float sq = 0.0f, th = 0.0f;
                    for (int d = 0; d < dc; ++d) {
                        const float x = qdx_min_nanprop(qdx_max_nanprop(s_x[d], 0.0f), 1.0f);
                        const float dev = x - mean;
                        sq = sq + dev * dev;
                        th = th + (0x1.921fb6p+2f * x - 0x1.921fb6p+1f);
                        s_a[d] = th;
                    }
                    stage_f[c] = -__fsqrt_rn(__fdiv_rn(sq, (float)D));
                }
                __syncwarp();
                for (int d = lane; d < dc; d += 32) { float sn, cs; qdx_sincosf(s_a[d], sn, cs); s_a[d] = cs; s_b[d] = sn; }
                __syncwarp();
                if (lane == 0) {
                    float cs = 0.0f, sn = 0.0f;
                    for (int d = 0; d < dc; ++d) { cs = cs + s_a[d]; sn = sn + s_b[d]; }
                    stage_d[c * 2 + 0] = __fdiv_rn(cs, (float)(2 * D)) + 0.5f;
                    stage_d[c * 2 + 1] = __fdiv_rn(sn, (float)(2 * D)) + 0.5f;
                }
            } else if (TASK != QDX_TASK_NONE) {         // standard_functions.py:9-48
                for (int d = lane; d < dc; d += 32) {
                    const float x = s_x[d] * 10.0f - 5.0f;
                    float term = x * x;
                    if (TASK == QDX_TASK_RASTRIGIN) { float sn, cs; qdx_sincosf(0x1.921fb6p+2f * x, sn, cs); term = term - 10.0f * cs; }
                    s_a[d] = term;
                    if (d0 + d < desc_dim) stage_d[c * desc_dim + d0 + d] = s_x[d];
                }
                __syncwarp();
                if (lane == 0) for (int d = 0; d < dc; ++d) acc = acc + s_a[d];
            }
        }
        if (TASK != QDX_TASK_ARM && TASK != QDX_TASK_NONE && lane == 0) {
            float f = acc;
            if (TASK == QDX_TASK_RASTRIGIN) f = (float)(10.0 * (double)D) + f;
            stage_f[c] = -f;
        }
    }
    }
#if QDX_XCHG_TRACE
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(&g_xchg_trace[0], 1ull); atomicAdd(&g_xchg_trace[1], tr_t1 - tr_t0); atomicAdd(&g_xchg_trace[2], qdx_now() - tr_t0);
    }
#endif
}

// =====================================================================================================
// small standalone ops of the preserved Python surface
// =====================================================================================================
// UniformSelector.select index stream (uniform_selector.py:48-55): key = key handed to select()
__global__ void __launch_bounds__(256) qdx_select_kernel(void* ws_raw, QdxKey key, int64_t num, int32_t* __restrict__ out) {
    __shared__ QdxSeg s_seg[QDX_MAX_SEG];
    __shared__ float s_last[QDX_MAX_SEG];
    const QdxWorkspace* ws = (const QdxWorkspace*)ws_raw;
    const int nseg = ws->sel.nseg;
    for (int i = threadIdx.x; i < nseg; i += blockDim.x) { s_seg[i] = ws->sel.seg[i]; s_last[i] = ws->sel.last[i]; }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num) return;
    if (nseg <= 0) {                           // empty repertoire (p = 0/0 in the reference): flag it, keep indices in range
        if (i == 0) qdx_set_error(ws_raw, QDX_ERR_EMPTY_REPERTOIRE);
        out[i] = 0; return;
    }
    const QdxKey sub = qdx_split(key, 1);
    const float u = qdx_unit_float(qdx_bits32(sub, (uint64_t)i));
    out[i] = qdx_ws_occ(ws_raw)[qdx_sel_rank(s_seg, s_last, nseg, ws->sel.total * (1.0f - u)) - 1];
}

// UniformSelector(select_with_replacement=False): jax.random.choice(subkey, arange(K), (n,), p=p, replace=False) is the Gumbel
// top-k trick: g_c = -log(-log(u_c)) + log(p_c), u = uniform(subkey, (K,), minval=tiny, maxval=1); the n largest g, equal
// values by ascending cell.  Empty cells have p = 0, log = -inf: they come last (and are returned when n exceeds the number
// of occupied cells, as in the reference).
__global__ void __launch_bounds__(256) qdx_gumbel_kernel(const float* __restrict__ rep_f, int64_t K, const void* ws_raw, QdxKey key,
                                                         float* __restrict__ g) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= K) return;
    const QdxWorkspace* ws = (const QdxWorkspace*)ws_raw;
    const int32_t M = ws->sel.M;
    if (M <= 0) { if (c == 0) qdx_set_error((void*)ws_raw, QDX_ERR_EMPTY_REPERTOIRE); g[c] = -INFINITY; return; }
    const QdxKey sub = qdx_split(key, 1);                                           // uniform_selector.py:48
    const float tiny = 0x1p-126f;
    float u = qdx_unit_float(qdx_bits32(sub, (uint64_t)c)) * 1.0f + tiny;            // f * (maxval - minval) + minval
    u = u < tiny ? tiny : u;
    const float gum = -qdx_logf(-qdx_logf(u));
    const float logq = qdx_logf(__fdiv_rn(1.0f, (float)M));
    g[c] = (__ldg(rep_f + c) != -INFINITY) ? gum + logq : -INFINITY;
}
// rank of every cell in the descending order of g (ties: ascending cell) by counting, candidate tiles through shared memory
__global__ void __launch_bounds__(256) qdx_topk_rank_kernel(const float* __restrict__ g, int64_t K, int64_t n, int32_t* __restrict__ out) {
    __shared__ float s_g[1024];
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const float mine = c < K ? g[c] : 0.0f;
    int64_t rank = 0;
    for (int64_t j0 = 0; j0 < K; j0 += 1024) {
        __syncthreads();
        for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_g[i] = j0 + i < K ? g[j0 + i] : -INFINITY;
        __syncthreads();
        const int m = (K - j0) < 1024 ? (int)(K - j0) : 1024;
        for (int i = 0; i < m; ++i) { const float v = s_g[i]; rank += (v > mine) || (v == mine && j0 + i < c); }
    }
    if (c < K && rank < n) out[rank] = (int32_t)c;
}

// out[i, :] = src[idx[i], :]
__global__ void __launch_bounds__(256) qdx_gather_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx,
                                                              int64_t B, int32_t D, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= B) return;
    const float* s = src + (int64_t)idx[row] * D; float* o = out + row * D;
    if ((D & 3) == 0 && ((((uintptr_t)src) | ((uintptr_t)out)) & 15u) == 0) {
        for (int q = lane; q < (D >> 2); q += 32) reinterpret_cast<float4*>(o)[q] = __ldg(reinterpret_cast<const float4*>(s) + q);
    } else {
        for (int d = lane; d < D; d += 32) o[d] = s[d];
    }
}

// isoline_variation(x1, x2, key) on dense parents (mutation_operators.py:175-226, one leaf)
__global__ void __launch_bounds__(256) qdx_isoline_kernel(const float* __restrict__ x1, const float* __restrict__ x2,
                                                          int64_t B, int32_t D, QdxKey key, float iso_sigma, float line_sigma,
                                                          int32_t has_min, float minv, int32_t has_max, float maxv,
                                                          float* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B * (int64_t)D) return;
    const int64_t row = e / D;
    const QdxKey k_line = qdx_split(key, 1);
    const QdxKey k_leaf = qdx_split(qdx_split(key, 0), 0);
    const float line = qdx_normal_from_bits(qdx_bits32(k_line, (uint64_t)row)) * line_sigma;
    const float iso = qdx_normal_from_bits(qdx_bits32(k_leaf, (uint64_t)e)) * iso_sigma;
    const float a = x1[e], b = x2[e];
    float t1 = a + iso, t2 = b - a, t3 = t2 * line;
    float x = t1 + t3;
    if (has_min) x = qdx_max_nanprop(x, minv);
    if (has_max) x = qdx_min_nanprop(x, maxv);
    out[e] = x;
}

// isoline_variation on dense parents of a pytree genotype, packed rows (mutation_operators.py:205-224): one shared line
// noise per individual, leaf l's iso noise = normal(keys[l], (B, *leaf_shape)) -> counter i * size_l + j
__global__ void __launch_bounds__(256) qdx_isoline_leaves_kernel(const float* __restrict__ x1, const float* __restrict__ x2,
                                                                 int64_t B, int32_t D, QdxKey k_line, const QdxLeafTab lt,
                                                                 float iso_sigma, float line_sigma, int32_t has_min, float minv,
                                                                 int32_t has_max, float maxv, float* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B * (int64_t)D) return;
    const int64_t row = e / D;
    const int32_t d = (int32_t)(e - row * D);
    const int l = qdx_leaf_of(lt, d);
    const int32_t o = lt.off[l], sz = lt.off[l + 1] - o;
    const float line = qdx_normal_from_bits(qdx_bits32(k_line, (uint64_t)row)) * line_sigma;
    const float iso = qdx_normal_from_bits(qdx_bits32(lt.key[l], (uint64_t)row * (uint64_t)sz + (uint64_t)(d - o))) * iso_sigma;
    const float a = x1[e], b = x2[e];
    float t1 = a + iso, t2 = b - a, t3 = t2 * line;
    float x = t1 + t3;
    if (has_min) x = qdx_max_nanprop(x, minv);
    if (has_max) x = qdx_min_nanprop(x, maxv);
    out[e] = x;
}

// strided 2-D copy: packs / unpacks the leaves of a pytree genotype into / out of the (N, D_total) row layout
__global__ void __launch_bounds__(256) qdx_copy_2d_kernel(const float* __restrict__ src, int64_t src_ld, float* __restrict__ dst,
                                                          int64_t dst_ld, int64_t rows, int64_t cols) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * cols) return;
    const int64_t r = e / cols, c = e - r * cols;
    dst[r * dst_ld + c] = src[r * src_ld + c];
}

// jax.random.{bits, uniform, normal, split} streams for the host-facing qdax_b200.random module
__global__ void __launch_bounds__(256) qdx_random_kernel(QdxKey key, int64_t n, int32_t kind, float minv, float maxv, void* out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t bits = qdx_bits32(key, (uint64_t)i);
    if (kind == 0) ((uint32_t*)out)[i] = bits;
    else if (kind == 1) { float v = qdx_unit_float(bits) * (maxv - minv) + minv; ((float*)out)[i] = v < minv ? minv : v; }
    else ((float*)out)[i] = qdx_normal_from_bits(bits);
}

// metrics only (default_qd_metrics on an arbitrary repertoire)
__global__ void __launch_bounds__(256) qdx_metrics_kernel(const float* __restrict__ rep_f, int64_t K, float qd_offset, float* out) {
    __shared__ double s_sum[8]; __shared__ float s_max[8]; __shared__ int s_cnt[8]; __shared__ int s_nan[8];
    const int lane = threadIdx.x & 31;
    double sum = 0.0; float mx = -INFINITY; int cnt = 0; int nan = 0;
    for (int64_t c = threadIdx.x; c < K; c += blockDim.x) {
        const float v = rep_f[c];
        if (v != -INFINITY) { sum += (double)v; ++cnt; }
        if (v != v) nan = 1; else if (v > mx) mx = v;
    }
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        nan |= __shfl_xor_sync(0xffffffffu, nan, o);
    }
    if (lane == 0) { s_sum[threadIdx.x >> 5] = sum; s_max[threadIdx.x >> 5] = mx; s_cnt[threadIdx.x >> 5] = cnt; s_nan[threadIdx.x >> 5] = nan; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0; float m = -INFINITY; int n = 0, nn = 0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) { s += s_sum[w]; m = fmaxf(m, s_max[w]); n += s_cnt[w]; nn |= s_nan[w]; }
        out[0] = (float)s + qd_offset * (float)n;
        out[1] = nn ? NAN : m;
        out[2] = 100.0f * __fdiv_rn((float)n, (float)K);
    }
}

// =====================================================================================================
// C ABI
// =====================================================================================================
static inline cudaStream_t S(void* s) { return (cudaStream_t)s; }

// ---- host-side key chain (control plane: a handful of Threefry blocks per generation, no device work) ----
static inline uint32_t h_rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static void h_threefry2x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t* o0, uint32_t* o1) {
    const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
    static const int rot[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
    uint32_t x0 = c0 + ks[0], x1 = c1 + ks[1];
    for (int g = 0; g < 5; ++g) {
        for (int j = 0; j < 4; ++j) { x0 += x1; x1 = h_rotl(x1, rot[g & 1][j]); x1 ^= x0; }
        x0 += ks[(g + 1) % 3]; x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
    }
    *o0 = x0; *o1 = x1;
}
static inline QdxKey h_split(QdxKey k, uint32_t i) { QdxKey o; h_threefry2x32(k.a, k.b, 0u, i, &o.a, &o.b); return o; }


static int fill_grid(const qdx_grid_desc* gd, int32_t desc_dim, QdxGrid* g) {
    memset(g, 0, sizeof(*g));
    if (!gd || gd->dd == 0) return 0;
    if (gd->dd != desc_dim || gd->dd < 1 || gd->dd > QDX_MAX_GRID_DIM || !gd->axes) return QDX_ERR_ARG;
    int off = 0;
    for (int d = 0; d < gd->dd; ++d) {
        if (gd->n[d] < 1) return QDX_ERR_ARG;
        g->n[d] = gd->n[d]; g->stride[d] = gd->stride[d]; g->off[d] = off; g->lo[d] = gd->lo[d]; g->hi[d] = gd->hi[d];
        off += gd->n[d];
    }
    if (off > QDX_MAX_AXES) return QDX_ERR_ARG;
    g->dd = gd->dd; g->axes = gd->axes; g->total_axes = off;
    return 0;
}

// Grid of the generate kernel: ceil(B / 128) CTAs, capped at ONE resident wave (SMs x CTAs per SM for this instantiation and
// this much dynamic shared memory): the kernel is persistent, warp w owns rows [B w / W, B (w + 1) / W).
static int32_t generate_tile_rows(int64_t, unsigned);
template <typename Kern>
static int generate_grid(Kern kern, size_t smem, int64_t B, unsigned* grid_out, int32_t* tile_rows_out) {
    // (kernel, shared memory, device) -> resident CTAs: queried once, then served from a small table (the attribute call and
    // the occupancy query cost microseconds each, on a path that enqueues a 100 us generation)
    struct Entry { const void* k; size_t smem; int dev; int64_t cap; };
    static Entry table[128];
    static int n_entries = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    int64_t cap = 0;
    size_t limit = 0;                               // largest dynamic shared memory size this kernel has been opted in for (never lowered)
    for (int i = 0; i < n_entries; ++i)
        if (table[i].k == (const void*)kern && table[i].dev == dev) {
            if (table[i].smem > limit) limit = table[i].smem;
            if (table[i].smem == smem) cap = table[i].cap;
        }
    if (cap == 0) {
        int sms = 0, per_sm = 0;
        if (smem > limit || n_entries >= 128) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > limit ? smem : limit));
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, QDX_GEN_WARPS * 32, smem);
        if (e != cudaSuccess) return (int)e;
        if (per_sm < 1) return QDX_ERR_UNSUPPORTED;
        cap = (int64_t)sms * per_sm;
        if (n_entries < 128) table[n_entries++] = Entry{(const void*)kern, smem, dev, cap};     // (a benign race: worst case an entry is queried twice)
    }
    // Batches of at most one 32-row tile per resident warp (README-sized runs, the 65 536-row configurations) are latency
    // problems: the kernel lasts as long as one warp needs for its tile, so the rows are spread evenly over ALL resident warps
    // in tiles of ceil(rows per warp) rounded up to 4, at least 8 (QDX_GEN_SMALL_TILES=0: off).
    static int small_tiles = -1;
    if (small_tiles < 0) { const char* e_ = getenv("QDX_GEN_SMALL_TILES"); small_tiles = e_ ? atoi(e_) : 1; }
    int32_t tile_rows = generate_tile_rows(B, 0);
    const int64_t per_warp = (B + cap * QDX_GEN_WARPS - 1) / (cap * QDX_GEN_WARPS);
    if (small_tiles && per_warp <= 32) { tile_rows = (int32_t)((per_warp + 3) / 4 * 4); if (tile_rows < 8) tile_rows = 8; }
    static int force_rows = -1;      // QDX_GEN_TILE_ROWS=n: timing experiments (tools/time_generate.py)
    if (force_rows < 0) { const char* e_ = getenv("QDX_GEN_TILE_ROWS"); force_rows = e_ ? atoi(e_) : 0; }
    if (force_rows >= 4 && force_rows <= 32) tile_rows = force_rows & ~3;
    const int64_t g = (B + QDX_GEN_WARPS * tile_rows - 1) / (QDX_GEN_WARPS * tile_rows);
    *grid_out = (unsigned)(g < cap ? g : cap);
    *tile_rows_out = tile_rows;
    return 0;
}

// Tile height.  Measured on B200 (tools/time_generate.py, arm 100-D, rows per launch 131 072 / 262 144 / 2^20; in-kernel span):
//   one 32-row tile per warp, ceil(B / 128) CTAs (hardware CTA scheduler)        93.4 / 177.3 / 677 us
//   persistent, static equal split of the rows                                   99.3 / 194.1 / 737 us   (warps progress unevenly)
//   persistent, dynamic, guided (tiles shrinking from 32 to 8 rows)             100.9 / 179.4 / 676 us   (row-serial scoring phase costs the same for 8 rows as for 32)
//   persistent, dynamic, equal tiles of 28 rows (whole tiles per warp)           96.0 / 203.1 / 674 us
//   persistent, dynamic, 32-row tiles                                            this build
//   same + end game: once less than one round of tiles is left, pieces of 8 / 16 rows fill a warp's 32-row tile before ONE
//   scoring pass (round 2, gpurun_out/r2r_gen_*.json)                           110.2 / 193.0 / 688 us (8 rows), 100.0 / 182.9 / 686 (16)
//                                                                          vs    98.6 / 176.9 / 680 us without, same run: the finer
//   forced tile heights for every batch size (QDX_GEN_TILE_ROWS, gpurun_out/r2x_gen_*.json), 2^20 rows: 32 rows 673 us, 28: 707, 24: 735,
//   20: 789, 16: 861 -- the lane = row phases cost ~28 % of a 32-row tile and do not shrink with its height.
//   grab for the NEXT tile issued at the start of the current one (the atomic's round trip hidden behind a tile of work): 2^20 rows
//   679 -> 690 us, c2 (one tile per warp) 72 -> 105 us -- a warp that commits to its next tile a tile early takes it from a warp
//   that would have been free sooner (gpurun_out/r3f_*).
//   The finer deal does not shorten the tail -- what is left at the end runs on SM sub-partitions with one or two warps, at their
//   latency-bound single-warp rate, however it is cut -- and pays a lane = row parent-selection pass per piece.
// The tail of the kernel is the lowest-priority warp of every SM sub-partition finishing its last tile alone, at
// single-warp issue rate; smaller last tiles shorten it but pay the fixed per-tile cost of the lane = row scoring phase.
static int32_t generate_tile_rows(int64_t, unsigned) { return 32; }

template <int TASK, bool ARM_CLIP>
static int launch_generate_task(const QdxGenParams& p, size_t smem, cudaStream_t st) {
#define QDX_LAUNCH_GEN(GD)                                                                                         \
    do {                                                                                                           \
        unsigned g_ = 0;                                                                                           \
        int32_t tr_ = 32;                                                                                          \
        int rc_ = generate_grid(qdx_generate_kernel<TASK, GD, ARM_CLIP>, smem, p.B, &g_, &tr_);                    \
        if (rc_) return rc_;                                                                                       \
        QdxGenParams q_ = p; q_.tile_rows = tr_;                                                                   \
        cudaError_t e_ = qdx_launch_pdl(qdx_generate_kernel<TASK, GD, ARM_CLIP>, dim3(g_), dim3(QDX_GEN_WARPS * 32), smem, st, q_); \
        if (e_ != cudaSuccess) return (int)e_;                                                                     \
    } while (0)
    const int gd = (TASK == QDX_TASK_NONE) ? 0 : (p.grid.dd ? p.grid.dd : -p.cvt.dd);
    if (TASK == QDX_TASK_ARM) {            // arm descriptors are 2-D: grid 2, bucket index 2, or none
        switch (gd) {
            case 0: QDX_LAUNCH_GEN(0); break;
            case 2: QDX_LAUNCH_GEN(2); break;
            case -2: QDX_LAUNCH_GEN(-2); break;
            default: return QDX_ERR_ARG;
        }
    } else if (TASK == QDX_TASK_NONE) {
        if (p.leaves.n > 1) {
            unsigned g_ = 0;
            int32_t tr_ = 32;
            int rc_ = generate_grid(qdx_generate_kernel<QDX_TASK_NONE, 0, false, true>, smem, p.B, &g_, &tr_);
            if (rc_) return rc_;
            QdxGenParams q_ = p; q_.tile_rows = tr_;
            cudaError_t e_ = qdx_launch_pdl(qdx_generate_kernel<QDX_TASK_NONE, 0, false, true>, dim3(g_), dim3(QDX_GEN_WARPS * 32), smem, st, q_);
            if (e_ != cudaSuccess) return (int)e_;
        } else {
            QDX_LAUNCH_GEN(0);
        }
    } else {
        switch (gd) {
            case 0: QDX_LAUNCH_GEN(0); break;
            case 1: QDX_LAUNCH_GEN(1); break;
            case 2: QDX_LAUNCH_GEN(2); break;
            case 3: QDX_LAUNCH_GEN(3); break;
            case 4: QDX_LAUNCH_GEN(4); break;
            case -1: QDX_LAUNCH_GEN(-1); break;
            case -2: QDX_LAUNCH_GEN(-2); break;
            case -3: QDX_LAUNCH_GEN(-3); break;
            default: return QDX_ERR_ARG;
        }
    }
#undef QDX_LAUNCH_GEN
    QDX_CHECK_LAUNCH();
    return 0;
}

// warp-per-cell commit kernel, ordinary launch: fallback of qdx_commit (qdx_commit.cu) when a cooperative launch or the
// 16-byte row alignment of the bulk-copy path is not available
int qdx_launch_commit_generic(void* ws, int64_t K, int64_t D, int32_t desc_dim, const float* off_genotypes, const float* off_fitness,
               const float* off_desc, uint32_t idx_base, int64_t B, int32_t first_wins, float* rep_genotypes,
               float* rep_fitness, float* rep_desc, float qd_offset, float* metrics_out4, int32_t* added_cells,
               int32_t mode, cudaStream_t stream) {
    if (!ws || !off_genotypes || !off_fitness || !off_desc || !rep_genotypes || !rep_fitness || !rep_desc) return QDX_ERR_ARG;
    if (K <= 0 || D <= 0 || desc_dim < 1 || B < 0 || mode < 0 || mode > 2) return QDX_ERR_ARG;
    int64_t ctas = (K + 7) / 8;                 // one warp per cell, 8 warps per CTA, grid-stride beyond one wave
    if (ctas > QDX_COMMIT_CTAS) ctas = QDX_COMMIT_CTAS;
    if (ctas < 1) ctas = 1;
    qdx_commit_kernel<<<(unsigned)ctas, 256, 0, stream>>>(ws, K, (int32_t)D, desc_dim, off_genotypes, off_fitness, off_desc,
                                                            idx_base, B, first_wins, rep_genotypes, rep_fitness, rep_desc,
                                                            qd_offset, metrics_out4, added_cells, mode);
    QDX_CHECK_LAUNCH();
    return 0;
}

#if QDX_GEN_TRACE
extern "C" int qdx_debug_gen_trace(unsigned long long* out4, int reset) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess && out4) e = cudaMemcpyFromSymbol(out4, g_gen_trace, sizeof(unsigned long long) * 4);
    if (e == cudaSuccess && reset) { unsigned long long init[4] = {~0ull, 0ull, ~0ull, 0ull}; e = cudaMemcpyToSymbol(g_gen_trace, init, sizeof(init)); }
    return (int)e;
}
#endif

#if QDX_XCHG_TRACE
extern "C" int qdx_debug_xchg_trace(unsigned long long* out8, int reset) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess && out8) e = cudaMemcpyFromSymbol(out8, g_xchg_trace, sizeof(unsigned long long) * 8);
    if (e == cudaSuccess && reset) { unsigned long long z[8] = {0}; e = cudaMemcpyToSymbol(g_xchg_trace, z, sizeof(z)); }
    return (int)e;
}
#endif

extern "C" {

int qdx_version(void) { return 100; }

int qdx_workspace_bytes(int64_t K, int64_t* bytes) {
    if (K <= 0 || K >= (1ll << 31) || !bytes) return QDX_ERR_ARG;
    *bytes = (int64_t)qdx_ws_total_bytes(K);
    return 0;
}

int qdx_workspace_keytab_offset(int64_t K, int64_t* offset) {
    if (K <= 0 || K >= (1ll << 31) || !offset) return QDX_ERR_ARG;
    *offset = (int64_t)qdx_ws_keytab_offset(K);
    return 0;
}

int qdx_workspace_init(void* ws, int64_t K, void* stream) {
    if (!ws || K <= 0) return QDX_ERR_ARG;
    return (int)cudaMemsetAsync(ws, 0, qdx_ws_total_bytes(K), S(stream));
}

int qdx_workspace_set_carry_key(void* ws, uint32_t k0, uint32_t k1, void* stream) {
    if (!ws) return QDX_ERR_ARG;
    uint32_t k[2] = {k0, k1};
    // 8-byte async copy from pageable memory is staged by the runtime before returning
    return (int)cudaMemcpyAsync((char*)ws + offsetof(QdxWorkspace, carry), k, sizeof(k), cudaMemcpyHostToDevice, S(stream));
}

int qdx_workspace_copy_carry_key(void* ws, uint32_t* key2_device, int32_t to_workspace, void* stream) {
    if (!ws || !key2_device) return QDX_ERR_ARG;
    char* carry = (char*)ws + offsetof(QdxWorkspace, carry);
    return (int)(to_workspace ? cudaMemcpyAsync(carry, key2_device, 2 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, S(stream))
                              : cudaMemcpyAsync(key2_device, carry, 2 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, S(stream)));
}

int qdx_workspace_set_error_mirror(void* ws, int32_t* host_pinned, void* stream) {
    if (!ws) return QDX_ERR_ARG;
    // the 8-byte pointer is staged by the runtime before the call returns (pageable source)
    return (int)cudaMemcpyAsync((char*)ws + offsetof(QdxWorkspace, err_host), &host_pinned, sizeof(host_pinned), cudaMemcpyHostToDevice, S(stream));
}

int qdx_workspace_read(void* ws, uint32_t* carry_key2, float* metrics4, int32_t* error, void* stream) {
    if (!ws) return QDX_ERR_ARG;
    QdxWorkspace h;   // header only (a few KB)
    cudaError_t e = cudaMemcpyAsync(&h, ws, sizeof(QdxWorkspace), cudaMemcpyDeviceToHost, S(stream));
    if (e != cudaSuccess) return (int)e;
    e = cudaStreamSynchronize(S(stream));
    if (e != cudaSuccess) return (int)e;
    if (carry_key2) { carry_key2[0] = h.carry.a; carry_key2[1] = h.carry.b; }
    if (metrics4) for (int j = 0; j < 4; ++j) metrics4[j] = h.metrics[j];
    if (error) *error = h.error;
    return 0;
}

int qdx_select_prepare(const float* rep_fitness, int64_t K, void* ws, int32_t key_mode, uint32_t k0, uint32_t k1,
                       int32_t rank_slot, void* stream) {
    if (!rep_fitness || !ws || K <= 0 || key_mode < 0 || key_mode > 4 || rank_slot >= QDX_MAX_RANKS) return QDX_ERR_ARG;
    qdx_prepare_kernel<<<1, 1024, 0, S(stream)>>>(rep_fitness, K, ws, QdxKey{k0, k1}, key_mode, rank_slot);
    QDX_CHECK_LAUNCH();
    return 0;
}

static int fill_leaves(const qdx_leaf_table* in, int64_t D, QdxLeafTab* out) {
    memset(out, 0, sizeof(*out));
    if (!in) return 0;
    if (in->n < 1 || in->n > QDX_MAX_LEAVES || in->off[0] != 0 || in->off[in->n] != D) return QDX_ERR_ARG;
    for (int l = 0; l < in->n; ++l) if (in->off[l + 1] < in->off[l]) return QDX_ERR_ARG;
    out->n = in->n;
    for (int l = 0; l <= in->n; ++l) out->off[l] = in->off[l];
    for (int l = 0; l < in->n; ++l) out->key[l] = QdxKey{in->key[2 * l], in->key[2 * l + 1]};
    return 0;
}

static int generate_impl(const float* rep_genotypes, const float* rep_fitness, const float* centroids, void* ws, int64_t K,
                 int64_t D, int64_t B, float iso_sigma, float line_sigma, int32_t has_min, float minval, int32_t has_max,
                 float maxval, int32_t task, int32_t desc_dim, const qdx_grid_desc* grid, int32_t offer,
                 uint32_t idx_base, int32_t first_wins, float* out_genotypes, float* out_fitness, float* out_desc,
                 int32_t* out_cells, int32_t* out_p1, int32_t* out_p2, const uint32_t* gen_keys8, const qdx_cvt_index* cvt,
                 const qdx_leaf_table* leaves, int32_t flags, void* stream) {
    if (!rep_genotypes || !rep_fitness || !ws || K <= 0 || D <= 0 || B < 0 || (D & 3)) return QDX_ERR_ARG;
    if (flags & ~(QDX_GEN_ROWS_FIRED_ONLY | QDX_GEN_OUT_XCHG)) return QDX_ERR_ARG;
    if ((flags & QDX_GEN_OUT_XCHG) && task == QDX_TASK_NONE) return QDX_ERR_ARG;
    if (flags & QDX_GEN_OUT_XCHG) out_genotypes = (float*)16;      // placeholder: the rows go to this rank's offspring block, resolved on the device
    if (task < QDX_TASK_NONE || task > QDX_TASK_SPHERE) return QDX_ERR_ARG;
    if (task != QDX_TASK_NONE && (!out_fitness || !out_desc || desc_dim < 1 || desc_dim > D || desc_dim > 128)) return QDX_ERR_ARG;
    if (task == QDX_TASK_ARM && desc_dim != 2) return QDX_ERR_ARG;
    if (task == QDX_TASK_NONE && (!out_genotypes || offer)) return QDX_ERR_ARG;
    if ((uint64_t)idx_base + (uint64_t)B > 0x7FFFFFFFull) return QDX_ERR_ARG;
    if ((((uintptr_t)rep_genotypes) | ((uintptr_t)out_genotypes)) & 15u) return QDX_ERR_ARG;   // 128-bit parent loads, bulk-copy stores
    if (B == 0) return 0;
    QdxGenParams p;
    memset(&p, 0, sizeof(p));
    int rc = fill_grid(task == QDX_TASK_NONE ? nullptr : grid, desc_dim, &p.grid);
    if (rc) return rc;
    rc = qdx_fill_cvt_index((task == QDX_TASK_NONE || p.grid.dd) ? nullptr : cvt, &p.cvt);
    if (rc) return rc;
    if (p.cvt.dd && p.cvt.dd != desc_dim) return QDX_ERR_ARG;
    if (offer && p.grid.dd == 0 && p.cvt.dd == 0) return QDX_ERR_ARG;   // offer needs cells: grid fast path or bucket index
    if (p.grid.dd && !centroids) return QDX_ERR_ARG;
    p.rep_g = rep_genotypes; p.rep_f = rep_fitness; p.centroids = centroids; p.ws = ws;
    p.B = B; p.K = K; p.D = (int32_t)D;
    // chunk of genes staged per warp tile: whole row when it fits 16 KB per warp, else 128 genes
    // genes per staged chunk: the whole row up to 128 genes; longer rows in chunks of 100 -- a 32-row tile of 100 floats per row
    // (stride 100, 100 / 4 odd: conflict-free) is 12.8 KB per warp, 4 CTAs per SM, where 128-gene chunks (stride 132) allowed
    // 3: 444 instead of 592 resident CTAs turned the 65 536-row batch of configuration c4 into 1.15 waves of 32-row tiles
    // (generate 0.455 -> 0.407 ms, gpurun_out/r3u_*; chunks of 64: 0.415).  QDX_GEN_CHUNK=n overrides (A/B).
    static int chunk = -1;
    if (chunk < 0) { const char* e_ = getenv("QDX_GEN_CHUNK"); chunk = e_ ? atoi(e_) : 100; if (chunk < 8 || chunk > 128 || (chunk & 3)) chunk = 100; }
    p.DC = (D <= 128) ? (int32_t)D : (desc_dim <= chunk ? chunk : 128);
    p.DS = ((p.DC & 7) == 4) ? p.DC : p.DC + 4;
    if (task == QDX_TASK_ARM && D > p.DC) return QDX_ERR_UNSUPPORTED;   // arm needs the whole row for its two passes
    p.out_xchg = (flags & QDX_GEN_OUT_XCHG) ? 1 : 0;
    // only the rows whose offer fired: needs the offer in this kernel and the whole row in the tile when the offer is made
    p.store_mode = ((flags & QDX_GEN_ROWS_FIRED_ONLY) && offer && D <= p.DC) ? 1 : 0;
    p.iso_sigma = iso_sigma; p.line_sigma = line_sigma; p.has_min = has_min; p.has_max = has_max; p.minv = minval; p.maxv = maxval;
    p.out_g = out_genotypes; p.out_f = out_fitness; p.out_d = out_desc; p.out_cell = out_cells; p.out_p1 = out_p1; p.out_p2 = out_p2;
    p.desc_dim = desc_dim; p.offer = offer; p.idx_base = idx_base; p.first_wins = first_wins;
    if (gen_keys8) {
        p.keys_by_value = 1;
        p.keys.sel1 = QdxKey{gen_keys8[0], gen_keys8[1]}; p.keys.sel2 = QdxKey{gen_keys8[2], gen_keys8[3]};
        p.keys.line = QdxKey{gen_keys8[4], gen_keys8[5]}; p.keys.leaf = QdxKey{gen_keys8[6], gen_keys8[7]};
    }
    if (leaves) {
        if (task != QDX_TASK_NONE || !gen_keys8) return QDX_ERR_ARG;
        rc = fill_leaves(leaves, D, &p.leaves);
        if (rc) return rc;
        if (p.leaves.n == 1) p.keys.leaf = p.leaves.key[0];
    }
    const size_t smem = ((size_t)QDX_GEN_WARPS * 32 * p.DS + (size_t)p.grid.total_axes) * sizeof(float);
    // arm.py:27 clip(params, 0, 1) is the identity when the variation already clipped into [0, 1]
    const bool arm_clip = !(has_min && has_max && minval >= 0.0f && maxval <= 1.0f);
    switch (task) {
        case QDX_TASK_NONE: return launch_generate_task<QDX_TASK_NONE, false>(p, smem, S(stream));
        case QDX_TASK_ARM: return arm_clip ? launch_generate_task<QDX_TASK_ARM, true>(p, smem, S(stream))
                                           : launch_generate_task<QDX_TASK_ARM, false>(p, smem, S(stream));
        case QDX_TASK_RASTRIGIN: return launch_generate_task<QDX_TASK_RASTRIGIN, false>(p, smem, S(stream));
        default: return launch_generate_task<QDX_TASK_SPHERE, false>(p, smem, S(stream));
    }
}

int qdx_score(int32_t task, const float* genotypes, int64_t B, int64_t D, int32_t desc_dim, float* out_fitness,
              float* out_desc, void* stream) {
    if (!genotypes || !out_fitness || !out_desc || B < 0 || D <= 0 || desc_dim < 1 || desc_dim > D) return QDX_ERR_ARG;
    if (task < QDX_TASK_ARM || task > QDX_TASK_SPHERE || (task == QDX_TASK_ARM && desc_dim != 2)) return QDX_ERR_ARG;
    if (B == 0) return 0;
    const size_t smem = 4 * 32 * (64 + 1) * sizeof(float);
    const dim3 g((unsigned)((B + 127) / 128));
    if (task == QDX_TASK_ARM) qdx_score_kernel<QDX_TASK_ARM><<<g, 128, smem, S(stream)>>>(genotypes, B, (int32_t)D, desc_dim, out_fitness, out_desc, QdxNoise{});
    else if (task == QDX_TASK_RASTRIGIN) qdx_score_kernel<QDX_TASK_RASTRIGIN><<<g, 128, smem, S(stream)>>>(genotypes, B, (int32_t)D, desc_dim, out_fitness, out_desc, QdxNoise{});
    else qdx_score_kernel<QDX_TASK_SPHERE><<<g, 128, smem, S(stream)>>>(genotypes, B, (int32_t)D, desc_dim, out_fitness, out_desc, QdxNoise{});
    QDX_CHECK_LAUNCH();
    return 0;
}

int qdx_score_noisy_arm(const float* genotypes, int64_t B, int64_t D, uint32_t k0, uint32_t k1, float fit_variance, float desc_variance,
                        float params_variance, float* out_fitness, float* out_desc, void* stream) {
    if (!genotypes || !out_fitness || !out_desc || B < 0 || D <= 0) return QDX_ERR_ARG;
    if (B == 0) return 0;
    const QdxKey key{k0, k1};
    QdxNoise nz;
    nz.kf = h_split(key, 1); nz.kd = h_split(key, 2); nz.kp = h_split(key, 3);            // arm.py:65: key, f_sub, d_sub, p_sub = split(key, 4)
    nz.fit_var = fit_variance; nz.desc_var = desc_variance; nz.params_var = params_variance;
    const size_t smem = 4 * 32 * (64 + 1) * sizeof(float);
    qdx_score_kernel<QDX_TASK_ARM, true><<<(unsigned)((B + 127) / 128), 128, smem, S(stream)>>>(genotypes, B, (int32_t)D, 2, out_fitness, out_desc, nz);
    QDX_CHECK_LAUNCH();
    return 0;
}

int qdx_cells(const float* desc, int64_t B, int32_t desc_dim, const float* centroids, int64_t K, const qdx_grid_desc* grid,
              int32_t* out_cells, void* ws, const float* rep_fitness, const float* fitness, int32_t offer,
              uint32_t idx_base, int32_t first_wins, void* stream) {
    if (!desc || !centroids || !out_cells || B < 0 || K <= 0 || desc_dim < 1) return QDX_ERR_ARG;
    if (offer && (!ws || !rep_fitness || !fitness)) return QDX_ERR_ARG;
    if ((uint64_t)idx_base + (uint64_t)B > 0x7FFFFFFFull) return QDX_ERR_ARG;
    if (B == 0) return 0;
    QdxGrid g;
    int rc = fill_grid(grid, desc_dim, &g);
    if (rc) return rc;
    cudaStream_t st = S(stream);
    if (g.dd) {
        const dim3 gr((unsigned)((B + 255) / 256));
        switch (g.dd) {
            case 1: qdx_cells_grid_kernel<1><<<gr, 256, 0, st>>>(desc, B, g, centroids, K, out_cells, ws, rep_fitness, fitness, offer, idx_base, first_wins); break;
            case 2: qdx_cells_grid_kernel<2><<<gr, 256, 0, st>>>(desc, B, g, centroids, K, out_cells, ws, rep_fitness, fitness, offer, idx_base, first_wins); break;
            case 3: qdx_cells_grid_kernel<3><<<gr, 256, 0, st>>>(desc, B, g, centroids, K, out_cells, ws, rep_fitness, fitness, offer, idx_base, first_wins); break;
            default: qdx_cells_grid_kernel<4><<<gr, 256, 0, st>>>(desc, B, g, centroids, K, out_cells, ws, rep_fitness, fitness, offer, idx_base, first_wins); break;
        }
    } else if (desc_dim <= 4) {
        const dim3 gr((unsigned)((B + 255) / 256));
        const size_t smem = (size_t)2048 * desc_dim * sizeof(float);
        switch (desc_dim) {
            case 1: qdx_cells_bf_kernel<1><<<gr, 256, smem, st>>>(desc, B, centroids, K, out_cells, ws, rep_fitness, fitness, offer, idx_base, first_wins); break;
            case 2: qdx_cells_bf_kernel<2><<<gr, 256, smem, st>>>(desc, B, centroids, K, out_cells, ws, rep_fitness, fitness, offer, idx_base, first_wins); break;
            case 3: qdx_cells_bf_kernel<3><<<gr, 256, smem, st>>>(desc, B, centroids, K, out_cells, ws, rep_fitness, fitness, offer, idx_base, first_wins); break;
            default: qdx_cells_bf_kernel<4><<<gr, 256, smem, st>>>(desc, B, centroids, K, out_cells, ws, rep_fitness, fitness, offer, idx_base, first_wins); break;
        }
    } else {
        const dim3 gr((unsigned)((B * 32 + 127) / 128));
        qdx_cells_bf_generic_kernel<<<gr, 128, 0, st>>>(desc, B, desc_dim, centroids, K, out_cells, ws, rep_fitness, fitness, offer, idx_base, first_wins);
    }
    QDX_CHECK_LAUNCH();
    return 0;
}

int qdx_offer_cells(const int32_t* cells, const float* fitness, int64_t B, int64_t K, void* ws, const float* rep_fitness,
                    uint32_t idx_base, int32_t first_wins, void* stream) {
    if (!cells || !fitness || !ws || !rep_fitness || B < 0 || K <= 0) return QDX_ERR_ARG;
    if ((uint64_t)idx_base + (uint64_t)B > 0x7FFFFFFFull) return QDX_ERR_ARG;
    if (B == 0) return 0;
    qdx_offer_kernel<<<(unsigned)((B + 255) / 256), 256, 0, S(stream)>>>(cells, fitness, B, K, ws, rep_fitness, idx_base, first_wins);
    QDX_CHECK_LAUNCH();
    return 0;
}

static int launch_elect(int32_t task, void* ws, int64_t K, int64_t D, int32_t desc_dim, int64_t B_dev, int32_t nranks,
                        const float* rep_g, float iso_sigma, float line_sigma, int32_t has_min, float minval, int32_t has_max,
                        float maxval, int32_t first_wins, float* sg, float* sf, float* sd, int32_t wait_peers, cudaStream_t st) {
    int64_t ctas = (K + 63) / 64;               // >= 64 cells per CTA, at most one wave (3 CTAs of 256 threads per SM)
    if (ctas > 148 * 3) ctas = 148 * 3;
#define QDX_ELECT(T) qdx_elect_kernel<T><<<(unsigned)ctas, 256, 0, st>>>(ws, K, (int32_t)D, desc_dim, B_dev, nranks, rep_g, iso_sigma, \
        line_sigma, has_min, minval, has_max, maxval, first_wins, sg, sf, sd, wait_peers)
    switch (task) {
        case QDX_TASK_NONE: QDX_ELECT(QDX_TASK_NONE); break;
        case QDX_TASK_ARM: QDX_ELECT(QDX_TASK_ARM); break;
        case QDX_TASK_RASTRIGIN: QDX_ELECT(QDX_TASK_RASTRIGIN); break;
        default: QDX_ELECT(QDX_TASK_SPHERE); break;
    }
#undef QDX_ELECT
    QDX_CHECK_LAUNCH();
    return 0;
}

int qdx_regenerate_winners(void* ws, int64_t K, int64_t D, int64_t B_dev, int32_t nranks, const float* rep_genotypes,
                           float iso_sigma, float line_sigma, int32_t has_min, float minval, int32_t has_max, float maxval,
                           int32_t first_wins, float* stage_genotypes, void* stream) {
    if (!ws || !rep_genotypes || !stage_genotypes || K <= 0 || D <= 0 || B_dev <= 0 || nranks < 1 || nranks > QDX_MAX_RANKS) return QDX_ERR_ARG;
    return launch_elect(QDX_TASK_NONE, ws, K, D, 1, B_dev, nranks, rep_genotypes, iso_sigma, line_sigma, has_min, minval, has_max,
                        maxval, first_wins, stage_genotypes, nullptr, nullptr, 0, S(stream));
}

int qdx_elect_winners(void* ws, int64_t K, int64_t D, int32_t task, int32_t desc_dim, int64_t B_dev, int32_t nranks,
                      const float* rep_genotypes, float iso_sigma, float line_sigma, int32_t has_min, float minval,
                      int32_t has_max, float maxval, int32_t first_wins, float* stage_genotypes, float* stage_fitness,
                      float* stage_desc, int32_t wait_peers, void* stream) {
    if (!ws || !rep_genotypes || !stage_genotypes || !stage_fitness || !stage_desc) return QDX_ERR_ARG;
    if (K <= 0 || D <= 0 || B_dev <= 0 || nranks < 1 || nranks > QDX_MAX_RANKS) return QDX_ERR_ARG;
    if (task < QDX_TASK_ARM || task > QDX_TASK_SPHERE || desc_dim < 1 || desc_dim > D || desc_dim > 128) return QDX_ERR_ARG;
    if (task == QDX_TASK_ARM && (desc_dim != 2 || D > 128)) return QDX_ERR_UNSUPPORTED;
    return launch_elect(task, ws, K, D, desc_dim, B_dev, nranks, rep_genotypes, iso_sigma, line_sigma, has_min, minval, has_max,
                        maxval, first_wins, stage_genotypes, stage_fitness, stage_desc, wait_peers, S(stream));
}

// ---- peer-memory exchange buffers (cudaMalloc + cudaIpc: one process per GPU, one NVLink / NVSwitch domain) ----
int qdx_xchg_bytes(int64_t K, int64_t B_dev, int64_t D, int32_t desc_dim, int64_t* bytes) {
    if (K <= 0 || K >= (1ll << 31) || B_dev < 0 || !bytes || (B_dev > 0 && (D <= 0 || desc_dim < 1))) return QDX_ERR_ARG;
    *bytes = (int64_t)qdx_xchg_total_bytes(K, B_dev, D, desc_dim);
    return 0;
}

int qdx_xchg_create(int64_t K, int64_t B_dev, int64_t D, int32_t desc_dim, void** buf, void* ipc_handle64) {
    if (K <= 0 || K >= (1ll << 31) || B_dev < 0 || !buf || !ipc_handle64 || (B_dev > 0 && (D <= 0 || desc_dim < 1))) return QDX_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    void* p = nullptr;
    const size_t n = qdx_xchg_total_bytes(K, B_dev, D, desc_dim);
    cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(p, 0, n);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaIpcGetMemHandle((cudaIpcMemHandle_t*)ipc_handle64, p);
    if (e != cudaSuccess) { cudaFree(p); return (int)e; }
    *buf = p;
    return 0;
}

int qdx_xchg_open(const void* ipc_handle64, void** peer_buf) {
    if (!ipc_handle64 || !peer_buf) return QDX_ERR_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle64, sizeof(h));
    return (int)cudaIpcOpenMemHandle(peer_buf, h, cudaIpcMemLazyEnablePeerAccess);
}

int qdx_xchg_close(void* peer_buf) { return peer_buf ? (int)cudaIpcCloseMemHandle(peer_buf) : QDX_ERR_ARG; }
int qdx_xchg_destroy(void* buf) { return buf ? (int)cudaFree(buf) : QDX_ERR_ARG; }

int qdx_xchg_attach(void* ws, int32_t rank, int32_t nranks, void* const* bufs, int64_t B_dev, int64_t D, int32_t desc_dim,
                    int32_t timeout_ms, void* stream) {
    if (!ws || nranks < 0 || nranks > QDX_MAX_PEERS || (nranks > 0 && (!bufs || rank < 0 || rank >= nranks))) return QDX_ERR_ARG;
    if (B_dev < 0 || (B_dev > 0 && (D <= 0 || desc_dim < 1))) return QDX_ERR_ARG;
    struct { uint32_t push_ticket, pad0; unsigned long long peer[QDX_MAX_PEERS]; int32_t rank, nranks; int64_t bdev; int32_t D, Dd, timeout_ms, pad1; } h;
    static_assert(offsetof(QdxWorkspace, pad1) - offsetof(QdxWorkspace, push_ticket) + sizeof(int32_t) == sizeof(h), "layout");
    memset(&h, 0, sizeof(h));
    for (int q = 0; q < nranks; ++q) { if (!bufs[q]) return QDX_ERR_ARG; h.peer[q] = (unsigned long long)bufs[q]; }
    h.rank = rank; h.nranks = nranks; h.bdev = nranks > 0 ? B_dev : 0; h.D = (int32_t)D; h.Dd = desc_dim; h.timeout_ms = timeout_ms;
    return (int)cudaMemcpyAsync((char*)ws + offsetof(QdxWorkspace, push_ticket), &h, sizeof(h), cudaMemcpyHostToDevice, S(stream));
}

int qdx_xchg_push(void* ws, int64_t K, const uint32_t* gen_keys8, void* stream) {
    if (!ws || K <= 0 || !gen_keys8) return QDX_ERR_ARG;
    QdxGenKeys g;
    g.sel1 = QdxKey{gen_keys8[0], gen_keys8[1]}; g.sel2 = QdxKey{gen_keys8[2], gen_keys8[3]};
    g.line = QdxKey{gen_keys8[4], gen_keys8[5]}; g.leaf = QdxKey{gen_keys8[6], gen_keys8[7]};
    qdx_publish_kernel<<<1, 32, 0, S(stream)>>>(ws, K, g);
    QDX_CHECK_LAUNCH();
    return 0;
}

// jax.random.split(key, n): out[2 i .. 2 i + 1] = threefry(key, counter = (0, i))
int qdx_generate(const float* rep_genotypes, const float* rep_fitness, const float* centroids, void* ws, int64_t K,
                 int64_t D, int64_t B, float iso_sigma, float line_sigma, int32_t has_min, float minval, int32_t has_max,
                 float maxval, int32_t task, int32_t desc_dim, const qdx_grid_desc* grid, int32_t offer,
                 uint32_t idx_base, int32_t first_wins, float* out_genotypes, float* out_fitness, float* out_desc,
                 int32_t* out_cells, int32_t* out_p1, int32_t* out_p2, const uint32_t* gen_keys8, const qdx_cvt_index* cvt,
                 int32_t flags, void* stream) {
    return generate_impl(rep_genotypes, rep_fitness, centroids, ws, K, D, B, iso_sigma, line_sigma, has_min, minval, has_max, maxval,
                         task, desc_dim, grid, offer, idx_base, first_wins, out_genotypes, out_fitness, out_desc, out_cells, out_p1,
                         out_p2, gen_keys8, cvt, nullptr, flags, stream);
}

int qdx_generate_leaves(const float* rep_genotypes, const float* rep_fitness, void* ws, int64_t K, int64_t D, int64_t B,
                        float iso_sigma, float line_sigma, int32_t has_min, float minval, int32_t has_max, float maxval,
                        float* out_genotypes, int32_t* out_p1, int32_t* out_p2, const uint32_t* gen_keys8,
                        const qdx_leaf_table* leaves, void* stream) {
    if (!leaves || !gen_keys8) return QDX_ERR_ARG;
    return generate_impl(rep_genotypes, rep_fitness, nullptr, ws, K, D, B, iso_sigma, line_sigma, has_min, minval, has_max, maxval,
                         QDX_TASK_NONE, 1, nullptr, 0, 0u, 1, out_genotypes, nullptr, nullptr, nullptr, out_p1, out_p2, gen_keys8,
                         nullptr, leaves, 0, stream);
}

int qdx_isoline_variation_leaves(const float* x1, const float* x2, int64_t B, int64_t D, uint32_t line_k0, uint32_t line_k1,
                                 const qdx_leaf_table* leaves, float iso_sigma, float line_sigma, int32_t has_min, float minval,
                                 int32_t has_max, float maxval, float* out, void* stream) {
    if (!x1 || !x2 || !out || !leaves || B < 0 || D <= 0) return QDX_ERR_ARG;
    QdxLeafTab lt;
    int rc = fill_leaves(leaves, D, &lt);
    if (rc) return rc;
    if (B == 0) return 0;
    const int64_t n = B * D;
    qdx_isoline_leaves_kernel<<<(unsigned)((n + 255) / 256), 256, 0, S(stream)>>>(x1, x2, B, (int32_t)D, QdxKey{line_k0, line_k1}, lt,
                                                                                iso_sigma, line_sigma, has_min, minval, has_max, maxval, out);
    QDX_CHECK_LAUNCH();
    return 0;
}

int qdx_copy_2d(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int64_t rows, int64_t cols, void* stream) {
    if (!src || !dst || rows < 0 || cols < 0 || src_ld < cols || dst_ld < cols) return QDX_ERR_ARG;
    if (rows == 0 || cols == 0) return 0;
    const int64_t n = rows * cols;
    qdx_copy_2d_kernel<<<(unsigned)((n + 255) / 256), 256, 0, S(stream)>>>(src, src_ld, dst, dst_ld, rows, cols);
    QDX_CHECK_LAUNCH();
    return 0;
}

int qdx_host_split(uint32_t k0, uint32_t k1, int32_t n, uint32_t* out) {
    if (n < 0 || (n > 0 && !out)) return QDX_ERR_ARG;
    for (int32_t i = 0; i < n; ++i) h_threefry2x32(k0, k1, 0u, (uint32_t)i, &out[2 * i], &out[2 * i + 1]);
    return 0;
}

// Generation keys {sel1, sel2, line, leaf} (2 words each) for the key handed to MAPElites.update (key_mode 1),
// one scan_update step on carry_io (2; carry_io advanced), DistributedMAPElites.update (3), MixingEmitter.emit (4).
int qdx_host_generation_keys(int32_t key_mode, uint32_t k0, uint32_t k1, uint32_t* carry_io2, uint32_t* out_keys8) {
    if (!out_keys8 || key_mode < 1 || key_mode > 4 || (key_mode == 2 && !carry_io2)) return QDX_ERR_ARG;
    QdxKey key{k0, k1};
    if (key_mode == 2) {                                                    // map_elites.py:214
        const QdxKey c{carry_io2[0], carry_io2[1]};
        key = h_split(c, 1);
        const QdxKey n = h_split(c, 0);
        carry_io2[0] = n.a; carry_io2[1] = n.b;
    }
    QdxKey emit = key;
    if (key_mode == 1 || key_mode == 2) emit = h_split(h_split(key, 1), 1);    // map_elites.py:177, :241
    else if (key_mode == 3) emit = h_split(key, 1);                          // distributed_map_elites.py:124
    const QdxKey e0 = h_split(emit, 0), e1 = h_split(emit, 1), kv = h_split(emit, 2);   // standard_emitters.py:55
    const QdxKey sel1 = h_split(e0, 1), sel2 = h_split(e1, 1);               // uniform_selector.py:48
    const QdxKey line = h_split(kv, 1), leaf = h_split(h_split(kv, 0), 0);   // mutation_operators.py:205, :220
    const uint32_t w[8] = {sel1.a, sel1.b, sel2.a, sel2.b, line.a, line.b, leaf.a, leaf.b};
    for (int j = 0; j < 8; ++j) out_keys8[j] = w[j];
    return 0;
}

int qdx_select_indices(void* ws, uint32_t k0, uint32_t k1, int64_t num, int32_t* out, void* stream) {
    if (!ws || !out || num < 0) return QDX_ERR_ARG;
    if (num == 0) return 0;
    qdx_select_kernel<<<(unsigned)((num + 255) / 256), 256, 0, S(stream)>>>(ws, QdxKey{k0, k1}, num, out);
    QDX_CHECK_LAUNCH();
    return 0;
}

int qdx_select_indices_without_replacement(const float* rep_fitness, int64_t K, void* ws, uint32_t k0, uint32_t k1, int64_t num, float* scratch_K,
                                           int32_t* out, void* stream) {
    if (!rep_fitness || !ws || !scratch_K || !out || K <= 0 || num < 0) return QDX_ERR_ARG;
    if (num > K) return QDX_ERR_ARG;                 // "Cannot take a larger sample than population when 'replace=False'"
    if (K > (1ll << 18)) return QDX_ERR_UNSUPPORTED;  // rank by counting: K^2 comparisons
    if (num == 0) return 0;
    const unsigned g = (unsigned)((K + 255) / 256);
    qdx_gumbel_kernel<<<g, 256, 0, S(stream)>>>(rep_fitness, K, ws, QdxKey{k0, k1}, scratch_K);
    qdx_topk_rank_kernel<<<g, 256, 0, S(stream)>>>(scratch_K, K, num, out);
    QDX_CHECK_LAUNCH();
    return 0;
}

int qdx_gather_rows(const float* src, const int32_t* idx, int64_t B, int64_t D, float* out, void* stream) {
    if (!src || !idx || !out || B < 0 || D <= 0) return QDX_ERR_ARG;
    if (B == 0) return 0;
    qdx_gather_rows_kernel<<<(unsigned)((B * 32 + 255) / 256), 256, 0, S(stream)>>>(src, idx, B, (int32_t)D, out);
    QDX_CHECK_LAUNCH();
    return 0;
}

int qdx_isoline_variation(const float* x1, const float* x2, int64_t B, int64_t D, uint32_t k0, uint32_t k1, float iso_sigma,
                          float line_sigma, int32_t has_min, float minval, int32_t has_max, float maxval, float* out,
                          void* stream) {
    if (!x1 || !x2 || !out || B < 0 || D <= 0) return QDX_ERR_ARG;
    if (B == 0) return 0;
    const int64_t n = B * D;
    qdx_isoline_kernel<<<(unsigned)((n + 255) / 256), 256, 0, S(stream)>>>(x1, x2, B, (int32_t)D, QdxKey{k0, k1}, iso_sigma, line_sigma,
                                                                         has_min, minval, has_max, maxval, out);
    QDX_CHECK_LAUNCH();
    return 0;
}

int qdx_random(uint32_t k0, uint32_t k1, int64_t n, int32_t kind, float minval, float maxval, void* out, void* stream) {
    if (!out || n < 0 || kind < 0 || kind > 2) return QDX_ERR_ARG;
    if (n == 0) return 0;
    qdx_random_kernel<<<(unsigned)((n + 255) / 256), 256, 0, S(stream)>>>(QdxKey{k0, k1}, n, kind, minval, maxval, out);
    QDX_CHECK_LAUNCH();
    return 0;
}

int qdx_metrics(const float* rep_fitness, int64_t K, float qd_offset, float* out3, void* stream) {
    if (!rep_fitness || !out3 || K <= 0) return QDX_ERR_ARG;
    qdx_metrics_kernel<<<1, 256, 0, S(stream)>>>(rep_fitness, K, qd_offset, out3);
    QDX_CHECK_LAUNCH();
    return 0;
}

// host-only helper (no GPU): expands the selection segments to T[1..M] for tests of the closed form
int qdx_host_select_table(int32_t M, float* out_T, int32_t* out_nseg) {
    if (M <= 0 || !out_T) return QDX_ERR_ARG;
    QdxSel sel;
    qdx_build_sel(M, &sel);
    if (sel.nseg <= 0) return QDX_ERR_UNSUPPORTED;
    for (int s = 0; s < sel.nseg; ++s)
        for (int i = 0; i <= sel.seg[s].n; ++i) out_T[sel.seg[s].j0 + i - 1] = qdx_seg_value(sel.seg[s], i);
    if (out_nseg) *out_nseg = sel.nseg;
    return 0;
}

// host-only helper: rank lookup through the segments (tests)
int qdx_host_select_rank(int32_t M, const float* r, int64_t n, int32_t* out_rank) {
    if (M <= 0 || !r || !out_rank) return QDX_ERR_ARG;
    QdxSel sel;
    qdx_build_sel(M, &sel);
    if (sel.nseg <= 0) return QDX_ERR_UNSUPPORTED;
    for (int64_t i = 0; i < n; ++i) out_rank[i] = qdx_sel_rank(sel.seg, sel.last, sel.nseg, r[i]);
    return 0;
}

}  // extern "C"
