// Shared device-side structures of the MAP-Elites generation step (B200 / sm_100a).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "qdx_math.cuh"
#include "qdx_select.cuh"

#define QDX_TASK_NONE (-1)
#define QDX_TASK_ARM 0
#define QDX_TASK_RASTRIGIN 1
#define QDX_TASK_SPHERE 2

#define QDX_MAX_GRID_DIM 4
#define QDX_MAX_AXES 4096   // total axis entries (sum of n_d) staged in shared memory

// Keys of one generation, derived from the key handed to MAPElites.update (qdax/core/map_elites.py:177,241;
// standard_emitters.py:55; uniform_selector.py:48; mutation_operators.py:205,220 under /root/reference).
struct QdxGenKeys {
    QdxKey sel1;    // split(split(emit,3)[0])[1]  -> parent-1 uniform draws
    QdxKey sel2;    // split(split(emit,3)[1])[1]  -> parent-2 uniform draws
    QdxKey line;    // split(split(emit,3)[2])[1]  -> line noise
    QdxKey leaf;    // split(split(split(emit,3)[2])[0], 1)[0] -> iso noise
};

// Pytree genotypes: an individual is the concatenation of its leaves, leaf l owning genes
// [off[l], off[l+1]) of the packed row; isoline_variation draws leaf l's noise from its own key with the counter
// i * size_l + j (mutation_operators.py:219-224: keys = split(key, nb_leaves), normal(key_l, x1_l.shape)).
#define QDX_MAX_LEAVES 32
struct QdxLeafTab {
    int32_t n;                              // 0 / 1 = single leaf (QdxGenKeys::leaf)
    int32_t off[QDX_MAX_LEAVES + 1];
    QdxKey key[QDX_MAX_LEAVES];
};
// leaf owning packed gene d (off[l] <= d < off[l+1]; empty leaves are skipped)
QDX_DEV int qdx_leaf_of(const QdxLeafTab& lt, int32_t d) {
    int lo = 0, hi = lt.n;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (lt.off[mid] <= d) lo = mid; else hi = mid; }
    return lo;
}

#define QDX_MAX_COMMIT_CTAS (148 * 8)
#ifndef QDX_COMMIT_CTAS
#define QDX_COMMIT_CTAS (148 * 8)      // grid cap of the commit kernel: measured better than one resident wave (148 * 4) for 4 KB rows (0.56 vs 0.42 of HBM peak)
#endif
#define QDX_MAX_PEERS 16      // ranks of one NVLink domain whose key tables are mapped into each other (cudaIpc)

// Device workspace header (one per repertoire); arrays follow at fixed offsets (qdx_ws_* below).
struct QdxWorkspace {
    QdxSel sel;             // selection segments of the current repertoire
    QdxGenKeys keys;        // keys of the current generation (device key chain)
    QdxKey carry;           // scan carry key (qdax/core/map_elites.py:213-214)
    float metrics[4];       // qd_score, max_fitness, coverage, num_added
    uint32_t ticket;        // last-CTA-done counter of the commit kernel
    int32_t error;          // sticky device-side error flag (e.g. empty repertoire)
    uint32_t push_ticket;   // last-CTA-done counter of the peer-memory push kernel
    uint32_t pad0;
    // ---- peer-memory exchange (multi-GPU, one process per GPU; qdx_xchg_* in include/qdx.h); 0 = not attached
    unsigned long long xchg_peer[QDX_MAX_PEERS];   // exchange buffer of every rank as mapped into THIS process
    int32_t xchg_rank;
    int32_t xchg_nranks;
    // offspring blocks of the exchange buffers (0 = none: the regen exchange): every rank keeps the rows of its fired offers
    // where its peers can read them, so the replicated insertion copies a winner straight from its owner over NVLink
    int64_t xchg_bdev;      // offspring per rank and generation
    int32_t xchg_D;
    int32_t xchg_Dd;
    int32_t xchg_timeout_ms;   // how long a consumer waits for its peers' arrival flags before it raises QDX_ERR_PEER_TIMEOUT
    int32_t pad1;
    // host mirror of `error` (pinned host memory, UVA; 0 = none): written by whoever raises the flag, so the host learns of a
    // device-side error without a blocking read-back and without a copy on the stream (qdx_workspace_set_error_mirror)
    int32_t* err_host;
    // per-CTA partial metrics of the commit kernel, summed in CTA order by the last CTA (deterministic)
    double part_sum[QDX_MAX_COMMIT_CTAS];
    float part_max[QDX_MAX_COMMIT_CTAS];
    int32_t part_cnt[QDX_MAX_COMMIT_CTAS];
    int32_t part_add[QDX_MAX_COMMIT_CTAS];
    int32_t part_nan[QDX_MAX_COMMIT_CTAS];
    int32_t part_new[QDX_MAX_COMMIT_CTAS];      // cells that turned from empty to occupied in this commit
    // streaming commit (qdx_commit.cu): per-CTA count of occupied cells, tagged with the launch sequence number
    unsigned long long occ_pub[QDX_MAX_COMMIT_CTAS];
    uint32_t commit_seq;
    uint32_t cta_arrived;   // grid barrier of the streaming commit (reset by its last CTA)
    uint32_t job_count;     // entries of the global list of changed cells (reset by the last CTA)
    uint32_t job_next;      // next batch of the list handed out to a streaming warp (reset by the last CTA)
    // generate kernel (persistent grid): next offspring row to hand out, CTAs done (both reset by the last CTA to finish)
    uint32_t gen_next_row;
    uint32_t gen_done;
};

// Raise the sticky device error flag (first error wins on the host mirror: later ones only overwrite the device copy).
__device__ __forceinline__ void qdx_set_error(void* ws_raw, int32_t code) {
    QdxWorkspace* ws = (QdxWorkspace*)ws_raw;
    ws->error = code;
    int32_t* h = ws->err_host;
    if (h) { *(volatile int32_t*)h = code; __threadfence_system(); }
}

__host__ __device__ inline size_t qdx_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
__host__ __device__ inline size_t qdx_ws_occ_offset() { return qdx_align_up(sizeof(QdxWorkspace), 256); }
__host__ __device__ inline size_t qdx_ws_keytab_offset(int64_t K) {
    return qdx_align_up(qdx_ws_occ_offset() + sizeof(int32_t) * (size_t)K, 256);
}
#define QDX_MAX_RANKS 64      // generation keys of up to 64 ranks ride in the tail of the key table (8 slots each)
__host__ __device__ inline size_t qdx_ws_jobs_offset(int64_t K) {      // (cell, source row) int32 pairs of the streaming commit
    return qdx_align_up(qdx_ws_keytab_offset(K) + sizeof(unsigned long long) * (size_t)(K + 8 * QDX_MAX_RANKS), 256);
}
__host__ __device__ inline size_t qdx_ws_total_bytes(int64_t K) {
    return qdx_align_up(qdx_ws_jobs_offset(K) + 2 * sizeof(int32_t) * (size_t)K, 256);
}
__host__ __device__ inline int32_t* qdx_ws_jobs(void* ws, int64_t K) { return (int32_t*)((char*)ws + qdx_ws_jobs_offset(K)); }
__host__ __device__ inline int32_t* qdx_ws_occ(void* ws) { return (int32_t*)((char*)ws + qdx_ws_occ_offset()); }
// Peer-memory exchange buffer of one rank (cudaMalloc'ed by qdx_xchg_create, mapped into every peer with cudaIpc):
//   [0, 512)             arrival flags: flag[r] = epoch + 1 once rank r's keys of that epoch have landed here
//   [512, 516)           epoch: generation counter of this rank (advanced by the commit kernel; lives here rather than in
//                        the workspace so that it survives when a cloned repertoire brings a fresh workspace)
//   then 2 key tables    (K keys + 8 * QDX_MAX_RANKS generation-key slots) x uint64, selected by epoch parity so a
//                        fast rank may push generation e+1 while this rank still consumes generation e
#define QDX_XCHG_EPOCH_OFFSET 512
__host__ __device__ inline size_t qdx_xchg_tab_entries(int64_t K) { return (size_t)K + 8 * QDX_MAX_RANKS; }
__host__ __device__ inline size_t qdx_xchg_tab_offset(int64_t K, int parity) {
    return 1024 + (size_t)parity * qdx_align_up(qdx_xchg_tab_entries(K) * sizeof(unsigned long long), 256);
}
// ... followed (B_dev > 0) by 2 offspring blocks, again selected by epoch parity: genotype rows (B_dev x D), fitness (B_dev),
// descriptors (B_dev x Dd) of this rank's offspring.  A rank starts generation e + 2 only after every rank has raised its
// flag of generation e + 1, i.e. after every rank has finished applying generation e -- so two blocks suffice.
__host__ __device__ inline size_t qdx_xchg_block_g_bytes(int64_t B_dev, int64_t D) { return qdx_align_up((size_t)B_dev * (size_t)D * 4, 256); }
__host__ __device__ inline size_t qdx_xchg_block_f_bytes(int64_t B_dev) { return qdx_align_up((size_t)B_dev * 4, 256); }
__host__ __device__ inline size_t qdx_xchg_block_bytes(int64_t B_dev, int64_t D, int64_t Dd) {
    return qdx_xchg_block_g_bytes(B_dev, D) + qdx_xchg_block_f_bytes(B_dev) + qdx_align_up((size_t)B_dev * (size_t)Dd * 4, 256);
}
__host__ __device__ inline size_t qdx_xchg_block_offset(int64_t K, int parity, int64_t B_dev, int64_t D, int64_t Dd) {
    return qdx_xchg_tab_offset(K, 2) + (size_t)parity * qdx_xchg_block_bytes(B_dev, D, Dd);
}
__host__ __device__ inline size_t qdx_xchg_total_bytes(int64_t K, int64_t B_dev = 0, int64_t D = 0, int64_t Dd = 0) {
    return B_dev > 0 ? qdx_xchg_block_offset(K, 2, B_dev, D, Dd) : qdx_xchg_tab_offset(K, 2);
}
// offspring block of rank `r` for the current generation, as mapped into THIS process
struct QdxOffBlock { float* g; float* f; float* d; };
__device__ __forceinline__ QdxOffBlock qdx_xchg_block(const void* ws_raw, int64_t K, int r, int parity) {
    const struct QdxWorkspace* ws = (const struct QdxWorkspace*)ws_raw;
    char* b = (char*)ws->xchg_peer[r] + qdx_xchg_block_offset(K, parity, ws->xchg_bdev, ws->xchg_D, ws->xchg_Dd);
    QdxOffBlock o;
    o.g = (float*)b;
    o.f = (float*)(b + qdx_xchg_block_g_bytes(ws->xchg_bdev, ws->xchg_D));
    o.d = (float*)(b + qdx_xchg_block_g_bytes(ws->xchg_bdev, ws->xchg_D) + qdx_xchg_block_f_bytes(ws->xchg_bdev));
    return o;
}
__device__ __forceinline__ int qdx_xchg_parity(const void* ws_raw) {
    const struct QdxWorkspace* ws = (const struct QdxWorkspace*)ws_raw;
    return (int)(*(const volatile uint32_t*)((const char*)ws->xchg_peer[ws->xchg_rank] + QDX_XCHG_EPOCH_OFFSET) & 1u);
}
// The insertion key table of the current generation: inside the workspace on one GPU, inside the exchange buffer
// (parity of the epoch) when the peer-memory exchange is attached.
__device__ __forceinline__ unsigned long long* qdx_ws_keytab(void* ws_raw, int64_t K) {
    const QdxWorkspace* ws = (const QdxWorkspace*)ws_raw;
    if (ws->xchg_nranks > 0) {
        char* base = (char*)ws->xchg_peer[ws->xchg_rank];
        return (unsigned long long*)(base + qdx_xchg_tab_offset(K, (int)(*(const uint32_t*)(base + QDX_XCHG_EPOCH_OFFSET) & 1u)));
    }
    return (unsigned long long*)((char*)ws_raw + qdx_ws_keytab_offset(K));
}

// Separable ("Euclidean grid") tessellation: centroid of cell sum_d idx_d*stride_d is (axis_0[idx_0], ...).
struct QdxGrid {
    int32_t dd;                          // 0 = not a grid (brute force)
    int32_t n[QDX_MAX_GRID_DIM];
    int32_t stride[QDX_MAX_GRID_DIM];
    int32_t off[QDX_MAX_GRID_DIM];       // offset of axis d inside `axes`
    float lo[QDX_MAX_GRID_DIM];          // fast path valid for lo <= x_d <= hi (else exact brute force per row)
    float hi[QDX_MAX_GRID_DIM];
    const float* axes;                   // device, sum n_d floats
    int32_t total_axes;
};

// Packed insertion key, 63 bits: (order_key(fitness) << 31) | (first_wins ? ~idx : idx) & 0x7FFFFFFF; 0 = empty slot.
// Bit 63 is never set, so unsigned order == signed order and the multi-GPU all-reduce(max) can run on int64.
QDX_DEV unsigned long long qdx_pack_key(float f, uint32_t idx, int first_wins) {
    return ((unsigned long long)qdx_order_key(f) << 31) | (unsigned long long)((first_wins ? ~idx : idx) & 0x7FFFFFFFu);
}
QDX_DEV bool qdx_key_is_nan(unsigned long long key) { return (uint32_t)(key >> 31) == 0xFFFFFFFFu; }
QDX_DEV uint32_t qdx_key_index(unsigned long long key, int first_wins) {
    const uint32_t lo = (uint32_t)key & 0x7FFFFFFFu;
    return first_wins ? (~lo & 0x7FFFFFFFu) : lo;
}

// Offer offspring `idx` with fitness f to cell c (MapElitesRepertoire.add, mapelites_repertoire.py:211-231):
// only candidates that can change the outcome touch the table (NaN poisons its cell; f <= current never wins
// and never blocks a winner because any winner has f > current >= f).
// With the peer-memory exchange attached (multi-GPU), an offer that becomes its cell's best so far on THIS rank is
// max-merged straight into every peer's table as well: system-scope 64-bit atomicMax into the peer's HBM over NVLink,
// issued from inside the generate / cells kernels, so the exchange overlaps the compute and needs no pass of its own.
// The global best of a cell is some rank's final local best, and a rank's final local best is always an improving
// record, so every table ends up holding the global maximum.  (~ln(offers per cell) records per cell at cold start,
// about one per changed cell in steady state.)  The fence makes the remote atomics visible before this thread's warp
// reports done (qdx_xchg_warp_done).
// Returns true when the offer was made (the offspring may be elected: its row must exist for the commit).
QDX_DEV bool qdx_offer(void* ws_raw, int64_t K, const float* rep_f, int32_t c, float f, uint32_t idx, int first_wins) {
    const float cur = __ldg(rep_f + c);
    const bool fire = (f != f) || f > cur;
    if (fire) {
        const QdxWorkspace* ws = (const QdxWorkspace*)ws_raw;
        const unsigned long long key = qdx_pack_key(f, idx, first_wins);
        const int R = ws->xchg_nranks;
        if (R > 0) {
            const int me = ws->xchg_rank;
            const size_t off = qdx_xchg_tab_offset(K, (int)(*(const uint32_t*)((const char*)ws->xchg_peer[me] + QDX_XCHG_EPOCH_OFFSET) & 1u));
            const unsigned long long old = atomicMax_system((unsigned long long*)((char*)ws->xchg_peer[me] + off) + c, key);
            if (key > old) {
                for (int q = 0; q < R; ++q)
                    if (q != me) atomicMax_system((unsigned long long*)((char*)ws->xchg_peer[q] + off) + c, key);
                __threadfence_system();
            }
        } else {
            atomicMax((unsigned long long*)((char*)ws_raw + qdx_ws_keytab_offset(K)) + c, key);
        }
    }
    return fire;
}

// Occupancy scan by ONE CTA (any multiple of 32 threads, <= 1024): ordered list of occupied cells -> occ[], M, and the
// selection segments (rebuilt only when M changed).  Warp w owns the contiguous cell range [w*chunk, (w+1)*chunk);
// 32 cells per step, ballot + popc, every load coalesced and independent of the previous step.  Used by the
// prepare kernel and by the last CTA of the commit kernel (which leaves the NEXT generation's selection ready, so a
// steady-state generation needs no prepare launch).  s_warp: >= 33 int32 of shared memory.
static __device__ __noinline__ void qdx_cta_occupancy_scan(const float* rep_f, int64_t K, void* ws_raw, int32_t* s_warp) {
    QdxWorkspace* ws = (QdxWorkspace*)ws_raw;
    int32_t* occ = qdx_ws_occ(ws_raw);
    constexpr int BAL = 2048;                       // ballot words kept in shared memory: K <= 65536 needs no second read
    __shared__ uint32_t s_bal[BAL];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5, nw = blockDim.x >> 5;
    const int64_t chunk = ((K + nw - 1) / nw + 31) / 32 * 32;
    const int64_t lo = (int64_t)w * chunk, hi = lo + chunk < K ? lo + chunk : K;
    const bool keep = (K + 31) / 32 <= BAL;
    // every step's load is independent of the previous step's ballot: 16 loads in flight per lane (the fitness array is
    // L2-resident, so the scan is latency- not bandwidth-bound)
    constexpr int U = 16;
    int32_t cnt = 0;
    for (int64_t c0 = lo; c0 < hi; c0 += 32 * U) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const int64_t c = c0 + 32 * u + lane; v[u] = c < hi ? __ldcg(rep_f + c) : -INFINITY; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned b = __ballot_sync(0xffffffffu, v[u] != -INFINITY);
            cnt += __popc(b);
            if (keep && lane == 0 && c0 + 32 * u < hi) s_bal[(c0 >> 5) + u] = b;
        }
    }
    __syncthreads();                       // s_warp may still be in use by the caller
    if (lane == 0) s_warp[w] = cnt;
    __syncthreads();
    if (w == 0) {
        int32_t v = lane < nw ? s_warp[lane] : 0, x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        __syncwarp();
        s_warp[lane] = x - v;
        if (lane == 31) s_warp[32] = x;
    }
    __syncthreads();
    int32_t pos = s_warp[w];
    if (keep) {
        for (int64_t c0 = lo; c0 < hi; c0 += 32) {
            const unsigned b = s_bal[c0 >> 5];
            if ((b >> lane) & 1u) occ[pos + __popc(b & ((1u << lane) - 1u))] = (int32_t)(c0 + lane);
            pos += __popc(b);
        }
    } else {
        for (int64_t c0 = lo; c0 < hi; c0 += 32 * U) {
            float v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { const int64_t c = c0 + 32 * u + lane; v[u] = c < hi ? __ldcg(rep_f + c) : -INFINITY; }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool o = v[u] != -INFINITY;
                const unsigned b = __ballot_sync(0xffffffffu, o);
                if (o) occ[pos + __popc(b & ((1u << lane) - 1u))] = (int32_t)(c0 + 32 * u + lane);
                pos += __popc(b);
            }
        }
    }
    if (t == 0) {
        const int32_t M = s_warp[32];
        if (M != ws->sel.M || ws->sel.nseg <= 0) qdx_build_sel(M, &ws->sel);
    }
}


// ---- programmatic dependent launch (generate -> lean commit -> generate ...): the next kernel of the generation chain is
// launched with cudaLaunchAttributeProgrammaticStreamSerialization, so its CTAs are scheduled while the previous kernel
// drains and only its first instruction waits for the previous grid to complete and flush (griddepcontrol.wait) -- the
// ~3-4 us launch gap between two dependent kernels disappears from a ~100 us generation.  QDX_PDL=0 in the environment
// turns the attribute off (A/B); kernels launched without it see both instructions as no-ops.
QDX_DEV void qdx_pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
static inline bool qdx_pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("QDX_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}
template <typename... KArgs, typename... Args>
static inline cudaError_t qdx_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = qdx_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

#ifndef QDX_XCHG_TRACE
#define QDX_XCHG_TRACE 0      // timing experiments only: where a multi-GPU generation's tail goes (qdx_debug_xchg_trace)
#endif
#if QDX_XCHG_TRACE
// [0] launches of the elect kernel, [1] sum ns waiting for the peers' flags, [2] sum ns elect kernel (CTA 0), [3] publishes,
// [4] sum ns publish (keys + fence + flags to every peer)
static __device__ unsigned long long g_xchg_trace[8];      // per translation unit; read from qdx_mapelites.cu
QDX_DEV unsigned long long qdx_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif

QDX_DEV void qdx_st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
QDX_DEV unsigned long long qdx_ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

QDX_DEV void qdx_st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// Publication of "rank `me`, epoch e: all my keys (and offspring rows) have landed": the arrival flag is raised in every
// peer.  One thread, after it has observed that every CTA of the offering kernel is done (ticket + fence): everything those
// CTAs wrote -- key pushes into the peers, rows / fitnesses / descriptors in this rank's offspring block -- is ordered
// before the flags by ONE system-scope fence (cumulativity), followed by relaxed system-scope stores; a consumer that
// acquires the flag sees all of it.  Keys-only exchange (no offspring blocks: the winners are REGENERATED by every rank
// from (owner's keys, local index)): this rank's generation keys go to slot `me` of every rank's table first.
QDX_DEV void qdx_xchg_publish(QdxWorkspace* ws, int64_t K, const QdxGenKeys& keys) {
    const int R = ws->xchg_nranks, me = ws->xchg_rank;
    const uint32_t epoch = *(const uint32_t*)((const char*)ws->xchg_peer[me] + QDX_XCHG_EPOCH_OFFSET);
    if (ws->xchg_bdev == 0) {
        const size_t off = qdx_xchg_tab_offset(K, (int)(epoch & 1u));
        const uint32_t w[8] = {keys.sel1.a, keys.sel1.b, keys.sel2.a, keys.sel2.b, keys.line.a, keys.line.b, keys.leaf.a, keys.leaf.b};
        for (int q = 0; q < R; ++q) {
            unsigned long long* slot = (unsigned long long*)((char*)ws->xchg_peer[q] + off) + K + 8 * me;
#pragma unroll
            for (int j = 0; j < 8; ++j) slot[j] = (unsigned long long)w[j];
        }
    }
    __threadfence_system();
    for (int q = 0; q < R; ++q) qdx_st_relaxed_sys((unsigned long long*)ws->xchg_peer[q] + me, (unsigned long long)(epoch + 1u));
}
