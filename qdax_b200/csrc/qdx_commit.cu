// Streaming commit (stage d of the generation): MapElitesRepertoire.add after the per-cell election
// (qdax/core/containers/mapelites_repertoire.py:217-266 under /root/reference) + default_qd_metrics
// (qdax/utils/metrics.py:74-98) + the parent-selection tables of the NEXT generation.
//
// HBM-bound byte work: per changed cell one offspring row (4 D bytes) is read and one repertoire row written.  The
// rows never touch registers: one elected lane per warp drives a 4-stage ring of shared-memory buffers with the
// bulk-copy engine -- cp.async.bulk global->shared (completion on an mbarrier), then cp.async.bulk shared->global --
// two loads and two stores of up to 4 KB in flight per warp, 12 warps per SM.  Everything around the row traffic is
// arranged so that nothing serial is left behind it (per-phase globaltimer stamps: -DQDX_COMMIT_TRACE=1):
//   * phase 1, thread = cell over contiguous blocks of cells (cooperative launch: all CTAs are co-resident): coalesced
//     key / fitness pass, winner descriptors copied with all loads ahead of all stores, changed cells appended to ONE
//     grid-wide job list; every CTA then publishes its partial metrics and its count of occupied cells;
//   * grid barrier (one fence, one atomic, one spin per CTA);
//   * phase 2, the row copies, by guided self-scheduling over the job list (static first batch, then batches from a
//     grid-wide counter that shrink with the work left), so the SMs run out of work together although their distance
//     to the memory differs; measured streaming rate = the HBM copy peak;
//   * a service CTA that owns no cells sums the published metrics and rebuilds the selection segments (qdx_build_sel, a
//     serial ~8 us computation when the number of occupied cells changed) while the others stream;
//   * afterwards every CTA writes its slice of the ordered occupied-cell list (offset = sum of its predecessors'
//     counts, one batched read) and the last CTA to finish (ticket) re-arms the workspace counters.
#include <cstring>
#include "qdx_common.cuh"
#include "../../include/qdx.h"

namespace {

#ifndef QDX_COMMIT_CW
#define QDX_COMMIT_CW 4
#endif
#ifndef QDX_COMMIT_NST
#define QDX_COMMIT_NST 4
#endif
#ifndef QDX_COMMIT_LEAD
#define QDX_COMMIT_LEAD 2
#endif
constexpr int CW = QDX_COMMIT_CW;      // warps per CTA
constexpr int NST = QDX_COMMIT_NST;    // ring stages per warp
constexpr int LEAD = QDX_COMMIT_LEAD;  // loads run this many pieces ahead of the stores
#ifndef QDX_COMMIT_JB
#define QDX_COMMIT_JB 8
#endif
#ifndef QDX_COMMIT_GS
#define QDX_COMMIT_GS 2
#endif
#ifndef QDX_COMMIT_GD
#define QDX_COMMIT_GD 2
#endif
constexpr unsigned GS = QDX_COMMIT_GS;  // static first batch = list / (GS * warps of the grid)
constexpr unsigned GD = QDX_COMMIT_GD;  // dynamic batch = remaining list / (GD * warps of the grid)
constexpr int JB = QDX_COMMIT_JB;      // most list entries (changed cells) per grab from the grid-wide counter (<= 32)
#ifndef QDX_COMMIT_CHUNK
#define QDX_COMMIT_CHUNK 4096
#endif
#ifndef QDX_COMMIT_OWN_HALF
#define QDX_COMMIT_OWN_HALF 1  // A/B switch: 0 = every winner goes through the grid-wide list (round-1 behaviour)
#endif
#ifndef QDX_COMMIT_OWN_NUM
#define QDX_COMMIT_OWN_NUM 2   // quarters of a warp's winners kept by the warp
#endif
#ifndef QDX_COMMIT_EXP
#define QDX_COMMIT_EXP 0      // timing experiments only: 1 = loads without stores, 2 = no row traffic at all
#endif
constexpr int CHUNK = QDX_COMMIT_CHUNK;        // bytes per stage (a row of D <= 1024 floats in one piece; longer rows in pieces)
#ifndef QDX_COMMIT_TRACE
#define QDX_COMMIT_TRACE 0    // timing experiments only: per-phase globaltimer stamps (tools/time_insert.py --trace)
#endif
#if QDX_COMMIT_TRACE
__device__ unsigned long long g_commit_trace[8];
#define QDX_TRACE_MIN(i) do { if (threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); atomicMin(&g_commit_trace[i], t_); } } while (0)
#define QDX_TRACE_MAX(i) do { if (threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); atomicMax(&g_commit_trace[i], t_); } } while (0)
#else
#define QDX_TRACE_MIN(i)
#define QDX_TRACE_MAX(i)
#endif
constexpr int MAX_SLABS = 64;      // occupancy ballots kept in shared memory for the list pass (block <= 8192 cells)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "CW_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra CW_DONE;\n\t"
        "bra CW_LOOP;\n\t"
        "CW_DONE:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n\tcp.async.bulk.commit_group;"
                 ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

// Spin until CTA b has published its count for launch `seq`; bounded (2 s) so that a scheduling accident can never hang
// the GPU: on timeout the sticky error flag is raised and 0 is returned.
__device__ __forceinline__ uint32_t wait_count(QdxWorkspace* ws, int b, uint32_t seq) {
    unsigned long long v = *(volatile unsigned long long*)&ws->occ_pub[b];
    if ((uint32_t)(v >> 32) == seq) return (uint32_t)v;
    unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        __nanosleep(32);
        v = *(volatile unsigned long long*)&ws->occ_pub[b];
        if ((uint32_t)(v >> 32) == seq) return (uint32_t)v;
        unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 2000000000ull) { qdx_set_error(ws, QDX_ERR_INTERNAL); return 0u; }
    }
}

// Per-thread partial sum of the occupied counts of CTAs [0, limit): the first reads go out together (one memory round
// trip for up to 4 counts per thread); only a count that has not been published yet is waited for.
__device__ __forceinline__ int sum_counts(QdxWorkspace* ws, int limit, uint32_t seq, int t, int nt) {      // thread t of nt
    int part = 0;
    for (int b0 = t; b0 < limit; b0 += 4 * nt) {
        unsigned long long v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int b = b0 + j * nt;
            v[j] = b < limit ? *(volatile unsigned long long*)&ws->occ_pub[b] : ((unsigned long long)seq << 32);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int b = b0 + j * nt;
            if ((uint32_t)(v[j] >> 32) != seq) v[j] = (unsigned long long)wait_count(ws, b, seq);
            part += (int)(uint32_t)v[j];
        }
    }
    return part;
}

struct CommitParams {
    void* ws; int64_t K; int32_t D; int32_t Dd;
    const float* off_g; const float* off_f; const float* off_d; uint32_t idx_base; int64_t B; int32_t first_wins;
    float* rep_g; float* rep_f; float* rep_d; float qd_offset; float* metrics_out; int32_t* added_cells; int32_t mode;
};

// One winner descriptor row, off_d[i] -> rep_d[c]: all loads first, then all stores (a load / store pair per element
// would serialise on the possible aliasing of the two arrays: one memory round trip per element).
__device__ __forceinline__ void copy_desc(const float* __restrict__ src, float* __restrict__ dst, int Dd, bool vec4) {
    if (vec4) {
        for (int d0 = 0; d0 < Dd; d0 += 32) {
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) if (d0 + 4 * j < Dd) v[j] = __ldg((const float4*)(src + d0) + j);
#pragma unroll
            for (int j = 0; j < 8; ++j) if (d0 + 4 * j < Dd) ((float4*)(dst + d0))[j] = v[j];
        }
    } else {
        for (int d0 = 0; d0 < Dd; d0 += 8) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) if (d0 + j < Dd) v[j] = __ldg(src + d0 + j);
#pragma unroll
            for (int j = 0; j < 8; ++j) if (d0 + j < Dd) dst[d0 + j] = v[j];
        }
    }
}

// Sum of the per-CTA partial metrics in CTA order (deterministic), by all threads of one CTA; result valid on thread 0.
struct MetricsAcc { double s; float m; int n, nn, a; };
__device__ __forceinline__ MetricsAcc reduce_partials(QdxWorkspace* ws, int nparts, double* s_sum, float* s_max, int* s_cnt, int* s_nan, int* s_add) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double s = 0.0; float m = -INFINITY; int n = 0, nn = 0, a = 0;
    for (int b = tid; b < nparts; b += CW * 32) {
        s += *(volatile double*)&ws->part_sum[b]; m = fmaxf(m, *(volatile float*)&ws->part_max[b]);
        n += *(volatile int32_t*)&ws->part_cnt[b]; nn |= *(volatile int32_t*)&ws->part_nan[b]; a += *(volatile int32_t*)&ws->part_add[b];
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o); m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        n += __shfl_xor_sync(0xffffffffu, n, o); nn |= __shfl_xor_sync(0xffffffffu, nn, o); a += __shfl_xor_sync(0xffffffffu, a, o);
    }
    __syncthreads();
    if (lane == 0) { s_sum[wid] = s; s_max[wid] = m; s_cnt[wid] = n; s_nan[wid] = nn; s_add[wid] = a; }
    __syncthreads();
    MetricsAcc r{0.0, -INFINITY, 0, 0, 0};
    if (tid == 0)
        for (int w = 0; w < CW; ++w) { r.s += s_sum[w]; r.m = fmaxf(r.m, s_max[w]); r.n += s_cnt[w]; r.nn |= s_nan[w]; r.a += s_add[w]; }
    return r;
}

// mode 0: commit winners whose offspring rows are in off_* (index = global idx - idx_base), reset keys, metrics, tables
// mode 1: stage -- copy only the winners owned by [idx_base, idx_base + B) into rep_* (= staging rows by cell), keep
//         the key table, nothing else                                        (multi-GPU winners-only exchange)
// mode 2: apply -- off_* are staging rows indexed by CELL; reset keys, metrics, tables
__global__ void __launch_bounds__(CW * 32) qdx_commit_stream_kernel(const CommitParams p) {
    extern __shared__ __align__(128) unsigned char s_stage[];        // [CW][NST][CHUNK]
    __shared__ uint64_t s_bar[CW * NST];
    __shared__ unsigned long long s_dst[CW * NST];                   // destination of the piece sitting in each stage
    __shared__ uint32_t s_bytes[CW * NST];
    __shared__ uint32_t s_occ[MAX_SLABS * CW];
    __shared__ int32_t s_base;
    __shared__ double s_sum[CW]; __shared__ float s_max[CW]; __shared__ int s_cnt[CW], s_nan[CW], s_add[CW], s_occn[CW];

    qdx_pdl_enter();         // keys, offspring rows and fitnesses read below are the previous kernel's output
    QdxWorkspace* ws = (QdxWorkspace*)p.ws;
    unsigned long long* keytab = qdx_ws_keytab(p.ws, p.K);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int mode = p.mode;
    const bool tail = mode != 1;
    const int nblk = (int)gridDim.x - 1;                     // streaming CTAs; CTA nblk is the service CTA
    const uint32_t seq = *(volatile uint32_t*)&ws->commit_seq + 1u;     // tag of this launch in occ_pub
    QDX_TRACE_MIN(0);                                         // first CTA starts
    // Multi-GPU apply after a peer timed out (qdx_elect_kernel raised QDX_ERR_PEER_TIMEOUT and the launch completed before
    // this one started, so every CTA reads the same value): nothing is applied, no table is cleared, the epoch stays.
    if (mode == 2 && *(volatile int32_t*)&ws->error == QDX_ERR_PEER_TIMEOUT) {
        if (p.metrics_out && blockIdx.x == 0 && tid < 4) p.metrics_out[tid] = __int_as_float(0x7fc00000);
        return;
    }

    if ((int)blockIdx.x == nblk) {
        // ---- service CTA: everything that needs the whole grid's phase-1 results but not the row traffic -- metrics
        // and the selection segments of the next generation -- runs here, behind the streaming of the other CTAs
        if (tail) {
            int part = sum_counts(ws, nblk, seq, tid, CW * 32);
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (lane == 0) s_occn[wid] = part;
            __threadfence();                                  // the partial metrics were published before the counts
            const MetricsAcc r = reduce_partials(ws, nblk, s_sum, s_max, s_cnt, s_nan, s_add);
            if (tid == 0) {
                float out[4];
                out[0] = (float)r.s + p.qd_offset * (float)r.n;            // qd_score   (metrics.py:92-93)
                out[1] = r.nn ? NAN : r.m;                                  // max_fitness (:95)
                out[2] = 100.0f * __fdiv_rn((float)r.n, (float)p.K);        // coverage   (:94)
                out[3] = (float)r.a;                                        // offspring inserted by this call
                for (int j = 0; j < 4; ++j) { ws->metrics[j] = out[j]; if (p.metrics_out) p.metrics_out[j] = out[j]; }
                QDX_TRACE_MAX(7);                                           // metrics written
                int M = 0;
                for (int w = 0; w < CW; ++w) M += s_occn[w];
                if (M != ws->sel.M || ws->sel.nseg <= 0) qdx_build_sel(M, &ws->sel);
            }
        }
    } else {
        const int64_t per_cta = (p.K + nblk - 1) / nblk;          // contiguous block of cells of this CTA
        const int64_t c_lo = (int64_t)blockIdx.x * per_cta < p.K ? (int64_t)blockIdx.x * per_cta : p.K;
        const int64_t c_hi = c_lo + per_cta < p.K ? c_lo + per_cta : p.K;
        const bool keep_bits = (c_hi - c_lo + CW * 32 - 1) / (CW * 32) <= MAX_SLABS;
        const bool vec4 = (p.Dd & 3) == 0 && ((((uintptr_t)p.off_d) | ((uintptr_t)p.rep_d)) & 15u) == 0;
        if (tid == 0) {
            s_base = 0;
            for (int s = 0; s < CW * NST; ++s) mbar_init(&s_bar[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        int32_t* job_cell = qdx_ws_jobs(p.ws, p.K);              // global list of changed cells: (cell, source row)
        int32_t* job_src = job_cell + p.K;
        double sum = 0.0; float mx = -INFINITY; int cnt = 0, nan = 0, added = 0;

        // ---- phase 1, thread = cell: election result, fitness / descriptor, key reset, metrics, occupancy; the changed
        // cells of the whole grid are appended to ONE global list (warp-aggregated atomics) that phase 2 deals out
        // A warp keeps the first half of its own winners (cell, source row in registers) and starts streaming them the moment its
        // CTA has arrived at the grid barrier, instead of waiting for the slowest CTA's phase 1 (~2 us of 9 before the first row
        // moved); only the other half goes on the grid-wide list, whose guided deal evens out what the static half leaves uneven.
        // (One slab of cells per CTA only: K <= 128 cells x CTAs.)
        const bool own_half = (c_hi - c_lo) <= (int64_t)(CW * 32) && QDX_COMMIT_OWN_HALF;
        int32_t st_cell = 0, st_src = 0; unsigned st_mask = 0u;
        int occ_after = 0, slab_i = 0;
        for (int64_t slab = c_lo; slab < c_hi; slab += CW * 32, ++slab_i) {
            const int64_t c = slab + tid;
            const bool in = c < c_hi;
            const unsigned long long key = in ? __ldcg(keytab + c) : 0ull;
            int64_t i = -1;
            if (key != 0ull && !qdx_key_is_nan(key)) {           // NaN-poisoned cells accept nobody
                if (mode == 2) i = c;
                else {
                    i = (int64_t)qdx_key_index(key, p.first_wins) - (int64_t)p.idx_base;
                    if (i < 0 || i >= p.B) { if (mode == 0) qdx_set_error(ws, QDX_ERR_BAD_INDEX); i = -1; }
                }
            }
            float fcell = -INFINITY;
            if (in) {
                if (i >= 0) {
                    fcell = __ldg(p.off_f + i);
                    copy_desc(p.off_d + i * p.Dd, p.rep_d + c * p.Dd, p.Dd, vec4);
                    p.rep_f[c] = fcell;
                    if (p.added_cells) p.added_cells[c] = (int32_t)i;
                    ++added;
                } else if (tail) {
                    fcell = __ldcg(p.rep_f + c);
                }
                if (key != 0ull && tail) keytab[c] = 0ull;
                if (fcell != -INFINITY) { sum += (double)fcell; ++cnt; }
                if (fcell != fcell) nan = 1; else if (fcell > mx) mx = fcell;
            }
            const unsigned ob = __ballot_sync(0xffffffffu, in && fcell != -INFINITY);
            if (keep_bits && lane == 0) s_occ[slab_i * CW + wid] = ob;
            occ_after += __popc(ob);
            unsigned wb = __ballot_sync(0xffffffffu, i >= 0);
            if (own_half) {
                const int keep = __popc(wb) * QDX_COMMIT_OWN_NUM / 4;
                const bool mine = i >= 0 && __popc(wb & ((1u << lane) - 1u)) < keep;
                st_mask = __ballot_sync(0xffffffffu, mine);
                if (mine) { st_cell = (int32_t)c; st_src = (int32_t)i; }
                wb &= ~st_mask;
            }
            unsigned base = 0;
            if (lane == 0 && wb) base = atomicAdd(&ws->job_count, (unsigned)__popc(wb));
            base = __shfl_sync(0xffffffffu, base, 0);
            if ((wb >> lane) & 1u) { const unsigned pos = base + __popc(wb & ((1u << lane) - 1u)); job_cell[pos] = (int32_t)c; job_src[pos] = (int32_t)i; }
        }
        // ---- this CTA's partial metrics and occupied count, published BEFORE the grid barrier: the service CTA sums them
        // while the rows stream
        for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o); nan |= __shfl_xor_sync(0xffffffffu, nan, o); added += __shfl_xor_sync(0xffffffffu, added, o);
        }
        if (lane == 0) { s_sum[wid] = sum; s_max[wid] = mx; s_cnt[wid] = cnt; s_nan[wid] = nan; s_add[wid] = added; s_occn[wid] = occ_after; }
        __syncthreads();                                      // the CTA's job-list entries and rows are ordered before thread 0's fence below
        QDX_TRACE_MIN(1); QDX_TRACE_MAX(2);                   // first / last CTA done with phase 1
        if (wid == 0) {
            // ---- warp 0: publish this CTA's partial metrics and occupied count (the service CTA sums them while the rows
            // stream), then the grid barrier (all CTAs are co-resident: cooperative launch): the job list is complete
            if (lane == 0) {
                int t = 0;
                if (tail) {
                    double s = 0.0; float m = -INFINITY; int n = 0, nn = 0, a = 0;
                    for (int w = 0; w < CW; ++w) { s += s_sum[w]; m = fmaxf(m, s_max[w]); n += s_cnt[w]; nn |= s_nan[w]; a += s_add[w]; t += s_occn[w]; }
                    ws->part_sum[blockIdx.x] = s; ws->part_max[blockIdx.x] = m; ws->part_cnt[blockIdx.x] = n; ws->part_nan[blockIdx.x] = nn;
                    ws->part_add[blockIdx.x] = a;
                }
                __threadfence();                              // ONE fence: partials + job list before the count and the arrival
                if (tail) *(volatile unsigned long long*)&ws->occ_pub[blockIdx.x] = ((unsigned long long)seq << 32) | (uint32_t)t;
                atomicAdd(&ws->cta_arrived, 1u);
            }
        }
        // ---- the row copies: the rows never touch registers -- one lane per warp drives a ring of shared-memory stages with the
        // bulk-copy engine (global -> shared on an mbarrier, then shared -> global)
        const uint32_t rowbytes = (uint32_t)p.D * 4u;
        const bool bulk = (p.D & 3) == 0;
        const int pieces = (int)((rowbytes + CHUNK - 1) / CHUNK);
        unsigned char* my_stage = s_stage + (size_t)wid * NST * CHUNK;
        uint64_t* bars = &s_bar[wid * NST];
        unsigned long long* sdst = &s_dst[wid * NST];
        uint32_t* sby = &s_bytes[wid * NST];
        uint32_t ql = 0, qs = 0;                                            // lane 0: pieces loaded / stored so far (ring positions)
        auto stream_row = [&](int32_t cell, int32_t src) {
            if (bulk) {
                if (lane == 0) {
                    for (int pc = 0; pc < pieces; ++pc) {
                        if (ql - qs >= (uint32_t)LEAD) {          // piece qs has been in flight long enough: send it on
                            const uint32_t sl = qs % NST;
                            mbar_wait(&bars[sl], (qs / NST) & 1u);
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            if (QDX_COMMIT_EXP != 1) bulk_s2g((void*)sdst[sl], my_stage + sl * CHUNK, sby[sl]);
                            ++qs;
                        }
                        const uint32_t sl = ql % NST;
                        const uint32_t bytes = rowbytes - (uint32_t)pc * CHUNK < (uint32_t)CHUNK ? rowbytes - (uint32_t)pc * CHUNK : (uint32_t)CHUNK;
                        asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NST - LEAD) : "memory");   // stage free again
                        sdst[sl] = (unsigned long long)((char*)p.rep_g + ((int64_t)cell * p.D) * 4 + (int64_t)pc * CHUNK);
                        sby[sl] = bytes;
                        mbar_expect_tx(&bars[sl], bytes);
                        bulk_g2s(my_stage + sl * CHUNK, (const char*)p.off_g + ((int64_t)src * p.D) * 4 + (int64_t)pc * CHUNK, bytes, &bars[sl]);
                        ++ql;
                    }
                }
                __syncwarp();
            } else {                                                  // D not a multiple of 4: plain per-lane copy
                const float* srow = p.off_g + (int64_t)src * p.D; float* drow = p.rep_g + (int64_t)cell * p.D;
                for (int d = lane; d < p.D; d += 32) drow[d] = srow[d];
            }
        };
        if (QDX_COMMIT_EXP != 2)
            for (unsigned m = st_mask; m; m &= m - 1u) {              // this warp's own half: no list, no barrier
                const int l = __ffs(m) - 1;
                stream_row(__shfl_sync(0xffffffffu, st_cell, l), __shfl_sync(0xffffffffu, st_src, l));
            }
        // ---- the grid barrier (all CTAs are co-resident: cooperative launch): the job list is complete
        if (wid == 0 && lane == 0) {
            unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            while (*(volatile unsigned*)&ws->cta_arrived < (unsigned)nblk) {
                unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 2000000000ull) { qdx_set_error(ws, QDX_ERR_INTERNAL); break; }
            }
            __threadfence();
        }
        __syncthreads();
        QDX_TRACE_MAX(3);                                     // last CTA through the grid barrier
        const bool aborted = *(volatile int32_t*)&ws->error == QDX_ERR_INTERNAL;    // a grid-barrier wait timed out: stream nothing more

        // ---- phase 2: the row copies, dealt out in batches of list entries.  SMs differ in their distance to the memory
        // they read and write, so a static deal leaves the slow ones streaming long after the fast ones are done (measured:
        // first CTA done at 36 us, last at 49 us).  Guided self-scheduling instead: warp g of the grid starts on batch g
        // (static, no atomic), every further batch comes from a grid-wide counter and shrinks with the work that is left
        // (JB entries while plenty remain, 1 at the end), so the warps run out of work within a row or two of each other.
        // The grab for the next batch is issued when a batch starts and consumed after its first row, and that batch's
        // list entries are fetched by the lanes then -- both latencies sit behind row traffic already in flight.
        const unsigned njobs = (QDX_COMMIT_EXP == 2 || aborted) ? 0u : *(volatile unsigned*)&ws->job_count;
        const unsigned g = blockIdx.x * CW + wid, G = (unsigned)nblk * CW;
        unsigned J0 = (njobs + GS * G - 1u) / (GS * G);                     // static first batch: about half of the list
        J0 = J0 < 1u ? 1u : (J0 > (unsigned)JB ? (unsigned)JB : J0);
        const unsigned dyn_base = G * J0;
        bool dyn = dyn_base < njobs;
        unsigned est = dyn_base;                                            // lane 0: where the counter stood at the last grab
        unsigned cur_base = g * J0, cur_n = cur_base < njobs ? (njobs - cur_base < J0 ? njobs - cur_base : J0) : 0u;
        int32_t my_cell = 0, my_src = 0;
        if ((unsigned)lane < cur_n) { my_cell = __ldcg(job_cell + cur_base + lane); my_src = __ldcg(job_src + cur_base + lane); }
        while (cur_n > 0u) {
            unsigned nxt_base = 0xFFFFFFFFu, nxt_n = 0u;
            if (dyn && lane == 0) {
                const unsigned rem = njobs > est ? njobs - est : 0u;
                unsigned want = rem / (GD * G);
                want = want < 1u ? 1u : (want > (unsigned)JB ? (unsigned)JB : want);
                nxt_base = dyn_base + atomicAdd(&ws->job_next, want);
                nxt_n = want;
            }
            int32_t nx_cell = 0, nx_src = 0;
            for (unsigned e = 0; e < cur_n; ++e) {
                stream_row(__shfl_sync(0xffffffffu, my_cell, (int)e), __shfl_sync(0xffffffffu, my_src, (int)e));
                if (e == 0u && dyn) {
                    nxt_base = __shfl_sync(0xffffffffu, nxt_base, 0); nxt_n = __shfl_sync(0xffffffffu, nxt_n, 0);
                    est = nxt_base + nxt_n;
                    if (nxt_base < njobs) {
                        if (njobs - nxt_base < nxt_n) nxt_n = njobs - nxt_base;
                        if ((unsigned)lane < nxt_n) { nx_cell = __ldcg(job_cell + nxt_base + lane); nx_src = __ldcg(job_src + nxt_base + lane); }
                    } else { nxt_n = 0u; dyn = false; }
                }
            }
            cur_base = nxt_base; cur_n = nxt_n; my_cell = nx_cell; my_src = nx_src;
        }
        if (bulk && lane == 0) {
            while (qs < ql) {
                const uint32_t sl = qs % NST;
                mbar_wait(&bars[sl], (qs / NST) & 1u);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (QDX_COMMIT_EXP != 1) bulk_s2g((void*)sdst[sl], my_stage + sl * CHUNK, sby[sl]);
                ++qs;
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        __syncwarp();
        QDX_TRACE_MIN(4); QDX_TRACE_MAX(5);                   // first / last CTA done streaming
        // ---- my slice of the ordered occupied-cell list: offset = sum of the predecessors' counts (all published before
        // the grid barrier, so this is one batched read; doing it before the barrier instead was measured 6-8 us slower --
        // thousands of threads polling for counts delay the CTAs that still have to publish theirs)
        if (tail) {
            int part = sum_counts(ws, (int)blockIdx.x, seq, tid, CW * 32);
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (lane == 0 && part) atomicAdd(&s_base, part);
            __syncthreads();
            int32_t* occ = qdx_ws_occ(p.ws);
            int pos = s_base;
            slab_i = 0;
            for (int64_t slab = c_lo; slab < c_hi; slab += CW * 32, ++slab_i) {
                unsigned bits[CW];
                if (keep_bits) {
#pragma unroll
                    for (int w = 0; w < CW; ++w) bits[w] = s_occ[slab_i * CW + w];
                } else {
#pragma unroll
                    for (int w = 0; w < CW; ++w) {
                        const int64_t c = slab + w * 32 + lane;
                        bits[w] = __ballot_sync(0xffffffffu, c < c_hi && __ldcg(p.rep_f + c) != -INFINITY);
                    }
                }
                int before = 0;
#pragma unroll
                for (int w = 0; w < CW; ++w) if (w < wid) before += __popc(bits[w]);
                const unsigned mine = bits[wid];
                if ((mine >> lane) & 1u) occ[pos + before + __popc(mine & ((1u << lane) - 1u))] = (int32_t)(slab + wid * 32 + lane);
#pragma unroll
                for (int w = 0; w < CW; ++w) pos += __popc(bits[w]);
            }
        }
    }
    // ---- the last CTA to finish re-arms the workspace for the next launch
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&ws->ticket, 1u) == gridDim.x - 1) {
            QDX_TRACE_MAX(6);                                     // end
            ws->ticket = 0u;
            ws->cta_arrived = 0u;
            ws->job_count = 0u;
            ws->job_next = 0u;
            ws->commit_seq = seq;
            if (mode == 2 && ws->xchg_nranks > 0) {                    // peer-memory exchange: next generation, other key table
                uint32_t* ep = (uint32_t*)((char*)ws->xchg_peer[ws->xchg_rank] + QDX_XCHG_EPOCH_OFFSET);
                *ep = *ep + 1u;
            }
        }
    }
}

// =====================================================================================================================
// Lean commit: the same insertion for the case where there is nothing to stream -- short rows and / or few winners, which
// is every steady-state generation of the 10^4-cell configurations (~20 winner rows of 400 B: the streaming kernel's
// cooperative launch, grid barrier, job list and service CTA are ~15 us of fixed latency around 8 KB of traffic).  Ordinary
// launch, thread = cell (one coalesced pass over keys and fitnesses), the winners of a warp's 32 cells are copied by the
// whole warp, four rows in flight; per-CTA partial metrics, the last CTA to finish (ticket) reduces them in CTA order and
// -- only when the set of occupied cells changed -- rebuilds the occupied-cell list and the selection segments.
//
// mode 0 / 2 as above.  mode 3 (multi-GPU, peer-memory exchange with offspring blocks): the winner of a cell is read
// straight out of its OWNER's offspring block (mapped with cudaIpc; NVLink loads), located by its global index
// rank * B_dev + i; the kernel first acquire-spins on this rank's arrival flags, so the key table is complete and every
// peer's rows have landed.  Every rank copies the same bits, so the replicas stay identical by construction.
// =====================================================================================================================
constexpr int LEAN_THREADS = 256;

template <typename T>
__device__ __forceinline__ T* shfl_ptr(T* p, int src) {
    return (T*)(uintptr_t)__shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)p, src);
}

__global__ void __launch_bounds__(LEAN_THREADS) qdx_commit_lean_kernel(const CommitParams p, int32_t) {
    qdx_pdl_enter();         // keys, offspring rows and fitnesses read below are the previous kernel's (generate / cells) output
    QdxWorkspace* ws = (QdxWorkspace*)p.ws;
    unsigned long long* keytab = qdx_ws_keytab(p.ws, p.K);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int mode = p.mode;
    __shared__ double s_sum[LEAN_THREADS / 32]; __shared__ float s_max[LEAN_THREADS / 32];
    __shared__ int s_cnt[LEAN_THREADS / 32], s_nan[LEAN_THREADS / 32], s_add[LEAN_THREADS / 32], s_new[LEAN_THREADS / 32];
    __shared__ bool s_last;
    __shared__ int32_t s_scan[33];
    int parity = 0, R = 1;
    bool dead = false;                                    // a peer timed out: apply nothing, keep the epoch
    if (mode == 3) {
        R = ws->xchg_nranks;
        parity = qdx_xchg_parity(p.ws);
        if (tid < R) {                                    // acquire-spin on the LOCAL arrival flags (bounded: wait_ms)
            const unsigned long long* flag = (const unsigned long long*)ws->xchg_peer[ws->xchg_rank] + tid;
            const uint32_t want = *(const volatile uint32_t*)((const char*)ws->xchg_peer[ws->xchg_rank] + QDX_XCHG_EPOCH_OFFSET) + 1u;
            unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            while ((int32_t)((uint32_t)qdx_ld_acquire_sys(flag) - want) < 0) {
                unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > (unsigned long long)(ws->xchg_timeout_ms > 0 ? ws->xchg_timeout_ms : 30000) * 1000000ull) { qdx_set_error(ws, QDX_ERR_PEER_TIMEOUT); break; }
                __nanosleep(32);
            }
        }
        __syncthreads();
        dead = *(volatile int32_t*)&ws->error == QDX_ERR_PEER_TIMEOUT;
    } else if (mode == 2) {
        dead = *(volatile int32_t*)&ws->error == QDX_ERR_PEER_TIMEOUT;
    }
    const bool vec_rows = (p.D & 3) == 0 && (((uintptr_t)p.rep_g) & 15u) == 0 && (mode == 3 || (((uintptr_t)p.off_g) & 15u) == 0);
    const int nq = p.D >> 2;
    double sum = 0.0; float mx = -INFINITY; int cnt = 0, nan = 0, added = 0, newly = 0;
    for (int64_t slab = (int64_t)blockIdx.x * LEAN_THREADS; slab < p.K && !dead; slab += (int64_t)gridDim.x * LEAN_THREADS) {
        const int64_t c = slab + tid;
        const bool in = c < p.K;
        const unsigned long long key = in ? __ldcg(keytab + c) : 0ull;
        const float cur = in ? __ldcg(p.rep_f + c) : -INFINITY;
        const float* sg = nullptr; const float* sd = nullptr; const float* sf = nullptr;
        int64_t i = -1;
        if (key != 0ull && !qdx_key_is_nan(key)) {           // NaN-poisoned cells accept nobody
            const int64_t idx = (int64_t)qdx_key_index(key, p.first_wins);
            if (mode == 2) { i = c; sg = p.off_g; sf = p.off_f; sd = p.off_d; }
            else if (mode == 3) {
                const int64_t r = idx / ws->xchg_bdev;
                if (r < R) { i = idx - r * ws->xchg_bdev; const QdxOffBlock ob = qdx_xchg_block(p.ws, p.K, (int)r, parity); sg = ob.g; sf = ob.f; sd = ob.d; }
                else qdx_set_error(ws, QDX_ERR_BAD_INDEX);
            } else {
                i = idx - (int64_t)p.idx_base; sg = p.off_g; sf = p.off_f; sd = p.off_d;
                if (i < 0 || i >= p.B) { qdx_set_error(ws, QDX_ERR_BAD_INDEX); i = -1; }
            }
        }
        float fcell = cur;
        if (i >= 0) {
            fcell = __ldcg(sf + i);
            const float* s = sd + i * p.Dd; float* d = p.rep_d + c * p.Dd;
            for (int d0 = 0; d0 < p.Dd; d0 += 8) {           // all loads of a group ahead of its stores
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) if (d0 + j < p.Dd) v[j] = __ldcg(s + d0 + j);
#pragma unroll
                for (int j = 0; j < 8; ++j) if (d0 + j < p.Dd) d[d0 + j] = v[j];
            }
            p.rep_f[c] = fcell;
            if (p.added_cells) p.added_cells[c] = (int32_t)i;
            ++added;
            if (cur == -INFINITY) ++newly;
        }
        if (key != 0ull) keytab[c] = 0ull;
        if (in) {
            if (fcell != -INFINITY) { sum += (double)fcell; ++cnt; }
            if (fcell != fcell) nan = 1; else if (fcell > mx) mx = fcell;
        }
        // ---- rows of this warp's winners: the whole warp copies them, up to four rows in flight
        unsigned wb = __ballot_sync(0xffffffffu, i >= 0);
        const float* my_src = i >= 0 ? sg + i * p.D : nullptr;
        float* my_dst = i >= 0 ? p.rep_g + c * p.D : nullptr;
        while (wb) {
            const float* src[4]; float* dst[4]; int n = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int l = wb ? __ffs(wb) - 1 : 0;
                src[j] = shfl_ptr(my_src, l); dst[j] = shfl_ptr(my_dst, l);
                if (wb) { ++n; wb &= wb - 1u; }
            }
            if (vec_rows) {
                for (int q0 = 0; q0 < nq; q0 += 32) {
                    const int q = q0 + lane;
                    float4 v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (j < n && q < nq) v[j] = __ldcg(reinterpret_cast<const float4*>(src[j]) + q);
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (j < n && q < nq) reinterpret_cast<float4*>(dst[j])[q] = v[j];
                }
            } else {
                for (int j = 0; j < n; ++j)
                    for (int d = lane; d < p.D; d += 32) dst[j][d] = __ldcg(src[j] + d);
            }
        }
    }
    // ---- metrics: per-CTA partials, summed in CTA order by the last CTA to finish (deterministic)
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o); nan |= __shfl_xor_sync(0xffffffffu, nan, o);
        added += __shfl_xor_sync(0xffffffffu, added, o); newly += __shfl_xor_sync(0xffffffffu, newly, o);
    }
    if (lane == 0) { s_sum[wid] = sum; s_max[wid] = mx; s_cnt[wid] = cnt; s_nan[wid] = nan; s_add[wid] = added; s_new[wid] = newly; }
    __syncthreads();
    if (tid == 0) {
        double s = 0.0; float m = -INFINITY; int n = 0, nn = 0, a = 0, nw = 0;
        for (int w = 0; w < LEAN_THREADS / 32; ++w) { s += s_sum[w]; m = fmaxf(m, s_max[w]); n += s_cnt[w]; nn |= s_nan[w]; a += s_add[w]; nw += s_new[w]; }
        ws->part_sum[blockIdx.x] = s; ws->part_max[blockIdx.x] = m; ws->part_cnt[blockIdx.x] = n; ws->part_nan[blockIdx.x] = nn;
        ws->part_add[blockIdx.x] = a; ws->part_new[blockIdx.x] = nw;
        __threadfence();
        s_last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double s = 0.0; float m = -INFINITY; int n = 0, nn = 0, a = 0, nw = 0;
    for (unsigned b = tid; b < gridDim.x; b += LEAN_THREADS) {
        s += *(volatile double*)&ws->part_sum[b]; m = fmaxf(m, *(volatile float*)&ws->part_max[b]); n += *(volatile int32_t*)&ws->part_cnt[b];
        nn |= *(volatile int32_t*)&ws->part_nan[b]; a += *(volatile int32_t*)&ws->part_add[b]; nw += *(volatile int32_t*)&ws->part_new[b];
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o); m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o)); n += __shfl_xor_sync(0xffffffffu, n, o);
        nn |= __shfl_xor_sync(0xffffffffu, nn, o); a += __shfl_xor_sync(0xffffffffu, a, o); nw += __shfl_xor_sync(0xffffffffu, nw, o);
    }
    __syncthreads();
    if (lane == 0) { s_sum[wid] = s; s_max[wid] = m; s_cnt[wid] = n; s_nan[wid] = nn; s_add[wid] = a; s_new[wid] = nw; }
    __syncthreads();
    int total_cnt = 0, total_new = 0;
    for (int w = 0; w < LEAN_THREADS / 32; ++w) { total_cnt += s_cnt[w]; total_new += s_new[w]; }
    if (tid == 0) {
        double ts = 0.0; float tm = -INFINITY; int tnn = 0, ta = 0;
        for (int w = 0; w < LEAN_THREADS / 32; ++w) { ts += s_sum[w]; tm = fmaxf(tm, s_max[w]); tnn |= s_nan[w]; ta += s_add[w]; }
        float out[4];
        out[0] = (float)ts + p.qd_offset * (float)total_cnt;            // qd_score   (metrics.py:92-93)
        out[1] = tnn ? NAN : tm;                                         // max_fitness (:95)
        out[2] = 100.0f * __fdiv_rn((float)total_cnt, (float)p.K);       // coverage   (:94)
        out[3] = (float)ta;                                              // offspring inserted by this call
        if (dead) for (int j = 0; j < 4; ++j) out[j] = __int_as_float(0x7fc00000);
        for (int j = 0; j < 4; ++j) { ws->metrics[j] = out[j]; if (p.metrics_out) p.metrics_out[j] = out[j]; }
        ws->ticket = 0u;
        if (!dead && (mode == 2 || mode == 3) && ws->xchg_nranks > 0) {  // peer-memory exchange: next generation, other tables / blocks
            uint32_t* ep = (uint32_t*)((char*)ws->xchg_peer[ws->xchg_rank] + QDX_XCHG_EPOCH_OFFSET);
            *ep = *ep + 1u;
        }
    }
    // the repertoire is final: leave the NEXT generation's parent selection ready.  The occupied-cell list only changes when a
    // cell turned from empty to occupied, or when the workspace has not seen this repertoire yet (M mismatch)
    if (!dead && (total_new != 0 || total_cnt != ws->sel.M || ws->sel.nseg <= 0)) qdx_cta_occupancy_scan(p.rep_f, p.K, p.ws, s_scan);
}

struct DeviceCaps { int sms; int ctas_per_sm; int coop; };

int device_caps(DeviceCaps* out) {
    static DeviceCaps cache[64];
    static bool have[64] = {false};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= 64 || !have[dev]) {
        DeviceCaps c;
        e = cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&c.coop, cudaDevAttrCooperativeLaunch, dev);
        const int smem = CW * NST * CHUNK;
        if (e == cudaSuccess) e = cudaFuncSetAttribute(qdx_commit_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.ctas_per_sm, qdx_commit_stream_kernel, CW * 32, smem);
        if (e != cudaSuccess) return (int)e;
        if (dev < 0 || dev >= 64) { *out = c; return 0; }
        cache[dev] = c; have[dev] = true;
    }
    *out = cache[dev];
    return 0;
}

}  // namespace

int qdx_launch_commit_generic(void* ws, int64_t K, int64_t D, int32_t desc_dim, const float* off_genotypes, const float* off_fitness,
                              const float* off_desc, uint32_t idx_base, int64_t B, int32_t first_wins, float* rep_genotypes,
                              float* rep_fitness, float* rep_desc, float qd_offset, float* metrics_out4, int32_t* added_cells,
                              int32_t mode, cudaStream_t stream);      // qdx_mapelites.cu: warp-per-cell kernel, ordinary launch

#if QDX_COMMIT_TRACE
extern "C" int qdx_debug_commit_trace(unsigned long long* out8, int reset) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess && out8) e = cudaMemcpyFromSymbol(out8, g_commit_trace, sizeof(unsigned long long) * 8);
    if (e == cudaSuccess && reset) {
        unsigned long long init[8] = {~0ull, ~0ull, 0ull, 0ull, ~0ull, 0ull, 0ull, 0ull};     // [7] = metrics written (service CTA)
        e = cudaMemcpyToSymbol(g_commit_trace, init, sizeof(init));
    }
    return (int)e;
}
#endif

extern "C" int qdx_commit(void* ws, int64_t K, int64_t D, int32_t desc_dim, const float* off_genotypes, const float* off_fitness,
                          const float* off_desc, uint32_t idx_base, int64_t B, int32_t first_wins, float* rep_genotypes,
                          float* rep_fitness, float* rep_desc, float qd_offset, float* metrics_out4, int32_t* added_cells,
                          int32_t mode, void* stream) {
    if (!ws || !rep_genotypes || !rep_fitness || !rep_desc) return QDX_ERR_ARG;
    if (mode != 3 && (!off_genotypes || !off_fitness || !off_desc)) return QDX_ERR_ARG;      // mode 3 reads the peers' offspring blocks
    if (K <= 0 || D <= 0 || desc_dim < 1 || B < 0 || mode < 0 || mode > 3) return QDX_ERR_ARG;
    CommitParams p;
    p.ws = ws; p.K = K; p.D = (int32_t)D; p.Dd = desc_dim; p.off_g = off_genotypes; p.off_f = off_fitness; p.off_d = off_desc;
    p.idx_base = idx_base; p.B = B; p.first_wins = first_wins; p.rep_g = rep_genotypes; p.rep_f = rep_fitness; p.rep_d = rep_desc;
    p.qd_offset = qd_offset; p.metrics_out = metrics_out4; p.added_cells = added_cells; p.mode = mode;
    // short rows (<= 1 KB) or peer sources: nothing to stream, the lean kernel (ordinary launch, thread = cell) is all latency saved
    if (mode == 3 || (mode != 1 && D <= 256)) {
        int64_t ctas = (K + LEAN_THREADS - 1) / LEAN_THREADS;
        if (ctas > QDX_MAX_COMMIT_CTAS) ctas = QDX_MAX_COMMIT_CTAS;
        return (int)qdx_launch_pdl(qdx_commit_lean_kernel, dim3((unsigned)ctas), dim3(LEAN_THREADS), 0, (cudaStream_t)stream, p, 0);
    }
    DeviceCaps caps;
    int rc = device_caps(&caps);
    if (rc) return rc;
    const bool aligned = (((uintptr_t)off_genotypes | (uintptr_t)rep_genotypes) & 15u) == 0;
#ifdef QDX_COMMIT_FORCE_GENERIC     // sanitizer experiment (tools/sanitize.sh): ordinary loads / stores instead of the bulk-copy ring
    caps.coop = 0;
#endif
    if (!caps.coop || caps.ctas_per_sm < 1 || ((D & 3) == 0 && !aligned))
        return qdx_launch_commit_generic(ws, K, D, desc_dim, off_genotypes, off_fitness, off_desc, idx_base, B, first_wins, rep_genotypes,
                                         rep_fitness, rep_desc, qd_offset, metrics_out4, added_cells, mode, (cudaStream_t)stream);
    int64_t nblk = (K + 31) / 32;                              // >= 32 cells per streaming CTA
    const int64_t cap = (int64_t)caps.sms * caps.ctas_per_sm - 1;
    if (nblk > cap) nblk = cap;
    if (nblk > QDX_MAX_COMMIT_CTAS - 1) nblk = QDX_MAX_COMMIT_CTAS - 1;
    if (nblk < 1) nblk = 1;
    // cooperative launch (the grid barrier needs co-residency) + programmatic dependent launch (its CTAs are scheduled while the
    // previous kernel drains; the first instruction waits for it).  If the runtime refuses the combination, cooperative only.
    static int pdl_ok = -1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(nblk + 1)); cfg.blockDim = dim3(CW * 32); cfg.dynamicSmemBytes = (size_t)CW * NST * CHUNK; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl_ok != 0 && qdx_pdl_enabled()) ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, qdx_commit_stream_kernel, p);
    if (e != cudaSuccess && cfg.numAttrs == 2) {
        (void)cudaGetLastError();
        pdl_ok = 0;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, qdx_commit_stream_kernel, p);
    } else if (e == cudaSuccess && cfg.numAttrs == 2) pdl_ok = 1;
    return e == cudaSuccess ? 0 : (int)e;
}
