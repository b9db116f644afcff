// Streaming commit (stage d of the generation): MapElitesRepertoire.add after the per-cell election
// (qdax/core/containers/mapelites_repertoire.py:217-266 under /root/reference) + default_qd_metrics
// (qdax/utils/metrics.py:74-98) + the parent-selection tables of the NEXT generation.
//
// HBM-bound byte work: per changed cell one offspring row (4 D bytes) is read and one repertoire row written.  The
// rows never touch registers: one elected lane per warp drives a 4-stage ring of shared-memory buffers with the
// bulk-copy engine -- cp.async.bulk global->shared (completion on an mbarrier), then cp.async.bulk shared->global --
// two loads and two stores of up to 4 KB in flight per warp, 12 warps per SM.  Everything around the row traffic is
// arranged so that nothing serial is left behind it:
//   * cells are split into contiguous blocks, one per CTA (cooperative launch: all CTAs are co-resident); thread = cell
//     for the coalesced key / fitness pass, winners compacted into a shared list, warps take them round-robin;
//   * every CTA first publishes how many of its cells are occupied after this commit; later it derives its offset in
//     the ordered occupied-cell list from its predecessors' counts and writes its slice (distributed scan, no tail);
//   * one extra CTA owns nothing but the selection segments: it sums all counts and rebuilds them (qdx_build_sel, a
//     serial ~8 us computation when the number of occupied cells changed) while the others stream rows;
//   * the last CTA to finish (ticket) only reduces <= 593 partial metrics.
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
constexpr int LEAD = QDX_COMMIT_LEAD;  // loads run this many jobs ahead of the stores
#ifndef QDX_COMMIT_CHUNK
#define QDX_COMMIT_CHUNK 4096
#endif
#ifndef QDX_COMMIT_EXP
#define QDX_COMMIT_EXP 0      // timing experiments only: 1 = loads without stores, 2 = no row traffic at all
#endif
constexpr int CHUNK = QDX_COMMIT_CHUNK;        // bytes per stage (a row of D <= 1024 floats in one piece; longer rows in pieces)
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
        if (t1 - t0 > 2000000000ull) { ws->error = QDX_ERR_INTERNAL; return 0u; }
    }
}

struct CommitParams {
    void* ws; int64_t K; int32_t D; int32_t Dd;
    const float* off_g; const float* off_f; const float* off_d; uint32_t idx_base; int64_t B; int32_t first_wins;
    float* rep_g; float* rep_f; float* rep_d; float qd_offset; float* metrics_out; int32_t* added_cells; int32_t mode;
};

// mode 0: commit winners whose offspring rows are in off_* (index = global idx - idx_base), reset keys, metrics, tables
// mode 1: stage -- copy only the winners owned by [idx_base, idx_base + B) into rep_* (= staging rows by cell), keep
//         the key table, nothing else                                        (multi-GPU winners-only exchange)
// mode 2: apply -- off_* are staging rows indexed by CELL; reset keys, metrics, tables
__global__ void __launch_bounds__(CW * 32) qdx_commit_stream_kernel(const CommitParams p) {
    extern __shared__ __align__(128) unsigned char s_stage[];        // [CW][NST][CHUNK]
    __shared__ uint64_t s_bar[CW * NST];
    __shared__ uint32_t s_occ[MAX_SLABS * CW];
    __shared__ int32_t s_base;
    __shared__ double s_sum[CW]; __shared__ float s_max[CW]; __shared__ int s_cnt[CW], s_nan[CW], s_add[CW];
    __shared__ bool s_last;

    QdxWorkspace* ws = (QdxWorkspace*)p.ws;
    unsigned long long* keytab = qdx_ws_keytab(p.ws, p.K);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int mode = p.mode;
    const bool tail = mode != 1;
    const int nblk = (int)gridDim.x - 1;                     // streaming CTAs; CTA nblk only builds the selection segments
    const uint32_t seq = *(volatile uint32_t*)&ws->commit_seq + 1u;     // tag of this launch in occ_pub
    double sum = 0.0; float mx = -INFINITY; int cnt = 0, nan = 0, added = 0;

    if ((int)blockIdx.x == nblk) {
        if (tail) {                                           // all threads collect the counts, thread 0 builds the segments
            int part = 0;
#pragma unroll 4
            for (int b = tid; b < nblk; b += CW * 32) part += (int)wait_count(ws, b, seq);
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (lane == 0) s_cnt[wid] = part;
            __syncthreads();
            if (tid == 0) {
                int M = 0;
                for (int w = 0; w < CW; ++w) M += s_cnt[w];
                if (M != ws->sel.M || ws->sel.nseg <= 0) qdx_build_sel(M, &ws->sel);
            }
            __syncthreads();
        }
    } else {
        const int64_t per_cta = (p.K + nblk - 1) / nblk;          // contiguous block of cells of this CTA
        const int64_t c_lo = (int64_t)blockIdx.x * per_cta < p.K ? (int64_t)blockIdx.x * per_cta : p.K;
        const int64_t c_hi = c_lo + per_cta < p.K ? c_lo + per_cta : p.K;
        const bool keep_bits = (c_hi - c_lo + CW * 32 - 1) / (CW * 32) <= MAX_SLABS;
        if (tid == 0) {
            for (int s = 0; s < CW * NST; ++s) mbar_init(&s_bar[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        int32_t* job_cell = qdx_ws_jobs(p.ws, p.K);              // global list of changed cells: (cell, source row)
        int32_t* job_src = job_cell + p.K;

        // ---- phase 1, thread = cell: election result, fitness / descriptor, key reset, metrics, occupancy; the changed
        // cells of the whole grid are appended to ONE global list (warp-aggregated atomics) so that phase 2 can deal the
        // row copies out evenly -- winners per block vary, and so does the distance of an SM to the memory it reads
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
                    if (i < 0 || i >= p.B) { if (mode == 0) ws->error = QDX_ERR_BAD_INDEX; i = -1; }
                }
            }
            float fcell = -INFINITY;
            if (in) {
                if (i >= 0) {
                    fcell = __ldg(p.off_f + i);
                    p.rep_f[c] = fcell;
                    if (p.added_cells) p.added_cells[c] = (int32_t)i;
                    for (int d = 0; d < p.Dd; ++d) p.rep_d[c * p.Dd + d] = __ldg(p.off_d + i * p.Dd + d);
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
            const unsigned wb = __ballot_sync(0xffffffffu, i >= 0);
            unsigned base = 0;
            if (lane == 0 && wb) base = atomicAdd(&ws->job_count, (unsigned)__popc(wb));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (i >= 0) { const unsigned pos = base + __popc(wb & ((1u << lane) - 1u)); job_cell[pos] = (int32_t)c; job_src[pos] = (int32_t)i; }
        }
        if (lane == 0) s_cnt[wid] = occ_after;                // identical on every lane of the warp
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < CW; ++w) t += s_cnt[w];
            if (tail) *(volatile unsigned long long*)&ws->occ_pub[blockIdx.x] = ((unsigned long long)seq << 32) | (uint32_t)t;
            __threadfence();
            atomicAdd(&ws->cta_arrived, 1u);
            // ---- grid barrier (all CTAs are co-resident: cooperative launch): the job list is complete
            unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            while (*(volatile unsigned*)&ws->cta_arrived < (unsigned)nblk) {
                unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 2000000000ull) { ws->error = QDX_ERR_INTERNAL; break; }
            }
            __threadfence();
        }
        __syncthreads();

        // ---- phase 2: warp g of the grid streams jobs g, g + G, g + 2G, ... through its ring of shared-memory stages
        const unsigned njobs_all = *(volatile unsigned*)&ws->job_count;
        const uint32_t rowbytes = (uint32_t)p.D * 4u;
        const bool bulk = (p.D & 3) == 0;
        const int pieces = (int)((rowbytes + CHUNK - 1) / CHUNK);
        unsigned char* my_stage = s_stage + (size_t)wid * NST * CHUNK;
        const unsigned g = blockIdx.x * CW + wid, G = (unsigned)nblk * CW;
        uint32_t jq = 0;                                      // pieces this warp has pushed through its ring so far
        for (unsigned j0 = g; j0 < (QDX_COMMIT_EXP == 2 ? 0u : njobs_all); j0 += 32u * G) {      // up to 32 jobs per round: lane l holds job j0 + l G
            const unsigned jmine = j0 + (unsigned)lane * G;
            const int32_t my_cell = jmine < njobs_all ? __ldcg(job_cell + jmine) : 0;
            const int32_t my_src = jmine < njobs_all ? __ldcg(job_src + jmine) : 0;
            const int n_e = (int)((njobs_all - j0 + G - 1) / G) < 32 ? (int)((njobs_all - j0 + G - 1) / G) : 32;
            if (bulk) {
                const int np = n_e * pieces;
                for (int t = 0; t < np + LEAD; ++t) {
                    if (t >= LEAD) {                                  // piece t - LEAD has landed: send it on
                        const int j = t - LEAD;
                        const int e = j / pieces, pc = j % pieces;
                        const int32_t cell = __shfl_sync(0xffffffffu, my_cell, e);
                        if (lane == 0) {
                            const uint32_t q = jq + (uint32_t)j;
                            const uint32_t bytes = rowbytes - (uint32_t)pc * CHUNK < (uint32_t)CHUNK ? rowbytes - (uint32_t)pc * CHUNK : (uint32_t)CHUNK;
                            mbar_wait(&s_bar[wid * NST + (q % NST)], (q / NST) & 1u);
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            if (QDX_COMMIT_EXP != 1) bulk_s2g((char*)p.rep_g + ((int64_t)cell * p.D) * 4 + (int64_t)pc * CHUNK, my_stage + (q % NST) * CHUNK, bytes);
                        }
                    }
                    if (t < np) {                                     // stage free again? (its previous store has read it)
                        const int e = t / pieces, pc = t % pieces;
                        const int32_t src = __shfl_sync(0xffffffffu, my_src, e);
                        if (lane == 0) {
                            const uint32_t q = jq + (uint32_t)t;
                            const uint32_t bytes = rowbytes - (uint32_t)pc * CHUNK < (uint32_t)CHUNK ? rowbytes - (uint32_t)pc * CHUNK : (uint32_t)CHUNK;
                            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NST - LEAD) : "memory");
                            mbar_expect_tx(&s_bar[wid * NST + (q % NST)], bytes);
                            bulk_g2s(my_stage + (q % NST) * CHUNK, (const char*)p.off_g + ((int64_t)src * p.D) * 4 + (int64_t)pc * CHUNK, bytes,
                                     &s_bar[wid * NST + (q % NST)]);
                        }
                    }
                }
                jq += (uint32_t)np;
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
            } else {                                                      // D not a multiple of 4: plain per-lane copy
                for (int e = 0; e < n_e; ++e) {
                    const int32_t cell = __shfl_sync(0xffffffffu, my_cell, e), src = __shfl_sync(0xffffffffu, my_src, e);
                    const float* srow = p.off_g + (int64_t)src * p.D; float* drow = p.rep_g + (int64_t)cell * p.D;
                    for (int d = lane; d < p.D; d += 32) drow[d] = srow[d];
                }
            }
        }
        if (bulk && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");

        // ---- my slice of the ordered occupied-cell list: offset = sum of the predecessors' counts
        if (tail) {
            if (tid == 0) s_base = 0;
            __syncthreads();
            int part = 0;
#pragma unroll 4
            for (int b = tid; b < (int)blockIdx.x; b += CW * 32) part += (int)wait_count(ws, b, seq);
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
    // ---- metrics: CTAs publish partials, the last CTA to finish sums them in CTA order (deterministic)
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o); nan |= __shfl_xor_sync(0xffffffffu, nan, o); added += __shfl_xor_sync(0xffffffffu, added, o);
    }
    __syncthreads();
    if (lane == 0) { s_sum[wid] = sum; s_max[wid] = mx; s_cnt[wid] = cnt; s_nan[wid] = nan; s_add[wid] = added; }
    __syncthreads();
    if (tid == 0) {
        double s = 0.0; float m = -INFINITY; int n = 0, nn = 0, a = 0;
        for (int w = 0; w < CW; ++w) { s += s_sum[w]; m = fmaxf(m, s_max[w]); n += s_cnt[w]; nn |= s_nan[w]; a += s_add[w]; }
        ws->part_sum[blockIdx.x] = s; ws->part_max[blockIdx.x] = m; ws->part_cnt[blockIdx.x] = n; ws->part_nan[blockIdx.x] = nn;
        ws->part_add[blockIdx.x] = a;
        __threadfence();
        s_last = (atomicAdd(&ws->ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    {
        double s = 0.0; float m = -INFINITY; int n = 0, nn = 0, a = 0;
        for (unsigned b = tid; b < gridDim.x; b += CW * 32) {
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
    }
    if (tid == 0) {
        double s = 0.0; float m = -INFINITY; int n = 0, nn = 0, a = 0;
        for (int w = 0; w < CW; ++w) { s += s_sum[w]; m = fmaxf(m, s_max[w]); n += s_cnt[w]; nn |= s_nan[w]; a += s_add[w]; }
        float out[4];
        out[0] = (float)s + p.qd_offset * (float)n;               // qd_score   (metrics.py:92-93)
        out[1] = nn ? NAN : m;                                     // max_fitness (:95)
        out[2] = 100.0f * __fdiv_rn((float)n, (float)p.K);         // coverage   (:94)
        out[3] = (float)a;                                         // offspring inserted by this call
        if (tail) for (int j = 0; j < 4; ++j) { ws->metrics[j] = out[j]; if (p.metrics_out) p.metrics_out[j] = out[j]; }
        ws->ticket = 0u;
        ws->cta_arrived = 0u;
        ws->job_count = 0u;
        ws->commit_seq = seq;
        if (mode == 2 && ws->xchg_nranks > 0) {                    // peer-memory exchange: next generation, other key table
            uint32_t* ep = (uint32_t*)((char*)ws->xchg_peer[ws->xchg_rank] + QDX_XCHG_EPOCH_OFFSET);
            *ep = *ep + 1u;
        }
    }
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

extern "C" int qdx_commit(void* ws, int64_t K, int64_t D, int32_t desc_dim, const float* off_genotypes, const float* off_fitness,
                          const float* off_desc, uint32_t idx_base, int64_t B, int32_t first_wins, float* rep_genotypes,
                          float* rep_fitness, float* rep_desc, float qd_offset, float* metrics_out4, int32_t* added_cells,
                          int32_t mode, void* stream) {
    if (!ws || !off_genotypes || !off_fitness || !off_desc || !rep_genotypes || !rep_fitness || !rep_desc) return QDX_ERR_ARG;
    if (K <= 0 || D <= 0 || desc_dim < 1 || B < 0 || mode < 0 || mode > 2) return QDX_ERR_ARG;
    DeviceCaps caps;
    int rc = device_caps(&caps);
    if (rc) return rc;
    const bool aligned = (((uintptr_t)off_genotypes | (uintptr_t)rep_genotypes) & 15u) == 0;
    if (!caps.coop || caps.ctas_per_sm < 1 || ((D & 3) == 0 && !aligned))
        return qdx_launch_commit_generic(ws, K, D, desc_dim, off_genotypes, off_fitness, off_desc, idx_base, B, first_wins, rep_genotypes,
                                         rep_fitness, rep_desc, qd_offset, metrics_out4, added_cells, mode, (cudaStream_t)stream);
    int64_t nblk = (K + 31) / 32;                              // >= 32 cells per streaming CTA
    const int64_t cap = (int64_t)caps.sms * caps.ctas_per_sm - 1;
    if (nblk > cap) nblk = cap;
    if (nblk > QDX_MAX_COMMIT_CTAS - 1) nblk = QDX_MAX_COMMIT_CTAS - 1;
    if (nblk < 1) nblk = 1;
    CommitParams p;
    p.ws = ws; p.K = K; p.D = (int32_t)D; p.Dd = desc_dim; p.off_g = off_genotypes; p.off_f = off_fitness; p.off_d = off_desc;
    p.idx_base = idx_base; p.B = B; p.first_wins = first_wins; p.rep_g = rep_genotypes; p.rep_f = rep_fitness; p.rep_d = rep_desc;
    p.qd_offset = qd_offset; p.metrics_out = metrics_out4; p.added_cells = added_cells; p.mode = mode;
    void* args[] = {(void*)&p};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)qdx_commit_stream_kernel, dim3((unsigned)(nblk + 1)), dim3(CW * 32), args,
                                                (size_t)CW * NST * CHUNK, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : (int)e;
}
