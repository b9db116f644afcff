// One whole generation behind ONE C entry point: the body of MAPElites.update (qdax/core/map_elites.py:148-195 under
// /root/reference) / DistributedMAPElites.update (qdax/core/distributed_map_elites.py:92-161) for the fused
// configuration -- host key chain -> generate (+ score + cells + offer) [-> cells] [-> elect] -> commit -- enqueued by a
// single call, so the per-generation host cost is one FFI crossing instead of one per kernel (8 GPUs: the Python / ctypes
// enqueue of three kernels was 80 us against a 143 us generation).  Host code only: every launch goes through the stage
// entry points of include/qdx.h, in the order the Python layer used to issue them.
#include <cstring>
#include "../../include/qdx.h"

extern "C" int qdx_map_elites_step(const qdx_step_desc* s, int32_t key_mode, uint32_t k0, uint32_t k1, uint32_t* carry_io2,
                                   float* metrics_out4, void* stream) {
    if (!s || !s->rep_genotypes || !s->rep_fitness || !s->rep_desc || !s->centroids || !s->ws) return QDX_ERR_ARG;
    if (s->nranks < 1 || s->rank < 0 || s->rank >= s->nranks) return QDX_ERR_ARG;
    const bool multi = s->nranks > 1;
    if (multi && s->exchange != QDX_EXCHANGE_P2P) return QDX_ERR_ARG;
    if (!s->off_cells || !s->off_fitness || !s->off_desc || (!multi && !s->off_genotypes)) return QDX_ERR_ARG;
    uint32_t keys[8];
    int rc = qdx_host_generation_keys(key_mode, k0, k1, carry_io2, keys);
    if (rc) return rc;
    const bool grid = s->grid && s->grid->dd != 0;
    const bool index = !grid && s->cvt && s->cvt->dd != 0;
    const bool fused_cells = grid || index;
    const uint32_t base = (uint32_t)((int64_t)s->rank * s->B);
    // only the rows that can be elected are written (fused offer); multi-GPU: into this rank's offspring block, where the peers read them
    const int32_t flags = QDX_GEN_ROWS_FIRED_ONLY | (multi ? QDX_GEN_OUT_XCHG : 0);
    rc = qdx_generate(s->rep_genotypes, s->rep_fitness, s->centroids, s->ws, s->K, s->D, s->B, s->iso_sigma, s->line_sigma, s->has_min,
                      s->minval, s->has_max, s->maxval, s->task, s->desc_dim, grid ? s->grid : nullptr, fused_cells ? 1 : 0, base,
                      s->first_wins, s->off_genotypes, s->off_fitness, s->off_desc, s->off_cells, nullptr, nullptr, keys,
                      index ? s->cvt : nullptr, flags, stream);
    if (rc) return rc;
    if (!fused_cells) {     // cell assignment stays sharded with the offspring: tensor-core pass or FP32 brute force, + offer
        if (s->tc_prep && s->tc_scratch)
            rc = qdx_cells_tc(s->off_desc, s->B, s->desc_dim, s->centroids, s->K, s->tc_prep, s->tc_scratch, s->off_cells, s->ws,
                              s->rep_fitness, s->off_fitness, 1, base, s->first_wins, stream);
        else
            rc = qdx_cells(s->off_desc, s->B, s->desc_dim, s->centroids, s->K, nullptr, s->off_cells, s->ws, s->rep_fitness,
                           s->off_fitness, 1, base, s->first_wins, stream);
        if (rc) return rc;
        if (multi) { rc = qdx_xchg_push(s->ws, s->K, keys, stream); if (rc) return rc; }
    }
    return qdx_commit(s->ws, s->K, s->D, s->desc_dim, s->off_genotypes, s->off_fitness, s->off_desc, base, s->B, s->first_wins,
                      s->rep_genotypes, s->rep_fitness, s->rep_desc, s->qd_offset, metrics_out4, nullptr, multi ? 3 : 0, stream);
}
