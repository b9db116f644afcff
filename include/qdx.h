/*
 * qdx.h -- C ABI of the B200-native MAP-Elites generation step (libqdx.so).
 *
 * This is the drop-in boundary: plain C, device pointers + sizes + a CUDA stream handle (void* =
 * cudaStream_t), no torch / XLA types.  Every entry point is stream-ordered, non-blocking (except
 * qdx_workspace_read, which returns host values) and re-entrant; the only mutable state is the
 * per-repertoire device workspace passed explicitly.  Return value: 0 = ok, > 0 = cudaError_t,
 * < 0 = QDX_ERR_*.  All arrays are float32 / int32, row-major, C-contiguous, resident in device memory.
 *
 * The reference (QDax 0.5.1, /root/reference) has no FFI for this path: the seam is Python-level and
 * the arithmetic lives in jax/XLA.  Each function below names the reference interface it replaces;
 * INTEGRATION.md shows the jax.ffi / ctypes binding a maintainer would add on the reference side.
 */
#ifndef QDX_H_
#define QDX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QDX_ERR_ARG (-1)               /* bad argument */
#define QDX_ERR_UNSUPPORTED (-2)       /* shape outside the fused path; use the unfused entry points */
#define QDX_ERR_EMPTY_REPERTOIRE (-3)  /* selection from an all-empty repertoire (p = 0/0 in the reference) */
#define QDX_ERR_BAD_CELL (-4)          /* cell index out of range handed to qdx_offer_cells */
#define QDX_ERR_BAD_INDEX (-5)         /* winner index outside the offspring buffer handed to qdx_commit */
#define QDX_ERR_PEER_TIMEOUT (-6)      /* a peer's keys did not arrive within the caller's time limit (peer-memory exchange) */
#define QDX_ERR_INTERNAL (-7)          /* an in-kernel wait between CTAs timed out (never expected; reported instead of hanging) */

/* task ids of the fused scoring functions */
#define QDX_TASK_ID_NONE (-1)
#define QDX_TASK_ID_ARM 0        /* qdax/tasks/arm.py:9-50 */
#define QDX_TASK_ID_RASTRIGIN 1  /* qdax/tasks/standard_functions.py:9-15,27-37 */
#define QDX_TASK_ID_SPHERE 2     /* qdax/tasks/standard_functions.py:18-24,40-48 */

/* Separable grid tessellation (compute_euclidean_centroids, mapelites_repertoire.py:75-108):
 * the centroid of cell sum_d idx_d * stride[d] is (axes_0[idx_0], axes_1[idx_1], ...).  dd = 0: not a grid.
 * The fast path is used for descriptors with lo[d] <= x_d <= hi[d]; other rows fall back to brute force. */
typedef struct qdx_grid_desc {
    int32_t dd;
    int32_t n[4];
    int32_t stride[4];
    float lo[4];
    float hi[4];
    const float* axes; /* device pointer, concatenated axis values (ascending), sum n[d] floats */
} qdx_grid_desc;

int qdx_version(void);

/* ---- per-repertoire device workspace (selection tables, key chain, 64-bit insertion key table) ---- */
int qdx_workspace_bytes(int64_t K, int64_t* bytes);
/* byte offset of the K-entry uint64 insertion-key table inside the workspace (the only part that crosses GPUs) */
int qdx_workspace_keytab_offset(int64_t K, int64_t* offset);
int qdx_workspace_init(void* ws, int64_t K, void* stream);
int qdx_workspace_set_carry_key(void* ws, uint32_t k0, uint32_t k1, void* stream);
/* blocking read-back of the scan carry key, last metrics {qd_score, max_fitness, coverage, num_added}, error flag */
/* device-to-device copy of the scan carry key (two uint32 words) into (to_workspace != 0) or out of the workspace, stream
 * ordered: lets a caller whose keys live on the device (jax.random.key_data under jit) drive the key_mode 2 chain of
 * qdx_select_prepare without a host round trip (csrc/qdx_xla_ffi.cc) */
int qdx_workspace_copy_carry_key(void* ws, uint32_t* key2_device, int32_t to_workspace, void* stream);
int qdx_workspace_read(void* ws, uint32_t* carry_key2, float* metrics4, int32_t* error, void* stream);
/* Host mirror of the sticky device error flag: `host_pinned` points at one int32 in pinned (page-locked, UVA) host memory,
 * initialised to 0 by the caller.  Whichever kernel raises the flag (QDX_ERR_EMPTY_REPERTOIRE, _BAD_CELL, _BAD_INDEX,
 * _PEER_TIMEOUT, _INTERNAL) also stores the code there, so the host can poll it between calls without a blocking
 * read-back and without a copy on the stream.  NULL detaches. */
int qdx_workspace_set_error_mirror(void* ws, int32_t* host_pinned, void* stream);

/* ---- stage (a) set-up.  Replaces the p / cumsum part of UniformSelector.select
 * (repertoire_selectors/uniform_selector.py:43-45) and the jax.random.split chain of
 * MAPElites.update/ask (map_elites.py:177,241), scan_update (:214), DistributedMAPElites.update
 * (distributed_map_elites.py:124), MixingEmitter.emit (standard_emitters.py:55),
 * isoline_variation (mutation_operators.py:205,220).
 * key_mode: 0 keep keys | 1 (k0,k1) = key of MAPElites.update | 2 advance the workspace carry key
 * (scan_update) | 3 (k0,k1) = key of DistributedMAPElites.update | 4 (k0,k1) = key of MixingEmitter.emit */
/* rank_slot >= 0 additionally publishes this rank's generation keys in tail slot `rank_slot` of the key table so that
 * they travel with the all-reduce of the multi-GPU exchange (-1: single GPU). */
int qdx_select_prepare(const float* rep_fitness, int64_t K, void* ws, int32_t key_mode, uint32_t k0, uint32_t k1,
                       int32_t rank_slot, void* stream);

/* ---- stages (a)+(b)[+(c grid)+(d offer)] fused.  Replaces MixingEmitter.emit with variation_percentage=1
 * (standard_emitters.py:51-62: two UniformSelector.select + isoline_variation), the task scoring function,
 * and -- when `grid` describes the tessellation and offer != 0 -- get_cells_indices + segment_max
 * (mapelites_repertoire.py:202-231).  task = QDX_TASK_ID_NONE: variation only (out_genotypes required).
 * Offspring i of this call has global index idx_base + i (rank * B_dev + i under DistributedMAPElites).
 * out_genotypes may be NULL when the offspring rows are not needed (winners are re-read by qdx_commit, so
 * pass NULL only with offer = 0).
 * gen_keys8: the generation keys {sel1, sel2, line, leaf} derived on the host (qdx_host_generation_keys), or NULL
 * to use the keys qdx_select_prepare left in the workspace.
 * cvt: bucket index over non-grid centroids (qdx_cvt_index below, desc_dim <= 3) or NULL; with it the cell assignment
 * (and the offer) of a CVT tessellation is fused like the grid fast path.
 * flags: QDX_GEN_ROWS_FIRED_ONLY -- write only the genotype rows of offspring whose offer fired (the only ones
 * qdx_commit can elect: an offspring that does not beat its cell's occupant is never inserted), instead of streaming all
 * B rows to HBM; honoured when the offer is fused and D <= 128, else every row is written.  QDX_GEN_OUT_XCHG -- write the
 * rows into this rank's offspring block of the peer-memory exchange buffer (current epoch parity) instead of out_genotypes
 * (ignored), and fitness / descriptors into the block as well as into out_fitness / out_desc: qdx_commit(mode 3) on every
 * rank reads the winners from the blocks. */
#define QDX_GEN_ROWS_FIRED_ONLY 1
#define QDX_GEN_OUT_XCHG 2
struct qdx_cvt_index;
int qdx_generate(const float* rep_genotypes, const float* rep_fitness, const float* centroids, void* ws, int64_t K,
                 int64_t D, int64_t B, float iso_sigma, float line_sigma, int32_t has_min, float minval, int32_t has_max,
                 float maxval, int32_t task, int32_t desc_dim, const qdx_grid_desc* grid, int32_t offer,
                 uint32_t idx_base, int32_t first_wins, float* out_genotypes, float* out_fitness, float* out_desc,
                 int32_t* out_cells, int32_t* out_p1, int32_t* out_p2, const uint32_t* gen_keys8,
                 const struct qdx_cvt_index* cvt, int32_t flags, void* stream);

/* ---- stage (b) standalone: arm_scoring_function / rastrigin_scoring_function / sphere_scoring_function */
int qdx_score(int32_t task, const float* genotypes, int64_t B, int64_t D, int32_t desc_dim, float* out_fitness,
              float* out_desc, void* stream);

/* noisy_arm_scoring_function (qdax/tasks/arm.py:53-81): (k0, k1) = the key handed to the scoring function; Gaussian noise on
 * the genotype before the arm is evaluated, on the fitness and on the descriptor afterwards. */
int qdx_score_noisy_arm(const float* genotypes, int64_t B, int64_t D, uint32_t k0, uint32_t k1, float fit_variance, float desc_variance,
                        float params_variance, float* out_fitness, float* out_desc, void* stream);

/* ---- stage (c): get_cells_indices (mapelites_repertoire.py:111-137); grid == NULL or grid->dd == 0: brute
 * force with first-index argmin.  offer != 0 additionally performs the per-cell best-offspring offer. */
int qdx_cells(const float* desc, int64_t B, int32_t desc_dim, const float* centroids, int64_t K, const qdx_grid_desc* grid,
              int32_t* out_cells, void* ws, const float* rep_fitness, const float* fitness, int32_t offer,
              uint32_t idx_base, int32_t first_wins, void* stream);

/* ---- stage (c) through a uniform bucket index, for low-dimensional CVT tessellations (desc_dim <= 3): same result as
 * qdx_cells (bit-exact first-index argmin of the reference expression) at ~30 instead of K distance evaluations per
 * descriptor.  The index is built once per tessellation on the host: qdx_cvt_index_plan fills g / lo / h and the
 * bucket count (QDX_ERR_UNSUPPORTED: no index applies, use qdx_cells), qdx_cvt_index_build fills start (n_buckets + 1),
 * ids (K) and pts (K * dd) host arrays, which the caller uploads and points the descriptor at. */
typedef struct qdx_cvt_index {
    int32_t dd;           /* 0 = none */
    int32_t g[3];         /* buckets per dimension; bucket id = b0 + g0 * (b1 + g1 * b2) */
    float lo[3];          /* lower corner of the centroids' bounding box */
    float h[3];           /* bucket width */
    const int32_t* start; /* device */
    const int32_t* ids;   /* device */
    const float* pts;     /* device */
} qdx_cvt_index;
int qdx_cvt_index_plan(const float* centroids_host, int64_t K, int32_t desc_dim, qdx_cvt_index* plan, int64_t* n_buckets);
int qdx_cvt_index_build(const float* centroids_host, int64_t K, const qdx_cvt_index* plan, int32_t* start_host, int32_t* ids_host,
                        float* pts_host);
int qdx_cells_indexed(const float* desc, int64_t B, const qdx_cvt_index* index, int64_t K, int32_t* out_cells, void* ws,
                      const float* rep_fitness, const float* fitness, int32_t offer, uint32_t idx_base, int32_t first_wins,
                      void* stream);

/* ---- stage (c) on the tensor cores, for high-dimensional CVT descriptors (desc_dim <= 32): same result as
 * qdx_cells (bit-exact argmin of the reference expression): tcgen05 TF32 pass -> candidates within a proven error
 * band -> exact FP32 re-rank; unresolved rows take an exact brute-force pass.
 * qdx_cells_tc_workspace: sizes (in elements) of the per-tessellation `prep` float buffer and the per-call int32
 * `scratch`.  qdx_cells_tc_prepare: one-off per tessellation (swizzled centroid copy, ||c||^2, max ||c||^2). */
int qdx_cells_tc_workspace(int64_t K, int64_t B, int64_t* prep_floats, int64_t* scratch_ints);
int qdx_cells_tc_prepare(const float* centroids, int64_t K, int32_t desc_dim, float* prep, void* stream);
int qdx_cells_tc(const float* desc, int64_t B, int32_t desc_dim, const float* centroids, int64_t K, const float* prep,
                 int32_t* scratch, int32_t* out_cells, void* ws, const float* rep_fitness, const float* fitness, int32_t offer,
                 uint32_t idx_base, int32_t first_wins, void* stream);

/* ---- stage (d): MapElitesRepertoire.add (mapelites_repertoire.py:173-266).
 * qdx_offer_cells: segment_max + tie-break as a packed (32-bit fitness key, 31-bit index) atomicMax per cell
 * (offspring indices idx_base + i must stay below 2^31).
 * qdx_commit: scatter of the winners' genotype / fitness / descriptor rows into the repertoire (in place),
 * key-table reset, and default_qd_metrics (qdax/utils/metrics.py:74-98) -> metrics_out4 (device, optional).
 * mode 0: offspring rows indexed by (winner index - idx_base).  Modes 1 / 2 split the commit around an exchange
 * for DistributedMAPElites (distributed_map_elites.py:133-146) when only winners travel: mode 1 copies the
 * winners owned by [idx_base, idx_base + B) into per-cell staging rows (rep_* = staging; keys kept, no metrics);
 * mode 2 applies staging rows indexed by cell (off_* = staging), resets the keys and writes the metrics.
 * mode 3 (peer-memory exchange with offspring blocks, see qdx_xchg_* below; off_* ignored): waits for every rank's arrival
 * flag, then commits each cell's winner from its owner's offspring block.
 * Rows of at most 1 KB (D <= 256) and mode 3 take a lean kernel (ordinary launch, thread = cell, winners copied by their
 * warp); longer rows stream through the bulk-copy engine under a cooperative launch. */
int qdx_offer_cells(const int32_t* cells, const float* fitness, int64_t B, int64_t K, void* ws, const float* rep_fitness,
                    uint32_t idx_base, int32_t first_wins, void* stream);
int qdx_commit(void* ws, int64_t K, int64_t D, int32_t desc_dim, const float* off_genotypes, const float* off_fitness,
               const float* off_desc, uint32_t idx_base, int64_t B, int32_t first_wins, float* rep_genotypes,
               float* rep_fitness, float* rep_desc, float qd_offset, float* metrics_out4, int32_t* added_cells,
               int32_t mode, void* stream);

/* Multi-GPU "regen" exchange (DistributedMAPElites.update, distributed_map_elites.py:133-146, without moving genotypes):
 * after an all-reduce(max) over the key table (K keys + 8 * 64 key slots, int64, all values < 2^63) every rank
 * recomputes the elected winners from (owner's generation keys, local index) into per-cell staging rows; score the
 * staging rows with qdx_score and apply them with qdx_commit(mode 2).  Global index = rank * B_dev + i. */
int qdx_regenerate_winners(void* ws, int64_t K, int64_t D, int64_t B_dev, int32_t nranks, const float* rep_genotypes,
                           float iso_sigma, float line_sigma, int32_t has_min, float minval, int32_t has_max, float maxval,
                           int32_t first_wins, float* stage_genotypes, void* stream);

/* Same, fused with the scoring of the regenerated rows (stage_fitness (K,), stage_desc (K, desc_dim) are written for
 * the elected cells only), one warp per elected cell.  wait_peers > 0: first acquire-spin until every rank's keys of
 * this generation have arrived in the local exchange buffer (peer-memory exchange below), for at most wait_peers
 * MILLISECONDS: on timeout QDX_ERR_PEER_TIMEOUT is raised (device flag + host mirror), nothing is elected, and the
 * qdx_commit(mode 2) that follows applies nothing and keeps the epoch -- replicas never diverge silently. */
int qdx_elect_winners(void* ws, int64_t K, int64_t D, int32_t task, int32_t desc_dim, int64_t B_dev, int32_t nranks,
                      const float* rep_genotypes, float iso_sigma, float line_sigma, int32_t has_min, float minval,
                      int32_t has_max, float maxval, int32_t first_wins, float* stage_genotypes, float* stage_fitness,
                      float* stage_desc, int32_t wait_peers, void* stream);

/* ---- peer-memory exchange: the all-gather + replicated add of DistributedMAPElites.update
 * (distributed_map_elites.py:133-146) as direct NVLink traffic between the ranks' key tables, no collective library.
 * Each rank owns an exchange buffer (arrival flags + two key tables, double-buffered by generation parity) created
 * with cudaMalloc and exported with cudaIpc (64-byte handle, exchanged by the host: torch.distributed
 * all_gather_object); qdx_xchg_attach records every rank's mapping in the workspace, after which offers go to the
 * exchange table AND -- when they improve the cell's local best -- straight into every peer's table (system-scope
 * atomicMax over NVLink, issued by the offering thread inside qdx_generate / qdx_cells*), the last CTA of a fused
 * qdx_generate (offer != 0, gen_keys8 given) publishes this rank's generation keys and raises its arrival flag in every
 * peer (qdx_xchg_push does the same from a 1-thread kernel after un-fused cells kernels).
 * With B_dev > 0 the buffer also holds two OFFSPRING BLOCKS (genotype rows B_dev x D, fitness, descriptors; selected by
 * epoch parity like the key tables): qdx_generate(flags & QDX_GEN_OUT_XCHG) leaves the rank's offspring there, and
 * qdx_commit(mode 3) on every rank waits for all arrival flags (at most timeout_ms, then QDX_ERR_PEER_TIMEOUT and nothing is
 * applied) and copies each elected winner straight out of its OWNER's block over NVLink (global offspring index =
 * rank * B_dev + i) -- the all_gather of the reference reduced to the rows that actually change the repertoire, with no
 * staging and no recomputation; replicas are bit-identical because every rank copies the same bits.
 * With B_dev = 0 (keys only) the consumer is qdx_elect_winners(wait_peers > 0), which regenerates the winners instead.
 * nranks <= 16.  All ranks must call the same sequence of generations (the epoch counter lives on the device).
 * qdx_xchg_attach(nranks = 0) detaches. */
int qdx_xchg_bytes(int64_t K, int64_t B_dev, int64_t D, int32_t desc_dim, int64_t* bytes);
int qdx_xchg_create(int64_t K, int64_t B_dev, int64_t D, int32_t desc_dim, void** buf, void* ipc_handle64);
int qdx_xchg_open(const void* ipc_handle64, void** peer_buf);
int qdx_xchg_close(void* peer_buf);
int qdx_xchg_destroy(void* buf);
int qdx_xchg_attach(void* ws, int32_t rank, int32_t nranks, void* const* bufs, int64_t B_dev, int64_t D, int32_t desc_dim,
                    int32_t timeout_ms, void* stream);
int qdx_xchg_push(void* ws, int64_t K, const uint32_t* gen_keys8, void* stream);

/* ---- one whole generation per call: the body of MAPElites.update (qdax/core/map_elites.py:148-195) and, with
 * nranks > 1, of DistributedMAPElites.update (qdax/core/distributed_map_elites.py:92-161) for the fused configuration
 * (MixingEmitter with isoline variation only + arm / rastrigin / sphere scoring + default_qd_metrics).  Enqueues, in this
 * order and on `stream`: the jax.random.split chain on the host (qdx_host_generation_keys with key_mode / (k0, k1) /
 * carry_io2), qdx_generate (+ the offer when `grid` or `cvt` describes the tessellation), otherwise qdx_cells_tc (tc_prep
 * and tc_scratch given) or qdx_cells with the offer, then qdx_commit -- or, with nranks > 1 (exchange must be
 * QDX_EXCHANGE_P2P, the workspace attached with qdx_xchg_attach and offspring blocks of B rows): qdx_generate writes into
 * this rank's offspring block, [qdx_xchg_push ->] qdx_commit(mode 3) reads the winners from their owners' blocks.
 * Rank r's offspring have global indices [r * B, (r + 1) * B).
 * The selection tables of the workspace must describe rep_fitness (qdx_select_prepare once, every qdx_commit afterwards).
 * metrics_out4 (device, optional): {qd_score, max_fitness, coverage, offspring inserted}.  Non-blocking. */
#define QDX_EXCHANGE_NONE 0
#define QDX_EXCHANGE_P2P 1
typedef struct qdx_step_desc {
    float* rep_genotypes;     /* (K, D)  in place */
    float* rep_fitness;       /* (K,)    in place, -inf = empty cell */
    float* rep_desc;          /* (K, desc_dim) in place */
    const float* centroids;   /* (K, desc_dim) */
    void* ws;                 /* qdx_workspace_* */
    int64_t K, D, B;          /* cells, genotype dimension, offspring of THIS rank per generation */
    int32_t desc_dim, task;   /* QDX_TASK_ID_* */
    float iso_sigma, line_sigma;
    int32_t has_min; float minval;
    int32_t has_max; float maxval;
    const qdx_grid_desc* grid;          /* grid fast path, or NULL */
    const struct qdx_cvt_index* cvt;    /* bucket index (desc_dim <= 3), or NULL */
    const float* tc_prep;               /* qdx_cells_tc_prepare buffer (8 <= desc_dim <= 32), or NULL */
    int32_t* tc_scratch;                /* qdx_cells_tc_workspace scratch_ints */
    int32_t first_wins; float qd_offset;
    float* off_genotypes; float* off_fitness; float* off_desc; int32_t* off_cells;   /* offspring buffers (B rows) */
    int32_t rank, nranks, exchange;     /* 0, 1, QDX_EXCHANGE_NONE on one GPU */
} qdx_step_desc;
int qdx_map_elites_step(const qdx_step_desc* step, int32_t key_mode, uint32_t k0, uint32_t k1, uint32_t* carry_io2,
                        float* metrics_out4, void* stream);

/* ---- pieces of the preserved Python surface ---- */
/* UniformSelector.select index stream for the key handed to select() (uniform_selector.py:48-55) */
int qdx_select_indices(void* ws, uint32_t k0, uint32_t k1, int64_t num, int32_t* out, void* stream);
/* UniformSelector(select_with_replacement=False).select index stream (uniform_selector.py:19-20, :49-55: jax.random.choice with
 * replace=False = Gumbel top-k); the workspace's selection tables must describe rep_fitness (qdx_select_prepare / qdx_commit);
 * scratch_K: K floats; num <= K <= 2^18. */
int qdx_select_indices_without_replacement(const float* rep_fitness, int64_t K, void* ws, uint32_t k0, uint32_t k1, int64_t num, float* scratch_K,
                                           int32_t* out, void* stream);
int qdx_gather_rows(const float* src, const int32_t* idx, int64_t B, int64_t D, float* out, void* stream);
/* isoline_variation(x1, x2, key, ...) on dense parents (mutation_operators.py:175-226) */
int qdx_isoline_variation(const float* x1, const float* x2, int64_t B, int64_t D, uint32_t k0, uint32_t k1, float iso_sigma,
                          float line_sigma, int32_t has_min, float minval, int32_t has_max, float maxval, float* out,
                          void* stream);
/* ---- sibling insertion rules on the same cells + election + commit machinery.
 * MELSRepertoire.add (qdax/core/containers/mels_repertoire.py:89-230): every individual comes with S evaluations.
 * cells_all (B*S) = qdx_cells of all descriptors; per individual the kernel takes the most frequent cell (smallest on
 * ties, _mode :51-57), the spread = mean pairwise descriptor distance (_dispersion :26-48; 0 when S == 1), the mean
 * fitness (:169) and the centroid of the cell as stored descriptor (:162-164), writes them to out_*, and offers the
 * individual to its cell iff fitness > rep_fitness[cell] && spread <= rep_spread[cell] (:181-187).  Collisions -- left
 * to scatter order by the reference -- go to the first / last offspring index.  Follow with qdx_commit(out_genotypes =
 * the batch, off_fitness = out_fitness, off_desc = out_desc, added_cells) and qdx_scatter_rows_by_source for the spreads. */
int qdx_mels_offer(const int32_t* cells_all, const float* desc_all, const float* fit_all, int64_t B, int32_t S, int32_t desc_dim,
                   const float* centroids, int64_t K, void* ws, const float* rep_fitness, const float* rep_spread,
                   int32_t first_wins, int32_t* out_cells, float* out_fitness, float* out_spread, float* out_desc, void* stream);
/* dst[c, :] = src[source_of_cell[c], :] where source_of_cell[c] >= 0 (qdx_commit's added_cells): per-cell side arrays
 * (spreads, extra_scores; mapelites_repertoire.py:250-257) */
int qdx_scatter_rows_by_source(const int32_t* source_of_cell, const float* src, int64_t K, int64_t W, float* dst, void* stream);

/* MOMERepertoire.add (qdax/core/containers/mome_repertoire.py:211-322): the batch is scanned in index order, every offspring
 * updating the Pareto front of its cell (_update_masked_pareto_front :72-209, masked dominance of qdax/utils/pareto_front.py:
 * 48-95), in place.  rep_fitness (K, front_len, num_criteria) with -inf rows for empty slots, rep_genotypes (K, front_len, D),
 * rep_desc (K, front_len, desc_dim); cells = qdx_cells of the batch descriptors.  One warp per cell (offspring of one cell are
 * applied sequentially, cells in parallel).  front_len <= 256, num_criteria <= 8.  ieee_literal: how the source's float * bool
 * products are read -- 0: XLA's compiled Select(mask, x, 0) (default), 1: IEEE product of the converted mask (inf * 0 = NaN). */
int qdx_mome_add(float* rep_fitness, float* rep_genotypes, float* rep_desc, int64_t K, int32_t front_len, int32_t num_criteria, int64_t D,
                 int32_t desc_dim, const int32_t* cells, const float* fitness, const float* genotypes, const float* desc, int64_t B,
                 int32_t ieee_literal, void* stream);
/* UnstructuredRepertoire.add (qdax/core/containers/unstructured_repertoire.py:162-337) in three calls around two gathers:
 * qdx_unstructured_plan: nearest / second-nearest occupied slot as the source computes them (see oracle/
 * qdax_containers_numpy.py), the l-value tests, the re-ordering of the batch (offspring that open a new slot first) ->
 * out_order (B int32: position -> batch index; `scratch` keeps the target slots and flags).  The caller gathers genotypes /
 * descriptors / fitnesses by out_order (qdx_gather_rows), then qdx_unstructured_offer runs intra_batch_comp (:69-129) and the
 * segment_max election into the workspace key table (N = max_size cells), and qdx_commit(off_* = the gathered batch) applies it.
 * qdx_unstructured_scratch: bytes of `scratch`. */
int qdx_unstructured_scratch(int64_t N, int64_t B, int64_t* bytes);
int qdx_unstructured_plan(const float* rep_fitness, const float* rep_desc, int64_t N, int32_t desc_dim, const float* fitness, const float* desc,
                          int64_t B, float l_value, void* scratch, int32_t* out_order, void* stream);
int qdx_unstructured_offer(const float* sorted_fitness, const float* sorted_desc, int64_t B, int32_t desc_dim, int64_t N, float l_value,
                           void* scratch, const int32_t* order, void* ws, const float* rep_fitness, int32_t first_wins, void* stream);

/* ---- compute_cvt_centroids on the GPU (mapelites_repertoire.py:30-72 calls scikit-learn KMeans on
 * the host).  One Lloyd iteration = qdx_cells (assignment) + qdx_kmeans_accumulate + qdx_kmeans_update.  The update is
 * order-free: coordinates (samples in [0, 1), as the reference clusters in the unit cube) are quantised to 32 fractional
 * bits and summed with 64-bit integer atomics, centroid = (double)sum / count * 2^-32 -> float32, an empty cluster keeps its
 * centroid.  acc (K * desc_dim), count (K) and changed (1) are zeroed by qdx_kmeans_accumulate; `changed` counts samples
 * whose cell differs from prev_cells (NULL on the first iteration). */
int qdx_kmeans_accumulate(const float* x, const int32_t* cells, const int32_t* prev_cells, int64_t N, int32_t desc_dim, int64_t K,
                          unsigned long long* acc, int32_t* count, int32_t* changed, void* stream);
int qdx_kmeans_update(const unsigned long long* acc, const int32_t* count, const float* old_centroids, int64_t K, int32_t desc_dim,
                      float* new_centroids, void* stream);

/* ---- pytree genotypes.  An individual is stored as ONE packed row: the concatenation of its
 * flattened leaves in jax.tree.leaves order, leaf l owning genes [off[l], off[l+1]).  isoline_variation on a pytree
 * (mutation_operators.py:205-224) shares the line noise across leaves and draws leaf l's iso noise as
 * normal(split(key', n_leaves)[l], (B,) + leaf_shape), i.e. with counter i * size_l + j inside the leaf. */
#define QDX_MAX_LEAVES 32
typedef struct qdx_leaf_table {
    int32_t n;                          /* 1 .. QDX_MAX_LEAVES */
    int32_t off[QDX_MAX_LEAVES + 1];    /* off[0] = 0 <= off[1] <= ... <= off[n] = D */
    uint32_t key[2 * QDX_MAX_LEAVES];   /* key words of split(key', n)[l] */
} qdx_leaf_table;
/* MixingEmitter.emit (variation only) for a pytree genotype: two selections + isoline over packed rows; gen_keys8 as for
 * qdx_generate (its leaf key is ignored, the table's keys are used). */
int qdx_generate_leaves(const float* rep_genotypes, const float* rep_fitness, void* ws, int64_t K, int64_t D, int64_t B,
                        float iso_sigma, float line_sigma, int32_t has_min, float minval, int32_t has_max, float maxval,
                        float* out_genotypes, int32_t* out_p1, int32_t* out_p2, const uint32_t* gen_keys8,
                        const qdx_leaf_table* leaves, void* stream);
/* isoline_variation(x1, x2, key) on dense packed parents; (line_k0, line_k1) = split(key)[1] */
int qdx_isoline_variation_leaves(const float* x1, const float* x2, int64_t B, int64_t D, uint32_t line_k0, uint32_t line_k1,
                                 const qdx_leaf_table* leaves, float iso_sigma, float line_sigma, int32_t has_min, float minval,
                                 int32_t has_max, float maxval, float* out, void* stream);
/* strided 2-D float copy (rows x cols, leading dimensions in floats): packs / unpacks leaves into / out of packed rows */
int qdx_copy_2d(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int64_t rows, int64_t cols, void* stream);
/* polynomial_mutation (mutation_operators.py:81-117; n_mutate = int(proportion_to_mutate * D), eta_plus_1 = 1 + eta,
 * mutpow = 1 / (1 + eta)) and polynomial_crossover (:139-172; n_change = int(proportion_var_to_change * D)).
 * out must not alias the inputs. */
int qdx_polynomial_mutation(const float* x, int64_t B, int64_t D, uint32_t k0, uint32_t k1, int32_t n_mutate, float eta_plus_1,
                            float mutpow, float minval, float maxval, float* out, void* stream);
int qdx_polynomial_crossover(const float* x1, const float* x2, int64_t B, int64_t D, uint32_t k0, uint32_t k1, int32_t n_change,
                             float* out, void* stream);
/* jax.random streams: kind 0 bits (uint32), 1 uniform(minval, maxval), 2 normal */
int qdx_random(uint32_t k0, uint32_t k1, int64_t n, int32_t kind, float minval, float maxval, void* out, void* stream);
/* default_qd_metrics -> {qd_score, max_fitness, coverage} */
int qdx_metrics(const float* rep_fitness, int64_t K, float qd_offset, float* out3, void* stream);

/* ---- stage (e): DominatedNoveltyRepertoire.add (qdax/core/containers/dns_repertoire.py:94-165) including
 * _novelty_and_dominated_novelty (:22-76).  Candidates = population rows followed by batch rows; outputs are
 * the new population (separate buffers, sorted exactly like argsort(meta_fitness)[::-1][:P]).
 * meta_scratch: (P+B) floats (receives the meta fitness), survivors_scratch: P int32 (receives survivor indices). */
int qdx_dns_add(const float* pop_genotypes, const float* pop_fitness, const float* pop_desc, int64_t P,
                const float* batch_genotypes, const float* batch_fitness, const float* batch_desc, int64_t B, int64_t D,
                int32_t desc_dim, int32_t k, float* out_genotypes, float* out_fitness, float* out_desc, float* meta_scratch,
                int32_t* survivors_scratch, void* stream);

/* ---- host-only helpers (no GPU needed; used by the CPU test-suite) ---- */
/* The jax.random.split chain from the key handed to MAPElites.update (key_mode 1; map_elites.py:177,241), one
 * scan_update step on carry_io2 (2; :214, carry advanced in place), DistributedMAPElites.update (3;
 * distributed_map_elites.py:124) or MixingEmitter.emit (4) down to the four stream keys of a generation
 * {sel1, sel2, line, leaf} (standard_emitters.py:55, uniform_selector.py:48, mutation_operators.py:205,220). */
/* jax.random.split(key, n) -> n keys of two words (partitionable Threefry: key i = threefry(key, counter (0, i))) */
int qdx_host_split(uint32_t k0, uint32_t k1, int32_t n, uint32_t* out);
int qdx_host_generation_keys(int32_t key_mode, uint32_t k0, uint32_t k1, uint32_t* carry_io2, uint32_t* out_keys8);
int qdx_host_select_table(int32_t M, float* out_T, int32_t* out_nseg);
int qdx_host_select_rank(int32_t M, const float* r, int64_t n, int32_t* out_rank);

#ifdef __cplusplus
}
#endif
#endif /* QDX_H_ */
