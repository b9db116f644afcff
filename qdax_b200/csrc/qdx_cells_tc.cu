// Nearest-centroid cell assignment for high-dimensional CVT descriptors on the 5th-gen tensor cores (sm_100a).
//
// Reference semantics (get_cells_indices, qdax/core/containers/mapelites_repertoire.py:111-137 under
// /root/reference): argmin_k sum_d (x_d - c_kd)^2 in float32, sequential over d, FIRST minimum.  The reference
// never expands the square; neither does the final decision here:
//
//   1. tensor pass   s~(i,k) = x_i . c_k - ||c_k||^2 / 2   (argmin of the distance = argmax of s) entirely on the tensor cores:
//                    tcgen05.mma kind::tf32 (M=128, N=128), four K=8 steps over the 32 descriptor dimensions plus a FIFTH
//                    K=8 step that multiplies [1, 1, 1, 0...] by [-||c||^2/2 split into three TF32-exact terms, 0...], so the
//                    epilogue is a bare running maximum over the accumulator (no ||c||^2 fetch, no FMA per pair);
//                    FP32 accumulators in TMEM, double buffered; centroid tiles stream through a 3-stage shared-memory
//                    ring filled by ONE 1-D bulk async copy per tile (cp.async.bulk + mbarrier complete_tx) from a copy
//                    of the centroids that was written ONCE in the layouts the UMMA descriptors expect: a 16 KB
//                    128-byte-swizzled K-major block (a 32-float row is exactly one swizzle row, so no tensor map is
//                    needed) followed by the 4 KB 32-byte-swizzled block of the extra K step; warp roles: 1 copy-issuer,
//                    1 MMA-issuer, 4 epilogue warps (one TMEM lane quadrant each, thread = descriptor row); two CTAs per SM.
//                    Both x and c are centred on the centroid mean mu first (distances are translation invariant):
//                    the TF32 error scales with ||x - mu|| * ||c - mu||, 4x smaller than the uncentred product for
//                    data in [0,1]^32, and it makes the bound usable for tessellations far from the origin.
//   2. candidates    every 5-column group holding a k with s~ >= max_k s~ - band_i is recorded (row's slots in shared memory), where band_i bounds twice the
//                    TF32 error (|x~c~ - xc| <= 2^-9 |xc| per product, Cauchy-Schwarz over the row) plus twice the
//                    rounding error of the reference's own float32 sum -- so the list provably contains every index
//                    that can attain the reference's computed minimum.
//   3. exact re-rank the candidates are re-evaluated with the reference expression (sub, mul, sequential add) and the
//                    minimum is taken with the lowest index on ties.  A row whose list is full (possible overflow)
//                    or whose descriptor is not finite is resolved by an exact brute-force pass.
//
// Bit-exact against the oracle (tests/test_gpu_parity.py::test_cells_tensor_core_path).
#include "qdx_common.cuh"
#include "../../include/qdx.h"

#define QDX_CHECK_LAUNCH() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return (int)e_; } while (0)

namespace tc {

constexpr int KD = 32;                 // padded descriptor dimension = one 128-byte swizzle row of float32
#ifndef QDX_TC_HALVES
#define QDX_TC_HALVES 1                // 1 = 128 rows per CTA, two CTAs per SM; 2 = 256 rows per CTA sharing every centroid tile, one CTA per
                                       // SM (half the L2 -> SM stream: measured 10 % SLOWER, the stream is 11 % of the L2 peak and two CTAs hide
                                       // each other's accumulator hand-over latency; profiles/r2_notes.md)
#endif
constexpr int TILE_M = 128;            // descriptor rows per MMA (UMMA M)
constexpr int HALVES = QDX_TC_HALVES;  // row tiles per CTA
constexpr int CTA_ROWS = HALVES * TILE_M;
constexpr int TILE_N = 128;            // centroids per MMA (UMMA N); 2 x 128 TMEM columns per CTA -> two CTAs per SM
constexpr int STAGES = HALVES == 1 ? 3 : 6;   // shared-memory ring of centroid tiles (the copy warp runs ahead of the MMAs with 3 already)
constexpr int ACC_STAGES = 2;          // TMEM accumulator double buffer (per stage: HALVES x 128 columns)
constexpr int TMEM_COLS = ACC_STAGES * HALVES * TILE_N;
constexpr int CAP = 16;                // recorded chunks per row (shared memory, append-only; compacted against the current threshold when full)
constexpr int KX = 8;                  // the extra K step (one UMMA_K of tf32 = 32 bytes per row, SWIZZLE_32B)
constexpr int A_BYTES = TILE_M * KD * 4;        // 16 KB
constexpr int AX_BYTES = TILE_M * KX * 4;       // 4 KB
constexpr int B_BYTES = TILE_N * KD * 4;        // 16 KB
constexpr int BX_BYTES = TILE_N * KX * 4;       // 4 KB
constexpr int STAGE_BYTES = B_BYTES + BX_BYTES; // one centroid tile: both blocks, contiguous in global memory too
constexpr int EPI_WARPS = 4 * HALVES;  // one per (row tile, TMEM lane quadrant)
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;   // warp 0: copies, warp 1: MMA + TMEM alloc, then the epilogue warps
constexpr int CBUF_BYTES = CAP * CTA_ROWS * 8;
constexpr int LIST_CAP = 48;           // centroid indices re-evaluated exactly per row at the end (list in the drained centroid ring)
static_assert(LIST_CAP * CTA_ROWS * 4 <= STAGES * STAGE_BYTES, "the index lists reuse the centroid ring");
constexpr int SMEM_BYTES = 1024 /*align slack*/ + HALVES * A_BYTES + AX_BYTES + STAGES * STAGE_BYTES + 256 /*barriers*/ + CBUF_BYTES;
constexpr float PAD_S = -1.0e30f;      // s of a padding centroid: never the maximum

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// (a suspend-time hint on try_wait was measured in round 2: 1.0646 vs 1.0661 ms, no effect -- profiles/r2_notes.md)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// ---- bulk async copy global -> shared, completion on an mbarrier ----------------------------------------------
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// ---- tcgen05 ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::tf32
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread (thread t of the warp <-> TMEM lane base + t).
// The load is asynchronous: v[] must not be read before tmem_wait_ld(v), which carries v as in/out operands so that
// neither the compiler nor ptxas can hoist a consumer above the wait.
#define QDX_V32(C) C(0) C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13) C(14) C(15) \
                   C(16) C(17) C(18) C(19) C(20) C(21) C(22) C(23) C(24) C(25) C(26) C(27) C(28) C(29) C(30) C(31)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
          "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]),
          "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]),
          "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld(float (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
        : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]),
          "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]), "+f"(v[16]),
          "+f"(v[17]), "+f"(v[18]), "+f"(v[19]), "+f"(v[20]), "+f"(v[21]), "+f"(v[22]), "+f"(v[23]), "+f"(v[24]),
          "+f"(v[25]), "+f"(v[26]), "+f"(v[27]), "+f"(v[28]), "+f"(v[29]), "+f"(v[30]), "+f"(v[31])
        :: "memory");
}

// Shared-memory matrix descriptors, K-major.  SWIZZLE_128B: rows of 128 bytes, 8-row atoms 1024 bytes apart (layout type 2);
// SWIZZLE_32B: rows of 32 bytes, 8-row atoms 256 bytes apart (layout type 6).  The leading byte offset is unused for
// swizzled K-major operands whose K extent per instruction fits in one swizzle row (1).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t atom_stride_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);           // start address, 16-byte units          bits [0,14)
    d |= (uint64_t)1 << 16;                                // leading byte offset (unused: 1)       bits [16,30)
    d |= (uint64_t)(atom_stride_bytes >> 4) << 32;         // stride byte offset between 8-row atoms bits [32,46)
    d |= (uint64_t)1 << 46;                                // descriptor version (Blackwell)        bits [46,48)
    d |= (uint64_t)layout_type << 61;                      // layout type                           bits [61,64)
    return d;
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) { return make_desc(smem_addr, 1024, 2); }
__device__ __forceinline__ uint64_t make_desc_sw32(uint32_t smem_addr) { return make_desc(smem_addr, 256, 6); }
// Instruction descriptor: D = F32, A = B = TF32, both K-major, N = 256, M = 128.
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TILE_N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);

// byte offset of 16-byte chunk `c` of row `r` inside a tile whose base is 1024-byte aligned (Swizzle<3,4,3>)
__device__ __host__ __forceinline__ uint32_t sw128_offset(uint32_t r, uint32_t c) { return r * 128u + ((c ^ (r & 7u)) << 4); }
// same for a 32-byte-swizzled tile (Swizzle<1,4,3>: address bit 7 = bit 2 of the row, XORed into the chunk bit)
__device__ __host__ __forceinline__ uint32_t sw32_offset(uint32_t r, uint32_t c) { return r * 32u + ((c ^ ((r >> 2) & 1u)) << 4); }

}  // namespace tc

// ---------------------------------------------------------------------------------------------------------------
// one-off per tessellation: swizzled, zero-padded copy of the centroids + ||c||^2 (+inf for padding) + max ||c||
// ---------------------------------------------------------------------------------------------------------------
// column means of the centroids (deterministic: one CTA per dimension, fixed-shape tree)
__global__ void __launch_bounds__(256) qdx_cells_tc_mean_kernel(const float* __restrict__ cent, int64_t K, int32_t Dd, float* __restrict__ mu) {
    __shared__ double s_part[256];
    const int d = blockIdx.x;
    double acc = 0.0;
    if (d < Dd) for (int64_t k = threadIdx.x; k < K; k += blockDim.x) acc += (double)cent[k * Dd + d];
    s_part[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) s_part[threadIdx.x] += s_part[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) mu[d] = d < Dd ? (float)(s_part[0] / (double)K) : 0.0f;
}

__global__ void __launch_bounds__(256) qdx_cells_tc_prepare_kernel(const float* __restrict__ cent, int64_t K, int32_t Dd,
                                                                   int64_t Kpad, const float* __restrict__ mu,
                                                                   float* __restrict__ cs, float* __restrict__ cmax2) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Kpad) return;
    float row[tc::KD];
    float n2 = 0.0f;
#pragma unroll
    for (int d = 0; d < tc::KD; ++d) {
        row[d] = (k < K && d < Dd) ? cent[k * Dd + d] - mu[d] : 0.0f;
        n2 = __fmaf_rn(row[d], row[d], n2);
    }
    const int64_t tile = k / tc::TILE_N;
    const uint32_t r = (uint32_t)(k % tc::TILE_N);
    char* base = (char*)cs + tile * tc::STAGE_BYTES;
#pragma unroll
    for (int c = 0; c < 8; ++c)
        *reinterpret_cast<float4*>(base + tc::sw128_offset(r, c)) = make_float4(row[4 * c], row[4 * c + 1], row[4 * c + 2], row[4 * c + 3]);
    // -||c-mu||^2 / 2 as the exact sum of three floats with 11 significant bits each (what a TF32 operand keeps): the tensor
    // core adds them to x . c against the ones of the A operand's extra columns
    const float h = (k < K) ? -0.5f * n2 : tc::PAD_S;
    const float h0 = __uint_as_float(__float_as_uint(h) & 0xFFFFE000u);
    const float r1 = h - h0;
    const float h1 = __uint_as_float(__float_as_uint(r1) & 0xFFFFE000u);
    const float h2 = r1 - h1;
    char* bx = base + tc::B_BYTES;
    *reinterpret_cast<float4*>(bx + tc::sw32_offset(r, 0)) = make_float4(h0, h1, h2, 0.0f);
    *reinterpret_cast<float4*>(bx + tc::sw32_offset(r, 1)) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (k < K) atomicMax(reinterpret_cast<int*>(cmax2), __float_as_int(n2));     // n2 >= 0: int order == float order
}

// ---------------------------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------------------------
struct QdxTcParams {
    const float* desc; int64_t B; int32_t Dd;
    const float* cent; int64_t K; int64_t Kpad;
    const float* cs; const float* cmax2; const float* mu;   // cmax2[0] = max ||c-mu||^2, mu = cmax2 + 32
    int32_t* cells; int32_t* fallback_rows; int32_t* fallback_count;
    void* ws; const float* rep_f; const float* fit; int32_t offer; uint32_t idx_base; int32_t first_wins;
};

// the reference expression, sequential over d; fully unrolled so that x[] stays in registers
template <int DDPAD>
__device__ __forceinline__ float qdx_exact_dist(const float (&x)[DDPAD], const float* __restrict__ c, int Dd) {
    float acc = 0.0f;
#pragma unroll
    for (int d = 0; d < DDPAD; ++d) {
        if (d < Dd) { float df = x[d] - __ldg(c + d); float s = df * df; acc = d ? acc + s : s; }
    }
    return acc;
}

// a row's slots are full: keep the recorded chunks whose maximum is still above the threshold (one copy in the instruction stream)
__device__ __noinline__ int qdx_tc_compact(uint2* my, float thr) {
    int w = 0;
    for (int e = 0; e < tc::CAP; ++e) {
        const uint2 x = my[e * tc::CTA_ROWS];
        if (__uint_as_float(x.x) > thr) { my[w * tc::CTA_ROWS] = x; ++w; }
    }
    return w;
}

__global__ void __launch_bounds__(tc::NUM_THREADS, tc::HALVES == 1 ? 2 : 1) qdx_cells_tc_kernel(const QdxTcParams p) {
    using namespace tc;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);     // SWIZZLE_128B needs 1024-byte alignment
    uint8_t* sA = smem;
    uint8_t* sAX = smem + HALVES * A_BYTES;        // extra K step of A: [1, 1, 1, 0, 0, 0, 0, 0] per row (the same for every row tile)
    uint8_t* sB = sAX + AX_BYTES;                  // ring of centroid tiles: [16 KB SW128 block | 4 KB SW32 block] per stage
    uint64_t* bars = (uint64_t*)(sB + STAGES * STAGE_BYTES);
    uint64_t* full = bars;                       // [STAGES]  copies landed
    uint64_t* empty = bars + STAGES;             // [STAGES]  MMA finished reading the stage
    uint64_t* acc_full = bars + 2 * STAGES;      // [ACC_STAGES] accumulator ready
    uint64_t* acc_empty = bars + 2 * STAGES + ACC_STAGES;   // [ACC_STAGES] accumulator drained by the epilogue
    uint32_t* tmem_base_slot = (uint32_t*)(bars + 2 * STAGES + 2 * ACC_STAGES);
    uint2* cbuf = (uint2*)((uint8_t*)bars + 256);          // [CAP][CTA_ROWS] (s~ bits, centroid index): entry-major, lanes side by side

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row0 = (int64_t)blockIdx.x * CTA_ROWS;
    const int ntiles = (int)(p.Kpad / TILE_N);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < ACC_STAGES; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_base_slot, TMEM_COLS);
    // A tile: 128 descriptor rows, zero-padded to 32 floats, written in the swizzled layout by all threads
    for (int i = threadIdx.x; i < CTA_ROWS * 8; i += NUM_THREADS) {     // row tile h at sA + h * A_BYTES (r * 128 bytes per row: contiguous)
        const int r = i >> 3, c = i & 7;
        const int64_t row = row0 + r;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { const int d = 4 * c + j; v[j] = (row < p.B && d < p.Dd) ? p.desc[row * p.Dd + d] - p.mu[d] : 0.0f; }
        *reinterpret_cast<float4*>(sA + sw128_offset(r, c)) = make_float4(v[0], v[1], v[2], v[3]);
    }
    for (int i = threadIdx.x; i < TILE_M * 2; i += NUM_THREADS) {
        const int r = i >> 1, c = i & 1;
        *reinterpret_cast<float4*>(sAX + sw32_offset(r, c)) = c == 0 ? make_float4(1.0f, 1.0f, 1.0f, 0.0f) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the MMA (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        // ===== copy issuer =====
        if (lane == 0) {
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % STAGES; const uint32_t ph = (t / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], STAGE_BYTES);
                bulk_g2s(sB + s * STAGE_BYTES, (const char*)p.cs + (int64_t)t * STAGE_BYTES, STAGE_BYTES, &full[s]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint64_t desc_ax = make_desc_sw32(smem_u32(sAX));
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % STAGES; const uint32_t ph = (t / STAGES) & 1;
                const int a = t % ACC_STAGES; const uint32_t aph = (t / ACC_STAGES) & 1;
                mbar_wait(&acc_empty[a], aph ^ 1);
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint64_t desc_b = make_desc_sw128(smem_u32(sB + s * STAGE_BYTES));
                const uint64_t desc_bx = make_desc_sw32(smem_u32(sB + s * STAGE_BYTES + B_BYTES));
#pragma unroll
                for (int h = 0; h < HALVES; ++h) {
                    const uint64_t desc_a = make_desc_sw128(smem_u32(sA + h * A_BYTES));
                    const uint32_t tacc = tmem_base + (uint32_t)((a * HALVES + h) * TILE_N);
#pragma unroll
                    for (int k = 0; k < KD / 8; ++k)     // UMMA_K = 8 tf32 = 32 bytes: advance the start address by 2 x 16 B
                        mma_tf32(tacc, desc_a + 2 * k, desc_b + 2 * k, IDESC, k > 0 ? 1u : 0u);
                    mma_tf32(tacc, desc_ax, desc_bx, IDESC, 1u);   // - ||c||^2 / 2
                }
                tc_commit(&empty[s]);        // frees the shared-memory stage once the MMAs have read it
                tc_commit(&acc_full[a]);     // accumulator complete
            }
        }
    } else {
        // ===== epilogue: thread = descriptor row; warp w may only touch TMEM lanes 32*(w%4) .. +31 =====
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int r = half * TILE_M + quad * 32 + lane;
        const int64_t row = row0 + r;
        const bool valid = row < p.B;
        float xn2 = 0.0f, xo2 = 0.0f, mu2 = 0.0f; bool finite = true;
#pragma unroll
        for (int d = 0; d < KD; ++d) {
            const float xd = (valid && d < p.Dd) ? p.desc[row * p.Dd + d] : 0.0f;
            const float m = d < p.Dd ? p.mu[d] : 0.0f;
            const float xc = xd - m;
            xn2 = __fmaf_rn(xc, xc, xn2); xo2 = __fmaf_rn(xd, xd, xo2); mu2 = __fmaf_rn(m, m, mu2);
            finite = finite && (fabsf(xd) <= 3.40282347e+38f);
        }
        const float cm2 = *p.cmax2;
        // band (in units of the distance d = ||x||^2 - 2 s) = 2*(error of d~ against the exact centred distance) + 2*(rounding
        // error of the reference's float32 sum):
        //   TF32:       |d~ - d| <= 2 * 2^-9 * 1.02 * ||x-mu|| * max||c-mu||   (two truncated operands per product, Cauchy-Schwarz)
        //   centring:   fl(x-mu), fl(c-mu) are off by <= 2^-24 (|x|+|mu|) per component -> <= 2^-19 (||x||+||mu||+1)(max||c-mu||+||x-mu||+1)
        //   reference:  <= 40 * 2^-24 * (||x-mu|| + max||c-mu||)^2
        // The tensor pass ranks s = x.c - ||c||^2/2 = (||x||^2 - d) / 2, so its band is half of that, plus the FP32 accumulation of
        // the 40 products inside the tensor core (<= 2^-16 of the largest partial sum, generously); -||c||^2/2 itself enters exactly.
        const float xn = __fsqrt_rn(xn2), cmx = __fsqrt_rn(cm2);
        const float band_d = 2.0f * (1.02f * 0x1p-8f * xn * cmx) + 0x1p-18f * (__fsqrt_rn(xo2) + __fsqrt_rn(mu2) + 1.0f) * (cmx + xn + 1.0f)
                           + 0x1p-17f * (xn + cmx) * (xn + cmx) + 1e-30f;
        const float band = 0.5f * band_d + 0x1p-16f * (xn * cmx + 0.5f * cm2);
        // A chunk with an admissible column is RECORDED in the row's slots in shared memory -- one 8-byte store of (chunk maximum,
        // chunk id, bit per column group that holds an admissible column) -- and the recorded groups are re-evaluated exactly at the
        // end.  Measured alternatives (r2_notes.md): a sorted register list of candidates costs a 16-level compare / select chain
        // per admission plus a 32-way register select, executed by the whole warp for one lane on the path that hands the
        // accumulator back to the MMA warp (half of the kernel's time); appending the columns themselves needs one store site per
        // column and unrolled copy, 7.9 k instructions, and stalls on instruction fetch.
        uint2* const my = cbuf + r;                        // slot e of this row: my[e * CTA_ROWS]
        int cnt = 0; bool lost = false;
        float best = -INFINITY, thr = -INFINITY;           // admit s~ > thr = best - band
        // One 32-column chunk.  Hot path: 16 three-input maxima (six groups of five columns, one of two) and one compare.
        // An admissible column turns up in ~10 % of the chunks for SOME lane of the warp (each row sees ~ln(chunks) running-maximum
        // records plus its band neighbours): the lane raises its threshold to the new maximum first, then records the chunk.
        auto process = [&](float (&acc)[32], int kbase) {
#define QDX_MX3(a_, b_, c_) fmaxf(fmaxf((a_), (b_)), (c_))
            float g[7];
#pragma unroll
            for (int q = 0; q < 6; ++q) g[q] = QDX_MX3(QDX_MX3(acc[5 * q], acc[5 * q + 1], acc[5 * q + 2]), acc[5 * q + 3], acc[5 * q + 4]);
            g[6] = fmaxf(acc[30], acc[31]);
            const float m = QDX_MX3(QDX_MX3(g[0], g[1], g[2]), QDX_MX3(g[3], g[4], g[5]), g[6]);
#undef QDX_MX3
            if (m > thr) {
                if (m > best) { best = m; thr = best - band; }
                uint32_t gm = 0;
#pragma unroll
                for (int q = 0; q < 7; ++q) gm |= (g[q] > thr ? 1u : 0u) << q;
                if (cnt == CAP) {                          // full: drop the chunks the threshold has overtaken since they were recorded
                    cnt = qdx_tc_compact(my, thr);
                    if (cnt == CAP) { lost = true; cnt = CAP - 1; }      // more than CAP chunks inside the band: exact fallback for this row
                }
                my[cnt * CTA_ROWS] = make_uint2(__float_as_uint(m), ((uint32_t)kbase << 2) | gm);     // kbase is a multiple of 32: (chunk id << 7) | groups
                ++cnt;
            }
        };

        // software pipeline over 32-column chunks (4 per tile, processed in ping-pong pairs): the TMEM load of chunk i+1 is in
        // flight while chunk i is processed
        float accA[32], accB[32];
        for (int t = 0; t < ntiles; ++t) {
            const int a = t % ACC_STAGES; const uint32_t aph = (t / ACC_STAGES) & 1;
            mbar_wait(&acc_full[a], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((a * HALVES + half) * TILE_N);
            const int kt = t * TILE_N;
            tmem_ld32(taddr, accA);
#pragma unroll
            for (int c0 = 0; c0 < TILE_N; c0 += 64) {
                tmem_wait_ld(accA);                                    // accA (chunk c0) has landed
                tmem_ld32(taddr + c0 + 32, accB);                      // chunk c0+32 in flight
                process(accA, kt + c0);
                tmem_wait_ld(accB);                                    // accB has landed
                if (c0 + 64 < TILE_N) tmem_ld32(taddr + c0 + 64, accA);
                process(accB, kt + c0 + 32);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[a]);
        }
        if (valid) {
            int32_t cell = 0; bool resolved = true;
            if (finite) {
                // every appended entry within the final band is a candidate; a row that overflowed its slots may have lost some
                if (lost || cnt == 0) resolved = false;
                else {
                    const float lim = best - band;
                    // flatten the recorded groups that survive the final band into a per-row list of centroid indices first (in
                    // the centroid ring: every copy has landed and every MMA has completed once this warp has seen the last
                    // accumulator), so that the lanes of the warp walk their lists in step -- evaluating straight out of the
                    // slots had each lane's one or two survivors at different slot / group positions and the warp paid for ~150
                    // exact distances instead of ~10 (r2_notes.md)
                    int32_t* const lst = reinterpret_cast<int32_t*>(sB) + r;       // element i of this row: lst[i * CTA_ROWS]
                    int n = 0;
#pragma unroll 1
                    for (int e = 0; e < cnt; ++e) {
                        const uint2 c = my[e * CTA_ROWS];
                        if (!(__uint_as_float(c.x) >= lim)) continue;          // overtaken by a later maximum
                        const int32_t k0 = (int32_t)(c.y >> 7) << 5;
#pragma unroll 1
                        for (int q = 0; q < 7; ++q) {
                            if (!((c.y >> q) & 1u)) continue;
                            const int32_t ke = min(k0 + (q < 6 ? 5 * q + 5 : 32), (int32_t)p.K);
                            for (int32_t ck = k0 + 5 * q; ck < ke; ++ck) {
                                if (n < LIST_CAP) { lst[n * CTA_ROWS] = ck; ++n; } else lost = true;
                            }
                        }
                    }
                    float x[KD];
#pragma unroll
                    for (int d = 0; d < KD; ++d) x[d] = d < p.Dd ? p.desc[row * p.Dd + d] : 0.0f;
                    float dbest = INFINITY; int32_t bk = 0x7fffffff;
#pragma unroll 1
                    for (int i = 0; i < n; ++i) {                               // the reference expression on every column of a recorded group
                        const int32_t ck = lst[i * CTA_ROWS];
                        const float dex = qdx_exact_dist<KD>(x, p.cent + (int64_t)ck * p.Dd, p.Dd);
                        if (dex < dbest || (dex == dbest && ck < bk)) { dbest = dex; bk = ck; }
                    }
                    if (lost) bk = 0x7fffffff;
                    if (bk == 0x7fffffff) resolved = false; else cell = bk;
                }
            }
            // non-finite descriptor: every distance is inf or NaN -> first index (centroids are finite)
            if (resolved) {
                p.cells[row] = cell;
                if (p.offer) qdx_offer(p.ws, p.K, p.rep_f, cell, p.fit[row], p.idx_base + (uint32_t)row, p.first_wins);
            } else {
                const int slot = atomicAdd(p.fallback_count, 1);
                p.fallback_rows[slot] = (int32_t)row;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// exact brute force for the (rare) rows the tensor pass could not resolve: one warp per listed row
__global__ void __launch_bounds__(128) qdx_cells_tc_fallback_kernel(const QdxTcParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int n = *p.fallback_count;
    for (int64_t i = w; i < n; i += nw) {
        const int64_t row = p.fallback_rows[i];
        const float* x = p.desc + row * p.Dd;
        float best = INFINITY; int64_t bk = 0x7fffffff;
        for (int64_t k = lane; k < p.K; k += 32) {
            const float* c = p.cent + k * p.Dd;
            float acc = 0.0f;
            for (int d = 0; d < p.Dd; ++d) { float df = x[d] - c[d]; float s = df * df; acc = d ? acc + s : s; }
            if (acc < best) { best = acc; bk = k; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const long long ok = __shfl_xor_sync(0xffffffffu, (long long)bk, o);
            if (ob < best || (ob == best && ok < bk)) { best = ob; bk = ok; }
        }
        if (lane == 0) {
            const int32_t cell = bk == 0x7fffffff ? 0 : (int32_t)bk;
            p.cells[row] = cell;
            if (p.offer) qdx_offer(p.ws, p.K, p.rep_f, cell, p.fit[row], p.idx_base + (uint32_t)row, p.first_wins);
        }
    }
}

extern "C" {

int qdx_cells_tc_workspace(int64_t K, int64_t B, int64_t* prep_floats, int64_t* scratch_ints) {
    if (K <= 0 || B < 0 || !prep_floats || !scratch_ints) return QDX_ERR_ARG;
    const int64_t Kpad = (K + tc::TILE_N - 1) / tc::TILE_N * tc::TILE_N;
    *prep_floats = Kpad * (tc::KD + tc::KX) + 64;  // per 128-centroid tile [swizzled centred centroids | -||c-mu||^2/2 block] | [max ||c-mu||^2, pad x31, mu x32]
    *scratch_ints = B + 64;                         // fallback rows | counter
    return 0;
}

int qdx_cells_tc_prepare(const float* centroids, int64_t K, int32_t desc_dim, float* prep, void* stream) {
    if (!centroids || !prep || K <= 0 || desc_dim < 1 || desc_dim > tc::KD) return QDX_ERR_ARG;
    const int64_t Kpad = (K + tc::TILE_N - 1) / tc::TILE_N * tc::TILE_N;
    float* cs = prep; float* cmax2 = prep + Kpad * (tc::KD + tc::KX);
    cudaError_t e = cudaMemsetAsync(cmax2, 0, 64 * sizeof(float), (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    float* mu = cmax2 + 32;
    qdx_cells_tc_mean_kernel<<<tc::KD, 256, 0, (cudaStream_t)stream>>>(centroids, K, desc_dim, mu);
    QDX_CHECK_LAUNCH();
    qdx_cells_tc_prepare_kernel<<<(unsigned)((Kpad + 255) / 256), 256, 0, (cudaStream_t)stream>>>(centroids, K, desc_dim, Kpad, mu, cs, cmax2);
    QDX_CHECK_LAUNCH();
    return 0;
}

int qdx_cells_tc(const float* desc, int64_t B, int32_t desc_dim, const float* centroids, int64_t K, const float* prep,
                 int32_t* scratch, int32_t* out_cells, void* ws, const float* rep_fitness, const float* fitness, int32_t offer,
                 uint32_t idx_base, int32_t first_wins, void* stream) {
    if (!desc || !centroids || !prep || !scratch || !out_cells || B < 0 || K <= 0 || desc_dim < 1 || desc_dim > tc::KD) return QDX_ERR_ARG;
    if (offer && (!ws || !rep_fitness || !fitness)) return QDX_ERR_ARG;
    if ((uint64_t)idx_base + (uint64_t)B > 0x7FFFFFFFull) return QDX_ERR_ARG;
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    QdxTcParams p;
    p.desc = desc; p.B = B; p.Dd = desc_dim; p.cent = centroids; p.K = K;
    p.Kpad = (K + tc::TILE_N - 1) / tc::TILE_N * tc::TILE_N;
    p.cs = prep; p.cmax2 = prep + p.Kpad * (tc::KD + tc::KX); p.mu = p.cmax2 + 32;
    p.cells = out_cells; p.fallback_rows = scratch; p.fallback_count = scratch + B;
    p.ws = ws; p.rep_f = rep_fitness; p.fit = fitness; p.offer = offer; p.idx_base = idx_base; p.first_wins = first_wins;
    cudaError_t e = cudaMemsetAsync(p.fallback_count, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(qdx_cells_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    qdx_cells_tc_kernel<<<(unsigned)((B + tc::CTA_ROWS - 1) / tc::CTA_ROWS), tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(p);
    QDX_CHECK_LAUNCH();
    qdx_cells_tc_fallback_kernel<<<148, 128, 0, st>>>(p);
    QDX_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
