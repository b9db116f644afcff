// XLA FFI (jax.ffi) handlers over the C ABI of include/qdx.h -- the "thin jax.ffi (XLA custom-call, C-ABI) layer" that
// lets QDax's own Python/JAX host code call libqdx.so.  The reference (QDax 0.5.1 under /root/reference) has no FFI of
// its own; each handler below names the reference code it stands in for.
//
// NOT COMPILED OR EXECUTED IN THIS REPOSITORY'S ENVIRONMENT: jax / jaxlib and their xla/ffi headers are absent from the
// image and from the GPU box (no wheel, no network).  The translation unit is empty unless the headers are found; build it
// next to a JAX install with
//     make -C qdax_b200/csrc ffi XLA_FFI_INCLUDE=$(python -c "import jax; print(jax.ffi.include_dir())")
// and register the handlers with qdax_b200/jax_ffi.py.  Everything here only unpacks XLA buffers / attributes / the
// platform stream and forwards to the qdx_* launchers, which are the code paths the tests and benches of this repository
// exercise through ctypes.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define QDX_HAVE_XLA_FFI 1
#endif
#endif

#ifdef QDX_HAVE_XLA_FFI
#include <cstdint>
#include <string>

#include <cuda_runtime_api.h>

#include "xla/ffi/api/c_api.h"
#include "xla/ffi/api/ffi.h"

#include "../../include/qdx.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error Status(const char* what, int rc) {
    if (rc == 0) return ffi::Error::Success();
    return ffi::Error(ffi::ErrorCode::kInternal, std::string(what) + " failed, rc=" + std::to_string(rc));
}

// get_cells_indices(batch_of_descriptors, centroids) -- qdax/core/containers/mapelites_repertoire.py:111-137
ffi::Error CellsImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> desc, ffi::Buffer<ffi::F32> centroids,
                     ffi::ResultBuffer<ffi::S32> cells) {
    const int64_t B = desc.dimensions()[0], Dd = desc.dimensions()[1], K = centroids.dimensions()[0];
    return Status("qdx_cells", qdx_cells(desc.typed_data(), B, (int32_t)Dd, centroids.typed_data(), K, /*grid=*/nullptr,
                                         cells->typed_data(), nullptr, nullptr, nullptr, /*offer=*/0, 0u, 1, stream));
}

// arm / rastrigin / sphere scoring -- qdax/tasks/arm.py:9-50, qdax/tasks/standard_functions.py:9-48
ffi::Error ScoreImpl(cudaStream_t stream, int32_t task, ffi::Buffer<ffi::F32> genotypes, ffi::ResultBuffer<ffi::F32> fitness,
                     ffi::ResultBuffer<ffi::F32> desc) {
    const int64_t B = genotypes.dimensions()[0], D = genotypes.dimensions()[1];
    return Status("qdx_score", qdx_score(task, genotypes.typed_data(), B, D, (int32_t)desc->dimensions()[1], fitness->typed_data(),
                                         desc->typed_data(), stream));
}

// MapElitesRepertoire.add -- mapelites_repertoire.py:173-266.  The repertoire arrays are input/output aliased
// (input_output_aliases on the Python side), so the scatter happens in place; `ws` is the per-repertoire workspace buffer
// (qdx_workspace_bytes bytes, zero-initialised once), carried like the repertoire.
ffi::Error AddImpl(cudaStream_t stream, int32_t first_wins, float qd_offset, ffi::Buffer<ffi::U8> ws, ffi::Buffer<ffi::F32> rep_g,
                   ffi::Buffer<ffi::F32> rep_f, ffi::Buffer<ffi::F32> rep_d, ffi::Buffer<ffi::F32> centroids,
                   ffi::Buffer<ffi::F32> g, ffi::Buffer<ffi::F32> f, ffi::Buffer<ffi::F32> d, ffi::ResultBuffer<ffi::U8> ws_out,
                   ffi::ResultBuffer<ffi::F32> out_g, ffi::ResultBuffer<ffi::F32> out_f, ffi::ResultBuffer<ffi::F32> out_d,
                   ffi::ResultBuffer<ffi::S32> cells, ffi::ResultBuffer<ffi::F32> metrics) {
    const int64_t K = centroids.dimensions()[0], Dd = centroids.dimensions()[1], B = g.dimensions()[0];
    const int64_t D = g.element_count() / (B > 0 ? B : 1);
    (void)ws; (void)rep_g; (void)rep_f; (void)rep_d;                 // aliased to ws_out / out_g / out_f / out_d
    int rc = qdx_cells(d.typed_data(), B, (int32_t)Dd, centroids.typed_data(), K, nullptr, cells->typed_data(), ws_out->typed_data(),
                       out_f->typed_data(), f.typed_data(), /*offer=*/1, 0u, first_wins, stream);
    if (rc) return Status("qdx_cells", rc);
    return Status("qdx_commit", qdx_commit(ws_out->typed_data(), K, D, (int32_t)Dd, g.typed_data(), f.typed_data(), d.typed_data(), 0u, B,
                                           first_wins, out_g->typed_data(), out_f->typed_data(), out_d->typed_data(), qd_offset,
                                           metrics->typed_data(), nullptr, /*mode=*/0, stream));
}

// isoline_variation(x1, x2, key, ...) -- qdax/core/emitters/mutation_operators.py:175-226 (single-leaf genotype).  The key
// is a device buffer under jit, so it is read back here (8 bytes, synchronises the stream): use the fused generation
// below on the hot path.
ffi::Error IsolineImpl(cudaStream_t stream, float iso_sigma, float line_sigma, int32_t has_min, float minval, int32_t has_max,
                       float maxval, ffi::Buffer<ffi::F32> x1, ffi::Buffer<ffi::F32> x2, ffi::Buffer<ffi::U32> key,
                       ffi::ResultBuffer<ffi::F32> out) {
    uint32_t k[2];
    cudaError_t e = cudaMemcpyAsync(k, key.typed_data(), sizeof(k), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return Status("key read-back", (int)e);
    const int64_t B = x1.dimensions()[0];
    const int64_t D = x1.element_count() / (B > 0 ? B : 1);
    return Status("qdx_isoline_variation", qdx_isoline_variation(x1.typed_data(), x2.typed_data(), B, D, k[0], k[1], iso_sigma, line_sigma,
                                                                 has_min, minval, has_max, maxval, out->typed_data(), stream));
}

// One MAPElites.scan_update step -- qdax/core/map_elites.py:197-225 with MixingEmitter(variation_percentage = 1,
// isoline_variation), UniformSelector, a natively scored task and a grid tessellation: the whole generation (select x2,
// isoline, scoring, cell, segment_max, scatter, metrics) on XLA's stream with the key chain on the device.
// carry key in (device, 2 x u32) -> new carry key out; repertoire + workspace aliased in place.
//   axes: the concatenated per-dimension centroid coordinates of compute_euclidean_centroids (ascending), n / stride /
//   lo / hi as in qdx_grid_desc (include/qdx.h), passed as attributes because they are static per tessellation.
ffi::Error ScanUpdateImpl(cudaStream_t stream, int64_t batch, int32_t task, float iso_sigma, float line_sigma, int32_t has_min,
                          float minval, int32_t has_max, float maxval, int32_t first_wins, float qd_offset, int32_t n0, int32_t n1,
                          int32_t stride0, int32_t stride1, float lo0, float lo1, float hi0, float hi1, ffi::Buffer<ffi::U8> ws,
                          ffi::Buffer<ffi::U32> key, ffi::Buffer<ffi::F32> rep_g, ffi::Buffer<ffi::F32> rep_f, ffi::Buffer<ffi::F32> rep_d,
                          ffi::Buffer<ffi::F32> centroids, ffi::Buffer<ffi::F32> axes, ffi::ResultBuffer<ffi::U8> ws_out,
                          ffi::ResultBuffer<ffi::U32> key_out, ffi::ResultBuffer<ffi::F32> out_g, ffi::ResultBuffer<ffi::F32> out_f,
                          ffi::ResultBuffer<ffi::F32> out_d, ffi::ResultBuffer<ffi::F32> off_g, ffi::ResultBuffer<ffi::F32> off_f,
                          ffi::ResultBuffer<ffi::F32> off_d, ffi::ResultBuffer<ffi::F32> metrics) {
    (void)ws; (void)rep_g; (void)rep_f; (void)rep_d;                 // aliased to ws_out / out_g / out_f / out_d
    const int64_t K = centroids.dimensions()[0], D = out_g->dimensions()[1];
    void* w = ws_out->typed_data();
    qdx_grid_desc grid{};
    grid.dd = 2;
    grid.n[0] = n0; grid.n[1] = n1; grid.stride[0] = stride0; grid.stride[1] = stride1;
    grid.lo[0] = lo0; grid.lo[1] = lo1; grid.hi[0] = hi0; grid.hi[1] = hi1;
    grid.axes = axes.typed_data();
    int rc = qdx_workspace_copy_carry_key(w, key.typed_data(), /*to_workspace=*/1, stream);
    if (rc) return Status("qdx_workspace_copy_carry_key", rc);
    rc = qdx_select_prepare(out_f->typed_data(), K, w, /*key_mode=*/2, 0u, 0u, /*rank_slot=*/-1, stream);      // map_elites.py:214
    if (rc) return Status("qdx_select_prepare", rc);
    rc = qdx_generate(out_g->typed_data(), out_f->typed_data(), centroids.typed_data(), w, K, D, batch, iso_sigma, line_sigma, has_min,
                      minval, has_max, maxval, task, /*desc_dim=*/2, &grid, /*offer=*/1, 0u, first_wins, off_g->typed_data(),
                      off_f->typed_data(), off_d->typed_data(), nullptr, nullptr, nullptr, /*gen_keys8=*/nullptr, /*cvt=*/nullptr,
                      /*flags=*/QDX_GEN_ROWS_FIRED_ONLY, stream);
    if (rc) return Status("qdx_generate", rc);
    rc = qdx_commit(w, K, D, 2, off_g->typed_data(), off_f->typed_data(), off_d->typed_data(), 0u, batch, first_wins, out_g->typed_data(),
                    out_f->typed_data(), out_d->typed_data(), qd_offset, metrics->typed_data(), nullptr, 0, stream);
    if (rc) return Status("qdx_commit", rc);
    return Status("qdx_workspace_copy_carry_key", qdx_workspace_copy_carry_key(w, key_out->typed_data(), /*to_workspace=*/0, stream));
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(QdxCells, CellsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(QdxScore, ScoreImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int32_t>("task")
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(QdxAdd, AddImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int32_t>("first_wins")
                                  .Attr<float>("qd_offset")
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(QdxIsolineVariation, IsolineImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<float>("iso_sigma")
                                  .Attr<float>("line_sigma")
                                  .Attr<int32_t>("has_min")
                                  .Attr<float>("minval")
                                  .Attr<int32_t>("has_max")
                                  .Attr<float>("maxval")
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(QdxScanUpdate, ScanUpdateImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("batch")
                                  .Attr<int32_t>("task")
                                  .Attr<float>("iso_sigma")
                                  .Attr<float>("line_sigma")
                                  .Attr<int32_t>("has_min")
                                  .Attr<float>("minval")
                                  .Attr<int32_t>("has_max")
                                  .Attr<float>("maxval")
                                  .Attr<int32_t>("first_wins")
                                  .Attr<float>("qd_offset")
                                  .Attr<int32_t>("n0")
                                  .Attr<int32_t>("n1")
                                  .Attr<int32_t>("stride0")
                                  .Attr<int32_t>("stride1")
                                  .Attr<float>("lo0")
                                  .Attr<float>("lo1")
                                  .Attr<float>("hi0")
                                  .Attr<float>("hi1")
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::U32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

#endif  // QDX_HAVE_XLA_FFI
