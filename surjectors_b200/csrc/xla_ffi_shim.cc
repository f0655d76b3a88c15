// XLA FFI handlers around the C ABI (include/surj_b200.h): what `jax.ffi.register_ffi_target`
// would bind so that the reference's JAX code calls these kernels.
//
// The XLA FFI headers ship with jaxlib (jax.ffi.include_dir()); neither JAX nor the headers exist in
// this image, so this translation unit compiles to nothing here and is NOT part of the tested
// boundary -- the C ABI is.  It is kept as source so a maintainer with a JAX install can build it:
//   g++ -shared -fPIC -I$(python -c 'import jax.ffi; print(jax.ffi.include_dir())') -Iinclude \
//       surjectors_b200/csrc/xla_ffi_shim.cc -L surjectors_b200/_lib -lsurj_b200 -o libsurj_b200_ffi.so
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#include <cstdint>

#include "surj_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

// The plan handle travels as an int64 attribute (created on the host by the Python layer with
// surj_plan_create and kept alive by it), buffers are XLA-owned, the stream is XLA's.
static ffi::Error LogProbFwdImpl(cudaStream_t stream, int64_t plan, ffi::Buffer<ffi::F32> y, ffi::Buffer<ffi::F32> x,
                                 ffi::Buffer<ffi::F32> params, ffi::ResultBuffer<ffi::F32> z,
                                 ffi::ResultBuffer<ffi::F32> lp, ffi::ResultBuffer<ffi::U8> workspace) {
  const SurjPlan* p = reinterpret_cast<const SurjPlan*>(plan);
  const int64_t batch = y.dimensions()[0];
  const float* xp = x.element_count() ? x.typed_data() : nullptr;
  const int st = surj_chain_logprob_fwd(p, stream, y.typed_data(), xp, params.typed_data(), z->typed_data(),
                                        lp->typed_data(), workspace->untyped_data(), workspace->size_bytes(), batch);
  if (st != SURJ_OK) return ffi::Error(st == SURJ_E_INVALID ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                                       surj_last_error());
  return ffi::Error::Success();
}

static ffi::Error LogProbBwdImpl(cudaStream_t stream, int64_t plan, ffi::Buffer<ffi::F32> y, ffi::Buffer<ffi::F32> x,
                                 ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> g_lp, ffi::Buffer<ffi::F32> g_z,
                                 ffi::ResultBuffer<ffi::F32> g_params, ffi::ResultBuffer<ffi::F32> g_y,
                                 ffi::ResultBuffer<ffi::F32> g_x, ffi::ResultBuffer<ffi::U8> workspace) {
  const SurjPlan* p = reinterpret_cast<const SurjPlan*>(plan);
  const int64_t batch = y.dimensions()[0];
  const float* xp = x.element_count() ? x.typed_data() : nullptr;
  float* gxp = g_x->element_count() ? g_x->typed_data() : nullptr;
  const int st = surj_chain_logprob_bwd(p, stream, y.typed_data(), xp, params.typed_data(), g_lp.typed_data(), 1.0f,
                                        g_z.typed_data(), g_params->typed_data(), nullptr, nullptr, g_y->typed_data(), gxp,
                                        workspace->untyped_data(), workspace->size_bytes(), batch);
  if (st != SURJ_OK) return ffi::Error(st == SURJ_E_INVALID ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                                       surj_last_error());
  return ffi::Error::Success();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(SurjLogProbFwd, LogProbFwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("plan")
                                  .Arg<ffi::Buffer<ffi::F32>>()   // y [B, D]
                                  .Arg<ffi::Buffer<ffi::F32>>()   // x [B, Dx] (or [B, 0])
                                  .Arg<ffi::Buffer<ffi::F32>>()   // flat params
                                  .Ret<ffi::Buffer<ffi::F32>>()   // z [B, D_out]
                                  .Ret<ffi::Buffer<ffi::F32>>()   // lp [B]
                                  .Ret<ffi::Buffer<ffi::U8>>());  // workspace (surj_plan_workspace_bytes)

XLA_FFI_DEFINE_HANDLER_SYMBOL(SurjLogProbBwd, LogProbBwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("plan")
                                  .Arg<ffi::Buffer<ffi::F32>>()   // y
                                  .Arg<ffi::Buffer<ffi::F32>>()   // x
                                  .Arg<ffi::Buffer<ffi::F32>>()   // flat params
                                  .Arg<ffi::Buffer<ffi::F32>>()   // g_lp [B]
                                  .Arg<ffi::Buffer<ffi::F32>>()   // g_z [B, D_out]
                                  .Ret<ffi::Buffer<ffi::F32>>()   // g_params
                                  .Ret<ffi::Buffer<ffi::F32>>()   // g_y
                                  .Ret<ffi::Buffer<ffi::F32>>()   // g_x
                                  .Ret<ffi::Buffer<ffi::U8>>());  // workspace
#endif
#endif
