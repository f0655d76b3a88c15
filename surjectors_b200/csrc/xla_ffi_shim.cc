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

// ---- the rest of the compute entry points (same conventions: plan handle as an int64 attribute, XLA-owned buffers,
// XLA's stream; in-place updates are expressed as outputs aliased to inputs with `input_output_aliases` on the Python
// side, so the handler writes the result buffers after copying nothing: XLA passes the same device pointers)
static ffi::Error status_of(int st) {
  if (st == SURJ_OK) return ffi::Error::Success();
  return ffi::Error(st == SURJ_E_INVALID ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal, surj_last_error());
}

// optax.adam update (examples/coupling_inference_surjection.py:78,88-89): params / mu / nu / count are aliased in -> out
static ffi::Error AdamStepImpl(cudaStream_t stream, float lr, float b1, float b2, float eps, float grad_scale,
                               ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> grads, ffi::Buffer<ffi::F32> mu,
                               ffi::Buffer<ffi::F32> nu, ffi::Buffer<ffi::S32> count, ffi::ResultBuffer<ffi::F32> params_out,
                               ffi::ResultBuffer<ffi::F32> mu_out, ffi::ResultBuffer<ffi::F32> nu_out,
                               ffi::ResultBuffer<ffi::S32> count_out) {
  if (params_out->typed_data() != params.typed_data() || mu_out->typed_data() != mu.typed_data() ||
      nu_out->typed_data() != nu.typed_data() || count_out->typed_data() != count.typed_data())
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "surj_adam_step updates in place: alias outputs to inputs");
  return status_of(surj_adam_step(stream, params_out->typed_data(), grads.typed_data(), mu_out->typed_data(), nu_out->typed_data(),
                                  count_out->typed_data(), (int64_t)params.element_count(), lr, b1, b2, eps, grad_scale));
}

// the whole jitted `step` (examples/coupling_inference_surjection.py:81-90): params / mu / nu / count aliased in -> out
static ffi::Error TrainStepImpl(cudaStream_t stream, int64_t plan, float g_scale, float lr, float b1, float b2, float eps,
                                ffi::Buffer<ffi::F32> y, ffi::Buffer<ffi::F32> x, ffi::Buffer<ffi::F32> params,
                                ffi::Buffer<ffi::F32> mu, ffi::Buffer<ffi::F32> nu, ffi::Buffer<ffi::S32> count,
                                ffi::ResultBuffer<ffi::F32> params_out, ffi::ResultBuffer<ffi::F32> mu_out,
                                ffi::ResultBuffer<ffi::F32> nu_out, ffi::ResultBuffer<ffi::S32> count_out,
                                ffi::ResultBuffer<ffi::F32> g_params, ffi::ResultBuffer<ffi::F32> loss,
                                ffi::ResultBuffer<ffi::U8> workspace) {
  if (params_out->typed_data() != params.typed_data() || mu_out->typed_data() != mu.typed_data() ||
      nu_out->typed_data() != nu.typed_data() || count_out->typed_data() != count.typed_data())
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "surj_train_step updates in place: alias outputs to inputs");
  const SurjPlan* p = reinterpret_cast<const SurjPlan*>(plan);
  const float* xp = x.element_count() ? x.typed_data() : nullptr;
  return status_of(surj_train_step(p, stream, y.typed_data(), xp, params_out->typed_data(), g_scale, g_params->typed_data(),
                                   mu_out->typed_data(), nu_out->typed_data(), count_out->typed_data(), lr, b1, b2, eps,
                                   loss->typed_data(), workspace->untyped_data(), workspace->size_bytes(), y.dimensions()[0]));
}

// sampling direction (chain.py:47-67, transformed_distribution.py:95-139); eps = the caller's standard-normal draws
static ffi::Error ChainForwardImpl(cudaStream_t stream, int64_t plan, ffi::Buffer<ffi::F32> z, ffi::Buffer<ffi::F32> x,
                                   ffi::Buffer<ffi::F32> eps, ffi::Buffer<ffi::F32> params, ffi::ResultBuffer<ffi::F32> y,
                                   ffi::ResultBuffer<ffi::F32> ld, ffi::ResultBuffer<ffi::U8> workspace) {
  const SurjPlan* p = reinterpret_cast<const SurjPlan*>(plan);
  const float* xp = x.element_count() ? x.typed_data() : nullptr;
  const float* ep = eps.element_count() ? eps.typed_data() : nullptr;
  return status_of(surj_chain_forward(p, stream, z.typed_data(), xp, ep, params.typed_data(), y->typed_data(), ld->typed_data(),
                                      workspace->untyped_data(), workspace->size_bytes(), z.dimensions()[0]));
}

// as_batch_iterator's batch selection (surjectors/util.py:88-95)
static ffi::Error IndexTakeImpl(cudaStream_t stream, int64_t first, ffi::Buffer<ffi::F32> src, ffi::Buffer<ffi::S32> idx,
                                ffi::ResultBuffer<ffi::F32> dst, ffi::ResultBuffer<ffi::S32> err_flag) {
  const int64_t n_rows = src.dimensions()[0], n_take = dst->dimensions()[0];
  const int64_t width = n_rows ? (int64_t)src.element_count() / n_rows : 0;
  return status_of(surj_index_take(stream, src.typed_data(), n_rows, (int32_t)width, idx.typed_data(), first, n_take,
                                   dst->typed_data(), err_flag->typed_data()));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(SurjAdamStep, AdamStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<float>("lr").Attr<float>("b1").Attr<float>("b2").Attr<float>("eps").Attr<float>("grad_scale")
                                  .Arg<ffi::Buffer<ffi::F32>>()   // params (aliased to output 0)
                                  .Arg<ffi::Buffer<ffi::F32>>()   // grads
                                  .Arg<ffi::Buffer<ffi::F32>>()   // mu (aliased to output 1)
                                  .Arg<ffi::Buffer<ffi::F32>>()   // nu (aliased to output 2)
                                  .Arg<ffi::Buffer<ffi::S32>>()   // count [1] (aliased to output 3)
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(SurjTrainStep, TrainStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("plan")
                                  .Attr<float>("g_scale").Attr<float>("lr").Attr<float>("b1").Attr<float>("b2").Attr<float>("eps")
                                  .Arg<ffi::Buffer<ffi::F32>>()   // y
                                  .Arg<ffi::Buffer<ffi::F32>>()   // x
                                  .Arg<ffi::Buffer<ffi::F32>>()   // params (aliased to output 0)
                                  .Arg<ffi::Buffer<ffi::F32>>()   // mu (aliased to output 1)
                                  .Arg<ffi::Buffer<ffi::F32>>()   // nu (aliased to output 2)
                                  .Arg<ffi::Buffer<ffi::S32>>()   // count (aliased to output 3)
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()   // g_params [n_params + 4]
                                  .Ret<ffi::Buffer<ffi::F32>>()   // loss [1]
                                  .Ret<ffi::Buffer<ffi::U8>>());  // workspace (surj_plan_workspace_bytes(plan, B, 1))

XLA_FFI_DEFINE_HANDLER_SYMBOL(SurjChainForward, ChainForwardImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("plan")
                                  .Arg<ffi::Buffer<ffi::F32>>()   // z [B, D_out]
                                  .Arg<ffi::Buffer<ffi::F32>>()   // x [B, Dx] (or [B, 0])
                                  .Arg<ffi::Buffer<ffi::F32>>()   // eps [B, surj_plan_noise_dim] (or [B, 0])
                                  .Arg<ffi::Buffer<ffi::F32>>()   // flat params
                                  .Ret<ffi::Buffer<ffi::F32>>()   // y [B, D]
                                  .Ret<ffi::Buffer<ffi::F32>>()   // forward log-det [B]
                                  .Ret<ffi::Buffer<ffi::U8>>());  // workspace (surj_plan_forward_workspace_bytes)

XLA_FFI_DEFINE_HANDLER_SYMBOL(SurjIndexTake, IndexTakeImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("first")
                                  .Arg<ffi::Buffer<ffi::F32>>()   // data set field [n, ...]
                                  .Arg<ffi::Buffer<ffi::S32>>()   // idxs [n]
                                  .Ret<ffi::Buffer<ffi::F32>>()   // batch [step, ...]
                                  .Ret<ffi::Buffer<ffi::S32>>()); // out-of-range flag [1]
#endif
#endif
