/* surj_b200.h -- C ABI of the B200 (sm_100a) implementation of surjectors' batched
 * flow log-likelihood and its gradient.
 *
 * This is the drop-in boundary for ONE path of dirmeier/surjectors v0.3.4 (all file:line
 * citations are relative to the reference checkout): `TransformedDistribution.log_prob`
 * through a `Chain` and its VJP.  In the reference the callee below this boundary is
 * XLA; the binding a maintainer would add is a `jax.ffi.ffi_call` per entry point
 * wrapped in `jax.custom_vjp`.  Everything is plain pointers and
 * sizes; there are no torch / JAX types in any signature.
 *
 * Conventions
 *   - all device arrays are float32, row-major, batch-leading: y[B, D], x[B, Dx], lp[B];
 *   - device pointers must be 16-byte aligned;
 *   - `params` is ONE flat float32 device buffer; every layer descriptor carries the float
 *     offsets of its haiku parameters inside it (w:[in,out] row-major, b:[out]);
 *   - every call only ENQUEUES work on `stream` (no host sync, no allocation on the hot path:
 *     scratch comes from the caller-provided workspace) and is CUDA-graph capturable;
 *   - functions return 0 on success, a negative SURJ_E_* code otherwise;
 *     `surj_last_error()` returns a thread-local message.
 */
#ifndef SURJ_B200_H_
#define SURJ_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SURJ_ABI_VERSION 1
#define SURJ_MAX_NET_LAYERS 8

enum SurjStatus {
  SURJ_OK = 0,
  SURJ_E_INVALID = -1,     /* bad descriptor / argument (reference: ValueError at construction) */
  SURJ_E_UNSUPPORTED = -2, /* structure the kernels do not implement (reference: NotImplementedError) */
  SURJ_E_CUDA = -3,        /* CUDA runtime error; message holds cudaGetErrorString */
  SURJ_E_WORKSPACE = -4    /* workspace too small / misaligned */
};

/* Layer kinds == the reference classes on the hot path. */
enum SurjLayerKind {
  SURJ_LAYER_COUPLING = 0,        /* MaskedCoupling                 _src/bijectors/masked_coupling.py:73-85 */
  SURJ_LAYER_COUPLING_FUNNEL = 1, /* MaskedCouplingInferenceFunnel  _src/surjectors/masked_coupling_inference_funnel.py:72-91 */
  SURJ_LAYER_MAF = 2,             /* MaskedAutoregressive           _src/bijectors/masked_autoregressive.py:73-79 */
  SURJ_LAYER_MAF_FUNNEL = 3,      /* MaskedAutoregressiveInferenceFunnel _src/surjectors/masked_autoregressive_inference_funnel.py:77-90 */
  SURJ_LAYER_SLICE = 4,           /* Slice                          _src/surjectors/slice.py:58-68 */
  SURJ_LAYER_LU = 5,              /* LULinear                       _src/bijectors/lu_linear.py:68-90 */
  SURJ_LAYER_PERMUTATION = 6,     /* Permutation                    _src/bijectors/permutation.py:30-37 */
  SURJ_LAYER_MLP_FUNNEL = 7       /* MLPInferenceFunnel             _src/surjectors/mlp.py:40-56: z = LULinear(y+) + Linear(y-);
                                     lu_* offsets = its LULinear, conditioner = the one hk.Linear [D-K -> K], decoder sees z only */
};

/* Inner bijector returned by the user's bijector_fn (distrax semantics). */
enum SurjInner {
  SURJ_INNER_AFFINE = 0,     /* ScalarAffine(m, exp(s)):           z = (y - m) e^{-s}, ld = -s */
  SURJ_INNER_AFFINE_INV = 1, /* Inverse(ScalarAffine(m, exp(s))):  z = e^{s} y + m,    ld = +s */
  SURJ_INNER_RQS = 2,        /* RationalQuadraticSpline(theta, lo, hi): spline inverse */
  SURJ_INNER_RQS_INV = 3     /* Inverse(RationalQuadraticSpline):  spline forward */
};

/* How bijector_fn slices the conditioner output for affine bijectors (quirk Q5). */
enum SurjLayout {
  SURJ_LAYOUT_SPLIT = 0,      /* jnp.split(theta, 2, -1): shift = theta[:D], log_scale = theta[D:] */
  SURJ_LAYOUT_INTERLEAVED = 1 /* unstack(theta[.., D, 2], -1): shift = theta[2d], log_scale = theta[2d+1] */
};

enum SurjActivation {
  SURJ_ACT_IDENTITY = 0, SURJ_ACT_RELU = 1, SURJ_ACT_SWISH = 2, SURJ_ACT_TANH = 3, SURJ_ACT_SIGMOID = 4,
  SURJ_ACT_SOFTPLUS = 5, SURJ_ACT_GELU = 6, SURJ_ACT_ELU = 7, SURJ_ACT_LEAKY_RELU = 8
};

/* hk.nets.MLP / make_mlp (_src/conditioners/mlp.py:7-29), possibly followed by hk.Linear
 * layers inside an hk.Sequential: a stack of n_layers Linear layers. */
typedef struct SurjMlpDesc {
  int32_t n_layers;                         /* 0 = absent */
  int32_t widths[SURJ_MAX_NET_LAYERS + 1];  /* widths[0] = input width; widths[l+1] = output of layer l */
  int32_t act[SURJ_MAX_NET_LAYERS];         /* activation applied after layer l */
  int64_t w_off[SURJ_MAX_NET_LAYERS];       /* float offset of w[widths[l], widths[l+1]] */
  int64_t b_off[SURJ_MAX_NET_LAYERS];       /* float offset of b[widths[l+1]]; -1 = no bias */
} SurjMlpDesc;

/* MADE (_src/conditioners/nn/made.py:26-91).  Masks are rebuilt from (input_size, hidden,
 * n_params) with TFP's _make_dense_autoregressive_masks algorithm (made.py:54-56). */
typedef struct SurjMadeDesc {
  int32_t input_size;                       /* 0 = absent */
  int32_t n_params;
  int32_t n_hidden;
  int32_t hidden[SURJ_MAX_NET_LAYERS];
  int32_t act;                              /* between masked layers, not after the last */
  int32_t ctx_dim;                          /* width of the conditioning input of the context Linear; 0 = none */
  int64_t w_off[SURJ_MAX_NET_LAYERS + 1];   /* n_hidden + 1 masked layers */
  int64_t b_off[SURJ_MAX_NET_LAYERS + 1];
  int64_t ctx_w_off, ctx_b_off;             /* hk.Linear(hidden[0]) on the context (made.py:80-84) */
} SurjMadeDesc;

typedef struct SurjLayerDesc {
  int32_t kind;        /* SurjLayerKind */
  int32_t dim_in;      /* event width entering the layer */
  int32_t dim_out;     /* event width leaving it (n_keep for Slice / funnels) */
  int32_t use_x;       /* 1 if the layer consumes the chain's conditioning input x */
  int32_t inner;       /* SurjInner */
  int32_t layout;      /* SurjLayout (affine inners on couplings; MADE output is always interleaved) */
  int32_t n_bins;      /* RQS: K, conditioner emits 3K+1 values per dim */
  float range_min, range_max;
  const uint8_t* mask;         /* COUPLING: [dim_in], 1 = passes through & feeds the conditioner */
  const int32_t* permutation;  /* PERMUTATION: [dim_in] (the forward permutation, as the reference stores it) */
  SurjMlpDesc conditioner;     /* COUPLING / COUPLING_FUNNEL */
  SurjMadeDesc made;           /* MAF / MAF_FUNNEL */
  SurjMlpDesc decoder;         /* SLICE / funnels: MLP -> split -> Independent(Normal(mu, exp(log_sigma))) */
  int64_t lu_lower_off, lu_upper_off, lu_diag_off; /* LU: haiku params lower_entries, upper_entries, diag_entries */
} SurjLayerDesc;

typedef struct SurjPlan SurjPlan;
typedef void* SurjStream; /* cudaStream_t */

enum SurjPlanMode {
  SURJ_MODE_AUTO = 0,  /* fused whole-chain kernel when one is compiled in for this chain, else per-layer */
  SURJ_MODE_FUSED = 1, /* require the fused whole-chain kernel (error if not compiled in) */
  SURJ_MODE_LAYERED = 2, /* generic per-layer kernels (any width) */
  SURJ_MODE_FUSED_GEN1 = 3, /* the build-time generated straight-line kernel for this chain (A/B comparison) */
  SURJ_MODE_FUSED_COMPACT = 4 /* the compact looped kernel for affine coupling / coupling-funnel chains (error if outside its family) */
};

/* What is added to the summed likelihood contributions of the layers. */
enum SurjBase {
  SURJ_BASE_STANDARD_NORMAL = 0, /* lp = Independent(Normal(0, 1)).log_prob(z) + lc   (TransformedDistribution.log_prob) */
  SURJ_BASE_NONE = 1             /* lp = lc   (Surjector.inverse_and_likelihood_contribution on its own) */
};

/* --- plan: replaces tracing the chain under jax.jit (examples/coupling_inference_surjection.py:81) */
int surj_plan_create(const SurjLayerDesc* layers, int32_t n_layers, int32_t dim_in, int32_t dim_x, int64_t n_params,
                     int32_t mode, int32_t base, SurjPlan** out);
void surj_plan_destroy(SurjPlan* plan);
/* canonical 64-bit key of the chain structure (what fused kernels are registered under) */
uint64_t surj_plan_key(const SurjPlan* plan);
int surj_chain_key(const SurjLayerDesc* layers, int32_t n_layers, int32_t dim_in, int32_t dim_x, int64_t n_params,
                   uint64_t* key_out);
/* 0 = per-layer kernels, 1 = generated straight-line fused kernel, 2 = fused interpreter kernel, 3 = compact looped fused kernel */
int32_t surj_plan_is_fused(const SurjPlan* plan);
int32_t surj_plan_dim_out(const SurjPlan* plan);
/* number of kernels one fwd / one bwd call launches (for launch accounting) */
int32_t surj_plan_launches(const SurjPlan* plan, int32_t backward);
/* MaskedAutoregressive layers whose sampling direction (masked_autoregressive.py:63-71, D MADE passes upstream) the
 * per-op executor runs as ONE incremental kernel (0 on the fused executors, which keep the D passes) */
int32_t surj_plan_sampling_recursions(const SurjPlan* plan);

/* bytes of device scratch a call with batch B needs (backward != 0 for the VJP) */
size_t surj_plan_workspace_bytes(const SurjPlan* plan, int64_t batch, int32_t backward);

/* --- TransformedDistribution.inverse_and_log_prob (_src/distributions/transformed_distribution.py:72-93)
 * with base = Independent(Normal(0, 1)).  z_out may be NULL.  x may be NULL iff dim_x == 0. */
int surj_chain_logprob_fwd(const SurjPlan* plan, SurjStream stream, const float* y, const float* x, const float* params,
                           float* z_out, float* lp_out, void* workspace, size_t workspace_bytes, int64_t batch);

/* --- VJP of (z, lp) w.r.t. params (and optionally y, x): what jax.value_and_grad differentiates at
 * examples/coupling_inference_surjection.py:83-87.  Cotangent of lp is g_lp[B], or the constant
 * g_scale when g_lp == NULL (g_scale = -1 gives loss = -sum(lp)); g_z[B, dim_out] is the optional
 * cotangent of the latent z (NULL = zero).  Outputs:
 *   g_params[n_params]  (overwritten),  loss_out[1] = g_scale * sum(lp) when g_lp == NULL (may be NULL),
 *   lp_out[B], g_y[B, D], g_x[B, Dx]  optional (NULL to skip). */
int surj_chain_logprob_bwd(const SurjPlan* plan, SurjStream stream, const float* y, const float* x, const float* params,
                           const float* g_lp, float g_scale, const float* g_z, float* g_params, float* loss_out,
                           float* lp_out,
                           float* g_y, float* g_x, void* workspace, size_t workspace_bytes, int64_t batch);

/* --- optax.adam(lr) update (optax 0.2.4: b1, b2, eps, bias-corrected, eps_root = 0), fused over the
 * flat buffers; `count` is a device int32 that is incremented by the kernel. */
int surj_adam_step(SurjStream stream, float* params, const float* grads, float* mu, float* nu, int32_t* count,
                   int64_t n, float lr, float b1, float b2, float eps, float grad_scale);

/* --- sampling direction: Surjector.forward_and_likelihood_contribution through the chain in reverse
 * (surjectors/_src/surjectors/chain.py:47-67): y_out[B, dim_in] and the summed forward log-det
 * ld_out[B] for latents z[B, dim_out].  Implemented for MaskedCoupling, MaskedAutoregressive (its D-pass
 * recursion, masked_autoregressive.py:63-71), Permutation, Slice (slice.py:70-78) and the coupling funnels
 * (masked_coupling_inference_funnel.py:93-106); SURJ_E_UNSUPPORTED for LULinear (no forward upstream either,
 * lu_linear.py:92-93) and the autoregressive funnels (broken upstream). */
int surj_chain_forward(const SurjPlan* plan, SurjStream stream, const float* z, const float* x, const float* eps,
                       const float* params, float* y_out, float* ld_out, void* workspace, size_t workspace_bytes,
                       int64_t batch);
/* standard-normal draws per sample that the sampling direction consumes (sum of the dropped widths of Slice /
 * coupling funnels, assigned in sampling order: last layer first); eps[B, noise_dim] comes from the caller's RNG --
 * the reference draws them with hk.next_rng_key() (slice.py:74-76), so a jax.ffi binding can pass its own stream */
int32_t surj_plan_noise_dim(const SurjPlan* plan);
size_t surj_plan_forward_workspace_bytes(const SurjPlan* plan, int64_t batch);

/* --- one training step = the jitted `step` of examples/coupling_inference_surjection.py:81-90:
 * loss = g_scale * sum(lp) (g_scale = -1: the examples' -sum; -1/B: the notebooks' -mean), its gradient
 * w.r.t. params (left in g_params) and the optax.adam update of params / mu / nu / count.  On the fused
 * executors the update is fused into the last stage of the gradient reduction (2 launches per step).
 * Data-parallel callers use surj_chain_logprob_bwd + all-reduce + surj_adam_step instead. */
int surj_train_step(const SurjPlan* plan, SurjStream stream, const float* y, const float* x, float* params,
                    float g_scale, float* g_params, float* mu, float* nu, int32_t* count, float lr, float b1, float b2,
                    float eps, float* loss_out, void* workspace, size_t workspace_bytes, int64_t batch);

/* --- data-parallel training: one process per GPU, batch sharded, parameters replicated.
 * For KB-sized gradients the all-reduce is fused into the gradient-reduction kernel over NVLink peer memory:
 * surj_dp_create allocates this rank's mailbox and returns its CUDA IPC handle; the handles of all ranks (in rank
 * order, exchanged by the host, e.g. torch.distributed.all_gather) go to surj_dp_connect; surj_train_step_dp is
 * surj_train_step on the local shard with the sum over ranks inside (2 launches, no NCCL call).  Every rank must
 * call it the same number of times.  A peer that does not arrive within the timeout (default 30 s,
 * SURJ_B200_DP_TIMEOUT_MS or surj_dp_set_timeout_ms; 0 = wait for ever, as a NCCL collective would) sets a sticky
 * error: the step that timed out and every later one leave params / mu / nu / count untouched (no stale value is
 * ever summed in), surj_dp_error() (synchronising) and surj_dp_error_nosync() (a host-mapped flag, cheap enough to
 * poll before every step) return 1, and the caller must stop. */
#define SURJ_IPC_HANDLE_BYTES 64
typedef struct SurjDp SurjDp;
int surj_dp_create(int32_t rank, int32_t world, int64_t n_params, SurjDp** out, unsigned char* handle_out);
int surj_dp_connect(SurjDp* dp, const unsigned char* handles);
void surj_dp_destroy(SurjDp* dp);
int32_t surj_dp_error(const SurjDp* dp);
int32_t surj_dp_error_nosync(const SurjDp* dp);
int surj_dp_set_timeout_ms(SurjDp* dp, int64_t ms);
int surj_train_step_dp(const SurjPlan* plan, SurjDp* dp, SurjStream stream, const float* y, const float* x, float* params,
                       float g_scale, float* g_params, float* mu, float* nu, int32_t* count, float lr, float b1,
                       float b2, float eps, float* loss_out, void* workspace, size_t workspace_bytes, int64_t batch);

/* --- helpers */
/* TFP _make_dense_autoregressive_masks restated; writes n_hidden+1 masks ([in, out] row-major,
 * bytes 0/1) back to back into out (size queried by passing out == NULL). */
int64_t surj_made_masks(int32_t n_params, int32_t input_size, const int32_t* hidden, int32_t n_hidden, uint8_t* out);
/* Permutation inverse index exactly as permutation.py:31-36 computes it (int32). */
int surj_permutation_inverse(const int32_t* permutation, int32_t n, int32_t* inverse_out);
/* as_batch_iterator's batch selection (surjectors/util.py:88-95, lax.index_take on axis 0) on the device:
 * dst[r][:] = src[idx[first + r]][:] for r < n_take; src is [n_rows, width] float32, idx int32 (all device pointers).
 * Bit-exact copy, enqueue-only.  An index outside [0, n_rows) writes a zero row and sets *err_flag (device int32, may be
 * NULL) to 1 instead of reading out of bounds. */
int surj_index_take(SurjStream stream, const float* src, int64_t n_rows, int32_t width, const int32_t* idx, int64_t first,
                    int64_t n_take, float* dst, int32_t* err_flag);

/* --- measurement helper (bench.py): a register-only FFMA loop, 16 independent chains per thread,
 * launched over every SM; *flops_out = floating-point operations one launch executes.  Gives the
 * measured FP32-pipe peak the small-network kernels are compared against. */
int surj_bench_fp32_peak(SurjStream stream, int32_t iters, float* sink, double* flops_out);
/* the same for the tensor pipe: one CTA per SM issues iters x 4 tcgen05.mma.kind::tf32 of 128 x 256 x 8 (A in tensor
 * memory, B in shared memory); *flops_out = dense TF32 FLOPs of one launch.  The wide-network kernels compute FP32-grade
 * products as 3 TF32 MMAs, so their roofline peak is this rate / 3. */
int surj_bench_tf32_peak(SurjStream stream, int32_t iters, float* sink, double* flops_out);

/* --- per-kernel timing of the per-op executor (bench.py's roofline leg; NOT for CUDA-graph capture): between
 * surj_profile_begin() and surj_profile_end() every launch of the per-op executor is bracketed by CUDA events on its
 * stream.  surj_profile_end synchronises, writes a JSON object {"kernel kind": {"launches", "ms", "flops"}, ...} into
 * out (flops = needed FLOPs: structurally non-zero multiply-adds x 2) and returns the number of bytes written
 * (0: buffer too small / profiling was not on). */
int surj_profile_begin(void);
int64_t surj_profile_end(char* out, int64_t out_bytes);

/* --- tuning / test knob of the per-op executor: the number of samples whose rows are resident in the workspace at a
 * time (rounded up to 128; default 262144, SURJ_B200_CHUNK at load time; 0 restores the default).  Returns the
 * previous value.  Changes surj_plan_workspace_bytes: query it again after setting. */
int64_t surj_set_chunk_samples(int64_t samples);
/* the fused conditioner + spline kernel of the per-op executor on (1, default; SURJ_B200_FUSION at load time) / off (0:
 * every op is its own launch, theta goes through the workspace).  Returns the previous setting.  Changes
 * surj_plan_workspace_bytes: query it again after setting. */
int32_t surj_set_wide_fusion(int32_t on);

const char* surj_last_error(void);
int32_t surj_abi_version(void);
/* number of fused whole-chain kernels compiled into this library */
int32_t surj_registry_size(void);

#ifdef __cplusplus
}
#endif
#endif /* SURJ_B200_H_ */
