// The step either side of the hot path in the reference's training loops: as_batch_iterator's batch selection
// (surjectors/util.py:88-95: `lax.index_take(array, (ret_idx,), axes=(0,))` for every field of the data set) as an
// on-device row gather, so that the shuffled batches of an epoch are cut out of a data set that already lives in HBM.
//
//   dst[r][:] = src[idx[first + r]][:]      r < n_take,  rows of `width` float32
//
// Pure data movement, bit-exact; HBM-bound (8 * width bytes per row, the indices are noise): a warp moves one row at a
// time with 128-bit loads / stores when the row pitch allows, consecutive warps take consecutive output rows.
#include <cuda_runtime.h>
#include <stdint.h>

#include "plan.h"

namespace {
template <typename V>
__global__ void __launch_bounds__(256)
surj_index_take_kernel(const V* __restrict__ src, int64_t n_rows, int width_v, const int32_t* __restrict__ idx, int64_t n_take,
                       V* __restrict__ dst, int* __restrict__ err) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_take; r += n_warps) {
    const int64_t s = idx[r];
    if (s < 0 || s >= n_rows) {  // never read out of bounds: flag it, leave zeros (kernels never abort)
      if (lane == 0 && err != nullptr) *err = 1;
      for (int c = lane; c < width_v; c += 32) dst[r * width_v + c] = V{};
      continue;
    }
    const V* p = src + s * width_v;
    V* q = dst + r * width_v;
    for (int c = lane; c < width_v; c += 32) q[c] = __ldg(p + c);
  }
}
}  // namespace

extern "C" int surj_index_take(SurjStream stream, const float* src, int64_t n_rows, int32_t width, const int32_t* idx, int64_t first,
                               int64_t n_take, float* dst, int32_t* err_flag) {
  if (!src || !idx || !dst || n_rows <= 0 || width <= 0 || first < 0 || n_take < 0) {
    surj_set_error("surj_index_take: bad arguments");
    return SURJ_E_INVALID;
  }
  if (n_take == 0) return SURJ_OK;
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const int64_t warps_wanted = n_take;
  const int grid = (int)std::min<int64_t>((warps_wanted + 7) / 8, (int64_t)n_sm * 8);
  const bool v4 = (width % 4 == 0) && (((uintptr_t)src | (uintptr_t)dst) % 16 == 0);
  if (v4)
    surj_index_take_kernel<float4><<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)src, n_rows, width / 4, idx + first, n_take,
                                                                         (float4*)dst, err_flag);
  else
    surj_index_take_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(src, n_rows, width, idx + first, n_take, dst, err_flag);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return surj_cuda_fail(e, "surj_index_take launch");
  return SURJ_OK;
}
