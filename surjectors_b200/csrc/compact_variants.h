// Instantiation table of the compact fused kernel (compact.cuh): hidden widths (conditioner, decoder) x activation.
#pragma once
#include "compact.cuh"

namespace cc {

typedef void (*Kernel)(const Program, const Io);
struct Variant {
  int hc, hd, act;
  Kernel fn;
};
constexpr int N_VARIANTS = 5;
#define CC_VARIANT(HC, HD, ACT) {HC, HD, ACT, &cc::chain_kernel<HC, HD, ACT>}
#define CC_VARIANTS_FOR(ACT) \
  {CC_VARIANT(4, 4, ACT), CC_VARIANT(8, 4, ACT), CC_VARIANT(8, 8, ACT), CC_VARIANT(16, 8, ACT), CC_VARIANT(16, 16, ACT)}

const Variant* variants_relu();
const Variant* variants_swish();
const Variant* variants_tanh();

}  // namespace cc
