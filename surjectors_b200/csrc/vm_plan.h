// Compiled chain program shared by the fused interpreter kernel (vm_plan.cu) and the per-op
// executor (vm_global.cu).
#pragma once
#include <string>
#include <vector>

#include "vm.cuh"

// Which OP_LINEAR ops of a program run on the pipelined tcgen05 kernel and where their tiled weight copy lives
// Structural zeros of a staged weight matrix at tile granularity (masked MADE weights are STAGE_ZERO in the gather table):
// K-block range [lo, hi) that can contribute to each 128-output tile of a LINEAR, and the all-zero 128x128 tiles of a
// LINEAR_DW.  Passed to the kernels by value; matrices with more tiles than fit simply run dense.
struct VmKRange {
  unsigned short lo[8], hi[8];
  int enabled;
};
// One fused conditioner + spline group of a program (vm_fused.cuh): ops [first, last] = LINEAR x (n_hidden + 1), RQS / RQS_BWD
struct VmFusedGroup {
  int first = 0, last = 0;       // op indices (inclusive)
  int mats[4] = {-1, -1, -1, -1}; // index into VmTcInfo::img_mats of every matrix
  double nnz = 0.0;              // structurally non-zero weights of its matrices (needed FLOPs per sample = 2 x this)
};
struct VmImgMat {  // mirrors vmf::ImgMat (kept POD here so that this header stays CUDA-free)
  int wt_off, ldt, n_rows, n_k, rows_per_block, n_blocks, N, k_pad, scale_slot;
  long long img_off;
};
// MaskedAutoregressive sampling recursion recognised in a program (vm_maf.cuh): ops [first, last] = OP_ZERO + D copies of the
// MADE block; desc = the kernel's descriptor (vmaf::Desc) as bytes so that this header stays CUDA-free
struct VmMafGroup {
  int first = 0, last = 0, gather = -1;  // gather: index of the context GATHER op of the first copy (-1: unconditional MADE)
  std::vector<unsigned char> desc;
  size_t smem = 0;
  int wbuf = 0;                          // floats of one weight staging buffer (the largest step's rows)
};
struct VmTcInfo {
  bool ready = false;
  std::vector<VmMafGroup> maf;             // by first op
  std::vector<int> maf_at;                 // per op: index into `maf` if a recursion starts there, else -1
  int fusion_flag = -1;                    // the fusion setting `fused` was computed under
  std::vector<VmFusedGroup> fused;         // fused conditioner + spline groups, by first op
  std::vector<int> fused_at;               // per op: index into `fused` if a group starts there, else -1
  std::vector<VmImgMat> img_mats;          // weight images of the fused groups (deduplicated by matrix)
  size_t img_bytes = 0;
  std::vector<VmKRange> kr;                // per op (LINEAR on the pipelined kernel)
  std::vector<unsigned long long> dw_skip; // per op (LINEAR_DW): bit (i_tile * n_j_tiles + j_tile) set = all-zero tile
  std::vector<double> nnz;                 // per op (LINEAR / LINEAR_DW): structurally non-zero weights of its matrix
  std::vector<long long> w_off;  // per op: float offset of its tiled [out][in] weights inside the weight-tile region, -1 = not a tensor-core op
  size_t w_floats = 0;           // size of the weight-tile region
  int max_kpad = 0;              // max over tensor-core ops of n_in rounded up to 32 (activation tiles need kpad * 2 floats per sample)
};

struct VmPlan {
  std::vector<vm::Op> ops_fwd, ops_bwd, ops_gen;  // log_prob, its VJP, sampling direction (empty: unsupported)
  std::string gen_why;
  int n_eps = 0;  // standard-normal draws per sample the sampling direction consumes (decoder samples)
  std::vector<int32_t> stage, grad_dst, grad_aux, itab;
  std::vector<vm::ScalarJob> sjobs;
  int n_scalars = 0, n_rows = 0;
  int d_in = 0, d_x = 0, d_out = 0, n_params = 0;
  // fused kernel launch configuration (fused_ok: parameters + one tile of rows fit in shared memory)
  bool fused_ok = false;
  std::string fused_why;
  int nt = 64, s = 1;
  size_t smem_fwd = 0, smem_bwd = 0, smem_gen = 0;
  // device state (uploaded lazily on the first launch)
  bool uploaded = false;
  void* d_blob = nullptr;
  vm::Program prog{};
  const vm::Op* d_ops_fwd = nullptr;
  const vm::Op* d_ops_bwd = nullptr;
  const vm::Op* d_ops_gen = nullptr;
  int n_sm = 148, occ_fwd = 1, occ_bwd = 1, occ_gen = 1;
  mutable VmTcInfo tc[3];  // per program (VM_PROG_*), filled on first use by the per-op executor
};

// vm_global.cu
enum { VM_PROG_FWD = 0, VM_PROG_BWD = 1, VM_PROG_GEN = 2 };
size_t vm_global_workspace_bytes(const VmPlan* v, int64_t B, bool backward);
int vm_global_launches(const VmPlan* v, int64_t B, bool backward);
int vm_global_maf_groups(const VmPlan* v);
int vm_global_run(VmPlan* v, cudaStream_t st, int prog, int with_base, const float* y, const float* x,
                  const float* params, const float* g_lp, float g_scale, const float* g_z, float* z_out, float* lp_out,
                  float* g_y, float* g_x, float* g_params, float* loss_out, void* ws, size_t ws_bytes, int64_t B,
                  const float* eps = nullptr);
