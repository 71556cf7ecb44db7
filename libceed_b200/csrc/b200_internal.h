// b200_internal.h -- private object model behind include/ceed_b200.h
#pragma once

#include "b200_driver.h"
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/ceed_b200.h"

struct B200Module {
  CUmodule    module = nullptr;
  std::string source;
  std::string log;
  // fused operator kernels: device address of the __constant__ argument block and the bytes last written to it
  CUdeviceptr       args_dptr = 0;
  std::vector<char> last_args;
};

// Shape of the generated operator kernel.  0 / -1 = let the plan heuristics decide.
struct B200Tuning {
  int epw = 0;          // elements per group
  int group_warps = 0;  // warps that share one element group (1, 2, 4)
  int cta_warps = 0;    // warps per CTA
  int minb = 0;         // __launch_bounds__ min blocks per SM
  int qf_mode = -1;     // 0 z-line QFunction stage, 1 pointwise
  int qf_unroll = 0;    // pointwise stage: points in flight per lane
  int stage = -1;       // cp.async staging mask
};

struct B200Ceed_ {
  int          device_id  = 0;
  cudaStream_t stream     = nullptr;
  int          num_sms    = 148;
  size_t       smem_optin = 0;  // max dynamic smem per block (opt-in)
  size_t       smem_sm    = 0;  // smem per SM
  int          cc_major = 0, cc_minor = 0;
  int          scatter_mode = B200_SCATTER_DETERMINISTIC;
  int64_t      launch_count = 0;
  std::string  last_error;
  std::vector<std::string>           jit_roots;
  std::vector<std::string>           jit_defines;
  std::map<std::string, B200Module *> module_cache;  // keyed on full source + options
  std::map<std::string, B200Tuning>   tune_table;    // kernel signature -> tuned shape (tuned/sm_100a.tune, autotuner)
  int                                 autotune = 0;  // 0 off, 1 tune operators missing from the table, 2 always
  // scratch for norms
  double *d_scratch = nullptr;
  // copy streams + events of the streamed host-buffer apply (ceedb200_operator_apply_streamed), created on first use
  cudaStream_t             s_h2d = nullptr, s_d2h = nullptr;
  std::vector<cudaEvent_t> ev_stream;
  size_t  scratch_len = 0;
  // grow-only work buffers of the standalone basis kernels (stream-ordered reuse: no allocation, no synchronisation per apply)
  double *d_basis_tmp[3]     = {nullptr, nullptr, nullptr};
  size_t  basis_tmp_bytes[3] = {0, 0, 0};
};

std::string b200_reduced_signature(const std::string &shape_signature);
int b200_error(B200Ceed ceed, int code, const char *fmt, ...) __attribute__((format(printf, 3, 4)));

#define B200_CHECK(cond, ceed, code, ...)                      \
  do {                                                         \
    if (!(cond)) return b200_error(ceed, code, __VA_ARGS__);   \
  } while (0)
#define B200_CALL(...)         \
  do {                         \
    int ierr_ = (__VA_ARGS__); \
    if (ierr_) return ierr_;   \
  } while (0)
#define B200_CUDA(ceed, ...)                                                                                             \
  do {                                                                                                                   \
    cudaError_t cerr_ = (__VA_ARGS__);                                                                                   \
    if (cerr_ != cudaSuccess)                                                                                            \
      return b200_error(ceed, B200_ERROR_BACKEND, "%s:%d CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(cerr_)); \
  } while (0)
#define B200_CU(ceed, ...)                                                                             \
  do {                                                                                                 \
    CUresult cres_ = (__VA_ARGS__);                                                                    \
    if (cres_ != CUDA_SUCCESS) {                                                                       \
      const char *msg_ = nullptr;                                                                      \
      cuGetErrorString(cres_, &msg_);                                                                  \
      return b200_error(ceed, B200_ERROR_BACKEND, "%s:%d CUDA driver error: %s", __FILE__, __LINE__, msg_ ? msg_ : "?"); \
    }                                                                                                  \
  } while (0)

// NVRTC compile (cached).  `defines` are extra -D options.
int b200_jit_compile(B200Ceed ceed, const std::string &source, const std::vector<std::string> &defines, B200Module **module);
int b200_jit_get_kernel(B200Ceed ceed, B200Module *module, const char *name, CUfunction *kernel);
int b200_launch(B200Ceed ceed, CUfunction kernel, unsigned grid, unsigned block, unsigned smem_bytes, void **args, bool cooperative = false);
std::string b200_jit_dir();
bool b200_compile_only();
int  b200_dmalloc(B200Ceed ceed, void **p, size_t bytes);
int  b200_dfree(B200Ceed ceed, void *p);
int  b200_h2d(B200Ceed ceed, void *d, const void *h, size_t bytes);
int  b200_d2h(B200Ceed ceed, void *h, const void *d, size_t bytes);
int  b200_d2d(B200Ceed ceed, void *dst, const void *src, size_t bytes);

// ---------------------------------------------------------------- objects
enum B200Valid { B200_VALID_NONE = 0, B200_VALID_HOST = 1, B200_VALID_DEVICE = 2, B200_VALID_BOTH = 3 };

struct B200Vector_ {
  B200Ceed ceed   = nullptr;
  int64_t  length = 0;
  double  *h_owned = nullptr, *h_borrowed = nullptr;
  double  *d_owned = nullptr, *d_borrowed = nullptr;
  double  *h_array = nullptr, *d_array = nullptr;  // the *valid* pointers (NULL when that side is stale)
};

struct B200Restriction_ {
  B200Ceed ceed = nullptr;
  int      num_elem = 0, elem_size = 0, num_comp = 0;
  int64_t  comp_stride = 0;  // offset restrictions
  int64_t  l_size = 0;
  bool     is_strided = false, backend_strides = false;
  int64_t  strides[3] = {0, 0, 0};  // node, comp, elem strides in the L-vector
  // offsets
  const int32_t *h_offsets = nullptr;        // valid host pointer
  int32_t       *h_offsets_owned = nullptr;  // allocated by us
  const int32_t *h_offsets_borrowed = nullptr;
  int32_t       *d_offsets = nullptr;
  bool           d_offsets_owned = false;
  // transpose CSR over referenced L-nodes, entries ordered by ascending E-index (elem, node)
  //   == summation order of backends/ref/ceed-ref-restriction.c:220-242
  int64_t  num_nodes = 0;  // distinct L-nodes referenced
  int32_t *d_lvec_indices = nullptr, *d_t_offsets = nullptr, *d_t_indices = nullptr;
  // owner/halo decomposition for the fused deterministic scatter (see b200_restriction.cu)
  int32_t *d_tgt = nullptr;         // per E-entry: >=0 L-index to store into (owner), <0: ~slot in halo buffer
  int32_t *d_halo_node = nullptr;   // per shared node: L-index
  int32_t *d_halo_ptr = nullptr;    // per shared node: start slot (CSR), size num_shared+1
  // dense form of the same tables for whole-mesh applies of an unpartitioned restriction (halo slots are then in ascending L-index
  // order): one BYTE per L-index (number of halo slots of that node, 0 for unshared nodes) + the first slot of every block of
  // kDenseBlockNodes L-indices -- 1 byte per node of table traffic in the finalize pass instead of 8 bytes per shared node
  uint8_t *d_halo_cnt = nullptr;
  int32_t *d_halo_block_base = nullptr;
  int64_t  dense_n = 0;             // L-indices covered by d_halo_cnt (0: dense form not available)
  int64_t  num_shared = 0, num_halo = 0;
  // element split of a partitioned mesh (multi-GPU overlap): elements [0, split_elem) touch the rank interface and are applied first.
  // Shared nodes all of whose touchers lie in that range come first in halo_node / halo_ptr (num_shared_first of them), so that
  // the boundary part of the scatter can be finalized -- and sent -- while the interior elements are still being applied.
  int      split_elem = -1;
  int64_t  num_shared_first = 0;
  // Generalisation to K element parts (streamed host-buffer apply, ceedb200_operator_apply_streamed): part p = elements
  // [part_ends[p - 1], part_ends[p]) (part_ends[-1] = 0); shared nodes are ordered by the part of their LAST toucher, shared_prefix[p]
  // = number of shared nodes completed by parts <= p.  The split above is the two-part case {split_elem, num_elem}.
  std::vector<int32_t> part_ends;
  std::vector<int64_t> shared_prefix;
  bool     transpose_built = false, owner_built = false;
};

struct B200Basis_ {
  B200Ceed ceed = nullptr;
  int      dim = 0, num_comp = 0, P = 0, Q = 0;
  std::vector<double> interp, grad, q_ref, q_weight, collo_grad;  // host copies (1-D)
  bool     has_collo_grad = false;  // Q >= P
  bool     is_collocated  = false;  // interp_1d == identity
  bool     is_tensor      = true;   // false: CeedBasisCreateH1 -- P nodes, Q points per element, interp [Q x P], grad [dim][Q x P]
  double  *d_interp = nullptr, *d_grad = nullptr, *d_q_weight = nullptr, *d_collo_grad = nullptr;
};

struct B200QFContext_ {
  B200Ceed ceed = nullptr;
  size_t   size = 0;
  void    *h_owned = nullptr, *h_borrowed = nullptr, *d_owned = nullptr, *d_borrowed = nullptr;
  void    *h_data = nullptr, *d_data = nullptr;  // valid pointers
};

struct B200QFField {
  std::string name;
  int         size = 0;
  int         eval_mode = 0;
};

struct B200QFunction_ {
  B200Ceed                 ceed = nullptr;
  std::string              source_path, kernel_name;
  std::vector<B200QFField> inputs, outputs;
  B200QFContext            ctx = nullptr;
  void                    *raw_ctx = nullptr;  // device pointer of a context owned by someone else (used when ctx == nullptr)
  B200Module              *module = nullptr;  // standalone apply kernel
  CUfunction               kernel = nullptr;
};

struct B200OpField {
  B200Restriction rstr = nullptr;
  B200Basis       basis = nullptr;
  B200Vector      vec = nullptr;
  bool            is_active = false, is_set = false;
};

// Tables of the ORDERED scatter mode (in-kernel completion of shared nodes, b200_restriction_build_ordered): like the
// owner/halo tables, but the LAST E-entry of a shared node is the direct one -- its element group waits for the groups of
// the earlier entries (flags) and adds their halo values in ascending E-order before storing.
struct B200OrderedScatter {
  int32_t *d_tgt = nullptr;       // per E-entry: >=0 plain L-index; <0: ~(slot | last << 27 | (touchers - 2) << 28)
  int32_t *d_node = nullptr;      // per shared node: L-index      } written once into the halo buffer (id slots),
  int32_t *d_ptr = nullptr;       // per shared node: its id slot  } see b200_ordered_init_halo
  bool     supported = true;      // false: too many slots / touchers for the 31-bit encoding -> two-pass scheme
  int32_t *d_pred_ptr = nullptr;  // per element group: predecessor groups (CSR, num_groups + 1)
  int32_t *d_pred_idx = nullptr;
  int32_t *d_flags = nullptr;     // per element group: epoch of the last launch whose halo values are complete
  int32_t *d_sync = nullptr;      // {epoch of the last finished launch, CTAs finished in the running launch}
  int64_t  num_shared = 0, num_halo = 0, num_groups = 0, num_pred = 0;
};
// Tables of the RUN scatter (deterministic, b200_restriction_build_runs): the elements of a launch are cut into one contiguous run
// per element group (warp); a group walks its run in nb = ceil(len / E) iterations, iteration i holding the E elements
// s + i + k * nb (k < E).  An E-entry whose L-node was first touched by the SAME group, with every earlier toucher processed in an
// earlier iteration of that group, is added straight into v (read-modify-write in program order: still the ascending E-order of
// the serial reference); only the other non-owner entries go through the halo buffer and the finalize pass.
//   tgt >= 0: L-index, bit 30 set = read-modify-write (clear = plain store of the owner entry);  tgt < 0: ~halo slot
struct B200RunScatter {
  int32_t *d_tgt = nullptr, *d_halo_node = nullptr, *d_halo_ptr = nullptr;
  int64_t  num_shared = 0, num_halo = 0, num_rmw = 0;
  int      num_groups = 0, group_elems = 0;  // key: the launch shape the tables were built for
  bool     built = false;
};
constexpr int32_t B200_RUN_RMW_BIT = 1 << 30;
int  b200_restriction_build_runs(B200Restriction r, int num_groups, int group_elems, B200RunScatter *out);
void b200_run_scatter_free(B200Ceed ceed, B200RunScatter *t);
int  b200_halo_finalize_lists(B200Restriction r, const int32_t *d_halo_node, const int32_t *d_halo_ptr, int64_t num_shared, int64_t num_halo, const double *d_halo,
                              double *d_v);
int  b200_restriction_build_ordered(B200Restriction r, int group_elems, B200OrderedScatter *out);
int  b200_ordered_init_halo(B200Restriction r, const B200OrderedScatter *t, double *d_halo);
void b200_ordered_scatter_free(B200Ceed ceed, B200OrderedScatter *t);

struct B200OpPlan;  // generated-kernel plan (b200_opgen.cpp)

struct B200Operator_ {
  B200Ceed                 ceed = nullptr;
  B200QFunction            qf = nullptr;
  std::vector<B200OpField> in_fields, out_fields;
  bool                     is_setup = false;
  B200OpPlan              *plan = nullptr;
  B200Tuning               tune;           // explicit overrides (ceedb200_operator_set_tuning, autotuner)
  bool                     tuned = false;  // autotuner has run (or was not applicable)
  bool                     no_tma = false; // an input of a bulk-copied field was not 16-byte aligned: generate without cp.async.bulk
  bool                     no_ladder = false; // autotuner trials: a candidate shape that cannot be built is an error, not a reason to fall back
  int                      build_rung = 0; // fallback ladder of the fused kernel: 0 tuned shape, 1 conservative shape, 2 unfused kernels
  bool                     timing = false;
  float                    last_fused_ms = 0.f, last_aux_ms = 0.f;
  cudaEvent_t              ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

// internal helpers shared between translation units
int b200_vector_device_read(B200Vector vec, const double **d);
int b200_vector_device_write(B200Vector vec, double **d, bool discard);
int b200_restriction_build_transpose(B200Restriction rstr);
int b200_restriction_build_owner(B200Restriction rstr);
int b200_restriction_e_size(B200Restriction rstr, int64_t *e_size);
// raw device-pointer restriction kernels (used by the unfused operator path and EVECTOR scatter mode)
int b200_restriction_apply_raw(B200Restriction rstr, int t_mode, const double *d_u, double *d_v);
// streamed apply: device side of `vec` handed out WITHOUT the host-to-device copy (the caller copies chunk by chunk); both sides valid after
int b200_vector_streamed_input(B200Vector vec, const double **h, double **d);
int b200_vector_streamed_output(B200Vector vec, double **h);  // host array the chunks of the result are copied into; valid on both sides after
int b200_halo_finalize(B200Restriction rstr, const double *d_halo, double *d_v, int part = 0);  // part 0 all; p >= 1: the shared nodes completed by element part p
int b200_restriction_set_parts(B200Restriction rstr, const std::vector<int32_t> &part_ends);     // K element parts (rebuilds the owner/halo tables if they change)
int b200_memset_async(B200Ceed ceed, void *d, size_t bytes);
