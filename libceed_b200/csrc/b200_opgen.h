// b200_opgen.h -- plan of one fused operator kernel (see b200_opgen.cpp)
#pragma once
#include <string>
#include <vector>

#include "b200_internal.h"

struct B200GenBasis {
  B200Basis basis;
  int       P, Q;
  bool      collocated;
};

// A group = all QFunction fields that share (restriction, vector, basis): gathered / scattered once.
struct B200GenGroup {
  bool            is_input   = true;
  B200Restriction rstr       = nullptr;
  B200Vector      vec        = nullptr;
  bool            is_active  = false;
  int             basis_id   = -1;
  int             nc         = 1;
  bool            use_interp = false, use_grad = false;
  int             plane0     = 0;  // first shared-memory plane
  int             slot       = 0;  // index into the kernel argument pointer tables
  // asynchronous staging buffers in shared memory (byte offsets from the dynamic smem base, -1 = not staged)
  int             uin_off = -1;    // inputs: gathered L-vector values of the NEXT batch   [nc][E][P^3] doubles
  int             idx_off = -1;    // inputs: element offsets of the next batch             [E][P^3] ints
  int             tgt_off = -1;    // outputs: scatter targets of the current batch         [E][P^3] ints
};

struct B200GenField {
  int             emode = 0, size = 0, nc = 0;
  int             group = -1;     // INTERP/GRAD fields
  int             basis_id = -1;  // WEIGHT fields
  B200Restriction rstr = nullptr; // NONE fields
  B200Vector      vec = nullptr;
  bool            is_active = false;
  int             slot = 0;
  int             qd_off = -1;    // EVAL_NONE inputs streamed with cp.async: [nc][E * Q^3] doubles (byte offset), -1 = direct loads
  int             ring_k = -1;    // EVAL_NONE inputs prefetched through the per-lane ring: first component row in a ring slot
  bool            qd_tma = false; // qd_off buffer is filled by cp.async.bulk (1-D TMA) + mbarrier instead of per-lane cp.async
  int             qd_cs  = 0;     // doubles between the components of the qd_off buffer (E * Q^3, padded for the 16-byte source alignment)
};

struct B200OpArgs {
  long long     num_elem;
  void         *ctx;
  const double *in_ptr[16];
  double       *out_ptr[16];
  const int    *in_idx[16];
  const int    *out_idx[16];
  double       *out_aux[16];  // halo buffer (deterministic) or E-vector (evector mode)
  // ordered scatter (B200_SCATTER_ORDERED): tables of the one offset-restricted output
  const int *ord_pred_ptr, *ord_pred_idx;
  int       *ord_flags, *ord_sync;
  long long  ord_num_halo;
};

struct B200KernelVariant {
  B200Module *module = nullptr;
  CUfunction  kernel = nullptr;
  std::string source;
  int         regs = 0, local_bytes = 0, static_smem = 0;
  int         blocks_per_sm = 1;  // resident CTAs per SM of THIS variant (occupancy query); the persistent grid is sized from it
  bool        built = false;
};

// chunk tables of the streamed host-buffer apply (ceedb200_operator_apply_streamed), cached per operator
struct B200StreamPlan {
  int                  num_chunks = 0;
  std::vector<int32_t> ends;             // element chunk ends
  std::vector<int64_t> in_hi, out_done;  // per chunk: (highest offset gathered by chunks <= c) + 1; offsets below out_done[c] are complete after chunk c
  bool                 per_comp = false; // components are blocks of comp_stride (each streams its own range); else one range over L-indices
};

struct B200OpPlan {
  bool                      fused = false;
  std::string               why_not_fused;
  int                       dim = 0, Q = 0, Qs = 0;
  int                       num_elem = 0;
  int                       epb = 1, threads = 256, blocks_per_sm = 1, grid = 1;
  int                       plane_size = 0, num_planes = 0, smem_bytes = 0;
  int                       scatter_mode = 0;
  bool                      warp_mode = true;    // one element group per warp, __syncwarp() only (see b200_opgen.cpp)
  int                       stage_mask = 1;      // which global reads are staged one batch ahead: cp.async 1 idx/tgt, 2 gather, 4 qdata, 8 idx only,
                                                 // 16 per-lane qdata ring; 32 qdata through cp.async.bulk (TMA) + mbarrier
  bool                      swz = false;         // conflict-free swizzled plane layout (stage bit 256), see b200_opgen_plan
  int                       swz_w = 0;           // its row width: 8 (Q, P <= 8) or 16
  bool                      lin = false;         // even-Q linear layout (stage bit 512): unpadded rows, z-stride = Q (mod 16), 16-byte x-line accesses
  bool                      qd_tma = false;      // some EVAL_NONE input is staged by bulk copies
  int                       mbar_off = -1;       // byte offset of the group's mbarrier (bulk-copy completion)
  bool                      no_tma = false;      // set by the host when an input pointer is not 16-byte aligned
  int                       group_smem_bytes = 0;  // shared memory of one element group (CTA in block mode, warp in warp mode)
  int                       group_warps = 1;      // warps sharing one element group (warp mode)
  bool                      qf_pointwise = false; // QFunction over independent points, d/dz as separate line stages
  bool                      async_copy = true;  // stage global reads through cp.async one batch ahead
  // z-line QFunction stage: per-lane cp.async ring for the streamed quadrature data.  A step = one z-layer of one task round;
  // ring_slots - 1 steps are always in flight, ACROSS the other stages and across element groups.
  int                       ring_off = -1, ring_slots = 0, ring_comps = 0, ring_rounds = 0;
  bool                      qf_xline = false;    // gradient-free operators: QFunction on x-lines inside the x-contraction stage
  bool                      lean = false;        // gradient-free operators on the lean in-place-plane kernel (b200_opgen_lean.cpp, QFunction layout 4)
  bool                      lean_runs = false;   // lean kernel generated with run-scatter support (experimental, CEED_B200_RUNS)
  int                       lean_es = 0;         // its element stride in shared memory (doubles)
  int                       lean_tg_off = 0;     // byte offset of the parked (or bulk-copied) scatter targets in a warp's shared-memory slice
  int                       lean_ip = 0;         // stage bit 1 (element-interleaved columns): element pitch (ints) of the staged offset / target tables
  int                       lean_off_off = -1;   // bulk pipeline (stage bit 32): byte offset of the element offsets of the batch
  int                       qf_pp = 1;           // pointwise QFunction stage: points per lane (2 = x-adjacent pair, 16-byte loads)
  int                       qf_ahead = 1;        // z-line QFunction stage: z-layers of streamed inputs in flight per lane
  int                       qf_unroll = 4;       // pointwise QFunction stage: points in flight per lane
  std::string               signature, shape_signature;  // tuning-table keys (with / without the QFunction name)
  B200Tuning                resolved;            // the shape actually used
  std::vector<B200GenBasis> bases;
  std::vector<B200GenGroup> in_groups, out_groups;
  std::vector<B200GenField> in_fields, out_fields;
  B200KernelVariant         variant[2];  // [0] overwrite (Apply), [1] add (ApplyAdd)
  // per output slot: auxiliary device buffer (halo or E-vector)
  double *aux[16]       = {nullptr};
  size_t  aux_bytes[16] = {0};
  B200OrderedScatter ordered;     // ordered scatter tables (scatter_mode == B200_SCATTER_ORDERED)
  B200StreamPlan     stream;      // streamed host-buffer apply
  // in-kernel finalize of the lean kernel (stage bit 128): device part table {K, batch starts[K+1], shared prefix[K+1]} and per-part counters
  int32_t           *d_fin_tab = nullptr, *d_fin_done = nullptr;
  int                fin_parts = 0;
  std::vector<int32_t> fin_tab_host;
  B200RunScatter     run[2][16];  // run scatter tables of the lean kernel, per kernel variant (their grids may differ) and output slot
  int                ordered_slot = -1;
  // unfused fallback scratch
  std::vector<B200Vector> e_in, q_in, e_out, q_out;
};

int b200_opgen_plan(B200Operator op, B200OpPlan *plan);
int b200_opgen_build(B200Operator op, B200OpPlan *plan, int add);
std::string b200_opgen_source(B200Operator op, B200OpPlan *plan, int add);
// lean kernel of gradient-free operators (b200_opgen_lean.cpp)
bool        b200_opgen_lean_eligible(const B200OpPlan *plan);
size_t      b200_opgen_lean_layout(B200OpPlan *plan, int E);
std::string b200_opgen_lean_source(B200Operator op, B200OpPlan *plan, int add);
int b200_opgen_grid(B200Ceed ceed, const B200OpPlan *plan, const B200KernelVariant &v, long long num_elem);
