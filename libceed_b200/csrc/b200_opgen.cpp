// b200_opgen.cpp -- generator of the fused CeedOperatorApply kernel for sm_100a (3-D tensor-product H1 operators).
//
// What it replaces: the reference's code generator backends/cuda-gen/ceed-cuda-gen-operator-build.cpp:1158-1681 and the
// device templates it strings together (include/ceed/jit-source/cuda/cuda-gen-templates.h:324-533,
// cuda-shared-basis-tensor-templates.h:321-654).  Same contract (one kernel per operator, user QFunction inlined through
// NVRTC with Q = 1 per call), different design:
//
//   * LINE-OWNED CONTRACTIONS.  A block processes a batch of E elements.  Every 1-D contraction stage assigns one whole
//     tensor line (P inputs -> Q outputs) to a thread: inputs are P shared-memory loads at immediate offsets, the P*Q FMAs
//     take their basis coefficient straight from the constant bank (matrices live in __constant__ memory and every loop
//     is fully unrolled, so the coefficient is an instruction operand: no register, no load).  That gives P*Q/(P+Q)
//     FMAs per shared-memory access (3.7 at p=6) where the reference's thread-per-(x,y)-column scheme needs one LDS per
//     DFMA plus 2 block barriers per z-layer per contraction (~105 barriers/element at p=6);
//     here a batch of E elements costs 9 barriers in total.
//   * Stages are fused where a thread already owns the data: x-interp + d/dx, d/dz + QFunction + (d/dz)^T in one
//     z-line pass, (d/dx)^T + x-interp^T.
//   * Gathers/scatters read element offsets with consecutive lanes on consecutive entries (coalesced 4-byte index
//     loads); quadrature data is streamed straight from HBM into registers at the quadrature point that consumes it.
//   * DETERMINISTIC SCATTER WITHOUT ATOMICS: the first E-entry referencing an L-node owns the store (plain st.global,
//     which also gives overwrite semantics without a memset + read-modify-write pass); other entries write a compact
//     halo buffer that a small finalize kernel folds in, in ascending (elem,node) order -- the order of the serial CPU
//     reference (backends/ref/ceed-ref-restriction.c:220-242).  Atomic and E-vector modes exist for comparison.
//   * shared-memory planes are padded to odd line strides (conflict-free line-strided access), sized so that several
//     blocks are resident per SM; grid = resident blocks x 148 SMs, grid-stride over element batches.
#include "b200_opgen.h"

#include <cstdlib>
#include <cstring>
#include <sstream>

using std::string;
typedef std::ostringstream oss;

namespace {

int odd_pad(int n) { return (n % 2) ? n : n + 1; }
// smallest m >= n with m = 8 (mod 16): the stride between z-layers of the conflict-free ("swizzled") shared-memory layouts
int pad8(int n) { return n <= 8 ? 8 : ((n - 8 + 15) / 16) * 16 + 8; }
// Strides of the swizzled layouts for row width w (8: Q, P <= 8; 16: larger).  With w = 16 a half-warp is exactly one row, so the
// z-strides need no residue and only the quadrature planes widen (pitch 16).
int swz_sz(int w, int n) { return w == 8 ? pad8(n) : n; }
// z-stride of the even-Q linear layout: smallest m >= Q^2 with m = Q (mod 16) -- y- and z-line stages then map consecutive lanes to
// consecutive banks (lane (qx, qz) -> qx + Q qz (mod 16)), x-lines move as 16-byte accesses (2-way conflicts instead of clean, but the
// y- and z-line stages carry 2.5x the traffic): scripts/model/lin_layout.py
int lin_sz(int Q) {
  int m = Q * Q;
  while (m % 16 != Q % 16) m++;
  return m;
}

string hexd(double v) {
  char buf[64];
  snprintf(buf, sizeof(buf), "%a", v);
  return buf;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ planning
int b200_opgen_plan(B200Operator op, B200OpPlan *plan) {
  B200Ceed      ceed = op->ceed;
  B200QFunction qf   = op->qf;
  plan->fused        = false;
  plan->no_tma       = op->no_tma || getenv("CEED_B200_NO_TMA") != nullptr;
  plan->scatter_mode = ceed->scatter_mode;
  plan->ordered_slot = -1;
  auto reject        = [&](const string &why) {
    plan->why_not_fused = why;
    return B200_SUCCESS;
  };
  int dim = 0, Q = 0, num_elem = -1;
  // bases
  auto basis_id = [&](B200Basis b) {
    for (size_t i = 0; i < plan->bases.size(); i++)
      if (plan->bases[i].basis == b) return (int)i;
    plan->bases.push_back({b, b->P, b->Q, b->is_collocated});
    return (int)plan->bases.size() - 1;
  };
  auto find_group = [&](std::vector<B200GenGroup> &groups, bool is_input, const B200OpField &f, int bid, int slot) {
    for (size_t i = 0; i < groups.size(); i++) {
      if (groups[i].rstr == f.rstr && groups[i].basis_id == bid && groups[i].is_active == f.is_active && (f.is_active || groups[i].vec == f.vec))
        return (int)i;
    }
    B200GenGroup g;
    g.is_input  = is_input;
    g.rstr      = f.rstr;
    g.vec       = f.vec;
    g.is_active = f.is_active;
    g.basis_id  = bid;
    g.nc        = f.rstr->num_comp;
    g.slot      = slot;
    groups.push_back(g);
    return (int)groups.size() - 1;
  };
  for (int io = 0; io < 2; io++) {
    const auto &qfields  = io == 0 ? qf->inputs : qf->outputs;
    const auto &opfields = io == 0 ? op->in_fields : op->out_fields;
    auto       &gfields  = io == 0 ? plan->in_fields : plan->out_fields;
    auto       &groups   = io == 0 ? plan->in_groups : plan->out_groups;
    for (size_t i = 0; i < qfields.size(); i++) {
      const B200OpField &f = opfields[i];
      B200GenField       g;
      g.emode     = qfields[i].eval_mode;
      g.size      = qfields[i].size;
      g.slot      = (int)i;
      g.vec       = f.vec;
      g.is_active = f.is_active;
      g.rstr      = f.rstr;
      if (f.basis) {
        if (!f.basis->is_tensor) return reject("non-tensor basis");
        if (dim == 0) {
          dim = f.basis->dim;
          Q   = f.basis->Q;
        }
        if (f.basis->dim != dim || f.basis->Q != Q) return reject("bases with different dim or Q_1d");
      }
      if (f.rstr) {
        if (num_elem < 0) num_elem = f.rstr->num_elem;
        if (f.rstr->num_elem != num_elem) return reject("restrictions with different numbers of elements");
      }
      switch (g.emode) {
        case B200_EVAL_NONE:
          if (!f.rstr) return reject("EVAL_NONE field without restriction");
          g.nc = f.rstr->num_comp;
          if (g.nc != g.size) return reject("EVAL_NONE field size != restriction components");
          break;
        case B200_EVAL_WEIGHT:
          if (io == 1 || !f.basis) return reject("EVAL_WEIGHT needs an input basis");
          g.basis_id = basis_id(f.basis);
          g.nc       = 1;
          break;
        case B200_EVAL_INTERP:
        case B200_EVAL_GRAD: {
          if (!f.basis || !f.rstr) return reject("INTERP/GRAD field without basis or restriction");
          if (g.emode == B200_EVAL_GRAD && !f.basis->has_collo_grad) return reject("GRAD with Q < P (no collocated gradient)");
          const int bid = basis_id(f.basis);
          g.group       = find_group(groups, io == 0, f, bid, (int)i);
          g.nc          = f.rstr->num_comp;
          if (g.emode == B200_EVAL_INTERP) groups[g.group].use_interp = true;
          else groups[g.group].use_grad = true;
          if (g.size != g.nc * (g.emode == B200_EVAL_GRAD ? dim : 1)) return reject("field size does not match components x q_comp");
          int nn = 1;
          for (int d = 0; d < dim; d++) nn *= f.basis->P;
          if (f.rstr->elem_size != nn) return reject("restriction element size != P^dim");
        } break;
        default: return reject("unsupported eval mode");
      }
      gfields.push_back(g);
    }
  }
  if (dim != 3) return reject("fused kernel is generated for dim == 3 only");
  if (num_elem < 0) return reject("no restriction");
  if (plan->out_groups.empty()) {
    bool any = false;
    for (auto &f : plan->out_fields) any = any || f.emode == B200_EVAL_NONE;
    if (!any) return reject("no output");
  }
  if (plan->in_groups.empty() && plan->out_groups.empty()) return reject("restriction-only operator (no basis action)");
  // EVAL_NONE fields must be sized Q^3 per element
  for (int io = 0; io < 2; io++)
    for (auto &f : (io ? plan->out_fields : plan->in_fields))
      if (f.emode == B200_EVAL_NONE && f.rstr->elem_size != Q * Q * Q) return reject("EVAL_NONE restriction element size != Q^dim");
  for (auto &b : plan->bases)
    if (b.P > 16 || b.Q > 16) return reject("P or Q > 16");

  if (plan->scatter_mode == B200_SCATTER_ORDERED) {
    // in-kernel ordered completion: one offset-restricted output group, no other offset-restricted outputs
    int n_offset = 0;
    for (auto &g : plan->out_groups)
      if (!g.rstr->is_strided) n_offset++, plan->ordered_slot = g.slot;
    for (auto &f : plan->out_fields)
      if (f.emode == B200_EVAL_NONE && !f.rstr->is_strided) n_offset += 2;
    if (n_offset != 1 || getenv("CEED_B200_BLOCK_MODE")) {
      plan->scatter_mode = B200_SCATTER_DETERMINISTIC;  // same results, two-pass scheme
      plan->ordered_slot = -1;
    }
  }
  plan->dim      = dim;
  plan->Q        = Q;
  plan->Qs       = odd_pad(Q);
  plan->num_elem = num_elem;
  // Kernel shape: explicit override (set_tuning / autotuner) > environment > tuning table > heuristic.
  {
    std::ostringstream sig;
    sig << "Q" << Q << "|sc" << plan->scatter_mode;
    auto group_sig = [&](const B200GenGroup &g) {
      const B200GenBasis &b = plan->bases[g.basis_id];
      sig << "P" << b.P << (b.collocated ? "c" : "") << "n" << g.nc << (g.use_interp ? "i" : "") << (g.use_grad ? "g" : "") << (g.rstr->is_strided ? "s" : "o")
          << (g.rstr->comp_stride == 1 && g.nc > 1 ? "l" : "");
    };
    sig << "|in:";
    for (auto &g : plan->in_groups) group_sig(g);
    for (auto &f : plan->in_fields)
      if (f.emode == B200_EVAL_NONE) sig << "N" << f.nc;
      else if (f.emode == B200_EVAL_WEIGHT) sig << "W";
    sig << "|out:";
    for (auto &g : plan->out_groups) group_sig(g);
    for (auto &f : plan->out_fields)
      if (f.emode == B200_EVAL_NONE) sig << "N" << f.nc;
    plan->shape_signature = sig.str();
    plan->signature       = plan->shape_signature + "|" + op->qf->kernel_name;
  }
  B200Tuning tn = op->tune;
  {
    B200Tuning table;
    auto       it = ceed->tune_table.find(plan->signature);
    if (it == ceed->tune_table.end()) it = ceed->tune_table.find(plan->shape_signature);
    if (it == ceed->tune_table.end()) it = ceed->tune_table.find(b200_reduced_signature(plan->shape_signature));
    if (it == ceed->tune_table.end() && plan->scatter_mode != B200_SCATTER_DETERMINISTIC) {
      // no entry for this scatter mode: the shape tuned for the default mode is the best guess
      auto as_default = [&](std::string key) {
        const size_t pos = key.find("|sc");
        if (pos != std::string::npos) key[pos + 3] = '0';
        return key;
      };
      it = ceed->tune_table.find(as_default(plan->signature));
      if (it == ceed->tune_table.end()) it = ceed->tune_table.find(as_default(plan->shape_signature));
      if (it == ceed->tune_table.end()) it = ceed->tune_table.find(b200_reduced_signature(as_default(plan->shape_signature)));
    }
    if (it != ceed->tune_table.end()) table = it->second;
    auto pick = [&](int &field, int unset, const char *env, int table_value) {
      if (field != unset) return;
      if (getenv(env)) field = atoi(getenv(env));
      else field = table_value;
    };
    pick(tn.epw, 0, "CEED_B200_EPB", table.epw);
    pick(tn.group_warps, 0, "CEED_B200_GROUP_WARPS", table.group_warps);
    pick(tn.cta_warps, 0, "CEED_B200_WARPS", table.cta_warps);
    pick(tn.minb, 0, "CEED_B200_MINB", table.minb);
    pick(tn.qf_mode, -1, "CEED_B200_QF_POINTWISE", table.qf_mode);
    pick(tn.qf_unroll, 0, "CEED_B200_QF_UNROLL", table.qf_unroll);
    pick(tn.stage, -1, "CEED_B200_STAGE", table.stage);
  }
  // QFunction stage layout: z-line (default) = d/dz, QFunction and (d/dz)^T fused per z-line in registers;
  // pointwise = d/dz and its transpose are separate line stages through shared memory and the QFunction runs over
  // independent points (short dependency chains on the streamed quadrature data, one more plane).
  // x-line = operators without gradients (mass-like: BP1/BP2): the QFunction runs on whole x-lines INSIDE the x-contraction
  // stage (X, QFunction with Q points, X^T in registers) -- no quadrature-point plane round trip and no separate stage.
  {
    bool no_grad = !plan->in_groups.empty() && !plan->out_groups.empty();
    for (auto &g : plan->in_groups) no_grad = no_grad && !g.use_grad && !plan->bases[g.basis_id].collocated;
    for (auto &g : plan->out_groups) no_grad = no_grad && !g.use_grad && !plan->bases[g.basis_id].collocated;
    for (auto &f : plan->in_fields) no_grad = no_grad && !(f.emode == B200_EVAL_NONE && !f.rstr->is_strided);
    for (auto &f : plan->out_fields) no_grad = no_grad && f.emode != B200_EVAL_NONE;
    if ((tn.qf_mode == 3 || tn.qf_mode == 4) && !no_grad) tn.qf_mode = 0;
    if (tn.qf_mode < 0 && no_grad && !getenv("CEED_B200_NO_XLINE")) tn.qf_mode = 3;
    // layout 4: the lean in-place-plane kernel (b200_opgen_lean.cpp); operators it does not cover run the general x-line layout
    if (tn.qf_mode == 4 && (!b200_opgen_lean_eligible(plan) || getenv("CEED_B200_BLOCK_MODE"))) tn.qf_mode = 3;
    plan->lean     = tn.qf_mode == 4;
    plan->lean_runs = plan->lean && getenv("CEED_B200_RUNS") != nullptr && plan->scatter_mode == B200_SCATTER_DETERMINISTIC;
    plan->qf_xline = tn.qf_mode == 3 || plan->lean;
  }
  // Shared-memory layout of the contraction planes.  Default: rows padded to an odd pitch.  Swizzled (stage bit 256, Q <= 8,
  // z-line / x-line QFunction stage): quadrature rows have pitch 8 with the x index XOR-ed by the row's y index, every
  // z-stride is 8 (mod 16) and every stage enumerates its lanes with a fastest extent of 8 -- with that, all stages
  // (x-, y- and z-lines) are free of shared-memory bank conflicts (scripts/model/swizzle_model.py), where the padded layout replays
  // a third of its wavefronts (ncu, profiles/r01_ncu_full_summary.txt).
  {
    const int stage_req = tn.stage >= 0 ? tn.stage : 1;
    bool      ok        = (stage_req & 256) && Q <= 16 && !getenv("CEED_B200_BLOCK_MODE") && !(tn.qf_mode == 1 || tn.qf_mode == 2) && !(stage_req & 16) &&
              plan->scatter_mode != B200_SCATTER_ORDERED;
    int w = Q <= 8 ? 8 : 16;
    for (auto &b : plan->bases) {
      ok = ok && b.P <= 16;
      if (b.P > 8) w = 16;
    }
    plan->swz   = ok;
    plan->swz_w = ok ? w : 0;
    plan->lin   = !ok && (stage_req & 512) && Q % 2 == 0 && !getenv("CEED_B200_BLOCK_MODE") && tn.qf_mode != 1 && tn.qf_mode != 2 && !(stage_req & 16) &&
                plan->scatter_mode != B200_SCATTER_ORDERED;
  }
  {
    const int w          = plan->swz_w;
    int       plane_size = plan->swz ? Q * swz_sz(w, Q * w) : (plan->lin ? Q * lin_sz(Q) : Q * Q * plan->Qs);
    for (auto &b : plan->bases) {
      const int wp = b.P <= 8 ? 8 : 16;  // lane width of the node-indexed (y) stages of this basis
      plane_size   = std::max(plane_size, Q * (plan->swz ? swz_sz(wp, b.P * b.P) : b.P * b.P));
      plane_size   = std::max(plane_size, Q * (plan->swz ? swz_sz(wp, Q * odd_pad(b.P)) : Q * odd_pad(b.P)));
    }
    if (plan->swz) plane_size = (plane_size + 15) / 16 * 16;  // element / plane strides keep the bank residues
    if (plan->lin) plane_size = (plane_size + 1) / 2 * 2;       // rows stay 16-byte aligned
    plan->plane_size = plane_size;
  }
  plan->qf_pointwise = tn.qf_mode == 1 || tn.qf_mode == 2;
  plan->qf_pp        = (tn.qf_mode == 2 && Q % 2 == 0) ? 2 : 1;  // point pairs need x-adjacent points in one row
  plan->qf_unroll    = tn.qf_unroll > 0 ? tn.qf_unroll : 4;
  plan->qf_ahead     = getenv("CEED_B200_QF_AHEAD") ? atoi(getenv("CEED_B200_QF_AHEAD")) : 1;
  auto planes_of = [&](const B200GenGroup &g) { return g.nc * (g.use_grad ? ((plan->qf_pointwise && g.use_interp) ? 4 : 3) : 2); };
  int  n_in = 0, n_out = 0;
  for (auto &g : plan->in_groups) {
    g.plane0 = n_in;
    n_in += planes_of(g);
  }
  for (auto &g : plan->out_groups) {
    g.plane0 = n_out;
    n_out += planes_of(g);
  }
  plan->num_planes = std::max(n_in, n_out);
  // elements per block / threads: one thread per quadrature line (Q^2 lines per element), ~256 threads per block
  const int lines = Q * Q;
  // Shared-memory layout for E elements per block: contraction planes, then the staging buffers that cp.async fills one
  // batch ahead (quadrature data, gathered inputs + their offsets, scatter targets).  Returns the total in bytes.
  plan->async_copy = !getenv("CEED_B200_NO_ASYNC");
  auto layout      = [&](int E) {
    size_t off  = ((size_t)plan->num_planes * plan->plane_size * 8 * E + 15) / 16 * 16;
    const int mask = plan->warp_mode ? plan->stage_mask : 7;
    auto   take = [&](size_t bytes) {
      size_t at = off;
      off += (bytes + 15) / 16 * 16;
      return (int)at;
    };
    plan->qd_tma   = false;
    plan->mbar_off = -1;
    for (auto &f : plan->in_fields) {
      f.qd_off = -1;
      f.qd_tma = false;
      f.qd_cs  = E * Q * Q * Q;
      // contiguous per component over the elements of a batch: unit node stride, element stride == element size
      const bool contiguous = f.emode == B200_EVAL_NONE && f.rstr->is_strided && f.rstr->strides[0] == 1 && f.rstr->strides[2] == f.rstr->elem_size;
      if (plan->async_copy && (mask & 32) && contiguous && !plan->no_tma) {
        // bulk copies need 16-byte aligned sources: an odd Q^3 makes every other (component, batch) start 8 bytes off, so the
        // copy starts one double early and the buffer of a component holds one double more (readers add the same shift)
        f.qd_tma = plan->qd_tma = true;
        f.qd_cs  = (E * Q * Q * Q + 1 + 1) / 2 * 2;
        f.qd_off = take((size_t)f.nc * f.qd_cs * 8);
      } else if (plan->async_copy && (mask & 4) && contiguous) {
        f.qd_off = take((size_t)f.nc * E * Q * Q * Q * 8);
      }
    }
    if (plan->qd_tma) plan->mbar_off = take(16);
    for (auto &g : plan->in_groups) {
      g.uin_off = g.idx_off = -1;
      if (plan->async_copy && (mask & 2) && !g.rstr->is_strided) {
        g.uin_off = take((size_t)g.nc * E * g.rstr->elem_size * 8);
        g.idx_off = take((size_t)E * g.rstr->elem_size * 4);
      } else if (plan->async_copy && (mask & 8) && !g.rstr->is_strided) {
        g.idx_off = take((size_t)E * g.rstr->elem_size * 4);  // offsets only: the gather itself stays a direct load
      }
    }
    for (auto &g : plan->out_groups) {
      g.tgt_off = -1;
      if (plan->async_copy && (mask & 1) && !g.rstr->is_strided && plan->scatter_mode != B200_SCATTER_EVECTOR) g.tgt_off = take((size_t)E * g.rstr->elem_size * 4);
    }
    plan->ring_off = -1;
    for (auto &f : plan->in_fields) f.ring_k = -1;
    if (plan->async_copy && (mask & 16) && plan->warp_mode && !plan->qf_pointwise && !plan->qf_xline) {
      int comps = 0;
      for (auto &f : plan->in_fields)
        if (f.emode == B200_EVAL_NONE && f.qd_off < 0 && f.rstr->is_strided) {
          f.ring_k = comps;
          comps += f.nc;
        }
      const int lanes  = 32 * plan->group_warps;
      const int rounds = (E * Q * Q + lanes - 1) / lanes, steps = rounds * Q;
      const int want   = getenv("CEED_B200_RING") ? atoi(getenv("CEED_B200_RING")) : 4;
      int       ns     = 0;
      for (int d = 2; d <= 8 && d <= steps; d++)  // slots must divide the steps of a batch so that slot numbers are compile-time constants
        if (steps % d == 0 && (ns == 0 || abs(d - want) < abs(ns - want))) ns = d;
      if (comps > 0 && ns >= 2 && steps <= 64) {
        plan->ring_slots  = ns;
        plan->ring_comps  = comps;
        plan->ring_rounds = rounds;
        plan->ring_off    = take((size_t)ns * lanes * comps * 8);
      } else {
        for (auto &f : plan->in_fields) f.ring_k = -1;
      }
    }
    return off;
  };
  plan->warp_mode = !getenv("CEED_B200_BLOCK_MODE");
  int epb, threads;
  if (plan->lean) {
    // one warp per element group, nothing staged by cp.async; E elements per warp and iteration (default 8)
    // stage bit 32: offsets, scatter targets and quadrature data through cp.async.bulk + mbarrier, one batch ahead
    // (bit 8: element offsets + scatter targets, bit 32: quadrature data)
    // bit 64: bulk L2 prefetch of the next batch's quadrature data
    // bit 128: in-kernel finalize of the deterministic scatter
    // bit 1: element-interleaved node columns in the gather / scatter stages; bit 2: element stride of the planes = P (mod 16);
    plan->stage_mask = (tn.stage >= 0 && plan->async_copy && !plan->no_tma) ? (tn.stage & (8 | 32 | 64 | 128)) : 0;
    // bit 4: directly loaded quadrature data of an x-line through 16-byte loads
    if (tn.stage >= 0) plan->stage_mask |= tn.stage & (plan->no_tma ? 6 : 7);  // (no_tma: a caller-owned table is not 16-byte aligned)
    if (plan->stage_mask & 1) plan->stage_mask &= ~8;  // (interleaved columns stage the index tables themselves)
    if (plan->lean_runs) plan->stage_mask &= ~1;       // (run scatter: the elements of a batch are not contiguous)
    if ((plan->stage_mask & 40) || plan->scatter_mode != B200_SCATTER_DETERMINISTIC || plan->lean_runs) plan->stage_mask &= ~128;
    plan->swz = false, plan->swz_w = 0, plan->group_warps = 1;
    plan->qd_tma = false, plan->mbar_off = -1, plan->ring_off = -1;
    for (auto &f : plan->in_fields) f.qd_off = -1, f.qd_tma = false, f.ring_k = -1;
    for (auto &g : plan->in_groups) g.uin_off = g.idx_off = -1;
    for (auto &g : plan->out_groups) g.tgt_off = -1;
    epb = tn.epw > 0 ? tn.epw : 8;
    if (epb > std::max(1, num_elem)) epb = std::max(1, num_elem);
    while (epb > 1 && b200_opgen_lean_layout(plan, epb) > (size_t)48 * 1024) epb--;
    if (b200_opgen_lean_layout(plan, epb) > ceed->smem_optin) return reject("element working set exceeds shared memory");
    int w = tn.cta_warps > 0 ? tn.cta_warps : 4;
    while (w > 1 && (size_t)w * b200_opgen_lean_layout(plan, epb) > ceed->smem_optin) w--;
    threads                = 32 * w;
    plan->group_smem_bytes = (int)b200_opgen_lean_layout(plan, epb);
    plan->epb              = epb;
    plan->threads          = threads;
    plan->smem_bytes       = w * plan->group_smem_bytes;
  } else if (plan->warp_mode) {
    // staging defaults in warp mode: offsets + scatter targets only (small); CEED_B200_STAGE=<bitmask 1 idx/tgt, 2 gather, 4 qdata>
    const int stage = tn.stage >= 0 ? tn.stage : 1;
    plan->stage_mask = stage;
    // warps that share one element group (1, 2 or 4): more resident warps per byte of shared memory for large Q
    int gw = tn.group_warps;
    if (gw < 1 || gw > 8) gw = 1;  // any width up to 8 warps (3 and 5 fit Q = 9, 10: 81 lines on 96 lanes, 160 swizzled tasks on 160)
    plan->group_warps = gw;
    plan->stage_mask  = stage;
    const int lanes = 32 * gw;
    // elements per group: fill the lanes in the line stages; prefer the best lane utilisation within ~14 KB per warp
    int    best = 1;
    double best_util = -1.0;
    const int Pmax = plan->bases.empty() ? Q : plan->bases[0].P;
    for (int e = 1; e <= 16; e++) {
      if (layout(e) > (size_t)14336 * gw && e > 1) break;
      double work = 0, slots = 0;
      for (int tasks : {Pmax * Pmax, Pmax * Q, lines, lines, lines}) {
        work += (double)e * tasks;
        slots += (double)lanes * ((e * tasks + lanes - 1) / lanes);
      }
      const double util = work / slots;
      if (util > best_util + 0.02) {
        best_util = util;
        best      = e;
      }
    }
    epb = tn.epw > 0 ? tn.epw : best;
    if (epb > std::max(1, num_elem)) epb = std::max(1, num_elem);
    if (layout(epb) > ceed->smem_optin) return reject("element working set exceeds shared memory");
    const int warps = tn.cta_warps > 0 ? tn.cta_warps : 4;
    int w = std::max(1, warps / gw);  // groups per CTA
    while (w > 1 && (size_t)w * layout(epb) > ceed->smem_optin) w--;
    threads                = 32 * gw * w;
    plan->group_warps      = gw;
    plan->group_smem_bytes = (int)layout(epb);
    plan->epb              = epb;
    plan->threads          = threads;
    plan->smem_bytes       = (int)std::max<size_t>((size_t)w * layout(epb), 16);
  } else {
    plan->stage_mask = 7;
    int target = getenv("CEED_B200_THREADS") ? atoi(getenv("CEED_B200_THREADS")) : 256;
    epb        = tn.epw > 0 ? tn.epw : std::max(1, target / lines);
    // aim for two resident blocks per SM so that one block's barrier / copy waits are covered by the other
    const bool   epb_forced = tn.epw > 0;
    const size_t smem_goal  = epb_forced ? ceed->smem_optin : (ceed->smem_sm - 2048) / 2;
    while (epb > 1 && layout(epb) > smem_goal) epb--;
    if (layout(epb) > ceed->smem_optin) {
      plan->async_copy = false;
      if (layout(epb) > ceed->smem_optin) return reject("element working set exceeds shared memory");
    }
    if (epb > std::max(1, num_elem)) epb = std::max(1, num_elem);
    plan->epb = epb;
    threads   = ((epb * lines + 31) / 32) * 32;
    if (threads > 1024) threads = 1024;
    if (threads < 64) threads = 64;
    plan->threads          = threads;
    plan->group_smem_bytes = (int)layout(epb);
    plan->smem_bytes       = (int)std::max<size_t>(layout(epb), 16);
  }
  {
    // occupancy target handed to __launch_bounds__: limited by shared memory and by a ~96 register/thread budget
    int by_smem = (int)(ceed->smem_sm / (size_t)(plan->smem_bytes + 1024));
    int by_regs = 65536 / (threads * 96);
    int minb    = std::max(1, std::min(by_smem, by_regs));
    if (tn.minb > 0) minb = tn.minb;
    // never ask ptxas for more resident blocks than shared memory / the thread limit allow: it would only cap registers
    minb = std::min(minb, std::max(1, by_smem));
    minb = std::min(minb, std::max(1, 2048 / threads));
    plan->blocks_per_sm = std::max(1, minb);
    tn.epw = plan->epb, tn.group_warps = plan->group_warps, tn.cta_warps = plan->threads / 32, tn.minb = plan->blocks_per_sm;
    tn.qf_mode = plan->lean ? 4 : (plan->qf_xline ? 3 : (plan->qf_pointwise ? plan->qf_pp : 0)), tn.qf_unroll = plan->qf_unroll, tn.stage = plan->stage_mask;
    plan->resolved = tn;
  }
  plan->fused      = true;
  return B200_SUCCESS;
}

// ------------------------------------------------------------------------------------------------ code generation
namespace {

struct Gen {
  B200Operator op;
  B200OpPlan  *plan;
  int          add;
  oss          c;
  int          Q, Qs, E, NT, S;
  // work distribution: block mode = all NT threads of the CTA share a batch of E elements and meet at __syncthreads();
  // warp mode = every warp owns its own group of E elements and its own shared-memory slice and only ever executes
  // __syncwarp(): warps drift apart, so the global-memory phases of some overlap the FP64 phases of others.
  bool         warp_mode = false;
  int          TS        = 0;            // task-loop stride (threads that share a group)
  string       TID, SMBASE, SYNC;        // lane id expression, shared-memory base, barrier statement
  string       QLD = "__ldg";            // load intrinsic of streamed EVAL_NONE inputs (stage bit 128: __ldcs = evict-first in L2)
  // plane layouts (see b200_opgen_plan): quadrature rows have pitch QP, z-layers are SZ apart, lanes enumerate the fastest
  // quadrature index with extent LQ; swizzled: QP = LQ = 8 and the x index of a point is XOR-ed with its y index
  bool         swz = false;
  bool         lin = false;  // even-Q linear layout: QP = Q, SZ = lin_sz(Q), x-lines through 16-byte accesses
  int          QP = 0, SZ = 0, LQ = 0;
  int          W = 0;                                                                                        // row width of the swizzled layouts (8 or 16)
  int          lp(const B200GenBasis &b) const { return swz ? (b.P <= 8 ? 8 : 16) : b.P; }                   // lane extent of the fastest node index
  int          sz1(const B200GenBasis &b) const { return swz ? swz_sz(lp(b), b.P * b.P) : b.P * b.P; }       // z-stride of T1 [qz][j][i]
  int          sz2(const B200GenBasis &b) const { return swz ? swz_sz(lp(b), Q * odd_pad(b.P)) : Q * odd_pad(b.P); }  // z-stride of T2 [qz][qy][i]
  string       xq(const string &x, const string &y) const { return swz ? "((" + x + ") ^ (" + y + "))" : "(" + x + ")"; }
  string       xq(int x, const string &y) const { return swz ? "(" + std::to_string(x) + " ^ (" + y + "))" : std::to_string(x); }
  // Q consecutive doubles of an x-line of a quadrature plane: scalar accesses (swizzled: XOR-ed index), or -- even-Q linear layout --
  // 16-byte accesses (every row starts at a 16-byte boundary there)
  void row_load(const string &ptr, const string &name, const string &y, const string &ind) {
    if (lin) {
      for (int q = 0; q < Q; q += 2) {
        c << ind << "const double2 " << name << "v" << q << " = *(const double2 *)(" << ptr << " + " << q << ");\n";
        c << ind << "const double " << name << q << " = " << name << "v" << q << ".x, " << name << q + 1 << " = " << name << "v" << q << ".y;\n";
      }
    } else {
      for (int q = 0; q < Q; q++) c << ind << "const double " << name << q << " = " << ptr << "[" << xq(q, y) << "];\n";
    }
  }
  void row_store(const string &ptr, const string &name, const string &y, const string &ind) {
    if (lin) {
      for (int q = 0; q < Q; q += 2) c << ind << "*(double2 *)(" << ptr << " + " << q << ") = make_double2(" << name << q << ", " << name << q + 1 << ");\n";
    } else {
      for (int q = 0; q < Q; q++) c << ind << ptr << "[" << xq(q, y) << "] = " << name << q << ";\n";
    }
  }
  string       smw_expr() const {
    return warp_mode ? "sm + (threadIdx.x / " + std::to_string(TS) + ") * " + std::to_string(plan->group_smem_bytes / 8) : "sm";
  }
  string       smw_decl() const {
    return warp_mode ? "  double *const smw = sm + (threadIdx.x / " + std::to_string(TS) + ") * " + std::to_string(plan->group_smem_bytes / 8) + ";\n" : "";
  }

  const B200GenBasis &basis(int id) const { return plan->bases[id]; }
  string plane(int pl, const string &le) const {
    // address (in doubles) of plane `pl` of local element `le`
    oss s;
    s << SMBASE << " + ((" << pl << ") * " << E << " + " << le << ") * " << S;
    return s.str();
  }

  // L-vector index expression for (entry n of element e, component c) of a restriction
  string lidx(B200Restriction r, const string &tbl, const string &e, const string &n, const string &comp) const {
    oss s;
    if (r->is_strided) {
      s << "((long long)(" << n << ") * " << r->strides[0] << "LL + (long long)(" << comp << ") * " << r->strides[1] << "LL + (" << e << ") * "
        << r->strides[2] << "LL)";
    } else {
      s << "((long long)" << tbl << "[(" << e << ") * " << r->elem_size << "LL + (" << n << ")] + (long long)(" << comp << ") * " << r->comp_stride
        << "LL)";
    }
    return s.str();
  }

  // Every stage is its own __noinline__ device function: ptxas then allocates registers / uniform registers per stage
  // instead of hoisting the coefficient loads of ALL stages to the kernel entry (which spills the uniform register file).
  std::vector<string> calls;  // kernel body: stage calls and barriers, in order
  int                 n_stage = 0;
  void task_loop_begin(const string &ntasks) {
    const string name = "b200_stage_" + std::to_string(n_stage++);
    c << "static __device__ __noinline__ void " << name << "(const long long e0) {\n";
    c << smw_decl();
    c << "    for (int t = " << TID << "; t < " << ntasks << "; t += " << TS << ") {\n";
    calls.push_back("    " + name + "(e0);\n");
  }
  void task_loop_end() { c << "    }\n}\n\n"; }
  void barrier() { calls.push_back("    " + SYNC + "\n"); }
  void comment(const string &text) { c << "// " << text << "\n"; }

  // out[0..n_out) = M (n_out x n_in, row-major cM[o*n_in+i]) * in   or transposed: out[o] = sum_i cM[i*n_out+o] in[i]
  void contract(const string &mat, int n_in, int n_out, bool transposed, const string &in, const string &out, const string &indent) {
    for (int o = 0; o < n_out; o++) {
      c << indent << "double " << out << o << " = ";
      for (int i = 0; i < n_in; i++) {
        const int idx = transposed ? i * n_out + o : o * n_in + i;
        if (i == 0) c << mat << "[" << idx << "] * " << in << i;
      }
      c << ";\n";
      for (int i = 1; i < n_in; i++) {
        const int idx = transposed ? i * n_out + o : o * n_in + i;
        c << indent << out << o << " = fma(" << mat << "[" << idx << "], " << in << i << ", " << out << o << ");\n";
      }
    }
  }

  void emit_header() {
    B200QFunction qf = op->qf;
    c << "// Fused operator kernel generated by ceed-b200 for QFunction " << qf->kernel_name << "\n";
    if (plan->qf_pointwise && plan->qf_pp > 1) c << "#define CEED_Q_VLA " << plan->qf_pp << "  // the QFunction is called with Q = " << plan->qf_pp << " points\n";
    if (plan->qf_xline) c << "#define CEED_Q_VLA " << Q << "  // the QFunction is called on whole x-lines (Q = " << Q << " points)\n";
    c << "#include <b200-jit.h>\n";
    if (plan->qd_tma || (plan->stage_mask & 64)) c << "#include <b200-tma.h>\n";
    c << "#include \"" << qf->source_path << "\"\n\n";
    for (size_t b = 0; b < plan->bases.size(); b++) {
      const B200Basis bs = plan->bases[b].basis;
      auto emit_mat      = [&](const string &name, const std::vector<double> &m) {
        c << "__constant__ double " << name << "[" << m.size() << "] = {";
        for (size_t i = 0; i < m.size(); i++) c << (i ? ", " : "") << hexd(m[i]);
        c << "};\n";
      };
      emit_mat("cB" + std::to_string(b), bs->interp);
      if (bs->has_collo_grad) emit_mat("cG" + std::to_string(b), bs->collo_grad);
      emit_mat("cW" + std::to_string(b), bs->q_weight);
    }
    c << "\nstruct B200OpArgs {\n  long long num_elem;\n  void *ctx;\n  const double *in_ptr[16];\n  double *out_ptr[16];\n"
      << "  const int *in_idx[16];\n  const int *out_idx[16];\n  double *out_aux[16];\n"
      << "  const int *ord_pred_ptr, *ord_pred_idx;\n  int *ord_flags, *ord_sync;\n  long long ord_num_halo;\n};\n"
      // the argument block lives in constant memory (written by the host before the launch): every stage function reads its
      // pointers with uniform constant loads instead of generic loads through a reference to the kernel parameter
      << "__constant__ B200OpArgs b200a;\n\n";
    c << "extern __shared__ __align__(16) double sm[];\n";
    // end of the element range of this launch (kernel parameter e_end, published by thread 0): stage functions clamp tail groups with it
    c << "__shared__ long long b200_ne;\n\n";
    // cp.async (LDGSTS): global -> shared without staging registers; completion tracked per thread with commit/wait groups
    c << "__device__ __forceinline__ void b200_cp4(void *dst, const void *src) {\n"
      << "  asm volatile(\"cp.async.ca.shared.global [%0], [%1], 4;\" ::\"r\"((unsigned)__cvta_generic_to_shared(dst)), \"l\"(src) : \"memory\");\n}\n"
      << "__device__ __forceinline__ void b200_cp8(void *dst, const void *src) {\n"
      << "  asm volatile(\"cp.async.ca.shared.global [%0], [%1], 8;\" ::\"r\"((unsigned)__cvta_generic_to_shared(dst)), \"l\"(src) : \"memory\");\n}\n"
      << "__device__ __forceinline__ void b200_cp16(void *dst, const void *src) {\n"
      << "  asm volatile(\"cp.async.cg.shared.global [%0], [%1], 16;\" ::\"r\"((unsigned)__cvta_generic_to_shared(dst)), \"l\"(src) : \"memory\");\n}\n"
      << "__device__ __forceinline__ void b200_cp_commit() { asm volatile(\"cp.async.commit_group;\" ::: \"memory\"); }\n"
      << "__device__ __forceinline__ void b200_cp_wait_all() { asm volatile(\"cp.async.wait_group 0;\" ::: \"memory\"); }\n\n";
  }

  string smem_at(int byte_off, const string &type) const {
    oss s;
    s << "((" << type << " *)((char *)" << SMBASE << " + " << byte_off << "))";
    return s.str();
  }

  // value of component cc of a staged EVAL_NONE input at point `pt` of local element `le` (group starting at element e0)
  string qd_ref(const B200GenField &fd, int cc, const string &le, const string &pt) const {
    const int Q3 = Q * Q * Q;
    oss       s;
    if (!fd.qd_tma) {
      s << smem_at(fd.qd_off, "const double") << "[(" << cc * E << " + " << le << ") * " << Q3 << " + " << pt << "]";
    } else {
      // bulk copies start at a 16-byte boundary: with an odd Q^3 the component block of this group may begin one double later
      s << smem_at(fd.qd_off, "const double") << "[" << cc * fd.qd_cs << " + (" << le << ") * " << Q3 << " + " << pt;
      if (Q3 % 2) s << " + (int)((" << (long long)cc * fd.rstr->strides[1] << "LL + e0 * " << Q3 << "LL) & 1)";
      s << "]";
    }
    return s.str();
  }

  // ---- asynchronous staging (issued one batch ahead) ---------------------------------------------
  // offsets of the gathers of batch e0 (all staged input groups) -> IDX buffers
  bool emit_issue_idx() {
    bool any = false;
    for (auto &g : plan->in_groups) any = any || g.idx_off >= 0;
    if (!any) return false;
    c << "static __device__ __noinline__ void b200_issue_idx(const long long e0) {\n";
    c << "  if (e0 >= b200_ne) return;\n";
    c << smw_decl();
    c << "  const int ne = (int)((b200_ne - e0 < " << E << ") ? b200_ne - e0 : " << E << ");\n";
    for (auto &g : plan->in_groups) {
      if (g.idx_off < 0) continue;
      const int es = g.rstr->elem_size;
      // tail groups: the gather stage runs branch-free on clamped element numbers, so every slot gets valid offsets
      c << "  { int *dst = " << smem_at(g.idx_off, "int") << "; const int *src = b200a.in_idx[" << g.slot << "] + e0 * " << es << "LL;\n";
      c << "    for (int i = " << TID << "; i < " << E * es << "; i += " << TS << ") b200_cp4(dst + i, src + (i < ne * " << es << " ? i : (ne - 1) * " << es
        << " + i % " << es << ")); }\n";
    }
    c << "}\n\n";
    return true;
  }
  // scatter targets of batch e0 (all staged output groups) -> TGT buffers
  bool emit_issue_tgt() {
    bool any = false;
    for (auto &g : plan->out_groups) any = any || g.tgt_off >= 0;
    if (!any) return false;
    c << "static __device__ __noinline__ void b200_issue_tgt(const long long e0) {\n";
    c << "  if (e0 >= b200_ne) return;\n";
    c << smw_decl();
    c << "  const int ne = (int)((b200_ne - e0 < " << E << ") ? b200_ne - e0 : " << E << ");\n";
    for (auto &g : plan->out_groups) {
      if (g.tgt_off < 0) continue;
      const int es = g.rstr->elem_size;
      c << "  { int *dst = " << smem_at(g.tgt_off, "int") << "; const int *src = b200a.out_idx[" << g.slot << "] + e0 * " << es << "LL;\n";
      c << "    for (int i = " << TID << "; i < ne * " << es << "; i += " << TS << ") b200_cp4(dst + i, src + i); }\n";
    }
    c << "}\n\n";
    return true;
  }
  // gathers of batch e0 through the staged offsets -> UIN buffers
  bool emit_issue_gather() {
    bool any = false;
    for (auto &g : plan->in_groups) any = any || g.uin_off >= 0;
    if (!any) return false;
    c << "static __device__ __noinline__ void b200_issue_gather(const long long e0) {\n";
    c << "  if (e0 >= b200_ne) return;\n";
    c << smw_decl();
    c << "  const int ne = (int)((b200_ne - e0 < " << E << ") ? b200_ne - e0 : " << E << ");\n";
    for (auto &g : plan->in_groups) {
      if (g.uin_off < 0) continue;
      const int es = g.rstr->elem_size;
      c << "  { const int *idx = " << smem_at(g.idx_off, "int") << "; double *dst = " << smem_at(g.uin_off, "double") << ";\n";
      c << "    const double *src = b200a.in_ptr[" << g.slot << "];\n";
      c << "    for (int i = " << TID << "; i < ne * " << es << "; i += " << TS << ") {\n";
      c << "      const long long l = idx[i];\n";
      for (int cc = 0; cc < g.nc; cc++)
        c << "      b200_cp8(dst + i + " << cc * E * es << ", src + l + " << (long long)cc * g.rstr->comp_stride << "LL);\n";
      c << "    } }\n";
    }
    c << "}\n\n";
    return true;
  }
  // streamed EVAL_NONE inputs (quadrature data) of batch e0 -> QD buffers; 16-byte copies when the source is aligned
  bool emit_issue_qd() {
    bool any = false;
    for (auto &f : plan->in_fields) any = any || f.qd_off >= 0;
    if (!any) return false;
    const int Q3 = Q * Q * Q;
    c << "static __device__ __noinline__ void b200_issue_qd(const long long e0) {\n";
    c << "  if (e0 >= b200_ne) return;\n";
    c << smw_decl();
    c << "  const int ne = (int)((b200_ne - e0 < " << E << ") ? b200_ne - e0 : " << E << ");\n";
    if (plan->qd_tma) {
      // one elected lane: a bulk copy per (field, component) = the contiguous block of the group's elements, completion on the group's mbarrier
      c << "  if (" << TID << " == 0) {\n";
      c << "    const unsigned bar = b200_smem_u32((char *)" << SMBASE << " + " << plan->mbar_off << ");\n";
      c << "    b200_fence_proxy_async();\n";
      int k = 0;
      for (auto &f : plan->in_fields) {
        if (!f.qd_tma) continue;
        for (int cc = 0; cc < f.nc; cc++, k++) {
          c << "    const double *src" << k << " = b200a.in_ptr[" << f.slot << "] + " << (long long)cc * f.rstr->strides[1] << "LL + e0 * " << Q3 << "LL;\n";
          if (Q3 % 2) {
            c << "    const unsigned sh" << k << " = (unsigned)((" << (long long)cc * f.rstr->strides[1] << "LL + e0 * " << Q3 << "LL) & 1);\n";
            c << "    unsigned nb" << k << " = ((ne * " << Q3 << " + sh" << k << ") * 8 + 15) & ~15u;\n";
            // the last block of the array: rounding up would read 8 bytes past its end -- copy up to the last 16-byte boundary and
            // let this lane move the final double itself
            c << "    { const double *end = b200a.in_ptr[" << f.slot << "] + " << (long long)cc * f.rstr->strides[1] << "LL + b200a.num_elem * " << Q3 << "LL;\n";
            c << "      if ((const char *)(src" << k << " - sh" << k << ") + nb" << k << " > (const char *)end) {\n";
            c << "        nb" << k << " -= 16;\n";
            c << "        ((double *)((char *)" << SMBASE << " + " << f.qd_off + cc * f.qd_cs * 8 << "))[(end - 1) - (src" << k << " - sh" << k << ")] = end[-1];\n      } }\n";
          } else {
            c << "    const unsigned sh" << k << " = 0, nb" << k << " = ne * " << Q3 * 8 << ";\n";
          }
        }
      }
      c << "    b200_mbar_expect_tx(bar, 0";
      for (int i = 0; i < k; i++) c << " + nb" << i;
      c << ");\n";
      k = 0;
      for (auto &f : plan->in_fields) {
        if (!f.qd_tma) continue;
        for (int cc = 0; cc < f.nc; cc++, k++)
          c << "    b200_bulk_g2s(b200_smem_u32((char *)" << SMBASE << " + " << f.qd_off + cc * f.qd_cs * 8 << "), src" << k << " - sh" << k << ", nb" << k
            << ", bar);\n";
      }
      c << "  }\n";
    }
    for (auto &f : plan->in_fields) {
      if (f.qd_off < 0 || f.qd_tma) continue;
      for (int cc = 0; cc < f.nc; cc++) {
        c << "  { double *dst = " << smem_at(f.qd_off, "double") << " + " << cc * E * Q3 << ";\n";
        c << "    const double *src = b200a.in_ptr[" << f.slot << "] + " << (long long)cc * f.rstr->strides[1] << "LL + e0 * " << Q3 << "LL;\n";
        c << "    const int n = ne * " << Q3 << ";\n";
        c << "    if (((((unsigned long long)src) | ((unsigned long long)dst)) & 15) == 0 && (n & 1) == 0) {\n";
        c << "      for (int i = " << TID << " * 2; i < n; i += " << 2 * TS << ") b200_cp16(dst + i, src + i);\n";
        c << "    } else {\n";
        c << "      for (int i = " << TID << "; i < n; i += " << TS << ") b200_cp8(dst + i, src + i);\n";
        c << "    } }\n";
      }
    }
    c << "}\n\n";
    return true;
  }

  // stage bit 64: bulk L2 prefetch (cp.async.bulk.prefetch.L2) of the directly loaded quadrature data of the group's NEXT batch;
  // one lane, one instruction per (field, component) block -- costs no shared memory, so the occupancy of the kernel is unchanged
  bool emit_prefetch_qd() {
    if (!(plan->stage_mask & 64) || !plan->warp_mode) return false;
    const int Q3 = Q * Q * Q;
    bool      any = false;
    for (auto &f : plan->in_fields)
      any = any || (f.emode == B200_EVAL_NONE && f.qd_off < 0 && f.rstr->is_strided && f.rstr->strides[0] == 1 && f.rstr->strides[2] == f.rstr->elem_size);
    if (!any) return false;
    c << "static __device__ __noinline__ void b200_prefetch_qd(const long long e0) {\n";
    c << "  if (e0 >= b200_ne || " << TID << " != 0) return;\n";
    c << "  const int ne = (int)((b200_ne - e0 < " << E << ") ? b200_ne - e0 : " << E << ");\n";
    for (auto &f : plan->in_fields) {
      if (!(f.emode == B200_EVAL_NONE && f.qd_off < 0 && f.rstr->is_strided && f.rstr->strides[0] == 1 && f.rstr->strides[2] == f.rstr->elem_size)) continue;
      for (int cc = 0; cc < f.nc; cc++) {
        c << "  { const unsigned long long a = (unsigned long long)(b200a.in_ptr[" << f.slot << "] + " << (long long)cc * f.rstr->strides[1] << "LL + e0 * " << Q3
          << "LL);\n";
        c << "    b200_bulk_prefetch_l2((const void *)(a & ~15ULL), (unsigned)(((a & 15ULL) + (unsigned long long)ne * " << Q3 * 8 << " + 15) & ~15ULL)); }\n";
      }
    }
    c << "}\n\n";
    return true;
  }

  // ---- input side -------------------------------------------------------------------------------
  void emit_gather_z(const B200GenGroup &g) {
    const B200GenBasis &b  = basis(g.basis_id);
    const int           P  = b.P;
    const string        sl = std::to_string(g.slot);
    const int           ntasks = E * g.nc * P * P;
    comment("gather + z-contraction, input group slot " + std::to_string(g.slot));
    auto emit_compute_store = [&](const string &sfx, const string &ind) {
      // u<k><sfx> -> z-contraction -> T1 plane (or straight into the Uq plane for collocated bases)
      if (b.collocated) {
        c << ind << "double *dst = " << plane(g.plane0, "le" + sfx) << " + cc" << sfx << " * " << E * S << " + (ij" << sfx << " / " << P << ") * " << QP
          << " + " << xq("ij" + sfx + " % " + std::to_string(P), "ij" + sfx + " / " + std::to_string(P)) << ";\n";
        for (int k = 0; k < P; k++) c << ind << "dst[" << k * SZ << "] = u" << k << sfx << ";\n";
      } else {
        c << ind << "double *dst = " << plane(g.plane0, "le" + sfx) << " + cc" << sfx << " * " << E * S << " + ij" << sfx << ";\n";  // T1 [qz][j][i]
        for (int o = 0; o < Q; o++) {
          c << ind << "double r" << o << " = cB" << g.basis_id << "[" << o * P << "] * u0" << sfx << ";\n";
          for (int i = 1; i < P; i++) c << ind << "r" << o << " = fma(cB" << g.basis_id << "[" << o * P + i << "], u" << i << sfx << ", r" << o << ");\n";
        }
        for (int q = 0; q < Q; q++) c << ind << "dst[" << q * sz1(b) << "] = r" << q << ";\n";
      }
    };
    const int rounds = (ntasks + TS - 1) / TS;
    if (warp_mode && g.uin_off < 0 && rounds * P <= 32 && !getenv("CEED_B200_NO_GATHER_BATCH")) {
      // Batched gather: every lane first issues the offset loads of ALL its tasks, then all value loads, and only then
      // computes -- two memory latencies per group instead of two per task round.
      const string name = "b200_stage_" + std::to_string(n_stage++);
      c << "static __device__ __noinline__ void " << name << "(const long long e0) {\n";
      c << smw_decl();
      calls.push_back("    " + name + "(e0);\n");
      c << "    const int lane = " << TID << ";\n";
      for (int r = 0; r < rounds; r++) {
        const string x = "_" + std::to_string(r);
        c << "    const int t" << x << " = lane + " << r * TS << ", tc" << x << " = t" << x << " < " << ntasks << " ? t" << x << " : " << ntasks - 1 << ";\n";
        c << "    const int ij" << x << " = tc" << x << " % " << P * P << ", cc" << x << " = (tc" << x << " / " << P * P << ") % " << g.nc << ", le" << x
          << " = tc" << x << " / " << P * P * g.nc << ";\n";
        c << "    const long long e" << x << " = (e0 + le" << x << " < b200_ne) ? e0 + le" << x << " : b200_ne - 1;\n";
      }
      if (!g.rstr->is_strided) {
        for (int r = 0; r < rounds; r++) {
          const string x = "_" + std::to_string(r);
          for (int k = 0; k < P; k++) {
            if (g.idx_off >= 0)
              c << "    const long long l" << k << x << " = " << smem_at(g.idx_off, "const int") << "[le" << x << " * " << P * P * P << " + ij" << x << " + "
                << k * P * P << "];\n";
            else
              c << "    const long long l" << k << x << " = __ldg(b200a.in_idx[" << sl << "] + e" << x << " * " << P * P * P << "LL + ij" << x << " + " << k * P * P
                << ");\n";
          }
        }
      }
      for (int r = 0; r < rounds; r++) {
        const string x = "_" + std::to_string(r);
        for (int k = 0; k < P; k++) {
          if (g.rstr->is_strided)
            c << "    const double u" << k << x << " = __ldg(b200a.in_ptr[" << sl << "] + "
              << lidx(g.rstr, "", "e" + x, "ij" + x + " + " + std::to_string(k * P * P), "cc" + x) << ");\n";
          else
            c << "    const double u" << k << x << " = __ldg(b200a.in_ptr[" << sl << "] + l" << k << x << " + (long long)cc" << x << " * " << g.rstr->comp_stride
              << "LL);\n";
        }
      }
      for (int r = 0; r < rounds; r++) {
        const string x = "_" + std::to_string(r);
        c << "    if (t" << x << " < " << ntasks << ") {\n";
        emit_compute_store(x, "      ");
        c << "    }\n";
      }
      c << "}\n\n";
      return;
    }
    task_loop_begin(std::to_string(ntasks));
    c << "      const int ij = t % " << P * P << ", cc = (t / " << P * P << ") % " << g.nc << ", le = t / " << P * P * g.nc << ";\n";
    c << "      const long long e = (e0 + le < b200_ne) ? e0 + le : b200_ne - 1;  // clamped: tail groups gather a valid element\n";
    if (g.uin_off >= 0) {
      // values were gathered into shared memory by cp.async while the previous batch was computing
      c << "      const double *uin = " << smem_at(g.uin_off, "double") << " + (cc * " << E << " + le) * " << P * P * P << " + ij;\n";
      for (int k = 0; k < P; k++) c << "      const double u" << k << " = uin[" << k * P * P << "];\n";
    } else {
      for (int k = 0; k < P; k++) {
        if (g.idx_off >= 0)  // offsets of this group were staged into shared memory while the previous group was computing
          c << "      const double u" << k << " = __ldg(b200a.in_ptr[" << sl << "] + (long long)" << smem_at(g.idx_off, "const int") << "[le * " << P * P * P
            << " + ij + " << k * P * P << "] + (long long)cc * " << g.rstr->comp_stride << "LL);\n";
        else
          c << "      const double u" << k << " = __ldg(b200a.in_ptr[" << sl << "] + "
            << lidx(g.rstr, "b200a.in_idx[" + sl + "]", "e", "ij + " + std::to_string(k * P * P), "cc") << ");\n";
      }
    }
    emit_compute_store("", "      ");
    task_loop_end();
  }

  void emit_interp_y(const B200GenGroup &g) {
    const B200GenBasis &b = basis(g.basis_id);
    if (b.collocated) return;
    const int P = b.P, Ps = odd_pad(P);
    comment("y-contraction, input group slot " + std::to_string(g.slot));
    const int LP = lp(b);
    task_loop_begin(std::to_string(E * g.nc * Q * LP));
    c << "      const int i = t % " << LP << ", qz = (t / " << LP << ") % " << Q << ", cc = (t / " << LP * Q << ") % " << g.nc << ", le = t / "
      << LP * Q * g.nc << ";\n";
    if (LP != P) c << "      if (i >= " << P << ") continue;\n";
    c << "      const double *src = " << plane(g.plane0, "le") << " + cc * " << E * S << " + qz * " << sz1(b) << " + i;\n";
    c << "      double *dst = " << plane(g.plane0 + g.nc, "le") << " + cc * " << E * S << " + qz * " << sz2(b) << " + i;\n";  // T2 in plane B
    for (int j = 0; j < P; j++) c << "      const double u" << j << " = src[" << j * P << "];\n";
    contract("cB" + std::to_string(g.basis_id), P, Q, false, "u", "r", "      ");
    for (int q = 0; q < Q; q++) c << "      dst[" << q * Ps << "] = r" << q << ";\n";
    task_loop_end();
  }

  void emit_interp_x(const B200GenGroup &g) {
    const B200GenBasis &b = basis(g.basis_id);
    const int           P = b.P, Ps = odd_pad(P);
    if (b.collocated && !g.use_grad) return;
    comment("x-contraction (+ d/dx), input group slot " + std::to_string(g.slot));
    task_loop_begin(std::to_string(E * g.nc * Q * LQ));
    c << "      const int qy = t % " << LQ << ", qz = (t / " << LQ << ") % " << Q << ", cc = (t / " << Q * LQ << ") % " << g.nc << ", le = t / " << Q * LQ * g.nc
      << ";\n";
    if (LQ != Q) c << "      if (qy >= " << Q << ") continue;\n";
    c << "      double *uq = " << plane(g.plane0, "le") << " + cc * " << E * S << " + qz * " << SZ << " + qy * " << QP << ";\n";
    if (b.collocated) {
      row_load("uq", "r", "qy", "      ");
    } else {
      c << "      const double *src = " << plane(g.plane0 + g.nc, "le") << " + cc * " << E * S << " + qz * " << sz2(b) << " + qy * " << Ps << ";\n";
      for (int i = 0; i < P; i++) c << "      const double u" << i << " = src[" << i << "];\n";
      contract("cB" + std::to_string(g.basis_id), P, Q, false, "u", "r", "      ");
      row_store("uq", "r", "qy", "      ");
    }
    if (g.use_grad) {
      c << "      double *gx = " << plane(g.plane0 + 2 * g.nc, "le") << " + cc * " << E * S << " + qz * " << SZ << " + qy * " << QP << ";\n";
      contract("cG" + std::to_string(g.basis_id), Q, Q, false, "r", "d", "      ");
      row_store("gx", "d", "qy", "      ");
    }
    task_loop_end();
  }

  void emit_grad_y(const B200GenGroup &g) {
    if (!g.use_grad) return;
    comment("d/dy, input group slot " + std::to_string(g.slot));
    task_loop_begin(std::to_string(E * g.nc * Q * LQ));
    c << "      const int qx = t % " << LQ << ", qz = (t / " << LQ << ") % " << Q << ", cc = (t / " << Q * LQ << ") % " << g.nc << ", le = t / "
      << Q * LQ * g.nc << ";\n";
    if (LQ != Q) c << "      if (qx >= " << Q << ") continue;\n";
    // swizzled rows: point (qx, m) of the line sits at m * QP + (qx ^ m)
    c << "      const double *src = " << plane(g.plane0, "le") << " + cc * " << E * S << " + qz * " << SZ << (swz ? "" : " + qx") << ";\n";
    c << "      double *dst = " << plane(g.plane0 + g.nc, "le") << " + cc * " << E * S << " + qz * " << SZ << (swz ? "" : " + qx") << ";\n";
    auto at = [&](int m) { return swz ? std::to_string(m * QP) + " + (qx ^ " + std::to_string(m) + ")" : std::to_string(m * QP); };
    for (int m = 0; m < Q; m++) c << "      const double u" << m << " = src[" << at(m) << "];\n";
    contract("cG" + std::to_string(g.basis_id), Q, Q, false, "u", "d", "      ");
    for (int q = 0; q < Q; q++) c << "      dst[" << at(q) << "] = d" << q << ";\n";
    task_loop_end();
  }

  // ---- quadrature-point stage -------------------------------------------------------------------
  // plane holding d/dz (inputs) or the z-component of the output gradient: in place of the values when they are not needed
  int gz_plane(const B200GenGroup &g, int cc) const { return g.use_interp ? g.plane0 + 3 * g.nc + cc : g.plane0 + cc; }

  void emit_grad_z(const B200GenGroup &g) {
    if (!g.use_grad) return;
    comment("d/dz (z-lines), input group slot " + std::to_string(g.slot));
    task_loop_begin(std::to_string(E * g.nc * Q * Q));
    c << "      const int pxy = (t % " << Q << ") + ((t / " << Q << ") % " << Q << ") * " << Qs << ", cc = (t / " << Q * Q << ") % " << g.nc << ", le = t / "
      << Q * Q * g.nc << ";\n";
    c << "      const double *src = " << plane(g.plane0, "le") << " + cc * " << E * S << " + pxy;\n";
    c << "      double *dst = " << plane(gz_plane(g, 0), "le") << " + cc * " << E * S << " + pxy;\n";
    for (int m = 0; m < Q; m++) c << "      const double u" << m << " = src[" << m * Q * Qs << "];\n";
    contract("cG" + std::to_string(g.basis_id), Q, Q, false, "u", "d", "      ");
    for (int q = 0; q < Q; q++) c << "      dst[" << q * Q * Qs << "] = d" << q << ";\n";
    task_loop_end();
  }

  void emit_gradT_z(const B200GenGroup &g) {
    if (!g.use_grad) return;
    comment("(d/dz)^T (z-lines), output group slot " + std::to_string(g.slot));
    task_loop_begin(std::to_string(E * g.nc * Q * Q));
    c << "      const int pxy = (t % " << Q << ") + ((t / " << Q << ") % " << Q << ") * " << Qs << ", cc = (t / " << Q * Q << ") % " << g.nc << ", le = t / "
      << Q * Q * g.nc << ";\n";
    c << "      const double *src = " << plane(gz_plane(g, 0), "le") << " + cc * " << E * S << " + pxy;\n";
    c << "      double *vq = " << plane(g.plane0, "le") << " + cc * " << E * S << " + pxy;\n";
    for (int m = 0; m < Q; m++) c << "      const double u" << m << " = src[" << m * Q * Qs << "];\n";
    contract("cG" + std::to_string(g.basis_id), Q, Q, true, "u", "d", "      ");
    for (int q = 0; q < Q; q++) c << "      vq[" << q * Q * Qs << "] " << (g.use_interp ? "+=" : "=") << " d" << q << ";\n";
    task_loop_end();
  }

  // QFunction over independent quadrature points (pointwise mode).  PP = points per lane: with PP == 2 a lane owns two
  // x-adjacent points, streams the quadrature data with 16-byte loads (the per-warp load-instruction rate, not bytes in
  // flight, limits HBM throughput at low occupancy -- scripts/ubench/stream_model.cu) and calls the QFunction with Q = 2.
  void emit_qf_points() {
    B200QFunction qf = op->qf;
    const int     Q3 = Q * Q * Q, PP = plan->qf_pp;
    comment("quadrature points: independent points, inputs from the planes / streamed from global memory");
    const string name = "b200_stage_" + std::to_string(n_stage++);
    c << "static __device__ __noinline__ void " << name << "(const long long e0) {\n";
    c << smw_decl();
    c << "    const CeedScalar *in[" << std::max<size_t>(1, qf->inputs.size()) << "];\n";
    c << "    CeedScalar *out[" << std::max<size_t>(1, qf->outputs.size()) << "];\n";
    const int unroll = plan->qf_unroll;
    c << "    #pragma unroll " << unroll << "\n";
    c << "    for (int t = " << TID << "; t < " << E * Q3 / PP << "; t += " << TS << ") {\n";
    calls.push_back("    " + name + "(e0);\n");
    c << "      const int pt = (t * " << PP << ") % " << Q3 << ", le = (t * " << PP << ") / " << Q3 << ";\n";
    c << "      const int qx = pt % " << Q << ", qy = (pt / " << Q << ") % " << Q << ", qz = pt / " << Q * Q << ";\n";
    c << "      const int p = (qz * " << Q << " + qy) * " << Qs << " + qx;\n";
    // tail groups: loads use a clamped (valid) element so that there is no branch in the loop body and the compiler can overlap
    // the global loads of several unrolled points; results of non-existent elements are simply never stored to global memory
    c << "      const long long e_real = e0 + le;\n";
    c << "      const long long e = e_real < b200_ne ? e_real : b200_ne - 1;\n";
    c << "      {\n";
    for (size_t f = 0; f < plan->in_fields.size(); f++) c << "      CeedScalar in_" << f << "[" << plan->in_fields[f].size * PP << "];\n";
    for (size_t f = 0; f < plan->out_fields.size(); f++) c << "      CeedScalar out_" << f << "[" << plan->out_fields[f].size * PP << "];\n";
    for (size_t f = 0; f < plan->in_fields.size(); f++) c << "      in[" << f << "] = in_" << f << ";\n";
    for (size_t f = 0; f < plan->out_fields.size(); f++) c << "      out[" << f << "] = out_" << f << ";\n";
    auto at = [&](int k, int h) { return "[" + std::to_string(k * PP + h) + "]"; };
    for (size_t f = 0; f < plan->in_fields.size(); f++) {
      const B200GenField &fd = plan->in_fields[f];
      const string        sl = std::to_string(fd.slot);
      switch (fd.emode) {
        case B200_EVAL_NONE:
          for (int cc = 0; cc < fd.nc; cc++) {
            if (fd.qd_off >= 0) {
              for (int h = 0; h < PP; h++) c << "      in_" << f << at(cc, h) << " = " << qd_ref(fd, cc, "le", "pt + " + std::to_string(h)) << ";\n";
            } else if (PP == 2 && fd.rstr->is_strided && fd.rstr->strides[0] == 1 && fd.rstr->strides[1] % 2 == 0 && fd.rstr->strides[2] % 2 == 0) {
              c << "      { const double2 v2 = " << QLD << "((const double2 *)(b200a.in_ptr[" << sl << "] + " << lidx(fd.rstr, "", "e", "pt", std::to_string(cc)) << "));\n";
              c << "        in_" << f << at(cc, 0) << " = v2.x; in_" << f << at(cc, 1) << " = v2.y; }\n";
            } else {
              for (int h = 0; h < PP; h++)
                c << "      in_" << f << at(cc, h) << " = " + QLD + "(b200a.in_ptr[" << sl << "] + "
                  << lidx(fd.rstr, "b200a.in_idx[" + sl + "]", "e", "pt + " + std::to_string(h), std::to_string(cc)) << ");\n";
            }
          }
          break;
        case B200_EVAL_WEIGHT:
          for (int h = 0; h < PP; h++)
            c << "      in_" << f << at(0, h) << " = cW" << fd.basis_id << "[qx + " << h << "] * cW" << fd.basis_id << "[qy] * cW" << fd.basis_id << "[qz];\n";
          break;
        case B200_EVAL_INTERP: {
          const B200GenGroup &g = plan->in_groups[fd.group];
          for (int cc = 0; cc < fd.nc; cc++)
            for (int h = 0; h < PP; h++) c << "      in_" << f << at(cc, h) << " = (" << plane(g.plane0 + cc, "le") << ")[p + " << h << "];\n";
        } break;
        case B200_EVAL_GRAD: {
          const B200GenGroup &g = plan->in_groups[fd.group];
          for (int cc = 0; cc < fd.nc; cc++)
            for (int h = 0; h < PP; h++) {
              c << "      in_" << f << at(cc, h) << " = (" << plane(g.plane0 + 2 * g.nc + cc, "le") << ")[p + " << h << "];\n";
              c << "      in_" << f << at(cc + fd.nc, h) << " = (" << plane(g.plane0 + g.nc + cc, "le") << ")[p + " << h << "];\n";
              c << "      in_" << f << at(cc + 2 * fd.nc, h) << " = (" << plane(gz_plane(g, cc), "le") << ")[p + " << h << "];\n";
            }
        } break;
      }
    }
    c << "      " << qf->kernel_name << "(b200a.ctx, " << PP << ", in, out);\n";
    for (size_t gi = 0; gi < plan->out_groups.size(); gi++) {
      const B200GenGroup &g = plan->out_groups[gi];
      for (int cc = 0; cc < g.nc; cc++)
        for (int h = 0; h < PP; h++) {
          string val, vx, vy, vz;
          for (size_t f = 0; f < plan->out_fields.size(); f++) {
            const B200GenField &fd = plan->out_fields[f];
            if (fd.group != (int)gi) continue;
            const string o = "out_" + std::to_string(f);
            if (fd.emode == B200_EVAL_INTERP) val += (val.empty() ? "" : " + ") + o + at(cc, h);
            if (fd.emode == B200_EVAL_GRAD) {
              vx += (vx.empty() ? "" : " + ") + o + at(cc, h);
              vy += (vy.empty() ? "" : " + ") + o + at(cc + fd.nc, h);
              vz += (vz.empty() ? "" : " + ") + o + at(cc + 2 * fd.nc, h);
            }
          }
          if (!val.empty()) c << "      (" << plane(g.plane0 + cc, "le") << ")[p + " << h << "] = " << val << ";\n";
          if (g.use_grad) {
            c << "      (" << plane(g.plane0 + 2 * g.nc + cc, "le") << ")[p + " << h << "] = " << vx << ";\n";
            c << "      (" << plane(g.plane0 + g.nc + cc, "le") << ")[p + " << h << "] = " << vy << ";\n";
            c << "      (" << plane(gz_plane(g, cc), "le") << ")[p + " << h << "] = " << vz << ";\n";
          }
        }
    }
    for (size_t f = 0; f < plan->out_fields.size(); f++) {
      const B200GenField &fd = plan->out_fields[f];
      if (fd.emode != B200_EVAL_NONE) continue;
      c << "      if (e_real < b200_ne) {\n";
      for (int h = 0; h < PP; h++)
        emit_scatter_value(fd.rstr, fd.slot, "e", "pt + " + std::to_string(h), fd.nc,
                           [&](int cc) { return "out_" + std::to_string(f) + at(cc, h); }, "        ");
      c << "      }\n";
    }
    c << "      }\n";
    c << "    }\n}\n\n";
  }

  void emit_qf_stage() {
    B200QFunction qf = op->qf;
    comment("quadrature points: one z-line per thread; d/dz, QFunction, (d/dz)^T in registers");
    task_loop_begin(std::to_string(E * Q * LQ));
    c << "      const int qx = t % " << LQ << ", qy = (t / " << LQ << ") % " << Q << ", le = t / " << Q * LQ << ";\n";
    c << "      const long long e = e0 + le;\n";
    c << "      if (e < b200_ne" << (LQ != Q ? " && qx < " + std::to_string(Q) : "") << ") {\n";
    c << "      const int pxy = qy * " << QP << " + " << xq("qx", "qy") << ";\n";
    // z-lines of input groups that need gradients (and their d/dz)
    for (size_t gi = 0; gi < plan->in_groups.size(); gi++) {
      const B200GenGroup &g = plan->in_groups[gi];
      if (!g.use_grad) continue;
      for (int cc = 0; cc < g.nc; cc++) {
        const string tag = "g" + std::to_string(gi) + "c" + std::to_string(cc) + "_";
        c << "      const double *uq_" << tag << " = " << plane(g.plane0 + cc, "le") << " + pxy;\n";
        for (int m = 0; m < Q; m++) c << "      const double uz_" << tag << m << " = uq_" << tag << "[" << m * SZ << "];\n";
        contract("cG" + std::to_string(g.basis_id), Q, Q, false, "uz_" + tag, "dz_" + tag, "      ");
      }
    }
    // accumulators of (d/dz)^T for output groups
    for (size_t gi = 0; gi < plan->out_groups.size(); gi++) {
      const B200GenGroup &g = plan->out_groups[gi];
      if (!g.use_grad) continue;
      for (int cc = 0; cc < g.nc; cc++)
        for (int m = 0; m < Q; m++) c << "      double vz_g" << gi << "c" << cc << "_" << m << " = 0.0;\n";
    }
    // weights
    for (size_t f = 0; f < plan->in_fields.size(); f++)
      if (plan->in_fields[f].emode == B200_EVAL_WEIGHT) {
        const string w = "cW" + std::to_string(plan->in_fields[f].basis_id);
        c << "      const double wxy_" << f << " = " << w << "[qx] * " << w << "[qy];\n";
      }
    c << "      const CeedScalar *in[" << std::max<size_t>(1, qf->inputs.size()) << "];\n";
    c << "      CeedScalar *out[" << std::max<size_t>(1, qf->outputs.size()) << "];\n";
    for (size_t f = 0; f < plan->in_fields.size(); f++) c << "      CeedScalar in_" << f << "[" << plan->in_fields[f].size << "];\n";
    for (size_t f = 0; f < plan->out_fields.size(); f++) c << "      CeedScalar out_" << f << "[" << plan->out_fields[f].size << "];\n";
    for (size_t f = 0; f < plan->in_fields.size(); f++) c << "      in[" << f << "] = in_" << f << ";\n";
    for (size_t f = 0; f < plan->out_fields.size(); f++) c << "      out[" << f << "] = out_" << f << ";\n";
    // direct-load EVAL_NONE inputs are fetched one z-layer ahead (software pipelining): the loads of layer qz+1 are in
    // flight while the QFunction and the transposed z-derivative of layer qz execute
    const bool qf_ahead = !getenv("CEED_B200_NO_QFPF");
    const int  depth    = std::max(1, std::min(Q, plan->qf_ahead));  // z-layers of streamed inputs in flight per lane
    auto none_load = [&](const B200GenField &fd, int cc, const string &pt_expr) {
      const string sl = std::to_string(fd.slot);
      return QLD + "(b200a.in_ptr[" + sl + "] + " + lidx(fd.rstr, "b200a.in_idx[" + sl + "]", "e", pt_expr, std::to_string(cc)) + ")";
    };
    c << "      const int pt0 = qy * " << Q << " + qx;\n";
    if (qf_ahead)
      for (int d = 0; d < depth; d++)
        for (size_t f = 0; f < plan->in_fields.size(); f++) {
          const B200GenField &fd = plan->in_fields[f];
          if (fd.emode != B200_EVAL_NONE || fd.qd_off >= 0) continue;
          for (int cc = 0; cc < fd.nc; cc++)
            c << "      double nx_" << f << "_" << cc << "_" << d << " = " << none_load(fd, cc, "pt0 + " + std::to_string(d * Q * Q)) << ";\n";
        }
    for (int qz = 0; qz < Q; qz++) {
      c << "      {  // qz = " << qz << "\n";
      c << "        const int p = pxy + " << qz * SZ << ";\n";
      c << "        const int pt = pt0 + " << qz * Q * Q << ";\n";
      for (size_t f = 0; f < plan->in_fields.size(); f++) {
        const B200GenField &fd = plan->in_fields[f];
        const string        sl = std::to_string(fd.slot);
        switch (fd.emode) {
          case B200_EVAL_NONE:
            for (int cc = 0; cc < fd.nc; cc++) {
              if (fd.qd_off >= 0)
                c << "        in_" << f << "[" << cc << "] = " << qd_ref(fd, cc, "le", "pt") << ";\n";
              else if (qf_ahead) {
                const string slot = "nx_" + std::to_string(f) + "_" + std::to_string(cc) + "_" + std::to_string(qz % depth);
                c << "        in_" << f << "[" << cc << "] = " << slot << ";\n";
                if (qz + depth < Q) c << "        " << slot << " = " << none_load(fd, cc, "pt + " + std::to_string(depth * Q * Q)) << ";\n";
              } else
                c << "        in_" << f << "[" << cc << "] = " << none_load(fd, cc, "pt") << ";\n";
            }
            break;
          case B200_EVAL_WEIGHT: c << "        in_" << f << "[0] = wxy_" << f << " * cW" << fd.basis_id << "[" << qz << "];\n"; break;
          case B200_EVAL_INTERP: {
            const B200GenGroup &g = plan->in_groups[fd.group];
            for (int cc = 0; cc < fd.nc; cc++) {
              if (g.use_grad) c << "        in_" << f << "[" << cc << "] = uz_g" << fd.group << "c" << cc << "_" << qz << ";\n";
              else c << "        in_" << f << "[" << cc << "] = (" << plane(g.plane0 + cc, "le") << ")[p];\n";
            }
          } break;
          case B200_EVAL_GRAD: {
            const B200GenGroup &g = plan->in_groups[fd.group];
            for (int cc = 0; cc < fd.nc; cc++) {
              c << "        in_" << f << "[" << cc << "] = (" << plane(g.plane0 + 2 * g.nc + cc, "le") << ")[p];\n";
              c << "        in_" << f << "[" << cc + fd.nc << "] = (" << plane(g.plane0 + g.nc + cc, "le") << ")[p];\n";
              c << "        in_" << f << "[" << cc + 2 * fd.nc << "] = dz_g" << fd.group << "c" << cc << "_" << qz << ";\n";
            }
          } break;
        }
      }
      c << "        " << qf->kernel_name << "(b200a.ctx, 1, in, out);\n";
      // outputs: first group contributions (INTERP stores, then GRAD parts), then EVAL_NONE
      for (size_t gi = 0; gi < plan->out_groups.size(); gi++) {
        const B200GenGroup &g = plan->out_groups[gi];
        for (int cc = 0; cc < g.nc; cc++) {
          // value part
          string val;
          for (size_t f = 0; f < plan->out_fields.size(); f++) {
            const B200GenField &fd = plan->out_fields[f];
            if (fd.group == (int)gi && fd.emode == B200_EVAL_INTERP) val += (val.empty() ? "" : " + ") + ("out_" + std::to_string(f) + "[" + std::to_string(cc) + "]");
          }
          if (!val.empty()) c << "        (" << plane(g.plane0 + cc, "le") << ")[p] = " << val << ";\n";
          if (g.use_grad) {
            string vx, vy, vz;
            for (size_t f = 0; f < plan->out_fields.size(); f++) {
              const B200GenField &fd = plan->out_fields[f];
              if (fd.group == (int)gi && fd.emode == B200_EVAL_GRAD) {
                const string o = "out_" + std::to_string(f);
                vx += (vx.empty() ? "" : " + ") + o + "[" + std::to_string(cc) + "]";
                vy += (vy.empty() ? "" : " + ") + o + "[" + std::to_string(cc + fd.nc) + "]";
                vz += (vz.empty() ? "" : " + ") + o + "[" + std::to_string(cc + 2 * fd.nc) + "]";
              }
            }
            c << "        (" << plane(g.plane0 + 2 * g.nc + cc, "le") << ")[p] = " << vx << ";\n";
            c << "        (" << plane(g.plane0 + g.nc + cc, "le") << ")[p] = " << vy << ";\n";
            c << "        { const double vz = " << vz << ";\n";
            for (int m = 0; m < Q; m++)
              c << "          vz_g" << gi << "c" << cc << "_" << m << " = fma(cG" << g.basis_id << "[" << qz * Q + m << "], vz, vz_g" << gi << "c" << cc
                << "_" << m << ");\n";
            c << "        }\n";
          }
        }
      }
      for (size_t f = 0; f < plan->out_fields.size(); f++) {
        const B200GenField &fd = plan->out_fields[f];
        if (fd.emode != B200_EVAL_NONE) continue;
        emit_scatter_value(fd.rstr, fd.slot, "e", "pt", fd.nc, [&](int cc) { return "out_" + std::to_string(f) + "[" + std::to_string(cc) + "]"; },
                           "        ");
      }
      c << "      }\n";
    }
    // fold the z-part of the transposed gradient into Vq
    for (size_t gi = 0; gi < plan->out_groups.size(); gi++) {
      const B200GenGroup &g = plan->out_groups[gi];
      if (!g.use_grad) continue;
      for (int cc = 0; cc < g.nc; cc++) {
        c << "      { double *vq = " << plane(g.plane0 + cc, "le") << " + pxy;\n";
        for (int m = 0; m < Q; m++)
          c << "        vq[" << m * SZ << "] " << (g.use_interp ? "+=" : "=") << " vz_g" << gi << "c" << cc << "_" << m << ";\n";
        c << "      }\n";
      }
    }
    c << "      }\n";
    task_loop_end();
  }

  // ---- x-line stage for operators without gradients: X contraction, QFunction on the line, X^T contraction --------------
  void emit_xline_qf() {
    B200QFunction qf = op->qf;
    comment("x-lines: x-contraction, QFunction on the Q points of the line, x-contraction^T (all components of one line per lane)");
    task_loop_begin(std::to_string(E * Q * LQ));
    c << "      const int qy = t % " << LQ << ", qz = (t / " << LQ << ") % " << Q << ", le = t / " << Q * LQ << ";\n";
    if (LQ != Q) c << "      if (qy >= " << Q << ") continue;\n";
    c << "      const int row = qz * " << Q << " + qy;\n";
    c << "      const long long e_real = e0 + le;\n";
    c << "      const long long e = e_real < b200_ne ? e_real : b200_ne - 1;\n      (void)qy; (void)qz; (void)e;\n";
    c << "      const CeedScalar *in[" << std::max<size_t>(1, qf->inputs.size()) << "];\n";
    c << "      CeedScalar *out[" << std::max<size_t>(1, qf->outputs.size()) << "];\n";
    for (size_t f = 0; f < plan->in_fields.size(); f++) c << "      CeedScalar in_" << f << "[" << plan->in_fields[f].size * Q << "];\n";
    for (size_t f = 0; f < plan->out_fields.size(); f++) c << "      CeedScalar out_" << f << "[" << plan->out_fields[f].size * Q << "];\n";
    for (size_t f = 0; f < plan->in_fields.size(); f++) c << "      in[" << f << "] = in_" << f << ";\n";
    for (size_t f = 0; f < plan->out_fields.size(); f++) c << "      out[" << f << "] = out_" << f << ";\n";
    // streamed inputs first (their latency overlaps the contractions below), then weights, then the interpolated fields
    for (size_t f = 0; f < plan->in_fields.size(); f++) {
      const B200GenField &fd = plan->in_fields[f];
      const string        sl = std::to_string(fd.slot);
      if (fd.emode == B200_EVAL_NONE) {
        for (int cc = 0; cc < fd.nc; cc++)
          for (int q = 0; q < Q; q++) {
            if (fd.qd_off >= 0)  // staged one group ahead by cp.async (b200_issue_qd)
              c << "      in_" << f << "[" << cc * Q + q << "] = " << qd_ref(fd, cc, "le", "row * " + std::to_string(Q) + " + " + std::to_string(q)) << ";\n";
            else
              c << "      in_" << f << "[" << cc * Q + q << "] = " + QLD + "(b200a.in_ptr[" << sl << "] + "
                << lidx(fd.rstr, "", "e", "row * " + std::to_string(Q) + " + " + std::to_string(q), std::to_string(cc)) << ");\n";
          }
      } else if (fd.emode == B200_EVAL_WEIGHT) {
        c << "      { const double wyz = cW" << fd.basis_id << "[qy] * cW" << fd.basis_id << "[qz];\n";
        for (int q = 0; q < Q; q++) c << "        in_" << f << "[" << q << "] = cW" << fd.basis_id << "[" << q << "] * wyz; }\n";
      }
    }
    for (size_t f = 0; f < plan->in_fields.size(); f++) {
      const B200GenField &fd = plan->in_fields[f];
      if (fd.emode != B200_EVAL_INTERP) continue;
      const B200GenGroup &g = plan->in_groups[fd.group];
      const B200GenBasis &b = basis(g.basis_id);
      const int           P = b.P, Ps = odd_pad(P);
      for (int cc = 0; cc < fd.nc; cc++) {
        c << "      { const double *src = " << plane(g.plane0 + g.nc + cc, "le") << " + qz * " << sz2(b) << " + qy * " << Ps << ";\n";
        for (int i = 0; i < P; i++) c << "        const double u" << i << " = src[" << i << "];\n";
        contract("cB" + std::to_string(g.basis_id), P, Q, false, "u", "r", "        ");
        for (int q = 0; q < Q; q++) c << "        in_" << f << "[" << cc * Q + q << "] = r" << q << ";\n";
        c << "      }\n";
      }
    }
    c << "      " << qf->kernel_name << "(b200a.ctx, " << Q << ", in, out);\n";
    for (size_t gi = 0; gi < plan->out_groups.size(); gi++) {
      const B200GenGroup &g = plan->out_groups[gi];
      const B200GenBasis &b = basis(g.basis_id);
      const int           P = b.P, Ps = odd_pad(P);
      for (int cc = 0; cc < g.nc; cc++) {
        c << "      {\n";
        for (int q = 0; q < Q; q++) {
          string val;
          for (size_t f = 0; f < plan->out_fields.size(); f++) {
            const B200GenField &fd = plan->out_fields[f];
            if (fd.group == (int)gi && fd.emode == B200_EVAL_INTERP) val += (val.empty() ? "" : " + ") + ("out_" + std::to_string(f) + "[" + std::to_string(cc * Q + q) + "]");
          }
          c << "        const double v" << q << " = " << val << ";\n";
        }
        contract("cB" + std::to_string(g.basis_id), Q, P, true, "v", "r", "        ");
        c << "        double *dst = " << plane(g.plane0 + g.nc + cc, "le") << " + qz * " << sz2(b) << " + qy * " << Ps << ";\n";
        for (int i = 0; i < P; i++) c << "        dst[" << i << "] = r" << i << ";\n";
        c << "      }\n";
      }
    }
    task_loop_end();
  }

  // ---- z-line QFunction stage with the per-lane quadrature-data ring ------------------------------
  // The streamed EVAL_NONE inputs are the bulk of the HBM traffic.  Every lane prefetches exactly the values it will consume
  // itself with cp.async into a small ring in shared memory: step s = (task round r, z-layer qz); ring_slots - 1 steps are
  // always in flight, also while the warp runs its contraction stages and across element groups (the last steps of a group
  // issue the first steps of the NEXT group this warp will process).  No registers are held by loads in flight and no
  // synchronisation is needed (a lane only reads what it copied itself; cp.async.wait_group orders it).
  struct RingStep { int round, qz; bool next; };
  void emit_ring_bases(const string &ind) {
    // per task round: clamped task decode and base pointers of the current (qc) and the next (qn) element group
    const int ntasks = E * Q * Q;
    for (int r = 0; r < plan->ring_rounds; r++) {
      const string x = "_" + std::to_string(r);
      c << ind << "const int t" << x << " = lane + " << r * TS << ", tc" << x << " = t" << x << " < " << ntasks << " ? t" << x << " : " << ntasks - 1 << ";\n";
      c << ind << "const int le" << x << " = tc" << x << " / " << Q * Q << ", pt0" << x << " = tc" << x << " % " << Q * Q << ";\n";
      c << ind << "const long long ec" << x << " = (e0 + le" << x << " < b200_ne) ? e0 + le" << x << " : b200_ne - 1;\n";
      c << ind << "const long long en" << x << " = (e0n + le" << x << " < b200_ne) ? e0n + le" << x << " : b200_ne - 1;\n";
      for (size_t f = 0; f < plan->in_fields.size(); f++) {
        const B200GenField &fd = plan->in_fields[f];
        if (fd.ring_k < 0) continue;
        const string sl = std::to_string(fd.slot);
        c << ind << "const double *qc" << f << x << " = b200a.in_ptr[" << sl << "] + " << lidx(fd.rstr, "", "ec" + x, "pt0" + x, "0") << ";\n";
        c << ind << "const double *qn" << f << x << " = b200a.in_ptr[" << sl << "] + " << lidx(fd.rstr, "", "en" + x, "pt0" + x, "0") << ";\n";
      }
    }
  }
  void emit_ring_issue(const RingStep &st, int slot, const string &ind) {
    const int    slot_doubles = TS * plan->ring_comps;
    const string x            = "_" + std::to_string(st.round);
    for (size_t f = 0; f < plan->in_fields.size(); f++) {
      const B200GenField &fd = plan->in_fields[f];
      if (fd.ring_k < 0) continue;
      for (int cc = 0; cc < fd.nc; cc++) {
        const long long goff = (long long)cc * fd.rstr->strides[1] + (long long)st.qz * Q * Q * fd.rstr->strides[0];
        c << ind << "b200_cp8(ring + " << slot * slot_doubles + (fd.ring_k + cc) * TS << " + lane, " << (st.next ? "qn" : "qc") << f << x << " + " << goff << "LL);\n";
      }
    }
    c << ind << "b200_cp_commit();\n";
  }
  RingStep ring_step(int s) const {
    const int S = plan->ring_rounds * Q;
    RingStep  st;
    st.next = s >= S;
    if (st.next) s -= S;
    st.round = s / Q;
    st.qz    = s % Q;
    return st;
  }
  // prologue (before the batch loop): the first D steps of the first group of this warp
  void emit_ring_prologue() {
    c << "static __device__ __noinline__ void b200_ring_prologue(const long long e0) {\n";
    c << smw_decl();
    c << "  const int lane = " << TID << ";\n  const long long e0n = e0;\n";
    c << "  double *const ring = " << smem_at(plan->ring_off, "double") << ";\n";
    emit_ring_bases("  ");
    for (int s = 0; s < plan->ring_slots - 1; s++) emit_ring_issue(ring_step(s), s % plan->ring_slots, "  ");
    c << "}\n\n";
  }

  void emit_qf_stage_ring() {
    B200QFunction qf     = op->qf;
    const int     D      = plan->ring_slots - 1, NS = plan->ring_slots;
    const int     ntasks = E * Q * Q, S = plan->ring_rounds * Q;
    const int     slot_doubles = TS * plan->ring_comps;
    comment("quadrature points: one z-line per thread; d/dz, QFunction, (d/dz)^T in registers; quadrature data through the cp.async ring");
    const string name = "b200_stage_" + std::to_string(n_stage++);
    c << "static __device__ __noinline__ void " << name << "(const long long e0, const long long e0n) {\n";
    c << smw_decl();
    calls.push_back("    " + name + "(e0, e0n);\n");
    c << "    const int lane = " << TID << ";\n";
    c << "    double *const ring = " << smem_at(plan->ring_off, "double") << ";\n";
    emit_ring_bases("    ");
    c << "    const CeedScalar *in[" << std::max<size_t>(1, qf->inputs.size()) << "];\n";
    c << "    CeedScalar *out[" << std::max<size_t>(1, qf->outputs.size()) << "];\n";
    for (size_t f = 0; f < plan->in_fields.size(); f++) c << "    CeedScalar in_" << f << "[" << plan->in_fields[f].size << "];\n";
    for (size_t f = 0; f < plan->out_fields.size(); f++) c << "    CeedScalar out_" << f << "[" << plan->out_fields[f].size << "];\n";
    for (size_t f = 0; f < plan->in_fields.size(); f++) c << "    in[" << f << "] = in_" << f << ";\n";
    for (size_t f = 0; f < plan->out_fields.size(); f++) c << "    out[" << f << "] = out_" << f << ";\n";
    for (int r = 0; r < plan->ring_rounds; r++) {
      const string x = "_" + std::to_string(r);
      c << "    {  // task round " << r << "\n";
      c << "      const int le = le" << x << ", qx = pt0" << x << " % " << Q << ", qy = pt0" << x << " / " << Q << ";\n";
      c << "      const long long e = ec" << x << ";\n";
      c << "      const bool act = t" << x << " < " << ntasks << " && e0 + le < b200_ne;\n";
      c << "      const int pxy = qy * " << Qs << " + qx, pt0 = pt0" << x << ";\n";
      c << "      (void)e; (void)pt0;\n";
      for (size_t gi = 0; gi < plan->in_groups.size(); gi++) {
        const B200GenGroup &g = plan->in_groups[gi];
        if (!g.use_grad) continue;
        for (int cc = 0; cc < g.nc; cc++) {
          const string tag = "g" + std::to_string(gi) + "c" + std::to_string(cc) + "_";
          c << "      const double *uq_" << tag << " = " << plane(g.plane0 + cc, "le") << " + pxy;\n";
          for (int m = 0; m < Q; m++) c << "      const double uz_" << tag << m << " = uq_" << tag << "[" << m * Q * Qs << "];\n";
          contract("cG" + std::to_string(g.basis_id), Q, Q, false, "uz_" + tag, "dz_" + tag, "      ");
        }
      }
      for (size_t gi = 0; gi < plan->out_groups.size(); gi++) {
        const B200GenGroup &g = plan->out_groups[gi];
        if (!g.use_grad) continue;
        for (int cc = 0; cc < g.nc; cc++)
          for (int m = 0; m < Q; m++) c << "      double vz_g" << gi << "c" << cc << "_" << m << " = 0.0;\n";
      }
      for (size_t f = 0; f < plan->in_fields.size(); f++)
        if (plan->in_fields[f].emode == B200_EVAL_WEIGHT) {
          const string w = "cW" + std::to_string(plan->in_fields[f].basis_id);
          c << "      const double wxy_" << f << " = " << w << "[qx] * " << w << "[qy];\n";
        }
      for (int qz = 0; qz < Q; qz++) {
        const int s = r * Q + qz;
        c << "      {  // step " << s << ": qz = " << qz << "\n";
        c << "        asm volatile(\"cp.async.wait_group " << D - 1 << ";\" ::: \"memory\");\n";
        c << "        const int p = pxy + " << qz * Q * Qs << ";\n";
        c << "        const int pt = pt0 + " << qz * Q * Q << ";\n        (void)pt;\n";
        // ring values into registers first (the slot is re-used D steps later), then refill the ring, then compute
        for (size_t f = 0; f < plan->in_fields.size(); f++) {
          const B200GenField &fd = plan->in_fields[f];
          if (fd.ring_k < 0) continue;
          for (int cc = 0; cc < fd.nc; cc++)
            c << "        in_" << f << "[" << cc << "] = ring[" << (s % NS) * slot_doubles + (fd.ring_k + cc) * TS << " + lane];\n";
        }
        emit_ring_issue(ring_step(s + D), (s + D) % NS, "        ");
        c << "        if (act) {\n";
        for (size_t f = 0; f < plan->in_fields.size(); f++) {
          const B200GenField &fd = plan->in_fields[f];
          const string        sl = std::to_string(fd.slot);
          switch (fd.emode) {
            case B200_EVAL_NONE:
              if (fd.ring_k >= 0) break;
              for (int cc = 0; cc < fd.nc; cc++) {
                if (fd.qd_off >= 0)
                  c << "        in_" << f << "[" << cc << "] = " << qd_ref(fd, cc, "le", "pt") << ";\n";
                else
                  c << "        in_" << f << "[" << cc << "] = " + QLD + "(b200a.in_ptr[" << sl << "] + " << lidx(fd.rstr, "b200a.in_idx[" + sl + "]", "e", "pt", std::to_string(cc)) << ");\n";
              }
              break;
            case B200_EVAL_WEIGHT: c << "        in_" << f << "[0] = wxy_" << f << " * cW" << fd.basis_id << "[" << qz << "];\n"; break;
            case B200_EVAL_INTERP: {
              const B200GenGroup &g = plan->in_groups[fd.group];
              for (int cc = 0; cc < fd.nc; cc++) {
                if (g.use_grad) c << "        in_" << f << "[" << cc << "] = uz_g" << fd.group << "c" << cc << "_" << qz << ";\n";
                else c << "        in_" << f << "[" << cc << "] = (" << plane(g.plane0 + cc, "le") << ")[p];\n";
              }
            } break;
            case B200_EVAL_GRAD: {
              const B200GenGroup &g = plan->in_groups[fd.group];
              for (int cc = 0; cc < fd.nc; cc++) {
                c << "        in_" << f << "[" << cc << "] = (" << plane(g.plane0 + 2 * g.nc + cc, "le") << ")[p];\n";
                c << "        in_" << f << "[" << cc + fd.nc << "] = (" << plane(g.plane0 + g.nc + cc, "le") << ")[p];\n";
                c << "        in_" << f << "[" << cc + 2 * fd.nc << "] = dz_g" << fd.group << "c" << cc << "_" << qz << ";\n";
              }
            } break;
          }
        }
        c << "        " << qf->kernel_name << "(b200a.ctx, 1, in, out);\n";
        for (size_t gi = 0; gi < plan->out_groups.size(); gi++) {
          const B200GenGroup &g = plan->out_groups[gi];
          for (int cc = 0; cc < g.nc; cc++) {
            string val;
            for (size_t f = 0; f < plan->out_fields.size(); f++) {
              const B200GenField &fd = plan->out_fields[f];
              if (fd.group == (int)gi && fd.emode == B200_EVAL_INTERP) val += (val.empty() ? "" : " + ") + ("out_" + std::to_string(f) + "[" + std::to_string(cc) + "]");
            }
            if (!val.empty()) c << "        (" << plane(g.plane0 + cc, "le") << ")[p] = " << val << ";\n";
            if (g.use_grad) {
              string vx, vy, vz;
              for (size_t f = 0; f < plan->out_fields.size(); f++) {
                const B200GenField &fd = plan->out_fields[f];
                if (fd.group == (int)gi && fd.emode == B200_EVAL_GRAD) {
                  const string o = "out_" + std::to_string(f);
                  vx += (vx.empty() ? "" : " + ") + o + "[" + std::to_string(cc) + "]";
                  vy += (vy.empty() ? "" : " + ") + o + "[" + std::to_string(cc + fd.nc) + "]";
                  vz += (vz.empty() ? "" : " + ") + o + "[" + std::to_string(cc + 2 * fd.nc) + "]";
                }
              }
              c << "        (" << plane(g.plane0 + 2 * g.nc + cc, "le") << ")[p] = " << vx << ";\n";
              c << "        (" << plane(g.plane0 + g.nc + cc, "le") << ")[p] = " << vy << ";\n";
              c << "        { const double vz = " << vz << ";\n";
              for (int m = 0; m < Q; m++)
                c << "          vz_g" << gi << "c" << cc << "_" << m << " = fma(cG" << g.basis_id << "[" << qz * Q + m << "], vz, vz_g" << gi << "c" << cc
                  << "_" << m << ");\n";
              c << "        }\n";
            }
          }
        }
        for (size_t f = 0; f < plan->out_fields.size(); f++) {
          const B200GenField &fd = plan->out_fields[f];
          if (fd.emode != B200_EVAL_NONE) continue;
          emit_scatter_value(fd.rstr, fd.slot, "e", "pt", fd.nc, [&](int cc) { return "out_" + std::to_string(f) + "[" + std::to_string(cc) + "]"; },
                             "        ");
        }
        c << "        }\n";  // act
        c << "      }\n";    // step
      }
      c << "      if (act) {\n";
      for (size_t gi = 0; gi < plan->out_groups.size(); gi++) {
        const B200GenGroup &g = plan->out_groups[gi];
        if (!g.use_grad) continue;
        for (int cc = 0; cc < g.nc; cc++) {
          c << "      { double *vq = " << plane(g.plane0 + cc, "le") << " + pxy;\n";
          for (int m = 0; m < Q; m++)
            c << "        vq[" << m * Q * Qs << "] " << (g.use_interp ? "+=" : "=") << " vz_g" << gi << "c" << cc << "_" << m << ";\n";
          c << "      }\n";
        }
      }
      c << "      }\n";
      c << "    }\n";  // round
    }
    (void)S;
    c << "}\n\n";
  }

  // ---- output side ------------------------------------------------------------------------------
  // Add `value(cc)` for E-entry n of element e into the L-vector of restriction r (output slot `slot`).
  template <typename F>
  void emit_scatter_value(B200Restriction r, int slot, const string &e, const string &n, int nc, F value, const string &ind) {
    const string sl = std::to_string(slot);
    if (r->is_strided) {
      for (int cc = 0; cc < nc; cc++)
        c << ind << "b200a.out_ptr[" << sl << "][" << lidx(r, "", e, n, std::to_string(cc)) << "] " << (add ? "+=" : "=") << " " << value(cc) << ";\n";
      return;
    }
    const long long e_entries = (long long)r->num_elem * r->elem_size;
    switch (plan->scatter_mode) {
      case B200_SCATTER_ATOMIC:
        for (int cc = 0; cc < nc; cc++)
          c << ind << "atomicAdd(b200a.out_ptr[" << sl << "] + " << lidx(r, "b200a.out_idx[" + sl + "]", e, n, std::to_string(cc)) << ", " << value(cc) << ");\n";
        break;
      case B200_SCATTER_EVECTOR:
        for (int cc = 0; cc < nc; cc++)
          c << ind << "b200a.out_aux[" << sl << "][(" << e << ") * " << r->elem_size << "LL + (" << n << ") + " << cc * e_entries << "LL] = " << value(cc)
            << ";\n";
        break;
      default:
        c << ind << "{ const int tg = b200a.out_idx[" << sl << "][(" << e << ") * " << r->elem_size << "LL + (" << n << ")];\n";
        c << ind << "  if (tg >= 0) {\n";
        for (int cc = 0; cc < nc; cc++)
          c << ind << "    b200a.out_ptr[" << sl << "][tg + " << (long long)cc * r->comp_stride << "LL] " << (add ? "+=" : "=") << " " << value(cc) << ";\n";
        c << ind << "  } else {\n";
        for (int cc = 0; cc < nc; cc++)
          c << ind << "    b200a.out_aux[" << sl << "][(long long)(~tg) + " << (long long)cc * r->num_halo << "LL] = " << value(cc) << ";\n";
        c << ind << "  } }\n";
    }
  }

  void emit_gradT_y(const B200GenGroup &g) {
    if (!g.use_grad) return;
    comment("(d/dy)^T, output group slot " + std::to_string(g.slot));
    task_loop_begin(std::to_string(E * g.nc * Q * LQ));
    c << "      const int qx = t % " << LQ << ", qz = (t / " << LQ << ") % " << Q << ", cc = (t / " << Q * LQ << ") % " << g.nc << ", le = t / "
      << Q * LQ * g.nc << ";\n";
    if (LQ != Q) c << "      if (qx >= " << Q << ") continue;\n";
    c << "      const double *src = " << plane(g.plane0 + g.nc, "le") << " + cc * " << E * S << " + qz * " << SZ << (swz ? "" : " + qx") << ";\n";
    c << "      double *vq = " << plane(g.plane0, "le") << " + cc * " << E * S << " + qz * " << SZ << (swz ? "" : " + qx") << ";\n";
    auto at = [&](int m) { return swz ? std::to_string(m * QP) + " + (qx ^ " + std::to_string(m) + ")" : std::to_string(m * QP); };
    for (int m = 0; m < Q; m++) c << "      const double u" << m << " = src[" << at(m) << "];\n";
    contract("cG" + std::to_string(g.basis_id), Q, Q, true, "u", "d", "      ");
    for (int q = 0; q < Q; q++) c << "      vq[" << at(q) << "] += d" << q << ";\n";
    task_loop_end();
  }

  void emit_interpT_x(const B200GenGroup &g) {
    const B200GenBasis &b = basis(g.basis_id);
    const int           P = b.P, Ps = odd_pad(P);
    if (b.collocated && !g.use_grad) return;
    comment("(d/dx)^T + x-contraction^T, output group slot " + std::to_string(g.slot));
    task_loop_begin(std::to_string(E * g.nc * Q * LQ));
    c << "      const int qy = t % " << LQ << ", qz = (t / " << LQ << ") % " << Q << ", cc = (t / " << Q * LQ << ") % " << g.nc << ", le = t / " << Q * LQ * g.nc
      << ";\n";
    if (LQ != Q) c << "      if (qy >= " << Q << ") continue;\n";
    c << "      double *vq = " << plane(g.plane0, "le") << " + cc * " << E * S << " + qz * " << SZ << " + qy * " << QP << ";\n";
    if (g.use_grad) {
      c << "      const double *vx = " << plane(g.plane0 + 2 * g.nc, "le") << " + cc * " << E * S << " + qz * " << SZ << " + qy * " << QP << ";\n";
      row_load("vx", "u", "qy", "      ");
      contract("cG" + std::to_string(g.basis_id), Q, Q, true, "u", "d", "      ");
      row_load("vq", "w", "qy", "      ");
      for (int q = 0; q < Q; q++) c << "      const double v" << q << " = w" << q << " + d" << q << ";\n";
    } else {
      row_load("vq", "v", "qy", "      ");
    }
    if (b.collocated) {
      row_store("vq", "v", "qy", "      ");
    } else {
      c << "      double *dst = " << plane(g.plane0 + g.nc, "le") << " + cc * " << E * S << " + qz * " << sz2(b) << " + qy * " << Ps << ";\n";
      contract("cB" + std::to_string(g.basis_id), Q, P, true, "v", "r", "      ");
      for (int i = 0; i < P; i++) c << "      dst[" << i << "] = r" << i << ";\n";
    }
    task_loop_end();
  }

  void emit_interpT_y(const B200GenGroup &g) {
    const B200GenBasis &b = basis(g.basis_id);
    if (b.collocated) return;
    const int P = b.P, Ps = odd_pad(P);
    // T1' goes to plane C when the group has a gradient (plane A still holds nothing live, but keep A/B/C rotation simple)
    const int t1_plane = g.plane0;  // Vq is dead after the x-stage
    comment("y-contraction^T, output group slot " + std::to_string(g.slot));
    const int LP = lp(b);
    task_loop_begin(std::to_string(E * g.nc * Q * LP));
    c << "      const int i = t % " << LP << ", qz = (t / " << LP << ") % " << Q << ", cc = (t / " << LP * Q << ") % " << g.nc << ", le = t / "
      << LP * Q * g.nc << ";\n";
    if (LP != P) c << "      if (i >= " << P << ") continue;\n";
    c << "      const double *src = " << plane(g.plane0 + g.nc, "le") << " + cc * " << E * S << " + qz * " << sz2(b) << " + i;\n";
    c << "      double *dst = " << plane(t1_plane, "le") << " + cc * " << E * S << " + qz * " << sz1(b) << " + i;\n";
    for (int q = 0; q < Q; q++) c << "      const double u" << q << " = src[" << q * Ps << "];\n";
    contract("cB" + std::to_string(g.basis_id), Q, P, true, "u", "r", "      ");
    for (int j = 0; j < P; j++) c << "      dst[" << j * P << "] = r" << j << ";\n";
    task_loop_end();
  }

  // Z^T + scatter with in-kernel ordered completion of shared nodes (B200_SCATTER_ORDERED), see B200OrderedScatter.
  // Phase A (this group): plain nodes are stored, every contribution to a shared node goes to its halo slot; then the group
  // publishes its flag.  Phase B runs ONE ITERATION LATER (b200_ordered_complete for the group this warp processed before):
  // by then the predecessor groups have long published, so the flag wait is free; the completing group adds the slots of
  // each shared node it holds the last E-entry of in ascending E-order and stores the sum.
  void emit_scatter_z_ordered(const B200GenGroup &g) {
    const B200GenBasis &b  = basis(g.basis_id);
    const int           P  = b.P;
    const string        sl = std::to_string(g.slot);
    const long long     cs = (long long)g.rstr->comp_stride;
    const int           ntasks = E * g.nc * P * P, es = g.rstr->elem_size;
    // ---- completion of a previously scattered group
    comment("ordered scatter, phase B: complete the shared nodes whose last E-entry lies in the group starting at element ep");
    c << "static __device__ __noinline__ void b200_ordered_complete(const long long ep, const int epoch) {\n";
    c << "    const long long grp = ep / " << E << ";\n";
    c << "    for (int i = b200a.ord_pred_ptr[grp] + " << TID << "; i < b200a.ord_pred_ptr[grp + 1]; i += " << TS << ") {\n";
    c << "      const int *flag = b200a.ord_flags + b200a.ord_pred_idx[i];\n      int seen;\n";
    c << "      do { asm volatile(\"ld.acquire.gpu.global.s32 %0, [%1];\" : \"=r\"(seen) : \"l\"(flag) : \"memory\"); } while (seen != epoch);\n";
    c << "    }\n";
    c << "    " << SYNC << "\n";
    c << "    const int nent = (int)((b200_ne - ep < " << E << ") ? b200_ne - ep : " << E << ") * " << es << ";\n";
    c << "    const int *tgt = b200a.out_idx[" << sl << "] + ep * " << es << "LL;\n";
    // U entries per lane at a time: all table loads first, then the first two contributions (every shared node has >= 2), ...
    const int U = 4;
    c << "    for (int t0 = " << TID << "; t0 < nent; t0 += " << U * TS << ") {\n";
    for (int u = 0; u < U; u++) c << "      const int x" << u << " = (t0 + " << u * TS << " < nent) ? ~__ldg(tgt + t0 + " << u * TS << ") : 0;\n";
    for (int u = 0; u < U; u++) {
      c << "      const bool last" << u << " = x" << u << " > 0 && ((x" << u << " >> 27) & 1);\n";
      c << "      const int cnt" << u << " = ((x" << u << " >> 28) & 7) + 2, s" << u << " = (x" << u << " & 0x7ffffff) - (cnt" << u << " - 1);  // first contribution\n";
    }
    for (int cc = 0; cc < g.nc; cc++) {
      c << "      { const double *h = b200a.out_aux[" << sl << "] + " << cc << " * b200a.ord_num_halo;\n";
      for (int u = 0; u < U; u++)
        c << "        double a" << u << " = 0.0, b" << u << " = 0.0; if (last" << u << ") { a" << u << " = __ldcg(h + s" << u << "); b" << u << " = __ldcg(h + s" << u
          << " + 1); }\n";
      for (int u = 0; u < U; u++) {
        c << "        if (last" << u << ") {\n";
        c << "          double acc = a" << u << " + b" << u << ";\n";
        c << "          for (int k = 2; k < cnt" << u << "; k++) acc += __ldcg(h + s" << u << " + k);\n";
        c << "          const long long node = __ldcg((const long long *)b200a.out_aux[" << sl << "] + s" << u << " - 1);\n";
        c << "          b200a.out_ptr[" << sl << "][node + " << cc * cs << "LL] " << (add ? "+=" : "=") << " acc;\n        }\n";
      }
      c << "      }\n";
    }
    c << "    }\n}\n\n";
    // ---- phase A
    comment("z-contraction^T + ordered scatter, phase A (plain stores + halo slots + flag), slot " + sl);
    const string name = "b200_stage_" + std::to_string(n_stage++);
    c << "static __device__ __noinline__ void " << name << "(const long long e0, const int epoch) {\n";
    c << smw_decl();
    calls.push_back("    " + name + "(e0, epoch);\n    if (e0p >= 0) b200_ordered_complete(e0p, epoch);\n    e0p = e0;\n");
    c << "    for (int t = " << TID << "; t < " << ntasks << "; t += " << TS << ") {\n";
    c << "      const int ij = t % " << P * P << ", cc = (t / " << P * P << ") % " << g.nc << ", le = t / " << P * P * g.nc << ";\n";
    c << "      const long long e = e0 + le;\n";
    if (b.collocated) {
      c << "      const double *col = " << plane(g.plane0, "le") << " + cc * " << E * S << " + (ij / " << P << ") * " << Qs << " + (ij % " << P << ");\n";
      for (int k = 0; k < P; k++) c << "      const double r" << k << " = col[" << k * Q * Qs << "];\n";
    } else {
      c << "      const double *col = " << plane(g.plane0, "le") << " + cc * " << E * S << " + ij;\n";
      for (int q = 0; q < Q; q++) c << "      const double u" << q << " = col[" << q * P * P << "];\n";
      contract("cB" + std::to_string(g.basis_id), Q, P, true, "u", "r", "      ");
    }
    c << "      if (e < b200_ne) {\n";
    for (int k = 0; k < P; k++) {
      const string n = "ij + " + std::to_string(k * P * P);
      string       tgt;
      if (g.tgt_off >= 0) tgt = smem_at(g.tgt_off, "const int") + "[le * " + std::to_string(es) + " + " + n + "]";
      else tgt = "b200a.out_idx[" + sl + "][e * " + std::to_string(es) + "LL + (" + n + ")]";
      c << "        { const int tg = " << tgt << ";\n";
      c << "          if (tg >= 0) b200a.out_ptr[" << sl << "][tg + cc * " << cs << "LL] " << (add ? "+=" : "=") << " r" << k << ";\n";
      c << "          else b200a.out_aux[" << sl << "][(long long)((~tg) & 0x7ffffff) + cc * b200a.ord_num_halo] = r" << k << "; }\n";
    }
    c << "      }\n    }\n";
    // one fence + release by one lane after the group barrier publishes the stores of the whole group (cumulativity)
    c << "    " << SYNC << "\n";
    c << "    if (" << TID << " == 0) {\n      __threadfence();\n      asm volatile(\"st.release.gpu.global.s32 [%0], %1;\" ::\"l\"(b200a.ord_flags + e0 / " << E
      << "), \"r\"(epoch) : \"memory\");\n    }\n";
    c << "}\n\n";
  }

  void emit_scatter_z(const B200GenGroup &g) {
    if (plan->scatter_mode == B200_SCATTER_ORDERED && g.slot == plan->ordered_slot) return emit_scatter_z_ordered(g);
    const B200GenBasis &b = basis(g.basis_id);
    const int           P = b.P;
    comment("z-contraction^T + scatter, output group slot " + std::to_string(g.slot));
    task_loop_begin(std::to_string(E * g.nc * P * P));
    c << "      const int ij = t % " << P * P << ", cc = (t / " << P * P << ") % " << g.nc << ", le = t / " << P * P * g.nc << ";\n";
    c << "      const long long e = e0 + le;\n";
    if (b.collocated) {
      c << "      const double *src = " << plane(g.plane0, "le") << " + cc * " << E * S << " + (ij / " << P << ") * " << QP << " + "
        << xq("ij % " + std::to_string(P), "ij / " + std::to_string(P)) << ";\n";
      for (int k = 0; k < P; k++) c << "      const double r" << k << " = src[" << k * SZ << "];\n";
    } else {
      c << "      const double *src = " << plane(g.plane0, "le") << " + cc * " << E * S << " + ij;\n";
      for (int q = 0; q < Q; q++) c << "      const double u" << q << " = src[" << q * sz1(b) << "];\n";
      contract("cB" + std::to_string(g.basis_id), Q, P, true, "u", "r", "      ");
    }
    c << "      if (e < b200_ne) {\n";
    // component handled through the (runtime) cc: emit with nc == 1 semantics and an explicit component offset
    for (int k = 0; k < P; k++) {
      const string n = "ij + " + std::to_string(k * P * P);
      // index of this E-entry: staged in shared memory by cp.async at the top of the batch, else read from global memory
      string index;
      if (g.tgt_off >= 0) index = smem_at(g.tgt_off, "const int") + "[le * " + std::to_string(g.rstr->elem_size) + " + " + n + "]";
      emit_scatter_comp(g.rstr, g.slot, "e", n, "cc", "r" + std::to_string(k), "        ", index);
    }
    c << "      }\n";
    task_loop_end();
  }

  // scatter of a single value for runtime component index `cc`; `index` (optional) is an expression for offsets/tgt[E-entry]
  void emit_scatter_comp(B200Restriction r, int slot, const string &e, const string &n, const string &cc, const string &val, const string &ind,
                         const string &index = "") {
    const string sl = std::to_string(slot);
    if (r->is_strided) {
      c << ind << "b200a.out_ptr[" << sl << "][" << lidx(r, "", e, n, cc) << "] " << (add ? "+=" : "=") << " " << val << ";\n";
      return;
    }
    const long long e_entries = (long long)r->num_elem * r->elem_size;
    const string    entry     = index.empty() ? "b200a.out_idx[" + sl + "][(" + e + ") * " + std::to_string(r->elem_size) + "LL + (" + n + ")]" : index;
    switch (plan->scatter_mode) {
      case B200_SCATTER_ATOMIC:
        c << ind << "atomicAdd(b200a.out_ptr[" << sl << "] + ((long long)" << entry << " + (long long)(" << cc << ") * " << r->comp_stride << "LL), " << val
          << ");\n";
        break;
      case B200_SCATTER_EVECTOR:
        c << ind << "b200a.out_aux[" << sl << "][(" << e << ") * " << r->elem_size << "LL + (" << n << ") + " << cc << " * " << e_entries << "LL] = " << val
          << ";\n";
        break;
      default:
        c << ind << "{ const int tg = " << entry << ";\n";
        c << ind << "  if (tg >= 0) b200a.out_ptr[" << sl << "][tg + " << cc << " * " << (long long)r->comp_stride << "LL] " << (add ? "+=" : "=") << " " << val
          << ";\n";
        c << ind << "  else b200a.out_aux[" << sl << "][(long long)(~tg) + " << cc << " * " << (long long)r->num_halo << "LL] = " << val << "; }\n";
    }
  }

  string generate() {
    Q  = plan->Q;
    Qs = plan->Qs;
    E  = plan->epb;
    NT = plan->threads;
    warp_mode = plan->warp_mode;
    TS        = warp_mode ? 32 * plan->group_warps : NT;
    TID       = !warp_mode ? "threadIdx.x" : ((TS & (TS - 1)) == 0 ? "(threadIdx.x & " + std::to_string(TS - 1) + ")" : "(threadIdx.x % " + std::to_string(TS) + ")");
    SMBASE    = warp_mode ? "smw" : "sm";
    // a group of one warp synchronises with __syncwarp(); wider groups meet at their own named barrier (ids 1..15)
    if (!warp_mode || NT == TS) SYNC = "__syncthreads();";
    else if (plan->group_warps == 1) SYNC = "__syncwarp();";
    else SYNC = "asm volatile(\"bar.sync %0, " + std::to_string(TS) + ";\" ::\"r\"((int)(threadIdx.x / " + std::to_string(TS) + ") + 1) : \"memory\");";
    S   = plan->plane_size;
    swz = plan->swz;
    W   = plan->swz_w;
    lin = plan->lin;
    QP  = swz ? W : (lin ? Q : Qs);
    SZ  = swz ? swz_sz(W, Q * W) : (lin ? lin_sz(Q) : Q * Qs);
    LQ  = swz ? W : Q;
    // quadrature data is read exactly once per apply: mark it evict-first so that u, v, the halo buffer and the offsets keep the L2
    if (plan->stage_mask & 128) QLD = "__ldcs";
    emit_header();
    bool any;
    // input side
    for (auto &g : plan->in_groups) emit_gather_z(g);
    if (!plan->in_groups.empty()) barrier();
    {
      // the offsets buffer has been consumed by the gather stage: stream in the next group's offsets
      bool any_idx = false;
      for (auto &g : plan->in_groups) any_idx = any_idx || g.idx_off >= 0;
      if (any_idx && !plan->in_groups.empty()) calls.push_back("    b200_issue_idx(e0n);\n    b200_cp_commit();\n");
    }
    any = false;
    for (auto &g : plan->in_groups) any = any || !basis(g.basis_id).collocated;
    if (any) {
      for (auto &g : plan->in_groups) emit_interp_y(g);
      barrier();
    }
    any = false;
    for (auto &g : plan->in_groups) any = any || !(basis(g.basis_id).collocated && !g.use_grad);
    if (any && !plan->qf_xline) {
      for (auto &g : plan->in_groups) emit_interp_x(g);
      barrier();
    }
    any = false;
    for (auto &g : plan->in_groups) any = any || g.use_grad;
    if (any) {
      for (auto &g : plan->in_groups) emit_grad_y(g);
      barrier();
    }
    // asynchronous staging: emit the issue functions (definitions go to the source stream, calls are placed below)
    const bool has_idx = emit_issue_idx(), has_tgt = emit_issue_tgt(), has_gather = emit_issue_gather(), has_qd = emit_issue_qd();
    const bool staged  = has_idx || has_tgt || has_gather || has_qd;
    if (has_gather) {
      // before the quadrature stage: the offsets of the next batch have landed (issued at the top of this batch) and the
      // gather buffer is free (the z-stage consumed it) -> start gathering the next batch's inputs
      if (!calls.empty() && calls.back() == "    " + SYNC + "\n") calls.pop_back();
      calls.push_back("    b200_cp_wait_all();\n    " + SYNC + "\n    b200_issue_gather(e0n);\n    b200_cp_commit();\n");
    } else if (has_tgt) {
      if (!calls.empty() && calls.back() == "    " + SYNC + "\n") calls.pop_back();
      calls.push_back("    b200_cp_wait_all();\n    " + SYNC + "\n");
    }
    // bulk-copied quadrature data of this batch: all lanes of the group wait on the mbarrier phase of this iteration
    if (plan->qd_tma) calls.push_back("    b200_mbar_wait(b200_smem_u32((char *)(" + smw_expr() + ") + " + std::to_string(plan->mbar_off) + "), it & 1);\n");
    if (plan->qf_xline) {
      emit_xline_qf();
    } else if (plan->qf_pointwise) {
      any = false;
      for (auto &g : plan->in_groups) any = any || g.use_grad;
      if (any) {
        for (auto &g : plan->in_groups) emit_grad_z(g);
        barrier();
      }
      emit_qf_points();
    } else if (plan->ring_off >= 0) {
      emit_ring_prologue();
      emit_qf_stage_ring();
    } else {
      emit_qf_stage();
    }
    if (!plan->out_groups.empty() || has_qd) barrier();
    // the quadrature-data buffer is free again: stream in the next batch's data behind the transpose stages
    if (has_qd) calls.push_back("    b200_issue_qd(e0n);\n    b200_cp_commit();\n");
    if (emit_prefetch_qd()) calls.push_back("    b200_prefetch_qd(e0n);\n");
    if (!plan->out_groups.empty()) {
      any = false;
      for (auto &g : plan->out_groups) any = any || g.use_grad;
      if (any && plan->qf_pointwise) {
        for (auto &g : plan->out_groups) emit_gradT_z(g);
        barrier();
      }
      if (any) {
        for (auto &g : plan->out_groups) emit_gradT_y(g);
        barrier();
      }
      any = false;
      for (auto &g : plan->out_groups) any = any || !(basis(g.basis_id).collocated && !g.use_grad);
      if (any && !plan->qf_xline) {
        for (auto &g : plan->out_groups) emit_interpT_x(g);
        barrier();
      }
      any = false;
      for (auto &g : plan->out_groups) any = any || !basis(g.basis_id).collocated;
      if (any) {
        for (auto &g : plan->out_groups) emit_interpT_y(g);
        barrier();
      }
      for (auto &g : plan->out_groups) emit_scatter_z(g);
    }
    if (!staged) barrier();  // staged kernels synchronise at the top of the batch loop instead
    const int minb = std::max(1, plan->blocks_per_sm);
    // L2 prefetch of the NEXT batch this block will process (grid-stride): the streamed quadrature data (the bulk of the
    // HBM traffic) and the element offsets, so that the demand loads a few microseconds later hit in the 126 MB L2.
    const bool prefetch = getenv("CEED_B200_PREFETCH") != nullptr;
    if (prefetch) {
      c << "static __device__ __noinline__ void b200_prefetch(const long long e0) {\n";
      c << "  if (e0 >= b200_ne) return;\n";
      c << "  const long long ne = (b200_ne - e0 < " << E << ") ? b200_ne - e0 : " << E << ";\n";
      c << smw_decl();
      auto emit_range = [&](const string &ptr, const string &first_elem_expr, long long bytes_per_elem) {
        // byte range [p0, p0 + ne * bytes_per_elem), one 128-byte line per thread per round
        c << "  { const char *p0 = (const char *)(" << ptr << ") + (" << first_elem_expr << ") * " << bytes_per_elem << "LL;\n";
        c << "    const long long nbytes = ne * " << bytes_per_elem << "LL;\n";
        c << "    for (long long off = (long long)" << TID << " * 128; off < nbytes; off += " << TS * 128 << ")\n";
        c << "      asm volatile(\"prefetch.global.L2 [%0];\" ::\"l\"(p0 + off));\n  }\n";
      };
      for (size_t f = 0; f < plan->in_fields.size(); f++) {
        const B200GenField &fd = plan->in_fields[f];
        if (fd.emode != B200_EVAL_NONE || !fd.rstr->is_strided || fd.rstr->strides[0] != 1) continue;
        // contiguous per (component, element) when the node stride is 1
        for (int cc = 0; cc < fd.nc; cc++) {
          if (fd.rstr->strides[2] == fd.rstr->elem_size) {
            emit_range("b200a.in_ptr[" + std::to_string(fd.slot) + "] + " + std::to_string((long long)cc * fd.rstr->strides[1]) + "LL", "e0", 8LL * fd.rstr->elem_size);
          }
        }
      }
      for (auto &g : plan->in_groups)
        if (!g.rstr->is_strided) emit_range("b200a.in_idx[" + std::to_string(g.slot) + "]", "e0", 4LL * g.rstr->elem_size);
      c << "}\n\n";
    }
    c << "extern \"C\" __global__ void __launch_bounds__(" << NT << ", " << minb << ") b200_operator_" << op->qf->kernel_name
      << "(const long long e_begin, const long long e_end) {\n";
    // the launch processes elements [e_begin, e_end): the whole mesh, or the boundary / interior part of a partitioned mesh
    c << "  if (threadIdx.x == 0) b200_ne = e_end;\n  __syncthreads();\n";
    c << "  const long long num_batches = (e_end - e_begin + " << E - 1 << ") / " << E << ";\n";
    if (plan->scatter_mode == B200_SCATTER_ORDERED)
      c << "  const int epoch = *(volatile const int *)b200a.ord_sync + 1;  // this launch\n  long long e0p = -1;  // group whose shared nodes are still to be completed\n";
    // block mode: one batch per CTA per iteration; warp mode: one group per warp per iteration
    const string first  = warp_mode ? "((long long)blockIdx.x * " + std::to_string(NT / TS) + " + (threadIdx.x / " + std::to_string(TS) + "))" : "(long long)blockIdx.x";
    const string stride = warp_mode ? "((long long)gridDim.x * " + std::to_string(NT / TS) + ")" : "(long long)gridDim.x";
    if (plan->qd_tma) {
      c << "  int it = 0;  // iteration of this group: parity of the mbarrier phase its bulk copies complete\n";
      c << "  if (" << TID << " == 0) b200_mbar_init(b200_smem_u32((char *)(" << smw_expr() << ") + " << plan->mbar_off << "), 1);\n  " << SYNC << "\n";
    }
    if (staged) {
      // prologue: stage the first batch
      c << "  if (" << first << " < num_batches) {\n";
      c << "    const long long e0 = e_begin + " << first << " * " << E << ";\n";
      if (has_idx) c << "    b200_issue_idx(e0);\n    b200_cp_commit();\n    b200_cp_wait_all();\n    " << SYNC << "\n";
      if (has_gather) c << "    b200_issue_gather(e0);\n";
      if (has_qd) c << "    b200_issue_qd(e0);\n";
      c << "    b200_cp_commit();\n  }\n";
    }
    if (plan->ring_off >= 0 && !plan->qf_pointwise) c << "  if (" << first << " < num_batches) b200_ring_prologue(e_begin + " << first << " * " << E << ");\n";
    c << "  for (long long batch = " << first << "; batch < num_batches; batch += " << stride << ") {\n";
    c << "    const long long e0 = e_begin + batch * " << E << ", e0n = e_begin + (batch + " << stride << ") * " << E << ";\n";
    if (staged) {
      // everything staged for this batch is visible after this point; all reads of the previous batch are done
      c << "    b200_cp_wait_all();\n    " << SYNC << "\n";
      if (has_tgt) c << "    b200_issue_tgt(e0);\n";
      c << "    b200_cp_commit();\n";
    } else {
      c << "    (void)e0n;\n";
    }
    if (prefetch) c << "    b200_prefetch(e0n);\n";
    for (auto &call : calls) c << call;
    if (plan->qd_tma) c << "    it++;\n";
    c << "  }\n";
    if (plan->scatter_mode == B200_SCATTER_ORDERED) {
      c << "  if (e0p >= 0) b200_ordered_complete(e0p, epoch);\n";
      // the last CTA to finish publishes the epoch of this launch (read by the next one) and resets the counter
      c << "  __syncthreads();\n  if (threadIdx.x == 0) {\n    __threadfence();\n";
      c << "    if (atomicAdd(b200a.ord_sync + 1, 1) == (int)gridDim.x - 1) {\n      b200a.ord_sync[1] = 0;\n      __threadfence();\n";
      c << "      *(volatile int *)b200a.ord_sync = epoch;\n    }\n  }\n";
    }
    c << "}\n";
    return c.str();
  }
};

}  // namespace

std::string b200_opgen_source(B200Operator op, B200OpPlan *plan, int add) {
  if (plan->lean) return b200_opgen_lean_source(op, plan, add);
  Gen g;
  g.op   = op;
  g.plan = plan;
  g.add  = add;
  return g.generate();
}

int b200_opgen_build(B200Operator op, B200OpPlan *plan, int add) {
  B200Ceed           ceed = op->ceed;
  B200KernelVariant &v    = plan->variant[add ? 1 : 0];
  if (v.built) return B200_SUCCESS;
  if (const char *inject = getenv("CEED_B200_FAIL_FUSED_BUILD")) {  // test hook of the fallback ladder: the first <n> fused builds fail
    static int failed = 0;
    if (failed < atoi(inject)) {
      failed++;
      return b200_error(ceed, B200_ERROR_BACKEND, "fused kernel build failure injected by CEED_B200_FAIL_FUSED_BUILD (%d)", failed);
    }
  }
  v.source = b200_opgen_source(op, plan, add);
  B200_CALL(b200_jit_compile(ceed, v.source, {}, &v.module));
  B200_CALL(b200_jit_get_kernel(ceed, v.module, ("b200_operator_" + op->qf->kernel_name).c_str(), &v.kernel));
  if (!b200_compile_only()) {
    if (!v.module->args_dptr) {
      size_t bytes = 0;
      B200_CU(ceed, cuModuleGetGlobal(&v.module->args_dptr, &bytes, v.module->module, "b200a"));
      B200_CHECK(bytes == sizeof(B200OpArgs), ceed, B200_ERROR_BACKEND, "generated argument block has %zu bytes, expected %zu", bytes, sizeof(B200OpArgs));
    }
    B200_CU(ceed, cuFuncSetAttribute(v.kernel, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, plan->smem_bytes));
    int val = 0;
    cuFuncGetAttribute(&val, CU_FUNC_ATTRIBUTE_NUM_REGS, v.kernel);
    v.regs = val;
    cuFuncGetAttribute(&val, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, v.kernel);
    v.local_bytes = val;
    cuFuncGetAttribute(&val, CU_FUNC_ATTRIBUTE_SHARED_SIZE_BYTES, v.kernel);
    v.static_smem = val;
    int nb = 0;
    B200_CU(ceed, cuOccupancyMaxActiveBlocksPerMultiprocessor(&nb, v.kernel, plan->threads, plan->smem_bytes));
    if (nb < 1) return b200_error(ceed, B200_ERROR_BACKEND, "fused kernel cannot be resident (regs %d, smem %d)", v.regs, plan->smem_bytes);
    v.blocks_per_sm = nb;  // per variant: the store and the accumulate kernels may differ in registers
    plan->grid      = b200_opgen_grid(ceed, plan, v, plan->num_elem);
  }
  v.built = true;
  return B200_SUCCESS;
}

// persistent grid of a launch over `num_elem` elements: resident CTAs of this variant x SMs, capped by the CTAs the range needs
int b200_opgen_grid(B200Ceed ceed, const B200OpPlan *plan, const B200KernelVariant &v, long long num_elem) {
  long long num_batches = (num_elem + plan->epb - 1) / plan->epb;
  if (plan->warp_mode) {
    const int groups = plan->threads / (32 * plan->group_warps);
    num_batches      = (num_batches + groups - 1) / groups;  // CTAs needed
  }
  long long grid = (long long)v.blocks_per_sm * ceed->num_sms;
  if (grid > num_batches) grid = num_batches;
  if (grid < 1) grid = 1;
  return (int)grid;
}
