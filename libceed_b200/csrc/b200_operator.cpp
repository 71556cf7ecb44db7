// b200_operator.cpp -- CeedOperator for the b200 backend: field wiring, lazy setup, fused apply, unfused fallback.
//
// Mirrors the host side of the reference's generated-kernel operator (backends/cuda-gen/ceed-cuda-gen-operator.c:105-300):
// on the first apply the operator is analysed and a fused kernel is generated + JIT-compiled (b200_opgen.cpp); every
// apply then packs device pointers (:131-171) and launches ONE kernel (+ a small halo-finalize kernel in deterministic
// scatter mode).  Operators the generator does not cover (1-D/2-D, mixed Q, Q < P gradients) run through an unfused
// sequence of this backend's own restriction / basis / QFunction kernels, like /gpu/cuda/ref does
// (backends/cuda-ref/ceed-cuda-ref-operator.c:519-638) -- never through the CPU.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>

#include "b200_opgen.h"

extern "C" int ceedb200_operator_create(B200Ceed ceed, B200QFunction qf, B200Operator *op_out) {
  B200_CHECK(qf, ceed, B200_ERROR_INCOMPLETE, "operator needs a QFunction");
  B200Operator op = new B200Operator_();
  op->ceed        = ceed;
  op->qf          = qf;
  op->in_fields.resize(qf->inputs.size());
  op->out_fields.resize(qf->outputs.size());
  *op_out = op;
  return B200_SUCCESS;
}

static void plan_free(B200Operator op) {
  B200OpPlan *plan = op->plan;
  if (!plan) return;
  for (int i = 0; i < 16; i++) b200_dfree(op->ceed, plan->aux[i]);
  b200_ordered_scatter_free(op->ceed, &plan->ordered);
  b200_dfree(op->ceed, plan->d_fin_tab);
  b200_dfree(op->ceed, plan->d_fin_done);
  for (int v = 0; v < 2; v++)
    for (int i = 0; i < 16; i++) b200_run_scatter_free(op->ceed, &plan->run[v][i]);
  for (auto *vecs : {&plan->e_in, &plan->q_in, &plan->e_out, &plan->q_out})
    for (auto v : *vecs) ceedb200_vector_destroy(v);
  delete plan;
  op->plan = nullptr;
}

extern "C" int ceedb200_operator_destroy(B200Operator op) {
  if (!op) return B200_SUCCESS;
  plan_free(op);
  for (int i = 0; i < 4; i++)
    if (op->ev[i]) cudaEventDestroy(op->ev[i]);
  delete op;
  return B200_SUCCESS;
}

extern "C" int ceedb200_operator_set_field(B200Operator op, const char *field_name, B200Restriction rstr, B200Basis basis, B200Vector vec) {
  B200Ceed      ceed = op->ceed;
  B200QFunction qf   = op->qf;
  B200_CHECK(!op->is_setup, ceed, B200_ERROR_MAJOR, "Operator cannot be changed after set as immutable");
  for (int io = 0; io < 2; io++) {
    auto &qfields = io ? qf->outputs : qf->inputs;
    auto &fields  = io ? op->out_fields : op->in_fields;
    for (size_t i = 0; i < qfields.size(); i++) {
      if (qfields[i].name != field_name) continue;
      const int emode = qfields[i].eval_mode;
      // same consistency rules as CeedOperatorCheckField (interface/ceed-operator.c:36-82)
      B200_CHECK((rstr == nullptr) == (emode == B200_EVAL_WEIGHT), ceed, B200_ERROR_INCOMPATIBLE,
                 "CEED_ELEMRESTRICTION_NONE and CEED_EVAL_WEIGHT must be used together (field %s)", field_name);
      B200_CHECK((basis == nullptr) == (emode == B200_EVAL_NONE), ceed, B200_ERROR_INCOMPATIBLE,
                 "CEED_BASIS_NONE and CEED_EVAL_NONE must be used together (field %s)", field_name);
      if (rstr && basis)
        B200_CHECK(rstr->num_comp == basis->num_comp, ceed, B200_ERROR_DIMENSION, "Field '%s': restriction has %d components, basis has %d", field_name,
                   rstr->num_comp, basis->num_comp);
      fields[i].rstr      = rstr;
      fields[i].basis     = basis;
      fields[i].is_active = vec == B200_VECTOR_ACTIVE;
      fields[i].vec       = (vec == B200_VECTOR_ACTIVE || vec == B200_VECTOR_NONE) ? nullptr : vec;
      fields[i].is_set    = true;
      return B200_SUCCESS;
    }
  }
  return b200_error(ceed, B200_ERROR_INCOMPLETE, "QFunction has no field named '%s'", field_name);
}

static void operator_reset(B200Operator op) {
  if (op->is_setup) {
    plan_free(op);
    op->is_setup = false;
  }
}

extern "C" int ceedb200_operator_set_tuning(B200Operator op, int elems_per_block, int blocks_per_sm) {
  op->tune.epw  = elems_per_block;
  op->tune.minb = blocks_per_sm;
  op->tuned     = elems_per_block > 0 || blocks_per_sm > 0;  // explicit shape: the autotuner keeps its hands off
  op->build_rung = 0;
  operator_reset(op);
  return B200_SUCCESS;
}

extern "C" int ceedb200_operator_set_kernel_shape(B200Operator op, const int *shape) {
  op->tune.epw = shape[0], op->tune.group_warps = shape[1], op->tune.cta_warps = shape[2], op->tune.minb = shape[3];
  op->tune.qf_mode = shape[4], op->tune.qf_unroll = shape[5], op->tune.stage = shape[6];
  op->tuned = true;
  op->build_rung = 0;
  operator_reset(op);
  return B200_SUCCESS;
}

extern "C" int ceedb200_operator_get_kernel_shape(B200Operator op, int *shape, char *signature, int signature_len) {
  B200_CHECK(op->is_setup && op->plan && op->plan->fused, op->ceed, B200_ERROR_UNSUPPORTED, "operator is not set up as a fused kernel");
  const B200Tuning &t = op->plan->resolved;
  const int         v[7] = {t.epw, t.group_warps, t.cta_warps, t.minb, t.qf_mode, t.qf_unroll, t.stage};
  if (shape) memcpy(shape, v, sizeof(v));
  if (signature && signature_len > 0) snprintf(signature, signature_len, "%s", op->plan->signature.c_str());
  return B200_SUCCESS;
}

static int operator_setup(B200Operator op) {
  if (op->is_setup) return B200_SUCCESS;
  B200Ceed ceed = op->ceed;
  for (auto *fields : {&op->in_fields, &op->out_fields})
    for (size_t i = 0; i < fields->size(); i++) B200_CHECK((*fields)[i].is_set, ceed, B200_ERROR_INCOMPLETE, "Not all operator fields set");
  B200OpPlan *plan = new B200OpPlan();
  op->plan         = plan;
  B200_CALL(b200_opgen_plan(op, plan));
  if (getenv("CEED_B200_NO_FUSE")) {
    plan->fused         = false;
    plan->why_not_fused = "CEED_B200_NO_FUSE set";
  }
  if (plan->fused) {
    // auxiliary buffers of offset-restricted outputs
    auto prepare = [&](B200Restriction r, int slot) -> int {
      if (!r || r->is_strided) return B200_SUCCESS;
      if (plan->scatter_mode == B200_SCATTER_ORDERED && slot == plan->ordered_slot) {
        B200_CALL(b200_restriction_build_ordered(r, plan->epb, &plan->ordered));
        if (!plan->ordered.supported) {  // encoding limits exceeded: same results through the two-pass scheme
          plan->scatter_mode = B200_SCATTER_DETERMINISTIC;
          plan->ordered_slot = -1;
        }
      }
      if (plan->scatter_mode == B200_SCATTER_ORDERED && slot == plan->ordered_slot) {
        plan->aux_bytes[slot] = (size_t)std::max<int64_t>(plan->ordered.num_halo, 1) * r->num_comp * sizeof(double);
      } else if (plan->scatter_mode == B200_SCATTER_DETERMINISTIC || plan->scatter_mode == B200_SCATTER_ORDERED) {
        B200_CALL(b200_restriction_build_owner(r));
        plan->aux_bytes[slot] = (size_t)r->num_halo * r->num_comp * sizeof(double);
      } else if (plan->scatter_mode == B200_SCATTER_EVECTOR) {
        B200_CALL(b200_restriction_build_transpose(r));
        plan->aux_bytes[slot] = (size_t)r->num_elem * r->elem_size * r->num_comp * sizeof(double);
      }
      if (plan->aux_bytes[slot]) B200_CALL(b200_dmalloc(ceed, (void **)&plan->aux[slot], plan->aux_bytes[slot]));
      if (plan->scatter_mode == B200_SCATTER_ORDERED && slot == plan->ordered_slot) B200_CALL(b200_ordered_init_halo(r, &plan->ordered, plan->aux[slot]));
      return B200_SUCCESS;
    };
    for (auto &g : plan->out_groups) B200_CALL(prepare(g.rstr, g.slot));
    for (auto &f : plan->out_fields)
      if (f.emode == B200_EVAL_NONE) B200_CALL(prepare(f.rstr, f.slot));
  } else {
    if (getenv("CEED_B200_DEBUG")) fprintf(stderr, "[ceed-b200] operator not fused: %s\n", plan->why_not_fused.c_str());
  }
  op->is_setup = true;
  return B200_SUCCESS;
}

// ------------------------------------------------------------------------------------------------ fused apply
static int apply_unfused(B200Operator op, B200Vector u, B200Vector v, int add);

// K contiguous element parts with ends at multiples of the kernel's batch width (shared by the streamed apply and the in-kernel finalize)
static std::vector<int32_t> chunk_ends(int num_elem, int epb, int K) {
  const long long      num_batches = ((long long)num_elem + epb - 1) / epb;
  std::vector<int32_t> ends;
  for (int c = 1; c <= K; c++) ends.push_back((int32_t)std::min<long long>(num_elem, num_batches * c / K * epb));
  ends.back() = num_elem;
  return ends;
}
static int default_parts() { return getenv("CEED_B200_PARTS") ? std::max(2, std::min(atoi(getenv("CEED_B200_PARTS")), 64)) : 8; }

// part: 0 = the whole operator; 1 / 2 = boundary / interior elements of a partitioned mesh (ceedb200_operator_apply_part):
// part 1 applies elements [0, split) and finalizes the shared nodes touched by those elements only, part 2 the rest.
static int apply_fused(B200Operator op, B200Vector u, B200Vector v, int add, int part = 0, B200DebugLaunch *desc = nullptr) {
  B200Ceed    ceed = op->ceed;
  B200OpPlan *plan = op->plan;
  B200OpArgs  args;
  memset(&args, 0, sizeof(args));
  args.num_elem = plan->num_elem;
  if (op->qf->ctx) B200_CALL(ceedb200_qfcontext_get_data(op->qf->ctx, B200_MEM_DEVICE, &args.ctx));
  else args.ctx = op->qf->raw_ctx;

  // In-kernel finalize (lean kernel, stage bit 128, whole-mesh launches with the deterministic scatter): the output restriction gets K
  // batch-aligned element parts -- before its tables are used below -- and the kernel folds the shared nodes of a part into v itself.
  bool fin_mode = false;
  if (plan->lean && (plan->stage_mask & 128) && plan->scatter_mode == B200_SCATTER_DETERMINISTIC && !part && !b200_compile_only()) {
    const B200Restriction r = op->out_fields[plan->out_groups[0].slot].rstr;
    const int             K = default_parts();
    const long long       num_batches = ((long long)plan->num_elem + plan->epb - 1) / plan->epb;
    if (r->split_elem < 0 && num_batches >= 64LL * K) {
      B200_CALL(b200_restriction_set_parts(r, chunk_ends(plan->num_elem, plan->epb, K)));
      B200_CALL(b200_restriction_build_owner(r));
      std::vector<int32_t> tab(2 * K + 3);
      tab[0] = K;
      for (int c = 0; c <= K; c++) tab[1 + c] = (int32_t)(c == K ? num_batches : num_batches * c / K);
      for (int c = 0; c <= K; c++) tab[K + 2 + c] = (int32_t)r->shared_prefix[c];
      if (tab != plan->fin_tab_host) {
        if (!plan->d_fin_tab || plan->fin_parts < K) {
          B200_CALL(b200_dfree(ceed, plan->d_fin_tab));
          B200_CALL(b200_dfree(ceed, plan->d_fin_done));
          plan->d_fin_tab = plan->d_fin_done = nullptr;
          B200_CALL(b200_dmalloc(ceed, (void **)&plan->d_fin_tab, tab.size() * sizeof(int32_t)));
          B200_CALL(b200_dmalloc(ceed, (void **)&plan->d_fin_done, K * sizeof(int32_t)));
          plan->fin_parts = K;
        }
        B200_CALL(b200_h2d(ceed, plan->d_fin_tab, tab.data(), tab.size() * sizeof(int32_t)));
        plan->fin_tab_host = tab;
      }
      fin_mode = true;
    }
  }
  // inputs
  for (size_t i = 0; i < op->in_fields.size(); i++) {
    const B200OpField &f = op->in_fields[i];
    if (!f.rstr) continue;
    B200Vector vec = f.is_active ? u : f.vec;
    B200_CHECK(vec && vec != B200_VECTOR_NONE, ceed, B200_ERROR_INCOMPLETE, "missing input vector for field %zu", i);
    B200_CHECK(vec->length >= f.rstr->l_size, ceed, B200_ERROR_DIMENSION, "input vector of field %zu shorter than restriction L-size", i);
    B200_CALL(b200_vector_device_read(vec, &args.in_ptr[i]));
    args.in_idx[i] = f.rstr->is_strided ? nullptr : f.rstr->d_offsets;
  }
  // bulk copies (cp.async.bulk) need 16-byte aligned sources: a caller-owned array that is only 8-byte aligned gets the
  // kernel generated without them (same results)
  if (plan->qd_tma) {
    bool misaligned = false;
    for (size_t i = 0; i < plan->in_fields.size(); i++)
      if (plan->in_fields[i].qd_tma && ((uintptr_t)args.in_ptr[i] & 15)) misaligned = true;
    if (misaligned) {
      op->no_tma = true;
      plan_free(op);
      op->is_setup = false;
      B200_CALL(operator_setup(op));
      B200_CHECK(op->plan->fused && !op->plan->qd_tma, ceed, B200_ERROR_BACKEND, "could not regenerate the operator kernel without bulk copies");
      return apply_fused(op, u, v, add, part, desc);
    }
  }
  // the vector loads that stage the offset table of the lean kernel's interleaved columns (stage bit 1) have the same requirement
  // (the owner / halo target table is always the backend's own allocation)
  if (plan->lean && (plan->stage_mask & 1)) {
    const int slot = plan->in_groups[0].slot;
    if ((uintptr_t)args.in_idx[slot] & 15) {
      op->no_tma = true;
      plan_free(op);
      op->is_setup = false;
      B200_CALL(operator_setup(op));
      B200_CHECK(op->plan->fused && !(op->plan->lean && (op->plan->stage_mask & 1)), ceed, B200_ERROR_BACKEND, "could not regenerate the operator kernel without vector table loads");
      return apply_fused(op, u, v, add, part, desc);
    }
  }
  // outputs.  Decide per distinct output vector whether the kernel can store (overwrite) or must accumulate.
  struct OutVec {
    B200Vector vec;
    int        writers;
    bool       covers;
  };
  std::vector<OutVec> outs;
  for (size_t i = 0; i < op->out_fields.size(); i++) {
    const B200OpField &f   = op->out_fields[i];
    B200Vector         vec = f.is_active ? v : f.vec;
    B200_CHECK(vec && vec != B200_VECTOR_NONE, ceed, B200_ERROR_INCOMPLETE, "missing output vector for field %zu", i);
    B200_CHECK(vec->length >= f.rstr->l_size, ceed, B200_ERROR_DIMENSION, "output vector of field %zu shorter than restriction L-size", i);
    // a field that is the first of its group (or EVAL_NONE) is a distinct writer
    bool is_writer = plan->out_fields[i].emode == B200_EVAL_NONE || plan->out_groups[plan->out_fields[i].group].slot == (int)i;
    if (!is_writer) continue;
    bool covers;
    if (f.rstr->is_strided) {
      covers = (int64_t)f.rstr->num_elem * f.rstr->elem_size * f.rstr->num_comp == vec->length;
    } else {
      B200_CALL(b200_restriction_build_owner(f.rstr));
      covers = f.rstr->num_nodes * f.rstr->num_comp == vec->length;
    }
    bool found = false;
    for (auto &o : outs)
      if (o.vec == vec) {
        o.writers++;
        found = true;
      }
    if (!found) outs.push_back({vec, 1, covers});
  }
  // kernel variant: stores are only safe when every output vector has one writer that covers it and the scatter is the
  // owner/halo scheme (or strided); otherwise zero first and accumulate.
  if (op->timing && !op->ev[0])
    for (int i = 0; i < 4; i++) B200_CUDA(ceed, cudaEventCreate(&op->ev[i]));
  if (op->timing) B200_CUDA(ceed, cudaEventRecord(op->ev[0], ceed->stream));  // aux time includes the memset of non-overwriting modes
  // element range of this launch
  long long e_begin = 0, e_end = plan->num_elem;
  if (part) {
    // element part `part` (1-based) of the output restriction's part table: the two parts of a partitioned mesh
    // (ceedb200_restriction_set_split) or the K chunks of a streamed host-buffer apply (ceedb200_operator_apply_streamed)
    const std::vector<int32_t> *ends = nullptr;
    for (size_t i = 0; i < op->out_fields.size(); i++) {
      const B200Restriction r = op->out_fields[i].rstr;
      if (r->is_strided) continue;
      B200_CHECK(!r->part_ends.empty(), ceed, B200_ERROR_INCOMPLETE, "apply_part needs ceedb200_restriction_set_split on the output restriction");
      B200_CHECK(!ends || *ends == r->part_ends, ceed, B200_ERROR_INCOMPATIBLE, "output restrictions with different element splits");
      ends = &r->part_ends;
    }
    B200_CHECK(ends, ceed, B200_ERROR_INCOMPLETE, "apply_part needs an offset-restricted output");
    B200_CHECK(part <= (int)ends->size(), ceed, B200_ERROR_DIMENSION, "element part %d of %d", part, (int)ends->size());
    B200_CHECK(plan->scatter_mode == B200_SCATTER_DETERMINISTIC, ceed, B200_ERROR_UNSUPPORTED, "apply_part needs the deterministic scatter mode");
    e_begin = part >= 2 ? (*ends)[part - 2] : 0;
    e_end   = (*ends)[part - 1];
  }
  int  kernel_add = add;
  bool zero_first = false;
  if (!add) {
    bool need_zero = false;
    for (auto &o : outs) need_zero = need_zero || o.writers > 1 || !o.covers;
    bool offset_non_det = false;
    for (size_t i = 0; i < op->out_fields.size(); i++)
      if (!op->out_fields[i].rstr->is_strided && plan->scatter_mode != B200_SCATTER_DETERMINISTIC && plan->scatter_mode != B200_SCATTER_ORDERED)
        offset_non_det = true;
    if (need_zero || offset_non_det) zero_first = true, kernel_add = 1;
  }
  // Fallback ladder (the reference's contract: try to compile the fused kernel, else fall back -- backends/cuda-gen/
  // ceed-cuda-gen-operator.c:291-298, backends/cuda/ceed-cuda-compile.cpp:206-224).  Before anything is written: build the kernel
  // variant this apply needs.  If NVRTC fails, the kernel cannot be resident, or it spills heavily (a register-hungry user
  // QFunction), regenerate it with a conservative shape (one element per group, z-line layout, nothing staged, no occupancy
  // target); if that fails too, the operator permanently drops to the unfused restriction / basis / QFunction kernels.  A valid
  // operator never returns an error because of its kernel shape.
  {
    constexpr int     kSpillLimit = 512;  // bytes of local memory per thread tolerated on the tuned shape
    const std::string saved_error = ceed->last_error;
    const int         berr        = b200_opgen_build(op, plan, kernel_add);
    const bool        spilled     = !berr && !b200_compile_only() && plan->variant[kernel_add ? 1 : 0].local_bytes > kSpillLimit && op->build_rung == 0;
    if (berr && op->no_ladder) return berr;
    if ((berr || spilled) && !op->no_ladder) {
      const std::string why = berr ? ceed->last_error : "tuned shape spills " + std::to_string(plan->variant[kernel_add ? 1 : 0].local_bytes) + " bytes per thread";
      if (getenv("CEED_B200_DEBUG")) fprintf(stderr, "[ceed-b200] fused kernel, rung %d: %s\n", op->build_rung, why.c_str());
      if (op->build_rung == 0) {
        const int lines = plan->Q * plan->Q;
        op->build_rung  = 1;
        op->tune        = B200Tuning();
        op->tune.epw = 1, op->tune.group_warps = lines > 64 ? 4 : (lines > 32 ? 2 : 1), op->tune.cta_warps = op->tune.group_warps, op->tune.minb = 1;
        op->tune.qf_mode = 0, op->tune.qf_unroll = 1, op->tune.stage = 0;
        op->tuned = true;  // no autotuning on the way down
        plan_free(op);
        op->is_setup = false;
        B200_CALL(operator_setup(op));
        ceed->last_error = saved_error;
        if (op->plan->fused) return apply_fused(op, u, v, add, part, desc);
        B200_CHECK(!part, ceed, B200_ERROR_UNSUPPORTED, "apply_part needs a fused operator: %s", op->plan->why_not_fused.c_str());
        return apply_unfused(op, u, v, add);
      }
      if (berr) {
        op->build_rung      = 2;
        plan->fused         = false;
        plan->why_not_fused = "fused kernel could not be built: " + why;
        B200_CHECK(!part, ceed, B200_ERROR_UNSUPPORTED, "apply_part needs a fused operator: %s", plan->why_not_fused.c_str());
        ceed->last_error = saved_error;
        return apply_unfused(op, u, v, add);
      }
    }
  }
  if (zero_first && part <= 1 && !desc)  // later parts continue what the first part started
    for (auto &o : outs) B200_CALL(ceedb200_vector_set_value(o.vec, 0.0));
  for (size_t i = 0; i < op->out_fields.size(); i++) {
    const B200OpField &f   = op->out_fields[i];
    B200Vector         vec = f.is_active ? v : f.vec;
    // discard previous contents only when the kernel overwrites everything
    B200_CALL(b200_vector_device_write(vec, &args.out_ptr[i], !kernel_add && part <= 1));
    if (!f.rstr->is_strided) {
      if (plan->scatter_mode == B200_SCATTER_ORDERED && (int)i == plan->ordered_slot) args.out_idx[i] = plan->ordered.d_tgt;
      else args.out_idx[i] = plan->scatter_mode == B200_SCATTER_DETERMINISTIC ? f.rstr->d_tgt : f.rstr->d_offsets;
    }
    args.out_aux[i] = plan->aux[i];
  }
  const bool ordered = plan->scatter_mode == B200_SCATTER_ORDERED;
  if (ordered) {
    args.ord_pred_ptr = plan->ordered.d_pred_ptr, args.ord_pred_idx = plan->ordered.d_pred_idx;
    args.ord_flags = plan->ordered.d_flags, args.ord_sync = plan->ordered.d_sync;
    args.ord_num_halo = plan->ordered.num_halo;
  }
  B200KernelVariant &var = plan->variant[kernel_add ? 1 : 0];  // built above (fallback ladder)
  // Lean kernel, deterministic scatter, whole-mesh launch: every warp owns one contiguous run of elements, and E-entries whose earlier
  // touchers were all processed earlier by the same warp are added straight into v (B200RunScatter) -- fewer halo entries, a smaller
  // finalize pass, same ascending E-order.  The tables depend on the launch shape (warps of the grid, elements per iteration).
  int             run_mode = 0;
  B200RunScatter *run      = nullptr;
  if (plan->lean && plan->scatter_mode == B200_SCATTER_DETERMINISTIC) {
    const int             slot = plan->out_groups[0].slot;
    const B200Restriction r    = op->out_fields[slot].rstr;
    args.ord_num_halo          = r->num_halo;
    if (fin_mode) {
      args.ord_pred_ptr = plan->d_fin_tab, args.ord_pred_idx = r->d_halo_node, args.ord_sync = r->d_halo_ptr, args.ord_flags = plan->d_fin_done;
      run_mode          = 2;
    }
    if (plan->lean_runs && !fin_mode && !part && !(plan->stage_mask & 40) && r->l_size < (int64_t)B200_RUN_RMW_BIT && !getenv("CEED_B200_NO_RUNS") && e_end > e_begin) {
      const int groups = b200_opgen_grid(ceed, plan, var, e_end - e_begin) * (plan->threads / 32);
      run              = &plan->run[kernel_add ? 1 : 0][slot];
      B200_CALL(b200_restriction_build_runs(r, groups, plan->epb, run));
      args.out_idx[slot] = run->d_tgt;
      args.ord_num_halo  = run->num_halo;
      run_mode           = 1;
    }
  }

  if (desc) {
    // host-logic tests (ceedb200_operator_debug_launch): describe the launch -- argument block, grid, element range, the second pass of
    // the deterministic scatter -- instead of performing it
    memset(desc, 0, sizeof(*desc));
    static_assert(sizeof(args) <= sizeof(desc->args), "argument block larger than the debug copy");
    memcpy(desc->args, &args, sizeof(args));
    desc->args_size = (int)sizeof(args);
    desc->source    = var.source.c_str();
    desc->grid = b200_opgen_grid(ceed, plan, var, e_end - e_begin), desc->threads = plan->threads, desc->smem_bytes = plan->smem_bytes;
    desc->run_mode = run_mode, desc->kernel_add = kernel_add, desc->zero_first = zero_first, desc->e_begin = e_begin, desc->e_end = e_end;
    desc->fin_slot = -1;
    for (size_t i = 0; i < op->out_fields.size() && plan->num_elem > 0; i++) {
      const B200OpField &f = op->out_fields[i];
      const bool is_writer = plan->out_fields[i].emode == B200_EVAL_NONE || plan->out_groups[plan->out_fields[i].group].slot == (int)i;
      if (f.rstr->is_strided || !is_writer) continue;
      B200_CHECK(!desc->v, ceed, B200_ERROR_UNSUPPORTED, "debug launch description holds one offset-restricted output");
      desc->v = args.out_ptr[i], desc->num_comp = f.rstr->num_comp, desc->comp_stride = f.rstr->comp_stride;
      desc->scatter_mode = plan->scatter_mode;
      if (plan->scatter_mode == B200_SCATTER_EVECTOR)
        desc->e_entries = (long long)f.rstr->num_elem * f.rstr->elem_size, desc->offsets = f.rstr->d_offsets, desc->evec = plan->aux[i];
      if (!plan->aux[i] || run || fin_mode || plan->scatter_mode != B200_SCATTER_DETERMINISTIC) continue;
      int64_t first = 0, count = f.rstr->num_shared;
      if (part >= 1) first = f.rstr->shared_prefix[part - 1], count = f.rstr->shared_prefix[part] - first;
      desc->fin_slot = (int)i, desc->num_comp = f.rstr->num_comp, desc->comp_stride = f.rstr->comp_stride;
      desc->num_shared = count, desc->num_halo = f.rstr->num_halo;
      desc->halo_node = f.rstr->d_halo_node + first, desc->halo_ptr = f.rstr->d_halo_ptr + first;
      desc->halo = plan->aux[i], desc->v = args.out_ptr[i];
    }
    return B200_SUCCESS;
  }
  if (op->timing) B200_CUDA(ceed, cudaEventRecord(op->ev[3], ceed->stream));
  B200_CHECK(!(ordered && part), ceed, B200_ERROR_UNSUPPORTED, "apply_part is not available with the in-kernel ordered scatter");
  if (e_end > e_begin) {
    // the argument block is a __constant__ object of the module: rewrite it only when a pointer changed (stream-ordered copy)
    B200Module *mod = var.module;
    if (!b200_compile_only() && (mod->last_args.size() != sizeof(args) || memcmp(mod->last_args.data(), &args, sizeof(args)) != 0)) {
      mod->last_args.assign((const char *)&args, (const char *)&args + sizeof(args));
      B200_CUDA(ceed, cudaMemcpyAsync((void *)mod->args_dptr, mod->last_args.data(), sizeof(args), cudaMemcpyHostToDevice, ceed->stream));
    }
    // element groups of the ordered scatter wait for each other: all CTAs must be resident (cooperative launch)
    void *params[3] = {&e_begin, &e_end, &run_mode};  // (the general kernel takes the first two)
    if (fin_mode) B200_CUDA(ceed, cudaMemsetAsync(plan->d_fin_done, 0, plan->fin_parts * sizeof(int32_t), ceed->stream));
    // (in-kernel finalize: warps wait for each other's parts, so all CTAs must be resident -- cooperative launch)
    B200_CALL(b200_launch(ceed, var.kernel, b200_opgen_grid(ceed, plan, var, e_end - e_begin), plan->threads, plan->smem_bytes, params, ordered || fin_mode));
  }
  if (op->timing) B200_CUDA(ceed, cudaEventRecord(op->ev[1], ceed->stream));
  // second phase of the scatter
  for (size_t i = 0; i < op->out_fields.size() && plan->num_elem > 0; i++) {
    const B200OpField &f = op->out_fields[i];
    if (f.rstr->is_strided || !plan->aux[i]) continue;
    bool is_writer = plan->out_fields[i].emode == B200_EVAL_NONE || plan->out_groups[plan->out_fields[i].group].slot == (int)i;
    if (!is_writer) continue;
    if (ordered && (int)i == plan->ordered_slot) continue;  // completed inside the kernel
    if (fin_mode && (int)i == plan->out_groups[0].slot) continue;  // folded in by the kernel
    if (run && (int)i == plan->out_groups[0].slot)
      B200_CALL(b200_halo_finalize_lists(f.rstr, run->d_halo_node, run->d_halo_ptr, run->num_shared, run->num_halo, plan->aux[i], args.out_ptr[i]));
    else if (plan->scatter_mode == B200_SCATTER_DETERMINISTIC || ordered) B200_CALL(b200_halo_finalize(f.rstr, plan->aux[i], args.out_ptr[i], part));
    else if (plan->scatter_mode == B200_SCATTER_EVECTOR) B200_CALL(b200_restriction_apply_raw(f.rstr, B200_TRANSPOSE, plan->aux[i], args.out_ptr[i]));
  }
  if (op->timing) {
    B200_CUDA(ceed, cudaEventRecord(op->ev[2], ceed->stream));
    B200_CUDA(ceed, cudaEventSynchronize(op->ev[2]));
    float pre_ms = 0.f;
    cudaEventElapsedTime(&pre_ms, op->ev[0], op->ev[3]);
    cudaEventElapsedTime(&op->last_fused_ms, op->ev[3], op->ev[1]);
    cudaEventElapsedTime(&op->last_aux_ms, op->ev[1], op->ev[2]);
    op->last_aux_ms += pre_ms;
  }
  return B200_SUCCESS;
}

// ------------------------------------------------------------------------------------------------ unfused apply
static int apply_unfused(B200Operator op, B200Vector u, B200Vector v, int add) {
  B200Ceed      ceed = op->ceed;
  B200QFunction qf   = op->qf;
  B200OpPlan   *plan = op->plan;
  int           num_elem = -1, nqpts = -1;
  auto          ipow     = [](int b, int e) {
    int r = 1;
    for (int i = 0; i < e; i++) r *= b;
    return r;
  };
  for (auto *fields : {&op->in_fields, &op->out_fields})
    for (auto &f : *fields) {
      if (f.rstr && num_elem < 0) num_elem = f.rstr->num_elem;
      if (f.basis && nqpts < 0) nqpts = f.basis->is_tensor ? ipow(f.basis->Q, f.basis->dim) : f.basis->Q;
    }
  if (nqpts < 0)
    for (auto *fields : {&op->in_fields, &op->out_fields})
      for (auto &f : *fields)
        if (f.rstr && nqpts < 0) nqpts = f.rstr->elem_size;
  B200_CHECK(num_elem >= 0 && nqpts > 0, ceed, B200_ERROR_INCOMPLETE, "operator has no restriction / quadrature space");
  const int64_t Qtot = (int64_t)num_elem * nqpts;
  B200_CHECK(Qtot < (1LL << 31), ceed, B200_ERROR_UNSUPPORTED, "Backend does not implement unfused operators with more than 2^31 quadrature points");
  // work vectors
  if (plan->q_in.empty() && plan->q_out.empty()) {
    plan->e_in.assign(op->in_fields.size(), nullptr);
    plan->q_in.assign(op->in_fields.size(), nullptr);
    plan->e_out.assign(op->out_fields.size(), nullptr);
    plan->q_out.assign(op->out_fields.size(), nullptr);
    for (size_t i = 0; i < op->in_fields.size(); i++) {
      const B200OpField &f = op->in_fields[i];
      if (f.rstr) B200_CALL(ceedb200_vector_create(ceed, (int64_t)f.rstr->num_elem * f.rstr->elem_size * f.rstr->num_comp, &plan->e_in[i]));
      if (qf->inputs[i].eval_mode != B200_EVAL_NONE) B200_CALL(ceedb200_vector_create(ceed, Qtot * qf->inputs[i].size, &plan->q_in[i]));
    }
    for (size_t i = 0; i < op->out_fields.size(); i++) {
      const B200OpField &f = op->out_fields[i];
      B200_CALL(ceedb200_vector_create(ceed, (int64_t)f.rstr->num_elem * f.rstr->elem_size * f.rstr->num_comp, &plan->e_out[i]));
      if (qf->outputs[i].eval_mode != B200_EVAL_NONE) B200_CALL(ceedb200_vector_create(ceed, Qtot * qf->outputs[i].size, &plan->q_out[i]));
    }
  }
  std::vector<B200Vector> U(op->in_fields.size()), V(op->out_fields.size());
  for (size_t i = 0; i < op->in_fields.size(); i++) {
    const B200OpField &f     = op->in_fields[i];
    const int          emode = qf->inputs[i].eval_mode;
    if (emode == B200_EVAL_WEIGHT) {
      B200_CALL(ceedb200_basis_apply(f.basis, num_elem, B200_NOTRANSPOSE, B200_EVAL_WEIGHT, B200_VECTOR_NONE, plan->q_in[i]));
      U[i] = plan->q_in[i];
      continue;
    }
    B200Vector vec = f.is_active ? u : f.vec;
    B200_CHECK(vec && vec != B200_VECTOR_NONE, ceed, B200_ERROR_INCOMPLETE, "missing input vector for field %zu", i);
    B200_CALL(ceedb200_restriction_apply(f.rstr, B200_NOTRANSPOSE, vec, plan->e_in[i]));
    if (emode == B200_EVAL_NONE) U[i] = plan->e_in[i];
    else {
      B200_CALL(ceedb200_basis_apply(f.basis, num_elem, B200_NOTRANSPOSE, emode, plan->e_in[i], plan->q_in[i]));
      U[i] = plan->q_in[i];
    }
  }
  for (size_t i = 0; i < op->out_fields.size(); i++) V[i] = qf->outputs[i].eval_mode == B200_EVAL_NONE ? plan->e_out[i] : plan->q_out[i];
  if (Qtot > 0) B200_CALL(ceedb200_qfunction_apply(qf, (b200_int)Qtot, U.data(), V.data()));
  // zero distinct output vectors for overwrite semantics
  if (!add) {
    std::vector<B200Vector> done;
    for (size_t i = 0; i < op->out_fields.size(); i++) {
      B200Vector vec = op->out_fields[i].is_active ? v : op->out_fields[i].vec;
      bool       seen = false;
      for (auto d : done) seen = seen || d == vec;
      if (!seen) {
        B200_CALL(ceedb200_vector_set_value(vec, 0.0));
        done.push_back(vec);
      }
    }
  }
  for (size_t i = 0; i < op->out_fields.size(); i++) {
    const B200OpField &f     = op->out_fields[i];
    const int          emode = qf->outputs[i].eval_mode;
    B200Vector         vec   = f.is_active ? v : f.vec;
    B200_CHECK(vec && vec != B200_VECTOR_NONE, ceed, B200_ERROR_INCOMPLETE, "missing output vector for field %zu", i);
    if (emode != B200_EVAL_NONE) B200_CALL(ceedb200_basis_apply(f.basis, num_elem, B200_TRANSPOSE, emode, plan->q_out[i], plan->e_out[i]));
    B200_CALL(ceedb200_restriction_apply(f.rstr, B200_TRANSPOSE, plan->e_out[i], vec));
  }
  return B200_SUCCESS;
}

// ------------------------------------------------------------------------------------------------ QFunction assembly
// CeedOperatorLinearAssembleQFunction (interface/ceed-preconditioning.c; backends/cuda-ref/ceed-cuda-ref-operator.c:1000-1130 for
// the GPU-side contract): the pointwise linear map of the QFunction, evaluated at every quadrature point by feeding unit vectors
// through the active inputs.  Layout of the result (this backend's choice, published through the strided restriction the host
// layer creates): assembled[((a * size_out + b) * num_elem + e) * Q + q] = d out_b / d in_a at point q of element e, i.e.
// strides {1, num_elem * Q, Q}.  Passive inputs (quadrature data, weights, coefficient fields) are evaluated once with the
// unfused restriction / basis kernels; the QFunction runs size_in times over all points.  This is what the interface needs
// to provide LinearAssembleDiagonal / LinearAssemble / multigrid on top.
static int assembly_sizes(B200Operator op, int *num_elem, int *nqpts, int *size_in, int *size_out) {
  B200QFunction qf = op->qf;
  *num_elem = -1, *nqpts = -1, *size_in = 0, *size_out = 0;
  for (auto *fields : {&op->in_fields, &op->out_fields})
    for (auto &f : *fields) {
      if (f.rstr && *num_elem < 0) *num_elem = f.rstr->num_elem;
      if (f.basis && *nqpts < 0) {
        *nqpts = f.basis->is_tensor ? 1 : f.basis->Q;
        for (int d = 0; d < f.basis->dim && f.basis->is_tensor; d++) *nqpts *= f.basis->Q;
      }
    }
  if (*nqpts < 0)
    for (auto *fields : {&op->in_fields, &op->out_fields})
      for (auto &f : *fields)
        if (f.rstr && *nqpts < 0) *nqpts = f.rstr->elem_size;
  for (size_t i = 0; i < op->in_fields.size(); i++)
    if (op->in_fields[i].is_active) *size_in += qf->inputs[i].size;
  for (size_t i = 0; i < op->out_fields.size(); i++)
    if (op->out_fields[i].is_active) *size_out += qf->outputs[i].size;
  B200_CHECK(*num_elem >= 0 && *nqpts > 0, op->ceed, B200_ERROR_INCOMPLETE, "operator has no restriction / quadrature space");
  B200_CHECK(*size_in > 0 && *size_out > 0, op->ceed, B200_ERROR_BACKEND, "Cannot assemble QFunction without active inputs and outputs");
  return B200_SUCCESS;
}

extern "C" int ceedb200_operator_assemble_qfunction_sizes(B200Operator op, b200_int *num_elem, b200_int *num_qpts, b200_int *size_in, b200_int *size_out) {
  for (auto *fields : {&op->in_fields, &op->out_fields})
    for (size_t i = 0; i < fields->size(); i++) B200_CHECK((*fields)[i].is_set, op->ceed, B200_ERROR_INCOMPLETE, "Not all operator fields set");
  return assembly_sizes(op, num_elem, num_qpts, size_in, size_out);
}

extern "C" int ceedb200_operator_assemble_qfunction(B200Operator op, B200Vector assembled) {
  B200Ceed      ceed = op->ceed;
  B200QFunction qf   = op->qf;
  int           num_elem, nqpts, size_in, size_out;
  B200_CALL(ceedb200_operator_assemble_qfunction_sizes(op, &num_elem, &nqpts, &size_in, &size_out));
  if (!b200_compile_only()) B200_CUDA(ceed, cudaSetDevice(ceed->device_id));
  const int64_t Qtot = (int64_t)num_elem * nqpts;
  B200_CHECK(Qtot < (1LL << 31), ceed, B200_ERROR_UNSUPPORTED, "Backend does not implement QFunction assembly with more than 2^31 quadrature points");
  B200_CHECK(assembled->length == Qtot * size_in * size_out, ceed, B200_ERROR_DIMENSION, "assembled vector has length %lld, expected %lld",
             (long long)assembled->length, (long long)(Qtot * size_in * size_out));
  double *d_asm;
  B200_CALL(b200_vector_device_write(assembled, &d_asm, true));
  if (Qtot == 0) return B200_SUCCESS;
  // inputs at the quadrature points: passive ones evaluated, active ones zero
  std::vector<B200Vector> owned;  // work vectors of this call
  auto cleanup = [&]() {
    for (auto v : owned) ceedb200_vector_destroy(v);
  };
  std::vector<const double *> in_ptr(std::max<size_t>(1, op->in_fields.size()), nullptr);
  std::vector<double *>       act_ptr(op->in_fields.size(), nullptr), out_ptr(std::max<size_t>(1, op->out_fields.size()), nullptr);
  auto work = [&](int64_t n, B200Vector *v) -> int {
    B200_CALL(ceedb200_vector_create(ceed, n, v));
    owned.push_back(*v);
    return B200_SUCCESS;
  };
  int ierr = B200_SUCCESS;
  for (size_t i = 0; i < op->in_fields.size() && !ierr; i++) {
    const B200OpField &f     = op->in_fields[i];
    const int          emode = qf->inputs[i].eval_mode;
    B200Vector         q     = nullptr;
    if (f.is_active) {
      if ((ierr = work(Qtot * qf->inputs[i].size, &q))) break;
      if ((ierr = ceedb200_vector_set_value(q, 0.0))) break;
      if ((ierr = b200_vector_device_write(q, &act_ptr[i], false))) break;
      in_ptr[i] = act_ptr[i];
      continue;
    }
    if (emode == B200_EVAL_WEIGHT) {
      if ((ierr = work(Qtot, &q))) break;
      if ((ierr = ceedb200_basis_apply(f.basis, num_elem, B200_NOTRANSPOSE, B200_EVAL_WEIGHT, B200_VECTOR_NONE, q))) break;
    } else {
      if (!f.vec || f.vec == B200_VECTOR_NONE) {
        ierr = b200_error(ceed, B200_ERROR_INCOMPLETE, "missing passive input vector for field %zu", i);
        break;
      }
      B200Vector e = nullptr;
      if ((ierr = work((int64_t)f.rstr->num_elem * f.rstr->elem_size * f.rstr->num_comp, &e))) break;
      if ((ierr = ceedb200_restriction_apply(f.rstr, B200_NOTRANSPOSE, f.vec, e))) break;
      if (emode == B200_EVAL_NONE) q = e;
      else {
        if ((ierr = work(Qtot * qf->inputs[i].size, &q))) break;
        if ((ierr = ceedb200_basis_apply(f.basis, num_elem, B200_NOTRANSPOSE, emode, e, q))) break;
      }
    }
    if ((ierr = b200_vector_device_read(q, &in_ptr[i]))) break;
  }
  // passive outputs are computed into scratch space and dropped
  for (size_t i = 0; i < op->out_fields.size() && !ierr; i++) {
    if (op->out_fields[i].is_active) continue;
    B200Vector q = nullptr;
    if ((ierr = work(Qtot * qf->outputs[i].size, &q))) break;
    ierr = b200_vector_device_write(q, &out_ptr[i], true);
  }
  // one QFunction sweep per active input component: unit input -> one block column of the assembled map
  B200Vector unit = nullptr;
  if (!ierr) ierr = work(0, &unit);
  int a = 0;
  for (size_t i = 0; i < op->in_fields.size() && !ierr; i++) {
    if (!op->in_fields[i].is_active) continue;
    for (int comp = 0; comp < qf->inputs[i].size && !ierr; comp++, a++) {
      double *slice = act_ptr[i] + (int64_t)comp * Qtot;
      ceedb200_vector_destroy(unit);
      owned.pop_back();
      if ((ierr = work(Qtot, &unit))) break;
      if ((ierr = ceedb200_vector_set_array(unit, B200_MEM_DEVICE, B200_USE_POINTER, slice))) break;
      if ((ierr = ceedb200_vector_set_value(unit, 1.0))) break;
      int64_t off = (int64_t)a * size_out * Qtot;
      for (size_t o = 0; o < op->out_fields.size(); o++) {
        if (!op->out_fields[o].is_active) continue;
        out_ptr[o] = d_asm + off;
        off += (int64_t)qf->outputs[o].size * Qtot;
      }
      if ((ierr = ceedb200_qfunction_apply_ptr(qf, (b200_int)Qtot, in_ptr.data(), out_ptr.data()))) break;
      if ((ierr = ceedb200_vector_set_value(unit, 0.0))) break;
    }
  }
  if (!ierr && !b200_compile_only()) {
    cudaError_t e = cudaStreamSynchronize(ceed->stream);  // the work vectors are released below
    if (e != cudaSuccess) ierr = b200_error(ceed, B200_ERROR_BACKEND, "CUDA error in QFunction assembly: %s", cudaGetErrorString(e));
  }
  cleanup();
  return ierr;
}

// ------------------------------------------------------------------------------------------------ autotuner
// Opt-in (ceedb200_set_autotune / CEED_B200_AUTOTUNE): on the first overwrite-apply of a fused operator, time a small set
// of kernel shapes on the caller's own vectors (Apply is idempotent) and keep the fastest.  Coordinate search: group width
// x QFunction stage layout, then elements per group, then the launch-bounds occupancy target.  The result goes into the
// context's tuning table (and, with CEED_B200_TUNE_SAVE=<file>, is appended to a table file that ceedb200_init loads).
static int operator_autotune(B200Operator op, B200Vector u, B200Vector v) {
  B200Ceed ceed = op->ceed;
  op->tuned     = true;
  if (!op->plan->fused || b200_compile_only()) return B200_SUCCESS;
  if ((long long)op->plan->num_elem * op->plan->Q * op->plan->Q * op->plan->Q < 200000) return B200_SUCCESS;  // too small to time
  if (ceed->autotune < 2 && (ceed->tune_table.count(op->plan->signature) || ceed->tune_table.count(op->plan->shape_signature))) return B200_SUCCESS;
  const bool        debug = getenv("CEED_B200_DEBUG") != nullptr;
  const bool        timing0 = op->timing;
  const std::string sig = op->plan->signature, shape_sig = op->plan->shape_signature;
  const std::string saved_error = ceed->last_error;
  op->timing    = true;
  op->no_ladder = true;
  B200Tuning best = op->plan->resolved;
  float      best_ms = 1e30f;
  auto trial = [&](B200Tuning t) -> float {
    operator_reset(op);
    op->tune = t;
    float ms = 1e30f;
    if (operator_setup(op) == B200_SUCCESS && op->plan->fused) {
      bool ok = true;
      for (int i = 0; i < 4 && ok; i++) {
        ok = apply_fused(op, u, v, 0) == B200_SUCCESS;
        if (ok && i > 0) ms = std::min(ms, op->last_fused_ms + op->last_aux_ms);
      }
      if (!ok) ms = 1e30f;
      t = op->plan->resolved;
    }
    if (debug)
      fprintf(stderr, "[ceed-b200] autotune %s: epw %d gw %d warps %d minb %d qf %d unroll %d -> %.4f ms\n", sig.c_str(), t.epw, t.group_warps, t.cta_warps,
              t.minb, t.qf_mode, t.qf_unroll, ms);
    if (ms < best_ms) best_ms = ms, best = t;
    return ms;
  };
  B200Tuning base;  // all heuristic
  bool       xline_ok = true;
  for (auto &g : op->plan->in_groups) xline_ok = xline_ok && !g.use_grad;
  for (auto &g : op->plan->out_groups) xline_ok = xline_ok && !g.use_grad;
  // 1. group width / warps per CTA x QFunction layout x plane layout (elements per group and occupancy target left to the heuristics)
  // (16-wide rows for Q = 9, 10 exist and are tested, but need 160-lane groups to pay: measured 2-3x slower with 128 lanes)
  bool swz_ok = op->plan->Q <= 8 && op->plan->scatter_mode != B200_SCATTER_ORDERED;
  for (auto &b : op->plan->bases) swz_ok = swz_ok && b.P <= 8;
  const bool full = ceed->autotune >= 3;  // level 3: also the rarely winning point-pair layout
  const int  shapes[][2] = {{1, 4}, {2, 2}, {2, 8}, {4, 4}, {1, 8}};
  for (auto &sh : shapes)
    for (int qf = 0; qf < 4; qf++) {
      if (qf == 2 && (op->plan->Q % 2 || !full)) continue;  // point pairs need an even number of points per row
      if (qf == 3 && !xline_ok) continue;                   // x-line fusion exists for gradient-free operators only
      if (qf == 0 && xline_ok && !full) continue;           // gradient-free operators: the x-line layout wins
      for (int swz = 0; swz < 3; swz++) {
        if (swz == 1 && (!swz_ok || qf == 1 || qf == 2)) continue;  // the conflict-free swizzled planes exist for the z-line / x-line layouts
        if (swz == 2 && (op->plan->Q % 2 || qf == 1 || qf == 2)) continue;  // even-Q linear layout (16-byte x-lines)
        B200Tuning t  = base;
        t.group_warps = sh[0], t.cta_warps = sh[1], t.qf_mode = qf;
        if (swz == 1) t.stage = 257;
        if (swz == 2) t.stage = 513;
        if (qf == 2) t.qf_unroll = 2;
        trial(t);
      }
    }
  // 1b. gradient-free operators: the lean in-place-plane kernel (layout 4; one warp per group, E elements per warp and iteration)
  //     -- batch widths by order: the planes of E elements must leave room for 16-20 resident warps (profiles/r02c_lean_sweep2.txt)
  if (xline_ok)
    for (int warps : {4, 8})
      for (int k = 0; k < 3; k++) {
        const int Pmax = op->plan->bases.empty() ? 4 : op->plan->bases[0].P;
        const int epw  = Pmax >= 6 ? 1 + k : (Pmax == 5 ? 2 + k : 4 + 2 * k);
        B200Tuning t  = base;
        t.group_warps = 1, t.cta_warps = warps, t.qf_mode = 4, t.epw = epw, t.stage = 0;
        trial(t);
      }
  // 2. elements per group around the winner
  {
    const B200Tuning win = best;
    for (int epw : {1, 2, 3, 4, 5, 6, 7, 8, 10, 12}) {
      if (win.qf_mode != 4 && (epw == 5 || epw == 7 || epw == 10)) continue;
      if (epw == win.epw) continue;
      B200Tuning t = base;
      t.group_warps = win.group_warps, t.cta_warps = win.cta_warps, t.qf_mode = win.qf_mode, t.epw = epw;
      if (win.stage >= 0 && (win.stage & (256 | 512))) t.stage = win.stage;
      trial(t);
    }
  }
  // 3. occupancy target (register cap) and points in flight
  {
    const B200Tuning win = best;
    for (int minb : {win.minb - 1, win.minb + 1, win.minb + 2, win.minb * 2}) {
      if (minb < 1 || minb == win.minb || minb * win.cta_warps > 64) continue;
      B200Tuning t = win;
      t.minb       = minb;
      trial(t);
    }
    if (best.qf_mode >= 1 && best.qf_mode <= 3) {
      const B200Tuning win2 = best;
      for (int unroll : {1, 2, 4, 8}) {
        if (unroll == win2.qf_unroll || (win2.qf_mode == 2 && unroll > 4)) continue;
        B200Tuning t = win2;
        t.qf_unroll  = unroll;
        trial(t);
      }
    }
  }
  // 4. cp.async staging: scatter targets only (1), + gather offsets (9), nothing (0) -- keeping the winner's plane layout bit
  {
    const B200Tuning win = best;
    const int        layout_bit = win.stage >= 0 ? (win.stage & (256 | 512)) : 0;
    for (int stage : {9, 0}) {
      if ((stage | layout_bit) == win.stage || win.qf_mode == 4) continue;
      B200Tuning t = win;
      t.stage      = stage | layout_bit;
      trial(t);
    }
  }
  // 4b. lean kernel: element-interleaved columns (1), + padded element stride (3), 16-byte quadrature-data loads (4) and their unions,
  //     at the winner's batch width and at 8 elements (one full node row of 8 elements per request when P = 4)
  if (best.qf_mode == 4) {
    const B200Tuning win = best;
    for (int pass = 0; pass < (win.epw == 8 ? 1 : 2); pass++)
      for (int stage : {1, 3, 4, 5, 7}) {
        const int epw = pass ? 8 : win.epw;
        B200Tuning t = win;
        t.epw = epw, t.stage = (win.stage > 0 ? win.stage & ~7 : 0) | stage;
        trial(t);
      }
  }
  operator_reset(op);
  op->tune      = best;
  op->timing    = timing0;
  op->no_ladder = false;
  ceed->last_error = saved_error;  // failed candidates are not errors of this call
  B200_CALL(operator_setup(op));
  ceed->tune_table[sig] = best;
  if (!ceed->tune_table.count(shape_sig)) ceed->tune_table[shape_sig] = best;
  if (const char *path = getenv("CEED_B200_TUNE_SAVE")) {
    if (FILE *f = fopen(path, "a")) {
      fprintf(f, "%s %d %d %d %d %d %d %d  # %.4f ms, %d elements\n", sig.c_str(), best.epw, best.group_warps, best.cta_warps, best.minb, best.qf_mode,
              best.qf_unroll, best.stage, best_ms, op->plan->num_elem);
      fclose(f);
    }
  }
  if (debug) fprintf(stderr, "[ceed-b200] autotune %s: best %.4f ms\n", sig.c_str(), best_ms);
  return B200_SUCCESS;
}

static int operator_apply(B200Operator op, B200Vector u, B200Vector v, int add) {
  B200_CALL(operator_setup(op));
  if (!b200_compile_only()) B200_CUDA(op->ceed, cudaSetDevice(op->ceed->device_id));
  if (op->ceed->autotune && !op->tuned && !add) B200_CALL(operator_autotune(op, u, v));
  if (op->plan->fused) return apply_fused(op, u, v, add);
  return apply_unfused(op, u, v, add);
}

extern "C" int ceedb200_operator_apply(B200Operator op, B200Vector u, B200Vector v) { return operator_apply(op, u, v, 0); }

// One half of an apply on a partitioned mesh (multi-GPU overlap): part 1 = the boundary elements
// [0, split) (+ the finalize pass of the nodes only they touch), part 2 = the interior elements.  After part 1 every interface
// DoF holds this rank's complete partial sum, so the exchange can run while part 2 executes.  part 1 followed by part 2 gives
// exactly the bits of ceedb200_operator_apply.
extern "C" int ceedb200_operator_apply_part(B200Operator op, B200Vector u, B200Vector v, int part) {
  B200_CHECK(part == 1 || part == 2, op->ceed, B200_ERROR_DIMENSION, "part must be 1 (boundary elements) or 2 (interior elements)");
  B200_CALL(operator_setup(op));
  if (!b200_compile_only()) B200_CUDA(op->ceed, cudaSetDevice(op->ceed->device_id));
  B200_CHECK(op->plan->fused, op->ceed, B200_ERROR_UNSUPPORTED, "apply_part needs a fused operator: %s", op->plan->why_not_fused.c_str());
  return apply_fused(op, u, v, 0, part);
}
extern "C" int ceedb200_operator_apply_add(B200Operator op, B200Vector u, B200Vector v) { return operator_apply(op, u, v, 1); }

// Chunk tables of the streamed apply for K chunks (cached in the plan): element chunk ends, per chunk how much of u must have arrived
// (prefix maximum of the gathered L-indices) and how much of v is final (suffix minimum of the L-indices later chunks still touch).
static void build_stream_plan(B200OpPlan *plan, B200Restriction rin, B200Restriction rout, int K) {
  const int ne = rout->num_elem;
  B200StreamPlan &sp = plan->stream;
  if (sp.num_chunks == K) return;
  sp = B200StreamPlan();
  sp.ends = chunk_ends(ne, plan->epb, K);  // (graded chunk sizes were measured: no gain, the PCIe duplex rate is the bound)
  auto ranges = [&](B200Restriction r, std::vector<int64_t> &lo, std::vector<int64_t> &hi) {
    lo.assign(K, INT64_MAX), hi.assign(K, -1);
    for (int c = 0; c < K; c++) {
      const int64_t b = (c ? sp.ends[c - 1] : 0) * (int64_t)r->elem_size, e = sp.ends[c] * (int64_t)r->elem_size;
      for (int64_t i = b; i < e; i++) {
        const int64_t l = r->h_offsets[i];
        lo[c] = std::min(lo[c], l), hi[c] = std::max(hi[c], l);
      }
    }
  };
  std::vector<int64_t> in_lo, in_hi, out_lo, out_hi;
  ranges(rin, in_lo, in_hi);
  ranges(rout, out_lo, out_hi);
  const auto blocked = [&](B200Restriction r, const std::vector<int64_t> &hi) {
    return r->comp_stride >= *std::max_element(hi.begin(), hi.end()) + 1 && (int64_t)r->num_comp * r->comp_stride == r->l_size;
  };
  sp.per_comp = rin->num_comp > 1 && rin->num_comp == rout->num_comp && blocked(rin, in_hi) && blocked(rout, out_hi);
  sp.in_hi.resize(K), sp.out_done.resize(K);
  for (int c = 0; c < K; c++) sp.in_hi[c] = std::max(c ? sp.in_hi[c - 1] : (int64_t)0, in_hi[c] + 1);
  int64_t later = INT64_MAX;  // lowest offset touched by a later chunk: everything below it is complete
  for (int c = K - 1; c >= 0; c--) {
    sp.out_done[c] = later;
    later          = std::min(later, out_lo[c]);
  }
  sp.num_chunks = K;
}

// host-logic tests: everything the fused apply would launch for (u, v) -- the kernel's argument block, launch shape, element range and the
// tables of the finalize pass -- without launching anything.  In compile-only mode (CEED_B200_COMPILE_ONLY: "device" memory is host memory)
// tests/test_kernel_emulation.py runs the generated source of the same kernel on CPU threads against these arguments.
extern "C" int ceedb200_operator_debug_launch(B200Operator op, B200Vector u, B200Vector v, int add, int part, B200DebugLaunch *desc) {
  B200_CALL(operator_setup(op));
  B200_CHECK(op->plan->fused, op->ceed, B200_ERROR_UNSUPPORTED, "operator is not fused: %s", op->plan->why_not_fused.c_str());
  B200_CHECK(desc, op->ceed, B200_ERROR_MAJOR, "no description to fill");
  return apply_fused(op, u, v, add, part, desc);
}

// host-logic tests: the chunk tables for K chunks (arrays of K entries each); active input / output restrictions as the streamed apply picks them
extern "C" int ceedb200_operator_debug_stream_plan(B200Operator op, int num_chunks, int32_t *ends, int64_t *in_hi, int64_t *out_done, int *per_comp) {
  B200_CALL(operator_setup(op));
  B200_CHECK(op->plan->fused, op->ceed, B200_ERROR_UNSUPPORTED, "operator is not fused");
  B200Restriction rin = nullptr, rout = nullptr;
  for (auto &f : op->in_fields)
    if (f.is_active && f.rstr && !f.rstr->is_strided) rin = f.rstr;
  for (auto &f : op->out_fields)
    if (f.is_active && f.rstr && !f.rstr->is_strided) rout = f.rstr;
  B200_CHECK(rin && rout, op->ceed, B200_ERROR_UNSUPPORTED, "no offset-restricted active input / output");
  op->plan->stream.num_chunks = 0;
  build_stream_plan(op->plan, rin, rout, num_chunks);
  const B200StreamPlan &sp = op->plan->stream;
  for (int c = 0; c < num_chunks; c++) ends[c] = sp.ends[c], in_hi[c] = sp.in_hi[c], out_done[c] = sp.out_done[c];
  *per_comp = sp.per_comp;
  return B200_SUCCESS;
}

// ------------------------------------------------------------------------------------------------ streamed host-buffer apply
// v = A u with u valid on the HOST only and v wanted on the host (the end-to-end call of an application that keeps its vectors in host
// memory: CeedVectorSetArray(HOST) -> CeedOperatorApply -> CeedVectorSyncArray(HOST)).  Instead of copy-in, apply, copy-out one after
// the other, the elements are cut into K contiguous chunks: chunk c is applied (fused kernel + the finalize pass of the shared nodes it
// completes) as soon as the part of u it gathers has arrived, and the part of v no later chunk touches is copied back while the next
// chunks run.  PCIe is full duplex, so the step costs about one transfer instead of two plus the kernels.  Which part of u a chunk
// needs / which part of v it completes comes from the offsets (prefix maximum / suffix minimum of the L-indices per chunk): element
// orders with locality (lexicographic, space-filling curves) stream well, a random order degenerates to copy-then-apply -- always
// correct.  Results are bitwise those of ceedb200_operator_apply.  Falls back to the plain apply (returns *streamed = 0) whenever a
// precondition does not hold: not fused, other scatter mode, u already on the device, pageable host memory, partial coverage, ...
namespace {
bool is_pinned_host(const void *p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeHost;
}
}  // namespace

extern "C" int ceedb200_operator_apply_streamed(B200Operator op, B200Vector u, B200Vector v, int num_chunks, int *streamed) {
  B200Ceed ceed = op->ceed;
  if (streamed) *streamed = 0;
  B200_CALL(operator_setup(op));
  B200OpPlan *plan = op->plan;
  // ---- preconditions
  B200Restriction rin = nullptr, rout = nullptr;
  bool            ok  = plan->fused && !b200_compile_only() && plan->scatter_mode == B200_SCATTER_DETERMINISTIC && u && v && u != v &&
            u != B200_VECTOR_NONE && v != B200_VECTOR_NONE && !(ceed->autotune && !op->tuned);
  for (size_t i = 0; ok && i < op->in_fields.size(); i++) {
    const B200OpField &f = op->in_fields[i];
    if (!f.is_active || !f.rstr) continue;
    ok  = ok && !f.rstr->is_strided && (!rin || rin == f.rstr);
    rin = f.rstr;
  }
  for (size_t i = 0; ok && i < op->out_fields.size(); i++) {
    const B200OpField &f = op->out_fields[i];
    ok   = ok && f.is_active && !f.rstr->is_strided && (!rout || rout == f.rstr);
    rout = f.rstr;
  }
  ok = ok && rin && rout && rout->split_elem < 0 && u->h_array && !u->d_array && (v->h_borrowed || v->h_owned) && u->length == rin->l_size && v->length == rout->l_size;
  ok = ok && rout->num_elem >= 4096 && is_pinned_host(u->h_array) && is_pinned_host(v->h_borrowed ? v->h_borrowed : v->h_owned);
  if (ok) {
    B200_CALL(b200_restriction_build_owner(rout));
    ok = rout->num_nodes * rout->num_comp == v->length;  // every entry of v is written: no zero-first pass
  }
  if (!ok) return operator_apply(op, u, v, 0);
  B200_CUDA(ceed, cudaSetDevice(ceed->device_id));
  // ---- chunk tables (cached in the plan)
  const int K  = std::max(2, std::min(num_chunks > 0 ? num_chunks : default_parts(), 64));
  build_stream_plan(plan, rin, rout, K);
  B200StreamPlan &sp = plan->stream;
  B200_CALL(b200_restriction_set_parts(rout, sp.ends));
  // ---- streams and events
  if (!ceed->s_h2d) {
    B200_CUDA(ceed, cudaStreamCreateWithFlags(&ceed->s_h2d, cudaStreamNonBlocking));
    B200_CUDA(ceed, cudaStreamCreateWithFlags(&ceed->s_d2h, cudaStreamNonBlocking));
  }
  while ((int)ceed->ev_stream.size() < 2 * K + 2) {
    cudaEvent_t ev;
    B200_CUDA(ceed, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    ceed->ev_stream.push_back(ev);
  }
  const double *h_u = nullptr;
  double       *d_u = nullptr, *h_v = nullptr;
  B200_CALL(b200_vector_streamed_input(u, &h_u, &d_u));
  // range [b, e) of L-indices of one component block (or of the whole vector) -> byte ranges to copy
  const int     ncb    = sp.per_comp ? rin->num_comp : 1;
  const int64_t in_len = sp.per_comp ? rin->comp_stride : rin->l_size, out_len = sp.per_comp ? rout->comp_stride : rout->l_size;
  const int64_t in_top = sp.per_comp ? 0 : (int64_t)(rin->num_comp - 1) * rin->comp_stride;  // one range: the other components sit above the offsets
  // everything queued on the copy-in stream must follow what the caller already queued on the backend's stream (e.g. a previous apply)
  cudaEvent_t ev_start = ceed->ev_stream[2 * K], ev_end = ceed->ev_stream[2 * K + 1];
  B200_CUDA(ceed, cudaEventRecord(ev_start, ceed->stream));
  B200_CUDA(ceed, cudaStreamWaitEvent(ceed->s_h2d, ev_start, 0));
  int64_t copied = 0;
  for (int c = 0; c < K; c++) {
    const int64_t upto = c == K - 1 ? in_len : std::min(in_len, sp.in_hi[c] + in_top);
    if (upto > copied)
      for (int cc = 0; cc < ncb; cc++)
        B200_CUDA(ceed, cudaMemcpyAsync(d_u + cc * rin->comp_stride + copied, h_u + cc * rin->comp_stride + copied, (upto - copied) * sizeof(double),
                                        cudaMemcpyHostToDevice, ceed->s_h2d));
    copied = std::max(copied, upto);
    B200_CUDA(ceed, cudaEventRecord(ceed->ev_stream[c], ceed->s_h2d));
  }
  int64_t done = 0;
  int     err  = B200_SUCCESS;
  double *d_v  = nullptr;
  for (int c = 0; c < K && !err; c++) {
    B200_CUDA(ceed, cudaStreamWaitEvent(ceed->stream, ceed->ev_stream[c], 0));
    err = apply_fused(op, u, v, 0, c + 1);
    if (err) break;
    if (c == 0) {
      d_v = v->d_array;
      B200_CALL(b200_vector_streamed_output(v, &h_v));
    }
    B200_CUDA(ceed, cudaEventRecord(ceed->ev_stream[K + c], ceed->stream));
    B200_CUDA(ceed, cudaStreamWaitEvent(ceed->s_d2h, ceed->ev_stream[K + c], 0));
    const int64_t upto = c == K - 1 ? out_len : std::min(out_len, sp.out_done[c]);
    if (upto > done)
      for (int cc = 0; cc < ncb; cc++)
        B200_CUDA(ceed, cudaMemcpyAsync(h_v + cc * rout->comp_stride + done, d_v + cc * rout->comp_stride + done, (upto - done) * sizeof(double),
                                        cudaMemcpyDeviceToHost, ceed->s_d2h));
    done = std::max(done, upto);
  }
  // the result is on the host when the call returns (the caller's CeedVectorSyncArray(HOST) is then a no-op), and later work on the
  // backend's stream is ordered after the copies
  B200_CUDA(ceed, cudaEventRecord(ev_end, ceed->s_d2h));
  B200_CUDA(ceed, cudaStreamWaitEvent(ceed->stream, ev_end, 0));
  B200_CUDA(ceed, cudaStreamSynchronize(ceed->s_d2h));
  B200_CUDA(ceed, cudaStreamSynchronize(ceed->s_h2d));
  if (err) return err;
  B200_CALL(b200_vector_streamed_output(v, &h_v));  // (every part re-marked the host side stale when it took the device array)
  if (streamed) *streamed = 1;
  return B200_SUCCESS;
}

extern "C" int ceedb200_operator_is_fused(B200Operator op, int *is_fused) {
  B200_CALL(operator_setup(op));
  *is_fused = op->plan->fused ? 1 : 0;
  return B200_SUCCESS;
}

extern "C" const char *ceedb200_operator_kernel_source(B200Operator op) {
  if (operator_setup(op) || !op->plan->fused) return "";
  if (b200_opgen_build(op, op->plan, 0)) return "";
  return op->plan->variant[0].source.c_str();
}

extern "C" int ceedb200_operator_kernel_info(B200Operator op, int *regs, int *smem_bytes, int *threads, int *elems_per_block, int *grid,
                                             int *local_bytes) {
  B200_CALL(operator_setup(op));
  B200_CHECK(op->plan->fused, op->ceed, B200_ERROR_UNSUPPORTED, "operator is not fused: %s", op->plan->why_not_fused.c_str());
  B200KernelVariant *var = op->plan->variant[0].built ? &op->plan->variant[0] : &op->plan->variant[1];
  if (!var->built) {
    B200_CALL(b200_opgen_build(op, op->plan, 0));
    var = &op->plan->variant[0];
  }
  if (regs) *regs = var->regs;
  if (local_bytes) *local_bytes = var->local_bytes;
  if (smem_bytes) *smem_bytes = op->plan->smem_bytes;
  if (threads) *threads = op->plan->threads;
  if (elems_per_block) *elems_per_block = op->plan->epb;
  if (grid) *grid = op->plan->grid;
  return B200_SUCCESS;
}

extern "C" int ceedb200_operator_set_timing(B200Operator op, int enabled) {
  op->timing = enabled != 0;
  return B200_SUCCESS;
}
extern "C" int ceedb200_operator_last_kernel_ms(B200Operator op, float *fused_ms, float *aux_ms) {
  if (fused_ms) *fused_ms = op->last_fused_ms;
  if (aux_ms) *aux_ms = op->last_aux_ms;
  return B200_SUCCESS;
}
