// ceed-cuda-b200-operator.c -- CeedOperator slots -> ceedb200_operator_*
//
// Replaces the host side of backends/cuda-gen/ceed-cuda-gen-operator.c:105-300,879-908.  On the first apply the field
// wiring of the interface object (CeedOperatorGetFields, interface/ceed-operator.c:931-1038) is mirrored into a core
// operator; every apply then fetches the DEVICE arrays of the active and passive vectors through the interface
// (CeedVectorGetArray*(CEED_MEM_DEVICE), so lazy host->device synchronisation and access locking stay with libCEED, as in
// ceed-cuda-gen-operator.c:131-171) and hands them to the core as borrowed device pointers.
// Both ApplyAdd (mandatory, interface/ceed-operator.c:2335) and Apply (overwrite; lets the fused kernel store instead of
// memset + read-modify-write, :2280-2282) are registered.
#include <stdlib.h>
#include <string.h>

#include "ceed-cuda-b200.h"

static int CeedOperatorSetup_B200(CeedOperator op) {
  Ceed               ceed = CeedOperatorReturnCeed(op);
  B200Ceed           core;
  CeedOperator_B200 *impl;
  CeedQFunction      qf;
  B200QFunction      core_qf;
  CeedInt            num_in, num_out;
  CeedOperatorField *in, *out;

  CeedCallBackend(CeedOperatorGetData(op, &impl));
  if (impl->is_setup) return CEED_ERROR_SUCCESS;
  {
    // objects created by the delegate backend (CEED_B200_FALLBACK): the whole operator goes through its fallback operator
    Ceed_B200 *data;

    CeedCallBackend(CeedGetData(ceed, &data));
    if (data->has_fallback) {
      CeedCallBackend(CeedOperatorGetFields(op, &num_in, &in, &num_out, &out));
      for (CeedInt i = 0; i < num_in + num_out; i++) {
        CeedElemRestriction rstr;
        CeedBasis           basis;

        CeedCallBackend(CeedOperatorFieldGetElemRestriction(i < num_in ? in[i] : out[i - num_in], &rstr));
        CeedCallBackend(CeedOperatorFieldGetBasis(i < num_in ? in[i] : out[i - num_in], &basis));
        if (rstr != CEED_ELEMRESTRICTION_NONE && CeedElemRestrictionReturnCeed(rstr) != ceed) impl->use_fallback = true;
        if (basis != CEED_BASIS_NONE && CeedBasisReturnCeed(basis) != ceed) impl->use_fallback = true;
        CeedCallBackend(CeedElemRestrictionDestroy(&rstr));
        CeedCallBackend(CeedBasisDestroy(&basis));
      }
      if (impl->use_fallback) {
        impl->is_setup = true;
        CeedCallBackend(CeedOperatorSetSetupDone(op));
        return CEED_ERROR_SUCCESS;
      }
    }
  }
  CeedCallBackend(CeedGetCore_B200(ceed, &core));
  CeedCallBackend(CeedSyncJitOptions_B200(ceed));
  CeedCallBackend(CeedOperatorGetQFunction(op, &qf));
  CeedCallBackend(CeedQFunctionGetCore_B200(qf, &core_qf));
  CeedCallBackend(CeedQFunctionDestroy(&qf));
  CeedCallB200(ceed, core, ceedb200_operator_create(core, core_qf, &impl->core));
  CeedCallB200(ceed, core, ceedb200_vector_create(core, 0, &impl->view_in));
  CeedCallB200(ceed, core, ceedb200_vector_create(core, 0, &impl->view_out));
  impl->view_in_len = impl->view_out_len = 0;
  CeedCallBackend(CeedOperatorGetFields(op, &num_in, &in, &num_out, &out));
  impl->num_in  = num_in;
  impl->num_out = num_out;
  for (CeedInt i = 0; i < num_in + num_out; i++) {
    const bool          is_in = i < num_in;
    CeedOperatorField   field = is_in ? in[i] : out[i - num_in];
    const char         *name;
    CeedElemRestriction rstr;
    CeedBasis           basis;
    CeedVector          vec;
    B200Restriction     core_rstr  = NULL;
    B200Basis           core_basis = NULL;
    B200Vector          core_vec   = B200_VECTOR_NONE;

    CeedCallBackend(CeedOperatorFieldGetData(field, &name, &rstr, &basis, &vec));
    if (rstr != CEED_ELEMRESTRICTION_NONE) {
      CeedElemRestriction_B200 *r;
      CeedCallBackend(CeedElemRestrictionGetData(rstr, &r));
      core_rstr = r->core;
    }
    if (basis != CEED_BASIS_NONE) {
      CeedBasis_B200 *b;
      CeedCallBackend(CeedBasisGetData(basis, &b));
      core_basis = b->core;
    }
    if (vec == CEED_VECTOR_ACTIVE) {
      core_vec = B200_VECTOR_ACTIVE;
    } else if (vec != CEED_VECTOR_NONE) {
      // passive vector: a zero-length core vector that borrows the device array at every apply
      B200Vector *slot = is_in ? &impl->passive_in[i] : &impl->passive_out[i - num_in];
      CeedSize    len;
      CeedCallBackend(CeedVectorGetLength(vec, &len));
      CeedCallB200(ceed, core, ceedb200_vector_create(core, len, slot));
      core_vec = *slot;
      if (is_in) impl->passive_in_vec[i] = vec;
      else impl->passive_out_vec[i - num_in] = vec;
      CeedCallBackend(CeedVectorReference(vec));  // keep alive as long as the operator
    }
    CeedCallB200(ceed, core, ceedb200_operator_set_field(impl->core, name, core_rstr, core_basis, core_vec));
    CeedCallBackend(CeedElemRestrictionDestroy(&rstr));
    CeedCallBackend(CeedBasisDestroy(&basis));
    CeedCallBackend(CeedVectorDestroy(&vec));
  }
  impl->is_setup = true;
  CeedCallBackend(CeedOperatorSetSetupDone(op));
  return CEED_ERROR_SUCCESS;
}

static int CeedOperatorApplyCore_B200(CeedOperator op, CeedVector in_vec, CeedVector out_vec, bool add) {
  Ceed               ceed = CeedOperatorReturnCeed(op);
  B200Ceed           core;
  CeedOperator_B200 *impl;
  const CeedScalar  *d_in = NULL, *d_pin[CEED_FIELD_MAX] = {NULL};
  CeedScalar        *d_out = NULL, *d_pout[CEED_FIELD_MAX] = {NULL};
  CeedSize           len;

  CeedCallBackend(CeedGetCore_B200(ceed, &core));
  CeedCallBackend(CeedOperatorSetup_B200(op));
  CeedCallBackend(CeedOperatorGetData(op, &impl));
  if (impl->use_fallback) {
    CeedOperator       op_fallback;
    CeedInt            num_in, num_out;
    CeedOperatorField *in, *out;

    CeedCallBackend(CeedOperatorGetFallback(op, &op_fallback));
    CeedCheck(op_fallback, ceed, CEED_ERROR_UNSUPPORTED, "Backend does not implement operators with objects of other backends");
    if (!add) {  // Apply = zero the outputs, then ApplyAdd (interface/ceed-operator.c:2357-2405)
      CeedCallBackend(CeedOperatorGetFields(op, &num_in, &in, &num_out, &out));
      for (CeedInt i = 0; i < num_out; i++) {
        CeedVector vec;

        CeedCallBackend(CeedOperatorFieldGetVector(out[i], &vec));
        if (vec != CEED_VECTOR_ACTIVE && vec != CEED_VECTOR_NONE) CeedCallBackend(CeedVectorSetValue(vec, 0.0));
        CeedCallBackend(CeedVectorDestroy(&vec));
      }
      if (out_vec != CEED_VECTOR_NONE) CeedCallBackend(CeedVectorSetValue(out_vec, 0.0));
    }
    return CeedOperatorApplyAdd(op_fallback, in_vec, out_vec, CEED_REQUEST_IMMEDIATE);
  }

  // Host-resident application (CeedVectorSetArray(HOST) -> CeedOperatorApply -> CeedVectorSyncArray(HOST)): the input is valid on the
  // host only and the output has a caller-owned host array.  The core then pipelines copy-in, apply and copy-out chunk by chunk
  // (ceedb200_operator_apply_streamed) instead of the reference's whole-vector copies around the apply; it falls back to the plain apply
  // by itself when a precondition does not hold.  Only when both vectors belong to this backend and there are no passive fields to bind
  // differently (their device arrays are bound below as usual).
  bool try_streamed = false;
  if (!add && in_vec != CEED_VECTOR_NONE && out_vec != CEED_VECTOR_NONE && in_vec != out_vec && !getenv("CEED_B200_NO_STREAMED") &&
      CeedVectorReturnCeed(in_vec) == ceed && CeedVectorReturnCeed(out_vec) == ceed) {
    CeedVector_B200 *vi, *vo;
    int              h_valid = 0, d_valid = 0, out_borrowed = 0;
    bool             passive_out = false;

    CeedCallBackend(CeedVectorGetData(in_vec, &vi));
    CeedCallBackend(CeedVectorGetData(out_vec, &vo));
    CeedCallB200(ceed, core, ceedb200_vector_valid_sides(vi->core, &h_valid, &d_valid));
    CeedCallB200(ceed, core, ceedb200_vector_has_borrowed_array_of_type(vo->core, B200_MEM_HOST, &out_borrowed));
    for (CeedInt i = 0; i < impl->num_out; i++) passive_out = passive_out || impl->passive_out_vec[i];
    try_streamed = h_valid && !d_valid && out_borrowed && !passive_out;
  }

  // Device arrays of every vector involved, through the interface.  From here on no early return: the first error is recorded,
  // every array that was acquired is restored (libCEED's access locks must be released whatever happens) and the error is
  // reported at the end.
  {
    int  status = CEED_ERROR_SUCCESS, core_err = 0;
    bool got_in = false, got_out = false, got_pin[CEED_FIELD_MAX] = {false}, got_pout[CEED_FIELD_MAX] = {false};
    bool has_passive_out = false, ctx_held = false;
    CeedQFunction qf = NULL;
    void         *held_ctx = NULL;
#define B200_TRY(call)                \
  do {                                \
    if (!status) status = (call);     \
  } while (0)
#define B200_TRY_CORE(call)                  \
  do {                                       \
    if (!status && !core_err) core_err = (call); \
  } while (0)

    if (try_streamed) {
      // interface bookkeeping only (access locks, state counters): the host arrays are valid, nothing is copied here
      B200_TRY(CeedVectorGetArrayRead(in_vec, CEED_MEM_HOST, &d_in));
      got_in = !status;
      B200_TRY(CeedVectorGetArrayWrite(out_vec, CEED_MEM_HOST, &d_out));
      got_out = !status;
    }
    if (in_vec != CEED_VECTOR_NONE && !try_streamed) {
      B200_TRY(CeedVectorGetLength(in_vec, &len));
      B200_TRY(CeedVectorGetArrayRead(in_vec, CEED_MEM_DEVICE, &d_in));
      got_in = !status;
      if (!status && impl->view_in_len != len) {  // the view vectors are cached: re-created only when the length changes
        ceedb200_vector_destroy(impl->view_in);
        impl->view_in = NULL;
        B200_TRY_CORE(ceedb200_vector_create(core, len, &impl->view_in));
        impl->view_in_len = len;
      }
      B200_TRY_CORE(ceedb200_vector_set_array(impl->view_in, B200_MEM_DEVICE, B200_USE_POINTER, (CeedScalar *)d_in));
    }
    if (out_vec != CEED_VECTOR_NONE && !try_streamed) {
      B200_TRY(CeedVectorGetLength(out_vec, &len));
      if (add) B200_TRY(CeedVectorGetArray(out_vec, CEED_MEM_DEVICE, &d_out));
      else B200_TRY(CeedVectorGetArrayWrite(out_vec, CEED_MEM_DEVICE, &d_out));
      got_out = !status;
      if (!status && impl->view_out_len != len) {
        ceedb200_vector_destroy(impl->view_out);
        impl->view_out = NULL;
        B200_TRY_CORE(ceedb200_vector_create(core, len, &impl->view_out));
        impl->view_out_len = len;
      }
      B200_TRY_CORE(ceedb200_vector_set_array(impl->view_out, B200_MEM_DEVICE, B200_USE_POINTER, d_out));
    }
    for (CeedInt i = 0; i < impl->num_in; i++) {
      if (!impl->passive_in_vec[i] || status || core_err) continue;
      B200_TRY(CeedVectorGetArrayRead(impl->passive_in_vec[i], CEED_MEM_DEVICE, &d_pin[i]));
      got_pin[i] = !status;
      B200_TRY_CORE(ceedb200_vector_set_array(impl->passive_in[i], B200_MEM_DEVICE, B200_USE_POINTER, (CeedScalar *)d_pin[i]));
    }
    for (CeedInt i = 0; i < impl->num_out; i++) {
      if (!impl->passive_out_vec[i]) continue;
      has_passive_out = true;
      if (status || core_err) continue;
      // passive outputs are overwritten by Apply (zero first, as CeedOperatorApplyAddActive does, interface/ceed-operator.c:2357-2405)
      // and accumulated into by ApplyAdd
      if (!add) B200_TRY(CeedVectorSetValue(impl->passive_out_vec[i], 0.0));
      B200_TRY(CeedVectorGetArray(impl->passive_out_vec[i], CEED_MEM_DEVICE, &d_pout[i]));
      got_pout[i] = !status;
      B200_TRY_CORE(ceedb200_vector_set_array(impl->passive_out[i], B200_MEM_DEVICE, B200_USE_POINTER, d_pout[i]));
    }
    // a QFunction context of another backend is handed over as a raw device pointer for this apply
    B200_TRY(CeedOperatorGetQFunction(op, &qf));
    if (!status && !core_err) {
      B200_TRY(CeedQFunctionContextAcquire_B200(qf, &held_ctx));
      ctx_held = !status;
    }
    if (!status && !core_err && try_streamed) {
      CeedVector_B200 *vi, *vo;
      int              used = 0;

      B200_TRY(CeedVectorGetData(in_vec, &vi));
      B200_TRY(CeedVectorGetData(out_vec, &vo));
      if (!status) core_err = ceedb200_operator_apply_streamed(impl->core, vi->core, vo->core, 0, &used);
    } else if (!status && !core_err) {
      // overwrite semantics can only be used when every output is the active vector
      if (!add && !has_passive_out) core_err = ceedb200_operator_apply(impl->core, in_vec != CEED_VECTOR_NONE ? impl->view_in : NULL, impl->view_out);
      else {
        if (!add && out_vec != CEED_VECTOR_NONE) core_err = ceedb200_vector_set_value(impl->view_out, 0.0);
        if (!core_err) core_err = ceedb200_operator_apply_add(impl->core, in_vec != CEED_VECTOR_NONE ? impl->view_in : NULL, impl->view_out);
      }
    }
    // the core's message has to be fetched before other core calls overwrite it
    if (core_err && !status) status = CeedError(ceed, CEED_ERROR_BACKEND, "%s", ceedb200_last_error(core));
    // release / restore everything that was acquired, whatever happened above
    {
      int ierr2;

      if (ctx_held && (ierr2 = CeedQFunctionContextRelease_B200(qf, &held_ctx)) && !status) status = ierr2;
      if (qf && (ierr2 = CeedQFunctionDestroy(&qf)) && !status) status = ierr2;
      if (got_in && (ierr2 = CeedVectorRestoreArrayRead(in_vec, &d_in)) && !status) status = ierr2;
      if (got_out && (ierr2 = CeedVectorRestoreArray(out_vec, &d_out)) && !status) status = ierr2;
      for (CeedInt i = 0; i < impl->num_in; i++)
        if (got_pin[i] && (ierr2 = CeedVectorRestoreArrayRead(impl->passive_in_vec[i], &d_pin[i])) && !status) status = ierr2;
      for (CeedInt i = 0; i < impl->num_out; i++)
        if (got_pout[i] && (ierr2 = CeedVectorRestoreArray(impl->passive_out_vec[i], &d_pout[i])) && !status) status = ierr2;
    }
#undef B200_TRY
#undef B200_TRY_CORE
    return status;
  }
}

static int CeedOperatorApplyAdd_B200(CeedOperator op, CeedVector in_vec, CeedVector out_vec, CeedRequest *request) {
  return CeedOperatorApplyCore_B200(op, in_vec, out_vec, true);
}
static int CeedOperatorApply_B200(CeedOperator op, CeedVector in_vec, CeedVector out_vec, CeedRequest *request) {
  return CeedOperatorApplyCore_B200(op, in_vec, out_vec, false);
}

static int CeedOperatorDestroy_B200(CeedOperator op) {
  CeedOperator_B200 *impl;

  CeedCallBackend(CeedOperatorGetData(op, &impl));
  if (impl->core) ceedb200_operator_destroy(impl->core);
  ceedb200_vector_destroy(impl->view_in);
  ceedb200_vector_destroy(impl->view_out);
  for (CeedInt i = 0; i < CEED_FIELD_MAX; i++) {
    ceedb200_vector_destroy(impl->passive_in[i]);
    ceedb200_vector_destroy(impl->passive_out[i]);
    if (impl->passive_in_vec[i]) CeedCallBackend(CeedVectorDestroy(&impl->passive_in_vec[i]));
    if (impl->passive_out_vec[i]) CeedCallBackend(CeedVectorDestroy(&impl->passive_out_vec[i]));
  }
  free(impl);
  return CEED_ERROR_SUCCESS;
}

// CeedOperatorLinearAssembleQFunction / ...Update (interface/ceed-preconditioning.c:2234-2330; GPU-side contract of
// backends/cuda-ref/ceed-cuda-ref-operator.c:1000-1130).  The core evaluates the QFunction's pointwise linear map on the device
// (ceedb200_operator_assemble_qfunction); this adapter creates the strided restriction + vector that describe its layout and
// binds the device arrays.  With this slot the interface's own LinearAssembleDiagonal / LinearAssemble / multigrid run on the backend.
static int CeedOperatorLinearAssembleQFunctionCore_B200(CeedOperator op, bool build_objects, CeedVector *assembled, CeedElemRestriction *rstr) {
  Ceed               ceed = CeedOperatorReturnCeed(op);
  B200Ceed           core;
  CeedOperator_B200 *impl;
  b200_int           num_elem, num_qpts, size_in, size_out;
  const CeedScalar  *d_pin[CEED_FIELD_MAX] = {NULL};
  bool               got_pin[CEED_FIELD_MAX] = {false}, got_asm = false;
  CeedScalar        *d_asm = NULL;
  B200Vector         view  = NULL;
  int                status = CEED_ERROR_SUCCESS, core_err = 0;

  CeedCallBackend(CeedGetCore_B200(ceed, &core));
  CeedCallBackend(CeedOperatorSetup_B200(op));
  CeedCallBackend(CeedOperatorGetData(op, &impl));
  CeedCheck(!impl->use_fallback, ceed, CEED_ERROR_UNSUPPORTED, "Backend does not implement QFunction assembly with objects of other backends");
  CeedCallB200(ceed, core, ceedb200_operator_assemble_qfunction_sizes(impl->core, &num_elem, &num_qpts, &size_in, &size_out));
  if (build_objects) {
    Ceed           ceed_parent;
    const CeedSize l_size     = (CeedSize)num_elem * num_qpts * size_in * size_out;
    CeedInt        strides[3] = {1, num_elem * num_qpts, num_qpts};

    CeedCallBackend(CeedOperatorGetFallbackParentCeed(op, &ceed_parent));
    CeedCallBackend(CeedElemRestrictionCreateStrided(ceed_parent, num_elem, num_qpts, size_in * size_out, l_size, strides, rstr));
    CeedCallBackend(CeedVectorCreate(ceed_parent, l_size, assembled));
    CeedCallBackend(CeedDestroy(&ceed_parent));
  }
  // no early return from here on: whatever was acquired is restored
  for (CeedInt i = 0; i < impl->num_in && !status && !core_err; i++) {
    if (!impl->passive_in_vec[i]) continue;
    status     = CeedVectorGetArrayRead(impl->passive_in_vec[i], CEED_MEM_DEVICE, &d_pin[i]);
    got_pin[i] = !status;
    if (!status) core_err = ceedb200_vector_set_array(impl->passive_in[i], B200_MEM_DEVICE, B200_USE_POINTER, (CeedScalar *)d_pin[i]);
  }
  if (!status && !core_err) {
    CeedSize len;

    status = CeedVectorGetLength(*assembled, &len);
    if (!status) status = CeedVectorGetArrayWrite(*assembled, CEED_MEM_DEVICE, &d_asm);
    got_asm = !status;
    if (!status) core_err = ceedb200_vector_create(core, len, &view);
    if (!status && !core_err) core_err = ceedb200_vector_set_array(view, B200_MEM_DEVICE, B200_USE_POINTER, d_asm);
  }
  if (!status && !core_err) {
    CeedQFunction qf       = NULL;
    void         *held_ctx = NULL;

    status = CeedOperatorGetQFunction(op, &qf);
    if (!status) status = CeedQFunctionContextAcquire_B200(qf, &held_ctx);
    if (!status) {
      core_err = ceedb200_operator_assemble_qfunction(impl->core, view);
      if (core_err) status = CeedError(ceed, CEED_ERROR_BACKEND, "%s", ceedb200_last_error(core));
      int ierr2 = CeedQFunctionContextRelease_B200(qf, &held_ctx);
      if (ierr2 && !status) status = ierr2;
    }
    if (qf) CeedQFunctionDestroy(&qf);
  }
  if (core_err && !status) status = CeedError(ceed, CEED_ERROR_BACKEND, "%s", ceedb200_last_error(core));
  if (view) ceedb200_vector_destroy(view);
  if (got_asm) {
    int ierr2 = CeedVectorRestoreArray(*assembled, &d_asm);
    if (ierr2 && !status) status = ierr2;
  }
  for (CeedInt i = 0; i < impl->num_in; i++) {
    if (!got_pin[i]) continue;
    int ierr2 = CeedVectorRestoreArrayRead(impl->passive_in_vec[i], &d_pin[i]);
    if (ierr2 && !status) status = ierr2;
  }
  return status;
}

static int CeedOperatorLinearAssembleQFunction_B200(CeedOperator op, CeedVector *assembled, CeedElemRestriction *rstr, CeedRequest *request) {
  return CeedOperatorLinearAssembleQFunctionCore_B200(op, true, assembled, rstr);
}
static int CeedOperatorLinearAssembleQFunctionUpdate_B200(CeedOperator op, CeedVector assembled, CeedElemRestriction rstr, CeedRequest *request) {
  return CeedOperatorLinearAssembleQFunctionCore_B200(op, false, &assembled, &rstr);
}

int CeedOperatorCreate_B200(CeedOperator op) {
  Ceed               ceed = CeedOperatorReturnCeed(op);
  CeedOperator_B200 *impl;

  impl = calloc(1, sizeof(*impl));
  CeedCallBackend(CeedOperatorSetData(op, impl));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Operator", op, "ApplyAdd", CeedOperatorApplyAdd_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Operator", op, "Apply", CeedOperatorApply_B200));
  {
    Ceed_B200 *data;

    CeedCallBackend(CeedGetData(ceed, &data));
    if (!data->has_fallback) {  // with a fallback Ceed the interface routes assembly to the fallback operator by itself
      CeedCallBackend(CeedSetBackendFunction(ceed, "Operator", op, "LinearAssembleQFunction", CeedOperatorLinearAssembleQFunction_B200));
      CeedCallBackend(CeedSetBackendFunction(ceed, "Operator", op, "LinearAssembleQFunctionUpdate", CeedOperatorLinearAssembleQFunctionUpdate_B200));
      // layout-aware twins of the interface's diagonal assembly (ceed-cuda-b200-assemble.c)
      CeedCallBackend(CeedSetBackendFunction(ceed, "Operator", op, "LinearAssembleAddDiagonal", CeedOperatorLinearAssembleAddDiagonal_B200));
      CeedCallBackend(CeedSetBackendFunction(ceed, "Operator", op, "LinearAssembleAddPointBlockDiagonal", CeedOperatorLinearAssembleAddPointBlockDiagonal_B200));
    }
  }
  CeedCallBackend(CeedSetBackendFunction(ceed, "Operator", op, "Destroy", CeedOperatorDestroy_B200));
  return CEED_ERROR_SUCCESS;
}
