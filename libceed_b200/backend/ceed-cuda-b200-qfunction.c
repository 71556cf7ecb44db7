// ceed-cuda-b200-qfunction.c -- CeedQFunction and CeedQFunctionContext slots -> ceedb200_qfunction_* / ceedb200_qfcontext_*
// (replaces the wiring of backends/cuda-ref/ceed-cuda-ref-qfunction.c:21-110 and ceed-cuda-ref-qfunctioncontext.c:152-333).
// The user's QFunction source (path + function name, interface/ceed-qfunction.c:259-320) is handed to the core, which
// JIT-compiles it with NVRTC either into the fused operator kernel or into a standalone apply kernel.
#include <stdlib.h>
#include <string.h>

#include "ceed-cuda-b200.h"

// ------------------------------------------------------------------------------------------------ context
static int CtxCore(CeedQFunctionContext ctx, B200QFContext *core) {
  CeedQFunctionContext_B200 *impl;

  CeedCallBackend(CeedQFunctionContextGetBackendData(ctx, &impl));
  *core = impl->core;
  return CEED_ERROR_SUCCESS;
}

#define CTX_CALL(ctx, ...)                                                            \
  do {                                                                                \
    B200Ceed core_ceed_;                                                              \
    CeedCallBackend(CeedGetCore_B200(CeedQFunctionContextReturnCeed(ctx), &core_ceed_)); \
    CeedCallB200(CeedQFunctionContextReturnCeed(ctx), core_ceed_, __VA_ARGS__);       \
  } while (0)

static int CeedQFunctionContextHasValidData_B200(CeedQFunctionContext ctx, bool *has_valid_data) {
  B200QFContext c;
  int           flag;

  CeedCallBackend(CtxCore(ctx, &c));
  CTX_CALL(ctx, ceedb200_qfcontext_has_valid_data(c, &flag));
  *has_valid_data = flag;
  return CEED_ERROR_SUCCESS;
}

static int CeedQFunctionContextHasBorrowedDataOfType_B200(CeedQFunctionContext ctx, CeedMemType mem_type, bool *has_borrowed) {
  B200QFContext c;
  int           flag;

  CeedCallBackend(CtxCore(ctx, &c));
  CTX_CALL(ctx, ceedb200_qfcontext_has_borrowed_data_of_type(c, mem_type, &flag));
  *has_borrowed = flag;
  return CEED_ERROR_SUCCESS;
}

static int CeedQFunctionContextSetData_B200(CeedQFunctionContext ctx, CeedMemType mem_type, CeedCopyMode copy_mode, void *data) {
  B200QFContext c;
  size_t        size;

  CeedCallBackend(CtxCore(ctx, &c));
  CeedCallBackend(CeedQFunctionContextGetContextSize(ctx, &size));
  CTX_CALL(ctx, ceedb200_qfcontext_set_data(c, mem_type, copy_mode, size, data));
  return CEED_ERROR_SUCCESS;
}

static int CeedQFunctionContextTakeData_B200(CeedQFunctionContext ctx, CeedMemType mem_type, void *data) {
  B200QFContext c;

  CeedCallBackend(CtxCore(ctx, &c));
  CTX_CALL(ctx, ceedb200_qfcontext_take_data(c, mem_type, (void **)data));
  return CEED_ERROR_SUCCESS;
}

static int CeedQFunctionContextGetData_B200(CeedQFunctionContext ctx, CeedMemType mem_type, void *data) {
  B200QFContext c;

  CeedCallBackend(CtxCore(ctx, &c));
  CTX_CALL(ctx, ceedb200_qfcontext_get_data(c, mem_type, (void **)data));
  return CEED_ERROR_SUCCESS;
}

static int CeedQFunctionContextGetDataRead_B200(CeedQFunctionContext ctx, CeedMemType mem_type, void *data) {
  B200QFContext c;

  CeedCallBackend(CtxCore(ctx, &c));
  CTX_CALL(ctx, ceedb200_qfcontext_get_data_read(c, mem_type, (void **)data));
  return CEED_ERROR_SUCCESS;
}

static int CeedQFunctionContextDestroy_B200(CeedQFunctionContext ctx) {
  CeedQFunctionContext_B200 *impl;

  CeedCallBackend(CeedQFunctionContextGetBackendData(ctx, &impl));
  ceedb200_qfcontext_destroy(impl->core);
  free(impl);
  return CEED_ERROR_SUCCESS;
}

int CeedQFunctionContextCreate_B200(CeedQFunctionContext ctx) {
  Ceed                       ceed = CeedQFunctionContextReturnCeed(ctx);
  B200Ceed                   core;
  CeedQFunctionContext_B200 *impl;

  CeedCallBackend(CeedGetCore_B200(ceed, &core));
  impl = calloc(1, sizeof(*impl));
  CeedCallB200(ceed, core, ceedb200_qfcontext_create(core, &impl->core));
  CeedCallBackend(CeedQFunctionContextSetBackendData(ctx, impl));
  CeedCallBackend(CeedSetBackendFunction(ceed, "QFunctionContext", ctx, "HasValidData", CeedQFunctionContextHasValidData_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "QFunctionContext", ctx, "HasBorrowedDataOfType", CeedQFunctionContextHasBorrowedDataOfType_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "QFunctionContext", ctx, "SetData", CeedQFunctionContextSetData_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "QFunctionContext", ctx, "TakeData", CeedQFunctionContextTakeData_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "QFunctionContext", ctx, "GetData", CeedQFunctionContextGetData_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "QFunctionContext", ctx, "GetDataRead", CeedQFunctionContextGetDataRead_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "QFunctionContext", ctx, "Destroy", CeedQFunctionContextDestroy_B200));
  return CEED_ERROR_SUCCESS;
}

// ------------------------------------------------------------------------------------------------ qfunction
// Mirror the (by now immutable) field list and the context of the interface object into the core object.
int CeedQFunctionGetCore_B200(CeedQFunction qf, B200QFunction *core_qf) {
  Ceed                 ceed = CeedQFunctionReturnCeed(qf);
  B200Ceed             core;
  CeedQFunction_B200  *impl;
  CeedQFunctionContext ctx;

  CeedCallBackend(CeedGetCore_B200(ceed, &core));
  CeedCallBackend(CeedQFunctionGetData(qf, &impl));
  if (!impl->fields_set) {
    CeedInt             num_in, num_out;
    CeedQFunctionField *in, *out;

    CeedCallBackend(CeedSyncJitOptions_B200(ceed));
    CeedCallBackend(CeedQFunctionGetFields(qf, &num_in, &in, &num_out, &out));
    for (CeedInt i = 0; i < num_in + num_out; i++) {
      const char  *name;
      CeedInt      size;
      CeedEvalMode eval_mode;

      CeedCallBackend(CeedQFunctionFieldGetData(i < num_in ? in[i] : out[i - num_in], &name, &size, &eval_mode));
      if (i < num_in) CeedCallB200(ceed, core, ceedb200_qfunction_add_input(impl->core, name, size, eval_mode));
      else CeedCallB200(ceed, core, ceedb200_qfunction_add_output(impl->core, name, size, eval_mode));
    }
    impl->fields_set = true;
  }
  CeedCallBackend(CeedQFunctionGetInnerContext(qf, &ctx));
  impl->foreign_ctx = NULL;
  if (ctx && CeedQFunctionContextReturnCeed(ctx) != ceed) {
    // context created on another Ceed (e.g. by the interface on the fallback Ceed, interface/ceed-preconditioning.c:3333): its
    // backend data is not ours; CeedQFunctionContextAcquire_B200 fetches the device pointer through the interface per apply
    impl->foreign_ctx = ctx;
    ceedb200_qfunction_set_context_ptr(impl->core, NULL);
  } else if (ctx) {
    B200QFContext c;

    CeedCallBackend(CtxCore(ctx, &c));  // borrowed reference: CeedQFunctionGetInnerContext does not add one (interface/ceed-qfunction.c:422-437)
    ceedb200_qfunction_set_context(impl->core, c);
  } else {
    ceedb200_qfunction_set_context(impl->core, NULL);
  }
  *core_qf = impl->core;
  return CEED_ERROR_SUCCESS;
}

// Foreign contexts: device pointer through the interface for the duration of one apply
int CeedQFunctionContextAcquire_B200(CeedQFunction qf, void **held) {
  CeedQFunction_B200 *impl;

  CeedCallBackend(CeedQFunctionGetData(qf, &impl));
  *held = NULL;
  if (impl->foreign_ctx) {
    CeedCallBackend(CeedQFunctionContextGetDataRead(impl->foreign_ctx, CEED_MEM_DEVICE, held));
    ceedb200_qfunction_set_context_ptr(impl->core, *held);
  }
  return CEED_ERROR_SUCCESS;
}
int CeedQFunctionContextRelease_B200(CeedQFunction qf, void **held) {
  CeedQFunction_B200 *impl;

  CeedCallBackend(CeedQFunctionGetData(qf, &impl));
  if (impl->foreign_ctx && *held) CeedCallBackend(CeedQFunctionContextRestoreDataRead(impl->foreign_ctx, held));
  return CEED_ERROR_SUCCESS;
}

static int CeedQFunctionApply_B200(CeedQFunction qf, CeedInt Q, CeedVector *U, CeedVector *V) {
  Ceed              ceed = CeedQFunctionReturnCeed(qf);
  B200Ceed          core;
  B200QFunction     core_qf;
  CeedInt           num_in, num_out;
  const CeedScalar *d_in[CEED_FIELD_MAX];
  CeedScalar       *d_out[CEED_FIELD_MAX];

  CeedCallBackend(CeedGetCore_B200(ceed, &core));
  CeedCallBackend(CeedQFunctionGetCore_B200(qf, &core_qf));
  CeedCallBackend(CeedQFunctionGetNumArgs(qf, &num_in, &num_out));
  {
    // the first error is recorded; every array and the context acquired before it are restored before it is reported
    int     status = CEED_ERROR_SUCCESS, ierr2;
    CeedInt got_in = 0, got_out = 0;
    void   *held = NULL;
    bool    ctx_held = false;

    for (CeedInt i = 0; i < num_in && !status; i++) {
      status = CeedVectorGetArrayRead(U[i], CEED_MEM_DEVICE, &d_in[i]);
      if (!status) got_in = i + 1;
    }
    for (CeedInt i = 0; i < num_out && !status; i++) {
      status = CeedVectorGetArrayWrite(V[i], CEED_MEM_DEVICE, &d_out[i]);
      if (!status) got_out = i + 1;
    }
    if (!status) {
      status   = CeedQFunctionContextAcquire_B200(qf, &held);
      ctx_held = !status;
    }
    if (!status && ceedb200_qfunction_apply_ptr(core_qf, Q, d_in, d_out)) status = CeedError(ceed, CEED_ERROR_BACKEND, "%s", ceedb200_last_error(core));
    if (ctx_held && (ierr2 = CeedQFunctionContextRelease_B200(qf, &held)) && !status) status = ierr2;
    for (CeedInt i = 0; i < got_in; i++)
      if ((ierr2 = CeedVectorRestoreArrayRead(U[i], &d_in[i])) && !status) status = ierr2;
    for (CeedInt i = 0; i < got_out; i++)
      if ((ierr2 = CeedVectorRestoreArray(V[i], &d_out[i])) && !status) status = ierr2;
    return status;
  }
}

static int CeedQFunctionDestroy_B200(CeedQFunction qf) {
  CeedQFunction_B200 *impl;

  CeedCallBackend(CeedQFunctionGetData(qf, &impl));
  ceedb200_qfunction_destroy(impl->core);
  free(impl);
  return CEED_ERROR_SUCCESS;
}

int CeedQFunctionCreate_B200(CeedQFunction qf) {
  Ceed                ceed = CeedQFunctionReturnCeed(qf);
  B200Ceed            core;
  CeedQFunction_B200 *impl;
  const char         *source_path, *kernel_name;

  CeedCallBackend(CeedGetCore_B200(ceed, &core));
  CeedCallBackend(CeedQFunctionGetSourcePath(qf, &source_path));
  CeedCallBackend(CeedQFunctionGetKernelName(qf, &kernel_name));
  CeedCheck(source_path && kernel_name, ceed, CEED_ERROR_BACKEND, "No QFunction source or CUfunction provided.");
  impl = calloc(1, sizeof(*impl));
  CeedCallB200(ceed, core, ceedb200_qfunction_create(core, source_path, kernel_name, &impl->core));
  CeedCallBackend(CeedQFunctionSetData(qf, impl));
  CeedCallBackend(CeedSetBackendFunction(ceed, "QFunction", qf, "Apply", CeedQFunctionApply_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "QFunction", qf, "Destroy", CeedQFunctionDestroy_B200));
  return CEED_ERROR_SUCCESS;
}
