// ceed-cuda-b200-basis.c -- tensor H1 CeedBasis slots -> ceedb200_basis_*
// (replaces the wiring of backends/cuda-shared/ceed-cuda-shared-basis.c:603-666)
#include <stdlib.h>

#include "ceed-cuda-b200.h"

static int CeedBasisApplyCore_B200(CeedBasis basis, bool apply_add, CeedInt num_elem, CeedTransposeMode t_mode, CeedEvalMode eval_mode, CeedVector u,
                                   CeedVector v) {
  Ceed              ceed = CeedBasisReturnCeed(basis);
  B200Ceed          core;
  CeedBasis_B200   *impl;
  const CeedScalar *d_u = NULL;
  CeedScalar       *d_v;

  CeedCallBackend(CeedGetCore_B200(ceed, &core));
  CeedCallBackend(CeedBasisGetData(basis, &impl));
  CeedCheck(eval_mode == CEED_EVAL_INTERP || eval_mode == CEED_EVAL_GRAD || eval_mode == CEED_EVAL_WEIGHT, ceed, CEED_ERROR_UNSUPPORTED,
            "Backend does not implement %s for tensor H1 bases", CeedEvalModes[eval_mode]);
  if (u != CEED_VECTOR_NONE) CeedCallBackend(CeedVectorGetArrayRead(u, CEED_MEM_DEVICE, &d_u));
  else CeedCheck(eval_mode == CEED_EVAL_WEIGHT, ceed, CEED_ERROR_BACKEND, "An input vector is required for this CeedEvalMode");
  if (apply_add) CeedCallBackend(CeedVectorGetArray(v, CEED_MEM_DEVICE, &d_v));
  else CeedCallBackend(CeedVectorGetArrayWrite(v, CEED_MEM_DEVICE, &d_v));
  CeedCallB200(ceed, core, ceedb200_basis_apply_ptr(impl->core, num_elem, t_mode, eval_mode, apply_add, d_u, d_v));
  if (u != CEED_VECTOR_NONE) CeedCallBackend(CeedVectorRestoreArrayRead(u, &d_u));
  CeedCallBackend(CeedVectorRestoreArray(v, &d_v));
  return CEED_ERROR_SUCCESS;
}

static int CeedBasisApply_B200(CeedBasis basis, CeedInt num_elem, CeedTransposeMode t_mode, CeedEvalMode eval_mode, CeedVector u, CeedVector v) {
  return CeedBasisApplyCore_B200(basis, false, num_elem, t_mode, eval_mode, u, v);
}
static int CeedBasisApplyAdd_B200(CeedBasis basis, CeedInt num_elem, CeedTransposeMode t_mode, CeedEvalMode eval_mode, CeedVector u, CeedVector v) {
  return CeedBasisApplyCore_B200(basis, true, num_elem, t_mode, eval_mode, u, v);
}

static int CeedBasisDestroy_B200(CeedBasis basis) {
  CeedBasis_B200 *impl;

  CeedCallBackend(CeedBasisGetData(basis, &impl));
  ceedb200_basis_destroy(impl->core);
  free(impl);
  return CEED_ERROR_SUCCESS;
}

int CeedBasisCreateTensorH1_B200(CeedInt dim, CeedInt P_1d, CeedInt Q_1d, const CeedScalar *interp_1d, const CeedScalar *grad_1d,
                                 const CeedScalar *q_ref_1d, const CeedScalar *q_weight_1d, CeedBasis basis) {
  Ceed            ceed = CeedBasisReturnCeed(basis);
  B200Ceed        core;
  CeedBasis_B200 *impl;
  CeedInt         num_comp;

  CeedCallBackend(CeedGetCore_B200(ceed, &core));
  CeedCallBackend(CeedBasisGetNumComponents(basis, &num_comp));
  impl = calloc(1, sizeof(*impl));
  CeedCallB200(ceed, core, ceedb200_basis_create_tensor_h1(core, dim, num_comp, P_1d, Q_1d, interp_1d, grad_1d, q_ref_1d, q_weight_1d, &impl->core));
  CeedCallBackend(CeedBasisSetData(basis, impl));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Basis", basis, "Apply", CeedBasisApply_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Basis", basis, "ApplyAdd", CeedBasisApplyAdd_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Basis", basis, "Destroy", CeedBasisDestroy_B200));
  return CEED_ERROR_SUCCESS;
}
