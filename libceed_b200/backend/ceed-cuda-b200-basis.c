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
  if (u == CEED_VECTOR_NONE) CeedCheck(eval_mode == CEED_EVAL_WEIGHT, ceed, CEED_ERROR_BACKEND, "An input vector is required for this CeedEvalMode");
  {
    // arrays acquired below are restored whatever happens in between (libCEED's access locks); the first error is reported
    int  status = CEED_ERROR_SUCCESS, ierr2;
    bool got_u = false, got_v = false;

    if (u != CEED_VECTOR_NONE) {
      status = CeedVectorGetArrayRead(u, CEED_MEM_DEVICE, &d_u);
      got_u  = !status;
    }
    if (!status) {
      status = apply_add ? CeedVectorGetArray(v, CEED_MEM_DEVICE, &d_v) : CeedVectorGetArrayWrite(v, CEED_MEM_DEVICE, &d_v);
      got_v  = !status;
    }
    if (!status && ceedb200_basis_apply_ptr(impl->core, num_elem, t_mode, eval_mode, apply_add, d_u, d_v))
      status = CeedError(ceed, CEED_ERROR_BACKEND, "%s", ceedb200_last_error(core));
    if (got_u && (ierr2 = CeedVectorRestoreArrayRead(u, &d_u)) && !status) status = ierr2;
    if (got_v && (ierr2 = CeedVectorRestoreArray(v, &d_v)) && !status) status = ierr2;
    return status;
  }
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

// CeedBasisCreateH1 (interface/ceed-basis.c:1434; reference GPU twin backends/cuda-ref/ceed-cuda-ref-basis.c:340-400): non-tensor
// H1 bases for mixed-topology / composite operators; operators using them run on the backend's unfused kernels.
int CeedBasisCreateH1_B200(CeedElemTopology topo, CeedInt dim, CeedInt num_nodes, CeedInt num_qpts, const CeedScalar *interp, const CeedScalar *grad,
                           const CeedScalar *q_ref, const CeedScalar *q_weight, CeedBasis basis) {
  Ceed            ceed = CeedBasisReturnCeed(basis);
  B200Ceed        core;
  CeedBasis_B200 *impl;
  CeedInt         num_comp;

  CeedCallBackend(CeedGetCore_B200(ceed, &core));
  CeedCallBackend(CeedBasisGetNumComponents(basis, &num_comp));
  impl = calloc(1, sizeof(*impl));
  CeedCallB200(ceed, core, ceedb200_basis_create_h1(core, dim, num_comp, num_nodes, num_qpts, interp, grad, q_ref, q_weight, &impl->core));
  CeedCallBackend(CeedBasisSetData(basis, impl));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Basis", basis, "Apply", CeedBasisApply_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Basis", basis, "ApplyAdd", CeedBasisApplyAdd_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Basis", basis, "Destroy", CeedBasisDestroy_B200));
  return CEED_ERROR_SUCCESS;
}
