// ceed-cuda-b200-assemble.c -- CeedOperatorLinearAssembleAddDiagonal / ...AddPointBlockDiagonal for the b200 backend.
//
// The interface's default implementation (interface/ceed-preconditioning.c:213-425) fills the element diagonals in the CPU
// E-vector layout [elem][comp][node]; this backend's E-vectors are [comp][elem][node] (like the reference GPU backends, which
// ship their own diagonal assembly for the same reason: backends/cuda-ref/ceed-cuda-ref-operator.c:1225-1520).  This file is
// the layout-aware twin: diag(B_out^T D B_in) per element from the assembled QFunction (computed on the device by
// ceedb200_operator_assemble_qfunction), written through the E-layout the restriction reports, then summed into the
// L-vector with the restriction's own (deterministic) transpose.  The per-element arithmetic is setup-time host work, as in the
// interface's version; the hot path is not involved.
#include <stdbool.h>
#include <stdlib.h>

#include "ceed-cuda-b200.h"

// basis matrix (num_qpts x num_nodes, row-major) of one evaluation-mode entry; `which` counts the entries of the same mode
// (the d-th derivative of CEED_EVAL_GRAD); identity for CEED_EVAL_NONE
static int EvalMatrix_B200(CeedBasis basis, CeedEvalMode mode, CeedInt which, const CeedScalar *identity, CeedInt num_qpts, CeedInt num_nodes,
                           const CeedScalar **mat) {
  switch (mode) {
    case CEED_EVAL_NONE: *mat = identity; break;
    case CEED_EVAL_INTERP: CeedCallBackend(CeedBasisGetInterp(basis, mat)); break;
    case CEED_EVAL_GRAD:
      CeedCallBackend(CeedBasisGetGrad(basis, mat));
      *mat += (size_t)which * num_qpts * num_nodes;
      break;
    default: return CeedError(CeedBasisReturnCeed(basis), CEED_ERROR_UNSUPPORTED, "Backend does not implement diagonal assembly for %s", CeedEvalModes[mode]);
  }
  return CEED_ERROR_SUCCESS;
}

static int CeedOperatorAssembleAddDiagonalCore_B200(CeedOperator op, CeedVector assembled, CeedRequest *request, bool point_block) {
  Ceed                     ceed = CeedOperatorReturnCeed(op);
  CeedVector               qf_vec  = NULL;
  CeedElemRestriction      qf_rstr = NULL;
  const CeedScalar        *qf;
  CeedInt                  lq[3];
  const CeedEvalMode     **modes_in, **modes_out;
  CeedInt                  nb_in, nb_out, *nmodes_in, *nmodes_out;
  CeedSize               **off_in, **off_out, ncomp_out_total;
  CeedBasis               *bases_in, *bases_out;
  CeedElemRestriction     *rstrs_in, *rstrs_out;
  CeedOperatorAssemblyData data;

  CeedCallBackend(CeedOperatorLinearAssembleQFunctionBuildOrUpdate(op, &qf_vec, &qf_rstr, request));
  CeedCallBackend(CeedElemRestrictionGetELayout(qf_rstr, lq));
  CeedCallBackend(CeedElemRestrictionDestroy(&qf_rstr));
  CeedCallBackend(CeedOperatorGetOperatorAssemblyData(op, &data));
  CeedCallBackend(CeedOperatorAssemblyDataGetEvalModes(data, &nb_in, &nmodes_in, &modes_in, &off_in, &nb_out, &nmodes_out, &modes_out, &off_out, &ncomp_out_total));
  CeedCallBackend(CeedOperatorAssemblyDataGetBases(data, NULL, &bases_in, NULL, NULL, &bases_out, NULL));
  CeedCallBackend(CeedOperatorAssemblyDataGetElemRestrictions(data, NULL, &rstrs_in, NULL, &rstrs_out));
  CeedCallBackend(CeedVectorGetArrayRead(qf_vec, CEED_MEM_HOST, &qf));

  for (CeedInt bi = 0; bi < nb_in; bi++) {
    CeedInt             bo, num_elem, num_nodes, num_qpts, num_comp, ld[3];
    CeedElemRestriction diag_rstr;
    CeedVector          elem_diag;
    CeedScalar         *d, *identity = NULL;

    // the diagonal only sees (input, output) pairs on the same basis
    for (bo = 0; bo < nb_out; bo++)
      if (bases_out[bo] == bases_in[bi]) break;
    if (bo == nb_out) continue;
    CeedCheck(rstrs_in[bi] == rstrs_out[bo], ceed, CEED_ERROR_UNSUPPORTED,
              "Cannot assemble operator diagonal with different input and output active element restrictions");
    if (point_block) CeedCallBackend(CeedOperatorCreateActivePointBlockRestriction(rstrs_in[bi], &diag_rstr));
    else CeedCallBackend(CeedElemRestrictionCreateUnsignedCopy(rstrs_in[bi], &diag_rstr));
    CeedCallBackend(CeedElemRestrictionCreateVector(diag_rstr, NULL, &elem_diag));
    CeedCallBackend(CeedElemRestrictionGetELayout(diag_rstr, ld));  // index of (node, component, element) in an E-vector of THIS backend
    CeedCallBackend(CeedElemRestrictionGetNumElements(diag_rstr, &num_elem));
    CeedCallBackend(CeedBasisGetNumComponents(bases_in[bi], &num_comp));
    if (bases_in[bi] == CEED_BASIS_NONE) {
      CeedCallBackend(CeedElemRestrictionGetElementSize(rstrs_in[bi], &num_nodes));
      num_qpts = num_nodes;
    } else {
      CeedCallBackend(CeedBasisGetNumNodes(bases_in[bi], &num_nodes));
      CeedCallBackend(CeedBasisGetNumQuadraturePoints(bases_in[bi], &num_qpts));
    }
    identity = calloc((size_t)num_qpts * num_nodes, sizeof(CeedScalar));
    for (CeedInt i = 0; i < (num_nodes < num_qpts ? num_nodes : num_qpts); i++) identity[(size_t)i * num_nodes + i] = 1.0;
    CeedCallBackend(CeedVectorSetValue(elem_diag, 0.0));
    CeedCallBackend(CeedVectorGetArray(elem_diag, CEED_MEM_HOST, &d));

    // every (output mode, input mode) pair contributes  sum_q Bt[q][n] * D[q] * B[q][n]  to node n
    CeedInt which_out = 0;
    for (CeedInt mo = 0; mo < nmodes_out[bo]; mo++) {
      const CeedScalar *Bt;
      CeedInt           which_in = 0;

      which_out = (mo > 0 && modes_out[bo][mo] == modes_out[bo][mo - 1]) ? which_out + 1 : 0;
      CeedCallBackend(EvalMatrix_B200(bases_out[bo], modes_out[bo][mo], which_out, identity, num_qpts, num_nodes, &Bt));
      for (CeedInt mi = 0; mi < nmodes_in[bi]; mi++) {
        const CeedScalar *B;

        which_in = (mi > 0 && modes_in[bi][mi] == modes_in[bi][mi - 1]) ? which_in + 1 : 0;
        CeedCallBackend(EvalMatrix_B200(bases_in[bi], modes_in[bi][mi], which_in, identity, num_qpts, num_nodes, &B));
        for (CeedInt c_out = 0; c_out < num_comp; c_out++) {
          for (CeedInt c_in = point_block ? 0 : c_out; c_in < (point_block ? num_comp : c_out + 1); c_in++) {
            const CeedSize d_comp  = point_block ? (CeedSize)c_out * num_comp + c_in : c_out;  // component index inside the diagonal E-vector
            const CeedSize qf_comp = (off_in[bi][mi] + c_in) * ncomp_out_total + off_out[bo][mo] + c_out;

            for (CeedInt e = 0; e < num_elem; e++) {
              for (CeedInt q = 0; q < num_qpts; q++) {
                const CeedScalar w = qf[(CeedSize)q * lq[0] + qf_comp * lq[1] + (CeedSize)e * lq[2]];

                if (w == 0.0) continue;
                for (CeedInt n = 0; n < num_nodes; n++)
                  d[(CeedSize)n * ld[0] + d_comp * ld[1] + (CeedSize)e * ld[2]] += Bt[(size_t)q * num_nodes + n] * w * B[(size_t)q * num_nodes + n];
              }
            }
          }
        }
      }
    }
    CeedCallBackend(CeedVectorRestoreArray(elem_diag, &d));
    CeedCallBackend(CeedElemRestrictionApply(diag_rstr, CEED_TRANSPOSE, elem_diag, assembled, request));
    free(identity);
    CeedCallBackend(CeedVectorDestroy(&elem_diag));
    CeedCallBackend(CeedElemRestrictionDestroy(&diag_rstr));
  }
  CeedCallBackend(CeedVectorRestoreArrayRead(qf_vec, &qf));
  CeedCallBackend(CeedVectorDestroy(&qf_vec));
  return CEED_ERROR_SUCCESS;
}

int CeedOperatorLinearAssembleAddDiagonal_B200(CeedOperator op, CeedVector assembled, CeedRequest *request) {
  return CeedOperatorAssembleAddDiagonalCore_B200(op, assembled, request, false);
}
int CeedOperatorLinearAssembleAddPointBlockDiagonal_B200(CeedOperator op, CeedVector assembled, CeedRequest *request) {
  return CeedOperatorAssembleAddDiagonalCore_B200(op, assembled, request, true);
}
