// bp_geo.h -- setup QFunctions for the CEED bake-off problems: geometric factors at quadrature points.
// Definitions follow examples/petsc/qfunctions/bps/bp1.h:26-50 (mass, 1 component: w detJ) and bp3.h:38-72
// (diffusion, 7 components: w detJ, then the upper triangle of w adj(J) adj(J)^T / detJ in row order 00 01 02 11 12 22).
// Fields: in[0] = x (INTERP, 3, unused), in[1] = dx/dX (GRAD, 3x3: J[d][c] = d x_c / d X_d), in[2] = quadrature weight.
#ifndef CEED_B200_BP_GEO_H
#define CEED_B200_BP_GEO_H
#include <ceed/types.h>

// adjugate A of the 3x3 Jacobian stored as J[(r * 3 + s) * Q + i]; returns det(J)
CEED_QFUNCTION_HELPER CeedScalar BPAdjugate3(const CeedScalar *J, const CeedInt Q, const CeedInt i, CeedScalar A[3][3]) {
  for (CeedInt r = 0; r < 3; r++) {
    const CeedInt r1 = (r + 1) % 3, r2 = (r + 2) % 3;
    for (CeedInt s = 0; s < 3; s++) {
      const CeedInt s1 = (s + 1) % 3, s2 = (s + 2) % 3;
      A[r][s] = J[(r1 * 3 + s1) * Q + i] * J[(r2 * 3 + s2) * Q + i] - J[(r1 * 3 + s2) * Q + i] * J[(r2 * 3 + s1) * Q + i];
    }
  }
  return J[0 * Q + i] * A[0][0] + J[1 * Q + i] * A[0][1] + J[2 * Q + i] * A[0][2];
}

CEED_QFUNCTION(BPSetupMassGeo)(void *ctx, const CeedInt Q, const CeedScalar *const *in, CeedScalar *const *out) {
  const CeedScalar *J = in[1], *w = in[2];
  CeedScalar       *qd = out[0];
  CeedPragmaSIMD for (CeedInt i = 0; i < Q; i++) {
    CeedScalar A[3][3];
    qd[i] = BPAdjugate3(J, Q, i, A) * w[i];
  }
  return 0;
}

CEED_QFUNCTION(BPSetupDiffGeo)(void *ctx, const CeedInt Q, const CeedScalar *const *in, CeedScalar *const *out) {
  const CeedScalar *J = in[1], *w = in[2];
  CeedScalar       *qd = out[0];
  CeedPragmaSIMD for (CeedInt i = 0; i < Q; i++) {
    CeedScalar       A[3][3];
    const CeedScalar detJ = BPAdjugate3(J, Q, i, A);
    const CeedScalar s    = w[i] / detJ;
    qd[i]                 = w[i] * detJ;
    CeedInt n             = 1;
    for (CeedInt r = 0; r < 3; r++)
      for (CeedInt c = r; c < 3; c++) qd[i + Q * (n++)] = s * (A[r][0] * A[c][0] + A[r][1] * A[c][1] + A[r][2] * A[c][2]);
  }
  return 0;
}
#endif
