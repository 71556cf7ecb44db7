// bp_geo.h -- setup QFunctions for the CEED bake-off problems: geometric factors at quadrature points.
// Definitions follow examples/petsc/qfunctions/bps/bp1.h:26-50 (mass, 1 component: w detJ) and bp3.h:38-72
// (diffusion, 7 components: w detJ, then the upper triangle of w adj(J) adj(J)^T / detJ in row order 00 01 02 11 12 22).
// Fields: in[0] = x (INTERP, 3, unused), in[1] = dx/dX (GRAD, 3x3: J[d][c] = d x_c / d X_d), in[2] = quadrature weight.
#ifndef CEED_B200_BP_GEO_H
#define CEED_B200_BP_GEO_H
#include <ceed/types.h>

CEED_QFUNCTION_HELPER CeedScalar BPAdjugate3(const CeedScalar (*J)[3][CEED_Q_VLA], CeedInt i, CeedScalar A[3][3]) {
  for (CeedInt r = 0; r < 3; r++) {
    const CeedInt r1 = (r + 1) % 3, r2 = (r + 2) % 3;
    for (CeedInt s = 0; s < 3; s++) {
      const CeedInt s1 = (s + 1) % 3, s2 = (s + 2) % 3;
      A[r][s] = J[r1][s1][i] * J[r2][s2][i] - J[r1][s2][i] * J[r2][s1][i];
    }
  }
  return J[0][0][i] * A[0][0] + J[0][1][i] * A[0][1] + J[0][2][i] * A[0][2];
}

CEED_QFUNCTION(BPSetupMassGeo)(void *ctx, const CeedInt Q, const CeedScalar *const *in, CeedScalar *const *out) {
  const CeedScalar(*J)[3][CEED_Q_VLA] = (const CeedScalar(*)[3][CEED_Q_VLA])in[1];
  const CeedScalar *w                 = in[2];
  CeedScalar       *qd                = out[0];
  CeedPragmaSIMD for (CeedInt i = 0; i < Q; i++) {
    CeedScalar A[3][3];
    qd[i] = BPAdjugate3(J, i, A) * w[i];
  }
  return 0;
}

CEED_QFUNCTION(BPSetupDiffGeo)(void *ctx, const CeedInt Q, const CeedScalar *const *in, CeedScalar *const *out) {
  const CeedScalar(*J)[3][CEED_Q_VLA] = (const CeedScalar(*)[3][CEED_Q_VLA])in[1];
  const CeedScalar *w                 = in[2];
  CeedScalar(*qd)[CEED_Q_VLA]         = (CeedScalar(*)[CEED_Q_VLA])out[0];
  CeedPragmaSIMD for (CeedInt i = 0; i < Q; i++) {
    CeedScalar       A[3][3];
    const CeedScalar detJ = BPAdjugate3(J, i, A);
    const CeedScalar s    = w[i] / detJ;
    qd[0][i]              = w[i] * detJ;
    CeedInt n             = 1;
    for (CeedInt r = 0; r < 3; r++)
      for (CeedInt c = r; c < 3; c++) qd[n++][i] = s * (A[r][0] * A[c][0] + A[r][1] * A[c][1] + A[r][2] * A[c][2]);
  }
  return 0;
}
#endif
