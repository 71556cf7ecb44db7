// bp_apply.h -- apply QFunctions of the CEED bake-off problems BP1-BP6.
//   BPMass / BPMass3 : v = qdata * u                  (examples/petsc/qfunctions/bps/bp1.h:77-86, bp2.h)
//   BPDiff / BPDiff3 : dv = D du, D symmetric 3x3 from qdata[1..6] (row order 00 01 02 11 12 22)
//                                                      (examples/petsc/qfunctions/bps/bp3.h:105-128, bp4.h:59-84)
// Gradient layout for 3 components: ug[comp + 3 * direction] (doc/sphinx/source/libCEEDdev.md:113-118).
#ifndef CEED_B200_BP_APPLY_H
#define CEED_B200_BP_APPLY_H
#include <ceed/types.h>

CEED_QFUNCTION(BPMass)(void *ctx, const CeedInt Q, const CeedScalar *const *in, CeedScalar *const *out) {
  const CeedScalar *u = in[0], *qd = in[1];
  CeedScalar       *v = out[0];
  CeedPragmaSIMD for (CeedInt i = 0; i < Q; i++) v[i] = qd[i] * u[i];
  return 0;
}

CEED_QFUNCTION(BPMass3)(void *ctx, const CeedInt Q, const CeedScalar *const *in, CeedScalar *const *out) {
  const CeedScalar(*u)[CEED_Q_VLA] = (const CeedScalar(*)[CEED_Q_VLA])in[0];
  const CeedScalar *qd             = in[1];
  CeedScalar(*v)[CEED_Q_VLA]       = (CeedScalar(*)[CEED_Q_VLA])out[0];
  CeedPragmaSIMD for (CeedInt i = 0; i < Q; i++) {
    for (CeedInt c = 0; c < 3; c++) v[c][i] = qd[i] * u[c][i];
  }
  return 0;
}

CEED_QFUNCTION(BPDiff)(void *ctx, const CeedInt Q, const CeedScalar *const *in, CeedScalar *const *out) {
  const CeedScalar(*ug)[CEED_Q_VLA] = (const CeedScalar(*)[CEED_Q_VLA])in[0];
  const CeedScalar(*qd)[CEED_Q_VLA] = (const CeedScalar(*)[CEED_Q_VLA])in[1];
  CeedScalar(*vg)[CEED_Q_VLA]       = (CeedScalar(*)[CEED_Q_VLA])out[0];
  CeedPragmaSIMD for (CeedInt i = 0; i < Q; i++) {
    const CeedScalar D[3][3] = {
        {qd[1][i], qd[2][i], qd[3][i]},
        {qd[2][i], qd[4][i], qd[5][i]},
        {qd[3][i], qd[5][i], qd[6][i]}
    };
    for (CeedInt j = 0; j < 3; j++) vg[j][i] = ug[0][i] * D[0][j] + ug[1][i] * D[1][j] + ug[2][i] * D[2][j];
  }
  return 0;
}

CEED_QFUNCTION(BPDiff3)(void *ctx, const CeedInt Q, const CeedScalar *const *in, CeedScalar *const *out) {
  const CeedScalar(*ug)[3][CEED_Q_VLA] = (const CeedScalar(*)[3][CEED_Q_VLA])in[0];  // [direction][component]
  const CeedScalar(*qd)[CEED_Q_VLA]    = (const CeedScalar(*)[CEED_Q_VLA])in[1];
  CeedScalar(*vg)[3][CEED_Q_VLA]       = (CeedScalar(*)[3][CEED_Q_VLA])out[0];
  CeedPragmaSIMD for (CeedInt i = 0; i < Q; i++) {
    const CeedScalar D[3][3] = {
        {qd[1][i], qd[2][i], qd[3][i]},
        {qd[2][i], qd[4][i], qd[5][i]},
        {qd[3][i], qd[5][i], qd[6][i]}
    };
    for (CeedInt c = 0; c < 3; c++)
      for (CeedInt j = 0; j < 3; j++) vg[j][c][i] = ug[0][c][i] * D[0][j] + ug[1][c][i] * D[1][j] + ug[2][c][i] * D[2][j];
  }
  return 0;
}
#endif
