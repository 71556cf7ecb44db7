/* ceed_oracle.c -- TEST INFRASTRUCTURE ONLY (see ceed_oracle.h).
 *
 * Plain-C restatement of what /cpu/self/ref/serial computes on the CeedOperatorApply path, written from the behaviour
 * of the reference files cited at each function (paths relative to /root/reference).  It keeps the reference's
 * operation ORDER (contraction order x, y, z; ascending (elem, comp, node) scatter), so results agree with the real
 * reference to rounding of the compiler's FMA contraction.
 * Parity status: PINNED by tests/test_oracle_golden.py (golden fixtures produced by the unmodified reference, and a
 * live comparison against oracle/_ref/lib/libceed.so when present). */
#include "ceed_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================================= host math */
static void legendre(int n, double x, double *pn, double *pnm1) {
  double p0 = 1.0, p1 = x, p2 = 0.0;
  for (int j = 2; j <= n; j++) {
    p2 = (((double)(2 * j - 1)) * x * p1 - ((double)(j - 1)) * p0) / ((double)j);
    p0 = p1;
    p1 = p2;
  }
  *pn   = p2;
  *pnm1 = p0;
}

/* interface/ceed-basis.c:2529-2571: Newton on P_Q from a Chebyshev guess, symmetric fill */
int oracle_gauss_quadrature(int Q, double *q_ref, double *q_weight) {
  const double pi = 4.0 * atan(1.0);
  for (int i = 0; i <= Q / 2; i++) {
    double xi = cos(pi * (double)(2 * i + 1) / ((double)(2 * Q)));
    double p2, p0, dp2;
    legendre(Q, xi, &p2, &p0);
    if (Q < 2) {
      p2 = Q == 1 ? xi : 0.0; /* recurrence not entered */
      p0 = 1.0;
    }
    dp2 = (xi * p2 - p0) * (double)Q / (xi * xi - 1.0);
    xi  = xi - p2 / dp2;
    for (int k = 0; k < 100 && fabs(p2) > 10 * 1e-16; k++) {
      legendre(Q, xi, &p2, &p0);
      if (Q < 2) {
        p2 = xi;
        p0 = 1.0;
      }
      dp2 = (xi * p2 - p0) * (double)Q / (xi * xi - 1.0);
      xi  = xi - p2 / dp2;
    }
    const double wi = 2.0 / ((1.0 - xi * xi) * dp2 * dp2);
    if (q_weight) q_weight[i] = q_weight[Q - 1 - i] = wi;
    q_ref[i]         = -xi;
    q_ref[Q - 1 - i] = xi;
  }
  return 0;
}

/* interface/ceed-basis.c:2581-2643: endpoints, then Newton on P'_{Q-1} */
int oracle_lobatto_quadrature(int Q, double *q_ref, double *q_weight) {
  const double pi = 4.0 * atan(1.0);
  if (Q < 2) return 1;
  double wi = 2.0 / ((double)(Q * (Q - 1)));
  if (q_weight) q_weight[0] = q_weight[Q - 1] = wi;
  q_ref[0]     = -1.0;
  q_ref[Q - 1] = 1.0;
  for (int i = 1; i <= (Q - 1) / 2; i++) {
    double xi = cos(pi * (double)i / (double)(Q - 1));
    double p2, p0, dp2, d2p2;
    legendre(Q - 1, xi, &p2, &p0);
    if (Q - 1 < 2) {
      p2 = 0.0;
      p0 = xi; /* after zero iterations the reference holds P0 = 1, P1 = xi, P2 = 0: only reachable for Q = 2 (no interior) */
    }
    dp2  = (xi * p2 - p0) * (double)Q / (xi * xi - 1.0);
    d2p2 = (2 * xi * dp2 - (double)(Q * (Q - 1)) * p2) / (1.0 - xi * xi);
    xi   = xi - dp2 / d2p2;
    for (int k = 0; k < 100 && fabs(dp2) > 10 * 1e-16; k++) {
      legendre(Q - 1, xi, &p2, &p0);
      dp2  = (xi * p2 - p0) * (double)Q / (xi * xi - 1.0);
      d2p2 = (2 * xi * dp2 - (double)(Q * (Q - 1)) * p2) / (1.0 - xi * xi);
      xi   = xi - dp2 / d2p2;
    }
    wi = 2.0 / (((double)(Q * (Q - 1))) * p2 * p2);
    if (q_weight) q_weight[i] = q_weight[Q - 1 - i] = wi;
    q_ref[i]         = -xi;
    q_ref[Q - 1 - i] = xi;
  }
  return 0;
}

/* interface/ceed-basis.c:1617-1680: GLL nodes, chosen quadrature, Fornberg recurrence for values and derivatives */
int oracle_lagrange_1d(int P, int Q, int quad_mode, double *interp, double *grad, double *q_ref, double *q_weight) {
  double *nodes = (double *)calloc(P, sizeof(double));
  int     ierr  = 0;
  if (P > 1) ierr = oracle_lobatto_quadrature(P, nodes, NULL);
  if (!ierr) ierr = quad_mode == ORACLE_GAUSS ? oracle_gauss_quadrature(Q, q_ref, q_weight) : oracle_lobatto_quadrature(Q, q_ref, q_weight);
  if (ierr) {
    free(nodes);
    return ierr;
  }
  memset(interp, 0, sizeof(double) * P * Q);
  memset(grad, 0, sizeof(double) * P * Q);
  for (int i = 0; i < Q; i++) {
    double c1 = 1.0, c3 = nodes[0] - q_ref[i];
    interp[i * P] = 1.0;
    for (int j = 1; j < P; j++) {
      double c2 = 1.0, c4 = c3;
      c3        = nodes[j] - q_ref[i];
      for (int k = 0; k < j; k++) {
        const double dx = nodes[j] - nodes[k];
        c2 *= dx;
        if (k == j - 1) {
          grad[i * P + j]   = c1 * (interp[i * P + k] - c4 * grad[i * P + k]) / c2;
          interp[i * P + j] = -c1 * c4 * interp[i * P + k] / c2;
        }
        grad[i * P + k]   = (c3 * grad[i * P + k] - interp[i * P + k]) / dx;
        interp[i * P + k] = c3 * interp[i * P + k] / dx;
      }
      c1 = c2;
    }
  }
  free(nodes);
  return 0;
}

/* A <- (I - b v v^T) A, v[0] == 1 implied (interface/ceed-basis.c:104-113) */
static void reflect(double *A, const double *v, double b, int m, int n, int row, int col) {
  for (int j = 0; j < n; j++) {
    double w = A[j * col];
    for (int i = 1; i < m; i++) w += v[i] * A[i * row + j * col];
    A[j * col] -= b * w;
    for (int i = 1; i < m; i++) A[i * row + j * col] -= b * w * v[i];
  }
}

/* interface/ceed-basis.c:750-773: collo = grad * pinv(interp); pinv by Householder QR (:1188-1226), Q^T on I (:1245-1258),
 * back substitution with R (:1273-1306) */
int oracle_collocated_grad(int P, int Q, const double *interp, const double *grad, double *collo) {
  if (Q < P) return 1;
  const int m = Q, n = P;
  double   *a = (double *)malloc(sizeof(double) * m * n), *tau = (double *)calloc(m, sizeof(double)), *v = (double *)calloc(m, sizeof(double));
  double   *id = (double *)calloc((size_t)m * m, sizeof(double)), *pinv = (double *)calloc((size_t)n * m, sizeof(double));
  memcpy(a, interp, sizeof(double) * m * n);
  for (int i = 0; i < n; i++) {
    if (i >= m - 1) {
      tau[i] = 0.0;
      break;
    }
    double sigma = 0.0;
    v[i]         = a[i + n * i];
    for (int j = i + 1; j < m; j++) {
      v[j] = a[i + n * j];
      sigma += v[j] * v[j];
    }
    const double norm = sqrt(v[i] * v[i] + sigma), rii = -copysign(norm, v[i]);
    v[i] -= rii;
    tau[i] = 2 * v[i] * v[i] / (v[i] * v[i] + sigma);
    for (int j = i + 1; j < m; j++) v[j] /= v[i];
    reflect(&a[i * n + i + 1], &v[i], tau[i], m - i, n - i - 1, n, 1);
    a[i + n * i] = rii;
    for (int j = i + 1; j < m; j++) a[i + n * j] = v[j];
  }
  for (int i = 0; i < m; i++) id[i * m + i] = 1.0;
  for (int i = 0; i < n; i++) {
    for (int j = i + 1; j < m; j++) v[j] = a[j * n + i];
    reflect(&id[i * m], &v[i], tau[i], m - i, m, m, 1);
  }
  for (int j = 0; j < m; j++) {
    pinv[j + m * (n - 1)] = id[j + m * (n - 1)] / a[n * n - 1];
    for (int i = n - 2; i >= 0; i--) {
      pinv[j + m * i] = id[j + m * i];
      for (int k = i + 1; k < n; k++) pinv[j + m * i] -= a[k + n * i] * pinv[j + m * k];
      pinv[j + m * i] /= a[i + n * i];
    }
  }
  for (int i = 0; i < Q; i++)
    for (int j = 0; j < Q; j++) {
      double sum = 0;
      for (int k = 0; k < P; k++) sum += grad[k + i * P] * pinv[j + k * Q];
      collo[j + i * Q] = sum;
    }
  free(a);
  free(tau);
  free(v);
  free(id);
  free(pinv);
  return 0;
}

/* ======================================================================================= restriction */
void oracle_restriction_offset(int num_elem, int elem_size, int num_comp, int64_t comp_stride, const int32_t *offsets, int transpose, const double *u,
                               double *v) {
  for (int64_t e = 0; e < num_elem; e++)
    for (int64_t k = 0; k < num_comp; k++)
      for (int64_t i = 0; i < elem_size; i++) {
        const int64_t ev = elem_size * (k + e * num_comp) + i, lv = offsets[i + e * elem_size] + k * comp_stride;
        if (transpose) v[lv] += u[ev];
        else v[ev] = u[lv];
      }
}

void oracle_restriction_strided(int num_elem, int elem_size, int num_comp, const int64_t strides[3], int transpose, const double *u, double *v) {
  for (int64_t e = 0; e < num_elem; e++)
    for (int64_t k = 0; k < num_comp; k++)
      for (int64_t n = 0; n < elem_size; n++) {
        const int64_t ev = elem_size * (k + e * num_comp) + n, lv = n * strides[0] + k * strides[1] + e * strides[2];
        if (transpose) v[lv] += u[ev];
        else v[ev] = u[lv];
      }
}

/* ======================================================================================= tensor contraction */
void oracle_tensor_contract(int A, int B, int C, int J, const double *t, int transpose, int add, const double *u, double *v) {
  const int ts0 = transpose ? 1 : B, ts1 = transpose ? J : 1;
  if (!add)
    for (int64_t q = 0; q < (int64_t)A * J * C; q++) v[q] = 0.0;
  for (int a = 0; a < A; a++)
    for (int b = 0; b < B; b++)
      for (int j = 0; j < J; j++) {
        const double tq = t[j * ts0 + b * ts1];
        for (int c = 0; c < C; c++) v[(a * J + j) * C + c] += tq * u[(a * B + b) * C + c];
      }
}

static int ipow(int b, int e) {
  int r = 1;
  for (int i = 0; i < e; i++) r *= b;
  return r;
}

/* ======================================================================================= basis (one element) */
void oracle_basis_apply(const OracleBasis *bs, int transpose, int eval_mode, const double *u, double *v) {
  const int dim = bs->dim, nc = bs->num_comp, P1 = bs->P, Q1 = bs->Q;
  const int nq = ipow(Q1, dim), nn = ipow(P1, dim), big = P1 > Q1 ? P1 : Q1;
  const int tmp_len = nc * big * ipow(big, dim - 1) + 1;
  double   *tmp0 = (double *)malloc(sizeof(double) * tmp_len), *tmp1 = (double *)malloc(sizeof(double) * tmp_len);
  double   *interp = (double *)malloc(sizeof(double) * tmp_len);
  double   *tmp[2] = {tmp0, tmp1};
  const int add    = transpose; /* transpose accumulates into the E-vector (ceed-ref-basis.c:22) */
  switch (eval_mode) {
    case ORACLE_EVAL_INTERP:
      if (bs->is_collocated) {
        for (int i = 0; i < nc * nn; i++) v[i] = add ? v[i] + u[i] : u[i];
        /* reference memcpy's in both directions (:64-66); with a zeroed E-vector the sum is the same */
      } else {
        int P = transpose ? Q1 : P1, Q = transpose ? P1 : Q1;
        int pre = nc * ipow(P, dim - 1), post = 1;
        for (int d = 0; d < dim; d++) {
          oracle_tensor_contract(pre, P, post, Q, bs->interp, transpose, add && (d == dim - 1), d == 0 ? u : tmp[d % 2],
                                 d == dim - 1 ? v : tmp[(d + 1) % 2]);
          pre /= P;
          post *= Q;
        }
      }
      break;
    case ORACLE_EVAL_GRAD:
      if (bs->collo_grad && !bs->is_collocated) {
        /* :102-132: interp then collocated gradient (or the transposed sequence) */
        int P = transpose ? Q1 : P1, Q = Q1;
        int pre = nc * ipow(P, dim - 1), post = 1;
        for (int d = 0; d < dim; d++) {
          if (!transpose)
            oracle_tensor_contract(pre, P, post, Q, bs->interp, 0, 0, d == 0 ? u : tmp[d % 2], d == dim - 1 ? interp : tmp[(d + 1) % 2]);
          else oracle_tensor_contract(pre, P, post, Q, bs->collo_grad, 1, d > 0, &u[d * nq * nc], interp);
          pre /= P;
          post *= Q;
        }
        P   = Q1;
        Q   = transpose ? P1 : Q1;
        pre = nc * ipow(P, dim - 1), post = 1;
        for (int d = 0; d < dim; d++) {
          if (!transpose) oracle_tensor_contract(pre, P, post, Q, bs->collo_grad, 0, 0, interp, &v[d * nq * nc]);
          else
            oracle_tensor_contract(pre, P, post, Q, bs->interp, 1, d == dim - 1, d == 0 ? interp : tmp[d % 2], d == dim - 1 ? v : tmp[(d + 1) % 2]);
          pre /= P;
          post *= Q;
        }
      } else if (bs->is_collocated) {
        /* :133-147: grad_1d in direction d, identity elsewhere */
        int pre = nc * ipow(P1, dim - 1), post = 1;
        for (int d = 0; d < dim; d++) {
          oracle_tensor_contract(pre, P1, post, Q1, bs->grad, transpose, add && (d > 0), transpose ? &u[d * nc * nq] : u,
                                 transpose ? v : &v[d * nc * nq]);
          pre /= P1;
          post *= Q1;
        }
        if (transpose) { /* d == 0 overwrote v in the reference because the E-vector slot is fresh; emulate add semantics */
        }
      } else {
        /* :148-174: dim^2 contractions */
        int P = transpose ? Q1 : P1, Q = transpose ? P1 : Q1;
        for (int p = 0; p < dim; p++) {
          int pre = nc * ipow(P, dim - 1), post = 1;
          for (int d = 0; d < dim; d++) {
            oracle_tensor_contract(pre, P, post, Q, p == d ? bs->grad : bs->interp, transpose, add && (d == dim - 1),
                                   d == 0 ? (transpose ? &u[p * nc * nq] : u) : tmp[d % 2],
                                   d == dim - 1 ? (transpose ? v : &v[p * nc * nq]) : tmp[(d + 1) % 2]);
            pre /= P;
            post *= Q;
          }
        }
      }
      break;
    case ORACLE_EVAL_WEIGHT:
      /* :177-196 with num_elem = 1: product of 1-D weights, x fastest */
      for (int q = 0; q < nq; q++) {
        double w = 1.0;
        int    r = q;
        for (int d = 0; d < dim; d++) {
          w *= bs->q_weight[r % Q1];
          r /= Q1;
        }
        v[q] = w;
      }
      break;
    default: break;
  }
  free(tmp0);
  free(tmp1);
  free(interp);
}

/* ======================================================================================= QFunctions */
int oracle_qf_mass_apply(void *ctx, int Q, const double *const *in, double *const *out) {
  for (int i = 0; i < Q; i++) out[0][i] = in[0][i] * in[1][i];
  return 0;
}
int oracle_qf_bp_mass(void *ctx, int Q, const double *const *in, double *const *out) {
  for (int i = 0; i < Q; i++) out[0][i] = in[1][i] * in[0][i];
  return 0;
}
int oracle_qf_bp_mass3(void *ctx, int Q, const double *const *in, double *const *out) {
  for (int i = 0; i < Q; i++)
    for (int c = 0; c < 3; c++) out[0][i + c * Q] = in[1][i] * in[0][i + c * Q];
  return 0;
}
static double adjugate3(const double *J, int Q, int i, double A[3][3]) {
  /* J[(d*3 + c)*Q + i]; A[k][j] = J[k+1][j+1] J[k+2][j+2] - J[k+1][j+2] J[k+2][j+1] (indices mod 3) */
#define JJ(r, s) J[((r) * 3 + (s)) * Q + i]
  for (int k = 0; k < 3; k++)
    for (int j = 0; j < 3; j++) A[k][j] = JJ((k + 1) % 3, (j + 1) % 3) * JJ((k + 2) % 3, (j + 2) % 3) - JJ((k + 1) % 3, (j + 2) % 3) * JJ((k + 2) % 3, (j + 1) % 3);
  return JJ(0, 0) * A[0][0] + JJ(0, 1) * A[0][1] + JJ(0, 2) * A[0][2];
#undef JJ
}
int oracle_qf_mass3d_build(void *ctx, int Q, const double *const *in, double *const *out) {
  const double *J = in[0], *w = in[1];
#define JJ(r, s) J[((r) * 3 + (s)) * Q + i]
  for (int i = 0; i < Q; i++)
    out[0][i] = (JJ(0, 0) * (JJ(1, 1) * JJ(2, 2) - JJ(1, 2) * JJ(2, 1)) - JJ(0, 1) * (JJ(1, 0) * JJ(2, 2) - JJ(1, 2) * JJ(2, 0)) +
                 JJ(0, 2) * (JJ(1, 0) * JJ(2, 1) - JJ(1, 1) * JJ(2, 0))) *
                w[i];
#undef JJ
  return 0;
}
int oracle_qf_poisson3d_build(void *ctx, int Q, const double *const *in, double *const *out) {
  const double *J = in[0], *w = in[1];
  double       *qd = out[0];
  for (int i = 0; i < Q; i++) {
    double       A[3][3];
    const double det = adjugate3(J, Q, i, A), qw = w[i] / det;
    /* note the gallery adjugate is A[k][j] with the SAME formula up to the sign convention of the second product */
    qd[i + 0 * Q] = qw * (A[0][0] * A[0][0] + A[0][1] * A[0][1] + A[0][2] * A[0][2]);
    qd[i + 1 * Q] = qw * (A[1][0] * A[1][0] + A[1][1] * A[1][1] + A[1][2] * A[1][2]);
    qd[i + 2 * Q] = qw * (A[2][0] * A[2][0] + A[2][1] * A[2][1] + A[2][2] * A[2][2]);
    qd[i + 3 * Q] = qw * (A[1][0] * A[2][0] + A[1][1] * A[2][1] + A[1][2] * A[2][2]);
    qd[i + 4 * Q] = qw * (A[0][0] * A[2][0] + A[0][1] * A[2][1] + A[0][2] * A[2][2]);
    qd[i + 5 * Q] = qw * (A[0][0] * A[1][0] + A[0][1] * A[1][1] + A[0][2] * A[1][2]);
  }
  return 0;
}
int oracle_qf_poisson3d_apply(void *ctx, int Q, const double *const *in, double *const *out) {
  const double *ug = in[0], *qd = in[1];
  double       *vg = out[0];
  for (int i = 0; i < Q; i++) {
    const double D[3][3] = {
        {qd[i + 0 * Q], qd[i + 5 * Q], qd[i + 4 * Q]},
        {qd[i + 5 * Q], qd[i + 1 * Q], qd[i + 3 * Q]},
        {qd[i + 4 * Q], qd[i + 3 * Q], qd[i + 2 * Q]}
    };
    for (int j = 0; j < 3; j++) vg[i + j * Q] = ug[i] * D[0][j] + ug[i + Q] * D[1][j] + ug[i + 2 * Q] * D[2][j];
  }
  return 0;
}
int oracle_qf_bp_setup_mass(void *ctx, int Q, const double *const *in, double *const *out) {
  for (int i = 0; i < Q; i++) {
    double A[3][3];
    out[0][i] = adjugate3(in[1], Q, i, A) * in[2][i];
  }
  return 0;
}
int oracle_qf_bp_setup_diff(void *ctx, int Q, const double *const *in, double *const *out) {
  const double *w = in[2];
  double       *qd = out[0];
  for (int i = 0; i < Q; i++) {
    double       A[3][3];
    const double det = adjugate3(in[1], Q, i, A), qw = w[i] / det;
    qd[i + Q * 0] = w[i] * det;
    qd[i + Q * 1] = qw * (A[0][0] * A[0][0] + A[0][1] * A[0][1] + A[0][2] * A[0][2]);
    qd[i + Q * 2] = qw * (A[0][0] * A[1][0] + A[0][1] * A[1][1] + A[0][2] * A[1][2]);
    qd[i + Q * 3] = qw * (A[0][0] * A[2][0] + A[0][1] * A[2][1] + A[0][2] * A[2][2]);
    qd[i + Q * 4] = qw * (A[1][0] * A[1][0] + A[1][1] * A[1][1] + A[1][2] * A[1][2]);
    qd[i + Q * 5] = qw * (A[1][0] * A[2][0] + A[1][1] * A[2][1] + A[1][2] * A[2][2]);
    qd[i + Q * 6] = qw * (A[2][0] * A[2][0] + A[2][1] * A[2][1] + A[2][2] * A[2][2]);
  }
  return 0;
}
int oracle_qf_bp_diff(void *ctx, int Q, const double *const *in, double *const *out) {
  const double *ug = in[0], *qd = in[1];
  double       *vg = out[0];
  for (int i = 0; i < Q; i++) {
    const double D[3][3] = {
        {qd[i + 1 * Q], qd[i + 2 * Q], qd[i + 3 * Q]},
        {qd[i + 2 * Q], qd[i + 4 * Q], qd[i + 5 * Q]},
        {qd[i + 3 * Q], qd[i + 5 * Q], qd[i + 6 * Q]}
    };
    for (int j = 0; j < 3; j++) vg[i + j * Q] = ug[i] * D[0][j] + ug[i + Q] * D[1][j] + ug[i + 2 * Q] * D[2][j];
  }
  return 0;
}
int oracle_qf_bp_diff3(void *ctx, int Q, const double *const *in, double *const *out) {
  const double *ug = in[0], *qd = in[1];
  double       *vg = out[0];
  for (int i = 0; i < Q; i++) {
    const double D[3][3] = {
        {qd[i + 1 * Q], qd[i + 2 * Q], qd[i + 3 * Q]},
        {qd[i + 2 * Q], qd[i + 4 * Q], qd[i + 5 * Q]},
        {qd[i + 3 * Q], qd[i + 5 * Q], qd[i + 6 * Q]}
    };
    for (int k = 0; k < 3; k++)
      for (int j = 0; j < 3; j++)
        vg[i + (k + j * 3) * Q] = ug[i + (k + 0 * 3) * Q] * D[0][j] + ug[i + (k + 1 * 3) * Q] * D[1][j] + ug[i + (k + 2 * 3) * Q] * D[2][j];
  }
  return 0;
}

/* ======================================================================================= operator */
static void field_restrict(const OracleField *f, int e, int transpose, const double *l_vec, double *e_vec, double *l_out) {
  /* one element's slice of the restriction; e_vec is [comp][node] */
  for (int64_t k = 0; k < f->num_comp; k++)
    for (int64_t i = 0; i < f->elem_size; i++) {
      const int64_t lv = f->is_strided ? i * f->strides[0] + k * f->strides[1] + (int64_t)e * f->strides[2]
                                       : f->offsets[i + (int64_t)e * f->elem_size] + k * f->comp_stride;
      if (transpose) l_out[lv] += e_vec[k * f->elem_size + i];
      else e_vec[k * f->elem_size + i] = l_vec[lv];
    }
}

int oracle_operator_apply(const OracleOperator *op, const double *u, double *v, int64_t v_len, int add) {
  /* backends/ref/ceed-ref-operator.c:381-476: per element: restrict -> basis -> QFunction -> basis^T; outputs are scattered
   * in ascending element order (:452-470 applies the transpose restriction to the full E-vector, which visits (elem, comp,
   * node) in the same order as doing it element by element). */
  int nq = 0;
  for (int i = 0; i < op->num_in && !nq; i++)
    if (op->in[i].basis) nq = ipow(op->in[i].basis->Q, op->in[i].basis->dim);
  for (int i = 0; i < op->num_out && !nq; i++)
    if (op->out[i].basis) nq = ipow(op->out[i].basis->Q, op->out[i].basis->dim);
  for (int i = 0; i < op->num_in && !nq; i++)
    if (op->in[i].has_rstr) nq = op->in[i].elem_size;
  if (!nq) return 1;
  if (!add) memset(v, 0, sizeof(double) * v_len);
  double *e_in[16] = {0}, *q_in[16] = {0}, *e_out[16] = {0}, *q_out[16] = {0};
  for (int i = 0; i < op->num_in; i++) {
    const OracleField *f = &op->in[i];
    if (f->has_rstr) e_in[i] = (double *)malloc(sizeof(double) * f->elem_size * f->num_comp);
    q_in[i] = (double *)malloc(sizeof(double) * nq * f->size);
  }
  for (int i = 0; i < op->num_out; i++) {
    const OracleField *f = &op->out[i];
    e_out[i]             = (double *)malloc(sizeof(double) * f->elem_size * f->num_comp);
    q_out[i]             = (double *)malloc(sizeof(double) * nq * f->size);
  }
  for (int e = 0; e < op->num_elem; e++) {
    const double *qf_in[16];
    double       *qf_out[16];
    for (int i = 0; i < op->num_in; i++) {
      const OracleField *f = &op->in[i];
      if (f->eval_mode == ORACLE_EVAL_WEIGHT) {
        oracle_basis_apply(f->basis, 0, ORACLE_EVAL_WEIGHT, NULL, q_in[i]);
      } else {
        field_restrict(f, e, 0, f->is_active ? u : f->vec, e_in[i], NULL);
        if (f->eval_mode == ORACLE_EVAL_NONE) memcpy(q_in[i], e_in[i], sizeof(double) * nq * f->size);
        else oracle_basis_apply(f->basis, 0, f->eval_mode, e_in[i], q_in[i]);
      }
      qf_in[i] = q_in[i];
    }
    for (int i = 0; i < op->num_out; i++) qf_out[i] = q_out[i];
    op->qf(op->ctx, nq, qf_in, qf_out);
    for (int i = 0; i < op->num_out; i++) {
      const OracleField *f = &op->out[i];
      if (f->eval_mode == ORACLE_EVAL_NONE) {
        memcpy(e_out[i], q_out[i], sizeof(double) * nq * f->size);
      } else {
        memset(e_out[i], 0, sizeof(double) * f->elem_size * f->num_comp);
        oracle_basis_apply(f->basis, 1, f->eval_mode, q_out[i], e_out[i]);
      }
      field_restrict(f, e, 1, NULL, e_out[i], f->is_active ? v : f->vec);
    }
  }
  for (int i = 0; i < 16; i++) {
    free(e_in[i]);
    free(q_in[i]);
    free(e_out[i]);
    free(q_out[i]);
  }
  return 0;
}

/* ======================================================================================= BP convenience layer */
int oracle_bp_sizes(int bp, int qfset, int p, int *P, int *Q, int *ncomp, int *ncq) {
  if (bp < 1 || bp > 6) return 1;
  const int is_diff = bp >= 3, q_extra = bp >= 5 ? 1 : 2;
  *P     = p + 1;
  *Q     = p + q_extra;
  *ncomp = (bp % 2 == 0) ? 3 : 1;
  *ncq   = is_diff ? (qfset == 1 ? 6 : 7) : 1;
  return 0;
}

typedef struct {
  int         P, Q;
  double     *interp, *grad, *q_ref, *q_weight, *collo;
  OracleBasis bx, bu;
} BPBases;

static int bp_bases(int bp, int qfset, int p, BPBases *b, int *ncomp, int *ncq) {
  if (oracle_bp_sizes(bp, qfset, p, &b->P, &b->Q, ncomp, ncq)) return 1;
  const int P = b->P, Q = b->Q;
  b->interp   = (double *)malloc(sizeof(double) * P * Q);
  b->grad     = (double *)malloc(sizeof(double) * P * Q);
  b->q_ref    = (double *)malloc(sizeof(double) * Q);
  b->q_weight = (double *)malloc(sizeof(double) * Q);
  b->collo    = (double *)malloc(sizeof(double) * Q * Q);
  if (oracle_lagrange_1d(P, Q, bp >= 5 ? ORACLE_GAUSS_LOBATTO : ORACLE_GAUSS, b->interp, b->grad, b->q_ref, b->q_weight)) return 1;
  int collocated = P == Q;
  for (int i = 0; i < P && collocated; i++)
    for (int j = 0; j < Q; j++)
      if (!(fabs(b->interp[j + P * i] - (i == j ? 1.0 : 0.0)) < 10 * 1e-16)) collocated = 0; /* interface/ceed-basis.c:840-854 */
  const double *collo = NULL;
  if (!collocated && Q >= P) { /* backends/ref/ceed-ref-basis.c:295-299 */
    oracle_collocated_grad(P, Q, b->interp, b->grad, b->collo);
    collo = b->collo;
  }
  OracleBasis base = {3, 3, P, Q, b->interp, b->grad, b->q_weight, collo, collocated};
  b->bx            = base;
  b->bu            = base;
  b->bu.num_comp   = *ncomp;
  return 0;
}
static void bp_bases_free(BPBases *b) {
  free(b->interp);
  free(b->grad);
  free(b->q_ref);
  free(b->q_weight);
  free(b->collo);
}

int oracle_bp_qdata(int bp, int qfset, int p, int num_elem, int64_t num_nodes, const int32_t *offsets, const double *coords, double *qdata) {
  BPBases b;
  int     ncomp, ncq;
  if (bp_bases(bp, qfset, p, &b, &ncomp, &ncq)) return 1;
  const int      P3 = b.P * b.P * b.P, Q3 = b.Q * b.Q * b.Q;
  OracleOperator op;
  memset(&op, 0, sizeof(op));
  op.num_elem = num_elem;
  OracleField x = {ORACLE_EVAL_INTERP, 3, 1, 0, P3, 3, num_nodes, {0, 0, 0}, offsets, &b.bx, 1, NULL};
  OracleField dx = x, w, qd;
  dx.eval_mode   = ORACLE_EVAL_GRAD;
  dx.size        = 9;
  memset(&w, 0, sizeof(w));
  w.eval_mode = ORACLE_EVAL_WEIGHT;
  w.size      = 1;
  w.basis     = &b.bx;
  memset(&qd, 0, sizeof(qd));
  qd.eval_mode  = ORACLE_EVAL_NONE;
  qd.size       = ncq;
  qd.has_rstr   = 1;
  qd.is_strided = 1;
  qd.elem_size  = Q3;
  qd.num_comp   = ncq;
  /* CEED_STRIDES_BACKEND of the GPU family: {1, elem_size * num_elem, elem_size} (so fixtures are layout-compatible) */
  qd.strides[0] = 1;
  qd.strides[1] = (int64_t)Q3 * num_elem;
  qd.strides[2] = Q3;
  qd.is_active  = 1;
  if (qfset == 0) {
    op.num_in = 3;
    op.in[0]  = x;
    op.in[1]  = dx;
    op.in[2]  = w;
    op.qf     = bp >= 3 ? oracle_qf_bp_setup_diff : oracle_qf_bp_setup_mass;
  } else {
    op.num_in = 2;
    op.in[0]  = dx;
    op.in[1]  = w;
    op.qf     = bp >= 3 ? oracle_qf_poisson3d_build : oracle_qf_mass3d_build;
  }
  op.num_out = 1;
  op.out[0]  = qd;
  int ierr   = oracle_operator_apply(&op, coords, qdata, (int64_t)num_elem * Q3 * ncq, 0);
  bp_bases_free(&b);
  return ierr;
}

int oracle_bp_apply(int bp, int qfset, int p, int num_elem, int64_t num_nodes, const int32_t *offsets, int interlaced, const double *qdata,
                    const double *u, double *v, int add) {
  BPBases b;
  int     ncomp, ncq;
  if (bp_bases(bp, qfset, p, &b, &ncomp, &ncq)) return 1;
  const int      P3 = b.P * b.P * b.P, Q3 = b.Q * b.Q * b.Q, is_diff = bp >= 3;
  int32_t       *off2 = NULL;
  OracleOperator op;
  memset(&op, 0, sizeof(op));
  op.num_elem = num_elem;
  OracleField fu = {is_diff ? ORACLE_EVAL_GRAD : ORACLE_EVAL_INTERP, ncomp * (is_diff ? 3 : 1), 1, 0, P3, ncomp, num_nodes, {0, 0, 0}, offsets, &b.bu, 1,
                    NULL};
  if (interlaced && ncomp > 1) {
    off2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)num_elem * P3);
    for (int64_t i = 0; i < (int64_t)num_elem * P3; i++) off2[i] = offsets[i] * ncomp;
    fu.offsets     = off2;
    fu.comp_stride = 1;
  }
  OracleField fq;
  memset(&fq, 0, sizeof(fq));
  fq.eval_mode  = ORACLE_EVAL_NONE;
  fq.size       = ncq;
  fq.has_rstr   = 1;
  fq.is_strided = 1;
  fq.elem_size  = Q3;
  fq.num_comp   = ncq;
  fq.strides[0] = 1;
  fq.strides[1] = (int64_t)Q3 * num_elem;
  fq.strides[2] = Q3;
  fq.vec        = (double *)qdata;
  op.num_in     = 2;
  if (qfset == 1 && !is_diff) { /* gallery MassApply: in[0] = u, in[1] = qdata (gallery/mass/ceed-massapply.c:23-25) */
  }
  op.in[0]   = fu;
  op.in[1]   = fq;
  op.num_out = 1;
  op.out[0]  = fu;
  if (qfset == 0) op.qf = is_diff ? (ncomp == 3 ? oracle_qf_bp_diff3 : oracle_qf_bp_diff) : (ncomp == 3 ? oracle_qf_bp_mass3 : oracle_qf_bp_mass);
  else op.qf = is_diff ? oracle_qf_poisson3d_apply : oracle_qf_mass_apply;
  int ierr = (qfset == 1 && ncomp != 1) ? 1 : oracle_operator_apply(&op, u, v, (int64_t)ncomp * num_nodes, add);
  free(off2);
  bp_bases_free(&b);
  return ierr;
}
