// b200_basis.cu -- tensor-product H1 CeedBasis for the b200 backend.
//
//  * host construction of 1-D Lagrange matrices / quadrature / collocated gradient (what the reference interface layer
//    computes in interface/ceed-basis.c:1617-1680, 2529-2650, 750-773; restated here so the C ABI is usable standalone)
//  * device copies of interp_1d / grad_1d / q_weight_1d / collo_grad_1d (as backends/cuda-shared/ceed-cuda-shared-basis.c:603-666)
//  * standalone CeedBasisApply[Add] (E-vector [comp][elem][node] <-> Q-vector [dim][comp][elem][qpt]) as a sequence of
//    1-D contractions, using the same formulation as the CPU reference (backends/ref/ceed-ref-basis.c:65-175):
//    interpolate, then collocated gradient when Q >= P.  The fused operator kernel (b200_opgen.cpp) does NOT go through
//    this path; this is the unfused twin used by CeedBasisApply callers (t3xx tests, multigrid, fallbacks).
#include <cmath>
#include <cstring>

#include "b200_internal.h"

// =================================================================================================== host math
extern "C" int ceedb200_host_gauss_quadrature(b200_int Q, b200_scalar *x, b200_scalar *w) {
  // Gauss-Legendre nodes by Newton iteration on P_Q (interface/ceed-basis.c:2529-2571): Chebyshev initial guess,
  // three-term recurrence for P_n, derivative from (x P_n - P_{n-1}) n / (x^2 - 1).
  const double pi = 4.0 * atan(1.0);
  for (int i = 0; i <= Q / 2; i++) {
    double xi = cos(pi * (double)(2 * i + 1) / ((double)(2 * Q)));
    double p0 = 1.0, p1 = xi, p2 = 0.0;
    for (int j = 2; j <= Q; j++) {
      p2 = (((double)(2 * j - 1)) * xi * p1 - ((double)(j - 1)) * p0) / ((double)j);
      p0 = p1;
      p1 = p2;
    }
    double dp2 = (xi * p2 - p0) * (double)Q / (xi * xi - 1.0);
    xi         = xi - p2 / dp2;
    for (int k = 0; k < 100 && fabs(p2) > 1e-15; k++) {
      p0 = 1.0;
      p1 = xi;
      for (int j = 2; j <= Q; j++) {
        p2 = (((double)(2 * j - 1)) * xi * p1 - ((double)(j - 1)) * p0) / ((double)j);
        p0 = p1;
        p1 = p2;
      }
      dp2 = (xi * p2 - p0) * (double)Q / (xi * xi - 1.0);
      xi  = xi - p2 / dp2;
    }
    const double wi = 2.0 / ((1.0 - xi * xi) * dp2 * dp2);
    if (w) w[i] = w[Q - 1 - i] = wi;
    x[i]         = -xi;
    x[Q - 1 - i] = xi;
  }
  return B200_SUCCESS;
}

extern "C" int ceedb200_host_lobatto_quadrature(b200_int Q, b200_scalar *x, b200_scalar *w) {
  // Gauss-Legendre-Lobatto: endpoints +-1, interior nodes are roots of P'_{Q-1} (interface/ceed-basis.c:2581-2643)
  if (Q < 2) return B200_ERROR_DIMENSION;
  const double pi = 4.0 * atan(1.0);
  double       wi = 2.0 / ((double)(Q * (Q - 1)));
  if (w) w[0] = w[Q - 1] = wi;
  x[0]     = -1.0;
  x[Q - 1] = 1.0;
  for (int i = 1; i <= (Q - 1) / 2; i++) {
    double xi = cos(pi * (double)i / (double)(Q - 1));
    double p0 = 1.0, p1 = xi, p2 = 0.0;
    for (int j = 2; j < Q; j++) {
      p2 = (((double)(2 * j - 1)) * xi * p1 - ((double)(j - 1)) * p0) / ((double)j);
      p0 = p1;
      p1 = p2;
    }
    double dp2  = (xi * p2 - p0) * (double)Q / (xi * xi - 1.0);
    double d2p2 = (2 * xi * dp2 - (double)(Q * (Q - 1)) * p2) / (1.0 - xi * xi);
    xi          = xi - dp2 / d2p2;
    for (int k = 0; k < 100 && fabs(dp2) > 1e-15; k++) {
      p0 = 1.0;
      p1 = xi;
      for (int j = 2; j < Q; j++) {
        p2 = (((double)(2 * j - 1)) * xi * p1 - ((double)(j - 1)) * p0) / ((double)j);
        p0 = p1;
        p1 = p2;
      }
      dp2  = (xi * p2 - p0) * (double)Q / (xi * xi - 1.0);
      d2p2 = (2 * xi * dp2 - (double)(Q * (Q - 1)) * p2) / (1.0 - xi * xi);
      xi   = xi - dp2 / d2p2;
    }
    wi = 2.0 / (((double)(Q * (Q - 1))) * p2 * p2);
    if (w) w[i] = w[Q - 1 - i] = wi;
    x[i]         = -xi;
    x[Q - 1 - i] = xi;
  }
  return B200_SUCCESS;
}

extern "C" int ceedb200_host_lagrange_1d(b200_int P, b200_int Q, int quad_mode, b200_scalar *interp, b200_scalar *grad, b200_scalar *q_ref,
                                         b200_scalar *q_weight) {
  // Lagrange basis on P GLL nodes evaluated (value + derivative) at the Q quadrature points with Fornberg's recurrence
  // (interface/ceed-basis.c:1646-1669).
  std::vector<double> nodes(P);
  if (P == 1) nodes[0] = 0.0;
  else if (ceedb200_host_lobatto_quadrature(P, nodes.data(), nullptr)) return B200_ERROR_DIMENSION;
  int ierr = quad_mode == B200_GAUSS ? ceedb200_host_gauss_quadrature(Q, q_ref, q_weight) : ceedb200_host_lobatto_quadrature(Q, q_ref, q_weight);
  if (ierr) return ierr;
  for (int i = 0; i < P * Q; i++) interp[i] = grad[i] = 0.0;
  for (int i = 0; i < Q; i++) {
    double c1 = 1.0, c3 = nodes[0] - q_ref[i];
    interp[i * P + 0] = 1.0;
    for (int j = 1; j < P; j++) {
      double c2 = 1.0;
      const double c4 = c3;
      c3              = nodes[j] - q_ref[i];
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
  return B200_SUCCESS;
}

namespace {
// A (rows m, row stride `row`, col stride `col`) <- (I - b v v^T) A with v[0] = 1 implied
void householder_reflect(double *A, const double *v, double b, int m, int n, int row, int col) {
  for (int j = 0; j < n; j++) {
    double w = A[0 * row + j * col];
    for (int i = 1; i < m; i++) w += v[i] * A[i * row + j * col];
    A[0 * row + j * col] -= b * w;
    for (int i = 1; i < m; i++) A[i * row + j * col] -= b * w * v[i];
  }
}
}  // namespace

extern "C" int ceedb200_host_collocated_grad_1d(b200_int P, b200_int Q, const b200_scalar *interp, const b200_scalar *grad, b200_scalar *collo) {
  // collo_grad = grad_1d * pinv(interp_1d), pinv from a Householder QR of the Q x P interp matrix
  // (interface/ceed-basis.c:750-773 with CeedMatrixPseudoinverse :1273-1306, CeedQRFactorization :1188-1226).
  if (Q < P) return B200_ERROR_DIMENSION;
  const int           m = Q, n = P;
  std::vector<double> qr(interp, interp + m * n), tau(m, 0.0), v(m), eye(m * m, 0.0), pinv(n * m, 0.0);
  for (int i = 0; i < n; i++) {
    if (i >= m - 1) {
      tau[i] = 0.0;
      break;
    }
    double sigma = 0.0;
    v[i]         = qr[i + n * i];
    for (int j = i + 1; j < m; j++) {
      v[j] = qr[i + n * j];
      sigma += v[j] * v[j];
    }
    const double norm = sqrt(v[i] * v[i] + sigma);
    const double r_ii = -copysign(norm, v[i]);
    v[i] -= r_ii;
    tau[i] = 2 * v[i] * v[i] / (v[i] * v[i] + sigma);
    for (int j = i + 1; j < m; j++) v[j] /= v[i];
    householder_reflect(&qr[i * n + i + 1], &v[i], tau[i], m - i, n - i - 1, n, 1);
    qr[i + n * i] = r_ii;
    for (int j = i + 1; j < m; j++) qr[i + n * j] = v[j];
  }
  // eye <- Q^T
  for (int i = 0; i < m; i++) eye[i * m + i] = 1.0;
  for (int i = 0; i < n; i++) {
    for (int j = i + 1; j < m; j++) v[j] = qr[j * n + i];
    householder_reflect(&eye[i * m], &v[i], tau[i], m - i, m, m, 1);
  }
  // pinv = R^{-1} Q^T (back substitution, column by column)
  for (int j = 0; j < m; j++) {
    pinv[j + m * (n - 1)] = eye[j + m * (n - 1)] / qr[n * n - 1];
    for (int i = n - 2; i >= 0; i--) {
      pinv[j + m * i] = eye[j + m * i];
      for (int k = i + 1; k < n; k++) pinv[j + m * i] -= qr[k + n * i] * pinv[j + m * k];
      pinv[j + m * i] /= qr[i + n * i];
    }
  }
  for (int i = 0; i < Q; i++) {
    for (int j = 0; j < Q; j++) {
      double sum = 0;
      for (int k = 0; k < P; k++) sum += grad[k + i * P] * pinv[j + k * Q];
      collo[j + i * Q] = sum;
    }
  }
  return B200_SUCCESS;
}

// =================================================================================================== device side
namespace {
constexpr int kThreads = 256;

// v[(a*J + j)*C + c] (+)= sum_b t[j*ts0 + b*ts1] * u[(a*B + b)*C + c]   (same index convention as
// CeedTensorContractApply, backends/ref/ceed-ref-tensor.c:16-38)
__global__ void k_contract(int64_t A, int B, int64_t C, int J, const double *__restrict__ t, int ts0, int ts1, int add, const double *__restrict__ u,
                           double *__restrict__ v) {
  const int64_t total  = A * J * C;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t c  = i % C;
    const int64_t aj = i / C;
    const int     j  = (int)(aj % J);
    const int64_t a  = aj / J;
    double        s  = add ? v[i] : 0.0;
    for (int b = 0; b < B; b++) s += t[j * ts0 + b * ts1] * u[(a * B + b) * C + c];
    v[i] = s;
  }
}

__global__ void k_weight(int dim, int Q, int64_t num_elem, const double *__restrict__ w, double *__restrict__ v) {
  int64_t nq = 1;
  for (int d = 0; d < dim; d++) nq *= Q;
  const int64_t total  = nq * num_elem;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int64_t q   = i % nq;
    double  val = 1.0;
    for (int d = 0; d < dim; d++) {
      val *= w[q % Q];
      q /= Q;
    }
    v[i] = val;
  }
}

inline unsigned grid_for(B200Ceed ceed, int64_t n) {
  int64_t blocks = (n + kThreads - 1) / kThreads;
  int64_t cap    = (int64_t)ceed->num_sms * 32;
  return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

int contract(B200Ceed ceed, int64_t A, int B, int64_t C, int J, const double *t, bool transpose, bool add, const double *u, double *v) {
  B200_CHECK(!b200_compile_only(), ceed, B200_ERROR_BACKEND, "CEED_B200_COMPILE_ONLY is set; kernels cannot run");
  // t is [Jn x Bn] row-major when not transposed (t[j*B + b]); transposed reads t[b*J + j]
  const int ts0 = transpose ? 1 : B, ts1 = transpose ? J : 1;
  k_contract<<<grid_for(ceed, A * J * C), kThreads, 0, ceed->stream>>>(A, B, C, J, t, ts0, ts1, add ? 1 : 0, u, v);
  ceed->launch_count++;
  B200_CUDA(ceed, cudaGetLastError());
  return B200_SUCCESS;
}

int64_t ipow(int64_t b, int e) {
  int64_t r = 1;
  for (int i = 0; i < e; i++) r *= b;
  return r;
}

// Apply `mat` (Qn x Pn, or its transpose) along every one of `dim` tensor directions of u -> v, for `nvec` stacked tensors.
// Direction 0 is the fastest index.  tmp0/tmp1 are scratch of sufficient size.  add applies to the final write.
int apply_all_dims(B200Ceed ceed, int dim, int64_t nvec, int Pn, int Qn, const double *mat, bool transpose, bool add, const double *u, double *v,
                   double *tmp0, double *tmp1) {
  // per-tensor sizes: before contracting direction d, tensor is [Pn^(dim-d)] slow x [Qn^d] fast
  const double *src = u;
  for (int d = 0; d < dim; d++) {
    const int64_t pre  = nvec * ipow(Pn, dim - 1 - d);
    const int64_t post = ipow(Qn, d);
    double       *dst  = d == dim - 1 ? v : (d % 2 == 0 ? tmp0 : tmp1);
    B200_CALL(contract(ceed, pre, Pn, post, Qn, mat, transpose, add && d == dim - 1, src, dst));
    src = dst;
  }
  return B200_SUCCESS;
}

// Apply `mat` along direction d only (others untouched; tensor extents are all n = rows = cols of mat).
int apply_one_dim(B200Ceed ceed, int dim, int64_t nvec, int n, int d, const double *mat, bool transpose, bool add, const double *u, double *v) {
  const int64_t pre = nvec * ipow(n, dim - 1 - d), post = ipow(n, d);
  return contract(ceed, pre, n, post, n, mat, transpose, add, u, v);
}

// grow-only work buffers cached on the context: consecutive applies reuse them in stream order (no cudaMalloc / cudaFree /
// synchronisation per apply, so unfused operators can be captured in CUDA graphs like the fused path)
int get_scratch(B200Ceed ceed, int which, double **buf, size_t bytes) {
  if (ceed->basis_tmp_bytes[which] < bytes) {
    if (ceed->d_basis_tmp[which]) {
      if (!b200_compile_only()) B200_CUDA(ceed, cudaStreamSynchronize(ceed->stream));  // kernels in flight may still use the old buffer
      B200_CALL(b200_dfree(ceed, ceed->d_basis_tmp[which]));
      ceed->d_basis_tmp[which] = nullptr, ceed->basis_tmp_bytes[which] = 0;
    }
    B200_CALL(b200_dmalloc(ceed, (void **)&ceed->d_basis_tmp[which], bytes));
    ceed->basis_tmp_bytes[which] = bytes;
  }
  *buf = ceed->d_basis_tmp[which];
  return B200_SUCCESS;
}

int basis_apply_ptr(B200Basis basis, bool apply_add, int num_elem, int t_mode, int eval_mode, const double *d_u, double *d_v);

int basis_apply_core(B200Basis basis, bool apply_add, int num_elem, int t_mode, int eval_mode, B200Vector U, B200Vector V) {
  B200Ceed      ceed = basis->ceed;
  const double *d_u  = nullptr;
  double       *d_v  = nullptr;
  if (eval_mode != B200_EVAL_WEIGHT) {
    B200_CHECK(U && U != B200_VECTOR_NONE, ceed, B200_ERROR_BACKEND, "An input vector is required for this CeedEvalMode");
    B200_CALL(b200_vector_device_read(U, &d_u));
  }
  B200_CALL(b200_vector_device_write(V, &d_v, false));
  return basis_apply_ptr(basis, apply_add, num_elem, t_mode, eval_mode, d_u, d_v);
}

int basis_apply_ptr(B200Basis basis, bool apply_add, int num_elem, int t_mode, int eval_mode, const double *d_u, double *d_v) {
  B200Ceed      ceed = basis->ceed;
  const int     dim = basis->dim, nc = basis->num_comp, P = basis->P, Q = basis->Q;
  const int64_t nvec   = (int64_t)nc * num_elem;
  const int64_t n_node = ipow(P, dim), n_qpt = ipow(Q, dim);
  if (eval_mode == B200_EVAL_WEIGHT)
    B200_CHECK(t_mode == B200_NOTRANSPOSE, ceed, B200_ERROR_BACKEND, "CEED_EVAL_WEIGHT incompatible with CEED_TRANSPOSE");
  else B200_CHECK(d_u, ceed, B200_ERROR_BACKEND, "An input vector is required for this CeedEvalMode");
  if (num_elem == 0) return B200_SUCCESS;
  if (!basis->is_tensor) {
    // CeedBasisCreateH1: dense element matrices, one contraction per evaluation (backends/ref/ceed-ref-basis.c:176-230)
    const int64_t nq = (int64_t)nvec * Q;
    switch (eval_mode) {
      case B200_EVAL_INTERP: return contract(ceed, nvec, t_mode == B200_NOTRANSPOSE ? P : Q, 1, t_mode == B200_NOTRANSPOSE ? Q : P, basis->d_interp,
                                             t_mode == B200_TRANSPOSE, apply_add, d_u, d_v);
      case B200_EVAL_GRAD:
        for (int d = 0; d < dim; d++) {
          const double *g = basis->d_grad + (int64_t)d * P * Q;
          if (t_mode == B200_NOTRANSPOSE) B200_CALL(contract(ceed, nvec, P, 1, Q, g, false, apply_add, d_u, d_v + d * nq));
          else B200_CALL(contract(ceed, nvec, Q, 1, P, g, true, apply_add || d > 0, d_u + d * nq, d_v));
        }
        return B200_SUCCESS;
      case B200_EVAL_WEIGHT:
        B200_CHECK(!b200_compile_only(), ceed, B200_ERROR_BACKEND, "CEED_B200_COMPILE_ONLY is set; kernels cannot run");
        k_weight<<<grid_for(ceed, (int64_t)Q * num_elem), kThreads, 0, ceed->stream>>>(1, Q, num_elem, basis->d_q_weight, d_v);
        ceed->launch_count++;
        B200_CUDA(ceed, cudaGetLastError());
        return B200_SUCCESS;
      default: return b200_error(ceed, B200_ERROR_UNSUPPORTED, "Backend does not implement eval mode %d for non-tensor H1 bases", eval_mode);
    }
  }
  const int     big      = P > Q ? P : Q;
  const size_t  tmp_size = (size_t)nvec * ipow(big, dim) * sizeof(double);
  double       *tmp0 = nullptr, *tmp1 = nullptr, *tmp2 = nullptr;
  int           ierr = B200_SUCCESS;
  B200_CALL(get_scratch(ceed, 0, &tmp0, tmp_size));
  B200_CALL(get_scratch(ceed, 1, &tmp1, tmp_size));
  B200_CALL(get_scratch(ceed, 2, &tmp2, tmp_size));
  switch (eval_mode) {
    case B200_EVAL_INTERP:
      if (t_mode == B200_NOTRANSPOSE) ierr = apply_all_dims(ceed, dim, nvec, P, Q, basis->d_interp, false, apply_add, d_u, d_v, tmp0, tmp1);
      else ierr = apply_all_dims(ceed, dim, nvec, Q, P, basis->d_interp, true, apply_add, d_u, d_v, tmp0, tmp1);
      break;
    case B200_EVAL_GRAD:
      if (basis->has_collo_grad) {
        if (t_mode == B200_NOTRANSPOSE) {
          // interp to quadrature points, then collocated derivative per direction
          ierr = apply_all_dims(ceed, dim, nvec, P, Q, basis->d_interp, false, false, d_u, tmp2, tmp0, tmp1);
          for (int d = 0; d < dim && !ierr; d++)
            ierr = apply_one_dim(ceed, dim, nvec, Q, d, basis->d_collo_grad, false, apply_add, tmp2, d_v + (int64_t)d * nvec * n_qpt);
        } else {
          for (int d = 0; d < dim && !ierr; d++)
            ierr = apply_one_dim(ceed, dim, nvec, Q, d, basis->d_collo_grad, true, d > 0, d_u + (int64_t)d * nvec * n_qpt, tmp2);
          if (!ierr) ierr = apply_all_dims(ceed, dim, nvec, Q, P, basis->d_interp, true, apply_add, tmp2, d_v, tmp0, tmp1);
        }
      } else {
        // P > Q: dim^2 contractions, grad_1d in direction p, interp_1d elsewhere (ceed-ref-basis.c:150-174)
        for (int p = 0; p < dim && !ierr; p++) {
          const bool    tr   = t_mode == B200_TRANSPOSE;
          const int     Pn   = tr ? Q : P, Qn = tr ? P : Q;
          const double *src  = tr ? d_u + (int64_t)p * nvec * n_qpt : d_u;
          double       *dst  = tr ? d_v : d_v + (int64_t)p * nvec * n_qpt;
          const double *cur  = src;
          for (int d = 0; d < dim && !ierr; d++) {
            const int64_t pre = nvec * ipow(Pn, dim - 1 - d), post = ipow(Qn, d);
            double       *out = d == dim - 1 ? dst : (d % 2 == 0 ? tmp0 : tmp1);
            const bool    add = d == dim - 1 && (apply_add || (tr && p > 0));
            ierr = contract(ceed, pre, Pn, post, Qn, p == d ? basis->d_grad : basis->d_interp, tr, add, cur, out);
            cur  = out;
          }
        }
      }
      break;
    case B200_EVAL_WEIGHT:
      B200_CHECK(!b200_compile_only(), ceed, B200_ERROR_BACKEND, "CEED_B200_COMPILE_ONLY is set; kernels cannot run");
      k_weight<<<grid_for(ceed, n_qpt * num_elem), kThreads, 0, ceed->stream>>>(dim, Q, num_elem, basis->d_q_weight, d_v);
      ceed->launch_count++;
      break;
    default:
      ierr = b200_error(ceed, B200_ERROR_UNSUPPORTED, "Backend does not implement eval mode %d for tensor H1 bases", eval_mode);
  }
  (void)n_node;
  return ierr;
}
}  // namespace

extern "C" int ceedb200_basis_create_tensor_h1(B200Ceed ceed, b200_int dim, b200_int num_comp, b200_int P, b200_int Q, const b200_scalar *interp,
                                               const b200_scalar *grad, const b200_scalar *q_ref, const b200_scalar *q_weight,
                                               B200Basis *basis_out) {
  B200_CHECK(dim >= 1 && dim <= 3 && P >= 1 && Q >= 1 && num_comp >= 1, ceed, B200_ERROR_DIMENSION, "invalid basis dimensions");
  B200Basis b = new B200Basis_();
  b->ceed     = ceed;
  b->dim      = dim;
  b->num_comp = num_comp;
  b->P        = P;
  b->Q        = Q;
  b->interp.assign(interp, interp + P * Q);
  b->grad.assign(grad, grad + P * Q);
  if (q_ref) b->q_ref.assign(q_ref, q_ref + Q);
  else b->q_ref.assign(Q, 0.0);
  if (q_weight) b->q_weight.assign(q_weight, q_weight + Q);  // projection bases come without quadrature (interface/ceed-basis.c:1905)
  else b->q_weight.assign(Q, 0.0);
  // collocated: interp_1d is the identity (interface/ceed-basis.c:840-854 uses a 1e-14-ish tolerance on |B - I|)
  b->is_collocated = P == Q;
  for (int i = 0; i < Q && b->is_collocated; i++)
    for (int j = 0; j < P; j++)
      if (!(fabs(interp[i * P + j] - (i == j ? 1.0 : 0.0)) < 10 * 1e-16)) b->is_collocated = false;
  if (b->is_collocated) {
    // nodes == quadrature points: the derivative matrix is grad_1d itself (backends/ref/ceed-ref-basis.c:133-147, :295-299)
    b->collo_grad     = b->grad;
    b->has_collo_grad = true;
  } else if (Q >= P) {
    b->collo_grad.resize(Q * Q);
    if (ceedb200_host_collocated_grad_1d(P, Q, interp, grad, b->collo_grad.data()) == B200_SUCCESS) b->has_collo_grad = true;
  }
  const size_t mb = (size_t)P * Q * sizeof(double);
  B200_CALL(b200_dmalloc(ceed, (void **)&b->d_interp, mb));
  B200_CALL(b200_dmalloc(ceed, (void **)&b->d_grad, mb));
  B200_CALL(b200_dmalloc(ceed, (void **)&b->d_q_weight, Q * sizeof(double)));
  B200_CALL(b200_h2d(ceed, b->d_interp, b->interp.data(), mb));
  B200_CALL(b200_h2d(ceed, b->d_grad, b->grad.data(), mb));
  B200_CALL(b200_h2d(ceed, b->d_q_weight, b->q_weight.data(), Q * sizeof(double)));
  if (b->has_collo_grad) {
    B200_CALL(b200_dmalloc(ceed, (void **)&b->d_collo_grad, (size_t)Q * Q * sizeof(double)));
    B200_CALL(b200_h2d(ceed, b->d_collo_grad, b->collo_grad.data(), (size_t)Q * Q * sizeof(double)));
  }
  *basis_out = b;
  return B200_SUCCESS;
}

// CeedBasisCreateH1 (interface/ceed-basis.c:1434; backends/cuda-ref/ceed-cuda-ref-basis.c:340-400): non-tensor H1 basis with
// num_nodes nodes and num_qpts quadrature points per element, interp [num_qpts x num_nodes], grad [dim][num_qpts x num_nodes].
// Used by mixed-topology / composite operators (triangles, tets, ...) through the unfused operator path.
extern "C" int ceedb200_basis_create_h1(B200Ceed ceed, b200_int dim, b200_int num_comp, b200_int num_nodes, b200_int num_qpts, const b200_scalar *interp,
                                        const b200_scalar *grad, const b200_scalar *q_ref, const b200_scalar *q_weight, B200Basis *basis_out) {
  B200_CHECK(dim >= 1 && dim <= 3 && num_nodes >= 1 && num_qpts >= 1 && num_comp >= 1, ceed, B200_ERROR_DIMENSION, "invalid basis dimensions");
  B200Basis b  = new B200Basis_();
  b->ceed      = ceed;
  b->dim       = dim;
  b->num_comp  = num_comp;
  b->P         = num_nodes;
  b->Q         = num_qpts;
  b->is_tensor = false;
  const size_t n = (size_t)num_nodes * num_qpts;
  b->interp.assign(interp, interp + n);
  b->grad.assign(grad, grad + n * dim);
  if (q_ref) b->q_ref.assign(q_ref, q_ref + (size_t)num_qpts * dim);
  if (q_weight) b->q_weight.assign(q_weight, q_weight + num_qpts);
  else b->q_weight.assign(num_qpts, 0.0);
  B200_CALL(b200_dmalloc(ceed, (void **)&b->d_interp, n * sizeof(double)));
  B200_CALL(b200_dmalloc(ceed, (void **)&b->d_grad, n * dim * sizeof(double)));
  B200_CALL(b200_dmalloc(ceed, (void **)&b->d_q_weight, num_qpts * sizeof(double)));
  B200_CALL(b200_h2d(ceed, b->d_interp, b->interp.data(), n * sizeof(double)));
  B200_CALL(b200_h2d(ceed, b->d_grad, b->grad.data(), n * dim * sizeof(double)));
  B200_CALL(b200_h2d(ceed, b->d_q_weight, b->q_weight.data(), num_qpts * sizeof(double)));
  *basis_out = b;
  return B200_SUCCESS;
}

extern "C" int ceedb200_basis_create_tensor_h1_lagrange(B200Ceed ceed, b200_int dim, b200_int num_comp, b200_int P, b200_int Q, int quad_mode,
                                                        B200Basis *basis) {
  B200_CHECK(P >= 1 && Q >= 1, ceed, B200_ERROR_DIMENSION, "invalid P/Q");
  std::vector<double> interp(P * Q), grad(P * Q), q_ref(Q), q_weight(Q);
  int                 ierr = ceedb200_host_lagrange_1d(P, Q, quad_mode, interp.data(), grad.data(), q_ref.data(), q_weight.data());
  B200_CHECK(!ierr, ceed, ierr, "Lagrange basis construction failed");
  return ceedb200_basis_create_tensor_h1(ceed, dim, num_comp, P, Q, interp.data(), grad.data(), q_ref.data(), q_weight.data(), basis);
}

extern "C" int ceedb200_basis_destroy(B200Basis b) {
  if (!b) return B200_SUCCESS;
  b200_dfree(b->ceed, b->d_interp);
  b200_dfree(b->ceed, b->d_grad);
  b200_dfree(b->ceed, b->d_q_weight);
  b200_dfree(b->ceed, b->d_collo_grad);
  delete b;
  return B200_SUCCESS;
}

extern "C" int ceedb200_basis_apply(B200Basis basis, b200_int num_elem, int t_mode, int eval_mode, B200Vector u, B200Vector v) {
  if (t_mode == B200_TRANSPOSE) {
    // transpose apply overwrites the E-vector: contributions of the dim Q-components are summed internally
    return basis_apply_core(basis, false, num_elem, t_mode, eval_mode, u, v);
  }
  return basis_apply_core(basis, false, num_elem, t_mode, eval_mode, u, v);
}
extern "C" int ceedb200_basis_apply_add(B200Basis basis, b200_int num_elem, int t_mode, int eval_mode, B200Vector u, B200Vector v) {
  return basis_apply_core(basis, true, num_elem, t_mode, eval_mode, u, v);
}

extern "C" int ceedb200_basis_apply_ptr(B200Basis basis, b200_int num_elem, int t_mode, int eval_mode, int add, const b200_scalar *d_u,
                                        b200_scalar *d_v) {
  return basis_apply_ptr(basis, add != 0, num_elem, t_mode, eval_mode, d_u, d_v);
}

extern "C" int ceedb200_basis_get_matrix(B200Basis b, int which, b200_scalar *out) {
  const std::vector<double> *src = nullptr;
  switch (which) {
    case 0: src = &b->interp; break;
    case 1: src = &b->grad; break;
    case 2: src = &b->q_ref; break;
    case 3: src = &b->q_weight; break;
    case 4:
      B200_CHECK(b->has_collo_grad, b->ceed, B200_ERROR_UNSUPPORTED, "no collocated gradient (Q < P)");
      src = &b->collo_grad;
      break;
    default: return b200_error(b->ceed, B200_ERROR_UNSUPPORTED, "unknown matrix id %d", which);
  }
  memcpy(out, src->data(), src->size() * sizeof(double));
  return B200_SUCCESS;
}
