/* ceed_b200.h -- C ABI of the B200-native libCEED operator-apply path ("/gpu/cuda/b200").
 *
 * This is the drop-in boundary: a plain C interface (opaque handles, raw pointers, sizes, int error
 * codes; no C++/torch types) that a libCEED backend binds into the reference's per-object function
 * pointer slots (CeedSetBackendFunction, /root/reference/include/ceed/backend.h:246-261,
 * slot table /root/reference/include/ceed-impl.h:101-399).  `libceed_b200/backend/` contains that binding
 * (INTEGRATION.md shows how it is registered); everything below is what it calls.
 *
 * Each entry point cites the reference interface it replaces (paths relative to /root/reference).
 * Enumerations use the reference's numeric values (include/ceed/types.h:186-252) so they can be passed through.
 *
 * Error convention (include/ceed/types.h:162-181): 0 = success, >0 recoverable, <0 fatal.  The message
 * of the last failure on a context is available from ceedb200_last_error().
 *
 * Threading: like the reference (README.md:74-80) a context and its objects are single-threaded;
 * use one context per GPU / thread.  All work is enqueued on one CUDA stream per context.
 */
#ifndef CEED_B200_H
#define CEED_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CEEDB200_EXPORT __attribute__((visibility("default")))

typedef double  b200_scalar; /* CeedScalar (include/ceed/ceed-f64.h:17) */
typedef int32_t b200_int;    /* CeedInt    (include/ceed/types.h:131)   */
typedef int64_t b200_size;   /* CeedSize   (include/ceed/types.h:136)   */

typedef struct B200Ceed_        *B200Ceed;
typedef struct B200Vector_      *B200Vector;
typedef struct B200Restriction_ *B200Restriction;
typedef struct B200Basis_       *B200Basis;
typedef struct B200QFContext_   *B200QFContext;
typedef struct B200QFunction_   *B200QFunction;
typedef struct B200Operator_    *B200Operator;

enum { B200_MEM_HOST = 0, B200_MEM_DEVICE = 1 };                               /* CeedMemType       */
enum { B200_COPY_VALUES = 0, B200_USE_POINTER = 1, B200_OWN_POINTER = 2 };      /* CeedCopyMode      */
enum { B200_NORM_1 = 0, B200_NORM_2 = 1, B200_NORM_MAX = 2 };                   /* CeedNormType      */
enum { B200_NOTRANSPOSE = 0, B200_TRANSPOSE = 1 };                              /* CeedTransposeMode */
enum { B200_EVAL_NONE = 0, B200_EVAL_INTERP = 1, B200_EVAL_GRAD = 2, B200_EVAL_WEIGHT = 16 }; /* CeedEvalMode */
enum { B200_GAUSS = 0, B200_GAUSS_LOBATTO = 1 };                                /* CeedQuadMode      */
enum {
  B200_SUCCESS = 0, B200_ERROR_MINOR = 1, B200_ERROR_DIMENSION = 2, B200_ERROR_INCOMPLETE = 3, B200_ERROR_INCOMPATIBLE = 4,
  B200_ERROR_ACCESS = 5, B200_ERROR_MAJOR = -1, B200_ERROR_BACKEND = -2, B200_ERROR_UNSUPPORTED = -3
};
/* How the fused operator kernel adds element contributions into the output L-vector. */
enum {
  B200_SCATTER_DETERMINISTIC = 0, /* owner-store + ordered halo sum; bitwise reproducible, no atomics (default) */
  B200_SCATTER_ATOMIC        = 1, /* red.global.add.f64, like /gpu/cuda/gen (cuda-gen-templates.h:385-396)   */
  B200_SCATTER_EVECTOR       = 2, /* E-vector to HBM + CSR transpose kernel, like /gpu/cuda/ref               */
  B200_SCATTER_ORDERED       = 3  /* deterministic, completed INSIDE the operator kernel: the element group holding the last
                                     E-entry of a shared node waits for the groups of the earlier entries (per-group flags)
                                     and adds their values in ascending E-order; same bits as DETERMINISTIC, no second pass */
};

/* ---------------------------------------------------------------- context (Ceed)
 * replaces CeedInit_Cuda / Ceed_Cuda: backends/cuda/ceed-cuda-common.c:19-35, ceed-cuda-common.h:67-72 */
CEEDB200_EXPORT int         ceedb200_init(int device_id, B200Ceed *ceed);
CEEDB200_EXPORT int         ceedb200_destroy(B200Ceed ceed);
CEEDB200_EXPORT const char *ceedb200_last_error(B200Ceed ceed);
CEEDB200_EXPORT const char *ceedb200_version(void);
/* CeedAddJitSourceRoot / CeedAddJitDefine: interface/ceed.c:1515,1579; consumed by NVRTC as -I / -D
 * (backends/cuda/ceed-cuda-compile.cpp:99-131) */
CEEDB200_EXPORT int ceedb200_add_jit_source_root(B200Ceed ceed, const char *path);
CEEDB200_EXPORT int ceedb200_add_jit_define(B200Ceed ceed, const char *define);
/* CeedSetStream (interface/ceed.c:1408): `stream` is a cudaStream_t; NULL = legacy default stream */
CEEDB200_EXPORT int ceedb200_set_stream(B200Ceed ceed, void *stream);
CEEDB200_EXPORT int ceedb200_get_stream(B200Ceed ceed, void **stream);
CEEDB200_EXPORT int ceedb200_synchronize(B200Ceed ceed);
CEEDB200_EXPORT int ceedb200_set_scatter_mode(B200Ceed ceed, int mode);
/* number of kernels launched by this context so far (bench.py's gpu_launches) */
CEEDB200_EXPORT int64_t ceedb200_launch_count(B200Ceed ceed);

/* ---------------------------------------------------------------- vector (CeedVector)
 * replaces CeedVector_Cuda: backends/cuda-ref/ceed-cuda-ref.h:15-22, ceed-cuda-ref-vector.c:21-130,177-228,307-485,
 * kernels backends/cuda-ref/kernels/cuda-ref-vector.cu:14-215 */
CEEDB200_EXPORT int ceedb200_vector_create(B200Ceed ceed, b200_size length, B200Vector *vec);
CEEDB200_EXPORT int ceedb200_vector_destroy(B200Vector vec);
CEEDB200_EXPORT int ceedb200_vector_length(B200Vector vec, b200_size *length);
CEEDB200_EXPORT int ceedb200_vector_has_valid_array(B200Vector vec, int *has_valid);
CEEDB200_EXPORT int ceedb200_vector_has_borrowed_array_of_type(B200Vector vec, int mem_type, int *has_borrowed);
/* which mirrors currently hold valid data (the lazy-sync state of backends/cuda-ref/ceed-cuda-ref-vector.c:20-38 CeedVectorNeedSync_Cuda) */
CEEDB200_EXPORT int ceedb200_vector_valid_sides(B200Vector vec, int *host_valid, int *device_valid);
CEEDB200_EXPORT int ceedb200_vector_set_array(B200Vector vec, int mem_type, int copy_mode, b200_scalar *array);
CEEDB200_EXPORT int ceedb200_vector_take_array(B200Vector vec, int mem_type, b200_scalar **array);
CEEDB200_EXPORT int ceedb200_vector_set_value(B200Vector vec, b200_scalar value);
CEEDB200_EXPORT int ceedb200_vector_set_value_strided(B200Vector vec, b200_size start, b200_size stop, b200_size step, b200_scalar value);
CEEDB200_EXPORT int ceedb200_vector_sync_array(B200Vector vec, int mem_type);
CEEDB200_EXPORT int ceedb200_vector_get_array(B200Vector vec, int mem_type, b200_scalar **array);             /* read-write */
CEEDB200_EXPORT int ceedb200_vector_get_array_read(B200Vector vec, int mem_type, const b200_scalar **array);  /* read-only  */
CEEDB200_EXPORT int ceedb200_vector_get_array_write(B200Vector vec, int mem_type, b200_scalar **array);       /* write-only */
CEEDB200_EXPORT int ceedb200_vector_copy_strided(B200Vector src, b200_size start, b200_size stop, b200_size step, B200Vector dst);
CEEDB200_EXPORT int ceedb200_vector_norm(B200Vector vec, int norm_type, b200_scalar *norm);
CEEDB200_EXPORT int ceedb200_vector_scale(B200Vector x, b200_scalar alpha);
CEEDB200_EXPORT int ceedb200_vector_reciprocal(B200Vector x);
CEEDB200_EXPORT int ceedb200_vector_filter(B200Vector x, b200_scalar epsilon);
CEEDB200_EXPORT int ceedb200_vector_axpy(B200Vector y, b200_scalar alpha, B200Vector x);                       /* y = alpha x + y        */
CEEDB200_EXPORT int ceedb200_vector_axpby(B200Vector y, b200_scalar alpha, b200_scalar beta, B200Vector x);    /* y = alpha x + beta y   */
CEEDB200_EXPORT int ceedb200_vector_pointwise_mult(B200Vector w, B200Vector x, B200Vector y);                  /* w = x .* y             */

/* ---------------------------------------------------------------- element restriction (CeedElemRestriction)
 * replaces CeedElemRestrictionCreate_Cuda / Apply: backends/cuda-ref/ceed-cuda-ref-restriction.c:24-291,416-661,
 * kernels include/ceed/jit-source/cuda/cuda-ref-restriction-offset.h:15-66, -strided.h:15-39.
 * E-vector layout is {1, num_elem*elem_size, elem_size} = [comp][elem][node] (cuda-ref-restriction.c:532-534). */
CEEDB200_EXPORT int ceedb200_restriction_create(B200Ceed ceed, b200_int num_elem, b200_int elem_size, b200_int num_comp, b200_int comp_stride,
                                                b200_size l_size, int mem_type, int copy_mode, const b200_int *offsets, B200Restriction *rstr);
/* strides = NULL selects CEED_STRIDES_BACKEND (interface/ceed-elemrestriction.c:630): {1, elem_size*num_elem, elem_size} */
CEEDB200_EXPORT int ceedb200_restriction_create_strided(B200Ceed ceed, b200_int num_elem, b200_int elem_size, b200_int num_comp, b200_size l_size,
                                                        const b200_int strides[3], B200Restriction *rstr);
CEEDB200_EXPORT int ceedb200_restriction_destroy(B200Restriction rstr);
CEEDB200_EXPORT int ceedb200_restriction_apply(B200Restriction rstr, int t_mode, B200Vector u, B200Vector v);
/* same, on raw device pointers (the libCEED plugin passes arrays obtained from CeedVectorGetArray*(CEED_MEM_DEVICE)) */
CEEDB200_EXPORT int ceedb200_restriction_apply_ptr(B200Restriction rstr, int t_mode, const b200_scalar *d_u, b200_scalar *d_v);
CEEDB200_EXPORT int ceedb200_restriction_get_offsets(B200Restriction rstr, int mem_type, const b200_int **offsets);
CEEDB200_EXPORT int ceedb200_restriction_get_e_layout(B200Restriction rstr, b200_int layout[3]);
/* elements [0, split_elem) touch the rank interface of a partitioned mesh and are applied first (ceedb200_operator_apply_part) */
CEEDB200_EXPORT int ceedb200_restriction_set_split(B200Restriction rstr, b200_int split_elem);
CEEDB200_EXPORT int ceedb200_restriction_get_info(B200Restriction rstr, b200_int *num_elem, b200_int *elem_size, b200_int *num_comp,
                                                  b200_size *l_size, b200_size *e_size);
/* diagnostics: host copy of the tables behind the fused kernel's deterministic scatter (no reference counterpart; the
 * reference scatters with atomics, cuda-gen-templates.h:385-396).  mode B200_SCATTER_DETERMINISTIC: tgt[e-entry] >= 0 L-index of
 * the owning (first) entry, < 0 ~halo slot.  mode B200_SCATTER_ORDERED (element groups of group_elems): tgt as documented at
 * B200_SCATTER_ORDERED plus the predecessor-group CSR.  counts = {shared nodes, halo slots, groups, predecessors, supported}. */
CEEDB200_EXPORT int ceedb200_restriction_debug_scatter_tables(B200Restriction rstr, int mode, int group_elems, int32_t *tgt, int64_t *counts,
                                                              int32_t *pred_ptr, int32_t *pred_idx, int64_t pred_capacity);

/* ---------------------------------------------------------------- basis (CeedBasis, tensor H1)
 * replaces CeedBasisCreateTensorH1_Cuda_shared / CeedBasisApply: backends/cuda-shared/ceed-cuda-shared-basis.c:24-196,603-666,
 * device code include/ceed/jit-source/cuda/cuda-shared-basis-tensor.h:18-514.
 * Matrices are row-major interp_1d[q*P+p], grad_1d[q*P+p] (include/ceed-impl.h:224-228).
 * Q-vector layout [dim][comp][elem][qpt]; E-vector layout [comp][elem][node]. */
CEEDB200_EXPORT int ceedb200_basis_create_tensor_h1(B200Ceed ceed, b200_int dim, b200_int num_comp, b200_int P_1d, b200_int Q_1d,
                                                    const b200_scalar *interp_1d, const b200_scalar *grad_1d, const b200_scalar *q_ref_1d,
                                                    const b200_scalar *q_weight_1d, B200Basis *basis);
/* host construction of Lagrange matrices: restates interface/ceed-basis.c:1617-1680 (Fornberg), :2529 (Gauss), :2581 (Lobatto) */
/* CeedBasisCreateH1 (interface/ceed-basis.c:1434): non-tensor H1 basis, interp [num_qpts x num_nodes], grad [dim][num_qpts x num_nodes].
 * Operators using it run on the unfused path (mixed-topology / composite operators). */
CEEDB200_EXPORT int ceedb200_basis_create_h1(B200Ceed ceed, b200_int dim, b200_int num_comp, b200_int num_nodes, b200_int num_qpts,
                                             const b200_scalar *interp, const b200_scalar *grad, const b200_scalar *q_ref, const b200_scalar *q_weight,
                                             B200Basis *basis);
CEEDB200_EXPORT int ceedb200_basis_create_tensor_h1_lagrange(B200Ceed ceed, b200_int dim, b200_int num_comp, b200_int P, b200_int Q, int quad_mode,
                                                             B200Basis *basis);
CEEDB200_EXPORT int ceedb200_basis_destroy(B200Basis basis);
CEEDB200_EXPORT int ceedb200_basis_apply(B200Basis basis, b200_int num_elem, int t_mode, int eval_mode, B200Vector u, B200Vector v);
CEEDB200_EXPORT int ceedb200_basis_apply_add(B200Basis basis, b200_int num_elem, int t_mode, int eval_mode, B200Vector u, B200Vector v);
CEEDB200_EXPORT int ceedb200_basis_apply_ptr(B200Basis basis, b200_int num_elem, int t_mode, int eval_mode, int add, const b200_scalar *d_u,
                                             b200_scalar *d_v);
/* which = 0 interp_1d [Q*P], 1 grad_1d [Q*P], 2 q_ref_1d [Q], 3 q_weight_1d [Q], 4 collocated grad [Q*Q] (interface/ceed-basis.c:750-773) */
CEEDB200_EXPORT int ceedb200_basis_get_matrix(B200Basis basis, int which, b200_scalar *out);
/* host utilities, no GPU needed (used by host-side tests): */
CEEDB200_EXPORT int ceedb200_host_gauss_quadrature(b200_int Q, b200_scalar *q_ref_1d, b200_scalar *q_weight_1d);
CEEDB200_EXPORT int ceedb200_host_lobatto_quadrature(b200_int Q, b200_scalar *q_ref_1d, b200_scalar *q_weight_1d);
CEEDB200_EXPORT int ceedb200_host_lagrange_1d(b200_int P, b200_int Q, int quad_mode, b200_scalar *interp_1d, b200_scalar *grad_1d,
                                              b200_scalar *q_ref_1d, b200_scalar *q_weight_1d);
CEEDB200_EXPORT int ceedb200_host_collocated_grad_1d(b200_int P, b200_int Q, const b200_scalar *interp_1d, const b200_scalar *grad_1d,
                                                     b200_scalar *collo_grad_1d);

/* ---------------------------------------------------------------- QFunction context (CeedQFunctionContext)
 * replaces CeedQFunctionContext_Cuda: backends/cuda-ref/ceed-cuda-ref.h:107-114, ceed-cuda-ref-qfunctioncontext.c:152-333 */
CEEDB200_EXPORT int ceedb200_qfcontext_create(B200Ceed ceed, B200QFContext *ctx);
CEEDB200_EXPORT int ceedb200_qfcontext_destroy(B200QFContext ctx);
CEEDB200_EXPORT int ceedb200_qfcontext_set_data(B200QFContext ctx, int mem_type, int copy_mode, size_t size, void *data);
CEEDB200_EXPORT int ceedb200_qfcontext_take_data(B200QFContext ctx, int mem_type, void **data);
CEEDB200_EXPORT int ceedb200_qfcontext_get_data(B200QFContext ctx, int mem_type, void **data);       /* read-write: invalidates the other copy */
CEEDB200_EXPORT int ceedb200_qfcontext_get_data_read(B200QFContext ctx, int mem_type, void **data);
CEEDB200_EXPORT int ceedb200_qfcontext_has_valid_data(B200QFContext ctx, int *has_valid);
CEEDB200_EXPORT int ceedb200_qfcontext_has_borrowed_data_of_type(B200QFContext ctx, int mem_type, int *has_borrowed);

/* ---------------------------------------------------------------- QFunction (CeedQFunction)
 * replaces CeedQFunctionCreate_Cuda / Apply: backends/cuda-ref/ceed-cuda-ref-qfunction.c:21-63, -qfunction-load.cpp:22-113.
 * `source_path` is the user's header (resolved absolute path or relative to a JIT source root), `kernel_name` the function
 * defined there with CEED_QFUNCTION(name) (interface/ceed-qfunction.c:259-320).  The source is JIT-compiled with NVRTC. */
CEEDB200_EXPORT int ceedb200_qfunction_create(B200Ceed ceed, const char *source_path, const char *kernel_name, B200QFunction *qf);
CEEDB200_EXPORT int ceedb200_qfunction_destroy(B200QFunction qf);
CEEDB200_EXPORT int ceedb200_qfunction_add_input(B200QFunction qf, const char *field_name, b200_int size, int eval_mode);
CEEDB200_EXPORT int ceedb200_qfunction_add_output(B200QFunction qf, const char *field_name, b200_int size, int eval_mode);
CEEDB200_EXPORT int ceedb200_qfunction_set_context(B200QFunction qf, B200QFContext ctx);
/* context owned by another backend (a libCEED operator may carry a CeedQFunctionContext of its fallback backend): raw device
 * pointer, valid for the applies that follow; obtained through CeedQFunctionGetInnerContextData(qf, CEED_MEM_DEVICE, ...) like
 * backends/cuda-gen/ceed-cuda-gen-operator.c:207 does */
CEEDB200_EXPORT int ceedb200_qfunction_set_context_ptr(B200QFunction qf, void *d_ctx);
/* standalone apply over Q points: U[i]/V[i] hold field i as [size_i][Q] (doc/sphinx/source/libCEEDdev.md:113-118) */
CEEDB200_EXPORT int ceedb200_qfunction_apply(B200QFunction qf, b200_int Q, const B200Vector *U, const B200Vector *V);

CEEDB200_EXPORT int ceedb200_qfunction_apply_ptr(B200QFunction qf, b200_int Q, const b200_scalar *const *d_in, b200_scalar *const *d_out);

/* ---------------------------------------------------------------- operator (CeedOperator)
 * replaces CeedOperatorCreate_Cuda_gen / ApplyAdd: backends/cuda-gen/ceed-cuda-gen-operator.c:105-300,879-908 and the kernel
 * generator backends/cuda-gen/ceed-cuda-gen-operator-build.cpp:1158-1681.
 * Field wiring follows CeedOperatorSetField (interface/ceed-operator.c:931-1038):
 *   rstr == NULL   <=> CEED_ELEMRESTRICTION_NONE (only with EVAL_WEIGHT)
 *   basis == NULL  <=> CEED_BASIS_NONE (EVAL_NONE fields)
 *   vec == B200_VECTOR_ACTIVE (the apply arguments) / B200_VECTOR_NONE (EVAL_WEIGHT) / a passive vector */
#define B200_VECTOR_ACTIVE ((B200Vector)(uintptr_t)1)
#define B200_VECTOR_NONE ((B200Vector)(uintptr_t)0)
CEEDB200_EXPORT int ceedb200_operator_create(B200Ceed ceed, B200QFunction qf, B200Operator *op);
CEEDB200_EXPORT int ceedb200_operator_destroy(B200Operator op);
CEEDB200_EXPORT int ceedb200_operator_set_field(B200Operator op, const char *field_name, B200Restriction rstr, B200Basis basis, B200Vector vec);
/* v = A u (overwrite; CeedOperatorApply, interface/ceed-operator.c:2271-2292) */
CEEDB200_EXPORT int ceedb200_operator_apply(B200Operator op, B200Vector u, B200Vector v);
/* v += A u (CeedOperatorApplyAdd, interface/ceed-operator.c:2313-2338) */
CEEDB200_EXPORT int ceedb200_operator_apply_add(B200Operator op, B200Vector u, B200Vector v);
/* introspection for tests / benchmarks */
CEEDB200_EXPORT int         ceedb200_operator_is_fused(B200Operator op, int *is_fused);
CEEDB200_EXPORT const char *ceedb200_operator_kernel_source(B200Operator op);
CEEDB200_EXPORT int         ceedb200_operator_kernel_info(B200Operator op, int *regs, int *smem_bytes, int *threads, int *elems_per_block,
                                                          int *grid, int *local_bytes);
/* device time (ms) of the most recent apply's kernels, measured with CUDA events on the context stream when enabled */
/* CeedOperatorLinearAssembleQFunction[Update] (interface/ceed-preconditioning.c; the reference's GPU twin is
 * backends/cuda-ref/ceed-cuda-ref-operator.c:1000-1130): the pointwise linear map of the QFunction at every quadrature point.
 * assembled (length num_elem * num_qpts * size_in * size_out) is laid out with strides {1, num_elem * num_qpts, num_qpts}:
 * entry ((a * size_out + b) * num_elem + e) * num_qpts + q = d out_b / d in_a, a / b counting the components of the ACTIVE
 * QFunction inputs / outputs in field order.  The interface builds LinearAssembleDiagonal, LinearAssemble and multigrid on it. */
CEEDB200_EXPORT int ceedb200_operator_assemble_qfunction_sizes(B200Operator op, b200_int *num_elem, b200_int *num_qpts, b200_int *size_in,
                                                               b200_int *size_out);
CEEDB200_EXPORT int ceedb200_operator_assemble_qfunction(B200Operator op, B200Vector assembled);
/* One half of an apply on a partitioned mesh (multi-GPU overlap; the reference pattern is VecScatterBegin / local work /
 * VecScatterEnd, examples/petsc/bpsraw.c:240-262): part 1 = elements [0, split) -- the ones touching the rank interface, see
 * ceedb200_restriction_set_split -- plus the finalize pass of the nodes only they touch; part 2 = the interior elements.
 * part 1 followed by part 2 produces exactly the bits of ceedb200_operator_apply. */
CEEDB200_EXPORT int ceedb200_operator_apply_part(B200Operator op, B200Vector u, B200Vector v, int part);
/* v = A u for an application that keeps its vectors in HOST memory: u valid on the host only, v with a host array (pinned memory both).
 * Replaces the sequence the reference performs for CeedVectorSetArray(HOST) -> CeedOperatorApply -> CeedVectorSyncArray(HOST): a whole-
 * vector host-to-device copy when the operator asks for the device array (backends/cuda-ref/ceed-cuda-ref-vector.c:40-76
 * CeedVectorSyncH2D_Cuda), the apply (backends/cuda-gen/ceed-cuda-gen-operator.c:99-300), and a whole-vector copy back
 * (ceed-cuda-ref-vector.c:101-130 CeedVectorSyncD2H_Cuda), one after the other.  Here the elements are cut into num_chunks (0 = default)
 * contiguous chunks; a chunk is applied as soon as the part of u it gathers has arrived and the part of v that no later chunk touches
 * is copied back while the next chunks run (two copy streams, PCIe is full duplex).  Bitwise the result of ceedb200_operator_apply; on
 * return v is valid on BOTH sides.  Falls back to ceedb200_operator_apply (*streamed = 0) when a precondition does not hold (operator not
 * fused, other scatter mode, u already valid on the device, pageable host memory, partially covered output, fewer than 4096 elements). */
CEEDB200_EXPORT int ceedb200_operator_apply_streamed(B200Operator op, B200Vector u, B200Vector v, int num_chunks, int *streamed);
/* Host-logic tests: description of the launch the fused apply would perform for (u, v) -- nothing is launched.  `args` is the kernel's
   argument block (the __constant__ B200OpArgs object of the generated source, see ceedb200_operator_kernel_source); the halo_* / v fields
   describe the second pass of the deterministic scatter (v[halo_node[i] + c comp_stride] += halo[j + c num_halo], j in
   [halo_ptr[i], halo_ptr[i + 1]), ascending) for output field `fin_slot` (-1: none); scatter_mode as ceedb200_set_scatter_mode. */
typedef struct {
  unsigned char args[1024];
  int           args_size, grid, threads, smem_bytes, run_mode, kernel_add, zero_first, fin_slot, num_comp;
  long long     e_begin, e_end, comp_stride, num_shared, num_halo;
  const int    *halo_node, *halo_ptr;
  const double *halo;
  double       *v;
  const char   *source; /* generated source of the kernel variant (store / accumulate) this apply launches */
  /* E-vector scatter mode: the kernel writes evec[e * elem_size + n + c * e_entries]; the transpose restriction that follows adds
     evec entry i of component c to v[offsets[i] + c comp_stride] in ascending i */
  int           scatter_mode;
  long long     e_entries;
  const int    *offsets;
  const double *evec;
} B200DebugLaunch;
CEEDB200_EXPORT int ceedb200_operator_debug_launch(B200Operator op, B200Vector u, B200Vector v, int add, int part, B200DebugLaunch *desc);
/* host-logic tests: the chunk tables of the streamed apply for num_chunks chunks -- element chunk ends, per chunk the number of leading
 * L-indices of u that must have arrived before it runs, and the number of leading L-indices of v that are final after it (INT64_MAX for
 * the last chunk); *per_comp = 1 when blocked components stream their own ranges */
CEEDB200_EXPORT int ceedb200_operator_debug_stream_plan(B200Operator op, int num_chunks, int32_t *ends, int64_t *in_hi, int64_t *out_done, int *per_comp);
CEEDB200_EXPORT int ceedb200_operator_set_timing(B200Operator op, int enabled);
CEEDB200_EXPORT int ceedb200_operator_last_kernel_ms(B200Operator op, float *fused_ms, float *aux_ms);
/* tuning override: elems_per_block (0 = heuristic), blocks_per_sm (0 = heuristic) */
CEEDB200_EXPORT int ceedb200_operator_set_tuning(B200Operator op, int elems_per_block, int blocks_per_sm);
/* full kernel shape: int[7] = {elems_per_group, group_warps, cta_warps, min_blocks_per_sm, qf_mode (0 z-line, 1 pointwise),
 * qf_unroll, stage_mask}; 0 (-1 for qf_mode / stage_mask) = heuristic.  The reference tunes the same knobs by hand per backend
 * (BlockGridCalculate, backends/cuda-gen/ceed-cuda-gen-operator.c:40-100); here they are data. */
CEEDB200_EXPORT int ceedb200_operator_set_kernel_shape(B200Operator op, const int *shape);
CEEDB200_EXPORT int ceedb200_operator_get_kernel_shape(B200Operator op, int *shape, char *signature, int signature_len);
/* autotuner: 0 off (default), 1 time candidate kernel shapes on the first Apply of every fused operator that has no entry
 * in the tuning table, 2 always.  Same as the environment variable CEED_B200_AUTOTUNE. */
CEEDB200_EXPORT int ceedb200_set_autotune(B200Ceed ceed, int level);

/* ---- multi-GPU interface exchange (raw device pointers; the transport between the two calls is NCCL send/recv) ----------
 * The reference has no communication layer: its examples delegate the interface sum to PETSc VecScatter ADD_VALUES
 * (examples/petsc/bpsraw.c:240-262).  pack: send[i] = v[idx[i]].  unpack_sum: for interface entry i (L-index node[i]) sum the
 * contributions src[ptr[i]..ptr[i+1]) in the given (ascending rank) order, src < 0 = this rank's own value v[node[i]],
 * src >= 0 = recv[src]; the result overwrites v[node[i]]. */
CEEDB200_EXPORT int ceedb200_iface_pack(B200Ceed ceed, const double *d_v, const long long *d_idx, long long n, double *d_send);
CEEDB200_EXPORT int ceedb200_iface_unpack_sum(B200Ceed ceed, double *d_v, long long n, const long long *d_node, const int *d_ptr, const int *d_src,
                                              const double *d_recv);

/* The same exchange over NVLink peer memory (no NCCL on the data path): `put` gathers the interface values and stores them
 * directly into the neighbours' receive areas (peer-mapped device pointers obtained through ceedb200_ipc_*), then raises one
 * flag per neighbour; `wait_unpack_sum` on the receiver waits for its flags and forms the rank-ordered sums.  Layout of a
 * rank's exposed allocation: double recv[2][half] (halves alternate with the parity of the device-side step counter), then
 * long long flags[num_nb].  d_ctr = {last completed step, CTAs done} (zero-initialised, private to the rank).  `stream` (nullable:
 * the context stream) lets the put run on a high-priority stream concurrently with the interior elements. */
CEEDB200_EXPORT int ceedb200_iface_put(B200Ceed ceed, void *stream, const double *d_v, const long long *d_idx, const int *d_nb_of, const long long *d_seg,
                                       long long n, int num_nb, double *const *d_peer_recv, const long long *d_peer_half, long long *const *d_peer_flag,
                                       long long *d_ctr);
CEEDB200_EXPORT int ceedb200_iface_wait_unpack_sum(B200Ceed ceed, double *d_v, long long n, const long long *d_node, const int *d_ptr, const int *d_src,
                                                   const double *d_recv, long long half, const long long *d_flags, int num_nb, const long long *d_ctr);
/* CUDA IPC plumbing (one process per GPU): allocate a zeroed device buffer and export its 64-byte handle / map a peer's buffer */
CEEDB200_EXPORT int ceedb200_ipc_alloc(B200Ceed ceed, size_t bytes, void **d_ptr, unsigned char *handle64);
CEEDB200_EXPORT int ceedb200_ipc_open(B200Ceed ceed, const unsigned char *handle64, void **d_ptr);
CEEDB200_EXPORT int ceedb200_ipc_close(B200Ceed ceed, void *d_ptr);
CEEDB200_EXPORT int ceedb200_ipc_free(B200Ceed ceed, void *d_ptr);

/* ---- device-resident conjugate-gradient pieces (the reference's figure of merit is
 * "DoFs/sec in CG", examples/petsc/bps.c:218-288, there with PETSc KSPCG).  Raw device pointers; all scalars live on the
 * device, so one CG iteration = operator apply + these calls needs no host synchronisation.  d_w (nullable) weights the dot
 * products (1 = entry owned by this rank, 0 = copy of an interface node).
 *   dot:       *d_out = sum_i w_i x_i y_i                                       (deterministic two-pass reduction)
 *   update:    alpha = *d_rr / *d_pAp;  x += alpha p;  r -= alpha Ap;  *d_rr_new = sum_i w_i r_i^2
 *   direction: beta = *d_rr_new / *d_rr;  p = r + beta p */
CEEDB200_EXPORT int ceedb200_cg_dot(B200Ceed ceed, const double *d_x, const double *d_y, const double *d_w, long long n, double *d_out);
CEEDB200_EXPORT int ceedb200_cg_update(B200Ceed ceed, double *d_x, double *d_r, const double *d_p, const double *d_Ap, const double *d_w, long long n,
                                       const double *d_rr, const double *d_pAp, double *d_rr_new);
/* essential boundary conditions in the CG loop: Ap[i] = p[i] wherever free_mask[i] == 0 (identity rows for constrained DoFs; with b and
 * the start vector zero there, the iterates stay in the constrained subspace -- what PETSc's DMPlex does for examples/petsc/bps.c by
 * leaving the constrained DoFs out of the global vector, examples/petsc/src/petscutils.c) */
CEEDB200_EXPORT int ceedb200_cg_constrain(B200Ceed ceed, double *d_Ap, const double *d_p, const double *d_free_mask, long long n);
CEEDB200_EXPORT int ceedb200_cg_direction(B200Ceed ceed, double *d_p, const double *d_r, long long n, const double *d_rr_new, const double *d_rr);

#ifdef __cplusplus
}
#endif
#endif /* CEED_B200_H */
