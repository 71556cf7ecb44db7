// ceed-cuda-b200.h -- private data of the "/gpu/cuda/b200" libCEED backend.
//
// This directory is the reference-side binding of the C ABI in include/ceed_b200.h: plain C that fills libCEED's
// per-object function-pointer slots (CeedSetBackendFunction, include/ceed/backend.h:246) with thin adapters around the
// ceedb200_* entry points.  It can be built in-tree (drop the files into backends/cuda-b200/ and add one CEED_BACKEND
// line to backends/ceed-backend-list-cuda.h) or, as done here, out-of-tree as a plugin shared library whose
// constructor calls CeedRegister() before the first CeedInit().
#ifndef CEED_CUDA_B200_H
#define CEED_CUDA_B200_H

#include <ceed.h>
#include <ceed/backend.h>
#include <stdbool.h>

#include "../../include/ceed_b200.h"

typedef struct {
  B200Ceed core;
  int      device_id;
  bool     has_fallback;  // /gpu/cuda/ref is the delegate for object types and the fallback for operators this backend does not cover
  int      num_roots_seen, num_defines_seen;
} Ceed_B200;

typedef struct {
  B200Vector core;
} CeedVector_B200;

typedef struct {
  B200Restriction core;
} CeedElemRestriction_B200;

typedef struct {
  B200Basis core;
} CeedBasis_B200;

typedef struct {
  B200QFContext core;
} CeedQFunctionContext_B200;

typedef struct {
  B200QFunction        core;
  bool                 fields_set;
  CeedQFunctionContext foreign_ctx;  // context of another backend (borrowed): accessed through the interface at apply time
} CeedQFunction_B200;

typedef struct {
  B200Operator core;
  B200Vector   view_in, view_out;     // device-pointer views of the active vectors of the current apply (cached per length)
  CeedSize     view_in_len, view_out_len;
  B200Vector   passive_in[CEED_FIELD_MAX], passive_out[CEED_FIELD_MAX];
  CeedVector   passive_in_vec[CEED_FIELD_MAX], passive_out_vec[CEED_FIELD_MAX];
  CeedInt      num_in, num_out;
  bool         is_setup;
  bool         use_fallback;  // operator with objects of the delegate backend (non-tensor basis, ...): applied through CeedOperatorGetFallback
} CeedOperator_B200;

// error plumbing: turn a ceedb200_* failure into a CeedError carrying the core's message
#define CeedCallB200(ceed, core, ...)                                                          \
  do {                                                                                         \
    int ierr_b200_ = (__VA_ARGS__);                                                            \
    if (ierr_b200_) return CeedError((ceed), CEED_ERROR_BACKEND, "%s", ceedb200_last_error(core)); \
  } while (0)

CEED_INTERN int CeedGetCore_B200(Ceed ceed, B200Ceed *core);
CEED_INTERN int CeedSyncJitOptions_B200(Ceed ceed);

CEED_INTERN int CeedVectorCreate_B200(CeedSize n, CeedVector vec);
CEED_INTERN int CeedElemRestrictionCreate_B200(CeedMemType mem_type, CeedCopyMode copy_mode, const CeedInt *offsets, const bool *orients,
                                               const CeedInt8 *curl_orients, CeedElemRestriction rstr);
CEED_INTERN int CeedBasisCreateTensorH1_B200(CeedInt dim, CeedInt P_1d, CeedInt Q_1d, const CeedScalar *interp_1d, const CeedScalar *grad_1d,
                                             const CeedScalar *q_ref_1d, const CeedScalar *q_weight_1d, CeedBasis basis);
CEED_INTERN int CeedBasisCreateH1_B200(CeedElemTopology topo, CeedInt dim, CeedInt num_nodes, CeedInt num_qpts, const CeedScalar *interp, const CeedScalar *grad,
                                       const CeedScalar *q_ref, const CeedScalar *q_weight, CeedBasis basis);
CEED_INTERN int CeedQFunctionCreate_B200(CeedQFunction qf);
CEED_INTERN int CeedQFunctionContextCreate_B200(CeedQFunctionContext ctx);
CEED_INTERN int CeedOperatorCreate_B200(CeedOperator op);
CEED_INTERN int CeedOperatorLinearAssembleAddDiagonal_B200(CeedOperator op, CeedVector assembled, CeedRequest *request);
CEED_INTERN int CeedOperatorLinearAssembleAddPointBlockDiagonal_B200(CeedOperator op, CeedVector assembled, CeedRequest *request);
CEED_INTERN int CeedQFunctionContextAcquire_B200(CeedQFunction qf, void **held);
CEED_INTERN int CeedQFunctionContextRelease_B200(CeedQFunction qf, void **held);
CEED_INTERN int CeedQFunctionGetCore_B200(CeedQFunction qf, B200QFunction *core);

#endif
