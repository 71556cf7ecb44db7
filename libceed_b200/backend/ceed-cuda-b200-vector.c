// ceed-cuda-b200-vector.c -- CeedVector slots -> ceedb200_vector_* (replaces backends/cuda-ref/ceed-cuda-ref-vector.c:650-700 wiring)
#include <stdlib.h>

#include "ceed-cuda-b200.h"

static inline int GetCore(CeedVector vec, B200Vector *core) {
  CeedVector_B200 *impl;

  CeedCallBackend(CeedVectorGetData(vec, &impl));
  *core = impl->core;
  return CEED_ERROR_SUCCESS;
}

#define VEC_CALL(vec, ...)                                                                                           \
  do {                                                                                                               \
    B200Ceed core_ceed_;                                                                                             \
    CeedCallBackend(CeedGetCore_B200(CeedVectorReturnCeed(vec), &core_ceed_));                                       \
    CeedCallB200(CeedVectorReturnCeed(vec), core_ceed_, __VA_ARGS__);                                                \
  } while (0)

static int CeedVectorHasValidArray_B200(CeedVector vec, bool *has_valid_array) {
  B200Vector v;
  int        flag;

  CeedCallBackend(GetCore(vec, &v));
  VEC_CALL(vec, ceedb200_vector_has_valid_array(v, &flag));
  *has_valid_array = flag;
  return CEED_ERROR_SUCCESS;
}

static int CeedVectorHasBorrowedArrayOfType_B200(CeedVector vec, CeedMemType mem_type, bool *has_borrowed) {
  B200Vector v;
  int        flag;

  CeedCallBackend(GetCore(vec, &v));
  VEC_CALL(vec, ceedb200_vector_has_borrowed_array_of_type(v, mem_type, &flag));
  *has_borrowed = flag;
  return CEED_ERROR_SUCCESS;
}

static int CeedVectorSetArray_B200(CeedVector vec, CeedMemType mem_type, CeedCopyMode copy_mode, CeedScalar *array) {
  B200Vector v;

  CeedCallBackend(GetCore(vec, &v));
  VEC_CALL(vec, ceedb200_vector_set_array(v, mem_type, copy_mode, array));
  return CEED_ERROR_SUCCESS;
}

static int CeedVectorTakeArray_B200(CeedVector vec, CeedMemType mem_type, CeedScalar **array) {
  B200Vector v;

  CeedCallBackend(GetCore(vec, &v));
  VEC_CALL(vec, ceedb200_vector_take_array(v, mem_type, array));
  return CEED_ERROR_SUCCESS;
}

static int CeedVectorSetValue_B200(CeedVector vec, CeedScalar value) {
  B200Vector v;

  CeedCallBackend(GetCore(vec, &v));
  VEC_CALL(vec, ceedb200_vector_set_value(v, value));
  return CEED_ERROR_SUCCESS;
}

static int CeedVectorSetValueStrided_B200(CeedVector vec, CeedSize start, CeedSize stop, CeedSize step, CeedScalar value) {
  B200Vector v;

  CeedCallBackend(GetCore(vec, &v));
  VEC_CALL(vec, ceedb200_vector_set_value_strided(v, start, stop, step, value));
  return CEED_ERROR_SUCCESS;
}

static int CeedVectorCopyStrided_B200(CeedVector vec, CeedSize start, CeedSize stop, CeedSize step, CeedVector vec_copy) {
  B200Vector v, w;

  CeedCallBackend(GetCore(vec, &v));
  CeedCallBackend(GetCore(vec_copy, &w));
  VEC_CALL(vec, ceedb200_vector_copy_strided(v, start, stop, step, w));
  return CEED_ERROR_SUCCESS;
}

static int CeedVectorSyncArray_B200(CeedVector vec, CeedMemType mem_type) {
  B200Vector v;

  CeedCallBackend(GetCore(vec, &v));
  VEC_CALL(vec, ceedb200_vector_sync_array(v, mem_type));
  return CEED_ERROR_SUCCESS;
}

static int CeedVectorGetArray_B200(CeedVector vec, CeedMemType mem_type, CeedScalar **array) {
  B200Vector v;

  CeedCallBackend(GetCore(vec, &v));
  VEC_CALL(vec, ceedb200_vector_get_array(v, mem_type, array));
  return CEED_ERROR_SUCCESS;
}

static int CeedVectorGetArrayRead_B200(CeedVector vec, CeedMemType mem_type, const CeedScalar **array) {
  B200Vector v;

  CeedCallBackend(GetCore(vec, &v));
  VEC_CALL(vec, ceedb200_vector_get_array_read(v, mem_type, array));
  return CEED_ERROR_SUCCESS;
}

static int CeedVectorGetArrayWrite_B200(CeedVector vec, CeedMemType mem_type, CeedScalar **array) {
  B200Vector v;

  CeedCallBackend(GetCore(vec, &v));
  VEC_CALL(vec, ceedb200_vector_get_array_write(v, mem_type, array));
  return CEED_ERROR_SUCCESS;
}

static int CeedVectorNorm_B200(CeedVector vec, CeedNormType type, CeedScalar *norm) {
  B200Vector v;

  CeedCallBackend(GetCore(vec, &v));
  VEC_CALL(vec, ceedb200_vector_norm(v, type, norm));
  return CEED_ERROR_SUCCESS;
}

static int CeedVectorScale_B200(CeedVector x, CeedScalar alpha) {
  B200Vector v;

  CeedCallBackend(GetCore(x, &v));
  VEC_CALL(x, ceedb200_vector_scale(v, alpha));
  return CEED_ERROR_SUCCESS;
}

static int CeedVectorReciprocal_B200(CeedVector x) {
  B200Vector v;

  CeedCallBackend(GetCore(x, &v));
  VEC_CALL(x, ceedb200_vector_reciprocal(v));
  return CEED_ERROR_SUCCESS;
}

static int CeedVectorFilter_B200(CeedVector x, CeedScalar epsilon) {
  B200Vector v;

  CeedCallBackend(GetCore(x, &v));
  VEC_CALL(x, ceedb200_vector_filter(v, epsilon));
  return CEED_ERROR_SUCCESS;
}

static int CeedVectorAXPY_B200(CeedVector y, CeedScalar alpha, CeedVector x) {
  B200Vector vy, vx;

  CeedCallBackend(GetCore(y, &vy));
  CeedCallBackend(GetCore(x, &vx));
  VEC_CALL(y, ceedb200_vector_axpy(vy, alpha, vx));
  return CEED_ERROR_SUCCESS;
}

static int CeedVectorAXPBY_B200(CeedVector y, CeedScalar alpha, CeedScalar beta, CeedVector x) {
  B200Vector vy, vx;

  CeedCallBackend(GetCore(y, &vy));
  CeedCallBackend(GetCore(x, &vx));
  VEC_CALL(y, ceedb200_vector_axpby(vy, alpha, beta, vx));
  return CEED_ERROR_SUCCESS;
}

static int CeedVectorPointwiseMult_B200(CeedVector w, CeedVector x, CeedVector y) {
  B200Vector vw, vx, vy;

  CeedCallBackend(GetCore(w, &vw));
  CeedCallBackend(GetCore(x, &vx));
  CeedCallBackend(GetCore(y, &vy));
  VEC_CALL(w, ceedb200_vector_pointwise_mult(vw, vx, vy));
  return CEED_ERROR_SUCCESS;
}

static int CeedVectorDestroy_B200(CeedVector vec) {
  CeedVector_B200 *impl;

  CeedCallBackend(CeedVectorGetData(vec, &impl));
  ceedb200_vector_destroy(impl->core);
  free(impl);
  return CEED_ERROR_SUCCESS;
}

int CeedVectorCreate_B200(CeedSize n, CeedVector vec) {
  Ceed             ceed = CeedVectorReturnCeed(vec);
  B200Ceed         core;
  CeedVector_B200 *impl;

  CeedCallBackend(CeedGetCore_B200(ceed, &core));
  impl = calloc(1, sizeof(*impl));
  CeedCallB200(ceed, core, ceedb200_vector_create(core, n, &impl->core));
  CeedCallBackend(CeedVectorSetData(vec, impl));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Vector", vec, "HasValidArray", CeedVectorHasValidArray_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Vector", vec, "HasBorrowedArrayOfType", CeedVectorHasBorrowedArrayOfType_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Vector", vec, "CopyStrided", CeedVectorCopyStrided_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Vector", vec, "SetArray", CeedVectorSetArray_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Vector", vec, "TakeArray", CeedVectorTakeArray_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Vector", vec, "SetValue", CeedVectorSetValue_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Vector", vec, "SetValueStrided", CeedVectorSetValueStrided_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Vector", vec, "SyncArray", CeedVectorSyncArray_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Vector", vec, "GetArray", CeedVectorGetArray_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Vector", vec, "GetArrayRead", CeedVectorGetArrayRead_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Vector", vec, "GetArrayWrite", CeedVectorGetArrayWrite_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Vector", vec, "Norm", CeedVectorNorm_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Vector", vec, "Scale", CeedVectorScale_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Vector", vec, "Filter", CeedVectorFilter_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Vector", vec, "Reciprocal", CeedVectorReciprocal_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Vector", vec, "AXPY", CeedVectorAXPY_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Vector", vec, "AXPBY", CeedVectorAXPBY_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Vector", vec, "PointwiseMult", CeedVectorPointwiseMult_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Vector", vec, "Destroy", CeedVectorDestroy_B200));
  return CEED_ERROR_SUCCESS;
}
