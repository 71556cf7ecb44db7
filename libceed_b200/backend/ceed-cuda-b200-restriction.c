// ceed-cuda-b200-restriction.c -- CeedElemRestriction slots -> ceedb200_restriction_*
// (replaces the wiring of backends/cuda-ref/ceed-cuda-ref-restriction.c:498-661; standard and strided restrictions)
#include <stdbool.h>
#include <stdlib.h>

#include "ceed-cuda-b200.h"

static int CeedElemRestrictionApply_B200(CeedElemRestriction rstr, CeedTransposeMode t_mode, CeedVector u, CeedVector v, CeedRequest *request) {
  Ceed                      ceed = CeedElemRestrictionReturnCeed(rstr);
  B200Ceed                  core;
  CeedElemRestriction_B200 *impl;
  const CeedScalar         *d_u;
  CeedScalar               *d_v;

  CeedCallBackend(CeedGetCore_B200(ceed, &core));
  CeedCallBackend(CeedElemRestrictionGetData(rstr, &impl));
  // device arrays through the interface, so vectors of any CUDA-family backend are accepted; whatever was acquired is restored
  // on every path (libCEED's access locks), the first error is reported
  {
    int  status, ierr2;
    bool got_u = false, got_v = false;

    status = CeedVectorGetArrayRead(u, CEED_MEM_DEVICE, &d_u);
    got_u  = !status;
    if (!status) {
      status = t_mode == CEED_TRANSPOSE ? CeedVectorGetArray(v, CEED_MEM_DEVICE, &d_v)        // sums into v
                                        : CeedVectorGetArrayWrite(v, CEED_MEM_DEVICE, &d_v);  // overwrites the E-vector
      got_v  = !status;
    }
    if (!status && ceedb200_restriction_apply_ptr(impl->core, t_mode, d_u, d_v)) status = CeedError(ceed, CEED_ERROR_BACKEND, "%s", ceedb200_last_error(core));
    if (got_u && (ierr2 = CeedVectorRestoreArrayRead(u, &d_u)) && !status) status = ierr2;
    if (got_v && (ierr2 = CeedVectorRestoreArray(v, &d_v)) && !status) status = ierr2;
    return status;
  }
}

static int CeedElemRestrictionGetOffsets_B200(CeedElemRestriction rstr, CeedMemType mem_type, const CeedInt **offsets) {
  Ceed                      ceed = CeedElemRestrictionReturnCeed(rstr);
  B200Ceed                  core;
  CeedElemRestriction_B200 *impl;

  CeedCallBackend(CeedGetCore_B200(ceed, &core));
  CeedCallBackend(CeedElemRestrictionGetData(rstr, &impl));
  CeedCallB200(ceed, core, ceedb200_restriction_get_offsets(impl->core, mem_type, offsets));
  return CEED_ERROR_SUCCESS;
}

static int CeedElemRestrictionDestroy_B200(CeedElemRestriction rstr) {
  CeedElemRestriction_B200 *impl;

  CeedCallBackend(CeedElemRestrictionGetData(rstr, &impl));
  ceedb200_restriction_destroy(impl->core);
  free(impl);
  return CEED_ERROR_SUCCESS;
}

int CeedElemRestrictionCreate_B200(CeedMemType mem_type, CeedCopyMode copy_mode, const CeedInt *offsets, const bool *orients,
                                   const CeedInt8 *curl_orients, CeedElemRestriction rstr) {
  Ceed                      ceed = CeedElemRestrictionReturnCeed(rstr);
  B200Ceed                  core;
  CeedElemRestriction_B200 *impl;
  CeedRestrictionType       rstr_type;
  CeedInt                   num_elem, elem_size, num_comp, comp_stride = 0;
  CeedSize                  l_size;

  CeedCallBackend(CeedGetCore_B200(ceed, &core));
  CeedCallBackend(CeedElemRestrictionGetType(rstr, &rstr_type));
  CeedCheck(rstr_type == CEED_RESTRICTION_STANDARD || rstr_type == CEED_RESTRICTION_STRIDED, ceed, CEED_ERROR_UNSUPPORTED,
            "Backend does not implement oriented, curl-oriented or at-points element restrictions");
  CeedCallBackend(CeedElemRestrictionGetNumElements(rstr, &num_elem));
  CeedCallBackend(CeedElemRestrictionGetElementSize(rstr, &elem_size));
  CeedCallBackend(CeedElemRestrictionGetNumComponents(rstr, &num_comp));
  CeedCallBackend(CeedElemRestrictionGetLVectorSize(rstr, &l_size));
  impl = calloc(1, sizeof(*impl));
  if (rstr_type == CEED_RESTRICTION_STRIDED) {
    bool    has_backend_strides;
    CeedInt strides[3] = {0, 0, 0};

    CeedCallBackend(CeedElemRestrictionHasBackendStrides(rstr, &has_backend_strides));
    if (!has_backend_strides) CeedCallBackend(CeedElemRestrictionGetStrides(rstr, strides));
    CeedCallB200(ceed, core, ceedb200_restriction_create_strided(core, num_elem, elem_size, num_comp, l_size, has_backend_strides ? NULL : strides, &impl->core));
    if (has_backend_strides) {
      // publish the backend's choice, same as the CUDA family (backends/cuda-ref/ceed-cuda-ref-restriction.c:536-541)
      CeedInt l_layout[3] = {1, elem_size * num_elem, elem_size};
      CeedCallBackend(CeedElemRestrictionSetLLayout(rstr, l_layout));
    }
  } else {
    CeedCallBackend(CeedElemRestrictionGetCompStride(rstr, &comp_stride));
    // OWN_POINTER hands the array to the core, which frees it with free(): only valid for host malloc'ed arrays, which is
    // what the interface passes (interface/ceed-elemrestriction.c:657-687)
    CeedCallB200(ceed, core,
                 ceedb200_restriction_create(core, num_elem, elem_size, num_comp, comp_stride, l_size, mem_type, copy_mode, offsets, &impl->core));
  }
  CeedCallBackend(CeedElemRestrictionSetData(rstr, impl));
  {
    CeedInt e_layout[3] = {1, elem_size * num_elem, elem_size};  // [comp][elem][node], cuda-ref-restriction.c:530-534
    CeedCallBackend(CeedElemRestrictionSetELayout(rstr, e_layout));
  }
  CeedCallBackend(CeedSetBackendFunction(ceed, "ElemRestriction", rstr, "Apply", CeedElemRestrictionApply_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "ElemRestriction", rstr, "ApplyUnsigned", CeedElemRestrictionApply_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "ElemRestriction", rstr, "ApplyUnoriented", CeedElemRestrictionApply_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "ElemRestriction", rstr, "GetOffsets", CeedElemRestrictionGetOffsets_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "ElemRestriction", rstr, "Destroy", CeedElemRestrictionDestroy_B200));
  return CEED_ERROR_SUCCESS;
}
