// ceed-cuda-b200.c -- registration and Ceed-level setup of the "/gpu/cuda/b200" backend.
//
// Replaces, for this resource, what backends/cuda-gen/ceed-cuda-gen.c:19-52 + backends/cuda/ceed-cuda-common.c:19-35 do for
// "/gpu/cuda/gen": validate the resource, pick the device (":device_id=N"), create the backend context and install the
// object constructors by name.  Nothing is delegated: Vector, ElemRestriction, tensor H1 Basis, QFunction, QFunctionContext
// and Operator are all implemented by this backend on top of include/ceed_b200.h.  Object types outside the operator-apply
// path (non-tensor bases, H(div)/H(curl), at-points, composite operators on streams, assembly) report
// "Backend does not implement ..." exactly like other backends do for their gaps (tests/junit.py:121-145 treats it as skip).
#include "ceed-cuda-b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int CeedGetCore_B200(Ceed ceed, B200Ceed *core) {
  Ceed_B200 *data;

  CeedCallBackend(CeedGetData(ceed, &data));
  *core = data->core;
  return CEED_ERROR_SUCCESS;
}

// Forward CeedAddJitSourceRoot / CeedAddJitDefine (interface/ceed.c:1515,1579) to the core before anything is compiled
int CeedSyncJitOptions_B200(Ceed ceed) {
  Ceed_B200    *data;
  CeedInt       num_roots = 0, num_defines = 0;
  const char  **roots, **defines;

  CeedCallBackend(CeedGetData(ceed, &data));
  CeedCallBackend(CeedGetJitSourceRoots(ceed, &num_roots, &roots));
  for (CeedInt i = data->num_roots_seen; i < num_roots; i++) ceedb200_add_jit_source_root(data->core, roots[i]);
  data->num_roots_seen = num_roots;
  CeedCallBackend(CeedRestoreJitSourceRoots(ceed, &roots));
  CeedCallBackend(CeedGetJitDefines(ceed, &num_defines, &defines));
  for (CeedInt i = data->num_defines_seen; i < num_defines; i++) ceedb200_add_jit_define(data->core, defines[i]);
  data->num_defines_seen = num_defines;
  CeedCallBackend(CeedRestoreJitDefines(ceed, &defines));
  return CEED_ERROR_SUCCESS;
}

static int CeedDestroy_B200(Ceed ceed) {
  Ceed_B200 *data;

  CeedCallBackend(CeedGetData(ceed, &data));
  if (data) {
    ceedb200_destroy(data->core);
    free(data);
  }
  return CEED_ERROR_SUCCESS;
}

static int CeedGetPreferredMemType_B200(CeedMemType *mem_type) {
  *mem_type = CEED_MEM_DEVICE;
  return CEED_ERROR_SUCCESS;
}

static int CeedSetStream_B200(Ceed ceed, void *handle) {
  B200Ceed core;

  CeedCallBackend(CeedGetCore_B200(ceed, &core));
  ceedb200_set_stream(core, handle);
  return CEED_ERROR_SUCCESS;
}

static int CeedInit_B200(const char *resource, Ceed ceed) {
  Ceed_B200  *data;
  const char *root = "/gpu/cuda/b200", *colon = strchr(resource, ':');
  size_t      root_len = colon ? (size_t)(colon - resource) : strlen(resource);
  int         device_id = -1;

  // "/gpu/cuda/b200", a prefix of it that resolved here ("/gpu/cuda"), optionally followed by ":device_id=N"
  CeedCheck(root_len <= strlen(root) && !strncmp(resource, root, root_len), ceed, CEED_ERROR_BACKEND, "B200 backend cannot use resource: %s", resource);
  if (colon) {
    const char *id = strstr(colon, "device_id=");
    if (id) device_id = atoi(id + strlen("device_id="));
  }
  data = calloc(1, sizeof(*data));
  CeedCheck(data, ceed, CEED_ERROR_MAJOR, "allocation failed");
  if (ceedb200_init(device_id, &data->core)) {
    free(data);
    return CeedError(ceed, CEED_ERROR_BACKEND, "%s", ceedb200_last_error(NULL));
  }
  data->device_id = device_id < 0 ? 0 : device_id;
  CeedCallBackend(CeedSetData(ceed, data));
  {
    // deterministic unless the caller opted into the atomic scatter mode (CEED_B200_SCATTER=atomic|1): ordered scatter, no atomics
    // (README.md:167-169 semantics)
    const char *mode = getenv("CEED_B200_SCATTER");
    CeedCallBackend(CeedSetDeterministic(ceed, !(mode && (!strcmp(mode, "atomic") || !strcmp(mode, "1")))));
  }

  CeedCallBackend(CeedSetBackendFunction(ceed, "Ceed", ceed, "Destroy", CeedDestroy_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Ceed", ceed, "GetPreferredMemType", CeedGetPreferredMemType_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Ceed", ceed, "SetStream", CeedSetStream_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Ceed", ceed, "VectorCreate", CeedVectorCreate_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Ceed", ceed, "ElemRestrictionCreate", CeedElemRestrictionCreate_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Ceed", ceed, "BasisCreateTensorH1", CeedBasisCreateTensorH1_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Ceed", ceed, "BasisCreateH1", CeedBasisCreateH1_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Ceed", ceed, "QFunctionCreate", CeedQFunctionCreate_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Ceed", ceed, "QFunctionContextCreate", CeedQFunctionContextCreate_B200));
  CeedCallBackend(CeedSetBackendFunction(ceed, "Ceed", ceed, "OperatorCreate", CeedOperatorCreate_B200));

  // Opt-in (CEED_B200_FALLBACK=1, needs a libCEED with the CUDA backends): like /gpu/cuda/gen delegates to /gpu/cuda/shared
  // and falls back to /gpu/cuda/ref (backends/cuda-gen/ceed-cuda-gen.c:33-39), object types this backend does not create
  // (non-tensor / H(div) / H(curl) bases, at-points restrictions, composite operators) come from /gpu/cuda/ref, operators that
  // contain such objects are applied through their fallback operator, and the LinearAssemble* family runs on the fallback.
  // The hot path -- 3-D tensor operators -- is unaffected.
  if (getenv("CEED_B200_FALLBACK") && atoi(getenv("CEED_B200_FALLBACK"))) {
    Ceed ceed_ref;
    char ref_resource[64];

    snprintf(ref_resource, sizeof(ref_resource), "/gpu/cuda/ref:device_id=%d", data->device_id);
    CeedCallBackend(CeedInit(ref_resource, &ceed_ref));
    CeedCallBackend(CeedSetDelegate(ceed, ceed_ref));
    CeedCallBackend(CeedSetOperatorFallbackCeed(ceed, ceed_ref));
    CeedCallBackend(CeedDestroy(&ceed_ref));
    data->has_fallback = true;
  }
  return CEED_ERROR_SUCCESS;
}

// In-tree: listed as CEED_BACKEND(CeedRegister_Cuda_B200, 1, "/gpu/cuda/b200") in backends/ceed-backend-list-cuda.h.
// Priority 45 > 40 (/gpu/cuda/ref, backends/cuda-ref/ceed-cuda-ref.c:72; lower wins): prefix matches such as "/gpu/cuda" keep
// resolving to the existing backends, this one is selected by its full name only.
CEED_EXTERN int CeedRegister_Cuda_B200(void);
int CeedRegister_Cuda_B200(void) { return CeedRegister("/gpu/cuda/b200", CeedInit_B200, 45); }

#ifdef CEED_B200_PLUGIN
// Out-of-tree plugin: register before main() / at dlopen() time; libCEED's registry is a static table, no init order issue.
__attribute__((constructor)) static void CeedRegister_Cuda_B200_Plugin(void) { CeedRegister_Cuda_B200(); }
#endif
