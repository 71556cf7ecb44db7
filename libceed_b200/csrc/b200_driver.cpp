// b200_driver.cpp -- lazy binding of the handful of CUDA driver API calls the JIT path needs (see b200_driver.h)
#include "b200_driver.h"

#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>

static CUresult missing_error_string(CUresult, const char **msg) {
  *msg = "libcuda.so.1 not found: ceed-b200 needs an NVIDIA driver";
  return CUDA_SUCCESS;
}
template <typename... A>
static CUresult missing(A...) {
  return CUDA_ERROR_NOT_INITIALIZED;
}

const B200Driver *b200_driver() {
  static B200Driver drv;
  static bool       loaded = false;
  if (loaded) return &drv;
  loaded       = true;
  void *handle = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!handle) handle = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
  if (!handle) {
    drv.GetErrorString                            = missing_error_string;
    drv.ModuleLoadData                            = missing<CUmodule *, const void *>;
    drv.ModuleUnload                              = missing<CUmodule>;
    drv.ModuleGetFunction                         = missing<CUfunction *, CUmodule, const char *>;
    drv.ModuleGetGlobal                           = missing<CUdeviceptr *, size_t *, CUmodule, const char *>;
    drv.LaunchKernel                              = missing<CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **, void **>;
    drv.LaunchCooperativeKernel                   = missing<CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **>;
    drv.FuncSetAttribute                          = missing<CUfunction, CUfunction_attribute, int>;
    drv.FuncGetAttribute                          = missing<int *, CUfunction_attribute, CUfunction>;
    drv.OccupancyMaxActiveBlocksPerMultiprocessor = missing<int *, CUfunction, int, size_t>;
    return &drv;
  }
#define B200_SYM(field, name)                                             \
  *(void **)(&drv.field) = dlsym(handle, name);                           \
  if (!drv.field) {                                                       \
    fprintf(stderr, "ceed-b200: symbol %s missing from libcuda\n", name); \
    abort();                                                              \
  }
  B200_SYM(GetErrorString, "cuGetErrorString")
  B200_SYM(ModuleLoadData, "cuModuleLoadData")
  B200_SYM(ModuleUnload, "cuModuleUnload")
  B200_SYM(ModuleGetFunction, "cuModuleGetFunction")
  B200_SYM(ModuleGetGlobal, "cuModuleGetGlobal_v2")
  B200_SYM(LaunchKernel, "cuLaunchKernel")
  B200_SYM(LaunchCooperativeKernel, "cuLaunchCooperativeKernel")
  B200_SYM(FuncSetAttribute, "cuFuncSetAttribute")
  B200_SYM(FuncGetAttribute, "cuFuncGetAttribute")
  B200_SYM(OccupancyMaxActiveBlocksPerMultiprocessor, "cuOccupancyMaxActiveBlocksPerMultiprocessor")
#undef B200_SYM
  return &drv;
}
