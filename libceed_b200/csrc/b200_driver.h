// b200_driver.h -- CUDA driver API entry points resolved at run time (dlopen of libcuda.so.1), so that the shared
// library itself loads on machines without a GPU driver (CPU-only CI checks the exported C ABI) and fails loudly,
// with a clear message, on first use when no driver is present.
#pragma once
#include <cuda.h>

struct B200Driver {
  CUresult (*GetErrorString)(CUresult, const char **);
  CUresult (*ModuleLoadData)(CUmodule *, const void *);
  CUresult (*ModuleUnload)(CUmodule);
  CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, const char *);
  CUresult (*ModuleGetGlobal)(CUdeviceptr *, size_t *, CUmodule, const char *);
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **, void **);
  CUresult (*LaunchCooperativeKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **);
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int);
  CUresult (*FuncGetAttribute)(int *, CUfunction_attribute, CUfunction);
  CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int *, CUfunction, int, size_t);
};
const B200Driver *b200_driver();

#define cuGetErrorString b200_driver()->GetErrorString
#define cuModuleLoadData b200_driver()->ModuleLoadData
#define cuModuleUnload b200_driver()->ModuleUnload
#define cuModuleGetFunction b200_driver()->ModuleGetFunction
#undef cuModuleGetGlobal
#define cuModuleGetGlobal b200_driver()->ModuleGetGlobal
#define cuLaunchKernel b200_driver()->LaunchKernel
#define cuLaunchCooperativeKernel b200_driver()->LaunchCooperativeKernel
#define cuFuncSetAttribute b200_driver()->FuncSetAttribute
#define cuFuncGetAttribute b200_driver()->FuncGetAttribute
#define cuOccupancyMaxActiveBlocksPerMultiprocessor b200_driver()->OccupancyMaxActiveBlocksPerMultiprocessor
