// b200_runtime.cpp -- context, error handling, NVRTC JIT with an in-process module cache, kernel launch.
//
// Plays the role of the reference's CUDA common layer (device selection backends/cuda/ceed-cuda-common.c:19-35,
// NVRTC driver backends/cuda/ceed-cuda-compile.cpp:70-240, launch wrappers :551-598) for the b200 backend.
// Differences by design: always targets sm_100a, caches modules by source hash (the reference recompiles per
// object), and launches on a per-context stream.
#include <dlfcn.h>
#include <nvrtc.h>

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "b200_internal.h"

int b200_error(B200Ceed ceed, int code, const char *fmt, ...) {
  char    buf[4096];
  va_list args;
  va_start(args, fmt);
  vsnprintf(buf, sizeof(buf), fmt, args);
  va_end(args);
  if (ceed) ceed->last_error = buf;
  if (getenv("CEED_B200_DEBUG")) fprintf(stderr, "[ceed-b200] error %d: %s\n", code, buf);
  return code;
}

static std::string g_global_error;

extern "C" const char *ceedb200_last_error(B200Ceed ceed) { return ceed ? ceed->last_error.c_str() : g_global_error.c_str(); }
extern "C" const char *ceedb200_version(void) { return "ceed-b200 0.1 (sm_100a)"; }

// Directory holding the JIT device headers (csrc/jit), located relative to this shared library.
std::string b200_jit_dir() {
  const char *env = getenv("CEED_B200_JIT_DIR");
  if (env) return env;
  Dl_info info;
  if (dladdr((void *)&b200_jit_dir, &info) && info.dli_fname) {
    std::string path(info.dli_fname);
    size_t      pos = path.rfind('/');
    std::string dir = pos == std::string::npos ? "." : path.substr(0, pos);
    return dir + "/../csrc/jit";
  }
  return "libceed_b200/csrc/jit";
}

// Tuning table: one line per kernel signature, "<signature> epw group_warps cta_warps minb qf_mode qf_unroll stage [# comment]".
// Loaded from <library dir>/../tuned/sm_100a.tune (shipped, produced by scripts/gpu_autotune.py on a B200) and from
// CEED_B200_TUNE_FILE; later entries win.  CEED_B200_NO_TUNE_TABLE=1 ignores both (pure heuristics).
// "Q8|sc0|in:P7n1goN7|out:P7n1go" -> "Q8|sc0|in:P7n1go|out:P7n1go~": drops the N<comps> / W tokens of EVAL_NONE / EVAL_WEIGHT fields
std::string b200_reduced_signature(const std::string &shape_signature) {
  std::string out;
  for (size_t i = 0; i < shape_signature.size(); i++) {
    const char ch = shape_signature[i];
    if (ch == 'W') continue;
    if (ch == 'N') {
      while (i + 1 < shape_signature.size() && isdigit((unsigned char)shape_signature[i + 1])) i++;
      continue;
    }
    out += ch;
  }
  return out + "~";
}

static void b200_load_tune_file(B200Ceed ceed, const std::string &path) {
  FILE *f = fopen(path.c_str(), "r");
  if (!f) return;
  char line[1024], sig[512];
  while (fgets(line, sizeof(line), f)) {
    if (line[0] == '#' || line[0] == '\n') continue;
    B200Tuning t;
    if (sscanf(line, "%511s %d %d %d %d %d %d %d", sig, &t.epw, &t.group_warps, &t.cta_warps, &t.minb, &t.qf_mode, &t.qf_unroll, &t.stage) == 8) {
      ceed->tune_table[sig] = t;
      // shape-only key (signature without the trailing |<QFunction name>): first entry of a shape serves unknown QFunctions
      std::string s(sig);
      size_t      pos = s.rfind('|');
      if (pos != std::string::npos && !ceed->tune_table.count(s.substr(0, pos))) ceed->tune_table[s.substr(0, pos)] = t;
      // reduced key (additionally without the streamed-input component counts): e.g. the gallery Poisson operator with 6
      // quadrature-data components takes the shape tuned for the 7-component BP operator of the same P, Q and eval modes
      if (pos != std::string::npos) {
        const std::string red = b200_reduced_signature(s.substr(0, pos));
        if (!ceed->tune_table.count(red)) ceed->tune_table[red] = t;
      }
    }
  }
  fclose(f);
}

static void b200_init_tuning(B200Ceed ceed) {
  if (!getenv("CEED_B200_NO_TUNE_TABLE")) {
    b200_load_tune_file(ceed, b200_jit_dir() + "/../../tuned/sm_100a.tune");
    if (const char *path = getenv("CEED_B200_TUNE_FILE")) b200_load_tune_file(ceed, path);
  }
  if (const char *a = getenv("CEED_B200_AUTOTUNE")) ceed->autotune = atoi(a);
}

extern "C" int ceedb200_set_autotune(B200Ceed ceed, int level) {
  ceed->autotune = level;
  return B200_SUCCESS;
}

// Development aid: with CEED_B200_COMPILE_ONLY=1 a context can be created without a GPU so that the host-side
// setup analysis, the kernel generator and NVRTC can be exercised (register/SASS inspection, CPU-only CI).
// Device buffers are then host allocations and every kernel launch FAILS loudly -- nothing is ever computed on the CPU.
bool b200_compile_only() {
  static int flag = -1;
  if (flag < 0) flag = getenv("CEED_B200_COMPILE_ONLY") ? 1 : 0;
  return flag == 1;
}

int b200_dmalloc(B200Ceed ceed, void **p, size_t bytes) {
  if (bytes == 0) bytes = 8;
  if (b200_compile_only()) {
    *p = calloc(1, bytes);
    return B200_SUCCESS;
  }
  B200_CUDA(ceed, cudaMalloc(p, bytes));
  return B200_SUCCESS;
}
int b200_dfree(B200Ceed ceed, void *p) {
  if (!p) return B200_SUCCESS;
  if (b200_compile_only()) {
    free(p);
    return B200_SUCCESS;
  }
  B200_CUDA(ceed, cudaFree(p));
  return B200_SUCCESS;
}
int b200_h2d(B200Ceed ceed, void *d, const void *h, size_t bytes) {
  if (b200_compile_only()) {
    memcpy(d, h, bytes);
    return B200_SUCCESS;
  }
  B200_CUDA(ceed, cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ceed->stream));
  B200_CUDA(ceed, cudaStreamSynchronize(ceed->stream));
  return B200_SUCCESS;
}
int b200_d2h(B200Ceed ceed, void *h, const void *d, size_t bytes) {
  if (b200_compile_only()) {
    memcpy(h, d, bytes);
    return B200_SUCCESS;
  }
  B200_CUDA(ceed, cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ceed->stream));
  B200_CUDA(ceed, cudaStreamSynchronize(ceed->stream));
  return B200_SUCCESS;
}
int b200_d2d(B200Ceed ceed, void *dst, const void *src, size_t bytes) {
  if (b200_compile_only()) {
    memcpy(dst, src, bytes);
    return B200_SUCCESS;
  }
  B200_CUDA(ceed, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ceed->stream));
  return B200_SUCCESS;
}

extern "C" int ceedb200_init(int device_id, B200Ceed *ceed_out) {
  *ceed_out        = nullptr;
  if (b200_compile_only()) {
    B200Ceed ceed    = new B200Ceed_();
    ceed->device_id  = 0;
    ceed->num_sms    = 148;
    ceed->smem_optin = 232448;
    ceed->smem_sm    = 233472;
    ceed->cc_major   = 10;
    ceed->jit_roots.push_back(b200_jit_dir());
    b200_init_tuning(ceed);
    const char *mode = getenv("CEED_B200_SCATTER");
    if (mode && !strcmp(mode, "atomic")) ceed->scatter_mode = B200_SCATTER_ATOMIC;
    if (mode && !strcmp(mode, "evector")) ceed->scatter_mode = B200_SCATTER_EVECTOR;
    if (mode && !strcmp(mode, "ordered")) ceed->scatter_mode = B200_SCATTER_ORDERED;
    *ceed_out = ceed;
    return B200_SUCCESS;
  }
  int         count = 0;
  cudaError_t err   = cudaGetDeviceCount(&count);
  if (err != cudaSuccess || count == 0) {
    g_global_error = std::string("ceed-b200 requires a CUDA device: ") + cudaGetErrorString(err);
    return B200_ERROR_BACKEND;
  }
  if (device_id < 0) device_id = 0;
  if (device_id >= count) {
    g_global_error = "ceed-b200: device_id out of range";
    return B200_ERROR_BACKEND;
  }
  B200Ceed ceed   = new B200Ceed_();
  ceed->device_id = device_id;
  B200_CUDA(ceed, cudaSetDevice(device_id));
  B200_CUDA(ceed, cudaFree(0));  // force primary context creation
  cudaDeviceProp prop;
  B200_CUDA(ceed, cudaGetDeviceProperties(&prop, device_id));
  ceed->num_sms    = prop.multiProcessorCount;
  ceed->smem_optin = prop.sharedMemPerBlockOptin;
  ceed->smem_sm    = prop.sharedMemPerMultiprocessor;
  ceed->cc_major   = prop.major;
  ceed->cc_minor   = prop.minor;
  if (prop.major < 10 && !getenv("CEED_B200_ALLOW_ANY_ARCH")) {
    g_global_error = "ceed-b200 kernels are written for sm_100a (B200); found compute capability " + std::to_string(prop.major) + "." +
                     std::to_string(prop.minor);
    delete ceed;
    return B200_ERROR_BACKEND;
  }
  const char *mode = getenv("CEED_B200_SCATTER");
  if (mode) {
    if (!strcmp(mode, "atomic")) ceed->scatter_mode = B200_SCATTER_ATOMIC;
    else if (!strcmp(mode, "evector")) ceed->scatter_mode = B200_SCATTER_EVECTOR;
    else if (!strcmp(mode, "ordered")) ceed->scatter_mode = B200_SCATTER_ORDERED;
    else ceed->scatter_mode = B200_SCATTER_DETERMINISTIC;
  }
  ceed->jit_roots.push_back(b200_jit_dir());
  b200_init_tuning(ceed);
  *ceed_out = ceed;
  return B200_SUCCESS;
}

extern "C" int ceedb200_destroy(B200Ceed ceed) {
  if (!ceed) return B200_SUCCESS;
  if (!b200_compile_only()) cudaSetDevice(ceed->device_id);
  for (auto &kv : ceed->module_cache) {
    if (kv.second->module) cuModuleUnload(kv.second->module);
    delete kv.second;
  }
  b200_dfree(ceed, ceed->d_scratch);
  for (int i = 0; i < 3; i++) b200_dfree(ceed, ceed->d_basis_tmp[i]);
  for (auto ev : ceed->ev_stream) cudaEventDestroy(ev);  // streamed apply: events and copy streams
  if (ceed->s_h2d) cudaStreamDestroy(ceed->s_h2d);
  if (ceed->s_d2h) cudaStreamDestroy(ceed->s_d2h);
  delete ceed;
  return B200_SUCCESS;
}

extern "C" int ceedb200_add_jit_source_root(B200Ceed ceed, const char *path) {
  // user roots take precedence over the built-in shim headers (which stay last)
  ceed->jit_roots.insert(ceed->jit_roots.end() - 1, std::string(path));
  return B200_SUCCESS;
}
extern "C" int ceedb200_add_jit_define(B200Ceed ceed, const char *define) {
  ceed->jit_defines.push_back(define);
  return B200_SUCCESS;
}
extern "C" int ceedb200_set_stream(B200Ceed ceed, void *stream) {
  ceed->stream = (cudaStream_t)stream;
  return B200_SUCCESS;
}
extern "C" int ceedb200_get_stream(B200Ceed ceed, void **stream) {
  *stream = (void *)ceed->stream;
  return B200_SUCCESS;
}
extern "C" int ceedb200_synchronize(B200Ceed ceed) {
  if (b200_compile_only()) return B200_SUCCESS;
  B200_CUDA(ceed, cudaSetDevice(ceed->device_id));
  B200_CUDA(ceed, cudaStreamSynchronize(ceed->stream));
  return B200_SUCCESS;
}
extern "C" int ceedb200_set_scatter_mode(B200Ceed ceed, int mode) {
  B200_CHECK(mode >= 0 && mode <= 3, ceed, B200_ERROR_UNSUPPORTED, "unknown scatter mode %d", mode);
  ceed->scatter_mode = mode;
  return B200_SUCCESS;
}
extern "C" int64_t ceedb200_launch_count(B200Ceed ceed) { return ceed->launch_count; }

int b200_memset_async(B200Ceed ceed, void *d, size_t bytes) {
  if (b200_compile_only()) {
    memset(d, 0, bytes);
    return B200_SUCCESS;
  }
  B200_CUDA(ceed, cudaMemsetAsync(d, 0, bytes, ceed->stream));
  ceed->launch_count++;
  return B200_SUCCESS;
}

// ------------------------------------------------------------------------------------------------ JIT
int b200_jit_compile(B200Ceed ceed, const std::string &source, const std::vector<std::string> &defines, B200Module **module_out) {
  // Options follow the reference's contract for QFunction sources (backends/cuda/ceed-cuda-compile.cpp:70-134):
  //   -default-device, -Dint32_t=int, -DCEED_RUNNING_JIT_PASS=1, -I<jit roots>, -D<user defines>; arch is fixed to sm_100a.
  std::vector<std::string> opts;
  opts.push_back("-default-device");
  const char *arch = getenv("CEED_B200_ARCH");
  opts.push_back(std::string("-arch=") + (arch ? arch : "sm_100a"));
  opts.push_back("-std=c++17");
  opts.push_back("-Dint32_t=int");
  opts.push_back("-DCEED_RUNNING_JIT_PASS=1");
  opts.push_back("-lineinfo");
  if (getenv("CEED_B200_MAXRREG")) opts.push_back(std::string("-maxrregcount=") + getenv("CEED_B200_MAXRREG"));
  for (auto &r : ceed->jit_roots) opts.push_back("-I" + r);
  for (auto &d : ceed->jit_defines) opts.push_back("-D" + d);
  for (auto &d : defines) opts.push_back("-D" + d);

  std::string key = source;
  for (auto &o : opts) key += "\n//" + o;
  auto it = ceed->module_cache.find(key);
  if (it != ceed->module_cache.end()) {
    *module_out = it->second;
    return B200_SUCCESS;
  }

  if (getenv("CEED_B200_DEBUG_SOURCE")) fprintf(stderr, "---------- ceed-b200 JIT source (pre-compile) ----------\n%s\n", source.c_str());
  nvrtcProgram prog;
  nvrtcResult  res = nvrtcCreateProgram(&prog, source.c_str(), "ceed_b200_jit.cu", 0, nullptr, nullptr);
  B200_CHECK(res == NVRTC_SUCCESS, ceed, B200_ERROR_BACKEND, "nvrtcCreateProgram: %s", nvrtcGetErrorString(res));
  std::vector<const char *> copts;
  for (auto &o : opts) copts.push_back(o.c_str());
  res = nvrtcCompileProgram(prog, (int)copts.size(), copts.data());
  size_t log_size = 0;
  nvrtcGetProgramLogSize(prog, &log_size);
  std::string log(log_size, '\0');
  if (log_size > 1) nvrtcGetProgramLog(prog, &log[0]);
  if (getenv("CEED_B200_DEBUG")) {
    fprintf(stderr, "---------- ceed-b200 JIT source ----------\n%s\n---------- options:", source.c_str());
    for (auto &o : opts) fprintf(stderr, " %s", o.c_str());
    fprintf(stderr, "\n---------- log ----------\n%s\n", log.c_str());
  }
  if (res != NVRTC_SUCCESS) {
    nvrtcDestroyProgram(&prog);
    return b200_error(ceed, B200_ERROR_BACKEND, "NVRTC compile failed: %s\n%s", nvrtcGetErrorString(res), log.c_str());
  }
  size_t cubin_size = 0;
  res               = nvrtcGetCUBINSize(prog, &cubin_size);
  B200_CHECK(res == NVRTC_SUCCESS && cubin_size > 0, ceed, B200_ERROR_BACKEND, "nvrtcGetCUBINSize failed: %s", nvrtcGetErrorString(res));
  std::vector<char> cubin(cubin_size);
  nvrtcGetCUBIN(prog, cubin.data());
  nvrtcDestroyProgram(&prog);
  const char *dump = getenv("CEED_B200_DUMP_CUBIN");
  if (dump) {
    static int    counter = 0;
    std::string   base    = std::string(dump) + "/b200_jit_" + std::to_string(counter++);
    std::ofstream(base + ".cubin", std::ios::binary).write(cubin.data(), cubin.size());
    std::ofstream(base + ".cu") << source;
  }

  B200Module *mod = new B200Module();
  mod->source     = source;
  mod->log        = log;
  if (!b200_compile_only()) {
    B200_CUDA(ceed, cudaSetDevice(ceed->device_id));
    CUresult cres = cuModuleLoadData(&mod->module, cubin.data());
    if (cres != CUDA_SUCCESS) {
      const char *msg = nullptr;
      cuGetErrorString(cres, &msg);
      delete mod;
      return b200_error(ceed, B200_ERROR_BACKEND, "cuModuleLoadData failed: %s", msg ? msg : "?");
    }
  }
  ceed->module_cache[key] = mod;
  *module_out             = mod;
  return B200_SUCCESS;
}

int b200_jit_get_kernel(B200Ceed ceed, B200Module *module, const char *name, CUfunction *kernel) {
  if (b200_compile_only()) {
    *kernel = nullptr;
    return B200_SUCCESS;
  }
  B200_CU(ceed, cuModuleGetFunction(kernel, module->module, name));
  return B200_SUCCESS;
}

int b200_launch(B200Ceed ceed, CUfunction kernel, unsigned grid, unsigned block, unsigned smem_bytes, void **args, bool cooperative) {
  B200_CHECK(!b200_compile_only(), ceed, B200_ERROR_BACKEND, "ceed-b200: CEED_B200_COMPILE_ONLY is set; kernels cannot run without a GPU");
  // cooperative = all CTAs guaranteed co-resident (kernels whose element groups wait for each other)
  if (cooperative) B200_CU(ceed, cuLaunchCooperativeKernel(kernel, grid, 1, 1, block, 1, 1, smem_bytes, (CUstream)ceed->stream, args));
  else B200_CU(ceed, cuLaunchKernel(kernel, grid, 1, 1, block, 1, 1, smem_bytes, (CUstream)ceed->stream, args, nullptr));
  ceed->launch_count++;
  return B200_SUCCESS;
}
