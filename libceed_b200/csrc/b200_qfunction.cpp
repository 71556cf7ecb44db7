// b200_qfunction.cpp -- CeedQFunctionContext and CeedQFunction for the b200 backend.
//
// Context: host/device mirror of the user's opaque struct with the same validity rules as the vector
// (backends/cuda-ref/ceed-cuda-ref-qfunctioncontext.c:152-333).
// QFunction: holds the user's source path + function name and the field signature; the standalone apply JIT-compiles a
// wrapper kernel that evaluates one quadrature point per thread with Q = 1, fields laid out [size][Q]
// (same contract as backends/cuda-ref/ceed-cuda-ref-qfunction-load.cpp:22-113; layout doc/sphinx/source/libCEEDdev.md:113-118).
// The fused operator (b200_opgen.cpp) inlines the same user function into its own kernel instead.
#include <cstdlib>
#include <cstring>
#include <sstream>

#include "b200_internal.h"

// ------------------------------------------------------------------------------------------------ context
namespace {
inline void *ctx_host_ptr(B200QFContext c) { return c->h_borrowed ? c->h_borrowed : c->h_owned; }
inline void *ctx_device_ptr(B200QFContext c) { return c->d_borrowed ? c->d_borrowed : c->d_owned; }

int ctx_sync_to(B200QFContext c, int mem_type) {
  B200Ceed ceed = c->ceed;
  B200_CHECK(c->h_data || c->d_data, ceed, B200_ERROR_BACKEND, "No context data set");
  if (mem_type == B200_MEM_HOST && !c->h_data) {
    if (!c->h_owned && !c->h_borrowed) c->h_owned = calloc(1, c->size ? c->size : 1);
    B200_CALL(b200_d2h(ceed, ctx_host_ptr(c), c->d_data, c->size));
    c->h_data = ctx_host_ptr(c);
  } else if (mem_type == B200_MEM_DEVICE && !c->d_data) {
    if (!c->d_owned && !c->d_borrowed) B200_CALL(b200_dmalloc(ceed, &c->d_owned, c->size));
    B200_CALL(b200_h2d(ceed, ctx_device_ptr(c), c->h_data, c->size));
    c->d_data = ctx_device_ptr(c);
  }
  return B200_SUCCESS;
}
}  // namespace

extern "C" int ceedb200_qfcontext_create(B200Ceed ceed, B200QFContext *ctx) {
  B200QFContext c = new B200QFContext_();
  c->ceed         = ceed;
  *ctx            = c;
  return B200_SUCCESS;
}

extern "C" int ceedb200_qfcontext_destroy(B200QFContext c) {
  if (!c) return B200_SUCCESS;
  free(c->h_owned);
  b200_dfree(c->ceed, c->d_owned);
  delete c;
  return B200_SUCCESS;
}

extern "C" int ceedb200_qfcontext_set_data(B200QFContext c, int mem_type, int copy_mode, size_t size, void *data) {
  B200Ceed ceed = c->ceed;
  // a size change drops owned buffers of the old size
  if (size != c->size) {
    free(c->h_owned);
    c->h_owned = nullptr;
    B200_CALL(b200_dfree(ceed, c->d_owned));
    c->d_owned = nullptr;
  }
  c->size = size;
  if (mem_type == B200_MEM_HOST) {
    free(c->h_owned);
    c->h_owned    = nullptr;
    c->h_borrowed = nullptr;
    switch (copy_mode) {
      case B200_COPY_VALUES:
        c->h_owned = malloc(size ? size : 1);
        memcpy(c->h_owned, data, size);
        break;
      case B200_OWN_POINTER: c->h_owned = data; break;
      default: c->h_borrowed = data;
    }
    c->h_data = ctx_host_ptr(c);
    c->d_data = nullptr;
  } else {
    B200_CALL(b200_dfree(ceed, c->d_owned));
    c->d_owned    = nullptr;
    c->d_borrowed = nullptr;
    switch (copy_mode) {
      case B200_COPY_VALUES:
        B200_CALL(b200_dmalloc(ceed, &c->d_owned, size));
        B200_CALL(b200_d2d(ceed, c->d_owned, data, size));
        break;
      case B200_OWN_POINTER: c->d_owned = data; break;
      default: c->d_borrowed = data;
    }
    c->d_data = ctx_device_ptr(c);
    c->h_data = nullptr;
  }
  return B200_SUCCESS;
}

extern "C" int ceedb200_qfcontext_take_data(B200QFContext c, int mem_type, void **data) {
  B200_CALL(ctx_sync_to(c, mem_type));
  if (mem_type == B200_MEM_HOST) {
    *data         = c->h_borrowed;
    c->h_borrowed = nullptr;
    c->h_data     = nullptr;
  } else {
    *data         = c->d_borrowed;
    c->d_borrowed = nullptr;
    c->d_data     = nullptr;
  }
  return B200_SUCCESS;
}

extern "C" int ceedb200_qfcontext_get_data(B200QFContext c, int mem_type, void **data) {
  B200_CALL(ctx_sync_to(c, mem_type));
  if (mem_type == B200_MEM_HOST) {
    *data     = c->h_data;
    c->d_data = nullptr;
  } else {
    *data     = c->d_data;
    c->h_data = nullptr;
  }
  return B200_SUCCESS;
}

extern "C" int ceedb200_qfcontext_get_data_read(B200QFContext c, int mem_type, void **data) {
  B200_CALL(ctx_sync_to(c, mem_type));
  *data = mem_type == B200_MEM_HOST ? c->h_data : c->d_data;
  return B200_SUCCESS;
}

extern "C" int ceedb200_qfcontext_has_valid_data(B200QFContext c, int *has_valid) {
  *has_valid = (c->h_data || c->d_data) ? 1 : 0;
  return B200_SUCCESS;
}

extern "C" int ceedb200_qfcontext_has_borrowed_data_of_type(B200QFContext c, int mem_type, int *has_borrowed) {
  *has_borrowed = mem_type == B200_MEM_HOST ? (c->h_borrowed != nullptr) : (c->d_borrowed != nullptr);
  return B200_SUCCESS;
}

// ------------------------------------------------------------------------------------------------ qfunction
extern "C" int ceedb200_qfunction_create(B200Ceed ceed, const char *source_path, const char *kernel_name, B200QFunction *qf_out) {
  B200_CHECK(source_path && kernel_name, ceed, B200_ERROR_BACKEND, "No QFunction source or kernel name provided");
  B200QFunction qf = new B200QFunction_();
  qf->ceed         = ceed;
  qf->source_path  = source_path;
  qf->kernel_name  = kernel_name;
  *qf_out          = qf;
  return B200_SUCCESS;
}

extern "C" int ceedb200_qfunction_destroy(B200QFunction qf) {
  delete qf;
  return B200_SUCCESS;
}

extern "C" int ceedb200_qfunction_add_input(B200QFunction qf, const char *field_name, b200_int size, int eval_mode) {
  B200_CHECK((int)qf->inputs.size() < 16, qf->ceed, B200_ERROR_UNSUPPORTED, "too many input fields (max 16, backend.h:30)");
  qf->inputs.push_back({field_name, size, eval_mode});
  qf->kernel = nullptr;
  return B200_SUCCESS;
}
extern "C" int ceedb200_qfunction_add_output(B200QFunction qf, const char *field_name, b200_int size, int eval_mode) {
  B200_CHECK((int)qf->outputs.size() < 16, qf->ceed, B200_ERROR_UNSUPPORTED, "too many output fields (max 16, backend.h:30)");
  qf->outputs.push_back({field_name, size, eval_mode});
  qf->kernel = nullptr;
  return B200_SUCCESS;
}
extern "C" int ceedb200_qfunction_set_context_ptr(B200QFunction qf, void *d_ctx) {
  qf->ctx     = nullptr;
  qf->raw_ctx = d_ctx;
  return B200_SUCCESS;
}

extern "C" int ceedb200_qfunction_set_context(B200QFunction qf, B200QFContext ctx) {
  qf->raw_ctx = nullptr;
  qf->ctx = ctx;
  return B200_SUCCESS;
}

static int qfunction_build(B200QFunction qf) {
  if (qf->module) return B200_SUCCESS;
  std::ostringstream code;
  code << "#include <b200-jit.h>\n";
  code << "#include \"" << qf->source_path << "\"\n\n";
  code << "struct B200QFPointers { const CeedScalar *in[16]; CeedScalar *out[16]; };\n";
  code << "extern \"C\" __global__ void __launch_bounds__(256) b200_qfunction_apply(void *ctx, long long Q, B200QFPointers f) {\n";
  for (size_t i = 0; i < qf->inputs.size(); i++) code << "  CeedScalar in_" << i << "[" << qf->inputs[i].size << "];\n";
  for (size_t i = 0; i < qf->outputs.size(); i++) code << "  CeedScalar out_" << i << "[" << qf->outputs[i].size << "];\n";
  code << "  const CeedScalar *in[" << (qf->inputs.empty() ? 1 : qf->inputs.size()) << "];\n";
  code << "  CeedScalar *out[" << (qf->outputs.empty() ? 1 : qf->outputs.size()) << "];\n";
  for (size_t i = 0; i < qf->inputs.size(); i++) code << "  in[" << i << "] = in_" << i << ";\n";
  for (size_t i = 0; i < qf->outputs.size(); i++) code << "  out[" << i << "] = out_" << i << ";\n";
  code << "  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < Q; q += (long long)blockDim.x * gridDim.x) {\n";
  for (size_t i = 0; i < qf->inputs.size(); i++)
    code << "    for (int c = 0; c < " << qf->inputs[i].size << "; c++) in_" << i << "[c] = f.in[" << i << "][q + c * Q];\n";
  code << "    " << qf->kernel_name << "(ctx, 1, in, out);\n";
  for (size_t i = 0; i < qf->outputs.size(); i++)
    code << "    for (int c = 0; c < " << qf->outputs[i].size << "; c++) f.out[" << i << "][q + c * Q] = out_" << i << "[c];\n";
  code << "  }\n}\n";
  B200_CALL(b200_jit_compile(qf->ceed, code.str(), {}, &qf->module));
  B200_CALL(b200_jit_get_kernel(qf->ceed, qf->module, "b200_qfunction_apply", &qf->kernel));
  return B200_SUCCESS;
}

extern "C" int ceedb200_qfunction_apply_ptr(B200QFunction qf, b200_int Q, const b200_scalar *const *d_in, b200_scalar *const *d_out) {
  B200Ceed ceed = qf->ceed;
  B200_CALL(qfunction_build(qf));
  struct {
    const double *in[16];
    double       *out[16];
  } ptrs;
  memset(&ptrs, 0, sizeof(ptrs));
  for (size_t i = 0; i < qf->inputs.size(); i++) ptrs.in[i] = d_in[i];
  for (size_t i = 0; i < qf->outputs.size(); i++) ptrs.out[i] = d_out[i];
  void *d_ctx = qf->raw_ctx;
  if (qf->ctx) B200_CALL(ceedb200_qfcontext_get_data(qf->ctx, B200_MEM_DEVICE, &d_ctx));
  if (Q == 0) return B200_SUCCESS;
  long long q64    = Q;
  void     *args[] = {&d_ctx, &q64, &ptrs};
  int64_t   blocks = ((int64_t)Q + 255) / 256, cap = (int64_t)ceed->num_sms * 8;
  if (blocks > cap) blocks = cap;
  return b200_launch(ceed, qf->kernel, (unsigned)blocks, 256, 0, args);
}

extern "C" int ceedb200_qfunction_apply(B200QFunction qf, b200_int Q, const B200Vector *U, const B200Vector *V) {
  B200Ceed      ceed = qf->ceed;
  const double *in[16]  = {nullptr};
  double       *out[16] = {nullptr};
  for (size_t i = 0; i < qf->inputs.size(); i++) {
    B200_CHECK(U[i]->length >= (int64_t)Q * qf->inputs[i].size, ceed, B200_ERROR_DIMENSION, "QFunction input %zu too short", i);
    B200_CALL(b200_vector_device_read(U[i], &in[i]));
  }
  for (size_t i = 0; i < qf->outputs.size(); i++) {
    B200_CHECK(V[i]->length >= (int64_t)Q * qf->outputs[i].size, ceed, B200_ERROR_DIMENSION, "QFunction output %zu too short", i);
    B200_CALL(b200_vector_device_write(V[i], &out[i], V[i]->length == (int64_t)Q * qf->outputs[i].size));
  }
  return ceedb200_qfunction_apply_ptr(qf, Q, in, out);
}
