// b200_vector.cu -- CeedVector for the b200 backend: host/device mirrors with lazy synchronisation and BLAS-1 kernels.
//
// Semantics follow the reference CUDA vector (backends/cuda-ref/ceed-cuda-ref-vector.c): a vector has up to two
// storage sides (host, device), each either owned or borrowed; `h_array`/`d_array` are non-NULL only while that
// side holds valid data (:116-130); read access syncs lazily (:96-114), write access invalidates the other side.
// Kernels are grid-stride and vectorised (double2) -- HBM-bound streaming ops sized in multiples of the SM count.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include <algorithm>

#include "b200_internal.h"

namespace {

constexpr int kThreads = 256;

inline unsigned grid_for(B200Ceed ceed, int64_t n, int per_thread = 2) {
  int64_t blocks = (n + (int64_t)kThreads * per_thread - 1) / ((int64_t)kThreads * per_thread);
  int64_t cap    = (int64_t)ceed->num_sms * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

__global__ void k_set_value(double *__restrict__ v, int64_t n, double val) {
  int64_t i      = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // vectorised body when aligned
  if ((reinterpret_cast<uintptr_t>(v) & 15) == 0) {
    double2 *v2 = reinterpret_cast<double2 *>(v);
    int64_t  n2 = n >> 1;
    for (int64_t j = i; j < n2; j += stride) v2[j] = make_double2(val, val);
    if (i == 0 && (n & 1)) v[n - 1] = val;
  } else {
    for (int64_t j = i; j < n; j += stride) v[j] = val;
  }
}
__global__ void k_set_value_strided(double *__restrict__ v, int64_t start, int64_t stop, int64_t step, double val) {
  int64_t i      = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t count  = (stop - start + step - 1) / step;
  for (int64_t j = i; j < count; j += stride) v[start + j * step] = val;
}
__global__ void k_copy_strided(const double *__restrict__ src, double *__restrict__ dst, int64_t start, int64_t stop, int64_t step) {
  int64_t i      = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t count  = (stop - start + step - 1) / step;
  for (int64_t j = i; j < count; j += stride) dst[start + j * step] = src[start + j * step];
}
__global__ void k_scale(double *__restrict__ x, int64_t n, double alpha) {
  int64_t i      = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = i; j < n; j += stride) x[j] *= alpha;
}
__global__ void k_reciprocal(double *__restrict__ x, int64_t n) {
  int64_t i      = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = i; j < n; j += stride) {
    double v = x[j];
    if (fabs(v) > 1e-16) x[j] = 1.0 / v;  // CEED_EPSILON guard as cuda-ref-vector.cu:84-100
  }
}
__global__ void k_filter(double *__restrict__ x, int64_t n, double eps) {
  int64_t i      = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = i; j < n; j += stride) {
    if (fabs(x[j]) <= eps) x[j] = 0.0;  // inclusive threshold, cuda-ref-vector.cu:130-136
  }
}
__global__ void k_axpy(double *__restrict__ y, double alpha, const double *__restrict__ x, int64_t n) {
  int64_t i      = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = i; j < n; j += stride) y[j] += alpha * x[j];
}
__global__ void k_axpby(double *__restrict__ y, double alpha, double beta, const double *__restrict__ x, int64_t n) {
  int64_t i      = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = i; j < n; j += stride) y[j] = alpha * x[j] + beta * y[j];
}
__global__ void k_pointwise_mult(double *w, const double *x, const double *y, int64_t n) {
  int64_t i      = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = i; j < n; j += stride) w[j] = x[j] * y[j];
}

// norm reduction: mode 0 = sum|x|, 1 = sum x^2, 2 = max|x|.  Deterministic: fixed grid, fixed tree order.
__global__ void k_norm_partial(const double *__restrict__ x, int64_t n, int mode, double *__restrict__ partial) {
  __shared__ double s[kThreads];
  int64_t           i      = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t           stride = (int64_t)gridDim.x * blockDim.x;
  double            acc    = 0.0;
  for (int64_t j = i; j < n; j += stride) {
    double v = fabs(x[j]);
    if (mode == 0) acc += v;
    else if (mode == 1) acc += v * v;
    else acc = fmax(acc, v);
  }
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int w = kThreads / 2; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) s[threadIdx.x] = (mode == 2) ? fmax(s[threadIdx.x], s[threadIdx.x + w]) : s[threadIdx.x] + s[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = s[0];
}

int host_alloc(B200Vector vec) {
  if (!vec->h_owned && !vec->h_borrowed) {
    vec->h_owned = (double *)calloc(vec->length > 0 ? vec->length : 1, sizeof(double));
    B200_CHECK(vec->h_owned, vec->ceed, B200_ERROR_MAJOR, "host allocation of %lld doubles failed", (long long)vec->length);
  }
  return B200_SUCCESS;
}
int device_alloc(B200Vector vec) {
  if (!vec->d_owned && !vec->d_borrowed) B200_CALL(b200_dmalloc(vec->ceed, (void **)&vec->d_owned, vec->length * sizeof(double)));
  return B200_SUCCESS;
}
inline double *host_ptr(B200Vector vec) { return vec->h_borrowed ? vec->h_borrowed : vec->h_owned; }
inline double *device_ptr(B200Vector vec) { return vec->d_borrowed ? vec->d_borrowed : vec->d_owned; }

int sync_to(B200Vector vec, int mem_type) {
  B200Ceed ceed = vec->ceed;
  B200_CHECK(vec->h_array || vec->d_array, ceed, B200_ERROR_BACKEND, "Invalid data access: CeedVector has no valid data to sync");
  size_t bytes = vec->length * sizeof(double);
  if (mem_type == B200_MEM_HOST && !vec->h_array) {
    B200_CALL(host_alloc(vec));
    B200_CALL(b200_d2h(ceed, host_ptr(vec), vec->d_array, bytes));
    vec->h_array = host_ptr(vec);
  } else if (mem_type == B200_MEM_DEVICE && !vec->d_array) {
    B200_CALL(device_alloc(vec));
    B200_CALL(b200_h2d(ceed, device_ptr(vec), vec->h_array, bytes));
    vec->d_array = device_ptr(vec);
  }
  return B200_SUCCESS;
}

}  // namespace

#define LAUNCH(ceed, kernel, n, ...)                                                            \
  do {                                                                                          \
    B200_CHECK(!b200_compile_only(), ceed, B200_ERROR_BACKEND, "CEED_B200_COMPILE_ONLY is set; kernels cannot run"); \
    kernel<<<grid_for(ceed, n), kThreads, 0, (ceed)->stream>>>(__VA_ARGS__);                    \
    (ceed)->launch_count++;                                                                     \
    B200_CUDA(ceed, cudaGetLastError());                                                        \
  } while (0)

// ------------------------------------------------------------------------------------------------ internal access
int b200_vector_device_read(B200Vector vec, const double **d) {
  B200_CALL(sync_to(vec, B200_MEM_DEVICE));
  *d = vec->d_array;
  return B200_SUCCESS;
}
int b200_vector_streamed_input(B200Vector vec, const double **h, double **d) {
  B200_CHECK(vec->h_array && !vec->d_array, vec->ceed, B200_ERROR_BACKEND, "streamed input needs a vector that is valid on the host only");
  B200_CALL(device_alloc(vec));
  *h           = vec->h_array;
  *d           = device_ptr(vec);
  vec->d_array = device_ptr(vec);  // filled chunk by chunk on the copy stream; kernels wait for the chunk events
  return B200_SUCCESS;
}
int b200_vector_streamed_output(B200Vector vec, double **h) {
  B200_CHECK(host_ptr(vec), vec->ceed, B200_ERROR_BACKEND, "streamed output needs a vector with a host array");
  *h           = host_ptr(vec);
  vec->h_array = host_ptr(vec);
  return B200_SUCCESS;
}
int b200_vector_device_write(B200Vector vec, double **d, bool discard) {
  if (discard || (!vec->h_array && !vec->d_array)) {  // nothing valid to preserve: (allocate and) hand out the device side
    B200_CALL(device_alloc(vec));
    vec->d_array = device_ptr(vec);
  } else {
    B200_CALL(sync_to(vec, B200_MEM_DEVICE));
  }
  vec->h_array = nullptr;
  *d           = vec->d_array;
  return B200_SUCCESS;
}

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" int ceedb200_vector_create(B200Ceed ceed, b200_size length, B200Vector *vec) {
  B200_CHECK(length >= 0, ceed, B200_ERROR_DIMENSION, "negative vector length");
  B200Vector v = new B200Vector_();
  v->ceed      = ceed;
  v->length    = length;
  *vec         = v;
  return B200_SUCCESS;
}

extern "C" int ceedb200_vector_destroy(B200Vector vec) {
  if (!vec) return B200_SUCCESS;
  free(vec->h_owned);
  b200_dfree(vec->ceed, vec->d_owned);
  delete vec;
  return B200_SUCCESS;
}

extern "C" int ceedb200_vector_length(B200Vector vec, b200_size *length) {
  *length = vec->length;
  return B200_SUCCESS;
}

extern "C" int ceedb200_vector_has_valid_array(B200Vector vec, int *has_valid) {
  *has_valid = (vec->h_array || vec->d_array) ? 1 : 0;
  return B200_SUCCESS;
}

extern "C" int ceedb200_vector_has_borrowed_array_of_type(B200Vector vec, int mem_type, int *has_borrowed) {
  *has_borrowed = mem_type == B200_MEM_HOST ? (vec->h_borrowed != nullptr) : (vec->d_borrowed != nullptr);
  return B200_SUCCESS;
}

// CeedVectorSetArray_Cuda (ceed-cuda-ref-vector.c:177-228): the given side becomes the only valid one.
extern "C" int ceedb200_vector_valid_sides(B200Vector vec, int *host_valid, int *device_valid) {
  if (host_valid) *host_valid = vec->h_array != nullptr;
  if (device_valid) *device_valid = vec->d_array != nullptr;
  return B200_SUCCESS;
}

extern "C" int ceedb200_vector_set_array(B200Vector vec, int mem_type, int copy_mode, b200_scalar *array) {
  B200Ceed ceed  = vec->ceed;
  size_t   bytes = vec->length * sizeof(double);
  if (mem_type == B200_MEM_HOST) {
    switch (copy_mode) {
      case B200_COPY_VALUES:
        B200_CALL(host_alloc(vec));
        if (array) memcpy(host_ptr(vec), array, bytes);
        break;
      case B200_OWN_POINTER:
        free(vec->h_owned);
        vec->h_owned    = array;
        vec->h_borrowed = nullptr;
        break;
      case B200_USE_POINTER:
        free(vec->h_owned);
        vec->h_owned    = nullptr;
        vec->h_borrowed = array;
        break;
      default:
        return b200_error(ceed, B200_ERROR_UNSUPPORTED, "unknown copy mode");
    }
    vec->h_array = host_ptr(vec);
    vec->d_array = nullptr;
  } else {
    switch (copy_mode) {
      case B200_COPY_VALUES:
        B200_CALL(device_alloc(vec));
        if (array) B200_CALL(b200_d2d(ceed, device_ptr(vec), array, bytes));
        break;
      case B200_OWN_POINTER:
        B200_CALL(b200_dfree(ceed, vec->d_owned));
        vec->d_owned    = array;
        vec->d_borrowed = nullptr;
        break;
      case B200_USE_POINTER:
        B200_CALL(b200_dfree(ceed, vec->d_owned));
        vec->d_owned    = nullptr;
        vec->d_borrowed = array;
        break;
      default:
        return b200_error(ceed, B200_ERROR_UNSUPPORTED, "unknown copy mode");
    }
    vec->d_array = device_ptr(vec);
    vec->h_array = nullptr;
  }
  return B200_SUCCESS;
}

// CeedVectorTakeArray_Cuda (ceed-cuda-ref-vector.c:307-342): hand back the borrowed pointer of that side, synced.
extern "C" int ceedb200_vector_take_array(B200Vector vec, int mem_type, b200_scalar **array) {
  B200_CALL(sync_to(vec, mem_type));
  if (mem_type == B200_MEM_HOST) {
    if (array) *array = vec->h_borrowed;
    vec->h_borrowed = nullptr;
    vec->h_array    = nullptr;
  } else {
    if (array) *array = vec->d_borrowed;
    vec->d_borrowed = nullptr;
    vec->d_array    = nullptr;
  }
  return B200_SUCCESS;
}

extern "C" int ceedb200_vector_set_value(B200Vector vec, b200_scalar value) {
  B200Ceed ceed = vec->ceed;
  // Same side-selection rule as ceed-cuda-ref-vector.c:307-339: write into the currently valid side (device first);
  // with nothing valid prefer borrowed-device, borrowed-host, owned-device, owned-host, else allocate on the device.
  if (!vec->d_array && !vec->h_array) {
    if (vec->d_borrowed) vec->d_array = vec->d_borrowed;
    else if (vec->h_borrowed) vec->h_array = vec->h_borrowed;
    else if (vec->d_owned) vec->d_array = vec->d_owned;
    else if (vec->h_owned) vec->h_array = vec->h_owned;
    else {
      B200_CALL(device_alloc(vec));
      vec->d_array = device_ptr(vec);
    }
  }
  if (vec->d_array) {
    if (vec->length > 0) {
      if (value == 0.0) B200_CALL(b200_memset_async(ceed, vec->d_array, vec->length * sizeof(double)));
      else LAUNCH(ceed, k_set_value, vec->length, vec->d_array, vec->length, value);
    }
    vec->h_array = nullptr;
  } else {
    for (int64_t i = 0; i < vec->length; i++) vec->h_array[i] = value;
    vec->d_array = nullptr;
  }
  return B200_SUCCESS;
}

extern "C" int ceedb200_vector_set_value_strided(B200Vector vec, b200_size start, b200_size stop, b200_size step, b200_scalar value) {
  B200Ceed ceed = vec->ceed;
  if (stop < 0) stop = vec->length;
  B200_CHECK(step > 0 && start >= 0 && stop <= vec->length, ceed, B200_ERROR_DIMENSION, "invalid strided range");
  if (vec->d_array) {
    vec->h_array = nullptr;
    if (stop > start) LAUNCH(ceed, k_set_value_strided, (stop - start) / step + 1, vec->d_array, start, stop, step, value);
  } else if (vec->h_array) {
    for (int64_t i = start; i < stop; i += step) vec->h_array[i] = value;
  } else {
    return b200_error(ceed, B200_ERROR_BACKEND, "CeedVector must have valid data set");
  }
  return B200_SUCCESS;
}

extern "C" int ceedb200_vector_sync_array(B200Vector vec, int mem_type) { return sync_to(vec, mem_type); }

extern "C" int ceedb200_vector_get_array(B200Vector vec, int mem_type, b200_scalar **array) {
  B200_CALL(sync_to(vec, mem_type));
  if (mem_type == B200_MEM_HOST) {
    *array       = vec->h_array;
    vec->d_array = nullptr;
  } else {
    *array       = vec->d_array;
    vec->h_array = nullptr;
  }
  return B200_SUCCESS;
}

extern "C" int ceedb200_vector_get_array_read(B200Vector vec, int mem_type, const b200_scalar **array) {
  B200_CALL(sync_to(vec, mem_type));
  *array = mem_type == B200_MEM_HOST ? vec->h_array : vec->d_array;
  return B200_SUCCESS;
}

extern "C" int ceedb200_vector_get_array_write(B200Vector vec, int mem_type, b200_scalar **array) {
  if (mem_type == B200_MEM_HOST) {
    B200_CALL(host_alloc(vec));
    vec->h_array = host_ptr(vec);
    vec->d_array = nullptr;
    *array       = vec->h_array;
  } else {
    B200_CALL(device_alloc(vec));
    vec->d_array = device_ptr(vec);
    vec->h_array = nullptr;
    *array       = vec->d_array;
  }
  return B200_SUCCESS;
}

extern "C" int ceedb200_vector_copy_strided(B200Vector src, b200_size start, b200_size stop, b200_size step, B200Vector dst) {
  B200Ceed ceed = src->ceed;
  if (stop < 0) stop = src->length < dst->length ? src->length : dst->length;
  B200_CHECK(step > 0 && start >= 0, ceed, B200_ERROR_DIMENSION, "invalid strided range");
  const double *d_src;
  double       *d_dst;
  B200_CALL(b200_vector_device_read(src, &d_src));
  B200_CALL(b200_vector_device_write(dst, &d_dst, false));
  if (stop > start) LAUNCH(ceed, k_copy_strided, (stop - start) / step + 1, d_src, d_dst, start, stop, step);
  return B200_SUCCESS;
}

extern "C" int ceedb200_vector_norm(B200Vector vec, int norm_type, b200_scalar *norm) {
  B200Ceed      ceed = vec->ceed;
  const double *d;
  B200_CALL(b200_vector_device_read(vec, &d));
  *norm = 0.0;
  if (vec->length == 0) return B200_SUCCESS;
  B200_CHECK(!b200_compile_only(), ceed, B200_ERROR_BACKEND, "CEED_B200_COMPILE_ONLY is set; kernels cannot run");
  unsigned grid = grid_for(ceed, vec->length, 8);
  if (ceed->scratch_len < grid) {
    const size_t want = std::max<size_t>((size_t)ceed->num_sms * 16, grid);
    B200_CALL(b200_dfree(ceed, ceed->d_scratch));
    ceed->d_scratch = nullptr, ceed->scratch_len = 0;
    B200_CALL(b200_dmalloc(ceed, (void **)&ceed->d_scratch, want * sizeof(double)));
    ceed->scratch_len = want;
  }
  int mode = norm_type == B200_NORM_1 ? 0 : (norm_type == B200_NORM_2 ? 1 : 2);
  k_norm_partial<<<grid, kThreads, 0, ceed->stream>>>(d, vec->length, mode, ceed->d_scratch);
  ceed->launch_count++;
  B200_CUDA(ceed, cudaGetLastError());
  std::vector<double> partial(grid);
  B200_CALL(b200_d2h(ceed, partial.data(), ceed->d_scratch, grid * sizeof(double)));
  double acc = 0.0;
  for (unsigned i = 0; i < grid; i++) acc = (mode == 2) ? fmax(acc, partial[i]) : acc + partial[i];
  *norm = (mode == 1) ? sqrt(acc) : acc;
  return B200_SUCCESS;
}

extern "C" int ceedb200_vector_scale(B200Vector x, b200_scalar alpha) {
  double *d;
  B200_CALL(b200_vector_device_write(x, &d, false));
  if (x->length > 0) LAUNCH(x->ceed, k_scale, x->length, d, x->length, alpha);
  return B200_SUCCESS;
}
extern "C" int ceedb200_vector_reciprocal(B200Vector x) {
  double *d;
  B200_CALL(b200_vector_device_write(x, &d, false));
  if (x->length > 0) LAUNCH(x->ceed, k_reciprocal, x->length, d, x->length);
  return B200_SUCCESS;
}
extern "C" int ceedb200_vector_filter(B200Vector x, b200_scalar epsilon) {
  double *d;
  B200_CALL(b200_vector_device_write(x, &d, false));
  if (x->length > 0) LAUNCH(x->ceed, k_filter, x->length, d, x->length, epsilon);
  return B200_SUCCESS;
}
extern "C" int ceedb200_vector_axpy(B200Vector y, b200_scalar alpha, B200Vector x) {
  B200_CHECK(x->length == y->length, y->ceed, B200_ERROR_DIMENSION, "AXPY length mismatch");
  const double *dx;
  double       *dy;
  B200_CALL(b200_vector_device_read(x, &dx));
  B200_CALL(b200_vector_device_write(y, &dy, false));
  if (y->length > 0) LAUNCH(y->ceed, k_axpy, y->length, dy, alpha, dx, y->length);
  return B200_SUCCESS;
}
extern "C" int ceedb200_vector_axpby(B200Vector y, b200_scalar alpha, b200_scalar beta, B200Vector x) {
  B200_CHECK(x->length == y->length, y->ceed, B200_ERROR_DIMENSION, "AXPBY length mismatch");
  const double *dx;
  double       *dy;
  B200_CALL(b200_vector_device_read(x, &dx));
  B200_CALL(b200_vector_device_write(y, &dy, false));
  if (y->length > 0) LAUNCH(y->ceed, k_axpby, y->length, dy, alpha, beta, dx, y->length);
  return B200_SUCCESS;
}
extern "C" int ceedb200_vector_pointwise_mult(B200Vector w, B200Vector x, B200Vector y) {
  B200_CHECK(x->length == w->length && y->length == w->length, w->ceed, B200_ERROR_DIMENSION, "PointwiseMult length mismatch");
  const double *dx, *dy;
  double       *dw;
  B200_CALL(b200_vector_device_read(x, &dx));
  B200_CALL(b200_vector_device_read(y, &dy));
  // w may alias x and/or y: keep its data if so
  B200_CALL(b200_vector_device_write(w, &dw, !(w == x || w == y)));
  if (w->length > 0) LAUNCH(w->ceed, k_pointwise_mult, w->length, dw, dx, dy, w->length);
  return B200_SUCCESS;
}

// ------------------------------------------------------------------------------------------------ interface exchange
// Multi-GPU element partition (behaviour reference examples/petsc/bpsraw.c:240-262, the PETSc
// VecScatter ADD_VALUES the reference examples delegate to): every rank owns a local L-vector with copies of the interface
// nodes.  After the local operator apply, k_iface_pack gathers this rank's partial sums of all interface entries into one
// send buffer (one contiguous segment per neighbour), the segments travel with NCCL send/recv, and k_iface_unpack_sum
// rebuilds every interface entry as the sum of all partial values in ASCENDING RANK ORDER (own value at its place), so all
// copies of a node carry identical bits on every rank.
namespace {
__global__ void k_iface_pack(const double *__restrict__ v, const long long *__restrict__ idx, long long n, double *__restrict__ send) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) send[i] = v[idx[i]];
}
__global__ void k_iface_unpack_sum(double *__restrict__ v, long long n, const long long *__restrict__ node, const int *__restrict__ ptr,
                                   const int *__restrict__ src, const double *__restrict__ recv) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long l = node[i];
    const int       b = ptr[i], e = ptr[i + 1];
    double          acc = src[b] < 0 ? v[l] : recv[src[b]];
    for (int k = b + 1; k < e; k++) acc += src[k] < 0 ? v[l] : recv[src[k]];
    v[l] = acc;
  }
}
}  // namespace

extern "C" int ceedb200_iface_pack(B200Ceed ceed, const double *d_v, const long long *d_idx, long long n, double *d_send) {
  if (n > 0) LAUNCH(ceed, k_iface_pack, n, d_v, d_idx, n, d_send);
  return B200_SUCCESS;
}

extern "C" int ceedb200_iface_unpack_sum(B200Ceed ceed, double *d_v, long long n, const long long *d_node, const int *d_ptr, const int *d_src,
                                         const double *d_recv) {
  if (n > 0) LAUNCH(ceed, k_iface_unpack_sum, n, d_v, n, d_node, d_ptr, d_src, d_recv);
  return B200_SUCCESS;
}

// ---- the same exchange over NVLink peer memory, without NCCL on the data path ---------------------------------------
// Every rank exposes ONE device allocation to its neighbours (CUDA IPC): a double-buffered receive area [2][total] followed by
// one arrival flag per neighbour.  k_iface_put gathers the interface values and stores them STRAIGHT INTO THE NEIGHBOURS'
// receive areas (st.global on peer-mapped addresses travel over NVLink 5 / NVSwitch), then the last CTA to finish publishes the
// step number in each neighbour's flag (fence + st.release.sys).  k_iface_wait_unpack_sum on the receiving GPU spins on its
// flags (ld.acquire.sys) and then forms the rank-ordered sums as above.  The step counter lives on the device and the buffer
// half alternates with its parity, so the pair needs no host-side arguments that change from step to step (CUDA-graph friendly)
// and a neighbour that is one step ahead writes the other half.  The put needs no SM resources to speak of (a few CTAs, no
// shared memory): launched at high priority right after the boundary elements it completes while the interior-element kernel runs.
namespace {
constexpr int kPutBlocks = 16;

__global__ void k_iface_put(const double *__restrict__ v, const long long *__restrict__ idx, const int *__restrict__ nb_of, const long long *__restrict__ seg,
                            long long n, int num_nb, double *const *__restrict__ peer_recv, const long long *__restrict__ peer_half,
                            long long *const *__restrict__ peer_flag, long long *__restrict__ ctr) {
  const long long step = ctr[0] + 1;  // ctr[0] = last completed put; every CTA reads it before the last one to finish updates it
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int k = nb_of[i];
    peer_recv[k][(step & 1) * peer_half[k] + (i - seg[k])] = v[idx[i]];
  }
  __threadfence_system();  // this thread's peer stores are ordered before what follows
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long done = atomicAdd((unsigned long long *)(ctr + 1), 1ULL);
    if (done == gridDim.x - 1) {  // all CTAs have stored and fenced
      __threadfence_system();
      for (int k = 0; k < num_nb; k++) asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(peer_flag[k]), "l"(step) : "memory");
      ctr[1] = 0;
      ctr[0] = step;
    }
  }
}

__global__ void k_iface_wait_unpack_sum(double *__restrict__ v, long long n, const long long *__restrict__ node, const int *__restrict__ ptr,
                                        const int *__restrict__ src, const double *__restrict__ recv, long long half, const long long *__restrict__ flags,
                                        int num_nb, const long long *__restrict__ ctr) {
  const long long step = ctr[0];  // the put of this step ran before this kernel (same GPU, stream/event ordered)
  if (threadIdx.x < num_nb) {
    long long seen;
    do {
      asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(seen) : "l"(flags + threadIdx.x) : "memory");
    } while (seen < step);
  }
  __syncthreads();
  const double *r = recv + (step & 1) * half;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long l = node[i];
    const int       b = ptr[i], e = ptr[i + 1];
    double          acc = src[b] < 0 ? v[l] : __ldcg(r + src[b]);
    for (int k = b + 1; k < e; k++) acc += src[k] < 0 ? v[l] : __ldcg(r + src[k]);
    v[l] = acc;
  }
}
}  // namespace

extern "C" int ceedb200_iface_put(B200Ceed ceed, void *stream, const double *d_v, const long long *d_idx, const int *d_nb_of, const long long *d_seg,
                                  long long n, int num_nb, double *const *d_peer_recv, const long long *d_peer_half, long long *const *d_peer_flag,
                                  long long *d_ctr) {
  B200_CHECK(!b200_compile_only(), ceed, B200_ERROR_BACKEND, "CEED_B200_COMPILE_ONLY is set; kernels cannot run");
  B200_CHECK(num_nb <= kThreads, ceed, B200_ERROR_UNSUPPORTED, "more than %d neighbour ranks", kThreads);
  if (num_nb == 0) return B200_SUCCESS;
  int blocks = (int)std::min<long long>(kPutBlocks, (n + kThreads - 1) / kThreads);
  if (blocks < 1) blocks = 1;
  k_iface_put<<<blocks, kThreads, 0, stream ? (cudaStream_t)stream : ceed->stream>>>(d_v, d_idx, d_nb_of, d_seg, n, num_nb, d_peer_recv, d_peer_half,
                                                                                     d_peer_flag, d_ctr);
  ceed->launch_count++;
  B200_CUDA(ceed, cudaGetLastError());
  return B200_SUCCESS;
}

extern "C" int ceedb200_iface_wait_unpack_sum(B200Ceed ceed, double *d_v, long long n, const long long *d_node, const int *d_ptr, const int *d_src,
                                              const double *d_recv, long long half, const long long *d_flags, int num_nb, const long long *d_ctr) {
  B200_CHECK(!b200_compile_only(), ceed, B200_ERROR_BACKEND, "CEED_B200_COMPILE_ONLY is set; kernels cannot run");
  if (num_nb == 0) return B200_SUCCESS;
  // few CTAs: every one of them waits for the flags itself (no grid-wide dependency)
  int blocks = (int)std::min<long long>(4 * ceed->num_sms, (n + kThreads - 1) / kThreads);
  if (blocks < 1) blocks = 1;
  k_iface_wait_unpack_sum<<<blocks, kThreads, 0, ceed->stream>>>(d_v, n, d_node, d_ptr, d_src, d_recv, half, d_flags, num_nb, d_ctr);
  ceed->launch_count++;
  B200_CUDA(ceed, cudaGetLastError());
  return B200_SUCCESS;
}

// CUDA IPC plumbing for the peer-memory exchange: one process per GPU, handles travel through the caller's bootstrap
// channel (torch.distributed all_gather_object in libceed_b200/parallel.py).
extern "C" int ceedb200_ipc_alloc(B200Ceed ceed, size_t bytes, void **d_ptr, unsigned char *handle64) {
  B200_CHECK(!b200_compile_only(), ceed, B200_ERROR_BACKEND, "CEED_B200_COMPILE_ONLY is set; no device");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  B200_CUDA(ceed, cudaSetDevice(ceed->device_id));
  B200_CUDA(ceed, cudaMalloc(d_ptr, bytes ? bytes : 8));
  B200_CUDA(ceed, cudaMemset(*d_ptr, 0, bytes ? bytes : 8));
  cudaIpcMemHandle_t h;
  B200_CUDA(ceed, cudaIpcGetMemHandle(&h, *d_ptr));
  memcpy(handle64, &h, 64);
  return B200_SUCCESS;
}
extern "C" int ceedb200_ipc_open(B200Ceed ceed, const unsigned char *handle64, void **d_ptr) {
  B200_CHECK(!b200_compile_only(), ceed, B200_ERROR_BACKEND, "CEED_B200_COMPILE_ONLY is set; no device");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  B200_CUDA(ceed, cudaSetDevice(ceed->device_id));
  B200_CUDA(ceed, cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return B200_SUCCESS;
}
extern "C" int ceedb200_ipc_close(B200Ceed ceed, void *d_ptr) {
  if (d_ptr) B200_CUDA(ceed, cudaIpcCloseMemHandle(d_ptr));
  return B200_SUCCESS;
}
extern "C" int ceedb200_ipc_free(B200Ceed ceed, void *d_ptr) {
  if (d_ptr) B200_CUDA(ceed, cudaFree(d_ptr));
  return B200_SUCCESS;
}

// ------------------------------------------------------------------------------------------------ device-resident CG pieces
// the CG loop around the operator (what the reference's published figure of merit measures,
// examples/petsc/bps.c:218-288, there with PETSc's KSPCG + VecDot/VecAXPY).  All scalars stay on the device, so an iteration
// needs no host synchronisation and can be replayed from a CUDA graph; reductions are two-pass with a fixed grid and a fixed
// tree order (deterministic).  `w` (optional) weights every entry of a dot product: 1 for entries this rank owns, 0 for its
// copies of interface nodes, so that a multi-GPU dot (local dot + all-reduce) counts every DoF once.
namespace {
constexpr int kCgBlocks = 592;  // 4 x 148: fixed, independent of n -> reproducible partial sums

__device__ __forceinline__ void block_sum_to(double acc, double *__restrict__ out) {
  __shared__ double s[kThreads];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int w = kThreads / 2; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) s[threadIdx.x] += s[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = s[0];
}
__global__ void k_cg_dot_partial(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ w, long long n,
                                 double *__restrict__ partial) {
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc += (w ? w[i] : 1.0) * x[i] * y[i];
  block_sum_to(acc, partial + blockIdx.x);
}
__global__ void k_cg_final(const double *__restrict__ partial, int n, double *__restrict__ out) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
  block_sum_to(acc, out);
}
// alpha = rr / pAp;  x += alpha p;  r -= alpha Ap;  partial sums of the new (weighted) r.r
__global__ void k_cg_update_partial(double *__restrict__ x, double *__restrict__ r, const double *__restrict__ p, const double *__restrict__ Ap,
                                    const double *__restrict__ w, long long n, const double *__restrict__ rr, const double *__restrict__ pAp,
                                    double *__restrict__ partial) {
  const double alpha = *rr / *pAp;
  double       acc   = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    x[i] = fma(alpha, p[i], x[i]);
    const double ri = fma(-alpha, Ap[i], r[i]);
    r[i] = ri;
    acc += (w ? w[i] : 1.0) * ri * ri;
  }
  block_sum_to(acc, partial + blockIdx.x);
}
// beta = rr_new / rr;  p = r + beta p
__global__ void k_cg_direction(double *__restrict__ p, const double *__restrict__ r, long long n, const double *__restrict__ rr_new,
                               const double *__restrict__ rr) {
  const double beta = *rr_new / *rr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = fma(beta, p[i], r[i]);
}

// essential (Dirichlet) boundary conditions: constrained rows of A become identity rows -- Ap[i] = p[i] where free[i] == 0
__global__ void k_cg_constrain(double *__restrict__ Ap, const double *__restrict__ p, const double *__restrict__ free_mask, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (free_mask[i] == 0.0) Ap[i] = p[i];
}

int cg_scratch(B200Ceed ceed) {
  // the partial sums use a FIXED grid of kCgBlocks CTAs (reproducible whatever the SM count): the buffer must hold them all
  const size_t want = std::max<size_t>((size_t)ceed->num_sms * 16, (size_t)kCgBlocks);
  if (ceed->scratch_len < want) {
    B200_CALL(b200_dfree(ceed, ceed->d_scratch));
    ceed->d_scratch = nullptr, ceed->scratch_len = 0;
    B200_CALL(b200_dmalloc(ceed, (void **)&ceed->d_scratch, want * sizeof(double)));
    ceed->scratch_len = want;
  }
  return B200_SUCCESS;
}
}  // namespace

#define CG_LAUNCH(ceed, kernel, grid, ...)                                                                              \
  do {                                                                                                                  \
    B200_CHECK(!b200_compile_only(), ceed, B200_ERROR_BACKEND, "CEED_B200_COMPILE_ONLY is set; kernels cannot run");    \
    kernel<<<grid, kThreads, 0, (ceed)->stream>>>(__VA_ARGS__);                                                         \
    (ceed)->launch_count++;                                                                                             \
    B200_CUDA(ceed, cudaGetLastError());                                                                                \
  } while (0)

extern "C" int ceedb200_cg_dot(B200Ceed ceed, const double *d_x, const double *d_y, const double *d_w, long long n, double *d_out) {
  B200_CALL(cg_scratch(ceed));
  CG_LAUNCH(ceed, k_cg_dot_partial, kCgBlocks, d_x, d_y, d_w, n, ceed->d_scratch);
  CG_LAUNCH(ceed, k_cg_final, 1, ceed->d_scratch, kCgBlocks, d_out);
  return B200_SUCCESS;
}
extern "C" int ceedb200_cg_update(B200Ceed ceed, double *d_x, double *d_r, const double *d_p, const double *d_Ap, const double *d_w, long long n,
                                  const double *d_rr, const double *d_pAp, double *d_rr_new) {
  B200_CALL(cg_scratch(ceed));
  CG_LAUNCH(ceed, k_cg_update_partial, kCgBlocks, d_x, d_r, d_p, d_Ap, d_w, n, d_rr, d_pAp, ceed->d_scratch);
  CG_LAUNCH(ceed, k_cg_final, 1, ceed->d_scratch, kCgBlocks, d_rr_new);
  return B200_SUCCESS;
}
extern "C" int ceedb200_cg_constrain(B200Ceed ceed, double *d_Ap, const double *d_p, const double *d_free_mask, long long n) {
  CG_LAUNCH(ceed, k_cg_constrain, kCgBlocks, d_Ap, d_p, d_free_mask, n);
  return B200_SUCCESS;
}
extern "C" int ceedb200_cg_direction(B200Ceed ceed, double *d_p, const double *d_r, long long n, const double *d_rr_new, const double *d_rr) {
  CG_LAUNCH(ceed, k_cg_direction, kCgBlocks, d_p, d_r, n, d_rr_new, d_rr);
  return B200_SUCCESS;
}
