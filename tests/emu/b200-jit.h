// tests/emu/b200-jit.h -- TEST INFRASTRUCTURE.  Stands in for libceed_b200/csrc/jit/b200-jit.h when the generated source of a fused
// operator kernel is compiled by g++ for the HOST (tests/kernel_emu.py): CUDA's execution-space keywords vanish, the built-in index
// variables become thread-local objects, every CUDA thread of a CTA is one OS thread and __syncwarp / __syncthreads are barriers.
// Only what the generated kernels without asynchronous-copy stages use is provided (no cp.async, mbarrier, named barriers).
#pragma once
#include <pthread.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define __global__
#define __device__
#define __host__
#define __constant__
#define __shared__
#define __noinline__ __attribute__((noinline))
#define __forceinline__ inline
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)
#define __restrict__ __restrict

struct b200_emu_dim3 {
  unsigned x, y, z;
};
static thread_local b200_emu_dim3 threadIdx, blockIdx;
static b200_emu_dim3              blockDim, gridDim;

struct int2 {
  int x, y;
} __attribute__((aligned(8)));
struct int4 {
  int x, y, z, w;
} __attribute__((aligned(16)));
struct double2 {
  double x, y;
} __attribute__((aligned(16)));
static inline int2    make_int2(int x, int y) { return int2{x, y}; }
static inline int4    make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __ldcg(const T *p) { return *p; }
template <typename T> static inline T __ldcs(const T *p) { return *p; }

static pthread_barrier_t b200_emu_cta_barrier, b200_emu_warp_barrier[32];
static inline void       __syncthreads() { pthread_barrier_wait(&b200_emu_cta_barrier); }
static inline void       __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&b200_emu_warp_barrier[threadIdx.x >> 5]); }
// named barrier `id` (bar.sync id, count) of a multi-warp element group: created on first use, torn down by the driver after the launch
static pthread_barrier_t b200_emu_named[16];
static int               b200_emu_named_count[16];
static pthread_mutex_t   b200_emu_named_lock = PTHREAD_MUTEX_INITIALIZER;
static inline void       b200_emu_named_barrier(int id, int count) {
  pthread_mutex_lock(&b200_emu_named_lock);
  if (!b200_emu_named_count[id]) {
    pthread_barrier_init(&b200_emu_named[id], nullptr, count);
    b200_emu_named_count[id] = count;
  }
  pthread_mutex_unlock(&b200_emu_named_lock);
  pthread_barrier_wait(&b200_emu_named[id]);
}
static inline void       __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline double     atomicAdd(double *p, double v) {
  static pthread_mutex_t m = PTHREAD_MUTEX_INITIALIZER;  // (one lock: tests are small)
  pthread_mutex_lock(&m);
  const double old = *p;
  *p               = old + v;
  pthread_mutex_unlock(&m);
  return old;
}
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }

// QFunction source contract of the JIT prelude (libceed_b200/csrc/jit/b200-jit.h)
#define CEED_QFUNCTION(name) inline int name
#define CEED_QFUNCTION_HELPER inline
#define CeedPragmaSIMD
#ifndef CEED_Q_VLA
#define CEED_Q_VLA 1
#endif
#include <ceed/types.h>
