// Microbenchmark: achieved HBM read bandwidth of the quadrature-data access patterns of the fused operator kernel.
// A buffer of NE "elements" x (NC comps x Q3 points) doubles is read once per launch by persistent warps, one element per warp
// iteration, in different orders.  build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/ubench/stream_patterns stream_patterns.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
constexpr int NC = 7, Q = 8, Q3 = Q * Q * Q, Q2 = Q * Q;
constexpr int ELEM_DOUBLES = NC * Q3;

// A: z-line order with one layer of look-ahead (what the z-line QFunction stage does)
__global__ void pat_zline1(const double *__restrict__ q, long long ne, double *out) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = (long long)gridDim.x * (blockDim.x >> 5);
  double acc = 0;
  for (long long e = w; e < ne; e += nw) {
    const double *b = q + e * ELEM_DOUBLES;
    for (int r = 0; r < Q2 / 32; r++) {
      double nx[NC];
#pragma unroll
      for (int c = 0; c < NC; c++) nx[c] = __ldg(b + c * Q3 + r * 32 + lane);
#pragma unroll
      for (int z = 0; z < Q; z++) {
        double cur[NC];
#pragma unroll
        for (int c = 0; c < NC; c++) cur[c] = nx[c];
        if (z + 1 < Q) {
#pragma unroll
          for (int c = 0; c < NC; c++) nx[c] = __ldg(b + c * Q3 + (z + 1) * Q2 + r * 32 + lane);
        }
#pragma unroll
        for (int c = 0; c < NC; c++) acc = fma(cur[c], 1.0000001, acc);
      }
    }
  }
  if (acc == 123.456) out[0] = acc;
}
// B: z-line order, all layers of a round issued at once
__global__ void pat_zline_all(const double *__restrict__ q, long long ne, double *out) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = (long long)gridDim.x * (blockDim.x >> 5);
  double acc = 0;
  for (long long e = w; e < ne; e += nw) {
    const double *b = q + e * ELEM_DOUBLES;
    for (int r = 0; r < Q2 / 32; r++) {
      double v[Q][NC];
#pragma unroll
      for (int z = 0; z < Q; z++)
#pragma unroll
        for (int c = 0; c < NC; c++) v[z][c] = __ldg(b + c * Q3 + z * Q2 + r * 32 + lane);
#pragma unroll
      for (int z = 0; z < Q; z++)
#pragma unroll
        for (int c = 0; c < NC; c++) acc = fma(v[z][c], 1.0000001, acc);
    }
  }
  if (acc == 123.456) out[0] = acc;
}
// C: pointwise order, U iterations (32 consecutive points each) in flight
template <int U>
__global__ void pat_points(const double *__restrict__ q, long long ne, double *out) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = (long long)gridDim.x * (blockDim.x >> 5);
  double acc = 0;
  for (long long e = w; e < ne; e += nw) {
    const double *b = q + e * ELEM_DOUBLES;
    for (int k = 0; k < Q3 / 32; k += U) {
      double v[U][NC];
#pragma unroll
      for (int u = 0; u < U; u++)
#pragma unroll
        for (int c = 0; c < NC; c++) v[u][c] = __ldg(b + c * Q3 + (k + u) * 32 + lane);
#pragma unroll
      for (int u = 0; u < U; u++)
#pragma unroll
        for (int c = 0; c < NC; c++) acc = fma(v[u][c], 1.0000001, acc);
    }
  }
  if (acc == 123.456) out[0] = acc;
}
// D: fully sequential 16-byte loads, U x 512 B per warp in flight
template <int U>
__global__ void pat_seq(const double *__restrict__ q, long long ne, double *out) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = (long long)gridDim.x * (blockDim.x >> 5);
  double acc = 0;
  for (long long e = w; e < ne; e += nw) {
    const double2 *b = (const double2 *)(q + e * ELEM_DOUBLES);
    for (int k = 0; k < ELEM_DOUBLES / 64; k += U) {
      double2 v[U];
#pragma unroll
      for (int u = 0; u < U; u++) v[u] = __ldg(b + (k + u) * 32 + lane);
#pragma unroll
      for (int u = 0; u < U; u++) acc = fma(v[u].x, 1.0000001, acc) + v[u].y;
    }
  }
  if (acc == 123.456) out[0] = acc;
}
// E: array-of-structures (point-major, NC contiguous doubles per point), U iterations in flight
template <int U>
__global__ void pat_aos(const double *__restrict__ q, long long ne, double *out) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = (long long)gridDim.x * (blockDim.x >> 5);
  double acc = 0;
  for (long long e = w; e < ne; e += nw) {
    const double *b = q + e * ELEM_DOUBLES;
    for (int k = 0; k < Q3 / 32; k += U) {
      double v[U][NC];
#pragma unroll
      for (int u = 0; u < U; u++)
#pragma unroll
        for (int c = 0; c < NC; c++) v[u][c] = __ldg(b + ((k + u) * 32 + lane) * NC + c);
#pragma unroll
      for (int u = 0; u < U; u++)
#pragma unroll
        for (int c = 0; c < NC; c++) acc = fma(v[u][c], 1.0000001, acc);
    }
  }
  if (acc == 123.456) out[0] = acc;
}

template <typename K>
static void run(const char *name, K kernel, const double *q, long long ne, double *out, int blocks_per_sm, int threads) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const int grid = 148 * blocks_per_sm;
  for (int i = 0; i < 3; i++) kernel<<<grid, threads>>>(q, ne, out);
  float best = 1e30f;
  for (int i = 0; i < 10; i++) {
    cudaEventRecord(a);
    kernel<<<grid, threads>>>(q, ne, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  cudaError_t err = cudaGetLastError();
  printf("%-22s warps/SM %3d  %.3f ms  %7.1f GB/s  %s\n", name, blocks_per_sm * threads / 32, best, ne * ELEM_DOUBLES * 8.0 / best / 1e6,
         err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
  const long long ne = 45360;
  double *q, *out;
  cudaMalloc(&q, ne * ELEM_DOUBLES * 8);
  cudaMalloc(&out, 8);
  cudaMemset(q, 0, ne * ELEM_DOUBLES * 8);
  for (int bps : {3, 6, 12}) {
    run("zline look-ahead 1", pat_zline1, q, ne, out, bps, 128);
    run("zline all layers", pat_zline_all, q, ne, out, bps, 128);
    run("points U=1", pat_points<1>, q, ne, out, bps, 128);
    run("points U=4", pat_points<4>, q, ne, out, bps, 128);
    run("points U=8", pat_points<8>, q, ne, out, bps, 128);
    run("sequential U=4", pat_seq<4>, q, ne, out, bps, 128);
    run("sequential U=14", pat_seq<14>, q, ne, out, bps, 128);
    run("aos U=1", pat_aos<1>, q, ne, out, bps, 128);
    run("aos U=4", pat_aos<4>, q, ne, out, bps, 128);
  }
  return 0;
}
