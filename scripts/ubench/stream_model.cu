// Microbenchmark 2: what limits per-SM read bandwidth at low occupancy?  Each warp streams its own 28 KB blocks.
// Parameters: bytes per lane per load (8 / 16), loads in flight per lane (U), streams per warp (1 = sequential, 7 = SoA comps),
// resident warps per SM.  build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stream_model stream_model.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ELEM_BYTES = 7 * 512 * 8;

template <typename T, int U, int STREAMS>
__global__ void k(const char *__restrict__ q, long long ne, double *out) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = (long long)gridDim.x * (blockDim.x >> 5);
  constexpr int STREAM_BYTES = ELEM_BYTES / STREAMS;        // bytes per stream per element
  constexpr int STEP = 32 * sizeof(T);                      // bytes per warp-wide load
  constexpr int NSTEP = STREAM_BYTES / STEP;                // loads per stream per element
  double acc = 0;
  for (long long e = w; e < ne; e += nw) {
    const char *b = q + e * ELEM_BYTES + lane * sizeof(T);
    for (int s = 0; s < NSTEP; s += U) {
      T v[U][STREAMS];
#pragma unroll
      for (int u = 0; u < U; u++)
#pragma unroll
        for (int c = 0; c < STREAMS; c++) v[u][c] = __ldg((const T *)(b + c * STREAM_BYTES + (s + u) * STEP));
#pragma unroll
      for (int u = 0; u < U; u++)
#pragma unroll
        for (int c = 0; c < STREAMS; c++) acc += *(const double *)&v[u][c];
    }
  }
  if (acc == 123.456) out[0] = acc;
}

template <typename K>
static void run(const char *name, K kernel, const char *q, long long ne, double *out, int warps_per_sm) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const int threads = 64, grid = 148 * warps_per_sm / 2;
  for (int i = 0; i < 2; i++) kernel<<<grid, threads>>>(q, ne, out);
  float best = 1e30f;
  for (int i = 0; i < 6; i++) {
    cudaEventRecord(a);
    kernel<<<grid, threads>>>(q, ne, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  printf("%-34s warps/SM %3d  %.3f ms  %7.1f GB/s  %5.2f B/clk/warp@1.9GHz\n", name, warps_per_sm, best, ne * (double)ELEM_BYTES / best / 1e6,
         ne * (double)ELEM_BYTES / best / 1e6 / 148 / warps_per_sm / 1.9);
}

int main() {
  const long long ne = 45360;
  char *q;
  double *out;
  cudaMalloc(&q, ne * ELEM_BYTES);
  cudaMalloc(&out, 8);
  cudaMemset(q, 0, ne * ELEM_BYTES);
  for (int w : {4, 8, 12, 16, 24, 32}) {
    run("8B  seq      U=1", k<double, 1, 1>, q, ne, out, w);
    run("8B  seq      U=4", k<double, 4, 1>, q, ne, out, w);
    run("8B  seq      U=16", k<double, 16, 1>, q, ne, out, w);
    run("16B seq      U=1", k<double2, 1, 1>, q, ne, out, w);
    run("16B seq      U=4", k<double2, 4, 1>, q, ne, out, w);
    run("16B seq      U=16", k<double2, 16, 1>, q, ne, out, w);
    run("8B  7 streams U=1", k<double, 1, 7>, q, ne, out, w);
    run("8B  7 streams U=4", k<double, 4, 7>, q, ne, out, w);
    run("16B 7 streams U=1", k<double2, 1, 7>, q, ne, out, w);
    run("16B 7 streams U=4", k<double2, 4, 7>, q, ne, out, w);
  }
  return 0;
}
