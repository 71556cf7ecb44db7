// Microbenchmark 4: why did bulk_ring.cu not pipeline?  Direct probes of cp.async.bulk (UBLKCP) and of tensor-map TMA (UTMALDG).
//  A. burst probe: one CTA per SM issues N copies of S bytes back to back (distinct mbarriers), then waits for all: time(N) vs time(1)
//     tells whether copies of one CTA overlap.  Timed with clock64 inside the kernel (cycles) -- cold data each repetition.
//  B. ring with a DEDICATED producer warp (never consumes) vs the combined producer/consumer of bulk_ring.cu.
//  C. the same ring fed by ONE 2-D tensor-map copy per slot (box = Q^2 points x ncq rows) instead of ncq bulk copies.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bulk_probe bulk_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(void *bar, int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(void *bar, int parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, int bytes, void *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma2d_g2s(void *dst, const CUtensorMap *map, int c0, int c1, void *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"(map),
               "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}

extern __shared__ __align__(128) char smem[];

// ---- A. burst probe
__global__ void k_burst(const char *__restrict__ q, long long stride_cta, int S, int N, int rep, long long *cycles) {
  unsigned long long *bar = (unsigned long long *)smem;
  char *buf = smem + 128;
  if (threadIdx.x == 0) {
    for (int i = 0; i < N; i++) mbar_init(bar + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const char *src = q + (long long)blockIdx.x * stride_cta + (long long)rep * N * S;
    const long long t0 = clock64();
    for (int i = 0; i < N; i++) {
      mbar_expect_tx(bar + i, S);
      bulk_g2s(buf + i * S, src + (long long)i * S, S, bar + i);
    }
    const long long t1 = clock64();
    for (int i = 0; i < N; i++) mbar_wait(bar + i, 0);
    const long long t2 = clock64();
    cycles[blockIdx.x * 2] = t1 - t0;
    cycles[blockIdx.x * 2 + 1] = t2 - t0;
  }
}

// ---- B/C. ring.  mode 0: combined producer/consumer (bulk copies); 1: dedicated producer warp (bulk copies);
//                  2: dedicated producer warp, one 2-D tensor copy per slot; 3: combined, lane 0 issues all copies itself
__global__ void k_ring(const char *__restrict__ q, const __grid_constant__ CUtensorMap map, long long nchunks, int Q, int ncq, int NS, int mode, double *out) {
  const int Q2 = Q * Q, Q3 = Q2 * Q;
  const int copy_bytes = Q2 * 8, slot_bytes = ncq * copy_bytes;
  unsigned long long *full = (unsigned long long *)smem, *empty = full + NS;
  char *ring = smem + 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool dedicated = mode == 1 || mode == 2;
  const int cwarps = (blockDim.x >> 5) - (dedicated ? 1 : 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; i++) mbar_init(full + i, 1), mbar_init(empty + i, cwarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long nsteps = (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x;
  auto issue = [&](long long t) {
    const int slot = (int)(t % NS);
    const long long k = t / NS;
    if (k > 0) mbar_wait(empty + slot, (int)((k - 1) & 1));
    const long long id = blockIdx.x + t * gridDim.x;
    if (lane == 0) mbar_expect_tx(full + slot, slot_bytes);
    __syncwarp();
    if (mode == 2) {
      if (lane == 0) tma2d_g2s(ring + slot * slot_bytes, &map, (int)(id % Q) * Q2, (int)(id / Q) * ncq, full + slot);
    } else if (mode == 3) {
      if (lane == 0)
        for (int c = 0; c < ncq; c++)
          bulk_g2s(ring + slot * slot_bytes + c * copy_bytes, q + (((id / Q) * ncq + c) * (long long)Q3 + (id % Q) * Q2) * 8, copy_bytes, full + slot);
    } else if (lane < ncq) {
      bulk_g2s(ring + slot * slot_bytes + lane * copy_bytes, q + (((id / Q) * ncq + lane) * (long long)Q3 + (id % Q) * Q2) * 8, copy_bytes, full + slot);
    }
  };
  if (dedicated && warp == 0) {
    for (long long t = 0; t < nsteps; t++) issue(t);
    return;
  }
  if (!dedicated && warp == 0)
    for (int t = 0; t < NS - 1 && t < nsteps; t++) issue(t);
  const int ctid = threadIdx.x - (dedicated ? 32 : 0);
  double acc = 0;
  for (long long s = 0; s < nsteps; s++) {
    const int slot = (int)(s % NS);
    mbar_wait(full + slot, (int)((s / NS) & 1));
    const double *v = (const double *)(ring + slot * slot_bytes);
    if (ctid < Q2)
      for (int c = 0; c < ncq; c++) acc += v[c * Q2 + ctid];
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + slot);
    if (!dedicated && warp == 0 && s + NS - 1 < nsteps) issue(s + NS - 1);
  }
  if (acc == 123.456) out[0] = acc;
}

int main() {
  const long long bytes = 3LL << 30;
  char *q;
  double *out;
  long long *cyc, hcyc[2 * 148];
  cudaMalloc(&q, bytes);
  cudaMalloc(&out, 8);
  cudaMalloc(&cyc, sizeof(hcyc));
  cudaMemset(q, 0, bytes);
  cudaFuncSetAttribute(k_burst, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(k_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  printf("A. burst probe: cycles (median over CTAs) to issue / to complete N copies of S bytes from one thread\n");
  for (int grid : {1, 148})
    for (int S : {512, 4096, 16384})
      for (int N : {1, 2, 4, 8, 12}) {
        if (128 + N * S > 200 * 1024) continue;
        long long best_issue = 1LL << 60, best_done = 1LL << 60;
        for (int rep = 0; rep < 5; rep++) {
          k_burst<<<grid, 32, 128 + N * S>>>(q + (long long)rep * (400LL << 20), 2LL << 20, S, N, rep, cyc);
          cudaMemcpy(hcyc, cyc, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
          // median over CTAs
          long long a[148], b[148];
          for (int i = 0; i < grid; i++) a[i] = hcyc[2 * i], b[i] = hcyc[2 * i + 1];
          for (int i = 0; i < grid; i++)
            for (int j = i + 1; j < grid; j++) {
              if (a[j] < a[i]) { long long t = a[i]; a[i] = a[j]; a[j] = t; }
              if (b[j] < b[i]) { long long t = b[i]; b[i] = b[j]; b[j] = t; }
            }
          if (rep > 0 && a[grid / 2] < best_issue) best_issue = a[grid / 2];
          if (rep > 0 && b[grid / 2] < best_done) best_done = b[grid / 2];
        }
        printf("  grid %3d  S %6d  N %2d : issue %6lld cyc  done %7lld cyc  (%.1f B/cyc)\n", grid, S, N, best_issue, best_done, (double)N * S / best_done);
      }
  // tensor map: 2-D view [rows = nelem * ncq][Q^3] of doubles
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  printf("B/C. ring, layer slots: mode 0 combined, 1 dedicated producer warp, 2 dedicated + one 2-D tensor copy per slot, 3 combined + single issuing lane\n");
  printf("%4s %3s %3s %3s %7s %9s %9s\n", "mode", "Q", "NS", "cta", "threads", "ms", "GB/s");
  for (int Q : {6, 8, 10}) {
    const int ncq = 7, Q2 = Q * Q, Q3 = Q2 * Q;
    const long long nelem = bytes / (ncq * Q3 * 8);
    const long long nchunks = nelem * Q;
    CUtensorMap map;
    cuuint64_t gdim[2] = {(cuuint64_t)Q3, (cuuint64_t)(nelem * ncq)};
    cuuint64_t gstride[1] = {(cuuint64_t)Q3 * 8};
    cuuint32_t box[2] = {(cuuint32_t)Q2, (cuuint32_t)ncq};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, q, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
    for (int mode : {0, 1, 2, 3})
      for (int NS : {3, 6})
        for (int cta_per_sm : {1, 2, 4, 8}) {
          const int cthreads = Q2 <= 64 ? 64 : 128;
          const int threads = cthreads + ((mode == 1 || mode == 2) ? 32 : 0);
          const int smem_bytes = 128 + NS * ncq * Q2 * 8;
          if ((long long)smem_bytes * cta_per_sm > 220 * 1024) continue;
          if (mode == 2 && r != CUDA_SUCCESS) continue;
          float best = 1e30f;
          for (int i = 0; i < 4; i++) {
            cudaEventRecord(a);
            k_ring<<<148 * cta_per_sm, threads, smem_bytes>>>(q, map, nchunks, Q, ncq, NS, mode, out);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            if (i > 0 && ms < best) best = ms;
          }
          cudaError_t err = cudaGetLastError();
          if (err != cudaSuccess) {
            printf("error: %s\n", cudaGetErrorString(err));
            return 1;
          }
          printf("%4d %3d %3d %3d %7d %9.3f %9.1f\n", mode, Q, NS, cta_per_sm, threads, best, nelem * (double)(ncq * Q3 * 8) / best / 1e6);
        }
  }
  return 0;
}
