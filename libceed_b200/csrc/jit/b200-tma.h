// b200-tma.h -- device helpers of the fused operator kernel for bulk asynchronous copies (Hopper/Blackwell copy engine).
//
// cp.async.bulk (1-D TMA, SASS UBLKCP): ONE lane hands a whole contiguous block global -> shared memory to the copy engine:
// no load instructions, no registers held by loads in flight.  Completion is counted in bytes on an mbarrier in shared
// memory (SASS SYNCS.*) that the consuming lanes wait on with try_wait (hardware sleep, not a spin on a memory location).
// Measured on B200 (scripts/ubench/bulk_probe.cu): ~60 cycles to issue a copy, ~75 cycles of serialised engine time per copy
// and ~31 B/clk per SM streaming rate, so copies should be a few KB or larger -- the generator issues one copy per
// (quadrature-data component, element group).
#pragma once

__device__ __forceinline__ unsigned b200_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void b200_mbar_init(unsigned bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void b200_mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void b200_mbar_wait(unsigned bar, int parity) {
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " B200_WAIT_%=:\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra B200_DONE_%=;\n"
      " bra B200_WAIT_%=;\n"
      " B200_DONE_%=:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// generic-proxy reads of a shared buffer (by any lane, ordered before this point by a barrier) -> async-proxy writes into it
__device__ __forceinline__ void b200_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void b200_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// cp.async.bulk.prefetch.L2: one instruction asks the copy engine to pull a whole contiguous range into L2 (no shared memory,
// no registers): the demand loads that follow a few microseconds later hit L2 instead of HBM.
__device__ __forceinline__ void b200_bulk_prefetch_l2(const void *src, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
