// tests/emu/b200-tma.h -- TEST INFRASTRUCTURE.  Host stand-in for libceed_b200/csrc/jit/b200-tma.h (bulk asynchronous copies + mbarrier)
// for the CPU emulation of generated kernels (tests/kernel_emu.py): a bulk copy is performed at once and then completes its byte count on
// the mbarrier; the mbarrier lives in the 8 bytes of "shared memory" the kernel reserved for it (transaction bytes outstanding, arrivals
// left in the current phase, phase parity); waiting lanes spin on the phase.  Same call sequence as on the device, so the generator's
// arm / wait / re-arm protocol (which buffer is re-armed when, phase parities per iteration) is what gets tested.
#pragma once
#include <sched.h>

extern double sm[] __attribute__((aligned(16)));

struct b200_emu_mbar {
  int           tx;        // transaction bytes announced and not yet completed
  short         left;      // arrivals left in the current phase
  unsigned char count;     // arrivals per phase
  unsigned char phase;     // parity of the current (incomplete) phase
};
static_assert(sizeof(b200_emu_mbar) == 8, "an mbarrier is 8 bytes of shared memory");
static pthread_mutex_t b200_emu_mbar_lock = PTHREAD_MUTEX_INITIALIZER;
static inline b200_emu_mbar *b200_emu_mbar_at(unsigned bar) { return (b200_emu_mbar *)((char *)sm + bar); }
static inline void           b200_emu_mbar_check(b200_emu_mbar *m) {
  if (m->left == 0 && m->tx == 0) {
    m->left = m->count;
    __atomic_store_n(&m->phase, (unsigned char)(m->phase ^ 1), __ATOMIC_RELEASE);
  }
}

static inline unsigned b200_smem_u32(const void *p) { return (unsigned)((const char *)p - (const char *)sm); }
static inline void     b200_mbar_init(unsigned bar, int count) {
  b200_emu_mbar *m = b200_emu_mbar_at(bar);
  m->tx = 0, m->left = (short)count, m->count = (unsigned char)count, m->phase = 0;
}
static inline void b200_mbar_expect_tx(unsigned bar, unsigned bytes) {
  pthread_mutex_lock(&b200_emu_mbar_lock);
  b200_emu_mbar *m = b200_emu_mbar_at(bar);
  m->tx += (int)bytes;
  m->left -= 1;
  b200_emu_mbar_check(m);
  pthread_mutex_unlock(&b200_emu_mbar_lock);
}
static inline void b200_mbar_wait(unsigned bar, int parity) {
  b200_emu_mbar *m = b200_emu_mbar_at(bar);
  while (__atomic_load_n(&m->phase, __ATOMIC_ACQUIRE) == (unsigned char)parity) sched_yield();
}
static inline void b200_fence_proxy_async() {}
static inline void b200_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  if (((uintptr_t)src & 15) || (dst & 15) || (bytes & 15)) {
    fprintf(stderr, "b200 emulation: cp.async.bulk needs 16-byte aligned source, destination and size (src %p dst %u bytes %u)\n", src, dst, bytes);
    abort();
  }
  memcpy((char *)sm + dst, src, bytes);
  pthread_mutex_lock(&b200_emu_mbar_lock);
  b200_emu_mbar *m = b200_emu_mbar_at(bar);
  m->tx -= (int)bytes;
  b200_emu_mbar_check(m);
  pthread_mutex_unlock(&b200_emu_mbar_lock);
}
static inline void b200_bulk_prefetch_l2(const void *, unsigned) {}
