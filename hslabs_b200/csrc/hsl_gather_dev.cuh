// Device side of the peer-memory gather (hsl_gather.cu): flag accesses at system scope and the publishing tail of a kernel
// that stored into the peers' gather buffers.  Included by the .cu files only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "hsl_internal.h"

// release / acquire at system scope
__device__ __forceinline__ void hsl_st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long hsl_ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// Wait until the flag has reached `epoch`.  Bounded: a rank that never arrives (crashed process, mismatched call sequence)
// must not hang the GPU -- after HSL_GATHER_TIMEOUT_NS the wait gives up and records the epoch in *timed_out, which
// hsl_gather_check reports on the host.
#ifndef HSL_GATHER_TIMEOUT_NS
#define HSL_GATHER_TIMEOUT_NS 30000000000ull
#endif
__device__ __forceinline__ void hsl_wait_flag(const unsigned long long* flag, unsigned long long epoch, unsigned long long* timed_out) {
  unsigned long long t0 = 0;
  unsigned spins = 0;
  while (hsl_ld_acquire_sys(flag) < epoch) {
    __nanosleep(50);
    if ((++spins & 1023u) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > HSL_GATHER_TIMEOUT_NS) { *timed_out = epoch; break; }
    }
  }
}
// Tail of a kernel whose blocks stored into the peers' gather buffers: the block that finishes last pads the unused
// entries, then raises this rank's flag at every peer (the threadFenceReduction pattern at system scope).  Every thread of
// every block must call it.
__device__ __forceinline__ void hsl_gather_publish(const HslPeerOut& peers) {
  __shared__ bool last;
  __threadfence_system();
  asm volatile("bar.sync 0;" ::: "memory");   // reached from two places of the finish kernel by whole warps
  if (threadIdx.x == 0) last = (atomicAdd(peers.ticket, 1u) == gridDim.x - 1);
  asm volatile("bar.sync 0;" ::: "memory");   // reached from two places of the finish kernel by whole warps
  if (!last) return;
  const double nanv = __longlong_as_double(0x7ff8000000000000LL);
  for (int64_t i = peers.pad_lo + threadIdx.x; i < peers.pad_hi; i += blockDim.x)
    for (int r = 0; r < peers.n; r++) { peers.cot[r][i] = nanv; peers.status[r][i] = 0; }
  __threadfence_system();
  asm volatile("bar.sync 0;" ::: "memory");   // reached from two places of the finish kernel by whole warps
  if (threadIdx.x < peers.n) hsl_st_release_sys(peers.flag[threadIdx.x], peers.epoch);
  if (threadIdx.x == 0) *peers.ticket = 0;
}
