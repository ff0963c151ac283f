// The one collective of the path -- the all-gather of the per-candidate costs before selection -- over
// NVLink peer memory instead of a library call: every rank owns a gather buffer that all ranks of the job map (CUDA IPC),
// the finish kernel of the gait evaluation stores each cost straight into all of them (hsl_kernels.cu), and what is left of
// the collective is a flag per rank: hsl_gather_signal_kernel (release) after the stores, hsl_gather_wait_kernel (acquire)
// before the selection reads.  Buffers alternate by call parity, flags carry the call number, so a rank that runs ahead
// never overwrites what a slower rank is still reading (it cannot get two calls ahead: it waits for that rank's flag).
#include <cuda_runtime.h>
#include <stdint.h>

#include "hsl_internal.h"

namespace {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

struct FlagPtrs { unsigned long long* p[HSL_MAX_PEERS]; };

// One block.  Candidates n_used .. n_per_rank-1 of this rank's segment do not exist: NaN / 0 at every peer.  Then the flag:
// the stores of the finish kernel (earlier on this stream) and of this block are ordered before it by the fence + release.
__global__ void hsl_gather_signal_kernel(const __grid_constant__ HslPeerOut peers, int64_t n_used, int64_t n_per_rank,
                                         const __grid_constant__ FlagPtrs flags, unsigned long long epoch) {
  const double nanv = __longlong_as_double(0x7ff8000000000000LL);
  for (int64_t i = n_used + threadIdx.x; i < n_per_rank; i += blockDim.x)
    for (int r = 0; r < peers.n; r++) { peers.cot[r][i] = nanv; peers.status[r][i] = 0; }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < peers.n) st_release_sys(flags.p[threadIdx.x], epoch);
}

// One block, thread r waits for rank r's flag.  Flags only grow.
__global__ void hsl_gather_wait_kernel(const unsigned long long* flags, int nranks, unsigned long long epoch) {
  if (threadIdx.x < nranks) {
    while (ld_acquire_sys(flags + threadIdx.x) < epoch) __nanosleep(100);
  }
  __syncthreads();
  __threadfence_system();
}

}  // namespace

cudaError_t hsl_launch_gather_signal(const HslPeerOut& peers, int64_t n_used, int64_t n_per_rank, unsigned long long* const* flag_at_peer,
                                     unsigned long long epoch, cudaStream_t st) {
  FlagPtrs f;
  for (int r = 0; r < HSL_MAX_PEERS; r++) f.p[r] = r < peers.n ? flag_at_peer[r] : nullptr;
  hsl_gather_signal_kernel<<<1, 256, 0, st>>>(peers, n_used, n_per_rank, f, epoch);
  return cudaGetLastError();
}
cudaError_t hsl_launch_gather_wait(const unsigned long long* flags, int nranks, unsigned long long epoch, cudaStream_t st) {
  hsl_gather_wait_kernel<<<1, 32, 0, st>>>(flags, nranks, epoch);
  return cudaGetLastError();
}
