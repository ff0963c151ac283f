// The one collective of the path -- the all-gather of the per-candidate costs before selection -- over
// NVLink peer memory instead of a library call: every rank owns a gather buffer that all ranks of the job map (CUDA IPC),
// the finish kernel of the gait evaluation stores each cost straight into all of them and its last block raises this rank's
// flag at every peer (release; hsl_gather_publish, hsl_internal.h); the reader -- hsl_gather_wait_kernel, or the first
// instructions of the argmin kernel (hsl_select.cu) -- acquires the flags of all ranks before it touches the costs.  Buffers alternate by call parity, flags carry the call number, so a rank that runs ahead
// never overwrites what a slower rank is still reading (it cannot get two calls ahead: it waits for that rank's flag).
#include <cuda_runtime.h>
#include <stdint.h>

#include "hsl_gather_dev.cuh"
#include "hsl_internal.h"

namespace {

// A rank without candidates launches no finish kernel: one block pads its whole segment and raises the flags.
__global__ void hsl_gather_signal_kernel(const __grid_constant__ HslPeerOut peers) { hsl_gather_publish(peers); }

// One block, thread r waits for rank r's flag.  Flags only grow.
__global__ void hsl_gather_wait_kernel(const unsigned long long* flags, int nranks, unsigned long long epoch) {
  if (threadIdx.x < nranks) hsl_wait_flag(flags + threadIdx.x, epoch, (unsigned long long*)(flags + HSL_MAX_PEERS + 1));
  __syncthreads();
  __threadfence_system();
}

}  // namespace

cudaError_t hsl_launch_gather_signal(const HslPeerOut& peers, cudaStream_t st) {
  hsl_gather_signal_kernel<<<1, 256, 0, st>>>(peers);
  return cudaGetLastError();
}
cudaError_t hsl_launch_gather_wait(const unsigned long long* flags, int nranks, unsigned long long epoch, cudaStream_t st) {
  hsl_gather_wait_kernel<<<1, 32, 0, st>>>(flags, nranks, epoch);
  return cudaGetLastError();
}
