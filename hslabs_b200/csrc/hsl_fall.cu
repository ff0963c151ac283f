// Fall / perturbation sweep (BASELINE configs[4]): many copies of the reference's closed loop
//
//   position_control_test -> step() -> simulate_ode            player.cpp:358-364, 326-340
//     set_position_control_torques (PD about the evaluated gait)  player.cpp:393-432
//     kick_torso (a force dv/dt on the torso for one step)        player.cpp:585-605
//     fall_check (torso z < hc once play_t >= tmin)               player.cpp:669-681
//     visualizer::simulate_odeworld: collide, dWorldQuickStep     visualization.cpp:296-337
//
// integrated side by side.  Two kernels state the same arithmetic: hsl_fall_warp_kernel (hsl_fall_warp.cuh, the default:
// one warp per world, lane = body, the world in registers) and hsl_fall_kernel below (one thread per world; its source,
// hsl_fall_world.h, also compiles for the host, which is how the CPU test tier checks it).  The target angles, rates and feed-forward torques come from the
// gait-evaluation kernels of this library (the hot path); what the reference takes from ODE is restated here:
// maximal coordinates (one rigid body per model body), hinge / fixed / contact constraint rows, projected Gauss-Seidel
// with over-relaxation 1.3, 20 iterations, lambda from 0, rows re-shuffled with ODE's LCG every eighth iteration, ERP 0.8,
// contact soft CFM 1e-3, bounce 0.5 above 0.1, unbounded friction rows (mu = infinity), gravity 1, semi-implicit Euler with
// ODE's infinitesimal quaternion update.  The CPU counterpart the tests compare with is the reference's own player code
// on oracle/shim/ode_step.cpp, which this file follows row for row.  What is approximated with respect to a real ODE
// build: the order in which ODE visits joints (island traversal of its intrusive lists), hence agreement with real ODE is
// statistical; trimesh terrain, motors / limits and joint feedback do not exist here.
//
// The thread-per-world kernel keeps a world's working set (~10 KB live: 22 bodies, rows in structured form -- anchor
// vectors instead of 12-wide Jacobian rows) in thread-local memory and is bound by the latency of that memory (1.3e6
// world-steps/s); the warp-per-world kernel has no such traffic and runs 2.5x faster in bulk, 30x faster on few worlds.
#include <cuda_runtime.h>
#include <stdint.h>

#include "hsl_fall_world.h"
#include "hsl_fall_warp.cuh"

namespace {

__global__ void __launch_bounds__(HSL_FALL_THREADS)
hsl_fall_kernel(const __grid_constant__ HslSimPod S, const __grid_constant__ HslFallArgs A) {
  const int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (wi >= A.n_worlds) return;
  hsl_fall::World w;   // ~21 KB of thread-local memory: the world's whole working set
  hsl_fall::fall_world(S, A, wi, w);
}

// q_cm: joint values component-major [config_dim][n_t+4] (frames 0..n_t+3), tau_cm [nmotor][n_t] (solved frames 2..n_t+1)
__global__ void hsl_fall_ctrl_kernel(int n_t, int nmotor, double dt, const double* __restrict__ q_cm, const double* __restrict__ tau_cm,
                                     double* __restrict__ ctrl) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_t * nmotor) return;
  const int tm = g / nmotor, j = g - tm * nmotor;
  hsl_fall::fall_ctrl_entry(n_t, nmotor, dt, q_cm, 1, n_t + 4, tau_cm, 1, n_t, tm, j, ctrl);
}

}  // namespace
cudaError_t hsl_launch_fall(const HslSimPod& S, const HslFallArgs& A, int variant, cudaStream_t st) {
  if (variant >= 1) {
    const int64_t blocks = (A.n_worlds + hsl_fall_warp::WARPS_PER_BLOCK - 1) / hsl_fall_warp::WARPS_PER_BLOCK;
    const unsigned th = 32 * hsl_fall_warp::WARPS_PER_BLOCK;
    hsl_fall_warp::hsl_fall_warp_kernel<<<(unsigned)blocks, th, 0, st>>>(S, A);
    return cudaGetLastError();
  }
  const int64_t blocks = (A.n_worlds + HSL_FALL_THREADS - 1) / HSL_FALL_THREADS;
  cudaError_t e = cudaFuncSetAttribute(hsl_fall_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 0);  // all of it to L1: local memory is the working set
  (void)e;
  hsl_fall_kernel<<<(unsigned)blocks, HSL_FALL_THREADS, 0, st>>>(S, A);
  return cudaGetLastError();
}
cudaError_t hsl_launch_fall_ctrl(int n_t, int nmotor, double dt, const double* q_cm, const double* tau_cm, double* ctrl, cudaStream_t st) {
  const int total = n_t * nmotor;
  hsl_fall_ctrl_kernel<<<(total + 127) / 128, 128, 0, st>>>(n_t, nmotor, dt, q_cm, tau_cm, ctrl);
  return cudaGetLastError();
}
