// sm_100a kernels of the gait-evaluation hot path.
//
//   hsl_setup_kernel      one thread per candidate: candidate constants + frame-time table   (a1, a4)
//   hsl_frames_kernel     one thread per (frame, role): phases A-E of hsl_frame.h             (a2-a14)
//   hsl_gait_pipe_kernel  persistent, software-pipelined form of the cost-only path (hsl_pipe.h); default for
//                         six-limbed models
//   hsl_forces_kernel     contact forces of all feet from given motor torques (hsl_forces.h)  (a15)
//   hsl_finish_kernel     one warp per candidate: work = sum_f (power_f * dt), COT, statistics (a14)
//   (selection after the all-gather of the costs: hsl_select.cu)
//
// Block layout of hsl_frames_kernel: FB frame slots x (NF+1) roles, thread = role*FB + slot, so a warp
// is one role over 32 consecutive frames of (normally) one candidate.  Blocks overlap by 4 slots: the
// outer two slots on each side only run phase A (they are the +-2 finite-difference halo of their
// neighbours) and are solved as interior slots of the adjacent block.  Model constants arrive as a
// __grid_constant__ parameter, i.e. in the constant bank, and are indexed by the warp-uniform role.
#include <cuda_runtime.h>

#include <cstdio>
#include <mutex>

#include "hsl_forces.h"
#include "hsl_frame.h"
#include "hsl_gather_dev.cuh"
#include "hsl_internal.h"
#include "hsl_pipe.h"

// Block-wide barrier usable from role-divergent code: bar.sync counts arriving warps, it does not care which
// instruction they arrive from (__syncthreads() may not be placed in divergent code).
#define HSL_BLOCK_SYNC() asm volatile("bar.sync 0;" ::: "memory")

template <int NF, int FB, int MODE, bool DUMP, int MAXREG, int AXP = HSL_AXP_GENERIC>
__global__ void __maxnreg__(MAXREG)
hsl_frames_kernel(const __grid_constant__ HslModelPod M, const __grid_constant__ HslFrameArgs A) {
  extern __shared__ __align__(16) double hsl_smem_raw[];
  HslSmem<NF, FB> sm;
  sm.carve(hsl_smem_raw, M.ntrunk);
  const int role = threadIdx.x / FB;
  HslSlot sl;
  sl.s = threadIdx.x % FB;
  HSL_GRID_DEP_WAIT();     // inputs written by the kernel in front (candidate setup, transposes of the device-pointer entries)
  HSL_GRID_DEP_LAUNCH();
  if (MODE == HSL_MODE_FIELDS) {
    const int64_t g = (int64_t)blockIdx.x * FB + sl.s;
    sl.valid = g < A.n_frames;
    sl.i = (int32_t)(sl.valid ? g : A.n_frames - 1);
    sl.c = sl.i;  // status is per frame in this mode
    sl.fo = sl.i;
    sl.interior = sl.valid;
  } else {
    const int per = A.n_t + 4;
    const int64_t g = (int64_t)blockIdx.x * (FB - 4) + sl.s;
    if (g < 0x7fffffffLL) {  // multiply-shift (exact for g < 2^31, see hsl_set_divider); the 64-bit division is a
      const uint32_t c32 = (uint32_t)(((uint64_t)(uint32_t)g * A.div_magic) >> A.div_shift);  // ~100-instruction routine
      sl.c = c32;
      sl.i = (int32_t)((uint32_t)g - c32 * (uint32_t)per);
    } else {
      sl.c = g / per;
      sl.i = (int32_t)(g - sl.c * per);
    }
    sl.valid = sl.c < A.n_cand;
    if (!sl.valid) { sl.c = A.n_cand - 1; sl.i = 0; }
    sl.interior = sl.valid && sl.s >= 2 && sl.s < FB - 2 && sl.i >= 2 && sl.i <= A.n_t + 1;
    sl.fo = sl.c * A.n_t + (sl.i - 2);
  }
#ifdef HSL_PHASE_CLOCKS
  long long clk[8];
  int nclk = 0;
#define HSL_STAMP() clk[nclk++] = clock64()
#else
#define HSL_STAMP()
#endif
  int bad = 0;
  // The two kinds of role run separate instruction streams that meet at the same four block barriers.  Whole warps
  // share a role (FB is a multiple of 32), so each warp executes every bar.sync exactly once per phase; keeping the
  // streams apart keeps the limbs' phase B -> D state out of the register allocation of the trunk's phase C.
  HSL_STAMP();
  if (role < NF) {
    HslLegState<DUMP> lst;
    phase_a_leg<NF, FB, MODE, DUMP, AXP>(M, A, sm, sl, role, lst);
    HSL_STAMP();
    HSL_BLOCK_SYNC();
    HSL_STAMP();
    if (sl.interior) phase_b_leg<NF, FB, MODE, DUMP>(M, A, sm, sl, role, lst);
    bad = lst.bad;
    HSL_STAMP();
    HSL_BLOCK_SYNC();
    HSL_STAMP();
    HSL_STAMP();
    HSL_BLOCK_SYNC();
    HSL_STAMP();
    if (sl.interior) phase_d_leg<NF, FB, MODE, DUMP>(M, A, sm, sl, role, lst);
    HSL_BLOCK_SYNC();
    HSL_STAMP();
  } else {
    HslTrunkState tst;
    phase_a_trunk<NF, FB, MODE>(M, A, sm, sl, tst);
    HSL_STAMP();
    HSL_BLOCK_SYNC();
    HSL_STAMP();
    if (sl.interior) phase_b_trunk<NF, FB, MODE>(M, A, sm, sl, tst);
    HSL_STAMP();
    HSL_BLOCK_SYNC();
    HSL_STAMP();
    if (sl.interior) bad |= phase_c_trunk<NF, FB, MODE, DUMP>(M, A, sm, sl, tst);
    HSL_STAMP();
    HSL_BLOCK_SYNC();
    HSL_STAMP();
    HSL_BLOCK_SYNC();
    if (sl.interior) phase_e_trunk<NF, FB>(A, sm, sl);
    HSL_STAMP();
  }
  if (bad && sl.valid && A.status) atomicOr(&A.status[sl.c], bad);
#ifdef HSL_PHASE_CLOCKS
  if (A.phase_clk && (threadIdx.x & 31) == 0) {
    long long* dst = A.phase_clk + ((size_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32) * 8;
    for (int k = 0; k < 8; k++) dst[k] = clk[k];
  }
#endif
}

// Contact forces of all feet from given motor torques (hsl_forces.h): phases A | B | F1 | F2 | F3, 32 frame slots.
template <int NF, int MODE>
__global__ void hsl_forces_kernel(const __grid_constant__ HslModelPod M, const __grid_constant__ HslFrameArgs A) {
  constexpr int FB = 32;
  extern __shared__ __align__(16) double hsl_smem_raw[];
  HslSmem<NF, FB, HSL_FORCES_PART> sm;
  sm.carve(hsl_smem_raw, M.ntrunk);
  const int role = threadIdx.x / FB;
  HslSlot sl;
  sl.s = threadIdx.x % FB;
  if (MODE == HSL_MODE_FIELDS) {
    const int64_t g = (int64_t)blockIdx.x * FB + sl.s;
    sl.valid = g < A.n_frames;
    sl.i = (int32_t)(sl.valid ? g : A.n_frames - 1);
    sl.c = sl.i;
    sl.fo = sl.i;
    sl.interior = sl.valid;
  } else {
    const int per = A.n_t + 4;
    const int64_t g = (int64_t)blockIdx.x * (FB - 4) + sl.s;
    sl.c = g / per;
    sl.i = (int32_t)(g - sl.c * per);
    sl.valid = sl.c < A.n_cand;
    if (!sl.valid) { sl.c = A.n_cand - 1; sl.i = 0; }
    sl.interior = sl.valid && sl.s >= 2 && sl.s < FB - 2 && sl.i >= 2 && sl.i <= A.n_t + 1;
    sl.fo = sl.c * A.n_t + (sl.i - 2);
  }
  HslLegState<true> lst;
  HslForcesLeg fst;
  HslTrunkState tst;
  int bad = 0;
  if (role < NF) {
    phase_a_leg<NF, FB, MODE, true>(M, A, sm, sl, role, lst);
    bad = lst.bad;
  } else {
    phase_a_trunk<NF, FB, MODE>(M, A, sm, sl, tst);
  }
  __syncthreads();
  if (sl.interior) {
    if (role < NF) {
      phase_b_leg<NF, FB, MODE, true>(M, A, sm, sl, role, lst);
      lst.bad = 0;  // level-1 blocks of phase B are not used here
      forces_f1_leg<NF, FB, MODE, true>(M, A, sm, sl, role, lst, fst);
      bad |= lst.bad;
    } else {
      phase_b_trunk<NF, FB, MODE>(M, A, sm, sl, tst);
    }
  }
  __syncthreads();
  if (sl.interior && role == NF) bad |= forces_f2_trunk<NF, FB, MODE>(M, A, sm, sl, tst);
  __syncthreads();
  if (sl.interior && role < NF) forces_f3_leg<NF, FB, MODE, true>(M, A, sm, sl, role, lst, fst);
  if (bad && sl.valid && A.status) atomicOr(&A.status[sl.c], bad);
}

// Persistent, software-pipelined cost-only kernel (schedule and buffer hand-offs in hsl_pipe.h).
// Slot of a thread: tile t covers the linearised (candidate, frame) indices t (FB-4) + s.  Walked incrementally: a
// block's tiles are gridDim.x tiles apart, so (candidate, frame) advance by a fixed quotient / remainder per
// iteration -- two additions and a compare instead of a division at the head of the dependency chain of every
// phase A.
template <int NF, int FB>
struct HslPipeSlotIter {
  int64_t c, dq;
  int32_t i, dr, per, s;
  __device__ __forceinline__ void init(const HslFrameArgs& A, int s_) {
    s = s_;
    per = A.n_t + 4;
    const int64_t g = (int64_t)blockIdx.x * (FB - 4) + s, step = (int64_t)gridDim.x * (FB - 4);
    c = g / per; i = (int32_t)(g - c * per);
    dq = step / per; dr = (int32_t)(step - dq * per);
  }
  __device__ __forceinline__ HslSlot get(const HslFrameArgs& A) const {
    HslSlot sl;
    sl.s = s; sl.c = c; sl.i = i;
    sl.valid = c < A.n_cand;
    if (!sl.valid) { sl.c = A.n_cand - 1; sl.i = 0; }
    sl.interior = sl.valid && s >= 2 && s < FB - 2 && sl.i >= 2 && sl.i <= A.n_t + 1;
    sl.fo = sl.c * A.n_t + (sl.i - 2);
    return sl;
  }
  __device__ __forceinline__ void next() {
    i += dr; c += dq;
    if (i >= per) { i -= per; c += 1; }
  }
};

template <int NF, int FB, int AXP = HSL_AXP_GENERIC>
__global__ void __maxnreg__(128)
hsl_gait_pipe_kernel(const __grid_constant__ HslModelPod M, const __grid_constant__ HslFrameArgs A, const int64_t n_tiles) {
  extern __shared__ __align__(16) double hsl_smem_raw[];
  HslPipeSmem<NF, FB> sm;
  sm.carve(hsl_smem_raw, M.ntrunk);
  const int role = threadIdx.x / FB;
  const int s = threadIdx.x % FB;
  HSL_GRID_DEP_WAIT();     // the candidate constants and frame times of the setup kernel are complete
  HSL_GRID_DEP_LAUNCH();   // the finish kernel may take the SMs this grid frees at its tail
  // p1 / p2 below: slots of the two previous tiles of this block (t-1, t-2)
#ifdef HSL_PHASE_CLOCKS
  long long acc[4] = {0, 0, 0, 0};
  long long t0c = 0, t1c = 0;
  bool first_tile = true;
#define HSL_CLK(x) asm volatile("mov.u64 %0, %%clock64;" : "=l"(x)::"memory")
#define HSL_T0() do { if (first_tile) HSL_CLK(t0c); } while (0)
#define HSL_T1(k) do { HSL_CLK(t1c); acc[k] += t1c - t0c; t0c = t1c; } while (0)
#else
#define HSL_T0()
#define HSL_T1(k)
#endif
  // Limb and trunk warps run separate loops that meet at the same two block barriers per tile (see
  // HSL_BLOCK_SYNC): neither role's live state enters the other's register allocation.
  if (role < NF) {
    bool p1_int = false;
    HslPipeSlotIter<NF, FB> it;
    it.init(A, s);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it.next()) {
      const HslSlot sl = it.get(A);
      HslLegState<false> lst;
      HSL_T0();
      phase_a_leg<NF, FB, HSL_MODE_GAIT, false, AXP>(M, A, sm, sl, role, lst);
      int bad = lst.bad;
      HSL_T1(0);
      HSL_BLOCK_SYNC();
      HSL_T1(1);
      if (p1_int) pipe_d_leg<NF, FB>(sm, s, role);
      if (sl.interior) {
        phase_b_leg<NF, FB, HSL_MODE_GAIT, false>(M, A, sm, sl, role, lst);
        pipe_store_dstate<NF, FB>(sm, s, role, lst);
        bad |= lst.bad;
      }
      if (bad && sl.valid && A.status) atomicOr(&A.status[sl.c], bad);
      HSL_T1(2);
      HSL_BLOCK_SYNC();
      HSL_T1(3);
#ifdef HSL_PHASE_CLOCKS
      first_tile = false;
#endif
      p1_int = sl.interior;
    }
    // drain: [trunk E(T-1), C(T)] | D(T) | [trunk E(T)]
    HSL_BLOCK_SYNC();
    if (p1_int) pipe_d_leg<NF, FB>(sm, s, role);
    HSL_BLOCK_SYNC();
  } else if (role == NF) {
    // solver warps: phase C of the previous tile in the first half, nothing in the second
    bool p1_int = false;
    int64_t p1_fo = 0, p1_c = 0;
    auto solve = [&]() {
      HslTrunkState tst;
#pragma unroll
      for (int k = 0; k < 3; k++) { tst.F0[k] = sm.twr[k * FB + s]; tst.T0[k] = sm.twr[(3 + k) * FB + s]; }
      HslSlot ps;
      ps.s = s; ps.fo = p1_fo; ps.c = p1_c; ps.i = 2; ps.valid = true; ps.interior = true;
      const int tb = phase_c_trunk<NF, FB, HSL_MODE_GAIT, false>(M, A, sm, ps, tst);
      if (tb && A.status) atomicOr(&A.status[p1_c], tb);
    };
    HslPipeSlotIter<NF, FB> it;
    it.init(A, s);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it.next()) {
      const HslSlot sl = it.get(A);
      HSL_T0();
      if (p1_int) solve();
      HSL_T1(0);
      HSL_BLOCK_SYNC();
      HSL_T1(1);
      HSL_T1(2);
      HSL_BLOCK_SYNC();
      HSL_T1(3);
#ifdef HSL_PHASE_CLOCKS
      first_tile = false;
#endif
      p1_int = sl.interior; p1_fo = sl.fo; p1_c = sl.c;
    }
    if (p1_int) solve();
    HSL_BLOCK_SYNC();
    HSL_BLOCK_SYNC();
  } else {
    // trunk warps: output pass of tile t-2 and kinematics of tile t in the first half, derivatives in the second
    bool p1_int = false, p2_int = false;
    int64_t p1_fo = 0, p2_fo = 0;
    HslTrunkState tst;
    HslPipeSlotIter<NF, FB> it;
    it.init(A, s);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it.next()) {
      const HslSlot sl = it.get(A);
      HSL_T0();
      if (p2_int) pipe_e_trunk<NF, FB>(A, sm, s, p2_fo);
      phase_a_trunk<NF, FB, HSL_MODE_GAIT>(M, A, sm, sl, tst);
      HSL_T1(0);
      HSL_BLOCK_SYNC();
      HSL_T1(1);
      if (sl.interior) {
        phase_b_trunk<NF, FB, HSL_MODE_GAIT>(M, A, sm, sl, tst);
#pragma unroll
        for (int k = 0; k < 3; k++) { sm.twr[k * FB + s] = tst.F0[k]; sm.twr[(3 + k) * FB + s] = tst.T0[k]; }
      }
      HSL_T1(2);
      HSL_BLOCK_SYNC();
      HSL_T1(3);
#ifdef HSL_PHASE_CLOCKS
      first_tile = false;
#endif
      p2_int = p1_int; p2_fo = p1_fo;
      p1_int = sl.interior; p1_fo = sl.fo;
    }
    if (p2_int) pipe_e_trunk<NF, FB>(A, sm, s, p2_fo);
    HSL_BLOCK_SYNC();
    HSL_BLOCK_SYNC();
    if (p1_int) pipe_e_trunk<NF, FB>(A, sm, s, p1_fo);
  }
#ifdef HSL_PHASE_CLOCKS
  if (A.phase_clk && (threadIdx.x & 31) == 0) {
    long long* dst = A.phase_clk + ((size_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32) * 8;
    for (int k = 0; k < 4; k++) dst[k] = acc[k];
    dst[4] = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  }
#endif
}

// Record-level entries: one thread per (record, role), role = limb 0..nf-1 | torso.
// rec [C][n_times][6+3nf] for candidates cand[C] at times[n_times]               (pergensetup::set_rec)
__global__ void hsl_gait_records_kernel(const __grid_constant__ HslModelPod M, const __grid_constant__ HslFrameArgs A, int n_times,
                                        const double* __restrict__ times, double* __restrict__ rec) {
  const int roles = M.nf + 1, rl = 6 + 3 * M.nf;
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = g / roles;
  if (r >= A.n_cand * n_times) return;
  const int role = (int)(g - r * roles);
  const int64_t c = r / n_times;
  gait_record(A, A.cand[c], M.nf, role, times[r - c * n_times], rec + r * rl);
}
// q [n][config_dim] for records rec [n][6+3nf]; status [n] gets HSL_ST_UNREACHABLE    (kinematicmodel::set_jvalues_with_lik)
__global__ void hsl_ik_records_kernel(const __grid_constant__ HslModelPod M, int64_t n, int flags, const double* __restrict__ rec,
                                      double* __restrict__ q, int32_t* __restrict__ status) {
  const int roles = M.nf + 1, rl = 6 + 3 * M.nf;
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = g / roles;
  if (r >= n) return;
  const int role = (int)(g - r * roles);
  if (!ik_record(M, role, rec + r * rl, (flags & HSL_FLAG_IGNORE_REACH) != 0, q + r * M.config_dim)) atomicOr(&status[r], HSL_ST_UNREACHABLE);
}

// A [n][nb][16], J [n][nb][16] for joint values q [n][config_dim]      (kinematicmodel::set_jvalues + recompute_modelnodes)
__global__ void hsl_fk_records_kernel(const __grid_constant__ HslModelPod M, int64_t n, const double* __restrict__ q,
                                      double* __restrict__ Aout, double* __restrict__ Jout) {
  const int roles = M.nf + 1;
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = g / roles;
  if (r >= n) return;
  const int role = (int)(g - r * roles);
  fk_record(M, role, q + r * M.config_dim, Aout ? Aout + r * M.n * 16 : nullptr, Jout ? Jout + r * M.n * 16 : nullptr);
}

__global__ void hsl_setup_kernel(const __grid_constant__ HslModelPod M, int64_t n_cand, int n_t, const double* __restrict__ params,
                                 HslCand* __restrict__ cand, double* __restrict__ ttab, int32_t* __restrict__ status) {
  // Two warps per block, lane = candidate (32 candidates per block).  Warp 0 computes the candidate constants.  Warp 1 writes
  // the frame times: an accumulated sum (t += dt, periodic.cpp:87-91), sequential per candidate; written lane by lane they
  // would be 32 scattered 8-byte stores per step, so the warp stages 32 steps of its 32 candidates in shared memory and
  // writes whole 256-byte rows.  The two run side by side (the kernel is a latency chain of one warp per SM otherwise).
  __shared__ double tile[32][33];
  HSL_GRID_DEP_LAUNCH();   // the per-frame kernel may start its prologue (it waits for this grid's end)
  const int lane = threadIdx.x & 31;
  const int64_t c0 = (int64_t)blockIdx.x * 32, c = c0 + lane;
  const bool valid = c < n_cand;
  if (threadIdx.x < 32) {
    if (valid) {
      double p[HSL_NPARAM];
#pragma unroll
      for (int k = 0; k < HSL_NPARAM; k++) p[k] = params[c * HSL_NPARAM + k];
      HslCand cd;
      setup_candidate(M, p, n_t, cd, nullptr);
      cand[c] = cd;
      status[c] = cd.status;
    }
    return;
  }
  const double dt = valid ? params[c * HSL_NPARAM + 7] / n_t : 0.0;   // cd.dt = period / n_t, as setup_candidate computes it
  const int per = n_t + 4;
  const int rows = (int)((n_cand - c0 < 32) ? n_cand - c0 : 32);
  double t = 0;
  for (int i0 = 0; i0 < per; i0 += 32) {
#pragma unroll
    for (int j = 0; j < 32; j++) { tile[lane][j] = t; t += dt; }   // the same recurrence as setup_candidate's own table
    __syncwarp();
    if (i0 + lane < per) {
      double* dst = ttab + c0 * per + i0 + lane;
      if (rows == 32) {
#pragma unroll
        for (int r = 0; r < 32; r++) dst[(int64_t)r * per] = tile[r][lane];
      } else {
        for (int r = 0; r < rows; r++) dst[(int64_t)r * per] = tile[r][lane];
      }
    }
    __syncwarp();
  }
}

// One warp per candidate.  periodic::work_over_period (periodic.cpp:285-307) and modelplayer::measure_cot
// (player.cpp:269-285): work = sum_frames (sum_motors max(tau*qdot,0)) * dt ; COT = work / (total mass * step length).
__global__ void hsl_finish_kernel(int64_t n_cand, int n_t, double total_mass, const HslCand* __restrict__ cand,
                                  const double* __restrict__ dt_in, const double* __restrict__ wframe,
                                  const double* __restrict__ fmin_in, const double* __restrict__ fmax_in,
                                  const int32_t* __restrict__ status, double* __restrict__ cot, double* __restrict__ work,
                                  double* __restrict__ min_cfz, double* __restrict__ max_mu, const __grid_constant__ HslPeerOut peers) {
  HSL_GRID_DEP_WAIT();     // launched with programmatic dependent launch behind the per-frame kernel
  HSL_GRID_DEP_LAUNCH();
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
  const int lane = threadIdx.x & 31;
  if (c >= n_cand) {   // a whole warp
    if (peers.n && peers.signal) hsl_gather_publish(peers);
    return;
  }
  const double dt = cand ? cand[c].dt : dt_in[c];
  double w = 0, mn = 1e10, mx = -1e10;  // periodic.cpp:380
  // four strides of loads in flight per lane (the kernel is a chain of L2 round trips otherwise); accumulated in the same order
  for (int f0 = lane; f0 < n_t; f0 += 128) {
    double a[4], b[4], d[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int f = f0 + 32 * u;
      const bool in = f < n_t;
      a[u] = in ? wframe[c * n_t + f] : 0.0;
      b[u] = in ? fmin_in[c * n_t + f] : 1e10;
      d[u] = in ? fmax_in[c * n_t + f] : -1e10;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (f0 + 32 * u < n_t) { w += a[u] * dt; mn = fmin(mn, b[u]); mx = fmax(mx, d[u]); }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    w += __shfl_xor_sync(0xffffffffu, w, o);
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  const int st = status ? status[c] : 0;
  const bool fatal = (st & (HSL_ST_BAD_PARAMS | HSL_ST_UNREACHABLE)) != 0;  // the reference exit(1)s here
  const double nanv = __longlong_as_double(0x7ff8000000000000LL);
  const double cv = (fatal || !cand) ? nanv : w / (total_mass * cand[c].step_length);
  if (lane == 0) {
    if (work) work[c] = fatal ? nanv : w;
    if (cot) cot[c] = cv;
    if (min_cfz) min_cfz[c] = fatal ? nanv : mn;
    if (max_mu) max_mu[c] = fatal ? nanv : mx;
  }
  // the all-gather of the costs, fused: lane r stores this candidate's cost and status into rank r's gather buffer (peer
  // memory over NVLink; the rank's own buffer among them); the last block to finish raises this rank's flag at every peer.
  if (lane < peers.n) { peers.cot[lane][c] = cv; peers.status[lane][c] = st; }
  if (peers.n && peers.signal) hsl_gather_publish(peers);
}

// Selection (argmin / top-k over the all-gathered costs) lives in hsl_select.cu.

// FP64 FMA throughput probe: register-resident dependent chains, 8 per thread.  Used by bench.py for the
// roofline denominator of this FP64-bound path (MEASURED_PEAKS.json has no FP64 figure).
__global__ void hsl_dfma_probe_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// Accuracy self-test of the branch-free primitives of hsl_fastmath.h against the IEEE / CUDA library operations.
__global__ void hsl_math_selftest_kernel(int n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = hsl_div(a[i], b[i]);
  out[n + i] = a[i] / b[i];
  out[2 * n + i] = hsl_sqrt(fabs(a[i]));
  out[3 * n + i] = sqrt(fabs(a[i]));
  out[4 * n + i] = hsl_atan2(a[i], b[i]);
  out[5 * n + i] = atan2(a[i], b[i]);
  double sn, cs, sr, cr;
  hsl_sincos_0_pi(fabs(a[i]), &sn, &cs);
  sincos(fabs(a[i]), &sr, &cr);
  out[6 * n + i] = sn; out[7 * n + i] = sr; out[8 * n + i] = cs; out[9 * n + i] = cr;
}

// Per-frame dumps leave the frame kernels component-major ([comp][frame]: the lanes of a warp are consecutive frames
// and write consecutive addresses); the reference-facing arrays are row-major per frame ([frame][comp]).  This tiled
// transpose turns one into the other at HBM speed on the device, so the host side is one plain D2H copy: 32 x 32
// tiles through shared memory (33-column padding: no bank conflicts), both the read and the write coalesced.
template <typename T>
__global__ void hsl_transpose_kernel(const T* __restrict__ src, T* __restrict__ dst, int comps, int64_t nfr) {
  __shared__ T tile[32][33];
  const int64_t f0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r;
    const int64_t f = f0 + threadIdx.x;
    if (c < comps && f < nfr) tile[r][threadIdx.x] = src[(int64_t)c * nfr + f];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t f = f0 + r;
    const int c = c0 + threadIdx.x;
    if (c < comps && f < nfr) dst[f * comps + c] = tile[threadIdx.x][r];
  }
}

// ------------------------------------------------------------------ launchers
namespace {
// floor(g / d) = (g * magic) >> shift for all 0 <= g < 2^31: shift = 31 + ceil(log2 d), magic = ceil(2^shift / d) < 2^32
void hsl_set_divider(HslFrameArgs& A) {
  const uint32_t d = (uint32_t)(A.n_t + 4);
  int l = 0;
  while ((1u << l) < d) l++;
  A.div_shift = 31 + l;
  A.div_magic = (uint32_t)((((uint64_t)1 << A.div_shift) + d - 1) / d);
}

template <int NF, int FB, int MODE, bool DUMP, int MAXREG = 255, int AXP = HSL_AXP_GENERIC>
cudaError_t launch_frames_t(const HslModelPod& M, const HslFrameArgs& A_in, cudaStream_t st) {
  HslFrameArgs A = A_in;
  hsl_set_divider(A);
  const size_t smem = (size_t)HslSmem<NF, FB>::doubles_per_slot(M.ntrunk) * FB * sizeof(double);
  auto kern = hsl_frames_kernel<NF, FB, MODE, DUMP, MAXREG, AXP>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int64_t blocks;
  if (MODE == HSL_MODE_FIELDS) blocks = (A.n_frames + FB - 1) / FB;
  else {
    const int64_t slots = A.n_cand * (A.n_t + 4);
    blocks = (slots - 4 + (FB - 4) - 1) / (FB - 4);  // interior ranges [b(FB-4)+2, b(FB-4)+FB-2) must cover [2, slots-2)
    if (blocks < 1) blocks = 1;
  }
  if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  HslPdlConfig pc(dim3((unsigned)blocks), dim3((NF + 1) * FB), smem, st);   // programmatic dependent launch, as for the pipelined kernel
  return cudaLaunchKernelEx(&pc.cfg, kern, M, A);
}
template <int NF, int FB>
cudaError_t launch_frames_nf(const HslModelPod& M, const HslFrameArgs& A, int mode, bool dump, cudaStream_t st) {
  (void)dump;
  if (mode == HSL_MODE_GAIT) return launch_frames_t<NF, FB, HSL_MODE_GAIT, true>(M, A, st);
  if (mode == HSL_MODE_TRAJ) return launch_frames_t<NF, FB, HSL_MODE_TRAJ, true>(M, A, st);
  return launch_frames_t<NF, FB, HSL_MODE_FIELDS, true>(M, A, st);
}
}  // namespace

template <int NF, int FB, int AXP = HSL_AXP_GENERIC>
cudaError_t launch_gait_pipe(const HslModelPod& M, const HslFrameArgs& A, cudaStream_t st) {
  const size_t smem = (size_t)HslPipeSmem<NF, FB>::doubles_per_slot(M.ntrunk) * FB * sizeof(double);
  auto kern = hsl_gait_pipe_kernel<NF, FB, AXP>;
  cudaError_t e = cudaSuccess;
  const int64_t slots = A.n_cand * (A.n_t + 4);
  int64_t n_tiles = (slots - 4 + (FB - 4) - 1) / (FB - 4);
  if (n_tiles < 1) n_tiles = 1;
  // persistent grid = SMs x co-resident blocks of THIS kernel on THE CURRENT device: queried once per (kernel instance, device)
  struct DevInfo { int sms, per_sm; size_t smem; };
  static DevInfo info[64];          // zero-initialised; slot = device ordinal
  static std::mutex info_mu;
  int dev = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  DevInfo di = {0, 0, 0};
  {
    std::lock_guard<std::mutex> lk(info_mu);
    if (dev >= 0 && dev < 64) di = info[dev];
    if (di.sms == 0 || di.smem != smem) {  // first use on this device, or a model with another trunk-body count
      di.smem = smem;
      if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
      cudaDeviceGetAttribute(&di.sms, cudaDevAttrMultiProcessorCount, dev);
      if (di.sms <= 0) di.sms = 148;
      if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&di.per_sm, kern, HSL_PIPE_ROLES(NF) * FB, smem)) != cudaSuccess) return e;
      if (di.per_sm < 1) return cudaErrorInvalidConfiguration;
      if (dev >= 0 && dev < 64) info[dev] = di;
    }
  }
  int64_t grid = (int64_t)di.sms * di.per_sm;
  if (grid > n_tiles) grid = n_tiles;
  // Programmatic dependent launch: the blocks of this kernel become resident and run their prologue (shared-memory carve-up,
  // slot arithmetic) while the candidate-setup kernel in front of it is still finishing; they wait for its results at
  // griddepcontrol.wait.  In front of any other kernel the wait is an ordinary stream dependency.
  HslPdlConfig pc(dim3((unsigned)grid), dim3(HSL_PIPE_ROLES(NF) * FB), smem, st);
  return cudaLaunchKernelEx(&pc.cfg, kern, M, A, (int64_t)n_tiles);
}

// Cost-only evaluation (the headline path) comes in a few occupancy variants: fb = frame slots per block,
// maxreg = register cap per thread.  The register file is 16K registers per SM sub-partition and warps are dealt
// round-robin to the 4 sub-partitions, so ceil(resident warps / 4) * 32 * maxreg must stay <= 16384:
// 4 warps per sub-partition -> 128 registers, 3 -> 168, 2 -> 255.
template <int NF>
cudaError_t launch_gait_fast(const HslModelPod& M, const HslFrameArgs& A, int fb, int maxreg, cudaStream_t st) {
  // the default variants (six limbs: pipelined, 64 slots; four limbs: plain, 32 slots, 128 registers) also exist
  // specialised for the model's hinge-axis pattern
  const int axp = hsl_axis_pattern(M);
  if (maxreg == 1) {  // software-pipelined persistent kernel
    if (fb == 64) {
      if (NF == 6 && axp == HSL_AXP_YXX) return launch_gait_pipe<NF, 64, (NF == 6 ? HSL_AXP_YXX : HSL_AXP_GENERIC)>(M, A, st);
      if (NF == 6 && axp == HSL_AXP_ZXX) return launch_gait_pipe<NF, 64, (NF == 6 ? HSL_AXP_ZXX : HSL_AXP_GENERIC)>(M, A, st);
      return launch_gait_pipe<NF, 64>(M, A, st);
    }
    return launch_gait_pipe<NF, 32>(M, A, st);
  }
  if (fb == 64) {  // 14 warps (nf=6): 4 per sub-partition
    if (NF == 6 && axp == HSL_AXP_YXX) return launch_frames_t<NF, 64, HSL_MODE_GAIT, false, 128, (NF == 6 ? HSL_AXP_YXX : HSL_AXP_GENERIC)>(M, A, st);
    return launch_frames_t<NF, 64, HSL_MODE_GAIT, false, 128>(M, A, st);
  }
  if (NF == 4 && maxreg > 96 && maxreg <= 128 && axp == HSL_AXP_YXX)
    return launch_frames_t<NF, 32, HSL_MODE_GAIT, false, 128, (NF == 4 ? HSL_AXP_YXX : HSL_AXP_GENERIC)>(M, A, st);
  if (maxreg <= 96) return launch_frames_t<NF, 32, HSL_MODE_GAIT, false, 80>(M, A, st);   // 3 blocks x 7 warps
  if (maxreg <= 128) return launch_frames_t<NF, 32, HSL_MODE_GAIT, false, 128>(M, A, st); // 2 blocks x 7 warps
  return launch_frames_t<NF, 32, HSL_MODE_GAIT, false, 255>(M, A, st);                    // 1 block x 7 warps
}

cudaError_t hsl_launch_frames(const HslModelPod& M, const HslFrameArgs& A, int mode, bool dump, int fb, int maxreg, cudaStream_t st) {
  if (M.nf == 6) {
    if (mode == HSL_MODE_GAIT && !dump) return launch_gait_fast<6>(M, A, fb, maxreg, st);
    return launch_frames_nf<6, 32>(M, A, mode, dump, st);
  }
  if (M.nf == 4) {
    if (mode == HSL_MODE_GAIT && !dump) return launch_gait_fast<4>(M, A, fb, maxreg, st);
    return launch_frames_nf<4, 32>(M, A, mode, dump, st);
  }
  return cudaErrorInvalidValue;
}

namespace {
template <int NF, int MODE>
cudaError_t launch_forces_t(const HslModelPod& M, const HslFrameArgs& A, cudaStream_t st) {
  constexpr int FB = 32;
  const size_t smem = (size_t)HslSmem<NF, FB, HSL_FORCES_PART>::doubles_per_slot(M.ntrunk) * FB * sizeof(double);
  auto kern = hsl_forces_kernel<NF, MODE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int64_t blocks;
  if (MODE == HSL_MODE_FIELDS) blocks = (A.n_frames + FB - 1) / FB;
  else {
    blocks = (A.n_cand * (A.n_t + 4) - 4 + (FB - 4) - 1) / (FB - 4);
    if (blocks < 1) blocks = 1;
  }
  if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  kern<<<(unsigned)blocks, (NF + 1) * FB, smem, st>>>(M, A);
  return cudaGetLastError();
}
}  // namespace

cudaError_t hsl_launch_forces(const HslModelPod& M, const HslFrameArgs& A, int mode, cudaStream_t st) {
  if (mode != HSL_MODE_GAIT && mode != HSL_MODE_FIELDS) return cudaErrorInvalidValue;
  if (M.nf == 6) return mode == HSL_MODE_GAIT ? launch_forces_t<6, HSL_MODE_GAIT>(M, A, st) : launch_forces_t<6, HSL_MODE_FIELDS>(M, A, st);
  if (M.nf == 4) return mode == HSL_MODE_GAIT ? launch_forces_t<4, HSL_MODE_GAIT>(M, A, st) : launch_forces_t<4, HSL_MODE_FIELDS>(M, A, st);
  return cudaErrorInvalidValue;
}

cudaError_t hsl_launch_fk_records(const HslModelPod& M, int64_t n, const double* q, double* Aout, double* Jout, cudaStream_t st) {
  const int64_t threads = n * (M.nf + 1);
  hsl_fk_records_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(M, n, q, Aout, Jout);
  return cudaGetLastError();
}
cudaError_t hsl_launch_gait_records(const HslModelPod& M, const HslFrameArgs& A, int n_times, const double* times, double* rec,
                                    cudaStream_t st) {
  const int64_t threads = A.n_cand * n_times * (M.nf + 1);
  hsl_gait_records_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(M, A, n_times, times, rec);
  return cudaGetLastError();
}
cudaError_t hsl_launch_ik_records(const HslModelPod& M, int64_t n, int flags, const double* rec, double* q, int32_t* status,
                                  cudaStream_t st) {
  const int64_t threads = n * (M.nf + 1);
  hsl_ik_records_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(M, n, flags, rec, q, status);
  return cudaGetLastError();
}

cudaError_t hsl_launch_setup(const HslModelPod& M, int64_t n_cand, int n_t, const double* params, HslCand* cand, double* ttab,
                             int32_t* status, cudaStream_t st) {
  // 32 candidates per block (one per lane, sequential time tables): spread over many SMs; two warps: constants | time table
  hsl_setup_kernel<<<(unsigned)((n_cand + 31) / 32), 64, 0, st>>>(M, n_cand, n_t, params, cand, ttab, status);
  return cudaGetLastError();
}

cudaError_t hsl_launch_finish(int64_t n_cand, int n_t, double total_mass, const HslCand* cand, const double* dt_in,
                              const double* wframe, const double* fmin_in, const double* fmax_in, const int32_t* status, double* cot,
                              double* work, double* min_cfz, double* max_mu, cudaStream_t st, const HslPeerOut* peers) {
  const int tpb = 256;
  const int64_t threads = n_cand * 32;
  HslPeerOut none;
  none.n = 0;
  HslPdlConfig pc(dim3((unsigned)((threads + tpb - 1) / tpb)), dim3(tpb), 0, st);
  return cudaLaunchKernelEx(&pc.cfg, hsl_finish_kernel, n_cand, n_t, total_mass, cand, dt_in, wframe, fmin_in, fmax_in, status, cot, work, min_cfz,
                            max_mu, peers ? *peers : none);
}


cudaError_t hsl_launch_math_selftest(int n, const double* a, const double* b, double* out, cudaStream_t st) {
  hsl_math_selftest_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, a, b, out);
  return cudaGetLastError();
}

cudaError_t hsl_launch_dfma_probe(double* out, int blocks, int threads, int iters, cudaStream_t st) {
  hsl_dfma_probe_kernel<<<blocks, threads, 0, st>>>(out, iters, 0.999999, 1e-9);
  return cudaGetLastError();
}

cudaError_t hsl_launch_transpose(const void* src, void* dst, int comps, int64_t nfr, int elem_size, cudaStream_t st) {
  if (comps < 1 || nfr < 1) return cudaSuccess;
  const int64_t bx = (nfr + 31) / 32;
  if (bx > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  const dim3 grid((unsigned)bx, (unsigned)((comps + 31) / 32)), block(32, 8);
  if (elem_size == 8) hsl_transpose_kernel<double><<<grid, block, 0, st>>>((const double*)src, (double*)dst, comps, nfr);
  else if (elem_size == 1) hsl_transpose_kernel<uint8_t><<<grid, block, 0, st>>>((const uint8_t*)src, (uint8_t*)dst, comps, nfr);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

