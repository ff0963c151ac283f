// Interface between the C-ABI layer and the fall / perturbation sweep kernels (hsl_fall.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "hsl_model.h"

#define HSL_FALL_THREADS 128
#define HSL_FALL_MAX_CONTACTS 16
#define HSL_FALL_ST_CONTACT_OVERFLOW 1  // more than HSL_FALL_MAX_CONTACTS bodies touched the ground in some step (extra ones ignored)

typedef struct HslFallArgs {
  int64_t n_worlds;
  int32_t n_steps, n_t, iterations, pad;
  double play_dt, play_t0;       // modelplayer::play_dt, play_t at the start (int(t0/dt + .5) * dt, player.cpp:376)
  double hc, tmin;               // fall_check: torso z < hc once play_t >= tmin (player.cpp:669-681)
  double erp, cfm, soft_cfm, bounce, bounce_vel, gravity;  // visualization.cpp:138-152, 296-325 and ODE's default world CFM
  double kp;                     // position-control gain k (player.cpp:394); damping 2 sqrt(k)
  const double* ctrl;            // [n_t][3][nmotor] target angle | target rate | feed-forward torque (hsl_fall_ctrl_kernel)
  const int32_t* kick_step;      // [W] step at which the torso is kicked (< 0: never); may be null
  const double* kick_dv;         // [W][3] velocity change of the kick; may be null
  double pos0[3 * HSL_MAX_BODIES], quat0[4 * HSL_MAX_BODIES];  // initial body poses (init_play_config, player.cpp:349-355); velocities 0
  uint8_t* fell;                 // [W]
  double* t_end;                 // [W] time of the fall, or the end time
  double* final_z;               // [W] torso COM height at the end
  int32_t* steps_done;           // [W]
  int32_t* status;               // [W] HSL_FALL_ST_*
  double* traj;                  // optional [W][n_steps][3] torso COM after every step
} HslFallArgs;

cudaError_t hsl_launch_fall(const HslSimPod& S, const HslFallArgs& A, int variant, cudaStream_t st);  // 1: a warp per world, 0: a thread per world
cudaError_t hsl_launch_fall_ctrl(int n_t, int nmotor, double dt, const double* q_cm, const double* tau_cm, double* ctrl, cudaStream_t st);
void hsl_sim_state_from_frames(const HslSimPod* sim, const double* A, double* pos, double* quat);
