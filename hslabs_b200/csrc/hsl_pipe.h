// Software-pipelined variant of the cost-only gait evaluation (used by hsl_gait_pipe_kernel).
//
// In the plain kernel (hsl_frames_kernel) the limb warps idle while the trunk warps assemble and solve the 6x6
// level-0 system of each frame (phase C), and the trunk warps idle during the limbs' kinematics; with one resident
// block per SM those waits are ~1/4 of all warp time (ncu: "barrier" is the largest stall reason).  Here a
// persistent block loops over tiles of FB frame slots and the phases of consecutive tiles are skewed so that
// neither role waits for the other:
//
//               | first half of iteration t  (until barrier 1) | second half (until barrier 2)
//   limb warps  | A(t): gait, IK, FK                            | D(t-1): contact force, torques, power ; B(t)
//   solver warps| C(t-1): 6x6 solve                             | -
//   trunk warps | E(t-2) ; A'(t)                                | B'(t)
//
// The trunk's work is split over two roles (solver: phase C; trunk: kinematics, derivatives and the output pass) so
// that the block has 16 warps, four per scheduler: with 14 warps two schedulers hosted a trunk warp next to three limb
// warps (6300 instructions per tile against 4600 on the other two) and their limb warps set the pace of both halves.
// The trunk wrench F0/T0 goes from B'(t) to C(t) through `twr`.
//
// Phase D needs 24 doubles of limb state from phase B of the same frame; instead of keeping them in registers across
// a whole tile they are parked in shared memory: 12 in `dstate` (written at the end of B(t), read back in D(t) one
// iteration later, by the same thread) and W, W g, r in the limb's own `part` entries, which nobody overwrites
// before that same thread runs B(t+1) -- after D(t).  Buffers and their hand-offs (w = written, r = read):
//   pos/ust/cs  w A(t),A'(t) | r B(t),B'(t)            part   w B(t)      | r C(t)   (next first half)
//   mu          w C(t-1)     | r D(t-1) (second half)  dstate w end B(t)  | r D(t)   (same thread, next iteration)
//   fin         w D(t-1)     | r E(t-1) (next first half)
// Every writer/reader pair is separated by exactly one of the two block barriers per iteration.
#pragma once
#include "hsl_frame.h"

#define HSL_DSTATE 15  // per limb: tau_p*qd [3], w*qd [9], results of phase D [3]  (W, W g, r are re-read from `part`)

template <int NF, int FB>
struct HslPipeSmem : HslSmem<NF, FB, 19> {
  double* dstate;  // [NF*HSL_DSTATE][FB]
  double* twr;     // [6][FB]  trunk bodies' own share of the root wrench: w B'(t) (second half) | r C(t) (next first half)
  HSL_HD static int doubles_per_slot(int ntrunk) { return HslSmem<NF, FB, 19>::doubles_per_slot(ntrunk) + NF * HSL_DSTATE + 6; }
  HSL_HD void carve(double* base, int ntrunk) {
    HslSmem<NF, FB, 19>::carve(base, ntrunk);
    dstate = this->mu + 7 * FB;
    twr = dstate + NF * HSL_DSTATE * FB;
  }
};
#define HSL_PIPE_ROLES(NF) ((NF) + 2)  // limbs, solver, trunk

// Limb thread, end of phase B: park what phase D needs besides W, W g, r.
template <int NF, int FB, class SM>
HSL_HD void pipe_store_dstate(const SM& sm, int s, int limb, const HslLegState<false>& st) {
  double* D = sm.dstate + (limb * HSL_DSTATE) * FB + s;
#pragma unroll
  for (int h = 0; h < 3; h++) {
    D[h * FB] = st.taup[h] * st.qd[h];
#pragma unroll
    for (int k = 0; k < 3; k++) D[(3 + 3 * h + k) * FB] = st.w[h][k] * st.qd[h];
  }
}

// Limb thread, phase D of the previous tile (phase_d_leg of the plain kernel, from the parked state):
// lambda = -(W g + W (mu_f + mu_t x r)),  power = sum_h max((tau_p,h - w_h . lambda) qd_h, 0),  contact statistics.
template <int NF, int FB, class SM>
HSL_HD void pipe_d_leg(const SM& sm, int s, int limb) {
  double* D = sm.dstate + (limb * HSL_DSTATE) * FB + s;
  const double* Pl = sm.part + (limb * SM::PART) * FB;
  const double* P = Pl + s;
  const bool con = (P[18 * FB] != 0.0) && (sm.mu[6 * FB + s] != 0.0);
  double lam[3] = {0, 0, 0};
  if (con) {
    double W[6], Wg[3], r[3], mu[6], y[3], Wy[3];
#pragma unroll
    for (int k = 0; k < 6; k++) mu[k] = sm.mu[k * FB + s];
    {
      const HslD2 e3 = part_pair_load<FB>(Pl, s, 3), e4 = part_pair_load<FB>(Pl, s, 4), e5 = part_pair_load<FB>(Pl, s, 5);
      const HslD2 e6 = part_pair_load<FB>(Pl, s, 6), e7 = part_pair_load<FB>(Pl, s, 7), e8 = part_pair_load<FB>(Pl, s, 8);
      W[0] = e3.x; W[1] = e3.y; W[2] = e4.x; W[3] = e4.y; W[4] = e5.x; W[5] = e5.y;
      Wg[0] = e6.x; Wg[1] = e6.y; Wg[2] = e7.x; r[0] = e7.y; r[1] = e8.x; r[2] = e8.y;
    }
    v3_cross(mu + 3, r, y);
#pragma unroll
    for (int k = 0; k < 3; k++) y[k] += mu[k];
    sym3_mul(W, y, Wy);
#pragma unroll
    for (int k = 0; k < 3; k++) lam[k] = -(Wg[k] + Wy[k]);
  }
  double work = 0;
#pragma unroll
  for (int h = 0; h < 3; h++) {
    const double dw = D[h * FB] - (D[(3 + 3 * h) * FB] * lam[0] + D[(4 + 3 * h) * FB] * lam[1] + D[(5 + 3 * h) * FB] * lam[2]);
    work += (dw > 0) ? dw : 0;  // periodic.cpp:291-304
  }
  double cfz = 1e300, mu_f = -1e300;
  if (con) {  // periodic.cpp:347-357, over the feet that are on the ground
    cfz = lam[2];
    mu_f = hsl_div(hsl_sqrt(lam[0] * lam[0] + lam[1] * lam[1]), lam[2]);
  }
  D[12 * FB] = work;
  D[13 * FB] = cfz;
  D[14 * FB] = mu_f;
}

// Trunk thread, phase E: add the limbs' results of a frame and write them out.
template <int NF, int FB, class SM>
HSL_HD void pipe_e_trunk(const HslFrameArgs& A, const SM& sm, int s, int64_t fo) {
  double work = 0, cfz = 1e300, mu = -1e300;
#pragma unroll
  for (int l = 0; l < NF; l++) {
    const double* D = sm.dstate + (l * HSL_DSTATE) * FB + s;
    work += D[12 * FB];
    cfz = fmin(cfz, D[13 * FB]);
    mu = fmax(mu, D[14 * FB]);
  }
  if (A.wframe) A.wframe[fo] = work;
  if (A.fmin_cfz) A.fmin_cfz[fo] = cfz;
  if (A.fmax_mu) A.fmax_mu[fo] = mu;
}
