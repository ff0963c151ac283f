// Software-pipelined variant of the cost-only gait evaluation (used by hsl_gait_pipe_kernel).
//
// In the plain kernel (hsl_frames_kernel) the limb warps wait while the trunk warp of each frame assembles and
// solves the 6x6 level-0 system, and the trunk warps wait during the limbs' kinematics: with one resident block per
// SM that serial section is ~30 % of a block's life.  Here a persistent block loops over tiles of FB frame slots and
// the roles are decoupled:
//
//   limb warps, tile t : phase A (gait, IK, FK, + COMs of the trunk bodies they are assigned)      -> barrier 1
//                        phase B (finite differences, Newton-Euler, 3x3 contact block, + the wrench of their trunk
//                        bodies) and publish EVERYTHING the per-frame finish needs                  -> barrier 2
//   trunk warps        : between barrier 2 of tile t-1 and barrier 1 of tile t (i.e. while the limbs run phase A of
//                        tile t): level-0 solve, contact forces, motor torques, positive power and contact statistics
//                        of tile t-1, written straight to global memory.
//
// The finish is possible on the trunk thread because, given the multiplier mu, every limb quantity is affine in it:
//   lambda_c = -(W_c g_c + W_c (mu_f + mu_t x r_c)),   tau_h = tau_p,h - w_h . lambda_c .
// The limb threads keep no state across tiles; the single `part` buffer is written by the limbs between barrier 1
// and barrier 2 and read by the trunk between barrier 2 and the next barrier 1.
#pragma once
#include "hsl_frame.h"

#define HSL_PPART 34  // part layout of the pipelined kernel, see hsl_frame.h

template <int NF, int FB>
using HslPipeSmem = HslSmem<NF, FB, HSL_PPART>;

// Limb thread, phase A extras: COM positions of the trunk bodies assigned to this limb (tb = limb, limb+NF, ...)
// and, for limb 0, the u*sin(theta) vector shared by all trunk bodies.
template <int NF, int FB, class SM>
HSL_HD void pipe_a_trunk_bodies(const HslModelPod& M, const HslFrameArgs& A, const SM& sm, const HslSlot& sl, int limb) {
  if (limb >= M.ntrunk) return;
  const HslCand& cd = A.cand[sl.c];
  const double t = A.ttab[sl.c * (A.n_t + 4) + sl.i];
  double qt[3], eul[3], R0[9], t0[3];
  if (torso_values(cd, t, qt, eul)) {
#pragma unroll
    for (int k = 0; k < 9; k++) R0[k] = cd.R0[k];
  } else {
    euler_to_R(eul[0], eul[1], eul[2], R0);
  }
  torso_frame(M, qt, R0, t0);
  for (int tb = limb; tb < M.ntrunk; tb += NF) {
    double ob[3], pb[3];
    m3_affine(R0, M.trunk[tb].off, t0, ob);
    m3_affine(R0, M.trunk[tb].com, ob, pb);
#pragma unroll
    for (int k = 0; k < 3; k++) sm.pos[((3 * NF + tb) * 3 + k) * FB + sl.s] = pb[k];
  }
  if (limb == 0) {
    sm.ust[((3 * NF) * 3 + 0) * FB + sl.s] = (R0[5] - R0[7]) / 2;
    sm.ust[((3 * NF) * 3 + 1) * FB + sl.s] = (R0[6] - R0[2]) / 2;
    sm.ust[((3 * NF) * 3 + 2) * FB + sl.s] = (R0[1] - R0[3]) / 2;
  }
}

// Limb thread, phase B extras (after phase_b_leg): add the wrench of the assigned trunk bodies to the published limb
// wrench and append tau_p, w, qdot to the partial.
template <int NF, int FB, class SM>
HSL_HD void pipe_b_extras(const HslModelPod& M, const HslFrameArgs& A, const SM& sm, const HslSlot& sl, int limb,
                          const HslLegState<false>& st) {
  double* P = sm.part + (limb * SM::PART) * FB + sl.s;
  if (limb < M.ntrunk) {
    const double hh = A.cand[sl.c].hh;
    double ref[3], F[3] = {0, 0, 0}, T[3] = {0, 0, 0};
    root_ref<NF, FB, HSL_MODE_GAIT>(M, A, sm, sl, ref);
    for (int tb = limb; tb < M.ntrunk; tb += NF) {
      double f[3], d[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        f[k] = fd2(sm.pos + ((3 * NF + tb) * 3 + k) * FB, sl.s, FB, hh, M.trunk[tb].mass);
        T[k] += fd2(sm.ust + ((3 * NF) * 3 + k) * FB, sl.s, FB, hh, M.trunk[tb].inertia);
        d[k] = sm.pos[((3 * NF + tb) * 3 + k) * FB + sl.s] - ref[k];
      }
      f[2] += M.trunk[tb].mass * M.g;
#pragma unroll
      for (int k = 0; k < 3; k++) F[k] += f[k];
      v3_cross_add(d, f, T);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) { P[k * FB] += F[k]; P[(3 + k) * FB] += T[k]; }
  }
#pragma unroll
  for (int h = 0; h < 3; h++) {
    P[(19 + h) * FB] = st.taup[h];
    P[(31 + h) * FB] = st.qd[h];
#pragma unroll
    for (int k = 0; k < 3; k++) P[(22 + 3 * h + k) * FB] = st.w[h][k];
  }
}

// Trunk thread: level-0 solve (as phase_c_trunk) + contact forces, motor torques, positive power, statistics of
// all limbs (as phase_d_leg / phase_e_trunk), from the published partials only.  Returns status bits.
template <int NF, int FB, class SM>
HSL_HD int pipe_trunk_finish(const HslFrameArgs& A, const SM& sm, const HslSlot& sl) {
  int bad = 0;
  double b[6] = {0, 0, 0, 0, 0, 0};
  double S[6][6];
#pragma unroll
  for (int i = 0; i < 6; i++)
#pragma unroll
    for (int j = 0; j < 6; j++) S[i][j] = 0;
  double v[6] = {0, 0, 0, 0, 0, 0};
  int nc = 0;
  double rA[3] = {0, 0, 0}, rB[3] = {0, 0, 0};
#pragma unroll
  for (int l = 0; l < NF; l++) {
    const double* P = sm.part + (l * SM::PART) * FB + sl.s;
#pragma unroll
    for (int k = 0; k < 6; k++) b[k] += P[k * FB];
    if (P[18 * FB] != 0.0) {
      double W[6], Wg[3], r[3];
#pragma unroll
      for (int k = 0; k < 6; k++) W[k] = P[(6 + k) * FB];
#pragma unroll
      for (int k = 0; k < 3; k++) { Wg[k] = P[(12 + k) * FB]; r[k] = P[(15 + k) * FB]; }
      if (nc == 0) { rA[0] = r[0]; rA[1] = r[1]; rA[2] = r[2]; }
      if (nc == 1) { rB[0] = r[0]; rB[1] = r[1]; rB[2] = r[2]; }
      nc++;
      const double Wc[3][3] = {{W[0], W[1], W[2]}, {W[1], W[3], W[4]}, {W[2], W[4], W[5]}};
      double K[3][3];
#pragma unroll
      for (int j = 0; j < 3; j++) {
        K[0][j] = r[1] * Wc[2][j] - r[2] * Wc[1][j];
        K[1][j] = r[2] * Wc[0][j] - r[0] * Wc[2][j];
        K[2][j] = r[0] * Wc[1][j] - r[1] * Wc[0][j];
      }
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) S[i][j] += Wc[i][j];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) S[3 + i][j] += K[i][j];
      S[3][3] += r[1] * K[0][2] - r[2] * K[0][1];
      S[4][3] += r[2] * K[0][0] - r[0] * K[0][2];
      S[5][3] += r[0] * K[0][1] - r[1] * K[0][0];
      S[4][4] += r[2] * K[1][0] - r[0] * K[1][2];
      S[5][4] += r[0] * K[1][1] - r[1] * K[1][0];
      S[5][5] += r[0] * K[2][1] - r[1] * K[2][0];
#pragma unroll
      for (int k = 0; k < 3; k++) v[k] += Wg[k];
      v3_cross_add(r, Wg, v + 3);
    }
  }
  double mu[6] = {0, 0, 0, 0, 0, 0};
  if (nc >= 2) {
    if (nc == 2) {
      double d[3] = {rA[0] - rB[0], rA[1] - rB[1], rA[2] - rB[2]}, nf_[3];
      v3_cross(rA, d, nf_);
      const double nn = hsl_rcp(hsl_sqrt(v3_dot(nf_, nf_) + v3_dot(d, d)));
      const double nv[6] = {nf_[0] * nn, nf_[1] * nn, nf_[2] * nn, d[0] * nn, d[1] * nn, d[2] * nn};
      double pb = 0;
#pragma unroll
      for (int k = 0; k < 6; k++) pb += nv[k] * b[k];
#pragma unroll
      for (int k = 0; k < 6; k++) mu[k] = -(b[k] - pb * nv[k] + v[k]);
#pragma unroll
      for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) S[i][j] += nv[i] * nv[j];
    } else {
#pragma unroll
      for (int k = 0; k < 6; k++) mu[k] = -(b[k] + v[k]);
    }
    if (!spd6_solve(S, mu)) bad |= HSL_ST_SOLVER;
  } else {
    bad |= HSL_ST_FEW_CONTACTS;
  }
  double work = 0, cfz = 1e300, mu_max = -1e300;
#pragma unroll
  for (int l = 0; l < NF; l++) {
    const double* P = sm.part + (l * SM::PART) * FB + sl.s;
    double lam[3] = {0, 0, 0};
    const bool con = (P[18 * FB] != 0.0) && (nc >= 2);
    if (con) {
      double W[6], r[3], y[3], Wy[3];
#pragma unroll
      for (int k = 0; k < 6; k++) W[k] = P[(6 + k) * FB];
#pragma unroll
      for (int k = 0; k < 3; k++) r[k] = P[(15 + k) * FB];
      v3_cross(mu + 3, r, y);
#pragma unroll
      for (int k = 0; k < 3; k++) y[k] += mu[k];
      sym3_mul(W, y, Wy);
#pragma unroll
      for (int k = 0; k < 3; k++) lam[k] = -(P[(12 + k) * FB] + Wy[k]);
      cfz = fmin(cfz, lam[2]);
      mu_max = fmax(mu_max, hsl_div(hsl_sqrt(lam[0] * lam[0] + lam[1] * lam[1]), lam[2]));
    }
    double wl = 0;
#pragma unroll
    for (int h = 0; h < 3; h++) {
      const double w0 = P[(22 + 3 * h) * FB], w1 = P[(23 + 3 * h) * FB], w2 = P[(24 + 3 * h) * FB];
      const double tau = P[(19 + h) * FB] - (w0 * lam[0] + w1 * lam[1] + w2 * lam[2]);
      const double dw = tau * P[(31 + h) * FB];
      wl += (dw > 0) ? dw : 0;
    }
    work += wl;
  }
  if (A.wframe) A.wframe[sl.fo] = work;
  if (A.fmin_cfz) A.fmin_cfz[sl.fo] = cfz;
  if (A.fmax_mu) A.fmax_mu[sl.fo] = mu_max;
  return bad;
}
