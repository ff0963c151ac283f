// Contact forces of ALL feet from given motor torques: forcetorquesolver::solve_forces, ftsolver.cpp:331-378
// (periodic::solve_contforces_given_torques, periodic.cpp:369-374; used by modelplayer::test_dynamics,
// playerexperim.cpp:95-121).
//
// The reference appends one force column per foot and one row  a_h . T_h = tau_h  per hinge to the force-torque
// matrix, zeroes the six torso columns and takes the least-squares solution of the (6n + nmj) x (6n - 6 + 3 nf)
// system with a sparse QR.  Structured equivalent used here (proved against the oracle by the test tiers):
//
//  * The residual of a least-squares problem lies in the left null space of the matrix.  Column by column, a left
//    null vector (phi_i, theta_i per body, sigma_h per hinge) satisfies: the joint point has the same velocity seen
//    from the body and from its parent; theta_j = theta_parent - sigma_j a_j; every foot point is at rest.  It is a
//    virtual motion of the mechanism with all feet pinned and hinge rates qd_h = -sigma_h, i.e. a 6-parameter family
//    indexed by the torso twist (v0, w0): per limb  qd = -J^-1 (v0 + w0 x (fpos - ref)),  J = [a_h x (fpos - jpos_h)].
//  * Residual = Y alpha with  (Y^T Y) alpha = Y^T g.  By virtual work  Y^T g = b - sum_c A_c lambda0_c : the torso
//    wrench left over when every limb carries lambda0_c = J_c^-T (tau_p - tau), the contact force its torque rows
//    alone ask for (b = torso wrench of the particular solution, A_c = [I ; [fpos_c - ref]x]).
//  * With the right-hand side g - Y alpha the system is consistent, so the non-root body rows and the torque rows
//    determine everything limb by limb:  lambda_c = J_c^-T (tau_p' - tau'),  tau_p' = tau_p - tau_p(Y alpha),
//    tau' = tau + qd(alpha).
//
// Phases (same thread layout as hsl_frames_kernel): A | B | F1 limbs (lambda0, Gram share) | F2 trunk (6x6 SPD
// solve) | F3 limbs (lambda).  part[] per limb: [0..18] as in phase B, [19..39] Gram share (lower triangle, row
// major), [40..45] A_c lambda0_c.
#pragma once
#include "hsl_frame.h"

#define HSL_FORCES_PART 46

struct HslForcesLeg {
  double Jinv[9];   // row-major inverse of J (columns w[h])
  double lam0[3];
};

// Inverse of the 3x3 whose COLUMNS are c0, c1, c2 (row-major result).  False when the columns are (numerically)
// coplanar: a stretched limb or a foot on the hip axis -- the reference's QR would drop a column there.
HSL_HD bool cols3_inverse(const double* c0, const double* c1, const double* c2, double* inv) {
  double r0[3], r1[3], r2[3];
  v3_cross(c1, c2, r0);
  v3_cross(c2, c0, r1);
  v3_cross(c0, c1, r2);
  const double det = v3_dot(c0, r0);
  const double scale = hsl_sqrt(v3_dot(c0, c0) * v3_dot(c1, c1) * v3_dot(c2, c2));
  const bool ok = fabs(det) > 1e-12 * scale;
  const double id = ok ? hsl_rcp(det) : 0.0;
#pragma unroll
  for (int k = 0; k < 3; k++) { inv[k] = r0[k] * id; inv[3 + k] = r1[k] * id; inv[6 + k] = r2[k] * id; }
  return ok;
}

// Virtual motion of one limb for the torso twist (v0, w0) with the foot pinned: hinge rates qd and the COM velocity
// ph / angular velocity th of its three bodies.
template <bool DUMP>
HSL_HD void leg_motion(const HslLegState<DUMP>& st, const HslForcesLeg& fs, const double* ref, const double* v0, const double* w0,
                       double* qd, double ph[3][3], double th[3][3]) {
  double rho[3], u[3];
#pragma unroll
  for (int k = 0; k < 3; k++) { rho[k] = st.fpos[k] - ref[k]; u[k] = v0[k]; }
  v3_cross_add(w0, rho, u);
#pragma unroll
  for (int h = 0; h < 3; h++) qd[h] = -(fs.Jinv[3 * h] * u[0] + fs.Jinv[3 * h + 1] * u[1] + fs.Jinv[3 * h + 2] * u[2]);
  double wacc[3] = {w0[0], w0[1], w0[2]};
#pragma unroll
  for (int h = 0; h < 3; h++) {
#pragma unroll
    for (int k = 0; k < 3; k++) { wacc[k] += qd[h] * st.axis[h][k]; th[h][k] = wacc[k]; }
    double d[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { d[k] = st.pos[h][k] - ref[k]; ph[h][k] = v0[k]; }
    v3_cross_add(w0, d, ph[h]);
#pragma unroll
    for (int hp = 0; hp <= h; hp++) {
      double e[3], c[3];
#pragma unroll
      for (int k = 0; k < 3; k++) e[k] = st.pos[h][k] - st.jpos[hp][k];
      v3_cross(st.axis[hp], e, c);
#pragma unroll
      for (int k = 0; k < 3; k++) ph[h][k] += qd[hp] * c[k];
    }
  }
}

template <int NF, int FB, int MODE, bool DUMP, class SM>
HSL_HD void forces_f1_leg(const HslModelPod& M, const HslFrameArgs& A, const SM& sm, const HslSlot& sl, int limb,
                          HslLegState<DUMP>& st, HslForcesLeg& fs) {
  double ref[3];
  root_ref<NF, FB, MODE>(M, A, sm, sl, ref);
  if (!cols3_inverse(st.w[0], st.w[1], st.w[2], fs.Jinv)) st.bad |= HSL_ST_SOLVER;
  // lambda0 = J^-T (tau_p - tau)
  const double* tin = A.tau_in + (int64_t)sl.fo * M.nmj + 3 * limb;
  double dl[3];
#pragma unroll
  for (int h = 0; h < 3; h++) dl[h] = st.taup[h] - tin[h];
#pragma unroll
  for (int k = 0; k < 3; k++) fs.lam0[k] = fs.Jinv[k] * dl[0] + fs.Jinv[3 + k] * dl[1] + fs.Jinv[6 + k] * dl[2];
  double* P = sm.part + (limb * SM::PART) * FB + sl.s;
  // Gram share over the six unit torso twists (21 motion components per twist: 3 x (ph, th) + qd)
  double mot[6][21];
#pragma unroll
  for (int k = 0; k < 6; k++) {
    double v0[3] = {0, 0, 0}, w0[3] = {0, 0, 0}, qd[3], ph[3][3], th[3][3];
    if (k < 3) v0[k] = 1; else w0[k - 3] = 1;
    leg_motion(st, fs, ref, v0, w0, qd, ph, th);
#pragma unroll
    for (int h = 0; h < 3; h++) {
#pragma unroll
      for (int c = 0; c < 3; c++) { mot[k][6 * h + c] = ph[h][c]; mot[k][6 * h + 3 + c] = th[h][c]; }
      mot[k][18 + h] = qd[h];
    }
  }
  int e = 19;
#pragma unroll
  for (int i = 0; i < 6; i++)
#pragma unroll
    for (int j = 0; j <= i; j++) {
      double s = 0;
#pragma unroll
      for (int c = 0; c < 21; c++) s += mot[i][c] * mot[j][c];
      P[(e++) * FB] = s;
    }
  double rho[3], tq[3];
#pragma unroll
  for (int k = 0; k < 3; k++) rho[k] = st.fpos[k] - ref[k];
  v3_cross(rho, fs.lam0, tq);
#pragma unroll
  for (int k = 0; k < 3; k++) { P[(40 + k) * FB] = fs.lam0[k]; P[(43 + k) * FB] = tq[k]; }
}

template <int NF, int FB, int MODE, class SM>
HSL_HD int forces_f2_trunk(const HslModelPod& M, const HslFrameArgs& A, const SM& sm, const HslSlot& sl, HslTrunkState& st) {
  double d[6] = {st.F0[0], st.F0[1], st.F0[2], st.T0[0], st.T0[1], st.T0[2]};
  double G[6][6];
#pragma unroll
  for (int i = 0; i < 6; i++)
#pragma unroll
    for (int j = 0; j < 6; j++) G[i][j] = 0;
  for (int l = 0; l < NF; l++) {
    const double* P = sm.part + (l * SM::PART) * FB + sl.s;
#pragma unroll
    for (int k = 0; k < 6; k++) d[k] += part_get<FB>(sm.part + (l * SM::PART) * FB, sl.s, k) - P[(40 + k) * FB];
    int e = 19;
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int j = 0; j <= i; j++) G[i][j] += P[(e++) * FB];
  }
  // the trunk bodies (torso and the jointless bodies fixed to it) move rigidly with the torso twist
  double ref[3];
  root_ref<NF, FB, MODE>(M, A, sm, sl, ref);
  for (int tb = 0; tb < M.ntrunk; tb++) {
    double r[3];
#pragma unroll
    for (int k = 0; k < 3; k++)
      r[k] = ((MODE == HSL_MODE_FIELDS) ? A.f_pos[((int64_t)sl.i * M.n + M.trunk[tb].body) * 3 + k]
                                        : sm.tpos[(tb * 3 + k) * FB + sl.s]) - ref[k];
    // ph = v0 + w0 x r, th = w0 for the six unit twists
    double mot[6][6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
      double v0[3] = {0, 0, 0}, w0[3] = {0, 0, 0};
      if (k < 3) v0[k] = 1; else w0[k - 3] = 1;
      v3_cross_add(w0, r, v0);
#pragma unroll
      for (int c = 0; c < 3; c++) { mot[k][c] = v0[c]; mot[k][3 + c] = w0[c]; }
    }
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int j = 0; j <= i; j++) {
        double acc = 0;
#pragma unroll
        for (int c = 0; c < 6; c++) acc += mot[i][c] * mot[j][c];
        G[i][j] += acc;
      }
  }
  int bad = 0;
  if (!spd6_solve(G, d)) bad |= HSL_ST_SOLVER;
#pragma unroll
  for (int k = 0; k < 6; k++) sm.mu[k * FB + sl.s] = d[k];
  return bad;
}

template <int NF, int FB, int MODE, bool DUMP, class SM>
HSL_HD void forces_f3_leg(const HslModelPod& M, const HslFrameArgs& A, const SM& sm, const HslSlot& sl, int limb,
                          const HslLegState<DUMP>& st, const HslForcesLeg& fs) {
  double ref[3], al[6];
  root_ref<NF, FB, MODE>(M, A, sm, sl, ref);
#pragma unroll
  for (int k = 0; k < 6; k++) al[k] = sm.mu[k * FB + sl.s];
  double qd[3], ph[3][3], th[3][3];
  leg_motion(st, fs, ref, al, al + 3, qd, ph, th);
  // hinge torques the residual motion field would ask for: a_h . sum_{b >= h} [th_b + (pos_b - jpos_h) x ph_b]
  double dl[3];
#pragma unroll
  for (int h = 0; h < 3; h++) {
    double s = 0;
#pragma unroll
    for (int b = h; b < 3; b++) {
      double e[3], c[3];
#pragma unroll
      for (int k = 0; k < 3; k++) e[k] = st.pos[b][k] - st.jpos[h][k];
      v3_cross(e, ph[b], c);
      s += v3_dot(st.axis[h], th[b]) + v3_dot(st.axis[h], c);
    }
    dl[h] = s + qd[h];
  }
  double lam[3];
#pragma unroll
  for (int k = 0; k < 3; k++) lam[k] = fs.lam0[k] - (fs.Jinv[k] * dl[0] + fs.Jinv[3 + k] * dl[1] + fs.Jinv[6 + k] * dl[2]);
  if (A.z) {
#pragma unroll
    for (int k = 0; k < 3; k++) A.z[(int64_t)(3 * limb + k) * A.n_frames + sl.fo] = lam[k];
  }
}
