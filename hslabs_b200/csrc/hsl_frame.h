// Per-thread frame math of the gait-evaluation kernels (host + device).
//
// One thread owns one (frame, role) pair: role 0..NF-1 is a limb (its three
// hinge bodies), role NF is the trunk (torso + jointless bodies).  The work of
// one frame is split in phases separated by block barriers; each phase is a
// function below so that hsl_kernels.cu (device) and tests/hostcheck (a serial
// CPU emulation used to validate the very same code without a GPU) share it.
//
//   phase A  gait target + closed-form IK + FK of the role's bodies     (a2,a3,a5,a6)
//            -> publish COM positions, u*sin(theta) vectors, cos/sin of the joint angles
//   ---- barrier ----
//   phase B  second differences over the +-2 neighbour frames, joint rates (a7)
//            recursive Newton-Euler pass up the limb                      (a8,a9)
//            per-contact 3x3 level-1 block                                (a10,a11)
//            -> publish limb wrench + contact block
//   ---- barrier ----
//   phase C  (trunk) level-0 6x6 Schur system over the contacts           (a11)
//            -> publish the multiplier
//   ---- barrier ----
//   phase D  contact force, motor torques, positive power of the limb     (a12-a14)
//   ---- barrier ----
//   phase E  (trunk) per-frame power and contact statistics -> global memory
//
// References (file:line in /root/reference): pergen.cpp:62-94,160-198,225-239,
// 386-397; lik.cpp:151-223,316-347; model.cpp:37-62,183-201; dynrec.cpp:134-155,
// 175-224,227-344; ftsolver.cpp:78-146,185-246; periodic.cpp:261-357.
//
// What replaces the reference's linear algebra.  The reference builds the
// 6n x 6n force-torque matrix B and factorises it twice per frame with a sparse
// QR.  Because the bodies form a tree, B^-1 f is the backward Newton-Euler
// recursion, and the null space added by the contact columns is
// {dF_j = -lambda_c, dT_j = -(fpos_c - jpos_j) x lambda_c for j on the chain
// foot -> root}.  With both torso penalties on, level 0 is "torso joint force = 0
// and torso joint torque about the torso COM = 0" (the root's torque row of B has
// no (jpos - pos) x F term, dynrec.cpp:282-287) -- 6 equations in the contact
// forces -- and the level-1 Hessian is block diagonal 3x3 per contact, so the
// lexicographic least-squares solution is one 6x6 symmetric solve per frame plus
// one 3x3 inverse per contact.  tests/test_hostcheck_parity.py checks this code
// against the reference-shaped algorithm of oracle/ to round-off.
#pragma once
#include <math.h>
#include <stdint.h>

#include "hsl_model.h"
#include "hsl_fastmath.h"

#if defined(__CUDACC__)
#define HSL_HD __host__ __device__ __forceinline__
#else
#define HSL_HD inline
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define HSL_MODE_GAIT 0    // candidates -> cost            (a1-a14)
#define HSL_MODE_TRAJ 1    // joint trajectories -> cost    (a5-a14)
#define HSL_MODE_FIELDS 2  // per-frame dynrec fields -> x,z,tau (a8-a13)

#define HSL_FLAG_IGNORE_REACH 1
// Hinge-axis patterns a kernel can be specialised for (all limbs of the model alike, every hinge axis-aligned):
// hip about the parent's y (yxx limbs: myant, hexapod) or z (zxx: spider), knee and ankle about x.
// The patterns also fix which body axis the link offsets lie on (all other components exactly zero):
//   YXX: hip COM, knee / ankle joint offsets and COMs, foot point on y ; ZXX: hip COM and knee joint offset on z,
//   knee COM, ankle joint offset and COM, foot point on y ; the hip joint offset is a general vector in both.
#define HSL_AXP_GENERIC 0
#define HSL_AXP_YXX 1
#define HSL_AXP_ZXX 2
#define HSL_FLAG_REC_TRANSFORM 2  // internal: set by the library when hsl_set_rec_transform is active

struct HslFrameArgs {
  int64_t n_cand;
  int32_t n_t, flags;
  uint32_t div_magic;  // floor(g / (n_t + 4)) = (g * div_magic) >> div_shift for 0 <= g < 2^31 (set by the launchers):
  int32_t div_shift;   // a multiply-shift instead of a division at the head of every block's dependency chain
  int64_t n_frames;         // n_cand * n_t solved frames (FIELDS mode: number of frames)
  const HslCand* cand;      // [C]            (GAIT)
  const double* ttab;       // [C][n_t+4]     (GAIT) accumulated frame times, periodic.cpp:87-91
  const double* traj;       // [C][n_t+5][config_dim] (TRAJ)
  const double* dt_in;      // [C]            (TRAJ)
  // FIELDS mode inputs, reference dynrecord layout [frame][body][3]
  const double *f_pos, *f_jpos, *f_jz, *f_momrate, *f_angrate, *f_fpos;
  const uint8_t* f_contacts;  // [frame][nf]
  const double* tau_in;       // [solved frame][nmj] given motor torques (forces-from-torques kernel, hsl_forces.h)
  // per-frame outputs reduced by the finishing kernel
  double *wframe, *fmin_cfz, *fmax_mu;  // [C][n_t]
  int32_t* status;                      // [C], OR of HSL_ST_*
  // optional dumps, component-major [comp][n_frames] (NULL = not written)
  double *x, *z, *tau;
  double* q_out;       // [config_dim][C*(n_t+4)] generated joint values (GAIT)
  uint8_t* contacts;   // [nf][n_frames]
  long long* phase_clk;  // [blocks][warps][8] cycle stamps, only written by -DHSL_PHASE_CLOCKS builds (profiling aid)
  // pergensetup::rec_transform (pergen.cpp:309-335): rigid map applied to every generated frame record
  double rec_R[9], rec_t[3];  // column-major rotation, translation; used when flags & HSL_FLAG_REC_TRANSFORM
};

// ------------------------------------------------------------------ small vector helpers
HSL_HD void v3_cross(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
HSL_HD void v3_cross_add(const double* a, const double* b, double* c) {
  c[0] += a[1] * b[2] - a[2] * b[1];
  c[1] += a[2] * b[0] - a[0] * b[2];
  c[2] += a[0] * b[1] - a[1] * b[0];
}
HSL_HD double v3_dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
// C = A*B, column-major 3x3
HSL_HD void m3_mul(const double* A, const double* B, double* C) {
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int i = 0; i < 3; i++) C[3 * j + i] = A[i] * B[3 * j] + A[3 + i] * B[3 * j + 1] + A[6 + i] * B[3 * j + 2];
}
// y = A*x + t
HSL_HD void m3_affine(const double* A, const double* x, const double* t, double* y) {
#pragma unroll
  for (int i = 0; i < 3; i++) y[i] = A[i] * x[0] + A[3 + i] * x[1] + A[6 + i] * x[2] + t[i];
}
// Rotation the reference gets from ODE's dRFromEulerAngles copied raw into its column-major
// affine (model.cpp:45-47, visualization.cpp:62-69): Rz(psi) Ry(theta) Rx(phi).
HSL_HD void euler_to_R(double phi, double theta, double psi, double* R) {
  double sphi, cphi, sth, cth, spsi, cpsi;
  sincos(phi, &sphi, &cphi);
  sincos(theta, &sth, &cth);
  sincos(psi, &spsi, &cpsi);
  R[0] = cpsi * cth;                      R[1] = spsi * cth;                      R[2] = -sth;
  R[3] = cpsi * sth * sphi - spsi * cphi; R[4] = spsi * sth * sphi + cpsi * cphi; R[5] = cth * sphi;
  R[6] = cpsi * sth * cphi + spsi * sphi; R[7] = spsi * sth * cphi - cpsi * sphi; R[8] = cth * cphi;
}
HSL_HD void wrap_pm_pi(double& a) {  // visualization.cpp:73-79
  if (a < -M_PI) { while (a < -M_PI) a += 2 * M_PI; }
  else if (a > M_PI) { while (a > M_PI) a -= 2 * M_PI; }
}

// Pattern shared by every limb of the model, or HSL_AXP_GENERIC.  Used by the launchers (and the host emulation) to
// pick a kernel specialised for it.
HSL_HD bool hsl_on_axis(const double* v, int ax) {
  return v[(ax + 1) % 3] == 0.0 && v[(ax + 2) % 3] == 0.0;
}
HSL_HD int hsl_axis_pattern(const HslModelPod& M) {
  int pat = -1;
  for (int l = 0; l < M.nf; l++) {
    const HslLimb& L = M.limb[l];
    const int a0 = L.h[0].aligned < 0 ? -L.h[0].aligned : L.h[0].aligned;
    const int a1 = L.h[1].aligned < 0 ? -L.h[1].aligned : L.h[1].aligned;
    const int a2 = L.h[2].aligned < 0 ? -L.h[2].aligned : L.h[2].aligned;
    int p = HSL_AXP_GENERIC;
    const bool tail = a1 == 1 && a2 == 1 && hsl_on_axis(L.h[1].com, 1) && hsl_on_axis(L.h[2].tjp, 1) && hsl_on_axis(L.h[2].com, 1) &&
                      hsl_on_axis(L.foot, 1);
    if (tail && a0 == 2 && L.kind == HSL_IK_YXX && hsl_on_axis(L.h[0].com, 1) && hsl_on_axis(L.h[1].tjp, 1)) p = HSL_AXP_YXX;
    if (tail && a0 == 3 && L.kind == HSL_IK_ZXX && hsl_on_axis(L.h[0].com, 2) && hsl_on_axis(L.h[1].tjp, 2)) p = HSL_AXP_ZXX;
    if (pat < 0) pat = p;
    if (p != pat) return HSL_AXP_GENERIC;
  }
  return pat < 0 ? HSL_AXP_GENERIC : pat;
}

// ------------------------------------------------------------------ gait generator (a2)
// Register copy of the candidate constants one thread needs.  Loading them all up front (before any branch that
// depends on them) puts every global load of the prologue in flight at once: one L2 round trip instead of one per
// dependent use.
struct HslCandView {
  double R0[9], tp0[3], eul[3];
  double period, step_length, step_height, v, t_step, curvature, max_radius;
  double pos0[3], ts, xs, turn_r, turn_sa, turn_ca;  // of the thread's limb (unused by the trunk)
};
HSL_HD void load_cand(const HslCand& c, int limb, HslCandView& o) {
#pragma unroll
  for (int k = 0; k < 9; k++) o.R0[k] = c.R0[k];
#pragma unroll
  for (int k = 0; k < 3; k++) { o.tp0[k] = c.tp0[k]; o.eul[k] = c.eul[k]; }
  o.period = c.period; o.step_length = c.step_length; o.step_height = c.step_height; o.v = c.v;
  o.t_step = c.t_step; o.curvature = c.curvature; o.max_radius = c.max_radius;
  if (limb >= 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) o.pos0[k] = c.pos0[limb][k];
    o.ts = c.ts[limb]; o.xs = c.xs[limb];
    o.turn_r = c.turn_r[limb]; o.turn_sa = c.turn_sa[limb]; o.turn_ca = c.turn_ca[limb];
  }
}

// Torso pose at time t: pergensetup::turn_torso, pergen.cpp:386-397.  qt = torso joint translation, R0 = torso rotation.
// Straight walking (no turn): the candidate's rotation, x advanced by t*v.  Curved walking: the reference composes
// the turn [Rz(psi) | arc] with the candidate orientation (pergen.cpp:187-198, 377-383), extracts Euler angles from
// the product (visualization.cpp:81-101) and the FK rebuilds the rotation from them; that round trip is the identity
// away from gimbal lock, so the rotation is taken directly as Rz(psi) * R0 and the Euler angles (asin / atan2) are
// only evaluated when a trajectory dump asks for them (eul != nullptr).
// euler_angles_from_affine, visualization.cpp:81-101 (column-major R)
HSL_HD void euler_from_R(const double* R, double* eul) {
  const double th = -asin(R[2]), ct = cos(th);
  eul[0] = atan2(R[5] / ct, R[8] / ct);
  eul[1] = th;
  eul[2] = atan2(R[1] / ct, R[0] / ct);
}
HSL_HD void torso_pose(const HslCandView& cd, double t, double* qt, double* R0, double* eul) {
  const double tv = t * cd.v;
  double psi = 0;
  if (cd.curvature != 0) {
    const int s = (cd.curvature > 0) ? 1 : -1;
    psi = s * tv / cd.max_radius;
  }
  if (psi == 0) {
    qt[0] = cd.tp0[0] + tv; qt[1] = cd.tp0[1]; qt[2] = cd.tp0[2];
#pragma unroll
    for (int k = 0; k < 9; k++) R0[k] = cd.R0[k];
    if (eul) { eul[0] = cd.eul[0]; eul[1] = cd.eul[1]; eul[2] = cd.eul[2]; }
    return;
  }
  const double rc = hsl_rcp(cd.curvature);
  double sp, cp;
  hsl_sincos_pm_pi(psi, &sp, &cp);
  // [Rz(psi) | (rc sin psi, rc (1 - cos psi), 0)] applied to the candidate pose
#pragma unroll
  for (int j = 0; j < 3; j++) {
    R0[3 * j] = cp * cd.R0[3 * j] - sp * cd.R0[3 * j + 1];
    R0[3 * j + 1] = sp * cd.R0[3 * j] + cp * cd.R0[3 * j + 1];
    R0[3 * j + 2] = cd.R0[3 * j + 2];
  }
  qt[0] = cp * cd.tp0[0] - sp * cd.tp0[1] + rc * sp;
  qt[1] = sp * cd.tp0[0] + cp * cd.tp0[1] + rc * (1 - cp);
  qt[2] = cd.tp0[2];
  if (eul) euler_from_R(R0, eul);
}
// pergensetup::transform_rec, pergen.cpp:325-335 with transform_orientation :377-383: the torso pose becomes
// A * pose (Euler angles re-extracted only for a trajectory dump, as in torso_pose) ...
HSL_HD void rec_transform_pose(const HslFrameArgs& A, double* qt, double* R0, double* eul) {
  double Rn[9], tn[3];
  m3_mul(A.rec_R, R0, Rn);
  m3_affine(A.rec_R, qt, A.rec_t, tn);
#pragma unroll
  for (int k = 0; k < 9; k++) R0[k] = Rn[k];
#pragma unroll
  for (int k = 0; k < 3; k++) qt[k] = tn[k];
  if (eul) euler_from_R(R0, eul);
}
// ... and every foot target becomes A * p.
HSL_HD void rec_transform_point(const HslFrameArgs& A, double* p) {
  double pn[3];
  m3_affine(A.rec_R, p, A.rec_t, pn);
#pragma unroll
  for (int k = 0; k < 3; k++) p[k] = pn[k];
}
// Torso body frame from its joint values: model.cpp:183-195 with the free joint (model.cpp:40-48).
HSL_HD void torso_frame(const HslModelPod& M, const double* qt, const double* R0, double* t0) {
#pragma unroll
  for (int i = 0; i < 3; i++) t0[i] = R0[i] * M.Qt[0] + R0[3 + i] * M.Qt[1] + R0[6 + i] * M.Qt[2] + (qt[i] + M.Pt[i]);
}
// Foot target of one limb at time t: periodicgenerator::limb_positions / turn_position, pergen.cpp:62-94,160-183.
HSL_HD void foot_target(const HslCandView& cd, double t, double* p) {
  const double tr = hsl_div(t, cd.period);
  const int t_int = (int)tr;
  const double tf = tr - t_int;
  const double tl = cd.ts;
  double sf;
  if (tf < tl) sf = 0;
  else if (tf < tl + cd.t_step) sf = hsl_div(tf - tl, cd.t_step);
  else sf = 1;
  double sn, cs;
  hsl_sincos_0_pi(M_PI * sf, &sn, &cs);
  double dx = (t_int + cd.xs + (1 - cs) / 2) * cd.step_length;
  double dy = 0;
  const double dz = sn * sn * cd.step_height;
  if (cd.curvature != 0) {
    // arc about the turning centre (pergen.cpp:160-183); r and the polar angle alpha of pos0 are per-candidate
    // constants (setup_candidate), so sin/cos(alpha - beta/2) come from one sincos of the small angle beta/2
    const int s = (cd.curvature > 0) ? 1 : -1;
    const double hb = -s * dx / cd.max_radius / 2;
    double shb, chb;
    hsl_sincos_pm_pi(hb, &shb, &chb);
    const double sg = cd.turn_sa * chb - cd.turn_ca * shb, cg = cd.turn_ca * chb + cd.turn_sa * shb;
    const double sb = 2 * shb;
    dx = cd.turn_r * sg * sb;
    dy += -cd.turn_r * cg * sb;
  }
  p[0] = dx + cd.pos0[0];
  p[1] = dy + cd.pos0[1];
  p[2] = dz + cd.pos0[2];
}

// ------------------------------------------------------------------ closed-form limb IK (a3)
// lik.cpp:151-223.  pl = foot target in the hip joint frame.  Returns false when out of reach.
// The reference computes the three joint angles with atan2 / acos and the FK that follows takes their
// sines and cosines again.  Here the cosines and sines of the joint angles are formed algebraically from
// the same intermediate quantities (cos(acos(c)) = c, sin(acos(c)) = sqrt((1-c)(1+c)), angle-sum formulas),
// so the hot path needs no inverse trigonometry at all; the angles themselves (ang != nullptr) are only
// evaluated when a trajectory dump is requested.
//   q0 = -phi,  q1 = -theta + beta,  q2 = -(beta + gamma)                       (lik.cpp:181,220)
// AXP: hinge-axis pattern of the model known at compile time (see HSL_AXP_* below), 0 = read L.kind at run time.
template <int AXP = 0>
HSL_HD bool limb_ik(const HslLimb& L, const double* pl, bool ignore_reach, double* cq, double* sq, double* ang) {
  const double l0 = L.ls[0], l1 = L.ls[1], l2 = L.ls[2];
  const int s0 = L.ysign, s1 = 2 * (L.bend != 0) - 1;
  const bool yxx = (AXP == 0) ? (L.kind == HSL_IK_YXX) : (AXP == HSL_AXP_YXX);
  const double zoff = yxx ? pl[2] - s0 * l0 : pl[2] + l0;
  const double rho2 = pl[0] * pl[0] + pl[1] * pl[1];
  const double l2sq = rho2 + zoff * zoff;
  // Quotients by l are formed with one reciprocal square root (products, ~1.5 ulp).  When the angles are requested
  // (trajectory dump) the reference's own operation order -- l = sqrt(.), then IEEE divisions -- is used instead, so
  // that the dumped angles match it bit for bit wherever acos / atan2 are well conditioned.
  double l, c, cb, cg;
  bool ok = true;
  const double del = l2 * l2 - l1 * l1;
  const int sg = yxx ? s1 * s0 : s1;
  if (ang) {
    l = hsl_sqrt(l2sq);
    if (l1 + l2 - l < 0) { if (ignore_reach) l = l1 + l2; else ok = false; }
    const double ll = l * l;
    c = hsl_div(zoff, l);
    cb = hsl_div(ll - del, 2 * l1 * l);
    cg = hsl_div(ll + del, 2 * l2 * l);
  } else {
    double rl = hsl_rsqrt(l2sq);
    l = l2sq * rl;
    if (l1 + l2 - l < 0) {
      if (ignore_reach) { l = l1 + l2; rl = hsl_rcp(l); } else ok = false;
    }
    const double ll = l * l;
    c = zoff * rl;
    cb = (ll - del) * rl * L.inv2l1;
    cg = (ll + del) * rl * L.inv2l2;
    // products can overshoot +-1 by an ulp exactly at the boundary (leg fully stretched, l == l1 + l2), where the
    // IEEE quotient is exactly 1; pull such values back so that the square roots below stay real
    c = (fabs(c) > 1.0 && fabs(c) < 1.0 + 1e-15) ? copysign(1.0, c) : c;
    cb = (cb > 1.0 && cb < 1.0 + 1e-15) ? 1.0 : cb;
    cg = (cg > 1.0 && cg < 1.0 + 1e-15) ? 1.0 : cg;
  }
  // phi = atan2(x, y): cos(phi) = y / rho, sin(phi) = x / rho  (atan2(0,0) = 0)
  const double rrho = hsl_rsqrt(rho2);
  const double cphi = (rho2 > 0) ? pl[1] * rrho : 1.0, sphi = (rho2 > 0) ? pl[0] * rrho : 0.0;
  const double st0 = hsl_sqrt((1 - c) * (1 + c));  // sin(acos(c)) >= 0 ; NaN when |c| > 1, as acos would be
  double cth, sth;
  if (yxx) {  // theta = acos(c) + (1 - s0) pi/2 : unchanged for s0 = 1, shifted by pi for s0 = -1
    cth = (s0 > 0) ? c : -c;
    sth = (s0 > 0) ? st0 : -st0;
  } else {    // theta = acos(c) - s0 pi/2
    cth = s0 * st0;
    sth = -s0 * c;
  }
  const double sb = sg * hsl_sqrt((1 - cb) * (1 + cb)), sgm = sg * hsl_sqrt((1 - cg) * (1 + cg));
  cq[0] = cphi;                 sq[0] = -sphi;
  cq[1] = cth * cb + sth * sb;  sq[1] = cth * sb - sth * cb;
  cq[2] = cb * cg - sb * sgm;   sq[2] = -(sb * cg + cb * sgm);
  if (ang) {
    double phi = atan2(pl[0], pl[1]);
    double theta = yxx ? acos(c) + (1 - s0) * M_PI / 2 : acos(c) - s0 * M_PI / 2;
    wrap_pm_pi(phi);
    wrap_pm_pi(theta);
    const double beta = sg * acos(cb), gamma = sg * acos(cg);
    ang[0] = -phi;
    ang[1] = -theta + beta;
    ang[2] = -(beta + gamma);
  }
  return ok;
}

// ------------------------------------------------------------------ shared-memory view
// All exchange arrays are [field][slot] so that consecutive lanes (= consecutive frames) hit
// consecutive 8-byte words: conflict-free LDS.64 / STS.64.
struct alignas(16) HslD2 { double x, y; };
template <int NF, int FB, int PARTN = 19>
struct HslSmem {
  static constexpr int PART = PARTN;  // doubles per limb in `part`
  HslD2* pu;     // [9*NF][FB]      limb bodies: (COM position component, u*sin(theta) component of the body rotation)
                 //                 as one 16-byte entry -- both get the same +-2 frame stencil, so phase B reads them
                 //                 with one LDS.128 per stencil point instead of two LDS.64 (pairing the cos/sin and
                 //                 the parked phase-D entries the same way measured no gain)
  double* tpos;  // [ntrunk*3][FB]  trunk bodies' COM positions
  double* tust;  // [3][FB]         u*sin(theta) of the trunk rotation (one entry for all trunk bodies)
  double* cs;    // [6*NF][FB]                cos, sin of the hinge angles
  double* part;  // [NF*PART][FB]             limb -> trunk partials ; reused for limb -> trunk results after phase D
  double* mu;    // [7][FB]                   trunk -> limb multiplier (+ validity)
  HSL_HD static int doubles_per_slot(int ntrunk) { return (3 * NF + ntrunk) * 3 + (3 * NF + 1) * 3 + 6 * NF + NF * PARTN + 7; }
  HSL_HD void carve(double* base, int ntrunk) {  // base must be 16-byte aligned
    pu = reinterpret_cast<HslD2*>(base);
    tpos = base + 18 * NF * FB;
    tust = tpos + ntrunk * 3 * FB;
    cs = tust + 3 * FB;
    part = cs + 6 * NF * FB;
    mu = part + NF * PARTN * FB;
  }
};
// part fields per limb: [0..2] Fl  [3..5] Tl  [6..11] W  [12..14] Wg  [15..17] r  [18] contact flag.  Fields 0..17 are
// stored as nine 16-byte pairs (field 2j, 2j+1 of a slot together: [j][slot][2]) so that the solver reads a limb with
// 9 LDS.128 + 1 LDS.64 instead of 19 LDS.64; fields >= 18 are single doubles at [field][slot] as before.
template <int FB>
HSL_HD HslD2 part_pair_load(const double* limb_base, int s, int j) {
  return reinterpret_cast<const HslD2*>(limb_base)[j * FB + s];
}
template <int FB>
HSL_HD void part_pair_store(double* limb_base, int s, int j, double a, double b) {
  HslD2 e;
  e.x = a; e.y = b;
  reinterpret_cast<HslD2*>(limb_base)[j * FB + s] = e;
}
template <int FB>
HSL_HD double part_get(const double* limb_base, int s, int k) {  // one field 0..17 (dump paths)
  return limb_base[2 * ((k >> 1) * FB + s) + (k & 1)];
}
// (pipelined kernel, PART = 34: + [19..21] tau_p  [22..30] w  [31..33] qdot)

// Where a thread sits.
struct HslSlot {
  int64_t c;     // candidate (FIELDS: 0)
  int32_t i;     // generated-frame index 0..n_t+3 (FIELDS: frame index)
  int32_t s;     // slot inside the block
  int64_t fo;    // index of the solved frame in the per-frame outputs (c*n_t + i-2), valid when interior
  bool valid;    // slot maps to an existing (candidate, frame)
  bool interior; // frame is solved by this block (has its +-2 neighbours in the block)
};

// Limb thread state carried across phases (lives in registers on the device).
template <bool DUMP>
struct HslLegState {
  double jpos[3][3], axis[3][3], pos[3][3], fpos[3];  // A -> B
  double taup[3], w[3][3], W[6], Wg[3], r[3], qd[3];  // B -> D
  double T[3][3], F[3][3];                            // B -> D, only read when DUMP
  int contact, bad;
};
struct HslTrunkState {
  double R0[9], t0[3];  // A -> B
  double F0[3], T0[3];  // B -> C : trunk bodies' own share of the root wrench
  double lam[HSL_MAX_LIMBS][3];  // C (DUMP only)
};

// ------------------------------------------------------------------ phase A
// FK of one hinge body given its parent's frame (model.cpp:183-195): joint frame, body frame, COM, ust.
// cs, sn = cosine and sine of the joint value.
template <int AX>  // rotation about the parent-frame coordinate axis AX (cyclic AX -> B -> C)
HSL_HD void rot_about_axis(const double* Rp, double cs, double sn, double* Rb) {
  constexpr int B = (AX + 1) % 3, C = (AX + 2) % 3;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    Rb[3 * AX + i] = Rp[3 * AX + i];
    Rb[3 * B + i] = Rp[3 * B + i] * cs + Rp[3 * C + i] * sn;
    Rb[3 * C + i] = Rp[3 * C + i] * cs - Rp[3 * B + i] * sn;
  }
}
// y = t + R[:, AXV] * v[AXV] when the vector v is known to lie on body axis AXV (AXV < 0: general vector)
template <int AXV>
HSL_HD void m3_affine_ax(const double* R, const double* v, const double* t, double* y) {
  if (AXV < 0) { m3_affine(R, v, t, y); return; }
#pragma unroll
  for (int i = 0; i < 3; i++) y[i] = R[3 * (AXV < 0 ? 0 : AXV) + i] * v[AXV < 0 ? 0 : AXV] + t[i];
}
// Hinge known at compile time to turn about +-(coordinate axis AX of the parent frame) with the joint at the body
// origin, joint offset along parent axis TAX and COM offset along body axis CAX (negative = general vector): same
// arithmetic as the aligned branch of hinge_fk below without its run-time dispatch (the three-way branch costs merge
// moves for the nine entries of Rb and keeps the scheduler from overlapping consecutive hinges) and without the
// products with the structural zeros of the offsets.
template <int AX, int TAX, int CAX>
HSL_HD void hinge_fk_aligned(const HslHinge& H, double cs, double sn, const double* Rp, const double* tp, double* Rb, double* tb,
                             double* jpos, double* axis, double* com, double* ust) {
  m3_affine_ax<TAX>(Rp, H.tjp, tp, jpos);
  const double sg = (H.aligned > 0) ? 1.0 : -1.0, s2 = sg * sn;
  rot_about_axis<AX>(Rp, cs, s2, Rb);
  axis[0] = sg * Rp[3 * AX]; axis[1] = sg * Rp[3 * AX + 1]; axis[2] = sg * Rp[3 * AX + 2];
  tb[0] = jpos[0]; tb[1] = jpos[1]; tb[2] = jpos[2];
  m3_affine_ax<CAX>(Rb, H.com, tb, com);
  ust[0] = (Rb[5] - Rb[7]) / 2;
  ust[1] = (Rb[6] - Rb[2]) / 2;
  ust[2] = (Rb[1] - Rb[3]) / 2;
}
HSL_HD void hinge_fk(const HslHinge& H, double cs, double sn, const double* Rp, const double* tp, double* Rb, double* tb,
                     double* jpos, double* axis, double* com, double* ust) {
  m3_affine(Rp, H.tjp, tp, jpos);
  if (H.aligned != 0) {
    // hinge axis is +-(a coordinate axis of the parent frame) and the joint sits at the body origin (all three
    // reference models): A_parent * Rz(q) * A_pj_body is the rotation by q about that axis -- two column mixes.
    const int a = (H.aligned > 0 ? H.aligned : -H.aligned) - 1;
    const double sg = (H.aligned > 0) ? 1.0 : -1.0, s2 = sg * sn;
    if (a == 0) rot_about_axis<0>(Rp, cs, s2, Rb);
    else if (a == 1) rot_about_axis<1>(Rp, cs, s2, Rb);
    else rot_about_axis<2>(Rp, cs, s2, Rb);
    const double* ca = (a == 0) ? Rp : (a == 1) ? Rp + 3 : Rp + 6;
    axis[0] = sg * ca[0]; axis[1] = sg * ca[1]; axis[2] = sg * ca[2];
    tb[0] = jpos[0]; tb[1] = jpos[1]; tb[2] = jpos[2];
  } else {
    double Rj[9];
    m3_mul(Rp, H.Rjp, Rj);
    axis[0] = Rj[6]; axis[1] = Rj[7]; axis[2] = Rj[8];  // dynpart::get_joint_zaxis, dynrec.cpp:84-93
    double Rz[9];
#pragma unroll
    for (int i = 0; i < 3; i++) {  // Rj * Rz(q), model.cpp:49-57
      Rz[i] = Rj[i] * cs + Rj[3 + i] * sn;
      Rz[3 + i] = Rj[3 + i] * cs - Rj[i] * sn;
      Rz[6 + i] = Rj[6 + i];
    }
    m3_mul(Rz, H.Rpb, Rb);
    m3_affine(Rz, H.tpb, jpos, tb);
  }
  m3_affine(Rb, H.com, tb, com);
  ust[0] = (Rb[5] - Rb[7]) / 2;  // dynrec.cpp:142-145: (A(2,1)-A(1,2))/2 ...
  ust[1] = (Rb[6] - Rb[2]) / 2;
  ust[2] = (Rb[1] - Rb[3]) / 2;
}

// Foot target p (world) into the hip joint frame of limb L (lik.cpp:341-347): R0, t0 = torso rotation / body origin.
HSL_HD void foot_into_hip_frame(const HslLimb& L, const double* R0, const double* t0, const double* p, double* pl) {
  double ta[3], th[3], d[3], e[3];
  m3_affine(R0, L.oatt, t0, ta);       // frame of the trunk body the limb hangs from
  m3_affine(R0, L.h[0].tjp, ta, th);   // hip joint origin
#pragma unroll
  for (int k = 0; k < 3; k++) d[k] = p[k] - th[k];
  // (R0 Rjp)^T d = Rjp^T (R0^T d): two matrix-vector products instead of a matrix-matrix product
#pragma unroll
  for (int k = 0; k < 3; k++) e[k] = R0[3 * k] * d[0] + R0[3 * k + 1] * d[1] + R0[3 * k + 2] * d[2];
  const double* Rj0 = L.h[0].Rjp;
#pragma unroll
  for (int k = 0; k < 3; k++) pl[k] = Rj0[3 * k] * e[0] + Rj0[3 * k + 1] * e[1] + Rj0[3 * k + 2] * e[2];
}

template <int NF, int FB, int MODE, bool DUMP, int AXP = 0, class SM>
HSL_HD void phase_a_leg(const HslModelPod& M, const HslFrameArgs& A, const SM& sm, const HslSlot& sl, int limb,
                        HslLegState<DUMP>& st) {
  const HslLimb& L = M.limb[limb];
  st.bad = 0;
  st.contact = 0;
  if (MODE == HSL_MODE_FIELDS) {
    const int64_t fb = sl.i;
#pragma unroll
    for (int h = 0; h < 3; h++) {
      const int b = L.h[h].body;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        st.pos[h][k] = A.f_pos[(fb * M.n + b) * 3 + k];
        st.jpos[h][k] = A.f_jpos[(fb * M.n + b) * 3 + k];
        st.axis[h][k] = A.f_jz[(fb * M.n + b) * 3 + k];
      }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) st.fpos[k] = A.f_fpos[(fb * M.nf + limb) * 3 + k];
    st.contact = A.f_contacts[fb * M.nf + limb] != 0;
    return;
  }
  double R0[9], t0[3], qt[3], eul[3], qa[3], cq[3], sq[3];
  if (MODE == HSL_MODE_GAIT) {
    HslCandView cd;
    load_cand(A.cand[sl.c], limb, cd);
    const double t = A.ttab[sl.c * (A.n_t + 4) + sl.i];
    torso_pose(cd, t, qt, R0, (DUMP && A.q_out != nullptr) ? eul : nullptr);
    if (A.flags & HSL_FLAG_REC_TRANSFORM) rec_transform_pose(A, qt, R0, (DUMP && A.q_out != nullptr) ? eul : nullptr);
    torso_frame(M, qt, R0, t0);
    // foot target into the hip joint frame (lik.cpp:341-347), then the closed-form solver
    double p[3], Rh[9], th[3], oj[3], d[3], pl[3];
    foot_target(cd, t, p);
    if (A.flags & HSL_FLAG_REC_TRANSFORM) rec_transform_point(A, p);
#pragma unroll
    for (int k = 0; k < 3; k++) oj[k] = L.oatt[k];
    foot_into_hip_frame(L, R0, t0, p, pl);
    (void)Rh; (void)th; (void)oj; (void)d;
    const bool want_angles = DUMP && A.q_out != nullptr;
    if (!limb_ik<AXP>(L, pl, (A.flags & HSL_FLAG_IGNORE_REACH) != 0, cq, sq, want_angles ? qa : nullptr)) st.bad |= HSL_ST_UNREACHABLE;
  } else {  // HSL_MODE_TRAJ
    const double* qrow = A.traj + (sl.c * (A.n_t + 5) + sl.i) * M.config_dim;
#pragma unroll
    for (int k = 0; k < 3; k++) { qt[k] = qrow[k]; eul[k] = qrow[3 + k]; qa[k] = qrow[6 + 3 * limb + k]; }
#pragma unroll
    for (int k = 0; k < 3; k++) sincos(qa[k], &sq[k], &cq[k]);
    euler_to_R(eul[0], eul[1], eul[2], R0);
    torso_frame(M, qt, R0, t0);
  }
  // FK down the limb
  double Rp[9], tp[3];
#pragma unroll
  for (int k = 0; k < 9; k++) Rp[k] = R0[k];
  m3_affine(R0, L.oatt, t0, tp);
  double Rb[9], tb[3];
#pragma unroll
  for (int h = 0; h < 3; h++) {
    double ust[3];
    // per-pattern (hinge axis, joint-offset axis, COM-offset axis) of the three hinges: see hsl_axis_pattern()
    if (AXP == 0) hinge_fk(L.h[h], cq[h], sq[h], Rp, tp, Rb, tb, st.jpos[h], st.axis[h], st.pos[h], ust);
    else if (h == 2) hinge_fk_aligned<0, 1, 1>(L.h[h], cq[h], sq[h], Rp, tp, Rb, tb, st.jpos[h], st.axis[h], st.pos[h], ust);
    else if (AXP == HSL_AXP_YXX && h == 0) hinge_fk_aligned<1, -1, 1>(L.h[h], cq[h], sq[h], Rp, tp, Rb, tb, st.jpos[h], st.axis[h], st.pos[h], ust);
    else if (AXP == HSL_AXP_YXX) hinge_fk_aligned<0, 1, 1>(L.h[h], cq[h], sq[h], Rp, tp, Rb, tb, st.jpos[h], st.axis[h], st.pos[h], ust);
    else if (h == 0) hinge_fk_aligned<2, -1, 2>(L.h[h], cq[h], sq[h], Rp, tp, Rb, tb, st.jpos[h], st.axis[h], st.pos[h], ust);
    else hinge_fk_aligned<0, 2, 1>(L.h[h], cq[h], sq[h], Rp, tp, Rb, tb, st.jpos[h], st.axis[h], st.pos[h], ust);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      HslD2 e;
      e.x = st.pos[h][k]; e.y = ust[k];
      sm.pu[((3 * limb + h) * 3 + k) * FB + sl.s] = e;
    }
    sm.cs[(6 * limb + 2 * h) * FB + sl.s] = cq[h];
    sm.cs[(6 * limb + 2 * h + 1) * FB + sl.s] = sq[h];
#pragma unroll
    for (int k = 0; k < 9; k++) Rp[k] = Rb[k];
#pragma unroll
    for (int k = 0; k < 3; k++) tp[k] = tb[k];
  }
  if (AXP == 0) m3_affine(Rb, L.foot, tb, st.fpos);   // odepart::get_foot_pos, visualization.cpp:565
  else m3_affine_ax<1>(Rb, L.foot, tb, st.fpos);
  st.contact = (st.fpos[2] < M.rcap + 1e-4);          // dynrec.cpp:149
  if (DUMP && MODE == HSL_MODE_GAIT && A.q_out && sl.valid) {
    const int64_t tot = A.n_cand * (A.n_t + 4), g = sl.c * (A.n_t + 4) + sl.i;
#pragma unroll
    for (int h = 0; h < 3; h++) A.q_out[(6 + 3 * limb + h) * tot + g] = qa[h];
    if (limb == 0) {
#pragma unroll
      for (int k = 0; k < 3; k++) { A.q_out[k * tot + g] = qt[k]; A.q_out[(3 + k) * tot + g] = eul[k]; }
    }
  }
}

template <int NF, int FB, int MODE, class SM>
HSL_HD void phase_a_trunk(const HslModelPod& M, const HslFrameArgs& A, const SM& sm, const HslSlot& sl,
                          HslTrunkState& st) {
  if (MODE == HSL_MODE_FIELDS) return;
  double qt[3], eul[3];
  if (MODE == HSL_MODE_GAIT) {
    HslCandView cd;
    load_cand(A.cand[sl.c], -1, cd);
    const double t = A.ttab[sl.c * (A.n_t + 4) + sl.i];
    torso_pose(cd, t, qt, st.R0, nullptr);
    if (A.flags & HSL_FLAG_REC_TRANSFORM) rec_transform_pose(A, qt, st.R0, nullptr);
  } else {
    const double* qrow = A.traj + (sl.c * (A.n_t + 5) + sl.i) * M.config_dim;
#pragma unroll
    for (int k = 0; k < 3; k++) { qt[k] = qrow[k]; eul[k] = qrow[3 + k]; }
    euler_to_R(eul[0], eul[1], eul[2], st.R0);
  }
  torso_frame(M, qt, st.R0, st.t0);
  for (int tb = 0; tb < M.ntrunk; tb++) {
    double ob[3], pb[3];
    m3_affine(st.R0, M.trunk[tb].off, st.t0, ob);
    m3_affine(st.R0, M.trunk[tb].com, ob, pb);
#pragma unroll
    for (int k = 0; k < 3; k++) sm.tpos[(tb * 3 + k) * FB + sl.s] = pb[k];
  }
  sm.tust[0 * FB + sl.s] = (st.R0[5] - st.R0[7]) / 2;
  sm.tust[1 * FB + sl.s] = (st.R0[6] - st.R0[2]) / 2;
  sm.tust[2 * FB + sl.s] = (st.R0[1] - st.R0[3]) / 2;
}

// ------------------------------------------------------------------ phase B
// Reference point of the torso (root) wrench.  The root's torque row of B has no (jpos - pos) x F term
// (the cross elements are only inserted when the body has a parent, dynrec.cpp:282-287), so its joint
// torque is taken about the torso COM, which moves with the frame.
template <int NF, int FB, int MODE, class SM>
HSL_HD void root_ref(const HslModelPod& M, const HslFrameArgs& A, const SM& sm, const HslSlot& sl, double* ref) {
#pragma unroll
  for (int k = 0; k < 3; k++)
    ref[k] = (MODE == HSL_MODE_FIELDS) ? A.f_pos[((int64_t)sl.i * M.n + M.trunk[0].body) * 3 + k]
                                       : sm.tpos[k * FB + sl.s];
}
// Second central difference over +-2 frames.  The reference stages it (dynrec.cpp:175-224): first differences
// (f(s+1)-f(s-1))*hh at frames s+-1, scaled by the mass / inertia, then their difference times hh again
// (hh = 1/(2 dt)).  The two first differences f(s+2)-f(s) and f(s)-f(s-2) are exact in floating point (neighbouring
// frames differ by far less than a factor 2), so the staged result equals ((p2-c0) + (m2-c0)) * hh^2 * scale up to
// the rounding of the staged products; this form needs 4 FP64 instructions instead of 7.
HSL_HD double fd2(const double* a, int s, int FB_, double hh, double scale) {
  const double m2 = a[s - 2], c0 = a[s], p2 = a[s + 2];
  (void)FB_;
  return ((p2 - c0) + (m2 - c0)) * (hh * hh * scale);
}
// The same for a (position, u sin theta) pair stored as one 16-byte entry: three 128-bit loads.
HSL_HD void fd2_pair(const HslD2* a, int s, double hh, double scale_x, double scale_y, double* dx, double* dy) {
  const HslD2 m2 = a[s - 2], c0 = a[s], p2 = a[s + 2];
  *dx = ((p2.x - c0.x) + (m2.x - c0.x)) * (hh * hh * scale_x);
  *dy = ((p2.y - c0.y) + (m2.y - c0.y)) * (hh * hh * scale_y);
}
// Inverse of a symmetric positive definite 3x3 stored as (00,01,02,11,12,22): adjugate over determinant (one
// reciprocal, all cofactors independent of each other -- a short dependency chain).  Returns false unless the
// leading minors are positive relative to the trace (the matrix is then not safely positive definite).
HSL_HD bool spd3_inverse(const double* H, double* W) {
  const double a = H[0], b = H[1], c = H[2], d = H[3], e = H[4], f = H[5];
  const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
  const double c11 = a * f - c * c, c12 = b * c - a * e, c22 = a * d - b * b;
  const double det = a * c00 + b * c01 + c * c02;
  const double id = hsl_rcp(det);
  W[0] = c00 * id; W[1] = c01 * id; W[2] = c02 * id;
  W[3] = c11 * id; W[4] = c12 * id; W[5] = c22 * id;
  const double tr = a + d + f, tol = 1e-13;
  return (a > tol * tr) && (c22 > tol * tr * tr) && (det > tol * tr * tr * tr);
}
HSL_HD void sym3_mul(const double* W, const double* x, double* y) {
  y[0] = W[0] * x[0] + W[1] * x[1] + W[2] * x[2];
  y[1] = W[1] * x[0] + W[3] * x[1] + W[4] * x[2];
  y[2] = W[2] * x[0] + W[4] * x[1] + W[5] * x[2];
}

template <int NF, int FB, int MODE, bool DUMP, class SM>
HSL_HD void phase_b_leg(const HslModelPod& M, const HslFrameArgs& A, const SM& sm, const HslSlot& sl, int limb,
                        HslLegState<DUMP>& st) {
  const HslLimb& L = M.limb[limb];
  double f[3][3], nn[3][3];
  if (MODE == HSL_MODE_FIELDS) {
#pragma unroll
    for (int h = 0; h < 3; h++) {
      const int b = L.h[h].body;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        f[h][k] = A.f_momrate[((int64_t)sl.i * M.n + b) * 3 + k];
        nn[h][k] = A.f_angrate[((int64_t)sl.i * M.n + b) * 3 + k];
      }
      f[h][2] += L.h[h].mass * M.g;
      st.qd[h] = 0;
    }
  } else {
    double hh, dt;
    if (MODE == HSL_MODE_GAIT) { hh = A.cand[sl.c].hh; dt = A.cand[sl.c].dt; }
    else { dt = A.dt_in[sl.c]; hh = 1. / (2 * dt); }
    double sdv[3], cdv[3], dang[3];
    bool slow[3];
#pragma unroll
    for (int h = 0; h < 3; h++) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        fd2_pair(sm.pu + ((3 * limb + h) * 3 + k) * FB, sl.s, hh, L.h[h].mass, L.h[h].inertia, &f[h][k], &nn[h][k]);
      }
      f[h][2] += L.h[h].mass * M.g;  // dynrec.cpp:293-297
      // joint rate, periodic.cpp:261-282: (q(s+1) - q(s-1)) wrapped to (-pi, pi], over 2 dt.  The wrapped
      // difference is the angle of the unit complex number e^{i q(s+1)} * conj(e^{i q(s-1)}).
      const double* csr = sm.cs + (6 * limb + 2 * h) * FB + sl.s;
      const double cp = csr[1], sp = csr[FB + 1], cm = csr[-1], sm_ = csr[FB - 1];
      sdv[h] = sp * cm - cp * sm_;
      cdv[h] = cp * cm + sp * sm_;
      dang[h] = hsl_small_angle(sdv[h], cdv[h], &slow[h]);
    }
    if (slow[0] || slow[1] || slow[2]) {  // large step between frames (rare): library arctangent
#pragma unroll
      for (int h = 0; h < 3; h++)
        if (slow[h]) dang[h] = atan2(sdv[h], cdv[h]);
    }
#pragma unroll
    for (int h = 0; h < 3; h++) st.qd[h] = dang[h] * hh;
  }
  // backward Newton-Euler recursion foot -> hip: F_j = f_j + F_child,
  // T_j = n_j + (pos_j - jpos_j) x f_j + T_child + (jpos_child - jpos_j) x F_child   (rows of B, dynrec.cpp:227-291)
  double F[3], T[3], d[3];
#pragma unroll
  for (int k = 0; k < 3; k++) { F[k] = 0; T[k] = 0; }
#pragma unroll
  for (int h = 2; h >= 0; h--) {
    if (h < 2) {
#pragma unroll
      for (int k = 0; k < 3; k++) d[k] = st.jpos[h + 1][k] - st.jpos[h][k];
      v3_cross_add(d, F, T);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) { d[k] = st.pos[h][k] - st.jpos[h][k]; T[k] += nn[h][k]; F[k] += f[h][k]; }
    v3_cross_add(d, f[h], T);
#pragma unroll
    for (int k = 0; k < 3; k++) { st.T[h][k] = T[k]; st.F[h][k] = F[k]; }
  }
  // limb wrench referred to the root reference point
  double ref[3];
  root_ref<NF, FB, MODE>(M, A, sm, sl, ref);
#pragma unroll
  for (int k = 0; k < 3; k++) d[k] = st.jpos[0][k] - ref[k];
  v3_cross_add(d, F, T);
  double* Pl = sm.part + (limb * SM::PART) * FB;
  double* P = Pl + sl.s;
  part_pair_store<FB>(Pl, sl.s, 0, F[0], F[1]);
  part_pair_store<FB>(Pl, sl.s, 1, F[2], T[0]);
  part_pair_store<FB>(Pl, sl.s, 2, T[1], T[2]);
  P[18 * FB] = st.contact ? 1.0 : 0.0;
  // motor torque of the particular solution and its sensitivity to the contact force:
  // tau_h = a_h . T_h - (a_h x rho_h) . lambda ,  rho_h = fpos - jpos_h     (periodic.cpp:328-343)
  double Hs[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
#pragma unroll
  for (int h = 0; h < 3; h++) {
    double rho[3];
#pragma unroll
    for (int k = 0; k < 3; k++) rho[k] = st.fpos[k] - st.jpos[h][k];
    st.taup[h] = v3_dot(st.axis[h], st.T[h]);
    v3_cross(st.axis[h], rho, st.w[h]);
    // level-1 cost sum_k (a_k (T_k - (rho x lambda)_k))^2, component-wise weights (ftsolver.cpp:239-246)
    const double ax = st.axis[h][0], ay = st.axis[h][1], az = st.axis[h][2];
    const double ux[3] = {0, -ax * rho[2], ax * rho[1]};
    const double uy[3] = {ay * rho[2], 0, -ay * rho[0]};
    const double uz[3] = {-az * rho[1], az * rho[0], 0};
    const double cx = ax * st.T[h][0], cy = ay * st.T[h][1], cz = az * st.T[h][2];
    Hs[0] += uy[0] * uy[0] + uz[0] * uz[0];
    Hs[1] += uz[0] * uz[1];
    Hs[2] += uy[0] * uy[2];
    Hs[3] += ux[1] * ux[1] + uz[1] * uz[1];
    Hs[4] += ux[1] * ux[2];
    Hs[5] += ux[2] * ux[2] + uy[2] * uy[2];
    g[0] -= cy * uy[0] + cz * uz[0];
    g[1] -= cx * ux[1] + cz * uz[1];
    g[2] -= cx * ux[2] + cy * uy[2];
  }
  if (st.contact) {
    if (!spd3_inverse(Hs, st.W)) st.bad |= HSL_ST_SOLVER;
    sym3_mul(st.W, g, st.Wg);
#pragma unroll
    for (int k = 0; k < 3; k++) st.r[k] = st.fpos[k] - ref[k];
    part_pair_store<FB>(Pl, sl.s, 3, st.W[0], st.W[1]);
    part_pair_store<FB>(Pl, sl.s, 4, st.W[2], st.W[3]);
    part_pair_store<FB>(Pl, sl.s, 5, st.W[4], st.W[5]);
    part_pair_store<FB>(Pl, sl.s, 6, st.Wg[0], st.Wg[1]);
    part_pair_store<FB>(Pl, sl.s, 7, st.Wg[2], st.r[0]);
    part_pair_store<FB>(Pl, sl.s, 8, st.r[1], st.r[2]);
  }
}

template <int NF, int FB, int MODE, class SM>
HSL_HD void phase_b_trunk(const HslModelPod& M, const HslFrameArgs& A, const SM& sm, const HslSlot& sl,
                          HslTrunkState& st) {
#pragma unroll
  for (int k = 0; k < 3; k++) { st.F0[k] = 0; st.T0[k] = 0; }
  double hh = 0;
  if (MODE == HSL_MODE_GAIT) hh = A.cand[sl.c].hh;
  if (MODE == HSL_MODE_TRAJ) hh = 1. / (2 * A.dt_in[sl.c]);
  double ref[3];
  root_ref<NF, FB, MODE>(M, A, sm, sl, ref);
  for (int tb = 0; tb < M.ntrunk; tb++) {
    double f[3], nn[3], pb[3], d[3];
    if (MODE == HSL_MODE_FIELDS) {
      const int b = M.trunk[tb].body;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        f[k] = A.f_momrate[((int64_t)sl.i * M.n + b) * 3 + k];
        nn[k] = A.f_angrate[((int64_t)sl.i * M.n + b) * 3 + k];
        pb[k] = A.f_pos[((int64_t)sl.i * M.n + b) * 3 + k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        f[k] = fd2(sm.tpos + (tb * 3 + k) * FB, sl.s, FB, hh, M.trunk[tb].mass);
        nn[k] = fd2(sm.tust + k * FB, sl.s, FB, hh, M.trunk[tb].inertia);
        pb[k] = sm.tpos[(tb * 3 + k) * FB + sl.s];
      }
    }
    f[2] += M.trunk[tb].mass * M.g;
#pragma unroll
    for (int k = 0; k < 3; k++) { d[k] = pb[k] - ref[k]; st.F0[k] += f[k]; st.T0[k] += nn[k]; }
    v3_cross_add(d, f, st.T0);
  }
}

// ------------------------------------------------------------------ phase C (trunk)
// LDL^T solve of a symmetric positive definite 6x6 (lower triangle in S[i][j], i>=j). Returns false on breakdown.
// *ill (optional) is set when a pivot falls below HSL_ILLCOND_PIVOT of the trace (nearly rank deficient).
HSL_HD bool spd6_solve(double S[6][6], double* b, bool* ill = nullptr) {
  double dinv[6], dd[6];
  double tr = 0;
#pragma unroll
  for (int i = 0; i < 6; i++) tr += S[i][i];
  const double tol = 1e-13 * tr, tol_ill = HSL_ILLCOND_PIVOT * tr;
  bool ok = true, weak = false;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    // column j: u_i = S_ij - sum_k L_ik (L_jk d_k) for i >= j ; d_j = u_j ; L_ij = u_i / d_j
    double ld[6];
#pragma unroll
    for (int k = 0; k < j; k++) ld[k] = S[j][k] * dd[k];
    double d = S[j][j];
#pragma unroll
    for (int k = 0; k < j; k++) d -= S[j][k] * ld[k];
    ok = ok && (d > tol);
    weak = weak || (d < tol_ill);
    dd[j] = d;
    dinv[j] = hsl_rcp(d);
#pragma unroll
    for (int i = j + 1; i < 6; i++) {
      double v = S[i][j];
#pragma unroll
      for (int k = 0; k < j; k++) v -= S[i][k] * ld[k];
      S[i][j] = v * dinv[j];
    }
  }
#pragma unroll
  for (int i = 0; i < 6; i++)
#pragma unroll
    for (int k = 0; k < i; k++) b[i] -= S[i][k] * b[k];
#pragma unroll
  for (int i = 0; i < 6; i++) b[i] *= dinv[i];
#pragma unroll
  for (int i = 5; i >= 0; i--)
#pragma unroll
    for (int k = i + 1; k < 6; k++) b[i] -= S[k][i] * b[k];
  if (ill) *ill = weak;
  return ok;
}

template <int NF, int FB, int MODE, bool DUMP, class SM>
HSL_HD int phase_c_trunk(const HslModelPod& M, const HslFrameArgs& A, const SM& sm, const HslSlot& sl,
                         HslTrunkState& st) {
  int bad = 0;
  double b[6] = {st.F0[0], st.F0[1], st.F0[2], st.T0[0], st.T0[1], st.T0[2]};
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
    const double* Pl = sm.part + (l * SM::PART) * FB;
    const double* P = Pl + sl.s;
    {
      const HslD2 e0 = part_pair_load<FB>(Pl, sl.s, 0), e1 = part_pair_load<FB>(Pl, sl.s, 1), e2 = part_pair_load<FB>(Pl, sl.s, 2);
      b[0] += e0.x; b[1] += e0.y; b[2] += e1.x; b[3] += e1.y; b[4] += e2.x; b[5] += e2.y;
    }
    if (P[18 * FB] != 0.0) {
      double W[6], Wg[3], r[3];
      {
        const HslD2 e3 = part_pair_load<FB>(Pl, sl.s, 3), e4 = part_pair_load<FB>(Pl, sl.s, 4), e5 = part_pair_load<FB>(Pl, sl.s, 5);
        const HslD2 e6 = part_pair_load<FB>(Pl, sl.s, 6), e7 = part_pair_load<FB>(Pl, sl.s, 7), e8 = part_pair_load<FB>(Pl, sl.s, 8);
        W[0] = e3.x; W[1] = e3.y; W[2] = e4.x; W[3] = e4.y; W[4] = e5.x; W[5] = e5.y;
        Wg[0] = e6.x; Wg[1] = e6.y; Wg[2] = e7.x; r[0] = e7.y; r[1] = e8.x; r[2] = e8.y;
      }
      if (nc == 0) { rA[0] = r[0]; rA[1] = r[1]; rA[2] = r[2]; }
      if (nc == 1) { rB[0] = r[0]; rB[1] = r[1]; rB[2] = r[2]; }
      nc++;
      // S += A W A^T with A = [I ; [r]x]:  K = [r]x W ; blocks [[W, K^T],[K, [r]x K^T]]
      const double Wc[3][3] = {{W[0], W[1], W[2]}, {W[1], W[3], W[4]}, {W[2], W[4], W[5]}};
      double K[3][3];  // K[i][j] = (r x W_col_j)_i
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
      // bottom-right: column j of [r]x K^T = r x (row j of K); symmetric, fill lower part
      {
        const double c0[3] = {r[1] * K[0][2] - r[2] * K[0][1], r[2] * K[0][0] - r[0] * K[0][2], r[0] * K[0][1] - r[1] * K[0][0]};
        const double c1[3] = {r[1] * K[1][2] - r[2] * K[1][1], r[2] * K[1][0] - r[0] * K[1][2], r[0] * K[1][1] - r[1] * K[1][0]};
        const double c2[3] = {r[1] * K[2][2] - r[2] * K[2][1], r[2] * K[2][0] - r[0] * K[2][2], r[0] * K[2][1] - r[1] * K[2][0]};
        S[3][3] += c0[0]; S[4][3] += c0[1]; S[5][3] += c0[2];
        S[4][4] += c1[1]; S[5][4] += c1[2];
        S[5][5] += c2[2];
      }
#pragma unroll
      for (int k = 0; k < 3; k++) v[k] += Wg[k];
      v3_cross_add(r, Wg, v + 3);
    }
  }
  double mu[6] = {0, 0, 0, 0, 0, 0};
  if (nc >= 2) {
    if (nc == 2) {
      // two contacts: level 0 has rank 5; the unreachable wrench direction is n = (rA x d, d), d = rA - rB
      // (torque about the line through both contact points).  Project it out of b and regularise S with it.
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
    bool ill = false;
    if (!spd6_solve(S, mu, &ill)) bad |= HSL_ST_SOLVER;
    if (ill) bad |= HSL_ST_ILLCOND;
  } else {
    bad |= HSL_ST_FEW_CONTACTS;
  }
#pragma unroll
  for (int k = 0; k < 6; k++) sm.mu[k * FB + sl.s] = mu[k];
  sm.mu[6 * FB + sl.s] = (nc >= 2) ? 1.0 : 0.0;

  if (DUMP && A.x) {
    // joint forces / torques of the trunk bodies (rows of x for bodies without a limb):
    // F_b = sum over the subtree of f - sum of contact forces in the subtree, same for torques about jpos_b.
    double lam[NF][3];
    bool con[NF];
#pragma unroll
    for (int l = 0; l < NF; l++) {
      const double* Pl = sm.part + (l * SM::PART) * FB;
      const double* P = Pl + sl.s;
      con[l] = (P[18 * FB] != 0.0) && (nc >= 2);
      lam[l][0] = lam[l][1] = lam[l][2] = 0;
      if (con[l]) {
        double W[6], y[3], r[3], Wy[3];
#pragma unroll
        for (int k = 0; k < 6; k++) W[k] = part_get<FB>(Pl, sl.s, 6 + k);
#pragma unroll
        for (int k = 0; k < 3; k++) r[k] = part_get<FB>(Pl, sl.s, 15 + k);
        v3_cross(mu + 3, r, y);
#pragma unroll
        for (int k = 0; k < 3; k++) y[k] += mu[k];
        sym3_mul(W, y, Wy);
#pragma unroll
        for (int k = 0; k < 3; k++) lam[l][k] = -(part_get<FB>(Pl, sl.s, 12 + k) + Wy[k]);
      }
    }
    double hh = 0;
    if (MODE == HSL_MODE_GAIT) hh = A.cand[sl.c].hh;
    if (MODE == HSL_MODE_TRAJ) hh = 1. / (2 * A.dt_in[sl.c]);
    double ref[3];
    root_ref<NF, FB, MODE>(M, A, sm, sl, ref);
    for (int tb = 0; tb < M.ntrunk; tb++) {
      // reference point: torso -> its COM (see root_ref) ; jointless body -> its frame origin (= jpos)
      double jp[3];
      if (tb == 0) { jp[0] = ref[0]; jp[1] = ref[1]; jp[2] = ref[2]; }
      else if (MODE == HSL_MODE_FIELDS) {
#pragma unroll
        for (int k = 0; k < 3; k++) jp[k] = A.f_jpos[((int64_t)sl.i * M.n + M.trunk[tb].body) * 3 + k];
      } else m3_affine(st.R0, M.trunk[tb].off, st.t0, jp);
      double Fb[3] = {0, 0, 0}, Tb[3] = {0, 0, 0};
      for (int t2 = 0; t2 < M.ntrunk; t2++) {  // trunk bodies in the subtree of tb
        int a = t2;
        while (a > tb) a = M.trunk[a].parent_trunk;
        if (a != tb) continue;
        double f[3], nn[3], pb[3], d[3];
        if (MODE == HSL_MODE_FIELDS) {
          const int bb = M.trunk[t2].body;
#pragma unroll
          for (int k = 0; k < 3; k++) {
            f[k] = A.f_momrate[((int64_t)sl.i * M.n + bb) * 3 + k];
            nn[k] = A.f_angrate[((int64_t)sl.i * M.n + bb) * 3 + k];
            pb[k] = A.f_pos[((int64_t)sl.i * M.n + bb) * 3 + k];
          }
        } else {
#pragma unroll
          for (int k = 0; k < 3; k++) {
            f[k] = fd2(sm.tpos + (t2 * 3 + k) * FB, sl.s, FB, hh, M.trunk[t2].mass);
            nn[k] = fd2(sm.tust + k * FB, sl.s, FB, hh, M.trunk[t2].inertia);
            pb[k] = sm.tpos[(t2 * 3 + k) * FB + sl.s];
          }
        }
        f[2] += M.trunk[t2].mass * M.g;
#pragma unroll
        for (int k = 0; k < 3; k++) { d[k] = pb[k] - jp[k]; Fb[k] += f[k]; Tb[k] += nn[k]; }
        v3_cross_add(d, f, Tb);
      }
#pragma unroll
      for (int l = 0; l < NF; l++) {  // limbs hanging in the subtree of tb
        int a = M.limb[l].attach;
        while (a > tb) a = M.trunk[a].parent_trunk;
        if (a != tb) continue;
        const double* Pl = sm.part + (l * SM::PART) * FB;
        double Fl[3], d[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { Fl[k] = part_get<FB>(Pl, sl.s, k); Fb[k] += Fl[k]; Tb[k] += part_get<FB>(Pl, sl.s, 3 + k); d[k] = ref[k] - jp[k]; }
        v3_cross_add(d, Fl, Tb);  // move the limb wrench from the root reference point to jp
        if (con[l]) {
          double rr[3];
#pragma unroll
          for (int k = 0; k < 3; k++) { rr[k] = part_get<FB>(Pl, sl.s, 15 + k) + ref[k] - jp[k]; Fb[k] -= lam[l][k]; }
          double cr[3];
          v3_cross(rr, lam[l], cr);
#pragma unroll
          for (int k = 0; k < 3; k++) Tb[k] -= cr[k];
        }
      }
      const int bb = M.trunk[tb].body;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        A.x[(int64_t)(3 * bb + k) * A.n_frames + sl.fo] = Fb[k];
        A.x[(int64_t)(3 * (M.n + bb) + k) * A.n_frames + sl.fo] = Tb[k];
      }
    }
  }
  return bad;
}

// ------------------------------------------------------------------ phase D (limb)
template <int NF, int FB, int MODE, bool DUMP, class SM>
HSL_HD void phase_d_leg(const HslModelPod& M, const HslFrameArgs& A, const SM& sm, const HslSlot& sl, int limb,
                        HslLegState<DUMP>& st) {
  const HslLimb& L = M.limb[limb];
  double lam[3] = {0, 0, 0};
  const bool con = st.contact && (sm.mu[6 * FB + sl.s] != 0.0);
  if (con) {
    double mu[6], y[3], Wy[3];
#pragma unroll
    for (int k = 0; k < 6; k++) mu[k] = sm.mu[k * FB + sl.s];
    v3_cross(mu + 3, st.r, y);  // A_c^T mu = mu_f + mu_t x r
#pragma unroll
    for (int k = 0; k < 3; k++) y[k] += mu[k];
    sym3_mul(st.W, y, Wy);
#pragma unroll
    for (int k = 0; k < 3; k++) lam[k] = -(st.Wg[k] + Wy[k]);
  }
  double work = 0, tau[3];
#pragma unroll
  for (int h = 0; h < 3; h++) {
    tau[h] = st.taup[h] - v3_dot(st.w[h], lam);
    const double dw = tau[h] * st.qd[h];
    work += (dw > 0) ? dw : 0;  // periodic.cpp:291-304
  }
  double cfz = 1e300, mu_f = -1e300;
  if (con) {  // periodic.cpp:347-357, over the feet that are on the ground
    cfz = lam[2];
    mu_f = hsl_div(hsl_sqrt(lam[0] * lam[0] + lam[1] * lam[1]), lam[2]);
  }
  double* P = sm.part + (limb * SM::PART) * FB + sl.s;
  P[0] = work;
  P[FB] = cfz;
  P[2 * FB] = mu_f;
  if (DUMP) {
    const int64_t nfr = A.n_frames, fo = sl.fo;
    if (A.tau) {
#pragma unroll
      for (int h = 0; h < 3; h++) A.tau[(int64_t)(3 * limb + h) * nfr + fo] = tau[h];
    }
    if (A.z) {
#pragma unroll
      for (int k = 0; k < 3; k++) A.z[(int64_t)(3 * limb + k) * nfr + fo] = lam[k];
    }
    if (A.contacts) A.contacts[(int64_t)limb * nfr + fo] = (uint8_t)st.contact;
    if (A.x) {
#pragma unroll
      for (int h = 0; h < 3; h++) {
        const int b = L.h[h].body;
        double rho[3], cr[3];
#pragma unroll
        for (int k = 0; k < 3; k++) rho[k] = st.fpos[k] - st.jpos[h][k];
        v3_cross(rho, lam, cr);
#pragma unroll
        for (int k = 0; k < 3; k++) {
          A.x[(int64_t)(3 * b + k) * nfr + fo] = st.F[h][k] - lam[k];
          A.x[(int64_t)(3 * (M.n + b) + k) * nfr + fo] = st.T[h][k] - cr[k];
        }
      }
    }
  }
}

// ------------------------------------------------------------------ phase E (trunk)
template <int NF, int FB, class SM>
HSL_HD void phase_e_trunk(const HslFrameArgs& A, const SM& sm, const HslSlot& sl) {
  double work = 0, cfz = 1e300, mu = -1e300;
#pragma unroll
  for (int l = 0; l < NF; l++) {
    const double* P = sm.part + (l * SM::PART) * FB + sl.s;
    work += P[0];
    cfz = fmin(cfz, P[FB]);
    mu = fmax(mu, P[2 * FB]);
  }
  if (A.wframe) A.wframe[sl.fo] = work;
  if (A.fmin_cfz) A.fmin_cfz[sl.fo] = cfz;
  if (A.fmax_mu) A.fmax_mu[sl.fo] = mu;
}

// ------------------------------------------------------------------ candidate setup (a1)
// pgssweeper::setup_pergen / partial_setup_pergen / setup_foot_shift / shift_pos0 (pergen.cpp:453-507),
// ------------------------------------------------------------------ record-level entries (a2, a3 on their own)
// pergensetup::set_rec (pergen.cpp:225-239): frame record of candidate cd at time t -- torso position and Euler
// angles, then the foot targets in LIK order.  One call per (record, role): role < nf writes that limb's target,
// role == nf the torso entries.
HSL_HD void gait_record(const HslFrameArgs& A, const HslCand& c, int nf, int role, double t, double* rec) {
  HslCandView cd;
  load_cand(c, role < nf ? role : -1, cd);
  if (role < nf) {
    double p[3];
    foot_target(cd, t, p);
    if (A.flags & HSL_FLAG_REC_TRANSFORM) rec_transform_point(A, p);
#pragma unroll
    for (int k = 0; k < 3; k++) rec[6 + 3 * role + k] = p[k];
  } else {
    double qt[3], R0[9], eul[3];
    torso_pose(cd, t, qt, R0, eul);
    if (A.flags & HSL_FLAG_REC_TRANSFORM) rec_transform_pose(A, qt, R0, eul);
#pragma unroll
    for (int k = 0; k < 3; k++) { rec[k] = qt[k]; rec[3 + k] = eul[k]; }
  }
}
// kinematicmodel::set_jvalues_with_lik / liksolver::place_limbs (model.cpp:354-359, lik.cpp:89-99,316-354): joint
// values of a record.  role < nf: the three hinge angles of that limb (false when out of reach); role == nf: the
// torso's six values.
HSL_HD bool ik_record(const HslModelPod& M, int role, const double* rec, bool ignore_reach, double* q) {
  if (role >= M.nf) {
#pragma unroll
    for (int k = 0; k < 6; k++) q[k] = rec[k];
    return true;
  }
  const HslLimb& L = M.limb[role];
  double R0[9], t0[3], pl[3], cq[3], sq[3], ang[3];
  euler_to_R(rec[3], rec[4], rec[5], R0);
  torso_frame(M, rec, R0, t0);
  foot_into_hip_frame(L, R0, t0, rec + 6 + 3 * role, pl);
  const bool ok = limb_ik<0>(L, pl, ignore_reach, cq, sq, ang);
#pragma unroll
  for (int h = 0; h < 3; h++) q[6 + 3 * role + h] = ang[h];
  return ok;
}

// kinematicmodel::set_jvalues + recompute_modelnodes (model.cpp:183-201,314-318,362-366): ground frames of the bodies
// and of their joints for given joint values q[config_dim], as column-major 4x4 `affine`s (matrix.cpp:138-146).
// role < nf: the three hinge bodies of that limb; role == nf: the torso and the jointless trunk bodies.
// A [n][16] = modelnode::A_ground, J [n][16] = modeljoint::A_ground (joint frame before the joint's own transform;
// zeros for bodies without a joint).  Either may be null.
HSL_HD void hsl_put_affine(double* dst, const double* R, const double* t) {
#pragma unroll
  for (int j = 0; j < 3; j++) {
#pragma unroll
    for (int i = 0; i < 3; i++) dst[4 * j + i] = R[3 * j + i];
    dst[4 * j + 3] = 0;
  }
  dst[12] = t[0]; dst[13] = t[1]; dst[14] = t[2]; dst[15] = 1;
}
HSL_HD void fk_record(const HslModelPod& M, int role, const double* q, double* Aout, double* Jout) {
  double R0[9], t0[3];
  euler_to_R(q[3], q[4], q[5], R0);
  torso_frame(M, q, R0, t0);
  if (role >= M.nf) {
    const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int tb = 0; tb < M.ntrunk; tb++) {
      double ob[3];
      m3_affine(R0, M.trunk[tb].off, t0, ob);
      const int b = M.trunk[tb].body;
      if (Aout) hsl_put_affine(Aout + 16 * b, R0, ob);
      if (Jout) {
        if (tb == 0) hsl_put_affine(Jout + 16 * b, I3, M.Pt);  // the free joint's A_ground = A_parent (model.cpp:158-161)
        else for (int k = 0; k < 16; k++) Jout[16 * b + k] = 0;
      }
    }
    return;
  }
  const HslLimb& L = M.limb[role];
  double Rp[9], tp[3], Rb[9], tb[3];
#pragma unroll
  for (int k = 0; k < 9; k++) Rp[k] = R0[k];
  m3_affine(R0, L.oatt, t0, tp);
  for (int h = 0; h < 3; h++) {
    double sn, cs, jpos[3], axis[3], com[3], ust[3], Rj[9];
    sincos(q[6 + 3 * role + h], &sn, &cs);
    m3_mul(Rp, L.h[h].Rjp, Rj);  // joint frame: parent frame times modeljoint::A_parent (model.cpp:64-66)
    hinge_fk(L.h[h], cs, sn, Rp, tp, Rb, tb, jpos, axis, com, ust);
    if (Aout) hsl_put_affine(Aout + 16 * L.h[h].body, Rb, tb);
    if (Jout) hsl_put_affine(Jout + 16 * L.h[h].body, Rj, jpos);
#pragma unroll
    for (int k = 0; k < 9; k++) Rp[k] = Rb[k];
#pragma unroll
    for (int k = 0; k < 3; k++) tp[k] = tb[k];
  }
}

// periodicgenerator::set_step_duration / compute_max_radius (pergen.cpp:30-51,144-154),
// periodic::record_trajectory time base (periodic.cpp:84-91).  p = 13 candidate scalars (include/hsl.h).
HSL_HD void setup_candidate(const HslModelPod& M, const double* p, int n_t, HslCand& cd, double* ttab) {
  cd.status = 0;
  cd.pad = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) { cd.tp0[k] = p[k]; cd.eul[k] = p[3 + k]; }
  euler_to_R(p[3], p[4], p[5], cd.R0);
  const double f = p[6];
  cd.period = p[7]; cd.step_length = p[8]; cd.step_height = p[9]; cd.curvature = p[10];
  const int shift_type = (int)p[11];
  const double shift_value = p[12];
  const int n = M.nf;
  if (!(f >= 0 && f <= 1) || !(cd.period > 0) || n_t < 1) cd.status |= HSL_ST_BAD_PARAMS;
  cd.t_step = f * (1. / 2 - 1. / n) + 1. / n;
  cd.v = cd.step_length / cd.period;
  cd.dt = cd.period / n_t;
  cd.hh = 1. / (2 * cd.dt);
  double t0[3];
  torso_frame(M, cd.tp0, cd.R0, t0);
  double lat[3] = {0, 0, 0};
  if (shift_type == 0) {  // A_ground(torso) * (0, shift, 0, 1): includes the torso translation (pergen.cpp:485-488)
    const double sh[3] = {0, shift_value, 0};
    m3_affine(cd.R0, sh, t0, lat);
  }
  const int jmax = n / 2, z = (jmax == 1) ? 1 : jmax - 1;
  cd.max_radius = 0;
  for (int i = 0; i < n; i++) {
    const HslLimb& L = M.limb[i];
    double ta[3], hp[3];
    m3_affine(cd.R0, L.oatt, t0, ta);
    m3_affine(cd.R0, L.h[0].tjp, ta, hp);  // hip position = origin of the top link's frame (lik.cpp:357-360)
    if (shift_type == 0) {
      const double sgn = (i % 2) ? -1.0 : 1.0;
#pragma unroll
      for (int k = 0; k < 3; k++) hp[k] += sgn * lat[k];
    } else if (shift_type == 1) {
      const double ff = shift_value / sqrt(hp[0] * hp[0] + hp[1] * hp[1]);
      hp[0] += hp[0] * ff;
      hp[1] += hp[1] * ff;
    }
    cd.pos0[i][0] = hp[0]; cd.pos0[i][1] = hp[1]; cd.pos0[i][2] = M.rcap;
    const int k = L.pg_index, grp = k / jmax, j = k % jmax;
    cd.ts[i] = j * (1. / 2 - cd.t_step) / z + double(grp) / 2;
    cd.xs[i] = cd.ts[i] + cd.t_step / 2 - 1. / 2;
    cd.turn_r[i] = 0; cd.turn_sa[i] = 0; cd.turn_ca[i] = 1;
    if (cd.curvature != 0) {
      const double dy = hp[1] - 1. / cd.curvature;
      const double rad = sqrt(hp[0] * hp[0] + dy * dy + M.rcap * M.rcap);
      if (rad > cd.max_radius) cd.max_radius = rad;
      // polar coordinates of the default foot position about the turning centre (0, 1/curvature)
      const double rr = sqrt(hp[0] * hp[0] + dy * dy);
      cd.turn_r[i] = rr;
      cd.turn_sa[i] = (rr > 0) ? dy / rr : 0.0;
      cd.turn_ca[i] = (rr > 0) ? hp[0] / rr : 1.0;
    }
  }
  for (int i = n; i < HSL_MAX_LIMBS; i++) {
    cd.pos0[i][0] = cd.pos0[i][1] = cd.pos0[i][2] = 0; cd.ts[i] = cd.xs[i] = 0;
    cd.turn_r[i] = 0; cd.turn_sa[i] = 0; cd.turn_ca[i] = 1;
  }
  if (ttab) {
    double t = 0;
    for (int i = 0; i < n_t + 4; i++) { ttab[i] = t; t += cd.dt; }
  }
}
