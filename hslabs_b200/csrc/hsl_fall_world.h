// One world of the fall / perturbation sweep: the per-thread code of hsl_fall_kernel (see hsl_fall.cu for the model it
// restates), host/device so that the CPU test tier can run the same source (tests/hostcheck) next to the reference's own
// player code on the ODE shim.
#pragma once
#include <math.h>
#include <stdint.h>

#include "hsl_fall.h"

#ifndef HSL_HD
#if defined(__CUDACC__)
#define HSL_HD __host__ __device__ __forceinline__
#else
#define HSL_HD inline
#endif
#endif
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace hsl_fall {

constexpr int NB = HSL_MAX_BODIES;
constexpr int MAXC = HSL_FALL_MAX_CONTACTS;
constexpr int MAXROWS = 6 * NB + 3 * MAXC;

HSL_HD void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
HSL_HD double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
HSL_HD void q_mul(const double* b, const double* c, double* a) {
  a[0] = b[0] * c[0] - b[1] * c[1] - b[2] * c[2] - b[3] * c[3];
  a[1] = b[0] * c[1] + b[1] * c[0] + b[2] * c[3] - b[3] * c[2];
  a[2] = b[0] * c[2] + b[2] * c[0] + b[3] * c[1] - b[1] * c[3];
  a[3] = b[0] * c[3] + b[3] * c[0] + b[1] * c[2] - b[2] * c[1];
}
// v (body axes) -> world: rotation by the unit quaternion q (the matrix of ODE's dQtoR applied to v)
HSL_HD void rot(const double* q, const double* v, double* o) {
  const double qq1 = 2 * q[1] * q[1], qq2 = 2 * q[2] * q[2], qq3 = 2 * q[3] * q[3];
  const double r00 = 1 - qq2 - qq3, r01 = 2 * (q[1] * q[2] - q[0] * q[3]), r02 = 2 * (q[1] * q[3] + q[0] * q[2]);
  const double r10 = 2 * (q[1] * q[2] + q[0] * q[3]), r11 = 1 - qq1 - qq3, r12 = 2 * (q[2] * q[3] - q[0] * q[1]);
  const double r20 = 2 * (q[1] * q[3] - q[0] * q[2]), r21 = 2 * (q[2] * q[3] + q[0] * q[1]), r22 = 1 - qq1 - qq2;
  o[0] = r00 * v[0] + r01 * v[1] + r02 * v[2];
  o[1] = r10 * v[0] + r11 * v[1] + r12 * v[2];
  o[2] = r20 * v[0] + r21 * v[1] + r22 * v[2];
}
HSL_HD void plane_space(const double* n, double* p, double* q) {  // ODE dPlaneSpace
  if (fabs(n[2]) > 0.70710678118654752440) {
    const double a = n[1] * n[1] + n[2] * n[2], k = 1.0 / sqrt(a);
    p[0] = 0; p[1] = -n[2] * k; p[2] = n[1] * k;
    q[0] = a * k; q[1] = -n[0] * p[2]; q[2] = n[0] * p[1];
  } else {
    const double a = n[0] * n[0] + n[1] * n[1], k = 1.0 / sqrt(a);
    p[0] = -n[1] * k; p[1] = n[0] * k; p[2] = 0;
    q[0] = -n[2] * p[1]; q[1] = n[2] * p[0]; q[2] = a * k;
  }
}

// Row kinds.  A row is (kind, owner, k): owner = joint index or contact index, k = component.
enum { R_BALL = 0, R_HANG = 1, R_FPOS = 2, R_FANG = 3, R_CONT = 4 };

struct World {
  // state
  double pos[NB][3], q[NB][4], lv[NB][3], av[NB][3];
  // per step
  double fe[NB][6], fc[NB][6];
  double ja1[NB][3], ja2[NB][3];   // hinge: world anchors a1, a2 ; fixed: ja1 = ofs
  double jp[NB][3], jq[NB][3];     // hinge: plane-space vectors of the axis
  double jax[NB][3];               // hinge: axis (world), for the motor torque
  double cd[MAXC][3][3], cc[MAXC][3];  // contact: directions (normal, t1, t2), lever arm c1
  int cbody[MAXC];
  double rhs[MAXROWS], Ad[MAXROWS], Adcfm[MAXROWS], lam[MAXROWS];
  unsigned char rkind[MAXROWS], rown[MAXROWS], rk[MAXROWS];
  short order[MAXROWS];
};

// J fc of one row and the two bodies it couples
HSL_HD double row_dot(const World& w, const HslSimPod& S, int i, int& b1, int& b2) {
  const int kind = w.rkind[i], o = w.rown[i], k = w.rk[i];
  double t[3];
  switch (kind) {
    case R_BALL: {
      b1 = S.joint[o].b1; b2 = S.joint[o].b2;
      cross3(&w.fc[b1][3], w.ja1[o], t);
      double s = w.fc[b1][k] + t[k];
      cross3(&w.fc[b2][3], w.ja2[o], t);
      return s - w.fc[b2][k] - t[k];
    }
    case R_HANG: {
      b1 = S.joint[o].b1; b2 = S.joint[o].b2;
      const double* u = k ? w.jq[o] : w.jp[o];
      return dot3(u, &w.fc[b1][3]) - dot3(u, &w.fc[b2][3]);
    }
    case R_FPOS: {
      b1 = S.joint[o].b1; b2 = S.joint[o].b2;
      cross3(w.ja1[o], &w.fc[b1][3], t);     // (e_k x ofs) . w = e_k . (ofs x w)
      return w.fc[b1][k] + t[k] - w.fc[b2][k];
    }
    case R_FANG: {
      b1 = S.joint[o].b1; b2 = S.joint[o].b2;
      return w.fc[b1][3 + k] - w.fc[b2][3 + k];
    }
    default: {
      b1 = w.cbody[o]; b2 = -1;
      cross3(w.cc[o], w.cd[o][k], t);
      return dot3(w.cd[o][k], &w.fc[b1][0]) + dot3(t, &w.fc[b1][3]);
    }
  }
}
// fc += M^-1 J^T delta
HSL_HD void row_apply(World& w, const HslSimPod& S, int i, double delta) {
  const int kind = w.rkind[i], o = w.rown[i], k = w.rk[i];
  double t[3], e[3] = {0, 0, 0};
  switch (kind) {
    case R_BALL: {
      const int b1 = S.joint[o].b1, b2 = S.joint[o].b2;
      const double im1 = 1.0 / S.body[b1].mass, ii1 = 1.0 / S.body[b1].inertia, im2 = 1.0 / S.body[b2].mass, ii2 = 1.0 / S.body[b2].inertia;
      e[k] = 1;
      w.fc[b1][k] += im1 * delta;
      cross3(w.ja1[o], e, t);                // J1a row k = a1 x e_k
      for (int c = 0; c < 3; c++) w.fc[b1][3 + c] += ii1 * t[c] * delta;
      w.fc[b2][k] -= im2 * delta;
      cross3(e, w.ja2[o], t);                // J2a row k = e_k x a2
      for (int c = 0; c < 3; c++) w.fc[b2][3 + c] += ii2 * t[c] * delta;
    } break;
    case R_HANG: {
      const int b1 = S.joint[o].b1, b2 = S.joint[o].b2;
      const double* u = k ? w.jq[o] : w.jp[o];
      const double ii1 = 1.0 / S.body[b1].inertia, ii2 = 1.0 / S.body[b2].inertia;
      for (int c = 0; c < 3; c++) { w.fc[b1][3 + c] += ii1 * u[c] * delta; w.fc[b2][3 + c] -= ii2 * u[c] * delta; }
    } break;
    case R_FPOS: {
      const int b1 = S.joint[o].b1, b2 = S.joint[o].b2;
      e[k] = 1;
      w.fc[b1][k] += delta / S.body[b1].mass;
      cross3(e, w.ja1[o], t);                // J1a row k = e_k x ofs
      for (int c = 0; c < 3; c++) w.fc[b1][3 + c] += t[c] * delta / S.body[b1].inertia;
      w.fc[b2][k] -= delta / S.body[b2].mass;
    } break;
    case R_FANG: {
      const int b1 = S.joint[o].b1, b2 = S.joint[o].b2;
      w.fc[b1][3 + k] += delta / S.body[b1].inertia;
      w.fc[b2][3 + k] -= delta / S.body[b2].inertia;
    } break;
    default: {
      const int b1 = w.cbody[o];
      cross3(w.cc[o], w.cd[o][k], t);
      for (int c = 0; c < 3; c++) { w.fc[b1][c] += w.cd[o][k][c] * delta / S.body[b1].mass; w.fc[b1][3 + c] += t[c] * delta / S.body[b1].inertia; }
    } break;
  }
}

// hinge angle of joint j (ODE getHingeAngle): rotation of body 1 against body 2 about axis1, relative to qrel
HSL_HD double hinge_angle(const World& w, const HslSimJoint& J) {
  double c1[4] = {w.q[J.b1][0], -w.q[J.b1][1], -w.q[J.b1][2], -w.q[J.b1][3]}, qq[4], cr[4] = {J.qrel[0], -J.qrel[1], -J.qrel[2], -J.qrel[3]}, qr[4];
  q_mul(c1, w.q[J.b2], qq);
  q_mul(qq, cr, qr);
  const double cost2 = qr[0], sint2 = sqrt(qr[1] * qr[1] + qr[2] * qr[2] + qr[3] * qr[3]);
  const double d = qr[1] * J.axis1[0] + qr[2] * J.axis1[1] + qr[3] * J.axis1[2];
  double th = (d >= 0) ? 2 * atan2(sint2, cost2) : 2 * atan2(sint2, -cost2);
  if (th > M_PI) th -= 2 * M_PI;
  return -th;
}


HSL_HD void fall_world(const HslSimPod& S, const HslFallArgs& A, int64_t wi, World& w) {
  const int n = S.n;
  for (int b = 0; b < n; b++) {
    for (int k = 0; k < 3; k++) { w.pos[b][k] = A.pos0[3 * b + k]; w.lv[b][k] = 0; w.av[b][k] = 0; }
    for (int k = 0; k < 4; k++) w.q[b][k] = A.quat0[4 * b + k];
  }
  const int kick_step = A.kick_step ? A.kick_step[wi] : -1;
  double kick[3] = {0, 0, 0};
  if (A.kick_dv) for (int k = 0; k < 3; k++) kick[k] = A.kick_dv[3 * wi + k];
  const double h = A.play_dt, fps = 1.0 / h, kerp = fps * A.erp, wcfm = A.cfm * fps, sor_w = 1.3;
  double play_t = A.play_t0;
  uint32_t seed = 0;  // ODE's dRand seed of a fresh process
  int fell = 0, status = 0, step = 0;
  double t_fall = 0;
  for (; step < A.n_steps; step++) {
    // fall_check (player.cpp:669-681)
    if (play_t >= A.tmin && w.pos[0][2] < A.hc) { fell = 1; t_fall = play_t; break; }
    // external forces: gravity, the kick, the position-control torques
    for (int b = 0; b < n; b++) {
      w.fe[b][0] = 0; w.fe[b][1] = 0; w.fe[b][2] = -S.body[b].mass * A.gravity;
      w.fe[b][3] = w.fe[b][4] = w.fe[b][5] = 0;
      for (int k = 0; k < 6; k++) w.fc[b][k] = 0;
    }
    const int tsi = (int)(play_t / h + .5);
    const int tm = tsi % A.n_t;
    const double* ctrl = A.ctrl + (size_t)tm * 3 * S.nmotor;  // q0 | dq0 | tau of the gait at this time step
    int m = 0;
    for (int j = 0; j < S.nj; j++) {
      const HslSimJoint& J = S.joint[j];
      const int b1 = J.b1, b2 = J.b2;
      if (J.kind == 1) {
        double ax2[3], bb[3];
        rot(w.q[b1], J.anchor1, w.ja1[j]);
        rot(w.q[b2], J.anchor2, w.ja2[j]);
        rot(w.q[b1], J.axis1, w.jax[j]);
        rot(w.q[b2], J.axis2, ax2);
        plane_space(w.jax[j], w.jp[j], w.jq[j]);
        cross3(w.jax[j], ax2, bb);
        // set_position_control_torques (player.cpp:393-432): tau = tau_ff - k (q - q0 wrapped) - 2 sqrt(k) (dq - dq0), k = 100
        const int mi = J.motor;
        double dq = dot3(w.jax[j], w.av[b1]) - dot3(w.jax[j], w.av[b2]);   // dJointGetHingeAngleRate
        double e = hinge_angle(w, J) - ctrl[mi];
        if (e > M_PI) e -= 2 * M_PI; else if (e <= -M_PI) e += 2 * M_PI;  // arrayops::modulus (core.cpp:120-131)
        double tau = ctrl[2 * S.nmotor + mi] + (-A.kp) * e + (-2.0 * sqrt(A.kp)) * (dq - ctrl[S.nmotor + mi]);
        for (int k = 0; k < 3; k++) { w.fe[b1][3 + k] += w.jax[j][k] * tau; w.fe[b2][3 + k] -= w.jax[j][k] * tau; }  // dJointAddHingeTorque
        for (int k = 0; k < 3; k++) {
          w.rkind[m] = R_BALL; w.rown[m] = (unsigned char)j; w.rk[m] = (unsigned char)k;
          w.rhs[m] = kerp * (w.ja2[j][k] + w.pos[b2][k] - w.ja1[j][k] - w.pos[b1][k]);
          m++;
        }
        w.rkind[m] = R_HANG; w.rown[m] = (unsigned char)j; w.rk[m] = 0; w.rhs[m] = kerp * dot3(bb, w.jp[j]); m++;
        w.rkind[m] = R_HANG; w.rown[m] = (unsigned char)j; w.rk[m] = 1; w.rhs[m] = kerp * dot3(bb, w.jq[j]); m++;
      } else {
        double ofs_w[3];
        rot(w.q[b1], J.offset, ofs_w);  // ofs = R1 * offset, offset = R1^T (p1 - p2) in the zero configuration
        for (int k = 0; k < 3; k++) w.ja1[j][k] = ofs_w[k];
        for (int k = 0; k < 3; k++) {
          w.rkind[m] = R_FPOS; w.rown[m] = (unsigned char)j; w.rk[m] = (unsigned char)k;
          w.rhs[m] = kerp * (w.pos[b2][k] - w.pos[b1][k] + ofs_w[k]);
          m++;
        }
        double c1[4] = {w.q[b1][0], -w.q[b1][1], -w.q[b1][2], -w.q[b1][3]}, qq[4], cr[4] = {J.qrel[0], -J.qrel[1], -J.qrel[2], -J.qrel[3]}, qe[4], e[3];
        q_mul(c1, w.q[b2], qq);
        q_mul(qq, cr, qe);
        if (qe[0] < 0) { qe[1] = -qe[1]; qe[2] = -qe[2]; qe[3] = -qe[3]; }
        rot(w.q[b1], qe + 1, e);
        for (int k = 0; k < 3; k++) { w.rkind[m] = R_FANG; w.rown[m] = (unsigned char)j; w.rk[m] = (unsigned char)k; w.rhs[m] = 2 * kerp * e[k]; m++; }
      }
    }
    if (step == kick_step) for (int k = 0; k < 3; k++) w.fe[0][k] += kick[k] * fps;   // kick_torso: f = dv / dt for one step
    // contacts with the ground plane z = 0 (nearCallback + dCollide, visualization.cpp:296-325): spheres and capsules
    int nc = 0;
    const int m_joint = m;
    for (int b = 0; b < n; b++) {
      const HslSimBody& sb = S.body[b];
      if (sb.geom == 0) continue;
      double p[3];
      if (sb.geom == 2) {
        double d[3] = {sb.p1[0] - sb.p0[0], sb.p1[1] - sb.p0[1], sb.p1[2] - sb.p0[2]}, az[3], e0[3];
        rot(w.q[b], d, az);
        // the deeper capping sphere: the end whose offset along the plane normal is lower
        rot(w.q[b], (az[2] > 0) ? sb.p0 : sb.p1, e0);
        for (int k = 0; k < 3; k++) p[k] = w.pos[b][k] + e0[k];
      } else {
        double e0[3];
        rot(w.q[b], sb.p0, e0);
        for (int k = 0; k < 3; k++) p[k] = w.pos[b][k] + e0[k];
      }
      const double depth = -p[2] + sb.radius;
      if (depth < 0) continue;
      if (nc >= MAXC) { status |= HSL_FALL_ST_CONTACT_OVERFLOW; continue; }
      const double nrm[3] = {0, 0, 1};
      w.cbody[nc] = b;
      for (int k = 0; k < 3; k++) { w.cd[nc][0][k] = nrm[k]; w.cc[nc][k] = p[k] - nrm[k] * sb.radius - w.pos[b][k]; }
      plane_space(nrm, w.cd[nc][1], w.cd[nc][2]);
      // normal row: push-out at ERP, bounce 0.5 when approaching faster than 0.1 (dxJointContact::getInfo2)
      double t[3];
      cross3(w.cc[nc], nrm, t);
      const double outgoing = dot3(nrm, w.lv[b]) + dot3(t, w.av[b]);
      double c = kerp * depth;
      if (-outgoing > A.bounce_vel) { const double nc2 = -A.bounce * outgoing; if (nc2 > c) c = nc2; }
      for (int k = 0; k < 3; k++) { w.rkind[m] = R_CONT; w.rown[m] = (unsigned char)nc; w.rk[m] = (unsigned char)k; w.rhs[m] = (k == 0) ? c : 0.0; m++; }
      nc++;
    }
    // rhs = (c/h - J (v/h + M^-1 fe)) Ad,  Ad = w / (J M^-1 J^T + cfm/h)    (quickstep.cpp, SOR_LCP)
    for (int b = 0; b < n; b++) {  // fc temporarily holds v/h + M^-1 fe
      for (int k = 0; k < 3; k++) { w.fc[b][k] = w.lv[b][k] * fps + w.fe[b][k] / S.body[b].mass; w.fc[b][3 + k] = w.av[b][k] * fps + w.fe[b][3 + k] / S.body[b].inertia; }
    }
    for (int i = 0; i < m; i++) {
      int b1, b2;
      const double acc = row_dot(w, S, i, b1, b2);
      const int kind = w.rkind[i], o = w.rown[i], k = w.rk[i];
      double diag, t[3], e[3] = {0, 0, 0};
      e[k % 3] = 1;
      double cfm = wcfm;
      switch (kind) {
        case R_BALL: {
          cross3(w.ja1[o], e, t); diag = 1.0 / S.body[b1].mass + dot3(t, t) / S.body[b1].inertia;
          cross3(e, w.ja2[o], t); diag += 1.0 / S.body[b2].mass + dot3(t, t) / S.body[b2].inertia;
        } break;
        case R_HANG: diag = 1.0 / S.body[b1].inertia + 1.0 / S.body[b2].inertia; break;
        case R_FPOS: cross3(e, w.ja1[o], t); diag = 1.0 / S.body[b1].mass + dot3(t, t) / S.body[b1].inertia + 1.0 / S.body[b2].mass; break;
        case R_FANG: diag = 1.0 / S.body[b1].inertia + 1.0 / S.body[b2].inertia; break;
        default: cross3(w.cc[o], w.cd[o][k], t); diag = 1.0 / S.body[b1].mass + dot3(t, t) / S.body[b1].inertia; if (k == 0) cfm = A.soft_cfm * fps; break;
      }
      const double Ad = sor_w / (diag + cfm);
      w.Ad[i] = Ad;
      w.Adcfm[i] = Ad * cfm;
      w.rhs[i] = (w.rhs[i] * fps - acc) * Ad;
      w.lam[i] = 0;
      w.order[i] = (short)i;
    }
    for (int b = 0; b < n; b++) for (int k = 0; k < 6; k++) w.fc[b][k] = 0;
    for (int it = 0; it < A.iterations; it++) {
      if ((it & 7) == 0) {
        for (int i = 1; i < m; i++) {
          seed = 1664525u * seed + 1013904223u;   // ODE dRand
          const int s = (int)((double)seed * ((double)(i + 1) / 4294967296.0));
          const short tmp = w.order[i]; w.order[i] = w.order[s]; w.order[s] = tmp;
        }
      }
      for (int oi = 0; oi < m; oi++) {
        const int i = w.order[oi];
        int b1, b2;
        double delta = w.rhs[i] - w.lam[i] * w.Adcfm[i] - w.Ad[i] * row_dot(w, S, i, b1, b2);
        double nl = w.lam[i] + delta;
        if (w.rkind[i] == R_CONT && w.rk[i] == 0 && nl < 0) { delta = -w.lam[i]; nl = 0; }  // lo = 0 on the normal row; every other row unbounded
        w.lam[i] = nl;
        row_apply(w, S, i, delta);
      }
    }
    (void)m_joint;
    // v += h (M^-1 fe + fc); x += h v; q += h/2 [0, w] q, renormalised (dxStepBody)
    for (int b = 0; b < n; b++) {
      for (int k = 0; k < 3; k++) {
        w.lv[b][k] += h * (w.fe[b][k] / S.body[b].mass + w.fc[b][k]);
        w.av[b][k] += h * (w.fe[b][3 + k] / S.body[b].inertia + w.fc[b][3 + k]);
      }
      for (int k = 0; k < 3; k++) w.pos[b][k] += h * w.lv[b][k];
      const double* wv = w.av[b];
      const double* q = w.q[b];
      const double dq[4] = {0.5 * (-wv[0] * q[1] - wv[1] * q[2] - wv[2] * q[3]), 0.5 * (wv[0] * q[0] + wv[1] * q[3] - wv[2] * q[2]),
                            0.5 * (-wv[0] * q[3] + wv[1] * q[0] + wv[2] * q[1]), 0.5 * (wv[0] * q[2] - wv[1] * q[1] + wv[2] * q[0])};
      double nq[4], l = 0;
      for (int k = 0; k < 4; k++) { nq[k] = q[k] + h * dq[k]; l += nq[k] * nq[k]; }
      l = 1.0 / sqrt(l);
      for (int k = 0; k < 4; k++) w.q[b][k] = nq[k] * l;
    }
    play_t += h;
    if (A.traj) for (int k = 0; k < 3; k++) A.traj[((size_t)wi * A.n_steps + step) * 3 + k] = w.pos[0][k];
  }
  if (A.fell) A.fell[wi] = (uint8_t)fell;
  if (A.t_end) A.t_end[wi] = fell ? t_fall : play_t;
  if (A.final_z) A.final_z[wi] = w.pos[0][2];
  if (A.steps_done) A.steps_done[wi] = step;
  if (A.status) A.status[wi] = status;
}

// One entry of the [n_t][3][nmotor] control table from the evaluated gait: target angle, target rate, feed-forward torque
// at time step tm (periodic::get_motor_adas / get_computed_torques, periodic.cpp:394-404, periodic.h:49; compute_vel_traj
// :261-282).  q: joint values of frames 0..n_t+3, element (frame f, component c) at q[f*qfs + c*qcs]; tau: motor torques of the
// solved frames 2..n_t+1, element (frame index i, motor j) at tau[i*tfs + j*tcs].
HSL_HD void fall_ctrl_entry(int n_t, int nmotor, double dt, const double* q, int64_t qfs, int64_t qcs, const double* tau, int64_t tfs, int64_t tcs,
                            int tm, int j, double* ctrl) {
  const int f = (tm < 2) ? tm + n_t : tm;            // tsi %= n_t; if (tsi < 2) tsi += n_t
  const double* qj = q + (int64_t)(6 + j) * qcs;
  double d = qj[(int64_t)(f + 1) * qfs] - qj[(int64_t)(f - 1) * qfs];
  if (d > M_PI) d -= 2 * M_PI; else if (d < -M_PI) d += 2 * M_PI;
  ctrl[(size_t)tm * 3 * nmotor + j] = qj[(int64_t)f * qfs];
  ctrl[(size_t)tm * 3 * nmotor + nmotor + j] = d / (2 * dt);
  ctrl[(size_t)tm * 3 * nmotor + 2 * nmotor + j] = tau[(int64_t)(f - 2) * tfs + (int64_t)j * tcs];
}

}  // namespace hsl_fall
