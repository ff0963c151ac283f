// Fall / perturbation sweep, second kernel: ONE WARP PER WORLD, lane = body, the world in registers.
//
// hsl_fall_kernel (hsl_fall_world.h: one thread per world) keeps a world's working set -- body state, constraint-force
// accumulators, joint geometry, four numbers per row -- in thread-local memory and is bound by the latency of that
// memory (5 % issue utilisation).  Here lane b of a warp owns body b and everything attached to it:
// its pose and velocities, its constraint-force accumulator fc_b, the geometry of the joint that ties it to its parent
// (anchors and plane-space vectors of a hinge, the offset of a fixed joint) and of its ground contact -- 40 doubles in
// registers.  The Gauss-Seidel sweep stays sequential over the rows in ODE's (re-shuffled) order, but a row update
// touches only two lanes: the owner fetches its parent's accumulator with shuffles, computes the update, and hands the
// parent its share back the same way.  Per-row scalars (rhs, Ad, lambda) and the visiting order live in shared memory
// (they are indexed by a run-time row number).  Row construction, the right-hand sides, contact detection and the
// integration run on all lanes at once.  Arithmetic and row numbering are those of hsl_fall_world.h (and hence of the
// CPU stepper oracle/shim/ode_step.cpp); tests/test_gpu_fall.py compares the two kernels world by world.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "hsl_fall_world.h"

namespace hsl_fall_warp {

using hsl_fall::cross3;
using hsl_fall::dot3;
using hsl_fall::plane_space;
using hsl_fall::q_mul;
using hsl_fall::rot;

constexpr unsigned FULL = 0xffffffffu;
constexpr int WARPS_PER_BLOCK = 4;
constexpr int MAXR = 6 * HSL_MAX_BODIES + 3 * HSL_FALL_MAX_CONTACTS;  // 240 rows at most

struct WorldSmem {
  double rhs[MAXR], Ad[MAXR], lam[MAXR];
  uint32_t sched[MAXR];       // the visiting order: row | owner lane << 8 | parent lane << 13 | case << 18, shuffled in place
  unsigned char swp[MAXR];    // the swap partners of one shuffle
};

__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ void shfl3(const double* v, int src, double* o) { o[0] = shfl(v[0], src); o[1] = shfl(v[1], src); o[2] = shfl(v[2], src); }
__device__ __forceinline__ double sel3(const double* v, int k) { return k == 0 ? v[0] : (k == 1 ? v[1] : v[2]); }
__device__ __forceinline__ void add_at(double* v, int k, double x) { v[0] += (k == 0) ? x : 0.0; v[1] += (k == 1) ? x : 0.0; v[2] += (k == 2) ? x : 0.0; }
__device__ __forceinline__ void unit(int k, double* e) { e[0] = (k == 0); e[1] = (k == 1); e[2] = (k == 2); }

// Row cases of the sweep: kind and component in one number, so that every case is straight-line code on named registers.
enum { C_BALL = 0, C_HANG = 3, C_FPOS = 5, C_FANG = 8, C_CONT = 11 };
__device__ __forceinline__ uint32_t pack(int row, int owner, int parent, int cs) { return (uint32_t)row | ((uint32_t)owner << 8) | ((uint32_t)parent << 13) | ((uint32_t)cs << 18); }

// What a lane keeps of its body during the sweep.  A joint row couples the lane that owns it (the child body) with the
// parent's lane: the owner reads the parent's accumulator components it needs through shuffles, updates lambda and its own
// accumulator, and hands the parent its finished increments (the owner knows the parent's inverse mass / inertia).
// The expressions are those of hsl_fall_world.h's row_dot / row_apply with the unit vector e_k multiplied out:
// (a x e_k) has components 0, a[k+2], -a[k+1] at k, k+1, k+2 (cyclic), (e_k x a) the negatives.
struct Lane {
  double fcl[3], fca[3];                 // constraint-force accumulator (M^-1 J^T lambda): linear, angular
  double ga[3], gb[3], gp[3], gq[3];     // hinge: anchors a1 (own), a2 (parent), plane-space vectors; fixed: ga = ofs
  double cc[3];                          // contact lever arm
  double im, ii, pim, pii;               // inverse mass / inertia: own, parent's
};

__device__ __forceinline__ double sor_delta(WorldSmem& W, int i, double cfm, double jf, double& lam) {
  const double Ad = W.Ad[i];
  lam = W.lam[i];
  return W.rhs[i] - lam * (Ad * cfm) - Ad * jf;
}

template <int K>
__device__ __forceinline__ void row_ball(Lane& L, WorldSmem& W, int i, int lane, int own, int po, double cfm) {
  constexpr int K1 = (K + 1) % 3, K2 = (K + 2) % 3;
  const double pl = shfl(L.fcl[K], po), pa1 = shfl(L.fca[K1], po), pa2 = shfl(L.fca[K2], po);
  double x0 = 0, x1 = 0, x2 = 0;
  if (lane == own) {
    const double t = L.fca[K1] * L.ga[K2] - L.fca[K2] * L.ga[K1];
    const double tp = pa1 * L.gb[K2] - pa2 * L.gb[K1];
    double lam;
    const double delta = sor_delta(W, i, cfm, (L.fcl[K] + t) - pl - tp, lam);
    W.lam[i] = lam + delta;
    L.fcl[K] += L.im * delta;
    L.fca[K1] += L.ii * L.ga[K2] * delta;
    L.fca[K2] -= L.ii * L.ga[K1] * delta;
    x0 = L.pim * delta; x1 = L.pii * L.gb[K2] * delta; x2 = L.pii * L.gb[K1] * delta;
  }
  x0 = shfl(x0, own); x1 = shfl(x1, own); x2 = shfl(x2, own);
  if (lane == po) { L.fcl[K] -= x0; L.fca[K1] -= x1; L.fca[K2] += x2; }
}
__device__ __forceinline__ void row_hang(Lane& L, const double (&u)[3], WorldSmem& W, int i, int lane, int own, int po, double cfm) {
  double pa[3], x[3] = {0, 0, 0};
  shfl3(L.fca, po, pa);
  if (lane == own) {
    double lam;
    const double delta = sor_delta(W, i, cfm, dot3(u, L.fca) - dot3(u, pa), lam);
    W.lam[i] = lam + delta;
#pragma unroll
    for (int c = 0; c < 3; c++) { L.fca[c] += L.ii * u[c] * delta; x[c] = L.pii * u[c] * delta; }
  }
  double y[3];
  shfl3(x, own, y);
  if (lane == po) { L.fca[0] -= y[0]; L.fca[1] -= y[1]; L.fca[2] -= y[2]; }
}
template <int K>
__device__ __forceinline__ void row_fpos(Lane& L, WorldSmem& W, int i, int lane, int own, int po, double cfm) {
  constexpr int K1 = (K + 1) % 3, K2 = (K + 2) % 3;
  const double pl = shfl(L.fcl[K], po), pa1 = shfl(L.fca[K1], po), pa2 = shfl(L.fca[K2], po);
  double x0 = 0, x1 = 0, x2 = 0;
  if (lane == own) {   // owner = b2 (the child), parent = b1; ga = ofs
    const double t = L.ga[K1] * pa2 - L.ga[K2] * pa1;
    double lam;
    const double delta = sor_delta(W, i, cfm, pl + t - L.fcl[K], lam);
    W.lam[i] = lam + delta;
    L.fcl[K] -= delta * L.im;
    x0 = delta * L.pim; x1 = L.ga[K2] * delta * L.pii; x2 = L.ga[K1] * delta * L.pii;
  }
  x0 = shfl(x0, own); x1 = shfl(x1, own); x2 = shfl(x2, own);
  if (lane == po) { L.fcl[K] += x0; L.fca[K1] -= x1; L.fca[K2] += x2; }
}
template <int K>
__device__ __forceinline__ void row_fang(Lane& L, WorldSmem& W, int i, int lane, int own, int po, double cfm) {
  const double pa = shfl(L.fca[K], po);
  double x0 = 0;
  if (lane == own) {
    double lam;
    const double delta = sor_delta(W, i, cfm, pa - L.fca[K], lam);
    W.lam[i] = lam + delta;
    L.fca[K] -= delta * L.ii;
    x0 = delta * L.pii;
  }
  x0 = shfl(x0, own);
  if (lane == po) L.fca[K] += x0;
}
// contact with the plane z = 0: directions n = (0,0,1), t1 = (0,-1,0), t2 = (1,0,0) (dPlaneSpace of n); one body, no exchange
template <int K>
__device__ __forceinline__ void row_cont(Lane& L, WorldSmem& W, int i, double cfm) {
  double jf;
  if (K == 0) jf = L.fcl[2] + (L.cc[1] * L.fca[0] - L.cc[0] * L.fca[1]);
  else if (K == 1) jf = -L.fcl[1] + (L.cc[2] * L.fca[0] - L.cc[0] * L.fca[2]);
  else jf = L.fcl[0] + (L.cc[2] * L.fca[1] - L.cc[1] * L.fca[2]);
  double lam;
  double delta = sor_delta(W, i, cfm, jf, lam);
  double nl = lam + delta;
  if (K == 0 && nl < 0) { delta = -lam; nl = 0; }   // lo = 0 on the normal row; friction rows unbounded (mu = infinity)
  W.lam[i] = nl;
  if (K == 0) { L.fcl[2] += delta * L.im; L.fca[0] += L.cc[1] * delta * L.ii; L.fca[1] -= L.cc[0] * delta * L.ii; }
  else if (K == 1) { L.fcl[1] -= delta * L.im; L.fca[0] += L.cc[2] * delta * L.ii; L.fca[2] -= L.cc[0] * delta * L.ii; }
  else { L.fcl[0] += delta * L.im; L.fca[1] += L.cc[2] * delta * L.ii; L.fca[2] -= L.cc[1] * delta * L.ii; }
}

// Blocks of 4 warps, HSL_FALL_MINB of them per SM.  Measured on 32768 hexapod worlds x 50 steps (world-steps/s):
//   2 blocks (255 registers, no spills) 1.77e6 | 3 (168) 2.07e6 | 4 (128) 2.56e6 | 5 (96) 2.92e6 | 6 (80) 3.21e6 | 7 (72) 3.30e6 |
//   8 (64) 3.33e6.
// A row update is a chain of dependent shuffles, shared-memory loads and FP64 operations, so warps in flight buy more than
// registers do -- the state the sweep does not touch (pose, velocities, external forces) is what ptxas spills -- until
// the issue slots fill (~75 % at 6..8 blocks; ~100 warp instructions per row update, 2.5e5 per world-step, profiles/r02_optimisation_log.md).
#ifndef HSL_FALL_MINB
#define HSL_FALL_MINB 6
#endif
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK, HSL_FALL_MINB)
hsl_fall_warp_kernel(const __grid_constant__ HslSimPod S, const __grid_constant__ HslFallArgs A) {
  __shared__ WorldSmem smem[WARPS_PER_BLOCK];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t wi = (int64_t)blockIdx.x * WARPS_PER_BLOCK + wib;
  if (wi >= A.n_worlds) return;  // whole warps leave together
  WorldSmem& W = smem[wib];
  const int n = S.n;
  const bool body = lane < n;
  const int b = body ? lane : 0;
  // ---- constants of this lane: body, the joint to its parent (joint index b-1 in body order), row numbering
  const HslSimBody sb = S.body[b];
  const double im = 1.0 / sb.mass, ii = 1.0 / sb.inertia;
  const bool has_joint = body && lane >= 1;
  const HslSimJoint J = S.joint[has_joint ? lane - 1 : 0];
  const int jkind = has_joint ? J.kind : 0;                  // 1 hinge (this lane = b1, parent = b2), 2 fixed (parent = b1, this lane = b2)
  const int par = has_joint ? (jkind == 1 ? J.b2 : J.b1) : 0;
  const int nrows_j = (jkind == 1) ? 5 : (jkind == 2 ? 6 : 0);
  int rbase = nrows_j;                                       // exclusive prefix sum over lanes = first row of this lane's joint
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, rbase, o); if (lane >= o) rbase += t; }
  const int m_joint = __shfl_sync(FULL, rbase, 31);
  rbase -= nrows_j;
  const double pmass = shfl(sb.mass, par), pinertia = shfl(sb.inertia, par);   // the parent's mass / inertia
  const double pim = 1.0 / pmass, pii = 1.0 / pinertia;
  // ---- state
  double pos[3], q[4], lv[3] = {0, 0, 0}, av[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < 3; k++) pos[k] = A.pos0[3 * b + k];
#pragma unroll
  for (int k = 0; k < 4; k++) q[k] = A.quat0[4 * b + k];
  const int kick_step = A.kick_step ? A.kick_step[wi] : -1;
  double kick[3] = {0, 0, 0};
  if (A.kick_dv) for (int k = 0; k < 3; k++) kick[k] = A.kick_dv[3 * wi + k];
  const double h = A.play_dt, fps = 1.0 / h, kerp = fps * A.erp, wcfm = A.cfm * fps, scfm = A.soft_cfm * fps, sor_w = 1.3;
  double play_t = A.play_t0;
  uint32_t seed = 0, lcg_a = 1, lcg_c = 0;   // ODE's dRand: seed' = a seed + c; after lane + 1 draws: lcg_a seed + lcg_c
  for (int j = 0; j <= lane; j++) { lcg_c = 1664525u * lcg_c + 1013904223u; lcg_a *= 1664525u; }
  int fell = 0, status = 0, step = 0;
  double t_fall = 0;
  for (; step < A.n_steps; step++) {
    const double z0 = shfl(pos[2], 0);
    if (play_t >= A.tmin && z0 < A.hc) { fell = 1; t_fall = play_t; break; }   // fall_check (warp-uniform)
    double fel[3] = {0, 0, -sb.mass * A.gravity}, fea[3] = {0, 0, 0};          // external force / torque on this body
    Lane L;
    L.im = im; L.ii = ii; L.pim = pim; L.pii = pii;
    double (&fcl)[3] = L.fcl, (&fca)[3] = L.fca, (&ga)[3] = L.ga, (&gb)[3] = L.gb, (&gp)[3] = L.gp, (&gq)[3] = L.gq, (&cc)[3] = L.cc;
#pragma unroll
    for (int k = 0; k < 3; k++) { fcl[k] = 0; fca[k] = 0; ga[k] = 0; gb[k] = 0; gp[k] = 0; gq[k] = 0; cc[k] = 0; }
    // ---- joint rows of this lane (all joints at once): geometry, control torque, right-hand sides
    double ppos[3], pq[4], pav[3];
    shfl3(pos, par, ppos);
    pq[0] = shfl(q[0], par); pq[1] = shfl(q[1], par); pq[2] = shfl(q[2], par); pq[3] = shfl(q[3], par);
    shfl3(av, par, pav);
    double crow[6] = {0, 0, 0, 0, 0, 0};                                               // c of this joint's rows
    double tq[3] = {0, 0, 0};                                                           // torque this lane's hinge puts on its parent (negated below)
    const int tsi = (int)(play_t / h + .5), tm = tsi % A.n_t;
    const double* ctrl = A.ctrl + (size_t)tm * 3 * S.nmotor;
    if (jkind == 1) {
      double ax1[3], ax2[3], bb[3];
      rot(q, J.anchor1, ga);
      rot(pq, J.anchor2, gb);
      rot(q, J.axis1, ax1);
      rot(pq, J.axis2, ax2);
      plane_space(ax1, gp, gq);
      cross3(ax1, ax2, bb);
      // hinge angle (ODE getHingeAngle) and rate, PD torque (player.cpp:393-432)
      double c1[4] = {q[0], -q[1], -q[2], -q[3]}, qq[4], cr[4] = {J.qrel[0], -J.qrel[1], -J.qrel[2], -J.qrel[3]}, qr[4];
      q_mul(c1, pq, qq);
      q_mul(qq, cr, qr);
      const double cost2 = qr[0], sint2 = sqrt(qr[1] * qr[1] + qr[2] * qr[2] + qr[3] * qr[3]);
      const double dd = qr[1] * J.axis1[0] + qr[2] * J.axis1[1] + qr[3] * J.axis1[2];
      double th = (dd >= 0) ? 2 * atan2(sint2, cost2) : 2 * atan2(sint2, -cost2);
      if (th > M_PI) th -= 2 * M_PI;
      const int mi = J.motor;
      const double dq = dot3(ax1, av) - dot3(ax1, pav);
      double e = -th - ctrl[mi];
      if (e > M_PI) e -= 2 * M_PI; else if (e <= -M_PI) e += 2 * M_PI;
      const double tau = ctrl[2 * S.nmotor + mi] + (-A.kp) * e + (-2.0 * sqrt(A.kp)) * (dq - ctrl[S.nmotor + mi]);
#pragma unroll
      for (int k = 0; k < 3; k++) { tq[k] = ax1[k] * tau; fea[k] += tq[k]; }
#pragma unroll
      for (int k = 0; k < 3; k++) crow[k] = kerp * (gb[k] + ppos[k] - ga[k] - pos[k]);
      crow[3] = kerp * dot3(bb, gp);
      crow[4] = kerp * dot3(bb, gq);
    } else if (jkind == 2) {
      rot(pq, J.offset, ga);  // ofs = R1 * offset (b1 = parent)
#pragma unroll
      for (int k = 0; k < 3; k++) crow[k] = kerp * (pos[k] - ppos[k] + ga[k]);
      double c1[4] = {pq[0], -pq[1], -pq[2], -pq[3]}, qq[4], cr[4] = {J.qrel[0], -J.qrel[1], -J.qrel[2], -J.qrel[3]}, qe[4], e3[3];
      q_mul(c1, q, qq);
      q_mul(qq, cr, qe);
      if (qe[0] < 0) { qe[1] = -qe[1]; qe[2] = -qe[2]; qe[3] = -qe[3]; }
      rot(pq, qe + 1, e3);
#pragma unroll
      for (int k = 0; k < 3; k++) crow[3 + k] = 2 * kerp * e3[k];
    }
    // the reaction of every hinge torque on the parent body, children in ascending order (the order hsl_fall_world.h adds them)
    for (int c = 1; c < n; c++) {
      double t3[3];
      shfl3(tq, c, t3);
      const int pc = __shfl_sync(FULL, par, c);
      if (lane == pc) { fea[0] -= t3[0]; fea[1] -= t3[1]; fea[2] -= t3[2]; }
    }
    if (lane == 0 && step == kick_step) { fel[0] += kick[0] * fps; fel[1] += kick[1] * fps; fel[2] += kick[2] * fps; }
    // ---- contact of this body with the ground plane z = 0
    bool hasc = false;
    double cnorm = 0;
    if (body && sb.geom != 0) {
      double e0[3], p[3];
      if (sb.geom == 2) {
        const double d[3] = {sb.p1[0] - sb.p0[0], sb.p1[1] - sb.p0[1], sb.p1[2] - sb.p0[2]};
        double az[3];
        rot(q, d, az);
        rot(q, (az[2] > 0) ? sb.p0 : sb.p1, e0);
      } else {
        rot(q, sb.p0, e0);
      }
#pragma unroll
      for (int k = 0; k < 3; k++) p[k] = pos[k] + e0[k];
      const double depth = -p[2] + sb.radius;
      if (depth >= 0) {
        hasc = true;
        cc[0] = p[0] - pos[0]; cc[1] = p[1] - pos[1]; cc[2] = p[2] - sb.radius - pos[2];
        const double nrm[3] = {0, 0, 1};
        double t[3];
        cross3(cc, nrm, t);
        const double outgoing = dot3(nrm, lv) + dot3(t, av);
        cnorm = kerp * depth;
        if (-outgoing > A.bounce_vel) { const double nc2 = -A.bounce * outgoing; if (nc2 > cnorm) cnorm = nc2; }
      }
    }
    const unsigned cmask = __ballot_sync(FULL, hasc);
    int cidx = __popc(cmask & ((1u << lane) - 1));
    if (hasc && cidx >= HSL_FALL_MAX_CONTACTS) { hasc = false; status |= HSL_FALL_ST_CONTACT_OVERFLOW; }
    int ncont = __popc(cmask);
    if (ncont > HSL_FALL_MAX_CONTACTS) ncont = HSL_FALL_MAX_CONTACTS;
    const int m = m_joint + 3 * ncont;
    const int cbase = m_joint + 3 * cidx;
    // plane space of the normal (0,0,1): t1 = (0,-1,0), t2 = (1,0,0)   (dPlaneSpace)
    const double cdir[3][3] = {{0, 0, 1}, {0, -1, 0}, {1, 0, 0}};
    // ---- rhs = (c/h - J (v/h + M^-1 fe)) Ad, Ad = w / (J M^-1 J^T + cfm/h); lambda = 0
    double vl[3], va[3], pvl[3], pva[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { vl[k] = lv[k] * fps + fel[k] / sb.mass; va[k] = av[k] * fps + fea[k] / sb.inertia; }
    shfl3(vl, par, pvl);
    shfl3(va, par, pva);
    if (jkind == 1) {
      double t1[3], t2[3];
      cross3(va, ga, t1);
      cross3(pva, gb, t2);
#pragma unroll
      for (int k = 0; k < 3; k++) {
        double e[3], u[3];
        unit(k, e);
        const double acc = (vl[k] + t1[k]) - pvl[k] - t2[k];
        cross3(ga, e, u);
        double diag = im + dot3(u, u) / sb.inertia;
        cross3(e, gb, u);
        diag += pim + dot3(u, u) / pinertia;
        const double Ad = sor_w / (diag + wcfm);
        W.Ad[rbase + k] = Ad; W.rhs[rbase + k] = (crow[k] * fps - acc) * Ad; W.lam[rbase + k] = 0;
        W.sched[rbase + k] = pack(rbase + k, lane, par, C_BALL + k);
      }
#pragma unroll
      for (int r = 0; r < 2; r++) {
        const double* u = r ? gq : gp;
        const double acc = dot3(u, va) - dot3(u, pva);
        const double Ad = sor_w / ((ii + pii) + wcfm);
        W.Ad[rbase + 3 + r] = Ad; W.rhs[rbase + 3 + r] = (crow[3 + r] * fps - acc) * Ad; W.lam[rbase + 3 + r] = 0;
        W.sched[rbase + 3 + r] = pack(rbase + 3 + r, lane, par, C_HANG + r);
      }
    } else if (jkind == 2) {
      double t1[3];
      cross3(ga, pva, t1);   // (e_k x ofs) . w = e_k . (ofs x w), w of b1 = the parent
#pragma unroll
      for (int k = 0; k < 3; k++) {
        double e[3], u[3];
        unit(k, e);
        const double acc = pvl[k] + t1[k] - vl[k];
        cross3(e, ga, u);
        const double diag = pim + dot3(u, u) / pinertia + im;
        const double Ad = sor_w / (diag + wcfm);
        W.Ad[rbase + k] = Ad; W.rhs[rbase + k] = (crow[k] * fps - acc) * Ad; W.lam[rbase + k] = 0;
        W.sched[rbase + k] = pack(rbase + k, lane, par, C_FPOS + k);
      }
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const double acc = pva[k] - va[k];
        const double Ad = sor_w / ((pii + ii) + wcfm);
        W.Ad[rbase + 3 + k] = Ad; W.rhs[rbase + 3 + k] = (crow[3 + k] * fps - acc) * Ad; W.lam[rbase + 3 + k] = 0;
        W.sched[rbase + 3 + k] = pack(rbase + 3 + k, lane, par, C_FANG + k);
      }
    }
    if (hasc) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        double t[3];
        cross3(cc, cdir[k], t);
        const double acc = dot3(cdir[k], vl) + dot3(t, va);
        const double diag = im + dot3(t, t) / sb.inertia;
        const double cfm = (k == 0) ? scfm : wcfm;
        const double Ad = sor_w / (diag + cfm);
        W.Ad[cbase + k] = Ad; W.rhs[cbase + k] = (((k == 0) ? cnorm : 0.0) * fps - acc) * Ad; W.lam[cbase + k] = 0;
        W.sched[cbase + k] = pack(cbase + k, lane, lane, C_CONT + k);
      }
    }
    __syncwarp();
    // ---- SOR projected Gauss-Seidel over the rows in ODE's order.  Inside the sweep x / mass is written x * (1 / mass):
    // the same bits for the unit masses and inertias every reference body has (dynrec.cpp:62-68 never sets a mass), one
    // rounding apart otherwise, and no division routine on the sequential path.
    for (int it = 0; it < A.iterations; it++) {
      if ((it & 7) == 0) {
        // ODE's shuffle: for i = 1 .. m-1 swap order[i] with order[dRandInt(i + 1)].  The m - 1 draws of the LCG come from all
        // lanes at once (jump-ahead constants per lane), the swaps themselves are a chain and stay on lane 0.
        for (int base = 0; base + 1 < m; base += 32) {
          const int i = base + lane + 1;
          const uint32_t si = lcg_a * seed + lcg_c;                       // the seed after lane + 1 more draws
          if (i < m) W.swp[i] = (unsigned char)(int)((double)si * ((double)(i + 1) / 4294967296.0));
          const int last = (m - 1 - base < 32) ? (m - 2 - base) : 31;    // lane that made the last draw of this round
          seed = __shfl_sync(FULL, si, last);
        }
        __syncwarp();
        if (lane == 0) {
          for (int i = 1; i < m; i++) {
            const int s = W.swp[i];
            const uint32_t tmp = W.sched[i]; W.sched[i] = W.sched[s]; W.sched[s] = tmp;
          }
        }
        __syncwarp();
      }
      for (int oi = 0; oi < m; oi++) {
        const uint32_t d = W.sched[oi];
        const int i = d & 255, own = (d >> 8) & 31, po = (d >> 13) & 31;
        switch (d >> 18) {
          case C_BALL + 0: row_ball<0>(L, W, i, lane, own, po, wcfm); break;
          case C_BALL + 1: row_ball<1>(L, W, i, lane, own, po, wcfm); break;
          case C_BALL + 2: row_ball<2>(L, W, i, lane, own, po, wcfm); break;
          case C_HANG + 0: row_hang(L, L.gp, W, i, lane, own, po, wcfm); break;
          case C_HANG + 1: row_hang(L, L.gq, W, i, lane, own, po, wcfm); break;
          case C_FPOS + 0: row_fpos<0>(L, W, i, lane, own, po, wcfm); break;
          case C_FPOS + 1: row_fpos<1>(L, W, i, lane, own, po, wcfm); break;
          case C_FPOS + 2: row_fpos<2>(L, W, i, lane, own, po, wcfm); break;
          case C_FANG + 0: row_fang<0>(L, W, i, lane, own, po, wcfm); break;
          case C_FANG + 1: row_fang<1>(L, W, i, lane, own, po, wcfm); break;
          case C_FANG + 2: row_fang<2>(L, W, i, lane, own, po, wcfm); break;
          case C_CONT + 0: if (lane == own) row_cont<0>(L, W, i, scfm); break;
          case C_CONT + 1: if (lane == own) row_cont<1>(L, W, i, wcfm); break;
          default: if (lane == own) row_cont<2>(L, W, i, wcfm); break;
        }
      }
    }
    __syncwarp();
    // ---- v += h (M^-1 fe + fc); x += h v; q += h/2 [0, w] q, renormalised
    if (body) {
#pragma unroll
      for (int k = 0; k < 3; k++) { lv[k] += h * (fel[k] / sb.mass + fcl[k]); av[k] += h * (fea[k] / sb.inertia + fca[k]); }
#pragma unroll
      for (int k = 0; k < 3; k++) pos[k] += h * lv[k];
      const double dq4[4] = {0.5 * (-av[0] * q[1] - av[1] * q[2] - av[2] * q[3]), 0.5 * (av[0] * q[0] + av[1] * q[3] - av[2] * q[2]),
                             0.5 * (-av[0] * q[3] + av[1] * q[0] + av[2] * q[1]), 0.5 * (av[0] * q[2] - av[1] * q[1] + av[2] * q[0])};
      double nq[4], l = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) { nq[k] = q[k] + h * dq4[k]; l += nq[k] * nq[k]; }
      l = 1.0 / sqrt(l);
#pragma unroll
      for (int k = 0; k < 4; k++) q[k] = nq[k] * l;
    }
    play_t += h;
    if (A.traj && lane == 0) for (int k = 0; k < 3; k++) A.traj[((size_t)wi * A.n_steps + step) * 3 + k] = pos[k];
  }
  status = __reduce_or_sync(FULL, status);
  if (lane == 0) {
    if (A.fell) A.fell[wi] = (uint8_t)fell;
    if (A.t_end) A.t_end[wi] = fell ? t_fall : play_t;
    if (A.final_z) A.final_z[wi] = pos[2];
    if (A.steps_done) A.steps_done[wi] = step;
    if (A.status) A.status[wi] = status;
  }
}

}  // namespace hsl_fall_warp
