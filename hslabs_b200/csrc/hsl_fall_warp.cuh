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
  unsigned char kind[MAXR], owner[MAXR], comp[MAXR];
  short order[MAXR];
};

__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ void shfl3(const double* v, int src, double* o) { o[0] = shfl(v[0], src); o[1] = shfl(v[1], src); o[2] = shfl(v[2], src); }
__device__ __forceinline__ double sel3(const double* v, int k) { return k == 0 ? v[0] : (k == 1 ? v[1] : v[2]); }
__device__ __forceinline__ void add_at(double* v, int k, double x) { v[0] += (k == 0) ? x : 0.0; v[1] += (k == 1) ? x : 0.0; v[2] += (k == 2) ? x : 0.0; }
__device__ __forceinline__ void unit(int k, double* e) { e[0] = (k == 0); e[1] = (k == 1); e[2] = (k == 2); }

enum { R_BALL = 0, R_HANG = 1, R_FPOS = 2, R_FANG = 3, R_CONT = 4 };

__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK)
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
  uint32_t seed = 0;
  int fell = 0, status = 0, step = 0;
  double t_fall = 0;
  for (; step < A.n_steps; step++) {
    const double z0 = shfl(pos[2], 0);
    if (play_t >= A.tmin && z0 < A.hc) { fell = 1; t_fall = play_t; break; }   // fall_check (warp-uniform)
    double fel[3] = {0, 0, -sb.mass * A.gravity}, fea[3] = {0, 0, 0};          // external force / torque on this body
    double fcl[3] = {0, 0, 0}, fca[3] = {0, 0, 0};                              // constraint-force accumulator (M^-1 J^T lambda)
    // ---- joint rows of this lane (all joints at once): geometry, control torque, right-hand sides
    double ppos[3], pq[4], pav[3];
    shfl3(pos, par, ppos);
    pq[0] = shfl(q[0], par); pq[1] = shfl(q[1], par); pq[2] = shfl(q[2], par); pq[3] = shfl(q[3], par);
    shfl3(av, par, pav);
    double ga[3] = {0, 0, 0}, gb[3] = {0, 0, 0}, gp[3] = {0, 0, 0}, gq[3] = {0, 0, 0};  // hinge: a1, a2, p, q ; fixed: ga = ofs
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
    double cc[3] = {0, 0, 0}, cnorm = 0;
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
        W.kind[rbase + k] = R_BALL; W.owner[rbase + k] = (unsigned char)lane; W.comp[rbase + k] = (unsigned char)k;
      }
#pragma unroll
      for (int r = 0; r < 2; r++) {
        const double* u = r ? gq : gp;
        const double acc = dot3(u, va) - dot3(u, pva);
        const double Ad = sor_w / ((ii + pii) + wcfm);
        W.Ad[rbase + 3 + r] = Ad; W.rhs[rbase + 3 + r] = (crow[3 + r] * fps - acc) * Ad; W.lam[rbase + 3 + r] = 0;
        W.kind[rbase + 3 + r] = R_HANG; W.owner[rbase + 3 + r] = (unsigned char)lane; W.comp[rbase + 3 + r] = (unsigned char)r;
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
        W.kind[rbase + k] = R_FPOS; W.owner[rbase + k] = (unsigned char)lane; W.comp[rbase + k] = (unsigned char)k;
      }
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const double acc = pva[k] - va[k];
        const double Ad = sor_w / ((pii + ii) + wcfm);
        W.Ad[rbase + 3 + k] = Ad; W.rhs[rbase + 3 + k] = (crow[3 + k] * fps - acc) * Ad; W.lam[rbase + 3 + k] = 0;
        W.kind[rbase + 3 + k] = R_FANG; W.owner[rbase + 3 + k] = (unsigned char)lane; W.comp[rbase + 3 + k] = (unsigned char)k;
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
        W.kind[cbase + k] = R_CONT; W.owner[cbase + k] = (unsigned char)lane; W.comp[cbase + k] = (unsigned char)k;
      }
    }
    for (int i = lane; i < m; i += 32) W.order[i] = (short)i;
    __syncwarp();
    // ---- SOR projected Gauss-Seidel over the rows in ODE's order.  Inside the sweep x / mass is written x * (1 / mass):
    // the same bits for the unit masses and inertias every reference body has (dynrec.cpp:62-68 never sets a mass), one
    // rounding apart otherwise, and no division routine on the sequential path.
    for (int it = 0; it < A.iterations; it++) {
      if ((it & 7) == 0) {
        if (lane == 0) {
          for (int i = 1; i < m; i++) {
            seed = 1664525u * seed + 1013904223u;   // ODE dRand
            const int s = (int)((double)seed * ((double)(i + 1) / 4294967296.0));
            const short tmp = W.order[i]; W.order[i] = W.order[s]; W.order[s] = tmp;
          }
        }
        seed = __shfl_sync(FULL, seed, 0);
        __syncwarp();
      }
      for (int oi = 0; oi < m; oi++) {
        const int i = W.order[oi];
        const int kind = W.kind[i], own = W.owner[i], k = W.comp[i];
        const bool mine = (lane == own);
        double delta = 0;
        if (kind == R_CONT) {   // one body: no exchange
          if (mine) {
            double t[3];
            cross3(cc, cdir[k], t);
            const double jf = dot3(cdir[k], fcl) + dot3(t, fca);
            const double Ad = W.Ad[i], lam = W.lam[i];
            delta = W.rhs[i] - lam * (Ad * ((k == 0) ? scfm : wcfm)) - Ad * jf;
            double nl = lam + delta;
            if (k == 0 && nl < 0) { delta = -lam; nl = 0; }
            W.lam[i] = nl;
#pragma unroll
            for (int c = 0; c < 3; c++) { fcl[c] += cdir[k][c] * delta * im; fca[c] += t[c] * delta * ii; }
          }
          continue;
        }
        const int po = __shfl_sync(FULL, par, own);          // the owner's parent lane
        double pfl[3], pfa[3];
        shfl3(fcl, po, pfl);
        shfl3(fca, po, pfa);
        double gv[3] = {0, 0, 0};                             // what the parent needs for its share of the update
        if (mine) {
          double jf;
          if (kind == R_BALL) {
            double t1[3], t2[3];
            cross3(fca, ga, t1);
            cross3(pfa, gb, t2);
            jf = (sel3(fcl, k) + sel3(t1, k)) - sel3(pfl, k) - sel3(t2, k);
            gv[0] = gb[0]; gv[1] = gb[1]; gv[2] = gb[2];
          } else if (kind == R_HANG) {
            const double* u = k ? gq : gp;
            jf = dot3(u, fca) - dot3(u, pfa);
            gv[0] = u[0]; gv[1] = u[1]; gv[2] = u[2];
          } else if (kind == R_FPOS) {
            double t1[3];
            cross3(ga, pfa, t1);
            jf = sel3(pfl, k) + sel3(t1, k) - sel3(fcl, k);
            gv[0] = ga[0]; gv[1] = ga[1]; gv[2] = ga[2];
          } else {
            jf = sel3(pfa, k) - sel3(fca, k);
          }
          const double Ad = W.Ad[i], lam = W.lam[i];
          delta = W.rhs[i] - lam * (Ad * wcfm) - Ad * jf;
          W.lam[i] = lam + delta;                             // joint rows are unbounded
          // the owner's share of fc += M^-1 J^T delta
          double e[3], t[3];
          unit(k, e);
          if (kind == R_BALL) {
            add_at(fcl, k, im * delta);
            cross3(ga, e, t);
#pragma unroll
            for (int c = 0; c < 3; c++) fca[c] += ii * t[c] * delta;
          } else if (kind == R_HANG) {
#pragma unroll
            for (int c = 0; c < 3; c++) fca[c] += ii * gv[c] * delta;
          } else if (kind == R_FPOS) {
            add_at(fcl, k, -(delta * im));
          } else {
            add_at(fca, k, -(delta * ii));
          }
        }
        delta = shfl(delta, own);
        double g3[3];
        shfl3(gv, own, g3);
        if (lane == po) {   // the parent's share
          double e[3], t[3];
          unit(k, e);
          if (kind == R_BALL) {
            add_at(fcl, k, -(im * delta));
            cross3(e, g3, t);
#pragma unroll
            for (int c = 0; c < 3; c++) fca[c] += ii * t[c] * delta;
          } else if (kind == R_HANG) {
#pragma unroll
            for (int c = 0; c < 3; c++) fca[c] -= ii * g3[c] * delta;
          } else if (kind == R_FPOS) {
            add_at(fcl, k, delta * im);
            cross3(e, g3, t);
#pragma unroll
            for (int c = 0; c < 3; c++) fca[c] += t[c] * delta * ii;
          } else {
            add_at(fca, k, delta * ii);
          }
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
