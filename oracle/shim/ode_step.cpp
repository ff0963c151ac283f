// TEST INFRASTRUCTURE ONLY -- joint kinematics and the world stepper of the ODE shim (see ode/ode.h).
// Restated from ODE's public source (joints/hinge.cpp, joints/fixed.cpp, joints/contact.cpp, quickstep.cpp,
// collision_std.cpp), ODE being absent from /root/reference and from this image.  Not on the gait-evaluation
// path; used by the reference's simulation modes (player.cpp:326-340) for the fall sweep.
#include <cstdio>
#include "ode/ode.h"

namespace {
inline void q_mul(dQuaternion qa, const dQuaternion qb, const dQuaternion qc) {  // qa = qb qc
  qa[0] = qb[0] * qc[0] - qb[1] * qc[1] - qb[2] * qc[2] - qb[3] * qc[3];
  qa[1] = qb[0] * qc[1] + qb[1] * qc[0] + qb[2] * qc[3] - qb[3] * qc[2];
  qa[2] = qb[0] * qc[2] + qb[2] * qc[0] + qb[3] * qc[1] - qb[1] * qc[3];
  qa[3] = qb[0] * qc[3] + qb[3] * qc[0] + qb[1] * qc[2] - qb[2] * qc[1];
}
inline void q_conj(dQuaternion qa, const dQuaternion qb) { qa[0] = qb[0]; qa[1] = -qb[1]; qa[2] = -qb[2]; qa[3] = -qb[3]; }
inline void to_body(const dxBody* b, const dReal* v, dReal* out, bool point) {  // out = R^T (v - pos)
  dReal d[3] = {v[0], v[1], v[2]};
  if (point) for (int i = 0; i < 3; i++) d[i] -= b->pos[i];
  for (int i = 0; i < 3; i++) out[i] = b->R[i] * d[0] + b->R[4 + i] * d[1] + b->R[8 + i] * d[2];
}
inline void to_world(const dxBody* b, const dReal* v, dReal* out, bool point) {  // out = R v (+ pos)
  for (int i = 0; i < 3; i++) out[i] = b->R[4 * i] * v[0] + b->R[4 * i + 1] * v[1] + b->R[4 * i + 2] * v[2] + (point ? b->pos[i] : 0);
}
}  // namespace

void dJointSetHingeAnchor(dJointID j, dReal x, dReal y, dReal z) {
  dReal a[3] = {x, y, z};
  if (j->b1) to_body(j->b1, a, j->anchor1, true);
  if (j->b2) to_body(j->b2, a, j->anchor2, true); else for (int i = 0; i < 3; i++) j->anchor2[i] = a[i];
}
void dJointSetHingeAxis(dJointID j, dReal x, dReal y, dReal z) {
  dReal l = std::sqrt(x * x + y * y + z * z);
  dReal a[3] = {x / l, y / l, z / l};
  if (j->b1) to_body(j->b1, a, j->axis1, false);
  if (j->b2) to_body(j->b2, a, j->axis2, false); else for (int i = 0; i < 3; i++) j->axis2[i] = a[i];
  // initial relative rotation body1 -> body2: qrel = q1^-1 q2
  dQuaternion c;
  if (j->b1 && j->b2) { q_conj(c, j->b1->q); q_mul(j->qrel, c, j->b2->q); }
  else if (j->b1) { q_conj(j->qrel, j->b1->q); }
}
void dJointSetFixed(dJointID j) {
  if (!j->b1) return;
  if (j->b2) {
    dReal d[3] = {j->b2->pos[0], j->b2->pos[1], j->b2->pos[2]};
    to_body(j->b1, d, j->offset, true);
    dQuaternion c;
    q_conj(c, j->b1->q);
    q_mul(j->qrel, c, j->b2->q);
  } else {
    for (int i = 0; i < 3; i++) j->offset[i] = j->b1->pos[i];
    q_conj(j->qrel, j->b1->q);
  }
}
dReal dJointGetHingeAngle(dJointID j) {
  if (!j->b1) return 0;
  dQuaternion qq, c, qrel;
  q_conj(c, j->b1->q);
  if (j->b2) q_mul(qq, c, j->b2->q); else for (int i = 0; i < 4; i++) qq[i] = c[i];  // q1^-1 q2
  q_conj(c, j->qrel);
  q_mul(qrel, qq, c);  // times the inverse of the initial relative rotation
  dReal cost2 = qrel[0];
  dReal sint2 = std::sqrt(qrel[1] * qrel[1] + qrel[2] * qrel[2] + qrel[3] * qrel[3]);
  dReal dot = qrel[1] * j->axis1[0] + qrel[2] * j->axis1[1] + qrel[3] * j->axis1[2];
  dReal theta = (dot >= 0) ? 2 * std::atan2(sint2, cost2) : 2 * std::atan2(sint2, -cost2);
  if (theta > M_PI) theta -= 2 * M_PI;
  return -theta;
}
dReal dJointGetHingeAngleRate(dJointID j) {
  if (!j->b1) return 0;
  dReal axis[3];
  to_world(j->b1, j->axis1, axis, false);
  dReal rate = axis[0] * j->b1->avel[0] + axis[1] * j->b1->avel[1] + axis[2] * j->b1->avel[2];
  if (j->b2) rate -= axis[0] * j->b2->avel[0] + axis[1] * j->b2->avel[1] + axis[2] * j->b2->avel[2];
  return rate;
}
void dJointAddHingeTorque(dJointID j, dReal torque) {
  dReal axis[3];
  if (!j->b1) return;
  to_world(j->b1, j->axis1, axis, false);
  for (int i = 0; i < 3; i++) {
    j->b1->tacc[i] += axis[i] * torque;
    if (j->b2) j->b2->tacc[i] -= axis[i] * torque;
  }
}

// ---------------------------------------------------------------------------------------------- collision
// Geoms with a body against the plane geoms of the space: the only pairs the reference's nearCallback keeps
// (capsule / sphere vs plane or trimesh, visualization.cpp:296-306; trimesh terrain is not implemented here).
void dSpaceCollide(dSpaceID s, void* data, dNearCallback* cb) {
  for (size_t i = 0; i < s->geoms.size(); i++) {
    dxGeom* g = s->geoms[i];
    if (!g->body) continue;
    for (size_t j = 0; j < s->geoms.size(); j++)
      if (s->geoms[j]->cls == dPlaneClass) cb(data, g, s->geoms[j]);
  }
}
// ODE collision_std.cpp: dCollideSpherePlane, dCollideCapsulePlane (first contact only: the reference asks for one).
int dCollide(dGeomID o1, dGeomID o2, int flags, dContactGeom* c, int) {
  if ((flags & 0xffff) < 1 || o2->cls != dPlaneClass || !o1->body) return 0;
  const dReal* n = o2->plane;
  const dxBody* b = o1->body;
  dReal p[3] = {b->pos[0], b->pos[1], b->pos[2]};
  if (o1->cls == dCapsuleClass) {
    // the deeper of the two capping spheres: the capsule's axis is the body's z axis
    const dReal az[3] = {b->R[2], b->R[6], b->R[10]};
    const dReal sign = (n[0] * az[0] + n[1] * az[1] + n[2] * az[2] > 0) ? -1.0 : 1.0;
    for (int k = 0; k < 3; k++) p[k] += az[k] * (o1->length * 0.5 * sign);
  } else if (o1->cls != dSphereClass) {
    return 0;
  }
  const dReal depth = n[3] - (n[0] * p[0] + n[1] * p[1] + n[2] * p[2]) + o1->radius;
  if (depth < 0) return 0;
  for (int k = 0; k < 3; k++) { c->normal[k] = n[k]; c->pos[k] = p[k] - n[k] * o1->radius; }
  c->depth = depth;
  c->g1 = o1;
  c->g2 = o2;
  return 1;
}

// ---------------------------------------------------------------------------------------------- quickstep
// Restatement of ODE's dWorldQuickStep (quickstep.cpp, step.cpp dxStepBody, joints/{hinge,fixed,contact}.cpp):
// maximal coordinates, one constraint row per removed degree of freedom, projected Gauss-Seidel with successive
// over-relaxation (w = 1.3, 20 iterations, lambda starts at 0), rows re-shuffled with ODE's linear congruential
// generator every eighth iteration, semi-implicit Euler.  What is NOT ODE's: the order in which bodies and joints
// are visited (ODE walks islands of its intrusive lists; here joints go in creation order, contacts after them), so a
// run agrees with a real ODE build statistically, not step for step.  hslabs_b200/csrc/hsl_fall.cu follows THIS
// arithmetic row for row.
namespace {
struct Row {
  dReal J1l[3], J1a[3], J2l[3], J2a[3];
  dReal c, cfm, lo, hi;
  dxBody *b1, *b2;
};
inline void cross(const dReal* a, const dReal* b, dReal* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
inline dReal dot3(const dReal* a, const dReal* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
void plane_space(const dReal* n, dReal* p, dReal* q) {  // ODE dPlaneSpace
  if (std::fabs(n[2]) > M_SQRT1_2) {
    const dReal a = n[1] * n[1] + n[2] * n[2], k = 1.0 / std::sqrt(a);
    p[0] = 0; p[1] = -n[2] * k; p[2] = n[1] * k;
    q[0] = a * k; q[1] = -n[0] * p[2]; q[2] = n[0] * p[1];
  } else {
    const dReal a = n[0] * n[0] + n[1] * n[1], k = 1.0 / std::sqrt(a);
    p[0] = -n[1] * k; p[1] = n[0] * k; p[2] = 0;
    q[0] = -n[2] * p[1]; q[1] = n[2] * p[0]; q[2] = a * k;
  }
}
Row blank_row(dxBody* b1, dxBody* b2, dReal cfm) {
  Row r;
  std::memset(&r, 0, sizeof r);
  r.b1 = b1; r.b2 = b2; r.cfm = cfm;
  r.lo = -dInfinity; r.hi = dInfinity;
  return r;
}
// three rows that keep the point b1.pos + a1 on b2.pos + a2 (setBall)
void ball_rows(std::vector<Row>& rows, dxBody* b1, dxBody* b2, const dReal* a1, const dReal* a2, dReal k, dReal cfm) {
  for (int i = 0; i < 3; i++) {
    Row r = blank_row(b1, b2, cfm);
    const int j = (i + 1) % 3, l = (i + 2) % 3;
    r.J1l[i] = 1;
    r.J1a[j] = a1[l]; r.J1a[l] = -a1[j];     // row i of -[a1]x
    if (b2) {
      r.J2l[i] = -1;
      r.J2a[j] = -a2[l]; r.J2a[l] = a2[j];   // row i of +[a2]x
      r.c = k * (a2[i] + b2->pos[i] - a1[i] - b1->pos[i]);
    } else {
      r.c = k * (a2[i] - a1[i] - b1->pos[i]);
    }
    rows.push_back(r);
  }
}
void hinge_rows(std::vector<Row>& rows, const dxJoint* j, dReal k, dReal cfm) {
  dxBody *b1 = j->b1, *b2 = j->b2;
  dReal a1[3], a2[3], ax1[3], ax2[3], p[3], q[3], b[3];
  to_world(b1, j->anchor1, a1, false);
  if (b2) to_world(b2, j->anchor2, a2, false); else for (int i = 0; i < 3; i++) a2[i] = j->anchor2[i];
  ball_rows(rows, b1, b2, a1, a2, k, cfm);
  to_world(b1, j->axis1, ax1, false);
  if (b2) to_world(b2, j->axis2, ax2, false); else for (int i = 0; i < 3; i++) ax2[i] = j->axis2[i];
  plane_space(ax1, p, q);
  cross(ax1, ax2, b);
  const dReal* pq[2] = {p, q};
  for (int r2 = 0; r2 < 2; r2++) {
    Row r = blank_row(b1, b2, cfm);
    for (int i = 0; i < 3; i++) { r.J1a[i] = pq[r2][i]; if (b2) r.J2a[i] = -pq[r2][i]; }
    r.c = k * dot3(b, pq[r2]);
    rows.push_back(r);
  }
}
void fixed_rows(std::vector<Row>& rows, const dxJoint* j, dReal k, dReal cfm) {
  dxBody *b1 = j->b1, *b2 = j->b2;
  if (!b2) return;  // the reference only fixes body to body
  dReal ofs[3];     // -(b2 origin seen from b1), in world axes: ofs = R1 * R1_0^T (p1 - p2)_0
  dReal negofs[3] = {-j->offset[0], -j->offset[1], -j->offset[2]};
  to_world(b1, negofs, ofs, false);
  for (int i = 0; i < 3; i++) {
    Row r = blank_row(b1, b2, cfm);
    const int jj = (i + 1) % 3, l = (i + 2) % 3;
    r.J1l[i] = 1;
    r.J1a[jj] = -ofs[l]; r.J1a[l] = ofs[jj];   // row i of +[ofs]x
    r.J2l[i] = -1;
    r.c = k * (b2->pos[i] - b1->pos[i] + ofs[i]);
    rows.push_back(r);
  }
  dQuaternion c1, qq, cr, qerr;
  q_conj(c1, b1->q);
  q_mul(qq, c1, b2->q);
  q_conj(cr, j->qrel);
  q_mul(qerr, qq, cr);
  if (qerr[0] < 0) { qerr[1] = -qerr[1]; qerr[2] = -qerr[2]; qerr[3] = -qerr[3]; }
  dReal e[3];
  to_world(b1, qerr + 1, e, false);
  for (int i = 0; i < 3; i++) {
    Row r = blank_row(b1, b2, cfm);
    r.J1a[i] = 1;
    r.J2a[i] = -1;
    r.c = 2 * k * e[i];
    rows.push_back(r);
  }
}
void contact_rows(std::vector<Row>& rows, const dxJoint* j, dReal k, dReal world_cfm) {
  dxBody *b1 = j->b1, *b2 = j->b2;
  const dContact& ct = j->contact;
  const dReal* n = ct.geom.normal;
  dReal c1[3], c2[3] = {0, 0, 0}, t1[3], t2[3];
  for (int i = 0; i < 3; i++) { c1[i] = ct.geom.pos[i] - b1->pos[i]; if (b2) c2[i] = ct.geom.pos[i] - b2->pos[i]; }
  Row r = blank_row(b1, b2, (ct.surface.mode & dContactSoftCFM) ? ct.surface.soft_cfm : world_cfm);
  for (int i = 0; i < 3; i++) r.J1l[i] = n[i];
  cross(c1, n, r.J1a);
  if (b2) { for (int i = 0; i < 3; i++) r.J2l[i] = -n[i]; dReal t[3]; cross(c2, n, t); for (int i = 0; i < 3; i++) r.J2a[i] = -t[i]; }
  dReal depth = ct.geom.depth;
  if (depth < 0) depth = 0;
  r.c = k * depth;
  if (ct.surface.mode & dContactBounce) {
    dReal outgoing = dot3(r.J1l, b1->lvel) + dot3(r.J1a, b1->avel);
    if (b2) outgoing += dot3(r.J2l, b2->lvel) + dot3(r.J2a, b2->avel);
    if (ct.surface.bounce_vel >= 0 && (-outgoing) > ct.surface.bounce_vel) {
      const dReal newc = -ct.surface.bounce * outgoing;
      if (newc > r.c) r.c = newc;
    }
  }
  r.lo = 0;
  r.hi = dInfinity;
  rows.push_back(r);
  if (!(ct.surface.mu > 0)) return;
  plane_space(n, t1, t2);
  const dReal* tt[2] = {t1, t2};
  for (int d = 0; d < 2; d++) {  // mu = dInfinity in the reference: unbounded friction rows
    Row f = blank_row(b1, b2, world_cfm);
    for (int i = 0; i < 3; i++) f.J1l[i] = tt[d][i];
    cross(c1, tt[d], f.J1a);
    if (b2) { for (int i = 0; i < 3; i++) f.J2l[i] = -tt[d][i]; dReal t[3]; cross(c2, tt[d], t); for (int i = 0; i < 3; i++) f.J2a[i] = -t[i]; }
    f.lo = -ct.surface.mu;
    f.hi = ct.surface.mu;
    rows.push_back(f);
  }
}
unsigned long g_ode_seed = 0;  // ODE misc.cpp: dRand / dRandInt
inline unsigned long ode_rand() { g_ode_seed = (1664525UL * g_ode_seed + 1013904223UL) & 0xffffffffUL; return g_ode_seed; }
inline int ode_rand_int(int n) { return (int)((double)ode_rand() * ((double)n / 4294967296.0)); }
// inverse inertia in world axes (I_body = diag-free 3x3 from dMass, row-major 3x4)
void world_inv_inertia(const dxBody* b, dReal* invI) {
  const dReal* I = b->mass.I;
  const dReal a = I[0], bb = I[1], c = I[2], d = I[4], e = I[5], f = I[6], g = I[8], h = I[9], i = I[10];
  const dReal det = a * (e * i - f * h) - bb * (d * i - f * g) + c * (d * h - e * g);
  dReal inv[9] = {(e * i - f * h) / det, (c * h - bb * i) / det, (bb * f - c * e) / det,
                  (f * g - d * i) / det, (a * i - c * g) / det, (c * d - a * f) / det,
                  (d * h - e * g) / det, (bb * g - a * h) / det, (a * e - bb * d) / det};
  dReal R[9] = {b->R[0], b->R[1], b->R[2], b->R[4], b->R[5], b->R[6], b->R[8], b->R[9], b->R[10]}, t[9];
  for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) { t[3 * r + cc] = 0; for (int kk = 0; kk < 3; kk++) t[3 * r + cc] += R[3 * r + kk] * inv[3 * kk + cc]; }
  for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) { invI[3 * r + cc] = 0; for (int kk = 0; kk < 3; kk++) invI[3 * r + cc] += t[3 * r + kk] * R[3 * cc + kk]; }
}
inline void mat3_vec(const dReal* M, const dReal* v, dReal* o) { for (int r = 0; r < 3; r++) o[r] = M[3 * r] * v[0] + M[3 * r + 1] * v[1] + M[3 * r + 2] * v[2]; }
}  // namespace

void dShimSeedRandom(unsigned long s) { g_ode_seed = s; }

int dWorldQuickStep(dWorldID w, dReal h) {
  const dReal fps = 1.0 / h, k = fps * w->erp;
  const size_t nb = w->bodies.size();
  // contact joints live in joint groups; every group joint of this world that is attached joins the step
  std::vector<Row> rows;
  for (size_t i = 0; i < w->joints.size(); i++) {
    const dxJoint* j = w->joints[i];
    if (!j->b1) continue;
    if (j->type == dShimJointHinge) hinge_rows(rows, j, k, w->cfm);
    else if (j->type == dShimJointFixed) fixed_rows(rows, j, k, w->cfm);
  }
  for (size_t i = 0; i < w->step_contacts.size(); i++)
    if (w->step_contacts[i]->b1) contact_rows(rows, w->step_contacts[i], k, w->cfm);
  const int m = (int)rows.size();
  // per body: inverse mass, inverse inertia, external force incl. gravity (the gyroscopic term vanishes for I = 1)
  std::vector<dReal> invI(9 * nb), fe(6 * nb, 0.0), fc(6 * nb, 0.0);
  for (size_t b = 0; b < nb; b++) {
    dxBody* bd = w->bodies[b];
    bd->tag = (int)b;
    world_inv_inertia(bd, &invI[9 * b]);
    for (int i = 0; i < 3; i++) { fe[6 * b + i] = bd->facc[i] + bd->mass.mass * w->gravity[i]; fe[6 * b + 3 + i] = bd->tacc[i]; }
  }
  if (m > 0) {
    // iMJ = M^-1 J^T, rhs = c/h - J (v/h + M^-1 fe), Ad = w / (J iMJ + cfm/h); then J, rhs scaled by Ad (SOR_LCP)
    std::vector<dReal> iMJ(12 * (size_t)m), Jm(12 * (size_t)m), rhs(m), Adcfm(m), lambda(m, 0.0);
    const dReal sor_w = 1.3;
    for (int i = 0; i < m; i++) {
      Row& r = rows[i];
      dReal* im = &iMJ[12 * (size_t)i];
      dReal* jm = &Jm[12 * (size_t)i];
      const int b1 = r.b1->tag, b2 = r.b2 ? r.b2->tag : -1;
      const dReal im1 = 1.0 / r.b1->mass.mass;
      for (int kk = 0; kk < 3; kk++) im[kk] = r.J1l[kk] * im1;
      mat3_vec(&invI[9 * b1], r.J1a, im + 3);
      if (b2 >= 0) {
        const dReal im2 = 1.0 / r.b2->mass.mass;
        for (int kk = 0; kk < 3; kk++) im[6 + kk] = r.J2l[kk] * im2;
        mat3_vec(&invI[9 * b2], r.J2a, im + 9);
      } else {
        for (int kk = 6; kk < 12; kk++) im[kk] = 0;
      }
      dReal sum = dot3(im, r.J1l) + dot3(im + 3, r.J1a) + dot3(im + 6, r.J2l) + dot3(im + 9, r.J2a);
      // tmp = v/h + M^-1 fe per body, dotted with J
      dReal acc = 0;
      for (int side = 0; side < 2; side++) {
        const dxBody* bd = side ? r.b2 : r.b1;
        if (!bd) continue;
        const int bi = bd->tag;
        const dReal* Jl = side ? r.J2l : r.J1l;
        const dReal* Ja = side ? r.J2a : r.J1a;
        dReal tl[3], ta[3], it[3];
        for (int kk = 0; kk < 3; kk++) tl[kk] = bd->lvel[kk] * fps + fe[6 * bi + kk] / bd->mass.mass;
        mat3_vec(&invI[9 * bi], &fe[6 * bi + 3], it);
        for (int kk = 0; kk < 3; kk++) ta[kk] = bd->avel[kk] * fps + it[kk];
        acc += dot3(Jl, tl) + dot3(Ja, ta);
      }
      const dReal cfm = r.cfm * fps;
      const dReal Ad = sor_w / (sum + cfm);
      rhs[i] = (r.c * fps - acc) * Ad;
      Adcfm[i] = Ad * cfm;
      for (int kk = 0; kk < 3; kk++) { jm[kk] = r.J1l[kk] * Ad; jm[3 + kk] = r.J1a[kk] * Ad; jm[6 + kk] = r.J2l[kk] * Ad; jm[9 + kk] = r.J2a[kk] * Ad; }
    }
    std::vector<int> order(m);
    for (int i = 0; i < m; i++) order[i] = i;
    for (int it = 0; it < w->qs_iterations; it++) {
      if ((it & 7) == 0)
        for (int i = 1; i < m; i++) { const int s = ode_rand_int(i + 1); std::swap(order[i], order[s]); }
      for (int oi = 0; oi < m; oi++) {
        const int i = order[oi];
        const Row& r = rows[i];
        const dReal* jm = &Jm[12 * (size_t)i];
        const dReal* im = &iMJ[12 * (size_t)i];
        const int b1 = r.b1->tag, b2 = r.b2 ? r.b2->tag : -1;
        dReal delta = rhs[i] - lambda[i] * Adcfm[i];
        delta -= dot3(jm, &fc[6 * b1]) + dot3(jm + 3, &fc[6 * b1 + 3]);
        if (b2 >= 0) delta -= dot3(jm + 6, &fc[6 * b2]) + dot3(jm + 9, &fc[6 * b2 + 3]);
        dReal nl = lambda[i] + delta;
        if (nl < r.lo) { delta = r.lo - lambda[i]; nl = r.lo; }
        else if (nl > r.hi) { delta = r.hi - lambda[i]; nl = r.hi; }
        lambda[i] = nl;
        for (int kk = 0; kk < 6; kk++) fc[6 * b1 + kk] += im[kk] * delta;
        if (b2 >= 0) for (int kk = 0; kk < 6; kk++) fc[6 * b2 + kk] += im[6 + kk] * delta;
      }
    }
  }
  // v += h (M^-1 fe + fc); x += h v; q += h/2 [0,w] q, renormalised (dxStepBody, infinitesimal rotation)
  for (size_t b = 0; b < nb; b++) {
    dxBody* bd = w->bodies[b];
    dReal it[3];
    mat3_vec(&invI[9 * b], &fe[6 * b + 3], it);
    for (int i = 0; i < 3; i++) {
      bd->lvel[i] += h * (fe[6 * b + i] / bd->mass.mass + fc[6 * b + i]);
      bd->avel[i] += h * (it[i] + fc[6 * b + 3 + i]);
    }
    for (int i = 0; i < 3; i++) bd->pos[i] += h * bd->lvel[i];
    const dReal* wv = bd->avel;
    const dReal* q = bd->q;
    dReal dq[4] = {0.5 * (-wv[0] * q[1] - wv[1] * q[2] - wv[2] * q[3]), 0.5 * (wv[0] * q[0] + wv[1] * q[3] - wv[2] * q[2]),
                   0.5 * (-wv[0] * q[3] + wv[1] * q[0] + wv[2] * q[1]), 0.5 * (wv[0] * q[2] - wv[1] * q[1] + wv[2] * q[0])};
    dReal nq[4], l = 0;
    for (int i = 0; i < 4; i++) { nq[i] = q[i] + h * dq[i]; l += nq[i] * nq[i]; }
    l = 1.0 / std::sqrt(l);
    for (int i = 0; i < 4; i++) bd->q[i] = nq[i] * l;
    dQtoR(bd->q, bd->R);
    for (int i = 0; i < 4; i++) bd->facc[i] = bd->tacc[i] = 0;
  }
  return 1;
}
