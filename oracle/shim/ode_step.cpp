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

void dSpaceCollide(dSpaceID, void*, dNearCallback*) {
  std::fprintf(stderr, "oracle/shim: dSpaceCollide is not implemented yet\n");
  std::abort();
}
int dCollide(dGeomID, dGeomID, int, dContactGeom*, int) { return 0; }
int dWorldQuickStep(dWorldID, dReal) {
  std::fprintf(stderr, "oracle/shim: dWorldQuickStep is not implemented yet\n");
  std::abort();
  return 0;
}
