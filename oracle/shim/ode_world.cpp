// TEST INFRASTRUCTURE ONLY -- out-of-line part of the ODE / drawstuff shims (see ode/ode.h, drawstuff/drawstuff.h).
// Object bookkeeping for worlds, spaces, joints; the rigid-body stepper behind dWorldQuickStep lives in
// ode_step.cpp.  None of this is on the gait-evaluation path; it lets visualization.cpp, model.cpp
// and player.cpp link unmodified.
#include <cstdio>
#include "drawstuff/drawstuff.h"
#include "ode/ode.h"

void dRtoQ(const dMatrix3 R, dQuaternion q) {  // ODE rotation.cpp
  dReal tr = R[0] + R[5] + R[10], s;
  if (tr >= 0) {
    s = std::sqrt(tr + 1);
    q[0] = 0.5 * s;
    s = 0.5 / s;
    q[1] = (R[9] - R[6]) * s;
    q[2] = (R[2] - R[8]) * s;
    q[3] = (R[4] - R[1]) * s;
  } else if (R[5] > R[0] && R[5] >= R[10]) {  // _R(1,1) largest
    s = std::sqrt((R[5] - (R[10] + R[0])) + 1);
    q[2] = 0.5 * s;
    s = 0.5 / s;
    q[3] = (R[9] + R[6]) * s;
    q[1] = (R[4] + R[1]) * s;
    q[0] = (R[2] - R[8]) * s;
  } else if (R[10] > R[0] && R[10] > R[5]) {  // _R(2,2) largest
    s = std::sqrt((R[10] - (R[0] + R[5])) + 1);
    q[3] = 0.5 * s;
    s = 0.5 / s;
    q[1] = (R[2] + R[8]) * s;
    q[2] = (R[9] + R[6]) * s;
    q[0] = (R[4] - R[1]) * s;
  } else {  // _R(0,0) largest
    s = std::sqrt((R[0] - (R[5] + R[10])) + 1);
    q[1] = 0.5 * s;
    s = 0.5 / s;
    q[2] = (R[4] + R[1]) * s;
    q[3] = (R[2] + R[8]) * s;
    q[0] = (R[9] - R[6]) * s;
  }
}

void dWorldDestroy(dWorldID w) {
  if (!w) return;
  for (size_t i = 0; i < w->bodies.size(); i++) delete w->bodies[i];
  for (size_t i = 0; i < w->joints.size(); i++) delete w->joints[i];
  delete w;
}
void dSpaceDestroy(dSpaceID s) {
  if (!s) return;
  for (size_t i = 0; i < s->geoms.size(); i++) delete s->geoms[i];
  delete s;
}
void dGeomDestroy(dGeomID g) {
  if (!g) return;
  if (g->space) {
    std::vector<dxGeom*>& v = g->space->geoms;
    for (size_t i = 0; i < v.size(); i++)
      if (v[i] == g) { v.erase(v.begin() + i); break; }
  }
  delete g;
}
void dJointGroupEmpty(dJointGroupID g) {
  for (size_t i = 0; i < g->joints.size(); i++) {
    dxJoint* j = g->joints[i];
    if (j->world) {
      std::vector<dxJoint*>& v = j->world->step_contacts;
      for (size_t k = 0; k < v.size(); k++) if (v[k] == j) { v.erase(v.begin() + k); break; }
    }
    delete j;
  }
  g->joints.clear();
}

static dxJoint* new_joint(dWorldID w, dJointGroupID g, int type) {
  dxJoint* j = new dxJoint;
  std::memset(j, 0, sizeof(dxJoint));
  j->type = type;
  j->world = w;
  j->qrel[0] = 1;
  if (g) g->joints.push_back(j); else if (w) w->joints.push_back(j);
  return j;
}
dJointID dJointCreateHinge(dWorldID w, dJointGroupID g) { return new_joint(w, g, dShimJointHinge); }
dJointID dJointCreateFixed(dWorldID w, dJointGroupID g) { return new_joint(w, g, dShimJointFixed); }
dJointID dJointCreateContact(dWorldID w, dJointGroupID g, const dContact* c) {
  dxJoint* j = new_joint(w, g, dShimJointContact);
  j->contact = *c;
  if (w) w->step_contacts.push_back(j);
  return j;
}
void dJointAttach(dJointID j, dBodyID b1, dBodyID b2) {
  // ODE: if only the second body is given the bodies are swapped (and the joint flagged reversed)
  if (!b1 && b2) { j->b1 = b2; j->b2 = 0; j->contact.geom.normal[0] = -j->contact.geom.normal[0];
    j->contact.geom.normal[1] = -j->contact.geom.normal[1]; j->contact.geom.normal[2] = -j->contact.geom.normal[2]; }
  else { j->b1 = b1; j->b2 = b2; }
}

long ds_shim_max_steps = 0;
long ds_shim_steps_done = 0;
void dsSimulationLoop(int, char**, int, int, dsFunctions* fn) {
  if (fn->start) fn->start();
  for (ds_shim_steps_done = 0; ds_shim_steps_done < ds_shim_max_steps; ds_shim_steps_done++) fn->step(0);
  if (fn->stop) fn->stop();
}
