// TEST INFRASTRUCTURE ONLY -- shim for <ode/ode.h> (Open Dynamics Engine), which is a third-party dependency of
// the reference that is absent from /root/reference and from this image (reference: matrix.h:10, visualization.h:13,
// geom.h:11; linked by makefile:7 as -lode, version unpinned).
//
// Purpose: let the reference's OWN sources compile unmodified into oracle/_ref (oracle/Makefile, target `ref`).
// Only the declarations the reference mentions exist here.  What the gait-evaluation path needs from ODE is
// functional and restated from ODE's public source (rotation.cpp, mass.cpp defaults):
//   dRFromAxisAndAngle, dRFromEulerAngles          (visualization.cpp:21,67  model.cpp:45)
//   dBodyCreate -> mass 1, inertia identity; dBodyGetMass   (visualization.cpp:458,485  dynrec.cpp:62-68)
//   body / geom pose setters and getters            (model.cpp:292-305, visualization.cpp:571-583)
// The reference pins the first two itself: rot_ztov() exits unless A*z == v (visualization.cpp:24) and
// pergensetup::turn_torso relies on the Euler round trip (pergen.cpp:377-397).
// World stepping and collision (dWorldQuickStep, dSpaceCollide, dCollide) are declared here and defined in
// oracle/shim/ode_world.cpp (the CPU side of the fall sweep); the evaluation path never calls them.
#ifndef ORACLE_SHIM_ODE_H
#define ORACLE_SHIM_ODE_H

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

typedef double dReal;  // the reference's makefile leaves PRECISION unset and the tree assumes double (matrix.cpp:107-113)
typedef dReal dVector3[4];
typedef dReal dVector4[4];
typedef dReal dMatrix3[12];  // row-major 3x4
typedef dReal dQuaternion[4];
typedef int dTriIndex;
#define dInfinity (std::numeric_limits<double>::infinity())

enum { dSphereClass = 0, dBoxClass, dCapsuleClass, dCylinderClass, dPlaneClass, dRayClass, dConvexClass,
       dGeomTransformClass, dTriMeshClass, dHeightfieldClass };
enum { dContactMu2 = 0x001, dContactFDir1 = 0x002, dContactBounce = 0x004, dContactSoftERP = 0x008,
       dContactSoftCFM = 0x010 };
enum { dShimJointHinge = 1, dShimJointFixed = 2, dShimJointContact = 3 };

struct dMass {
  dReal mass;
  dVector3 c;
  dMatrix3 I;
};

struct dxWorld;
struct dxSpace;
struct dxGeom;
struct dxJoint;
struct dxJointGroup;
struct dxTriMeshData;

struct dxBody {
  dxWorld* world;
  dVector3 pos;
  dMatrix3 R;
  dQuaternion q;
  dVector3 lvel, avel, facc, tacc;
  dMass mass;
  int tag;  // index in the world's body array during a step
};

struct dxGeom {
  int cls;
  dxBody* body;
  dReal radius, length;   // sphere / capsule / cylinder
  dReal plane[4];         // plane a x + b y + c z = d
  dVector3 pos;           // placeable geoms without a body
  dMatrix3 R;
  dxTriMeshData* tmdata;
  dxSpace* space;
};

struct dSurfaceParameters {
  int mode;
  dReal mu, mu2, bounce, bounce_vel, soft_erp, soft_cfm, motion1, motion2, slip1, slip2;
};
struct dContactGeom {
  dVector3 pos, normal;
  dReal depth;
  dxGeom *g1, *g2;
  int side1, side2;
};
struct dContact {
  dSurfaceParameters surface;
  dContactGeom geom;
  dVector3 fdir1;
};

struct dxJoint {
  int type;
  dxWorld* world;
  dxBody *b1, *b2;
  dVector3 anchor1, anchor2;  // hinge / fixed: anchor in each body's frame
  dVector3 axis1, axis2;      // hinge axis in each body's frame
  dQuaternion qrel;           // initial relative orientation (hinge angle zero / fixed)
  dVector3 offset;            // fixed: b2 origin in b1's frame
  dReal motor_torque;         // dJointAddHingeTorque accumulates here until the next step
  dContact contact;
};

struct dxJointGroup { std::vector<dxJoint*> joints; };
struct dxSpace { std::vector<dxGeom*> geoms; };
struct dxWorld {
  dVector3 gravity;
  dReal erp, cfm;
  int qs_iterations;
  std::vector<dxBody*> bodies;
  std::vector<dxJoint*> joints;         // persistent joints (hinge, fixed)
  std::vector<dxJoint*> step_contacts;  // contact joints created since the last dJointGroupEmpty (they live in their group)
};
struct dxTriMeshData {
  const dReal* vertices;
  int n_vert;
  const dTriIndex* indices;
  int n_ind;
};

typedef dxWorld* dWorldID;
typedef dxSpace* dSpaceID;
typedef dxBody* dBodyID;
typedef dxGeom* dGeomID;
typedef dxJoint* dJointID;
typedef dxJointGroup* dJointGroupID;
typedef dxTriMeshData* dTriMeshDataID;
typedef void dNearCallback(void* data, dGeomID o1, dGeomID o2);

// ---------------------------------------------------------------- rotations (ODE rotation.cpp)
inline void dRSetIdentity(dMatrix3 R) {
  for (int i = 0; i < 12; i++) R[i] = 0;
  R[0] = R[5] = R[10] = 1;
}
inline void dQtoR(const dQuaternion q, dMatrix3 R) {
  dReal qq1 = 2 * q[1] * q[1], qq2 = 2 * q[2] * q[2], qq3 = 2 * q[3] * q[3];
  R[0] = 1 - qq2 - qq3;
  R[1] = 2 * (q[1] * q[2] - q[0] * q[3]);
  R[2] = 2 * (q[1] * q[3] + q[0] * q[2]);
  R[3] = 0;
  R[4] = 2 * (q[1] * q[2] + q[0] * q[3]);
  R[5] = 1 - qq1 - qq3;
  R[6] = 2 * (q[2] * q[3] - q[0] * q[1]);
  R[7] = 0;
  R[8] = 2 * (q[1] * q[3] - q[0] * q[2]);
  R[9] = 2 * (q[2] * q[3] + q[0] * q[1]);
  R[10] = 1 - qq1 - qq2;
  R[11] = 0;
}
inline void dQFromAxisAndAngle(dQuaternion q, dReal ax, dReal ay, dReal az, dReal angle) {
  dReal l = ax * ax + ay * ay + az * az;
  if (l > 0) {
    angle *= 0.5;
    q[0] = std::cos(angle);
    l = std::sin(angle) * (1.0 / std::sqrt(l));
    q[1] = ax * l;
    q[2] = ay * l;
    q[3] = az * l;
  } else {
    q[0] = 1;
    q[1] = q[2] = q[3] = 0;
  }
}
inline void dRFromAxisAndAngle(dMatrix3 R, dReal ax, dReal ay, dReal az, dReal angle) {
  dQuaternion q;
  dQFromAxisAndAngle(q, ax, ay, az, angle);
  dQtoR(q, R);
}
inline void dRFromEulerAngles(dMatrix3 R, dReal phi, dReal theta, dReal psi) {
  dReal sphi = std::sin(phi), cphi = std::cos(phi), stheta = std::sin(theta), ctheta = std::cos(theta),
        spsi = std::sin(psi), cpsi = std::cos(psi);
  R[0] = cpsi * ctheta;
  R[1] = spsi * ctheta;
  R[2] = -stheta;
  R[3] = 0;
  R[4] = cpsi * stheta * sphi - spsi * cphi;
  R[5] = spsi * stheta * sphi + cpsi * cphi;
  R[6] = ctheta * sphi;
  R[7] = 0;
  R[8] = cpsi * stheta * cphi + spsi * sphi;
  R[9] = spsi * stheta * cphi - cpsi * sphi;
  R[10] = ctheta * cphi;
  R[11] = 0;
}
void dRtoQ(const dMatrix3 R, dQuaternion q);  // ode_world.cpp

// ---------------------------------------------------------------- world / space / bodies / geoms
inline void dInitODE() {}
inline void dCloseODE() {}
inline dWorldID dWorldCreate() {
  dxWorld* w = new dxWorld;
  w->gravity[0] = w->gravity[1] = w->gravity[2] = w->gravity[3] = 0;
  w->erp = 0.2;     // ODE defaults
  w->cfm = 1e-10;   // double-precision default
  w->qs_iterations = 20;
  return w;
}
void dWorldDestroy(dWorldID w);
inline void dWorldSetGravity(dWorldID w, dReal x, dReal y, dReal z) { w->gravity[0] = x; w->gravity[1] = y; w->gravity[2] = z; }
inline void dWorldSetERP(dWorldID w, dReal erp) { w->erp = erp; }
inline void dWorldSetCFM(dWorldID w, dReal cfm) { w->cfm = cfm; }
inline void dWorldSetQuickStepNumIterations(dWorldID w, int n) { w->qs_iterations = n; }
int dWorldQuickStep(dWorldID w, dReal stepsize);
void dShimSeedRandom(unsigned long s);  // ODE's dRandSetSeed: seed of the constraint re-ordering (0 in a fresh process)

inline dSpaceID dHashSpaceCreate(dSpaceID) { return new dxSpace; }
void dSpaceDestroy(dSpaceID s);
void dSpaceCollide(dSpaceID s, void* data, dNearCallback* cb);
int dCollide(dGeomID o1, dGeomID o2, int flags, dContactGeom* contact, int skip);

inline dJointGroupID dJointGroupCreate(int) { return new dxJointGroup; }
void dJointGroupEmpty(dJointGroupID g);
inline void dJointGroupDestroy(dJointGroupID g) { dJointGroupEmpty(g); delete g; }

inline dBodyID dBodyCreate(dWorldID w) {
  dxBody* b = new dxBody;
  std::memset(b, 0, sizeof(dxBody));
  b->world = w;
  dRSetIdentity(b->R);
  b->q[0] = 1;
  b->mass.mass = 1;          // ODE: dMassSetParameters(&b->mass,1,0,0,0,1,1,1,0,0,0)
  dRSetIdentity(b->mass.I);
  if (w) w->bodies.push_back(b);
  return b;
}
inline void dBodyGetMass(dBodyID b, dMass* m) { *m = b->mass; }
inline void dBodySetPosition(dBodyID b, dReal x, dReal y, dReal z) { b->pos[0] = x; b->pos[1] = y; b->pos[2] = z; }
inline void dBodySetRotation(dBodyID b, const dMatrix3 R) {
  for (int i = 0; i < 12; i++) b->R[i] = R[i];
  dRtoQ(b->R, b->q);
}
inline void dBodySetQuaternion(dBodyID b, const dQuaternion q) {
  for (int i = 0; i < 4; i++) b->q[i] = q[i];
  dQtoR(b->q, b->R);
}
inline void dBodySetLinearVel(dBodyID b, dReal x, dReal y, dReal z) { b->lvel[0] = x; b->lvel[1] = y; b->lvel[2] = z; }
inline void dBodySetAngularVel(dBodyID b, dReal x, dReal y, dReal z) { b->avel[0] = x; b->avel[1] = y; b->avel[2] = z; }
inline const dReal* dBodyGetPosition(dBodyID b) { return b->pos; }
inline const dReal* dBodyGetRotation(dBodyID b) { return b->R; }
inline const dReal* dBodyGetQuaternion(dBodyID b) { return b->q; }
inline const dReal* dBodyGetLinearVel(dBodyID b) { return b->lvel; }
inline const dReal* dBodyGetAngularVel(dBodyID b) { return b->avel; }
inline void dBodyAddForce(dBodyID b, dReal fx, dReal fy, dReal fz) { b->facc[0] += fx; b->facc[1] += fy; b->facc[2] += fz; }
inline void dBodyAddTorque(dBodyID b, dReal fx, dReal fy, dReal fz) { b->tacc[0] += fx; b->tacc[1] += fy; b->tacc[2] += fz; }

inline dxGeom* dShimNewGeom(dSpaceID s, int cls) {
  dxGeom* g = new dxGeom;
  std::memset(g, 0, sizeof(dxGeom));
  g->cls = cls;
  dRSetIdentity(g->R);
  g->space = s;
  if (s) s->geoms.push_back(g);
  return g;
}
inline dGeomID dCreateSphere(dSpaceID s, dReal r) { dxGeom* g = dShimNewGeom(s, dSphereClass); g->radius = r; return g; }
inline dGeomID dCreateCapsule(dSpaceID s, dReal r, dReal l) { dxGeom* g = dShimNewGeom(s, dCapsuleClass); g->radius = r; g->length = l; return g; }
inline dGeomID dCreateCylinder(dSpaceID s, dReal r, dReal l) { dxGeom* g = dShimNewGeom(s, dCylinderClass); g->radius = r; g->length = l; return g; }
inline dGeomID dCreatePlane(dSpaceID s, dReal a, dReal b, dReal c, dReal d) {
  dxGeom* g = dShimNewGeom(s, dPlaneClass);
  g->plane[0] = a; g->plane[1] = b; g->plane[2] = c; g->plane[3] = d;
  return g;
}
inline void dGeomSetBody(dGeomID g, dBodyID b) { g->body = b; }
inline dBodyID dGeomGetBody(dGeomID g) { return g->body; }
inline int dGeomGetClass(dGeomID g) { return g->cls; }
inline const dReal* dGeomGetPosition(dGeomID g) { return g->body ? g->body->pos : g->pos; }
inline const dReal* dGeomGetRotation(dGeomID g) { return g->body ? g->body->R : g->R; }
inline dReal dGeomSphereGetRadius(dGeomID g) { return g->radius; }
inline void dGeomCapsuleGetParams(dGeomID g, dReal* r, dReal* l) { *r = g->radius; *l = g->length; }
inline void dGeomCylinderGetParams(dGeomID g, dReal* r, dReal* l) { *r = g->radius; *l = g->length; }
inline void dGeomBoxGetLengths(dGeomID, dVector3 l) { l[0] = l[1] = l[2] = 0; }
void dGeomDestroy(dGeomID g);
inline void dGeomDisable(dGeomID) {}
inline void dGeomEnable(dGeomID) {}

// trimesh terrain (geom.cpp; uneven-ground tests only -- data is kept, collision against it is not implemented)
inline dTriMeshDataID dGeomTriMeshDataCreate() { dxTriMeshData* d = new dxTriMeshData; std::memset(d, 0, sizeof(*d)); return d; }
inline void dGeomTriMeshDataDestroy(dTriMeshDataID d) { delete d; }
inline void dGeomTriMeshDataBuildSimple(dTriMeshDataID d, const dReal* v, int nv, const dTriIndex* idx, int ni) {
  d->vertices = v; d->n_vert = nv; d->indices = idx; d->n_ind = ni;
}
inline dGeomID dCreateTriMesh(dSpaceID s, dTriMeshDataID d, void*, void*, void*) {
  dxGeom* g = dShimNewGeom(s, dTriMeshClass);
  g->tmdata = d;
  return g;
}
inline dTriMeshDataID dGeomTriMeshGetData(dGeomID g) { return g->tmdata; }
inline void dGeomTriMeshGetTriangle(dGeomID g, int i, dVector3* v0, dVector3* v1, dVector3* v2) {
  dVector3* vv[3] = {v0, v1, v2};
  for (int k = 0; k < 3; k++) {
    const dReal* p = g->tmdata->vertices + 4 * g->tmdata->indices[3 * i + k];
    for (int c = 0; c < 3; c++) (*vv[k])[c] = p[c];
  }
}

// ---------------------------------------------------------------- joints
dJointID dJointCreateHinge(dWorldID w, dJointGroupID g);
dJointID dJointCreateFixed(dWorldID w, dJointGroupID g);
dJointID dJointCreateContact(dWorldID w, dJointGroupID g, const dContact* c);
void dJointAttach(dJointID j, dBodyID b1, dBodyID b2);
void dJointSetHingeAnchor(dJointID j, dReal x, dReal y, dReal z);
void dJointSetHingeAxis(dJointID j, dReal x, dReal y, dReal z);
void dJointSetFixed(dJointID j);
dReal dJointGetHingeAngle(dJointID j);
dReal dJointGetHingeAngleRate(dJointID j);
void dJointAddHingeTorque(dJointID j, dReal torque);

#endif
