// Packed model constants and per-candidate constants shared by host and device.
//
// The reference keeps a pointer tree of modelnode / modeljoint / odepart
// objects (model.h:34-137, visualization.h:93-120) and walks it recursively
// for every frame.  Here the tree is flattened once at load time into a POD
// block that is passed to the kernels as a __grid_constant__ parameter, so
// every constant is a warp-uniform constant-bank operand.
//
// Topology the evaluation path supports (all three reference models,
// SURVEY.md Appendix A):  body 0 carries the free joint; "trunk" bodies are
// body 0 plus jointless bodies below it (their frames differ from the torso's
// by a constant translation, model.cpp:81-86,196-198); every limb is a chain
// of exactly three hinge bodies hanging off a trunk body (lik.cpp:304-311,
// 363-365).  Anything else is rejected by hsl_model_load_xml.
#pragma once
#include <stdint.h>

#define HSL_MAX_LIMBS 8
#define HSL_MAX_TRUNK 8
#define HSL_MAX_BODIES 32
#define HSL_NPARAM 13

#define HSL_IK_YXX 0  // lik.cpp:151-184
#define HSL_IK_ZXX 1  // lik.cpp:189-223

// 3x3 matrices are column-major: element (i,j) at m[3*j+i] (as the reference's affine, matrix.cpp:138-146).
typedef struct HslHinge {
  double Rjp[9], tjp[3];  // modeljoint::A_parent  (parent body frame -> joint frame), model.cpp:132-135
  double Rpb[9], tpb[3];  // modelnode::A_pj_body  (joint frame past the hinge -> body frame), model.cpp:138-141
  double com[3];          // translation of odepart::A_body_geom (COM in body frame), visualization.cpp:541-545
  double mass, inertia;   // ODE dBodyCreate defaults: 1 and identity (dynrec.cpp:62-68)
  int32_t body;           // DFS body id
  int32_t aligned;        // +-(1,2,3): hinge axis is +-x,y,z of the parent frame and the joint sits at the body origin; 0: general
} HslHinge;

typedef struct HslLimb {
  HslHinge h[3];
  double foot[3];   // odepart::capsule_to_pos of the last link (foot point), visualization.cpp:503,565
  double oatt[3];   // offset of the trunk body the limb hangs from, in the torso frame
  double ls[3];     // link lengths used by the closed-form IK (lik.cpp:226-227)
  double inv2l1, inv2l2;  // 1/(2 l1), 1/(2 l2)
  int32_t ysign;    // lik.cpp:231,237,243
  int32_t kind;     // HSL_IK_YXX / HSL_IK_ZXX
  int32_t bend;     // liklimb::limb_bend (true), lik.cpp:300
  int32_t pg_index; // pergen index of this LIK limb (pergen.cpp:243-262)
  int32_t attach;   // index into trunk[] of the body the limb hangs from
  int32_t pad;
} HslLimb;

typedef struct HslTrunkBody {
  double off[3];   // body-frame origin in the torso frame (sum of body@pos along the jointless chain)
  double com[3];
  double mass, inertia;
  int32_t body, parent_trunk;  // DFS body id; index into trunk[] of the parent (-1 for the torso)
} HslTrunkBody;

typedef struct HslModelPod {
  int32_t n, nf, nmj, ntrunk, config_dim, lik_index;
  double rcap, g;
  double Pt[3];  // torso joint frame origin = modeljoint::A_parent translation (model.cpp:158-161); jpos of body 0
  double Qt[3];  // torso A_pj_body translation (model.cpp:164-168)
  HslTrunkBody trunk[HSL_MAX_TRUNK];
  HslLimb limb[HSL_MAX_LIMBS];
  int32_t parent[HSL_MAX_BODIES];
  int32_t jkind[HSL_MAX_BODIES];   // 0 none, 1 free6, 2 hinge
  int32_t motor_of_body[HSL_MAX_BODIES];  // hinge index in joint-value order, -1 otherwise
} HslModelPod;

// Constants of the fall / perturbation sweep (hsl_fall.cu): what the reference hands to ODE when it
// builds the simulated robot -- one rigid body per model body at its COM (odepart::make, visualization.cpp:442-491),
// its first geom as collision shape (sphere and capsule collide with the ground plane, visualization.cpp:296-306), a
// hinge per hinge joint and a fixed joint per jointless body (kinematicmodel::set_ode_joints, model.cpp:375-400;
// odepart::make_hinge_joint / make_fixed_joint, visualization.cpp:506-533), all set up in the zero configuration.
// Body frames here: origin at the COM, axes of the model body frame.
typedef struct HslSimBody {
  int32_t geom;          // 0 does not collide (cylinder), 1 sphere, 2 capsule
  int32_t pad;
  double radius;
  double p0[3], p1[3];   // capsule end points / sphere centre (p0 = p1) relative to the COM, body axes
  double mass, inertia;
} HslSimBody;
typedef struct HslSimJoint {
  int32_t kind;          // 1 hinge, 2 fixed
  int32_t b1, b2;        // ODE's body order: hinge (child, parent), fixed (parent, child)
  int32_t motor;         // hinge: index in joint-value order
  double anchor1[3], anchor2[3], axis1[3], axis2[3];  // in the bodies' own frames
  double qrel[4];        // q1^-1 q2 in the zero configuration
  double offset[3];      // fixed: R1^T (p1 - p2)
} HslSimJoint;
typedef struct HslSimPod {
  int32_t n, nj, nmotor, pad;
  HslSimBody body[HSL_MAX_BODIES];
  HslSimJoint joint[HSL_MAX_BODIES];
  double com[HSL_MAX_BODIES][3];  // COM offset in the model body frame (initial poses: A_ground * com)
} HslSimPod;

// Per-candidate constants written by the setup kernel (reference: pgssweeper::setup_pergen and
// friends, pergen.cpp:453-507; periodicgenerator::set_step_duration, pergen.cpp:30-51).
typedef struct HslCand {
  double R0[9];      // torso rotation from torso_angles (constant over time when curvature == 0)
  double tp0[3];     // torso_pos
  double eul[3];     // torso_angles
  double period, step_length, step_height, v, t_step, curvature, max_radius, dt, hh;  // hh = 1/(2 dt)
  double pos0[HSL_MAX_LIMBS][3];  // default foot positions, LIK order
  double ts[HSL_MAX_LIMBS], xs[HSL_MAX_LIMBS];  // lift-off tables, LIK order
  double turn_r[HSL_MAX_LIMBS], turn_sa[HSL_MAX_LIMBS], turn_ca[HSL_MAX_LIMBS];  // curved gaits: radius and sin/cos of the
                                                                               // polar angle of pos0 about the turning centre
  int32_t status, pad;
} HslCand;

// per-candidate status bits
#define HSL_ST_BAD_PARAMS 1      // reference: "ERROR: f = ... out of bounds", exit(1) (pergen.cpp:31)
#define HSL_ST_UNREACHABLE 2     // reference: "LIK ERROR: limb position is unreachable", exit(1) (lik.cpp:161-164,321-330)
#define HSL_ST_SOLVER 4          // contact blocks not positive definite (reference: threshold retry loop, ftsolver.cpp:208-232)
#define HSL_ST_FEW_CONTACTS 8    // fewer than 2 feet on the ground in some frame (cannot occur for valid step_duration)
#define HSL_ST_ILLCOND 16        // informational: level 0 nearly rank deficient (contact points almost collinear) in some frame
#define HSL_ILLCOND_PIVOT 1e-4   // ... = an LDL^T pivot of the 6x6 level-0 matrix below this fraction of its trace
