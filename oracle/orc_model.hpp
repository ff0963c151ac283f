// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
// Nothing under hslabs_b200/ may include, link or call this file.
//
// CPU restatement of the reference's kinematic layer:
//   rigid transforms / 4-vectors      matrix.cpp:78-192, 250-322
//   ODE rotation helpers [ext]        restated from ODE's public rotation.cpp
//                                     (dRFromAxisAndAngle, dRFromEulerAngles);
//                                     ODE is absent from /root/reference
//   rotation / Euler helpers          visualization.cpp:11-25, 54-101
//   XML -> kinematic tree, FK         model.cpp:37-62, 81-201, 224-289, 314-372, 403-409
//   per-body geometry constants       visualization.cpp:442-504, 541-568
//   closed-form limb IK               lik.cpp:7-99, 142-245, 295-366
#pragma once
#include "orc_real.hpp"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

struct Failure : std::runtime_error {  // the reference prints "ERROR ..." and exit(1)s
  explicit Failure(const std::string& s) : std::runtime_error(s) {}
};

// ---------------------------------------------------------------- 4-vectors
struct V4 {  // reference extvec: (x,y,z,1)
  real v[4];
  V4() { v[0] = v[1] = v[2] = 0; v[3] = 1; }
  V4(real x, real y, real z) { v[0] = x; v[1] = y; v[2] = z; v[3] = 1; }
  void set3(const real* a) { v[0] = a[0]; v[1] = a[1]; v[2] = a[2]; }
  void sub4(const V4& u) { for (int i = 0; i < 4; i++) v[i] -= u.v[i]; }  // matrix.cpp:250-254 (all 4!)
  void add3(const V4& u) { for (int i = 0; i < 3; i++) v[i] += u.v[i]; }
  void scale3(real f) { for (int i = 0; i < 3; i++) v[i] *= f; }
  real norm3() const { real s = 0; for (int i = 0; i < 3; i++) s += v[i] * v[i]; return orc::m_sqrt(s); }
  real dot3(const V4& u) const { real s = 0; for (int i = 0; i < 3; i++) s += v[i] * u.v[i]; return s; }
  V4 cross(const V4& u) const {
    return V4(v[1] * u.v[2] - v[2] * u.v[1], v[2] * u.v[0] - v[0] * u.v[2], v[0] * u.v[1] - v[1] * u.v[0]);
  }
};

// ------------------------------------------------- rigid transform, col-major
struct M4 {  // reference affine: element (i,j) at a[4j+i]
  real a[16];
  M4() { identity(); }
  void identity() { for (int i = 0; i < 16; i++) a[i] = (i % 5 == 0) ? 1.0 : 0.0; }
  static M4 translation(const real* t) { M4 m; m.a[12] = t[0]; m.a[13] = t[1]; m.a[14] = t[2]; return m; }
  real at(int i, int j) const { return a[4 * j + i]; }
  // this <- this * b, sums accumulated k = 0..3 starting from 0 (matrix.cpp:78-97)
  void mul(const M4& b) {
    real r[16];
    for (int j = 0; j < 4; j++)
      for (int i = 0; i < 4; i++) {
        real s = 0;
        for (int k = 0; k < 4; k++) s += a[4 * k + i] * b.a[4 * j + k];
        r[4 * j + i] = s;
      }
    std::memcpy(a, r, sizeof r);
  }
  V4 apply(const V4& x) const {  // full 4x4 times (x,y,z,w), matrix.cpp:149-163
    V4 y;
    for (int i = 0; i < 4; i++) {
      real s = 0;
      for (int j = 0; j < 4; j++) s += a[4 * j + i] * x.v[j];
      y.v[i] = s;
    }
    return y;
  }
  void set_rot_raw12(const real* r12) {  // raw copy of an ODE dMatrix3 (=> transpose), matrix.cpp:107-113
    for (int i = 0; i < 12; i++) a[i] = r12[i];
    a[12] = a[13] = a[14] = 0; a[15] = 1;
  }
  void set_rot_from(const M4& m) { set_rot_raw12(m.a); }
  void shift(const V4& t) { for (int i = 0; i < 3; i++) a[12 + i] += t.v[i]; }  // matrix.cpp:124-128
  void transpose4() { for (int i = 0; i < 4; i++) for (int j = i + 1; j < 4; j++) std::swap(a[4 * j + i], a[4 * i + j]); }
  void invert_rigid() {  // matrix.cpp:182-192
    V4 t;
    for (int i = 0; i < 3; i++) { t.v[i] = -a[12 + i]; a[12 + i] = 0; }
    transpose4();
    V4 t1 = apply(t);
    shift(t1);
  }
  V4 translation_part() const { return V4(a[12], a[13], a[14]); }
};

// ---------------------------------------------------- ODE helpers [ext], dReal=real
// Row-major 3x4 ("dMatrix3") outputs.
inline void ode_R_from_axis_angle(real R[12], real ax, real ay, real az, real angle) {
  real q[4];
  real l = ax * ax + ay * ay + az * az;
  if (l > 0) {
    angle *= 0.5;
    q[0] = orc::m_cos(angle);
    l = orc::m_sin(angle) * (1.0 / orc::m_sqrt(l));
    q[1] = ax * l; q[2] = ay * l; q[3] = az * l;
  } else { q[0] = 1; q[1] = q[2] = q[3] = 0; }
  real qq1 = 2 * q[1] * q[1], qq2 = 2 * q[2] * q[2], qq3 = 2 * q[3] * q[3];
  R[0] = 1 - qq2 - qq3;                 R[1] = 2 * (q[1] * q[2] - q[0] * q[3]); R[2] = 2 * (q[1] * q[3] + q[0] * q[2]);  R[3] = 0;
  R[4] = 2 * (q[1] * q[2] + q[0] * q[3]); R[5] = 1 - qq1 - qq3;                 R[6] = 2 * (q[2] * q[3] - q[0] * q[1]);  R[7] = 0;
  R[8] = 2 * (q[1] * q[3] - q[0] * q[2]); R[9] = 2 * (q[2] * q[3] + q[0] * q[1]); R[10] = 1 - qq1 - qq2;                R[11] = 0;
}
inline void ode_R_from_euler(real R[12], real phi, real theta, real psi) {
  real sphi = orc::m_sin(phi), cphi = orc::m_cos(phi), sth = orc::m_sin(theta), cth = orc::m_cos(theta),
         spsi = orc::m_sin(psi), cpsi = orc::m_cos(psi);
  R[0] = cpsi * cth;                      R[1] = spsi * cth;                      R[2] = -sth;        R[3] = 0;
  R[4] = cpsi * sth * sphi - spsi * cphi; R[5] = spsi * sth * sphi + cpsi * cphi; R[6] = cth * sphi;  R[7] = 0;
  R[8] = cpsi * sth * cphi + spsi * sphi; R[9] = spsi * sth * cphi - cpsi * sphi; R[10] = cth * cphi; R[11] = 0;
}

// visualization.cpp:11-25 -- rotation (as ODE matrix) whose raw copy into M4 takes z to v.
inline void rot_z_to_v(real R[12], const V4& v) {
  const V4 z(0, 0, 1);
  V4 a = v.cross(z);
  real an = a.norm3();
  if (an < 1e-10) a = V4(0, 1, 0);
  real angle = orc::m_asin(an / v.norm3());
  if (v.dot3(z) < 0) angle = M_PI - angle;
  ode_R_from_axis_angle(R, a.v[0], a.v[1], a.v[2], angle);
  M4 A; A.set_rot_raw12(R);                       // the reference's inline self-check (line 24)
  V4 b = A.apply(z), c = v; real n = c.norm3();
  for (int i = 0; i < 3; i++) c.v[i] /= n;
  c.sub4(b);
  if (c.norm3() > 1e-3) throw Failure("rot_ztov self-check failed");
}
inline M4 m4_from_posrot(const real* pos, const real* R12) {  // visualization.cpp:54-60
  M4 A; A.set_rot_raw12(R12); A.a[12] = pos[0]; A.a[13] = pos[1]; A.a[14] = pos[2]; return A;
}
inline M4 m4_from_orientation(const V4 o[2]) {  // visualization.cpp:62-69
  real R[12]; ode_R_from_euler(R, o[1].v[0], o[1].v[1], o[1].v[2]);
  return m4_from_posrot(o[0].v, R);
}
inline void wrap_pm_pi(real& a) {  // visualization.cpp:73-79
  if (a < -M_PI) { while (a < -M_PI) a += 2 * M_PI; }
  else if (a > M_PI) { while (a > M_PI) a -= 2 * M_PI; }
}
inline void euler_from_m4(const M4& A, real* ang) {  // visualization.cpp:81-101
  real r11 = A.a[0], r21 = A.a[1], r31 = A.a[2], r32 = A.a[6], r33 = A.a[10];
  real th = -orc::m_asin(r31), ct = orc::m_cos(th);
  ang[0] = orc::m_atan2(r32 / ct, r33 / ct);
  ang[1] = th;
  ang[2] = orc::m_atan2(r21 / ct, r11 / ct);
}

// --------------------------------------------------------------- tiny XML reader
struct XNode {
  std::string name;
  std::vector<std::pair<std::string, std::string> > attrs;
  std::vector<XNode> kids;
  const std::string* attr(const char* k) const {
    for (size_t i = 0; i < attrs.size(); i++) if (attrs[i].first == k) return &attrs[i].second;
    return 0;
  }
  const XNode* first(const char* n) const {
    for (size_t i = 0; i < kids.size(); i++) if (kids[i].name == n) return &kids[i];
    return 0;
  }
};
class XParser {
  const std::string& s; size_t p;
  void ws() { while (p < s.size() && std::isspace((unsigned char)s[p])) p++; }
  bool starts(const char* t) const { return s.compare(p, std::strlen(t), t) == 0; }
  void skip_misc() {
    for (;;) {
      while (p < s.size() && s[p] != '<') p++;
      if (starts("<!--")) { p = s.find("-->", p); if (p == std::string::npos) throw Failure("xml: open comment"); p += 3; }
      else if (starts("<?")) { p = s.find("?>", p); if (p == std::string::npos) throw Failure("xml: open PI"); p += 2; }
      else return;
    }
  }
  std::string ident() { size_t b = p; while (p < s.size() && (std::isalnum((unsigned char)s[p]) || s[p] == '_' || s[p] == '-' || s[p] == ':' || s[p] == '.')) p++; return s.substr(b, p - b); }
 public:
  explicit XParser(const std::string& src) : s(src), p(0) {}
  bool element(XNode& out) {
    skip_misc();
    if (p >= s.size() || starts("</")) return false;
    p++;
    out.name = ident();
    for (;;) {
      ws();
      if (starts("/>")) { p += 2; return true; }
      if (s[p] == '>') { p++; break; }
      std::string k = ident(); ws();
      if (s[p] != '=') throw Failure("xml: expected '='");
      p++; ws();
      char qc = s[p++]; size_t e = s.find(qc, p);
      out.attrs.push_back(std::make_pair(k, s.substr(p, e - p))); p = e + 1;
    }
    for (;;) {
      XNode kid;
      if (element(kid)) out.kids.push_back(kid); else break;
    }
    skip_misc();
    if (!starts("</")) throw Failure("xml: expected closing tag");
    p = s.find('>', p) + 1;
    return true;
  }
};
inline int parse_doubles(const std::string& str, real* out, int maxn) {  // core.cpp:8-12
  std::stringstream ss(str); int n = 0; real v = 0;
  while (n < maxn && (ss >> RealIn(v))) out[n++] = v;
  return n;
}

// --------------------------------------------------------------- kinematic tree
enum JointKind { J_NONE, J_FREE6, J_HINGE };

struct Body {  // modelnode + modeljoint + odepart of one XML <body>
  int id, parent;
  std::vector<int> kids;
  M4 A_pj_body, A_ground;
  JointKind jk;
  M4 J_A_parent, J_A_ground;
  int qoff;                 // offset of this joint's values in the configuration vector
  M4 A_body_geom;           // odepart
  V4 capsule_to_pos;
  real rcap;
  std::string name;
  Body() : id(0), parent(-1), jk(J_NONE), qoff(-1), rcap(0) {}
};

class Model;
typedef bool (*LimbSolver)(int limbi, const V4& pos_limb, V4& angles, bool bend, bool ignore_reach);

class Model {  // kinematicmodel (+ liksolver, which the reference keys on the file name)
 public:
  std::vector<Body> b;
  std::vector<real> q;  // joint values: torso x,y,z,phi,theta,psi then hinges in DFS order
  std::string xmlname;
  int lik_index;          // 0 myant, 1 hexapod, 2 spider, -1 none (lik.cpp:8-16)
  std::vector<int> limb_top;
  std::vector<bool> limb_bend;
  real rcap;
  bool ignore_reach;      // the reference's global ignore_reach_flag (lik.cpp:142)

  explicit Model(const std::string& path) : lik_index(-1), rcap(0), ignore_reach(false) {
    std::ifstream f(path.c_str());
    if (!f) throw Failure("cannot open " + path);
    std::stringstream ss; ss << f.rdbuf();
    std::string src = ss.str();
    XParser xp(src); XNode root;
    if (!xp.element(root) || root.name != "mujoco") throw Failure("not a mujoco file");
    const XNode* wb = root.first("worldbody");
    const XNode* tb = wb ? wb->first("body") : 0;
    if (!tb) throw Failure("no worldbody/body");
    size_t sl = path.find_last_of('/');
    xmlname = (sl == std::string::npos) ? path : path.substr(sl + 1);
    M4 I;
    build(*tb, -1, I);
    setup_lik();
    fk();
  }
  int n() const { return (int)b.size(); }
  int config_dim() const { return (int)q.size(); }
  int nmj() const { return (int)q.size() - 6; }
  int nlimbs() const { return (int)limb_top.size(); }

  // model.cpp:183-201, 314-318
  void fk() { M4 I; fk_rec(0, I); }
  M4 joint_transform(const Body& bd) const {  // model.cpp:37-62
    M4 A;
    if (bd.jk == J_FREE6) {
      real R[12];
      ode_R_from_euler(R, q[bd.qoff + 3], q[bd.qoff + 4], q[bd.qoff + 5]);
      A.set_rot_raw12(R);
      A.shift(V4(q[bd.qoff], q[bd.qoff + 1], q[bd.qoff + 2]));
    } else {
      real val = q[bd.qoff], c = orc::m_cos(val), s = orc::m_sin(val);
      A.a[0] = c; A.a[5] = c; A.a[4] = -s; A.a[1] = s;
    }
    return A;
  }
  void set_jvalues(const real* v) { for (size_t i = 0; i < q.size(); i++) q[i] = v[i]; }
  void get_jvalues(real* v) const { for (size_t i = 0; i < q.size(); i++) v[i] = q[i]; }
  void orient_torso(const V4 o[2]) {  // model.cpp:403-409
    for (int i = 0; i < 2; i++) for (int j = 0; j < 3; j++) q[j + 3 * i] = o[i].v[j];
    fk();
  }
  // model.cpp:354-359 ; returns false where the reference would exit(1) (unreachable target)
  bool set_jvalues_with_lik(const real* rec) {
    for (int i = 0; i < 6; i++) q[i] = rec[i];
    fk();
    return place_limbs(rec + 6);
  }
  bool place_limbs(const real* p) {  // lik.cpp:89-99
    if (lik_index < 0) return true;
    for (int i = 0; i < nlimbs(); i++) if (!place_limb(i, V4(p[3 * i], p[3 * i + 1], p[3 * i + 2]))) return false;
    return true;
  }
  bool place_limb(int li, const V4& pos_ground);  // lik.cpp:316-336
  V4 limb_hip_pos(int li) const { return b[limb_top[li]].A_ground.translation_part(); }  // lik.cpp:357-360
  int limb_foot(int li) const { return b[b[limb_top[li]].kids[0]].kids[0]; }              // lik.cpp:363-365
  V4 com_pos(int i) const { return b[i].A_ground.apply(b[i].A_body_geom.translation_part()); }  // visualization.cpp:541-545
  V4 foot_pos(int i) const { return b[i].A_ground.apply(b[i].capsule_to_pos); }                  // visualization.cpp:565

 private:
  int build(const XNode& x, int parent, const M4& A_parent_ground) {  // model.cpp:246-289
    Body bd;
    bd.id = (int)b.size(); bd.parent = parent;
    real pos[3] = {0, 0, 0};
    if (const std::string* a = x.attr("pos")) parse_doubles(*a, pos, 3);
    if (const std::string* a = x.attr("name")) bd.name = *a;
    bd.A_pj_body = M4::translation(pos);
    bd.A_ground = A_parent_ground; bd.A_ground.mul(bd.A_pj_body);
    make_geom(x, bd);
    make_joint(x, bd);
    b.push_back(bd);
    const int id = bd.id;
    for (size_t i = 0; i < x.kids.size(); i++)
      if (x.kids[i].name == "body") {
        M4 Ag = b[id].A_ground;
        int kid = build(x.kids[i], id, Ag);
        b[id].kids.push_back(kid);
      }
    return id;
  }
  void make_geom(const XNode& x, Body& bd) {  // visualization.cpp:442-504 (first <geom> only)
    const XNode* g = x.first("geom");
    if (!g) throw Failure("body without geom");
    const std::string* type = g->attr("type");
    if (!type) throw Failure("geom without type");
    real r = 0;
    if (const std::string* a = g->attr("size")) parse_doubles(*a, &r, 1);
    if (*type == "sphere") {
      real pos[3] = {0, 0, 0};
      if (const std::string* a = g->attr("pos")) parse_doubles(*a, pos, 3);
      bd.A_body_geom = M4::translation(pos);
    } else if (*type == "capsule" || *type == "cylinder") {
      real ft[6] = {0, 0, 0, 0, 0, 0};
      if (const std::string* a = g->attr("fromto")) parse_doubles(*a, ft, 6);
      real pos[3], R[12];
      for (int i = 0; i < 3; i++) pos[i] = (ft[i] + ft[i + 3]) / 2.;
      V4 r1, r2; r1.set3(ft); r2.set3(ft + 3); r2.sub4(r1);
      rot_z_to_v(R, r2);
      bd.capsule_to_pos.set3(ft + 3);
      bd.A_body_geom = m4_from_posrot(pos, R);
      if (*type == "capsule") bd.rcap = r;
    }
  }
  void make_joint(const XNode& x, Body& bd) {  // model.cpp:119-174, 272-289
    const XNode* j = x.first("joint");
    if (!j) return;
    const std::string* type = j->attr("type");
    real pos[3] = {0, 0, 0}, axis[3] = {0, 0, 1};
    if (const std::string* a = j->attr("pos")) parse_doubles(*a, pos, 3);
    if (type && *type == "free") {
      bd.jk = J_FREE6; bd.qoff = (int)q.size(); q.resize(q.size() + 6, 0.0);
      V4 p; p.set3(pos);
      M4 A = bd.A_pj_body; A.shift(p);
      bd.J_A_parent = A;
      M4 B; p.scale3(-1); B.shift(p);
      bd.A_pj_body = B;
    } else if (type && *type == "hinge") {
      if (const std::string* a = j->attr("axis")) parse_doubles(*a, axis, 3);
      bd.jk = J_HINGE; bd.qoff = (int)q.size(); q.resize(q.size() + 1, 0.0);
      V4 p, v; p.set3(pos); v.set3(axis);
      real R[12]; rot_z_to_v(R, v);
      M4 A1; A1.set_rot_raw12(R);
      M4 A = bd.A_pj_body; A.mul(A1); A.shift(p);
      bd.J_A_parent = A;
      bd.A_pj_body = m4_from_posrot(pos, R);
      bd.A_pj_body.invert_rigid();
    }
  }
  void fk_rec(int i, const M4& A) {
    Body& bd = b[i];
    if (bd.jk != J_NONE) {
      bd.J_A_ground = A; bd.J_A_ground.mul(bd.J_A_parent);
      bd.A_ground = bd.J_A_ground;
      M4 J = joint_transform(bd);
      bd.A_ground.mul(J);
      bd.A_ground.mul(bd.A_pj_body);
    } else {
      bd.A_ground = A; bd.A_ground.mul(bd.A_pj_body);
    }
    for (size_t k = 0; k < bd.kids.size(); k++) { M4 Ag = b[i].A_ground; fk_rec(b[i].kids[k], Ag); }
  }
  void setup_lik() {  // lik.cpp:7-20, 44-78, 131-140
    if (xmlname == "myant.xml") lik_index = 0;
    else if (xmlname == "hexapod.xml") lik_index = 1;
    else if (xmlname == "spider.xml") lik_index = 2;
    else { lik_index = -1; return; }
    static const int t0[] = {2, 6, 10, 14}, t1[] = {2, 5, 9, 12, 16, 19}, t2[] = {1, 4, 7, 10, 13, 16};
    if (lik_index == 0) limb_top.assign(t0, t0 + 4);
    if (lik_index == 1) limb_top.assign(t1, t1 + 6);
    if (lik_index == 2) limb_top.assign(t2, t2 + 6);
    limb_bend.assign(limb_top.size(), true);
    for (size_t i = 0; i < limb_top.size(); i++) {
      real r1 = b[limb_top[i] + 2].rcap;
      if (rcap > 0 && r1 != rcap) throw Failure("rcaps must be same for all feet");
      rcap = r1;
    }
  }
};

// ------------------------------------------------------------- limb IK (lik.cpp:151-245)
static const real kLimbLs[3] = {.05, .4, .4};   // hexapod / quadruped links
static const real kLimbLs1[3] = {.1, .4, .4};   // spider links

inline bool limb_solver_yxx(const V4& pos_limb, V4& ja, const real* ls, int ysign, bool bend, bool ignore_reach) {
  real l0 = ls[0], l1 = ls[1], l2 = ls[2];
  int s0 = ysign, s1 = 2 * int(bend) - 1;
  V4 pos0(0, 0, s0 * l0), pos1(pos_limb);
  pos1.sub4(pos0);
  real l = pos1.norm3();
  if (l1 + l2 - l < 0) { if (ignore_reach) l = l1 + l2; else return false; }
  real x1 = pos_limb.v[0], y1 = pos_limb.v[1], z1 = pos_limb.v[2];
  real c = (z1 - s0 * l0) / l;
  real phi = orc::m_atan2(x1, y1);
  real theta = orc::m_acos(c) + (1 - s0) * M_PI / 2;
  wrap_pm_pi(phi); wrap_pm_pi(theta);
  real ll = l * l, del = l2 * l2 - l1 * l1;
  real beta = s1 * s0 * orc::m_acos((ll - del) / (2 * l1 * l));
  real gamma = s1 * s0 * orc::m_acos((ll + del) / (2 * l2 * l));
  ja = V4(-phi, -theta + beta, -(beta + gamma));
  return true;
}
inline bool limb_solver_zxx(const V4& pos_limb, V4& ja, const real* ls, int ysign, bool bend, bool ignore_reach) {
  real l0 = ls[0], l1 = ls[1], l2 = ls[2];
  int s0 = ysign, s1 = 2 * int(bend) - 1;
  V4 pos0(0, 0, l0), pos1(pos_limb);
  pos1.add3(pos0);
  real l = pos1.norm3();
  if (l1 + l2 - l < 0) { if (ignore_reach) l = l1 + l2; else return false; }
  real x = pos_limb.v[0], y = pos_limb.v[1], z = pos_limb.v[2];
  real c = (z + l0) / l;
  real phi = orc::m_atan2(x, y);
  real theta = orc::m_acos(c) - s0 * M_PI / 2;
  wrap_pm_pi(phi); wrap_pm_pi(theta);
  real ll = l * l, del = l2 * l2 - l1 * l1;
  real beta = s1 * orc::m_acos((ll - del) / (2 * l1 * l));
  real gamma = s1 * orc::m_acos((ll + del) / (2 * l2 * l));
  ja = V4(-phi, -theta + beta, -(beta + gamma));
  return true;
}
inline bool limb_solve(int lik_index, int limbi, const V4& p, V4& ja, bool bend, bool ign) {  // lik.cpp:230-245
  switch (lik_index) {
    case 0: return limb_solver_yxx(p, ja, kLimbLs, (limbi < 2) ? 1 : -1, bend, ign);
    case 1: return limb_solver_yxx(p, ja, kLimbLs, (limbi % 2 == 0) ? 1 : -1, bend, ign);
    default: return limb_solver_zxx(p, ja, kLimbLs1, (limbi % 2 == 0) ? 1 : -1, bend, ign);
  }
}
// Forward map of the y-x-x limb in the hip joint frame (lik.cpp:248-274, pos_flag branch); used by the
// IK round-trip pin (lik.cpp:371-404).
inline bool bend_solver_yxx(V4& pos, const V4& ang, const real* ls, int ysign, bool pos_flag) {
  int s0 = ysign;
  real a0 = ang.v[0], a1 = ang.v[1], a2 = ang.v[2];
  if (pos_flag) {
    real l0 = ls[0], l1 = ls[1], l2 = ls[2];
    int s1 = (a2 < 0) ? 1 : -1;
    real ca = orc::m_cos(a2), l = orc::m_sqrt(l1 * l1 + l2 * l2 + 2 * l1 * l2 * ca);
    real theta = s1 * orc::m_acos((l1 + l2 * ca) / l) - a1 + (1 - s0) * M_PI / 2, phi = -a0;
    real stl = orc::m_sin(theta) * l;
    pos = V4(stl * orc::m_sin(phi), stl * orc::m_cos(phi), s0 * l0 + orc::m_cos(theta) * l);
  }
  return (a2 * s0 < 0);
}

inline bool Model::place_limb(int li, const V4& pos_ground) {
  Body& top = b[limb_top[li]];
  // lik.cpp:341-347: hip joint frame from the parent's current ground transform
  top.J_A_ground = b[top.parent].A_ground; top.J_A_ground.mul(top.J_A_parent);
  M4 A = top.J_A_ground; A.invert_rigid();
  V4 pos_limb = A.apply(pos_ground);
  V4 ja;
  if (!limb_solve(lik_index, li, pos_limb, ja, limb_bend[li], ignore_reach)) return false;
  int id = limb_top[li];
  for (int k = 0; k < 3; k++) { q[b[id].qoff] = ja.v[k]; if (k < 2) id = b[id].kids[0]; }
  return true;
}

}  // namespace orc
