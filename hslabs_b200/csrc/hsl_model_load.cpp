// XML -> packed model block (host side, load time only).
//
// Replaces kinematicmodel::load_fromxml / mnode_from_xnode / make_joint (model.cpp:119-174,224-289),
// odepart::make / make_ccylinder / capsule_lenposrot_from_fromto (visualization.cpp:442-504),
// rot_ztov (visualization.cpp:11-25, with ODE's dRFromAxisAndAngle [ext]), liksolver::set_limbs /
// set_rcap (lik.cpp:44-78,131-140), pergensetup::set_likpergen_map (pergen.cpp:243-262) and
// periodic::set_dynparts (periodic.cpp:34-58).  Instead of a pointer tree it emits the flat constant
// block of hsl_model.h.  Only body@pos, the first <geom> (type,size,fromto|pos) and <joint>
// (type,pos,axis) are read, exactly as in the reference.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "hsl_internal.h"

namespace {

struct Elem {
  std::string tag;
  std::vector<std::string> keys, vals;
  std::vector<Elem> sub;
  const char* get(const char* k) const {
    for (size_t i = 0; i < keys.size(); i++)
      if (keys[i] == k) return vals[i].c_str();
    return nullptr;
  }
  const Elem* child(const char* t) const {
    for (const Elem& e : sub)
      if (e.tag == t) return &e;
    return nullptr;
  }
};

// Minimal recursive-descent reader for the element/attribute subset MuJoCo model files use.
struct Reader {
  const char* p;
  const char* end;
  bool fail = false;
  void skip_space() { while (p < end && (unsigned char)*p <= ' ') p++; }
  bool at(const char* lit) const { size_t n = strlen(lit); return (size_t)(end - p) >= n && memcmp(p, lit, n) == 0; }
  void skip_until(const char* lit) {
    size_t n = strlen(lit);
    while (p < end && !at(lit)) p++;
    if (p < end) p += n; else fail = true;
  }
  void skip_noise() {  // text, comments, declarations
    while (p < end) {
      while (p < end && *p != '<') p++;
      if (at("<!--")) skip_until("-->");
      else if (at("<?")) skip_until("?>");
      else if (at("<!")) skip_until(">");
      else break;
    }
  }
  std::string name() {
    const char* b = p;
    while (p < end && (isalnum((unsigned char)*p) || *p == '_' || *p == '-' || *p == ':' || *p == '.')) p++;
    return std::string(b, p);
  }
  bool read(Elem& e) {
    skip_noise();
    if (p >= end || at("</")) return false;
    p++;  // '<'
    e.tag = name();
    while (!fail) {
      skip_space();
      if (p >= end) { fail = true; return false; }
      if (at("/>")) { p += 2; return true; }
      if (*p == '>') { p++; break; }
      std::string k = name();
      skip_space();
      if (p >= end || *p != '=') { fail = true; return false; }
      p++;
      skip_space();
      if (p >= end || (*p != '"' && *p != '\'')) { fail = true; return false; }
      char qc = *p++;
      const char* b = p;
      while (p < end && *p != qc) p++;
      if (p >= end) { fail = true; return false; }
      e.keys.push_back(k);
      e.vals.push_back(std::string(b, p));
      p++;
    }
    for (;;) {
      Elem kid;
      if (!read(kid)) break;
      e.sub.push_back(std::move(kid));
    }
    skip_noise();
    if (!at("</")) { fail = true; return false; }
    skip_until(">");
    return !fail;
  }
};

int numbers(const char* s, double* out, int cap) {  // whitespace separated doubles (core.cpp:8-12)
  if (!s) return 0;
  std::istringstream ss(s);
  int n = 0;
  double v;
  while (n < cap && (ss >> v)) out[n++] = v;
  return n;
}

// Rotation taking z to v, as a column-major 3x3 (what the reference's affine holds after copying the
// ODE matrix raw): quaternion about a = v x z by the angle between them, transposed.
void rot_z_to(const double* v, double* R) {
  double a[3] = {v[1] * 1 - v[2] * 0, v[2] * 0 - v[0] * 1, v[0] * 0 - v[1] * 0};
  double an = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  if (an < 1e-10) { a[0] = 0; a[1] = 1; a[2] = 0; }
  double vn = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  double angle = std::asin(an / vn);
  if (v[2] < 0) angle = M_PI - angle;
  double q0, q1, q2, q3, l = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
  if (l > 0) {
    angle *= 0.5;
    q0 = std::cos(angle);
    l = std::sin(angle) * (1.0 / std::sqrt(l));
    q1 = a[0] * l; q2 = a[1] * l; q3 = a[2] * l;
  } else { q0 = 1; q1 = q2 = q3 = 0; }
  const double qq1 = 2 * q1 * q1, qq2 = 2 * q2 * q2, qq3 = 2 * q3 * q3;
  // ODE row-major rows become columns here
  R[0] = 1 - qq2 - qq3;           R[1] = 2 * (q1 * q2 - q0 * q3); R[2] = 2 * (q1 * q3 + q0 * q2);
  R[3] = 2 * (q1 * q2 + q0 * q3); R[4] = 1 - qq1 - qq3;           R[5] = 2 * (q2 * q3 - q0 * q1);
  R[6] = 2 * (q1 * q3 - q0 * q2); R[7] = 2 * (q2 * q3 + q0 * q1); R[8] = 1 - qq1 - qq2;
}

struct RawBody {
  int parent = -1;
  std::vector<int> kids;
  double pos[3] = {0, 0, 0};
  int jkind = 0;  // 0 none 1 free 2 hinge
  double jpos[3] = {0, 0, 0}, axis[3] = {0, 0, 1};
  double com[3] = {0, 0, 0}, tip[3] = {0, 0, 0};
  double rcap = 0;
  int geom = 0;  // 1 sphere, 2 capsule, 0 cylinder (does not collide with the ground in the reference)
  double gsize = 0, from[3] = {0, 0, 0};
};

bool flatten(const Elem& e, int parent, std::vector<RawBody>& out, std::string& err) {
  RawBody b;
  b.parent = parent;
  numbers(e.get("pos"), b.pos, 3);
  const Elem* g = e.child("geom");
  if (!g || !g->get("type")) { err = "body without a typed geom"; return false; }
  std::string type = g->get("type");
  double size = 0;
  numbers(g->get("size"), &size, 1);
  b.gsize = size;
  if (type == "sphere") { numbers(g->get("pos"), b.com, 3); b.geom = 1; for (int k = 0; k < 3; k++) b.from[k] = b.tip[k] = b.com[k]; }
  else if (type == "capsule" || type == "cylinder") {
    double ft[6] = {0, 0, 0, 0, 0, 0};
    numbers(g->get("fromto"), ft, 6);
    for (int k = 0; k < 3; k++) { b.com[k] = (ft[k] + ft[k + 3]) / 2.; b.tip[k] = ft[k + 3]; b.from[k] = ft[k]; }
    if (type == "capsule") { b.rcap = size; b.geom = 2; }
  } else { err = "unsupported geom type " + type; return false; }
  if (const Elem* j = e.child("joint")) {
    std::string jt = j->get("type") ? j->get("type") : "";
    numbers(j->get("pos"), b.jpos, 3);
    if (jt == "free") b.jkind = 1;
    else if (jt == "hinge") { b.jkind = 2; numbers(j->get("axis"), b.axis, 3); }
    else { err = "unsupported joint type " + jt; return false; }
  }
  const int id = (int)out.size();
  out.push_back(b);
  for (const Elem& k : e.sub)
    if (k.tag == "body") {
      const int kid = (int)out.size();
      out[id].kids.push_back(kid);
      if (!flatten(k, id, out, err)) return false;
    }
  return true;
}

}  // namespace

int hsl_build_model_pod(const char* xml_path, HslModelPod* pod, char* errbuf, int errlen) {
  auto fail = [&](int code, const std::string& msg) { snprintf(errbuf, errlen, "%s", msg.c_str()); return code; };
  std::ifstream f(xml_path, std::ios::binary);
  if (!f) return fail(-2, std::string("cannot open ") + xml_path);
  std::stringstream buf;
  buf << f.rdbuf();
  const std::string src = buf.str();
  Reader rd{src.data(), src.data() + src.size()};
  Elem root;
  if (!rd.read(root) || rd.fail || root.tag != "mujoco") return fail(-2, "not a mujoco file");
  const Elem* wb = root.child("worldbody");
  const Elem* tb = wb ? wb->child("body") : nullptr;
  if (!tb) return fail(-2, "no worldbody/body");
  std::vector<RawBody> B;
  std::string err;
  if (!flatten(*tb, -1, B, err)) return fail(-3, err);
  const int n = (int)B.size();
  if (n > HSL_MAX_BODIES) return fail(-3, "too many bodies");
  if (B[0].jkind != 1) return fail(-3, "body 0 must carry the free joint");

  HslModelPod& M = *pod;
  memset(&M, 0, sizeof M);
  M.n = n;
  M.g = 1.0;  // dynrec.cpp:294
  // the reference picks its IK solver by file name (lik.cpp:8-16)
  std::string base = xml_path;
  size_t sl = base.find_last_of('/');
  if (sl != std::string::npos) base = base.substr(sl + 1);
  if (base == "myant.xml") M.lik_index = 0;
  else if (base == "hexapod.xml") M.lik_index = 1;
  else if (base == "spider.xml") M.lik_index = 2;
  else return fail(-3, "no limb IK solver is defined for " + base + " (lik.cpp:12-16)");

  // torso: joint frame = translation(body pos + joint pos); body frame offset = -joint pos (model.cpp:151-170)
  for (int k = 0; k < 3; k++) { M.Pt[k] = B[0].pos[k] + B[0].jpos[k]; M.Qt[k] = -B[0].jpos[k]; }
  // trunk bodies and limbs in DFS order
  std::vector<int> trunk_of(n, -1);
  int motor = 0;
  for (int i = 0; i < n; i++) { M.parent[i] = B[i].parent; M.jkind[i] = B[i].jkind; M.motor_of_body[i] = -1; }
  for (int i = 0; i < n; i++) {
    if (i == 0 || (B[i].jkind == 0 && trunk_of[B[i].parent] >= 0)) {
      if (M.ntrunk >= HSL_MAX_TRUNK) return fail(-3, "too many trunk bodies");
      HslTrunkBody& t = M.trunk[M.ntrunk];
      t.body = i;
      t.parent_trunk = (i == 0) ? -1 : trunk_of[B[i].parent];
      for (int k = 0; k < 3; k++) {
        t.off[k] = (i == 0) ? 0.0 : M.trunk[t.parent_trunk].off[k] + B[i].pos[k];
        t.com[k] = B[i].com[k];
      }
      t.mass = 1.0; t.inertia = 1.0;  // ODE dBodyCreate defaults, dynrec.cpp:62-68
      trunk_of[i] = M.ntrunk++;
    } else if (B[i].jkind == 2 && trunk_of[B[i].parent] >= 0) {
      if (M.nf >= HSL_MAX_LIMBS) return fail(-3, "too many limbs");
      HslLimb& L = M.limb[M.nf];
      L.attach = trunk_of[B[i].parent];
      for (int k = 0; k < 3; k++) L.oatt[k] = M.trunk[L.attach].off[k];
      int id = i;
      for (int h = 0; h < 3; h++) {
        const RawBody& b = B[id];
        if (b.jkind != 2) return fail(-3, "limb link without hinge");
        HslHinge& H = L.h[h];
        H.body = id;
        rot_z_to(b.axis, H.Rjp);
        for (int k = 0; k < 3; k++) H.tjp[k] = b.pos[k] + b.jpos[k];
        for (int r = 0; r < 3; r++)
          for (int c = 0; c < 3; c++) H.Rpb[3 * c + r] = H.Rjp[3 * r + c];
        for (int r = 0; r < 3; r++) {  // inverse of [R | jpos]: R^T * (-jpos), summed as the reference does
          double s = 0;
          for (int c = 0; c < 3; c++) s += H.Rpb[3 * c + r] * (-b.jpos[c]);
          H.tpb[r] = s;
        }
        for (int k = 0; k < 3; k++) H.com[k] = b.com[k];
        H.mass = 1.0; H.inertia = 1.0;
        H.aligned = 0;
        {  // axis-aligned hinge at the body origin: lets the kernels rotate two columns instead of two 3x3 products
          double an = std::sqrt(b.axis[0] * b.axis[0] + b.axis[1] * b.axis[1] + b.axis[2] * b.axis[2]);
          for (int k = 0; k < 3 && an > 0; k++)
            if (std::fabs(std::fabs(b.axis[k]) / an - 1.0) < 1e-14 && b.jpos[0] == 0 && b.jpos[1] == 0 && b.jpos[2] == 0)
              H.aligned = (b.axis[k] > 0) ? (k + 1) : -(k + 1);
        }
        M.motor_of_body[id] = motor++;
        if (h < 2) {
          if (b.kids.size() != 1) return fail(-3, "limb link must have exactly one child");
          id = b.kids[0];
        } else {
          if (!b.kids.empty()) return fail(-3, "limbs must have exactly three links");
          for (int k = 0; k < 3; k++) L.foot[k] = b.tip[k];
          if (M.nf > 0 && b.rcap != M.rcap) return fail(-3, "rcaps must be same for all feet (lik.cpp:136)");
          M.rcap = b.rcap;
        }
      }
      if (L.h[0].tpb[0] != 0 || L.h[0].tpb[1] != 0 || L.h[0].tpb[2] != 0) return fail(-3, "hip joint pos must be 0 (hip position would depend on stale joint values)");
      M.nf++;
    } else if (trunk_of[B[i].parent] >= 0) {
      return fail(-3, "unsupported body below the trunk");
    }
  }
  M.nmj = motor;
  M.config_dim = 6 + motor;
  if (M.nf != 4 && M.nf != 6) return fail(-3, "LIK/pergen limb map is only defined for 4 or 6 limbs (pergen.cpp:243-262)");
  // hard-coded tables of the reference, checked against the detected topology
  static const int tops[3][6] = {{2, 6, 10, 14, -1, -1}, {2, 5, 9, 12, 16, 19}, {1, 4, 7, 10, 13, 16}};
  static const int map4[4] = {0, 3, 1, 2}, map6[6] = {0, 3, 4, 1, 2, 5};
  for (int l = 0; l < M.nf; l++) {
    HslLimb& L = M.limb[l];
    if (tops[M.lik_index][l] != L.h[0].body) return fail(-3, "limb top bodies differ from lik.cpp:50-62");
    L.pg_index = (M.nf == 4) ? map4[l] : map6[l];
    L.bend = 1;
    const double l0 = (M.lik_index == 2) ? .1 : .05;
    L.ls[0] = l0; L.ls[1] = .4; L.ls[2] = .4;  // lik.cpp:226-227
    L.inv2l1 = 1.0 / (2 * L.ls[1]); L.inv2l2 = 1.0 / (2 * L.ls[2]);
    L.kind = (M.lik_index == 2) ? HSL_IK_ZXX : HSL_IK_YXX;
    L.ysign = (M.lik_index == 0) ? ((l < 2) ? 1 : -1) : ((l % 2 == 0) ? 1 : -1);  // lik.cpp:231,237,243
  }
  if ((M.lik_index == 0 && M.nf != 4) || (M.lik_index != 0 && M.nf != 6)) return fail(-3, "limb count does not match the model's LIK table");
  return 0;
}


// ---------------------------------------------------------------------------------------------------- simulation constants
#include "hsl_frame.h"

namespace {
void quat_from_colmajor(const double* R, double* q) {  // ODE dRtoQ on the matrix whose element (i,j) is R[3*j+i]
  auto e = [&](int i, int j) { return R[3 * j + i]; };
  const double tr = e(0, 0) + e(1, 1) + e(2, 2);
  double s;
  if (tr >= 0) {
    s = std::sqrt(tr + 1); q[0] = 0.5 * s; s = 0.5 / s;
    q[1] = (e(2, 1) - e(1, 2)) * s; q[2] = (e(0, 2) - e(2, 0)) * s; q[3] = (e(1, 0) - e(0, 1)) * s;
  } else if (e(1, 1) > e(0, 0) && e(1, 1) >= e(2, 2)) {
    s = std::sqrt((e(1, 1) - (e(2, 2) + e(0, 0))) + 1); q[2] = 0.5 * s; s = 0.5 / s;
    q[3] = (e(2, 1) + e(1, 2)) * s; q[1] = (e(1, 0) + e(0, 1)) * s; q[0] = (e(0, 2) - e(2, 0)) * s;
  } else if (e(2, 2) > e(0, 0) && e(2, 2) > e(1, 1)) {
    s = std::sqrt((e(2, 2) - (e(0, 0) + e(1, 1))) + 1); q[3] = 0.5 * s; s = 0.5 / s;
    q[1] = (e(0, 2) + e(2, 0)) * s; q[2] = (e(2, 1) + e(1, 2)) * s; q[0] = (e(1, 0) - e(0, 1)) * s;
  } else {
    s = std::sqrt((e(0, 0) - (e(1, 1) + e(2, 2))) + 1); q[1] = 0.5 * s; s = 0.5 / s;
    q[2] = (e(1, 0) + e(0, 1)) * s; q[3] = (e(0, 2) + e(2, 0)) * s; q[0] = (e(2, 1) - e(1, 2)) * s;
  }
}
void qmul(const double* b, const double* c, double* a) {
  a[0] = b[0] * c[0] - b[1] * c[1] - b[2] * c[2] - b[3] * c[3];
  a[1] = b[0] * c[1] + b[1] * c[0] + b[2] * c[3] - b[3] * c[2];
  a[2] = b[0] * c[2] + b[2] * c[0] + b[3] * c[1] - b[1] * c[3];
  a[3] = b[0] * c[3] + b[3] * c[0] + b[1] * c[2] - b[2] * c[1];
}
// v (world) -> body axes: R^T v, R column-major with element (i,j) at [4*j+i] of an affine
void to_body_axes(const double* A, const double* v, double* o) { for (int j = 0; j < 3; j++) o[j] = A[4 * j] * v[0] + A[4 * j + 1] * v[1] + A[4 * j + 2] * v[2]; }
}  // namespace

// Body poses and quaternions of a configuration from its body frames A [n][16]: pos = A * com, q = quaternion of the rotation.
void hsl_sim_state_from_frames(const HslSimPod* sim, const double* A, double* pos, double* quat) {
  for (int b = 0; b < sim->n; b++) {
    const double* Ab = A + 16 * b;
    for (int i = 0; i < 3; i++) pos[3 * b + i] = Ab[i] * sim->com[b][0] + Ab[4 + i] * sim->com[b][1] + Ab[8 + i] * sim->com[b][2] + Ab[12 + i];
    double R[9];
    for (int j = 0; j < 3; j++) for (int i = 0; i < 3; i++) R[3 * j + i] = Ab[4 * j + i];
    quat_from_colmajor(R, quat + 4 * b);
  }
}

int hsl_build_sim_pod(const char* xml_path, const HslModelPod* pod, HslSimPod* sim, char* errbuf, int errlen) {
  auto fail = [&](int code, const std::string& msg) { snprintf(errbuf, errlen, "%s", msg.c_str()); return code; };
  std::ifstream f(xml_path, std::ios::binary);
  if (!f) return fail(-2, std::string("cannot open ") + xml_path);
  std::stringstream buf;
  buf << f.rdbuf();
  const std::string src = buf.str();
  Reader rd{src.data(), src.data() + src.size()};
  Elem root;
  if (!rd.read(root) || rd.fail) return fail(-2, "not a mujoco file");
  const Elem* wb = root.child("worldbody");
  const Elem* tb = wb ? wb->child("body") : nullptr;
  std::vector<RawBody> B;
  std::string err;
  if (!tb || !flatten(*tb, -1, B, err)) return fail(-3, err);
  const int n = (int)B.size();
  memset(sim, 0, sizeof *sim);
  sim->n = n;
  // zero configuration (every joint value 0): the state in which the reference creates the ODE joints (player.cpp:43-49)
  std::vector<double> q(pod->config_dim, 0.0), A(16 * (size_t)n), J(16 * (size_t)n);
  for (int role = 0; role <= pod->nf; role++) fk_record(*pod, role, q.data(), A.data(), J.data());
  std::vector<double> pos(3 * (size_t)n), quat(4 * (size_t)n);
  for (int b = 0; b < n; b++) {
    HslSimBody& sb = sim->body[b];
    sb.geom = B[b].geom;
    sb.radius = B[b].gsize;
    for (int k = 0; k < 3; k++) { sim->com[b][k] = B[b].com[k]; sb.p0[k] = B[b].from[k] - B[b].com[k]; sb.p1[k] = B[b].tip[k] - B[b].com[k]; }
    sb.mass = 1.0; sb.inertia = 1.0;  // ODE dBodyCreate defaults: the reference never sets a mass (dynrec.cpp:62-68)
  }
  hsl_sim_state_from_frames(sim, A.data(), pos.data(), quat.data());
  int nj = 0;
  for (int b = 1; b < n; b++) {  // kinematicmodel::set_ode_joints: every body with a parent, in body order
    const int p = pod->parent[b];
    HslSimJoint& jt = sim->joint[nj++];
    double qc[4], t4[4];
    if (pod->jkind[b] == 2) {
      jt.kind = 1; jt.b1 = b; jt.b2 = p; jt.motor = pod->motor_of_body[b];   // dJointAttach(hinge, odebody, parent_odebody)
      const double* Jb = J.data() + 16 * b;
      double anc[3] = {Jb[12], Jb[13], Jb[14]}, axis[3] = {Jb[8], Jb[9], Jb[10]}, d[3];
      for (int k = 0; k < 3; k++) d[k] = anc[k] - pos[3 * b + k];
      to_body_axes(A.data() + 16 * b, d, jt.anchor1);
      for (int k = 0; k < 3; k++) d[k] = anc[k] - pos[3 * p + k];
      to_body_axes(A.data() + 16 * p, d, jt.anchor2);
      to_body_axes(A.data() + 16 * b, axis, jt.axis1);
      to_body_axes(A.data() + 16 * p, axis, jt.axis2);
    } else if (pod->jkind[b] == 0) {
      jt.kind = 2; jt.b1 = p; jt.b2 = b; jt.motor = -1;                        // dJointAttach(joint, parent_odebody, odebody)
      double d[3];
      for (int k = 0; k < 3; k++) d[k] = pos[3 * jt.b1 + k] - pos[3 * jt.b2 + k];
      to_body_axes(A.data() + 16 * jt.b1, d, jt.offset);
    } else {
      return fail(-3, "unsupported joint below the torso");
    }
    qc[0] = quat[4 * jt.b1]; qc[1] = -quat[4 * jt.b1 + 1]; qc[2] = -quat[4 * jt.b1 + 2]; qc[3] = -quat[4 * jt.b1 + 3];
    qmul(qc, quat.data() + 4 * jt.b2, t4);
    for (int k = 0; k < 4; k++) jt.qrel[k] = t4[k];
  }
  sim->nj = nj;
  sim->nmotor = pod->nmj;
  return 0;
}
