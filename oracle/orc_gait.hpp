// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
// Nothing under hslabs_b200/ may include, link or call this file.
//
// CPU restatement of the reference's periodic gait generator:
//   periodicgenerator   pergen.cpp:14-198
//   pergensetup         pergen.cpp:201-397
//   pgssweeper          pergen.cpp:400-507   (candidate construction + 1-D sweep)
//   pgsconfigparams     pergen.h:137-146, pergen.cpp:510-519
//   preset-row parsing  player.cpp:170-208, 230-244
#pragma once
#include "orc_model.hpp"

namespace orc {

struct GaitParams {  // pgsconfigparams
  std::string fname;
  V4 orientation[2];
  real step_duration;
  real TLh[3];
  real curvature;
  int shift_type;      // -1 none, 0 lateral, 1 radial
  real shift_value;
  GaitParams() : step_duration(0), curvature(0), shift_type(-1), shift_value(0) { TLh[0] = TLh[1] = TLh[2] = 0; }
};

class FootPattern {  // periodicgenerator
 public:
  int n;
  real t_step;
  std::vector<real> ts, xs;
  real period, step_length, step_height;
  std::vector<V4> pos0;  // default foot positions, pergen order
  real step_duration, curvature, max_radius;

  explicit FootPattern(int n_) : n(n_), t_step(0), ts(n_), xs(n_), period(0), step_length(0), step_height(0),
                                 step_duration(0), curvature(0), max_radius(0) {
    if (n_ % 2) throw Failure("number of limbs must be even");
  }
  void set_step_duration(real f) {  // pergen.cpp:30-51
    if (f < 0 || f > 1) throw Failure("step_duration out of bounds");
    t_step = f * (1. / 2 - 1. / n) + 1. / n;
    for (int i = 0; i < 2; i++) {
      int jmax = n / 2, z = (jmax == 1) ? 1 : jmax - 1;
      for (int j = 0; j < jmax; j++) {
        int k = j + i * jmax;
        ts[k] = j * (1. / 2 - t_step) / z + real(i) / 2;
        xs[k] = ts[k] + t_step / 2 - 1. / 2;
      }
    }
    step_duration = f;
  }
  void set_pos0s(const std::vector<V4>& p) { pos0 = p; compute_max_radius(); }
  void set_curvature(real c) { curvature = c; compute_max_radius(); }
  real step_frac(int li, real t) const {  // pergen.cpp:73-78
    real tl = ts[li];
    if (t < tl) return 0;
    else if (t < tl + t_step) return (t - tl) / t_step;
    else return 1;
  }
  void limb_positions(real time, std::vector<V4>& out) const {  // pergen.cpp:82-94
    real t = time / period;
    int t_int = int(t);
    real t_frac = t - t_int;
    for (int i = 0; i < n; i++) {
      real sf = step_frac(i, t_frac);
      real delx = (t_int + xs[i] + (1 - orc::m_cos(M_PI * sf)) / 2) * step_length;
      real a = orc::m_sin(M_PI * sf);
      real delz = a * a * step_height;
      turn_position(pos0[i], V4(delx, 0, delz), out[i]);
    }
  }
  void turn_position(const V4& p0, const V4& del, V4& pos) const {  // pergen.cpp:160-183
    real dx = del.v[0], dy = del.v[1], dz = del.v[2];
    if (curvature != 0) {
      int s = (curvature > 0) ? 1 : -1;
      real x0 = p0.v[0], y0 = p0.v[1];
      real rc = 1. / curvature, rx = x0, ry = y0 - rc;
      real r = orc::m_sqrt(rx * rx + ry * ry);
      real alpha = orc::m_atan2(ry, rx), beta = -s * dx / max_radius, gamma = alpha - beta / 2;
      real sb = 2 * orc::m_sin(beta / 2);
      dx = r * orc::m_sin(gamma) * sb;
      dy += -r * orc::m_cos(gamma) * sb;
    }
    V4 d(dx, dy, dz);
    d.add3(p0);
    pos = d;
  }
  void turn_orientation(real dx, V4 o[2]) const {  // pergen.cpp:187-198
    if (curvature != 0) {
      int s = (curvature > 0) ? 1 : -1;
      real psi = s * dx / max_radius, rc = 1. / curvature;
      o[0] = V4(rc * orc::m_sin(psi), rc * (1 - orc::m_cos(psi)), 0);
      o[1] = V4(0, 0, psi);
    } else { o[0] = V4(dx, 0, 0); o[1] = V4(0, 0, 0); }
  }
 private:
  void compute_max_radius() {  // pergen.cpp:144-154
    if (curvature == 0) return;
    real c[3] = {0, 1. / curvature, 0};
    max_radius = 0;
    for (size_t i = 0; i < pos0.size(); i++) {
      real s = 0;
      for (int k = 0; k < 3; k++) { real d = pos0[i].v[k] - c[k]; s += d * d; }
      real rad = orc::m_sqrt(s);
      if (rad > max_radius) max_radius = rad;
    }
  }
};

class GaitSetup {  // pergensetup
 public:
  int n;
  FootPattern pattern;
  std::vector<int> lik2pg;  // LIK limb index -> pergen index (pergen.cpp:243-262)
  std::vector<V4> limb_poss;
  real v;
  V4 torso_pos0, euler;
  M4 rec_transform;
  bool rec_transform_flag;
  int shift_type; real shift_value;

  explicit GaitSetup(int n_) : n(n_), pattern(n_), limb_poss(n_), v(0), rec_transform_flag(false), shift_type(-1), shift_value(0) {
    static const int m4[] = {0, 3, 1, 2}, m6[] = {0, 3, 4, 1, 2, 5};
    if (n == 4) lik2pg.assign(m4, m4 + 4);
    else if (n == 6) lik2pg.assign(m6, m6 + 6);
    else throw Failure("likpergen map undefined for this limb count");
  }
  int config_dim() const { return 6 + 3 * n; }
  real period() const { return pattern.period; }
  void set_TLh(real T, real L, real h) { pattern.period = T; pattern.step_length = L; pattern.step_height = h; v = L / T; }
  void set_limb_pos0(int lik_i, const V4& pos, real rcap) { V4 p(pos); p.v[2] = rcap; limb_poss[lik2pg[lik_i]] = p; }  // pergen.cpp:268-275
  void commit_pos0s() { pattern.set_pos0s(limb_poss); }
  void set_rec_rotation(const V4& eas) {  // pergen.cpp:309-322
    V4 o[2] = {rec_transform.translation_part(), eas};
    rec_transform = m4_from_orientation(o);
    rec_transform_flag = true;
  }
  void set_rec_transform(const V4& transl, const V4& eas) {  // pergen.cpp:316-320 (private in the reference)
    V4 o[2] = {transl, eas};
    rec_transform = m4_from_orientation(o);
    rec_transform_flag = true;
  }
  // pergen.cpp:225-239
  void set_rec(real* rec, real t) {
    V4 o[2] = {torso_pos0, euler};
    turn_torso(t, o);
    for (int k = 0; k < 3; k++) { rec[k] = o[0].v[k]; rec[3 + k] = o[1].v[k]; }
    pattern.limb_positions(t, limb_poss);
    for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) rec[6 + 3 * i + k] = limb_poss[lik2pg[i]].v[k];
    if (rec_transform_flag) transform_rec(rec);
  }
  void params(GaitParams* p) const {  // pergen.cpp:297-306
    p->orientation[0] = torso_pos0; p->orientation[1] = euler;
    p->step_duration = pattern.step_duration;
    p->TLh[0] = pattern.period; p->TLh[1] = pattern.step_length; p->TLh[2] = pattern.step_height;
    p->curvature = pattern.curvature;
    p->shift_type = shift_type; p->shift_value = shift_value;
  }
 private:
  static void transform_orientation(const M4& A, V4 o[2]) {  // pergen.cpp:377-383
    M4 A0 = m4_from_orientation(o), A1 = A;
    A1.mul(A0);
    o[0] = A1.translation_part();
    euler_from_m4(A1, o[1].v);
  }
  void turn_torso(real t, V4 o[2]) const {  // pergen.cpp:386-397
    real tv = t * v;
    V4 to[2];
    pattern.turn_orientation(tv, to);
    if (to[1].v[2] == 0) o[0].v[0] += tv;
    else { M4 A = m4_from_orientation(to); transform_orientation(A, o); }
  }
  void transform_rec(real* rec) const {  // pergen.cpp:325-337
    V4 o[2]; o[0].set3(rec); o[1].set3(rec + 3);
    transform_orientation(rec_transform, o);
    for (int k = 0; k < 3; k++) { rec[k] = o[0].v[k]; rec[3 + k] = o[1].v[k]; }
    for (int i = 0; i < n; i++) {
      V4 p0; p0.set3(rec + 6 + 3 * i);
      V4 p = rec_transform.apply(p0);
      for (int k = 0; k < 3; k++) rec[6 + 3 * i + k] = p.v[k];
    }
  }
};

// pgssweeper::setup_pergen / partial_setup_pergen / setup_foot_shift / shift_pos0 (pergen.cpp:453-507).
// Mutates the model's torso values exactly as the reference does.
inline void setup_gait(GaitSetup& g, Model& model, const GaitParams& p) {
  g.shift_type = p.shift_type; g.shift_value = p.shift_value;
  model.orient_torso(p.orientation);
  g.pattern.set_step_duration(p.step_duration);
  if (g.n != model.nlimbs()) throw Failure("limb count mismatch");
  real rcap = model.rcap;
  V4 lat_shift; real rad_shift = 0;
  if (p.shift_type == 0) lat_shift = model.b[0].A_ground.apply(V4(0, p.shift_value, 0));  // includes the torso translation
  else if (p.shift_type == 1) rad_shift = p.shift_value;
  for (int i = 0; i < g.n; i++) {
    V4 pos = model.limb_hip_pos(i);
    if (p.shift_type == 0) { V4 d = lat_shift; if (i % 2) d.scale3(-1); pos.add3(d); }
    else if (p.shift_type == 1) {
      real x = pos.v[0], y = pos.v[1], f = rad_shift / orc::m_sqrt(x * x + y * y);
      pos.add3(V4(x * f, y * f, 0));
    }
    g.set_limb_pos0(i, pos, rcap);
  }
  g.commit_pos0s();
  g.torso_pos0 = p.orientation[0]; g.euler = p.orientation[1];
  g.set_TLh(p.TLh[0], p.TLh[1], p.TLh[2]);
  g.pattern.set_curvature(p.curvature);
}

// pgssweeper::sweep / next (pergen.cpp:417-449): n_val+1 candidates val0 + i*(val1-val0)/n_val.
class GaitSweeper {
 public:
  const GaitSetup* g0; Model* model; GaitSetup* g;
  int parami, n_val, vali; real val0, delval, val;
  GaitSweeper(const GaitSetup* g0_, Model* m) : g0(g0_), model(m), g(0), parami(-1), n_val(0), vali(0), val0(0), delval(0), val(0) {}
  ~GaitSweeper() { delete g; }
  void sweep(const std::string& name, real v0, real v1, int nv) {
    val0 = v0; n_val = nv; delval = (v1 - v0) / nv; vali = 0;
    static const char* names[] = {"step_duration", "period", "step_length", "step_height"};
    parami = -1;
    for (int i = 0; i < 4; i++) if (name == names[i]) parami = i;
    if (parami < 0) throw Failure("cannot sweep over " + name);
  }
  bool next() {
    if (vali > n_val) { vali = 0; return false; }
    val = val0 + vali * delval; vali++;
    delete g; g = 0;
    GaitParams p; g0->params(&p);
    if (parami == 0) p.step_duration = val; else p.TLh[parami - 1] = val;
    g = new GaitSetup(g0->n);
    setup_gait(*g, *model, p);
    g->rec_transform_flag = g0->rec_transform_flag; g->rec_transform = g0->rec_transform;
    return true;
  }
};

// player.cpp:230-244 + 170-208: preset row "id key value ..." -> GaitParams
inline bool parse_preset_row(const std::string& row, GaitParams& p) {
  std::stringstream ss(row);
  std::string key; real period = 0, sl = 0, sh = 0;
  while (ss >> key) {
    if (key == "xml_file") ss >> p.fname;
    else if (key == "torso_pos") { real x = 0, y = 0, z = 0; ss >> RealIn(x) >> RealIn(y) >> RealIn(z); p.orientation[0] = V4(x, y, z); }
    else if (key == "torso_angles") { real x = 0, y = 0, z = 0; ss >> RealIn(x) >> RealIn(y) >> RealIn(z); p.orientation[1] = V4(x, y, z); }
    else if (key == "step_duration") ss >> RealIn(p.step_duration);
    else if (key == "period") ss >> RealIn(period);
    else if (key == "step_length") ss >> RealIn(sl);
    else if (key == "step_height") ss >> RealIn(sh);
    else if (key == "curvature") ss >> RealIn(p.curvature);
    else if (key == "lateral_foot_shift") { p.shift_type = 0; ss >> RealIn(p.shift_value); }
    else if (key == "radial_foot_shift") { p.shift_type = 1; ss >> RealIn(p.shift_value); }
    else throw Failure("unknown key " + key);
  }
  p.TLh[0] = period; p.TLh[1] = sl; p.TLh[2] = sh;
  return true;
}
inline bool load_preset(const std::string& file, int id, GaitParams& p) {
  std::ifstream f(file.c_str());
  std::string line;
  while (std::getline(f, line)) {
    std::stringstream ss(line); int rid;
    if (!(ss >> rid) || rid != id) continue;
    return parse_preset_row(line.substr(line.find_first_of(" \t") + 1), p);
  }
  return false;
}

}  // namespace orc
