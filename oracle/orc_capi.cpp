// TEST INFRASTRUCTURE ONLY -- C entry points of the CPU oracle for ctypes
// (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference
// leg).  Nothing under hslabs_b200/ may include, link or call this file.
//
// Candidate parameter vector (13 doubles, same layout as include/hsl.h):
//   [0..2] torso_pos  [3..5] torso_angles (phi,theta,psi)  [6] step_duration
//   [7] period  [8] step_length  [9] step_height  [10] curvature
//   [11] foot-shift type (-1 none, 0 lateral, 1 radial)  [12] foot-shift value
#include <thread>

#include "orc_dynamics.hpp"

using namespace orc;

namespace {
struct Handle {
  std::string path;
  Model* model;
};
GaitParams params_from(const double* p) {
  GaitParams g;
  g.orientation[0] = V4(p[0], p[1], p[2]);
  g.orientation[1] = V4(p[3], p[4], p[5]);
  g.step_duration = p[6];
  g.TLh[0] = p[7]; g.TLh[1] = p[8]; g.TLh[2] = p[9];
  g.curvature = p[10];
  g.shift_type = (int)p[11];
  g.shift_value = p[12];
  return g;
}
void copy16(const M4& m, double* out) { for (int i = 0; i < 16; i++) out[i] = m.a[i]; }
}  // namespace

extern "C" {

void* orc_model_load(const char* path) {
  try {
    Handle* h = new Handle;
    h->path = path;
    h->model = new Model(path);
    return h;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "orc_model_load: %s\n", e.what());
    return 0;
  }
}
void orc_model_free(void* hv) {
  Handle* h = (Handle*)hv;
  if (!h) return;
  delete h->model;
  delete h;
}
void orc_model_dims(void* hv, int* out) {  // n, nf, nmj, config_dim
  Model& m = *((Handle*)hv)->model;
  out[0] = m.n(); out[1] = m.nlimbs(); out[2] = m.nmj(); out[3] = m.config_dim();
}
double orc_model_rcap(void* hv) { return ((Handle*)hv)->model->rcap; }
void orc_set_ignore_reach(void* hv, int flag) { ((Handle*)hv)->model->ignore_reach = (flag != 0); }

// Load-time constants, for checking the product's packed model block.
void orc_model_constants(void* hv, int* parent, int* jkind, double* A_pj_body, double* J_A_parent,
                         double* A_body_geom, double* capsule_to_pos, int* limb_top, int* limb_foot) {
  Model& m = *((Handle*)hv)->model;
  for (int i = 0; i < m.n(); i++) {
    parent[i] = m.b[i].parent; jkind[i] = (int)m.b[i].jk;
    copy16(m.b[i].A_pj_body, A_pj_body + 16 * i);
    copy16(m.b[i].J_A_parent, J_A_parent + 16 * i);
    copy16(m.b[i].A_body_geom, A_body_geom + 16 * i);
    for (int k = 0; k < 3; k++) capsule_to_pos[3 * i + k] = m.b[i].capsule_to_pos.v[k];
  }
  for (int i = 0; i < m.nlimbs(); i++) { limb_top[i] = m.limb_top[i]; limb_foot[i] = m.limb_foot(i); }
}

// FK at configuration q: body frames and joint frames (16 doubles each, column-major).
void orc_fk(void* hv, const double* q, double* A_ground, double* J_A_ground) {
  Model& m = *((Handle*)hv)->model;
  m.set_jvalues(q);
  m.fk();
  for (int i = 0; i < m.n(); i++) { copy16(m.b[i].A_ground, A_ground + 16 * i); copy16(m.b[i].J_A_ground, J_A_ground + 16 * i); }
}
// IK: rec[6+3nf] -> q[config_dim]; returns 0 ok, 1 unreachable.
int orc_ik(void* hv, const double* rec, double* q) {
  Model& m = *((Handle*)hv)->model;
  if (!m.set_jvalues_with_lik(rec)) return 1;
  m.get_jvalues(q);
  return 0;
}
// y-x-x limb round trip helpers (lik.cpp:248-274 forward map, 151-184 inverse)
void orc_limb_forward_yxx(int ysign, const double* ang, double* pos) {
  V4 p, a(ang[0], ang[1], ang[2]);
  bend_solver_yxx(p, a, kLimbLs, ysign, true);
  for (int k = 0; k < 3; k++) pos[k] = p.v[k];
}
int orc_limb_bend_yxx(int ysign, const double* ang) {
  V4 p, a(ang[0], ang[1], ang[2]);
  return bend_solver_yxx(p, a, kLimbLs, ysign, false) ? 1 : 0;
}
int orc_limb_inverse_yxx(int ysign, const double* pos, int bend, double* ang) {
  V4 p(pos[0], pos[1], pos[2]), a;
  bool ok = limb_solver_yxx(p, a, kLimbLs, ysign, bend != 0, false);
  for (int k = 0; k < 3; k++) ang[k] = a.v[k];
  return ok ? 0 : 1;
}
void orc_euler_roundtrip(const double* pos, const double* ang, double* pos_out, double* ang_out) {
  V4 o[2] = {V4(pos[0], pos[1], pos[2]), V4(ang[0], ang[1], ang[2])};
  M4 A = m4_from_orientation(o);
  for (int k = 0; k < 3; k++) pos_out[k] = A.a[12 + k];
  euler_from_m4(A, ang_out);
}
void orc_rot_z_to_v(const double* v, double* m16) {
  double R[12]; rot_z_to_v(R, V4(v[0], v[1], v[2]));
  M4 A; A.set_rot_raw12(R); copy16(A, m16);
}

// Candidate construction (a1): default foot positions (pergen order), lift-off tables.
int orc_gait_setup(void* hv, const double* params, double* pos0, double* ts, double* xs, double* scal /*t_step,v,max_radius*/) {
  Model& m = *((Handle*)hv)->model;
  try {
    GaitSetup g(m.nlimbs());
    setup_gait(g, m, params_from(params));
    for (int i = 0; i < g.n; i++) { for (int k = 0; k < 3; k++) pos0[3 * i + k] = g.pattern.pos0[i].v[k]; ts[i] = g.pattern.ts[i]; xs[i] = g.pattern.xs[i]; }
    scal[0] = g.pattern.t_step; scal[1] = g.v; scal[2] = g.pattern.max_radius;
    return 0;
  } catch (const std::exception&) { return -1; }
}
// Frame record (a2) at time t: rec[6+3nf].
int orc_gait_rec(void* hv, const double* params, double t, double* rec) {
  Model& m = *((Handle*)hv)->model;
  try {
    GaitSetup g(m.nlimbs());
    setup_gait(g, m, params_from(params));
    g.set_rec(rec, t);
    return 0;
  } catch (const std::exception&) { return -1; }
}

// measure_cot for one candidate.  out4 = cot, work, min_cfz, max_mu.  Optional dumps (may be NULL):
// traj[(n_t+5)][config_dim], x[n_t][6n], z[n_t][3nf], tau[n_t][nmj] (frames 2..n_t+1 in order).
// returns 0 ok, 1 IK unreachable, 2 solver breakdown, -1 bad parameters.
int orc_measure_cot(void* hv, const double* params, int n_t, double* out4, double* traj, double* x, double* z, double* tau) {
  Model& m = *((Handle*)hv)->model;
  try {
    GaitSetup g(m.nlimbs());
    setup_gait(g, m, params_from(params));
    CotResult r = measure_cot(m, g, n_t, traj, x, z, tau);
    out4[0] = r.cot; out4[1] = r.work; out4[2] = r.min_cfz; out4[3] = r.max_mu;
    return r.status;
  } catch (const std::exception&) { return -1; }
}

// measure_cot with pergensetup::rec_transform active (pergen.cpp:309-335): rec_transl / rec_eas [3] each.
int orc_measure_cot_rect(void* hv, const double* params, const double* rec_transl, const double* rec_eas, int n_t, double* out4,
                         double* traj, double* x, double* z, double* tau) {
  Model& m = *((Handle*)hv)->model;
  try {
    GaitSetup g(m.nlimbs());
    setup_gait(g, m, params_from(params));
    g.set_rec_transform(V4(rec_transl[0], rec_transl[1], rec_transl[2]), V4(rec_eas[0], rec_eas[1], rec_eas[2]));
    CotResult r = measure_cot(m, g, n_t, traj, x, z, tau);
    out4[0] = r.cot; out4[1] = r.work; out4[2] = r.min_cfz; out4[3] = r.max_mu;
    return r.status;
  } catch (const std::exception&) { return -1; }
}

// Per-frame dynrec fields of the solved frames 2..n_t+1 (inputs of the frame-solve entry).
// Layout [frame][body][3] for pos,jpos,jzaxis,mom_rate,ang_mom_rate; fpos [frame][nf][3]; contacts [frame][nf].
int orc_frame_fields(void* hv, const double* params, int n_t, double* pos, double* jpos, double* jzaxis,
                     double* mom_rate, double* ang_mom_rate, double* fpos, unsigned char* contacts) {
  Model& m = *((Handle*)hv)->model;
  try {
    GaitSetup g(m.nlimbs());
    setup_gait(g, m, params_from(params));
    GaitEvaluator ev(&m);
    if (!ev.record_trajectory(&g, n_t)) return 1;
    ev.compute_dynrecs();
    ev.compute_dynrec_ders();
    const int n = ev.n, nf = ev.nf;
    for (int t = 0; t < n_t; t++) {
      const FrameRecord& r = ev.recs[t + 2];
      for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) {
          size_t o = ((size_t)t * n + i) * 3 + k;
          pos[o] = r.pos[i].v[k]; jpos[o] = r.jpos[i].v[k]; jzaxis[o] = r.jzaxis[i].v[k];
          mom_rate[o] = r.mom_rate[i].v[k]; ang_mom_rate[o] = r.ang_mom_rate[i].v[k];
        }
      for (int fi = 0; fi < nf; fi++) {
        for (int k = 0; k < 3; k++) fpos[((size_t)t * nf + fi) * 3 + k] = r.fpos[fi].v[k];
        contacts[(size_t)t * nf + fi] = (unsigned char)r.contacts[fi];
      }
    }
    return 0;
  } catch (const std::exception&) { return -1; }
}

// Evaluate an externally supplied joint trajectory q[(n_t+5)][config_dim] with step dt (L2 entry).
int orc_eval_trajectory(void* hv, const double* q, int n_t, double dt, double* out3 /*work,min_cfz,max_mu*/,
                        double* x, double* z, double* tau) {
  Model& m = *((Handle*)hv)->model;
  try {
    GaitEvaluator ev(&m);
    ev.set_trajectory(q, n_t, dt);
    ev.compute_dynrecs();
    ev.compute_dynrec_ders();
    ev.switch_torso_penalty(true, true);
    double work = 0;
    if (!ev.work_over_period(work, x, z)) return 2;
    out3[0] = work; out3[1] = ev.min_cfz; out3[2] = ev.max_mu;
    if (tau) for (int i = 2; i < n_t + 2; i++) for (int j = 0; j < ev.nmj; j++) tau[(size_t)(i - 2) * ev.nmj + j] = ev.torques[i % n_t][j];
    return 0;
  } catch (const std::exception&) { return -1; }
}

// test_dynamics (playerexperim.cpp:95-121): solve frame `frame`, then recover the contact forces of
// all feet from the motor torques.  cf, cf1: [3nf]; tau: [nmj].
int orc_test_dynamics(void* hv, const double* params, int n_t, int frame, double* cf, double* cf1, double* tau) {
  Model& m = *((Handle*)hv)->model;
  try {
    GaitSetup g(m.nlimbs());
    setup_gait(g, m, params_from(params));
    GaitEvaluator ev(&m);
    if (!ev.record_trajectory(&g, n_t)) return 1;
    ev.compute_dynrecs();
    ev.compute_dynrec_ders();
    ev.switch_torso_penalty(true, true);
    Vec x, y;
    if (!ev.solve_forcetorques(ev.recs[frame], x, y)) return 2;
    ev.motor_torques(tau);
    for (size_t i = 0; i < y.size(); i++) cf[i] = y[i];
    ev.solve_forces(ev.recs[frame], tau, cf1);
    return 0;
  } catch (const std::exception&) { return -1; }
}

// periodic::solve_contforces_given_torques (periodic.cpp:369-374) on every solved frame 2..n_t+1 of a candidate:
// tau [n_t][nmj] given motor torques (any values), cf [n_t][3nf] least-squares contact forces of all feet.
int orc_solve_forces_frames(void* hv, const double* params, int n_t, const double* tau, double* cf) {
  Model& m = *((Handle*)hv)->model;
  try {
    GaitSetup g(m.nlimbs());
    setup_gait(g, m, params_from(params));
    GaitEvaluator ev(&m);
    if (!ev.record_trajectory(&g, n_t)) return 1;
    ev.compute_dynrecs();
    ev.compute_dynrec_ders();
    for (int t = 0; t < n_t; t++) ev.solve_forces(ev.recs[t + 2], tau + (size_t)t * ev.nmj, cf + (size_t)t * 3 * ev.nf);
    return 0;
  } catch (const std::exception&) { return -1; }
}

// measure_cot_sweep (player.cpp:311-321): n_val+1 candidates; vals/cots: [n_val+1].
int orc_measure_cot_sweep(void* hv, const double* params, int n_t, const char* name, double v0, double v1, int n_val,
                          double* vals, double* cots) {
  Model& m = *((Handle*)hv)->model;
  try {
    GaitSetup g0(m.nlimbs());
    setup_gait(g0, m, params_from(params));
    GaitSweeper sw(&g0, &m);
    sw.sweep(name, v0, v1, n_val);
    int i = 0;
    while (sw.next()) {
      CotResult r = measure_cot(m, *sw.g, n_t);
      vals[i] = sw.val; cots[i] = (r.status == 0) ? r.cot : NAN; i++;
    }
    return 0;
  } catch (const std::exception&) { return -1; }
}

// Batch of candidates (AoS params[C][13]) on `nthreads` host threads, one private model per thread
// (the reference mutates its model in place and is single-threaded; player.cpp:311-321).
int orc_eval_batch(void* hv, long n_cand, int n_t, const double* params, double* cot, double* work, double* min_cfz,
                   double* max_mu, int* status, int nthreads) {
  Handle* h = (Handle*)hv;
  if (nthreads < 1) nthreads = 1;
  std::vector<std::thread> pool;
  for (int w = 0; w < nthreads; w++)
    pool.push_back(std::thread([=]() {
      Model* m = 0;
      try { m = new Model(h->path); } catch (...) { return; }
      m->ignore_reach = h->model->ignore_reach;
      for (long c = w; c < n_cand; c += nthreads) {
        int st = -1; CotResult r; r.cot = r.work = r.min_cfz = r.max_mu = NAN;
        try {
          GaitSetup g(m->nlimbs());
          setup_gait(g, *m, params_from(params + 13 * c));
          r = measure_cot(*m, g, n_t);
          st = r.status;
        } catch (const std::exception&) { st = -1; }
        cot[c] = r.cot; work[c] = r.work; min_cfz[c] = r.min_cfz; max_mu[c] = r.max_mu; status[c] = st;
      }
      delete m;
    }));
  for (size_t i = 0; i < pool.size(); i++) pool[i].join();
  return 0;
}

// Preset row -> params[13]; model file name copied to xml_name (<= 63 chars). 0 ok, 1 not found.
int orc_load_preset(const char* file, int id, double* params, char* xml_name) {
  try {
    GaitParams p;
    if (!load_preset(file, id, p)) return 1;
    for (int k = 0; k < 3; k++) { params[k] = p.orientation[0].v[k]; params[3 + k] = p.orientation[1].v[k]; }
    params[6] = p.step_duration; params[7] = p.TLh[0]; params[8] = p.TLh[1]; params[9] = p.TLh[2];
    params[10] = p.curvature; params[11] = p.shift_type; params[12] = p.shift_value;
    std::strncpy(xml_name, p.fname.c_str(), 63); xml_name[63] = 0;
    return 0;
  } catch (const std::exception&) { return -1; }
}

}  // extern "C"
