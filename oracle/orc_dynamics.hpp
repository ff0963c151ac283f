// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
// Nothing under hslabs_b200/ may include, link or call this file.
//
// CPU restatement ("variant A", reference-shaped) of the reference's per-frame
// dynamics and contact force/torque solve:
//   dynpart / dynrecord       dynrec.cpp:5-93, 134-155, 175-224, 227-344
//   forcetorquesolver         ftsolver.cpp:5-14, 78-146, 150-303, 305-378
//   periodic                  periodic.cpp:10-58, 77-96, 149-160, 185-202, 261-391, 408-426
//   measure_cot(_sweep)       player.cpp:259-285, 311-321 ; test_dynamics playerexperim.cpp:95-121
// The force-torque matrix B is assembled entry by entry exactly as the
// reference inserts it into its sparse matrix, then handed to the dense
// stand-ins of orc_linalg.hpp.  Body mass = 1 and inertia = identity for every
// body: that is what ODE's dBodyCreate gives and the reference never calls a
// dMassSet* function (dynrec.cpp:62-68).
#pragma once
#include "orc_gait.hpp"
#include "orc_linalg.hpp"

namespace orc {

struct FrameRecord {  // dynrecord
  int n, nf;
  std::vector<V4> pos, jpos, vel, mom, mom_rate, acc, ust, ang_vel, ang_mom, ang_mom_rate, fpos, jzaxis;
  std::vector<M4> rot;
  std::vector<char> contacts;
  FrameRecord(int n_, int nf_) : n(n_), nf(nf_), pos(n_), jpos(n_), vel(n_), mom(n_), mom_rate(n_), acc(n_), ust(n_),
                                 ang_vel(n_), ang_mom(n_), ang_mom_rate(n_), fpos(nf_), jzaxis(n_), rot(n_), contacts(nf_, 0) {}
  int ncontacts() const { int s = 0; for (int i = 0; i < nf; i++) s += contacts[i]; return s; }
};

inline void central_diff(std::vector<V4>& der, const std::vector<V4>& prev, const std::vector<V4>& next, real dt) {
  for (size_t i = 0; i < der.size(); i++) {  // dynrec.cpp:192-203
    der[i] = next[i];
    der[i].sub4(prev[i]);
    der[i].scale3(1. / (2 * dt));
  }
}

// B entries for r x F (dynrec.cpp:253-262): rows 3(n+i).., columns 3j..
inline void add_cross(Mat& B, int i, int j, const V4& r, int n) {
  int k = 3 * (n + i), k1 = 3 * j;
  for (int l = 0; l < 3; l++) {
    int d0 = l % 3, d1 = (l + 1) % 3, d2 = (l + 2) % 3;
    B(k + d0, k1 + d1) = -r.v[d2];
    B(k + d1, k1 + d0) = r.v[d2];
  }
}
inline void add_unit(Mat& B, int i, int pi) {  // dynrec.cpp:239-247
  for (int j = 0; j < 3; j++) {
    B(3 * i + j, 3 * i + j) = 1;
    if (pi >= 0) B(3 * pi + j, 3 * i + j) = -1;
  }
}

class GaitEvaluator {  // periodic + forcetorquesolver
 public:
  Model* model;
  int n, nf, nmj, n_t, traj_size, config_dim;
  std::vector<int> parentis, footis, hinge_ids;
  std::vector<char> is_foot;
  std::vector<real> masses;
  std::vector<M4> inertia;
  real dt, rcap, min_cfz, max_mu;
  std::vector<std::vector<real> > traj, vel_traj, torques;
  std::vector<FrameRecord> recs;
  bool pen_force, pen_torque, mask_set;
  Vec fts, jz;  // last solution, last joint z axes
  int solver_warnings;

  explicit GaitEvaluator(Model* m) : model(m), n_t(0), traj_size(0), config_dim(0), dt(0), rcap(0), min_cfz(0), max_mu(0),
                                     pen_force(false), pen_torque(false), mask_set(false), solver_warnings(0) {
    n = m->n(); nf = m->nlimbs(); nmj = m->nmj();
    is_foot.assign(n, 0);
    for (int i = 0; i < nf; i++) is_foot[m->limb_foot(i)] = 1;
    for (int i = 0; i < n; i++) {  // periodic.cpp:34-58, 310-318
      parentis.push_back(m->b[i].parent);
      masses.push_back(1.0);
      inertia.push_back(M4());
      if (is_foot[i]) footis.push_back(i);
      if (m->b[i].jk == J_HINGE) hinge_ids.push_back(i);
    }
  }
  real total_mass() const { real s = 0; for (int i = 0; i < n; i++) s += masses[i]; return s; }

  // periodic.cpp:77-96 ; false where the reference would exit(1) in IK
  bool record_trajectory(GaitSetup* g, int n_t_) {
    n_t = n_t_; traj_size = n_t + 5; config_dim = g->config_dim();
    real t = 0;
    dt = g->period() / n_t;
    std::vector<real> rec(config_dim);
    traj.assign(traj_size, std::vector<real>(config_dim));
    for (int i = 0; i < traj_size; i++) {
      g->set_rec(&rec[0], t);
      if (!model->set_jvalues_with_lik(&rec[0])) return false;
      model->get_jvalues(&traj[i][0]);
      t += dt;
    }
    rcap = model->rcap;
    return true;
  }
  void set_trajectory(const real* q, int n_t_, real dt_) {  // externally supplied trajectory (L2 entry tests)
    n_t = n_t_; traj_size = n_t + 5; config_dim = model->config_dim(); dt = dt_; rcap = model->rcap;
    traj.assign(traj_size, std::vector<real>(config_dim));
    for (int i = 0; i < traj_size; i++) for (int j = 0; j < config_dim; j++) traj[i][j] = q[(size_t)i * config_dim + j];
  }
  void compute_dynrecs() {  // periodic.cpp:149-160 + dynrec.cpp:31-93, 134-155
    recs.assign(traj_size, FrameRecord(n, nf));
    for (int t = 0; t < traj_size; t++) {
      model->set_jvalues(&traj[t][0]);
      model->fk();
      FrameRecord& r = recs[t];
      int fi = 0;
      for (int i = 0; i < n; i++) {
        const Body& bd = model->b[i];
        r.pos[i] = model->com_pos(i);
        r.jpos[i] = (bd.jk != J_NONE) ? bd.J_A_ground.translation_part() : bd.A_ground.translation_part();
        const M4& A = bd.A_ground;
        r.ust[i] = V4((A.at(2, 1) - A.at(1, 2)) / 2, (A.at(0, 2) - A.at(2, 0)) / 2, (A.at(1, 0) - A.at(0, 1)) / 2);
        r.rot[i].set_rot_from(A);
        if (is_foot[i]) {
          r.fpos[fi] = model->foot_pos(i);
          r.contacts[fi] = (r.fpos[fi].v[2] < rcap + 1e-4);
          fi++;
        }
        if (bd.jk != J_NONE) r.jzaxis[i].set3(bd.J_A_ground.a + 8); else r.jzaxis[i] = V4(0, 0, 0);
      }
    }
  }
  void compute_dynrec_ders() {  // periodic.cpp:192-202
    for (int i = 0; i < traj_size + 2; i++)
      for (int j = 0; j < 2; j++)
        if (i > 2 * j && i < traj_size - 1) stage(j, recs[i - j], recs[i - 1 - j], recs[i + 1 - j]);
  }
  void switch_torso_penalty(bool f, bool t) { pen_force = f; pen_torque = t; mask_set = (f || t); }

  // ---- one frame: forcetorquesolver::solve_forcetorques (ftsolver.cpp:78-102)
  // returns false on a solver breakdown the reference would loop on forever
  bool solve_forcetorques(const FrameRecord& r, Vec& x, Vec& z) {
    load_jz(r);
    const int m0 = 6 * n;
    Mat B(m0, m0); Vec f(m0, 0.0);
    assemble(r, B, f);
    {  // particular solution with zero contact forces (ftsolver.cpp:107-113)
      HouseholderQR qr(B);
      x = qr.solve(f);
    }
    // null space of B extended by the contact columns (ftsolver.cpp:116-146)
    const int delm = 3 * r.ncontacts(), m = m0 + delm;
    Mat Bt(m, m0);  // transpose of the (padded) extended matrix, zero columns dropped
    for (int j = 0; j < m0; j++) for (int i = 0; i < m0; i++) Bt(j, i) = B(i, j);
    {
      int ci = 0;
      for (int fi = 0; fi < nf; fi++) {
        if (!r.contacts[fi]) continue;
        int i = footis[fi];
        Mat C(m0, 3);  // contact columns (dynrec.cpp:327-344)
        for (int j = 0; j < 3; j++) C(3 * i + j, j) = 1;
        V4 rr(r.fpos[fi]); rr.sub4(r.pos[i]);
        Mat Ctmp(m0, m0 + 3);
        add_cross(Ctmp, i, 2 * n, rr, n);  // writes rows 3(n+i).., columns 6n..6n+2
        for (int a = 0; a < m0; a++) for (int j = 0; j < 3; j++) C(a, j) += Ctmp(a, m0 + j);
        for (int a = 0; a < m0; a++) for (int j = 0; j < 3; j++) Bt(m0 + 3 * ci + j, a) = C(a, j);
        ci++;
      }
    }
    // No foot on the ground: the reference goes on with 0-column matrices, Eigen's kernel() / image() of the
    // 0 x 0 level-0 matrix come back 0 x 1 and the comma initialiser of ftsolver.cpp:222-223 asserts (abort);
    // oracle/_ref reproduces that, so it is a breakdown here too.
    if (delm == 0) return false;
    Mat N(m0, delm);
    if (delm > 0) {
      HouseholderQR qrt(Bt);
      for (int i = 0; i < delm; i++) {
        Vec col = qrt.q_times_unit(m - 1 - i);
        for (int a = 0; a < m0; a++) N(a, i) = col[a];  // conservativeResize keeps the top 6n rows
      }
    }
    Vec y;
    if (!solve_contact_forces(x, y, N)) return false;
    z.assign(3 * nf, 0.0);  // z = -N_cont y (ftsolver.cpp:90-92, 276-284)
    for (int i = 0; i < nf; i++)
      for (int j = 0; j < 3; j++) {
        real s = 0;
        for (int c = 0; c < delm; c++) s += N(3 * footis[i] + j, c) * y[c];
        z[3 * i + j] = -s;
      }
    Vec dx = matvec(N, y);
    for (int a = 0; a < m0; a++) x[a] += dx[a];
    fts = x;
    return true;
  }
  // periodic.cpp:328-343
  void motor_torques(real* out) const {
    for (size_t h = 0; h < hinge_ids.size(); h++) {
      int k = 3 * hinge_ids[h], k1 = 3 * n + k;
      real s = 0;
      for (int j = 0; j < 3; j++) s += jz[k + j] * fts[k1 + j];
      out[h] = s;
    }
  }
  // periodic.cpp:377-391 ; optional per-frame dumps (x [n_t][6n], z [n_t][3nf]) in solve order i = 2..n_t+1
  bool compute_torques_over_period(real* x_dump = 0, real* z_dump = 0) {
    torques.assign(n_t, std::vector<real>(nmj));
    min_cfz = 1e10; max_mu = -1e10;
    for (int i = 2; i < n_t + 2; i++) {
      Vec x, y;
      if (!solve_forcetorques(recs[i], x, y)) return false;
      for (int fi = 0; fi < nf; fi++) {  // periodic.cpp:347-357
        real cx = y[3 * fi], cy = y[3 * fi + 1], cz = y[3 * fi + 2];
        if (cz < min_cfz) min_cfz = cz;
        real mu = orc::m_sqrt(cx * cx + cy * cy) / cz;
        if (mu > max_mu) max_mu = mu;
      }
      motor_torques(&torques[i % n_t][0]);
      if (x_dump) for (int a = 0; a < 6 * n; a++) x_dump[(size_t)(i - 2) * 6 * n + a] = x[a];
      if (z_dump) for (int a = 0; a < 3 * nf; a++) z_dump[(size_t)(i - 2) * 3 * nf + a] = y[a];
    }
    return true;
  }
  void compute_vel_traj() {  // periodic.cpp:261-282
    vel_traj.assign(traj_size, std::vector<real>(config_dim, 0.0));
    for (int i = 2; i < traj_size; i++)
      for (int j = 0; j < config_dim; j++) {
        real d = traj[i][j] - traj[i - 2][j];
        if (d > M_PI) d -= 2 * M_PI; else if (d < -M_PI) d += 2 * M_PI;
        vel_traj[i - 1][j] = d / (2 * dt);
      }
  }
  bool work_over_period(real& work, real* x_dump = 0, real* z_dump = 0) {  // periodic.cpp:285-307
    if (!compute_torques_over_period(x_dump, z_dump)) return false;
    compute_vel_traj();
    real wp = 0;
    for (int i = 2; i < n_t + 2; i++) {
      real wd = 0;
      for (int j = 0; j < nmj; j++) {
        real dw = torques[i % n_t][j] * vel_traj[i][6 + j];
        dw = (dw > 0) ? dw : 0;
        wd += dw;
      }
      wd *= dt;
      wp += wd;
    }
    work = wp;
    return true;
  }
  // ftsolver.cpp:331-378 : contact forces of ALL feet for given motor torques, torso columns zeroed
  void solve_forces(const FrameRecord& r, const real* tau, real* cf) {
    load_jz(r);
    const int m0 = 6 * n, m1 = m0 + 3 * nf;
    Mat B0(m0, m0); Vec f(m0, 0.0);
    assemble(r, B0, f);
    Mat B(m0 + nmj, m1);
    for (int j = 0; j < m0; j++) for (int i = 0; i < m0; i++) B(i, j) = B0(i, j);
    for (int fi = 0; fi < nf; fi++) {
      int i = footis[fi];
      for (int j = 0; j < 3; j++) B(3 * i + j, m0 + 3 * fi + j) = 1;
      V4 rr(r.fpos[fi]); rr.sub4(r.pos[i]);
      Mat T(m0, m1);
      add_cross(T, i, 2 * n + fi, rr, n);
      for (int a = 0; a < m0; a++) for (int j = 0; j < 3; j++) B(a, m0 + 3 * fi + j) += T(a, m0 + 3 * fi + j);
    }
    f.resize(m0 + nmj);
    for (size_t h = 0; h < hinge_ids.size(); h++) {
      int k = 3 * hinge_ids[h], k1 = 3 * n + k;
      for (int j = 0; j < 3; j++) B(m0 + (int)h, k1 + j) = jz[k + j];
      f[m0 + h] = tau[h];
    }
    for (int i = 0; i < 2; i++) for (int c = 0; c < 3; c++) for (int a = 0; a < B.r; a++) B(a, 3 * n * i + c) = 0;
    HouseholderQR qr(B, true);
    Vec x = qr.solve(f);
    for (int a = 0; a < 3 * nf; a++) cf[a] = x[m0 + a];
  }

 private:
  void stage(int s, FrameRecord& r, const FrameRecord& prev, const FrameRecord& next) {  // dynrec.cpp:175-224
    if (s == 0) {
      central_diff(r.vel, prev.pos, next.pos, dt);
      central_diff(r.ang_vel, prev.ust, next.ust, dt);
      for (int i = 0; i < n; i++) { r.mom[i] = r.vel[i]; r.mom[i].scale3(masses[i]); }
      for (int i = 0; i < n; i++) {
        M4 rt = r.rot[i]; rt.transpose4();
        V4 v = rt.apply(r.ang_vel[i]);
        V4 u = inertia[i].apply(v);
        r.ang_mom[i] = r.rot[i].apply(u);
      }
    } else {
      central_diff(r.mom_rate, prev.mom, next.mom, dt);
      central_diff(r.acc, prev.vel, next.vel, dt);
      central_diff(r.ang_mom_rate, prev.ang_mom, next.ang_mom, dt);
    }
  }
  void load_jz(const FrameRecord& r) {
    jz.assign(3 * n, 0.0);
    for (int i = 0; i < n; i++) for (int j = 0; j < 3; j++) jz[3 * i + j] = r.jzaxis[i].v[j];
  }
  void assemble(const FrameRecord& r, Mat& B, Vec& f) const {  // dynrec.cpp:227-297
    for (int i = 0; i < n; i++) {
      int pi = parentis[i];
      add_unit(B, i, pi);
      for (int j = 0; j < 3; j++) f[3 * i + j] = r.mom_rate[i].v[j];
      add_unit(B, i + n, (pi < 0) ? pi : pi + n);
      if (pi >= 0) {
        V4 rr(r.jpos[i]); rr.sub4(r.pos[i]);
        add_cross(B, i, i, rr, n);
        rr = r.pos[pi]; rr.sub4(r.jpos[i]);
        add_cross(B, pi, i, rr, n);
      }
      for (int j = 0; j < 3; j++) f[3 * (n + i) + j] = r.ang_mom_rate[i].v[j];
      f[3 * i + 2] += masses[i] * 1.0;  // g = 1 (dynrec.cpp:293-297)
    }
  }
  // ftsolver.cpp:185-236 with 239-246 (penalties) and 253-303 (masks)
  bool solve_contact_forces(const Vec& x, Vec& y, const Mat& N) {
    if (!mask_set) throw Failure("mask0 not set");
    const int m0 = 6 * n, K = N.c;
    Vec c(m0, 1.0);
    for (int i = 3; i < 3 * n; i++) c[i] = 0;
    for (int i = 3; i < 3 * n; i++) c[3 * n + i] = jz[i];
    std::vector<int> rows0, rows1;
    std::vector<char> in0(m0, 0);
    if (pen_force) for (int j = 0; j < 3; j++) in0[j] = 1;
    if (pen_torque) for (int j = 0; j < 3; j++) in0[3 * n + j] = 1;
    for (int i = 0; i < m0; i++) (in0[i] ? rows0 : rows1).push_back(i);
    Mat N0((int)rows0.size(), K), N1((int)rows1.size(), K);
    Vec x0(rows0.size()), x1(rows1.size());
    for (size_t a = 0; a < rows0.size(); a++) { x0[a] = c[rows0[a]] * x[rows0[a]]; for (int j = 0; j < K; j++) N0((int)a, j) = c[rows0[a]] * N(rows0[a], j); }
    for (size_t a = 0; a < rows1.size(); a++) { x1[a] = c[rows1[a]] * x[rows1[a]]; for (int j = 0; j < K; j++) N1((int)a, j) = c[rows1[a]] * N(rows1[a], j); }
    Mat N0t = transpose(N0), N1t = transpose(N1);
    Vec ntx0 = matvec(N0t, x0), ntx1 = matvec(N1t, x1);
    Mat ntn0 = matmul(N0t, N0), ntn1 = matmul(N1t, N1);
    if (K == 0) { y.clear(); return true; }
    real rel_error;
    int rank0 = K, guard = 0;
    do {
      FullPivLU lu(ntn0);
      while (lu.rank() > rank0) {
        lu.setThreshold(2 * lu.threshold());
        if (!(lu.threshold() < 1e300)) return false;
      }
      Vec mb(K);
      for (int i = 0; i < K; i++) mb[i] = -ntx0[i];
      Vec y0 = lu.solve(mb);
      Mat Ny = lu.kernel(), Ry = lu.image(ntn0);
      // Eigen's kernel() of an invertible matrix is ONE zero column, so `m << ntn1*Ny, ntn0*Ry` (ftsolver.cpp:222-223)
      // then passes K+1 columns to a K x K comma initialiser and Eigen's assertion aborts the reference (one foot
      // on the ground: level 0 has full rank 3).  oracle/_ref reproduces the abort; report it as a breakdown here.
      if (Ny.c == 0) return false;
      if (rank0 == lu.rank()) solver_warnings++;  // "decomposition threshold increased"
      rank0 = lu.rank();
      Vec t1 = matvec(ntn1, y0), b(K);
      for (int i = 0; i < K; i++) b[i] = -(ntx1[i] + t1[i]);
      Mat m(K, K), a1 = matmul(ntn1, Ny), a2 = matmul(ntn0, Ry);
      for (int i = 0; i < K; i++) {
        for (int j = 0; j < a1.c; j++) m(i, j) = a1(i, j);
        for (int j = 0; j < a2.c; j++) m(i, a1.c + j) = a2(i, j);
      }
      Vec zz = colpiv_qr_solve(m, b);
      Vec res = matvec(m, zz);
      for (int i = 0; i < K; i++) res[i] -= b[i];
      rel_error = norm2(res) / norm2(b);
      rank0--;
      y = y0;
      for (int i = 0; i < K; i++) for (int j = 0; j < Ny.c; j++) y[i] += Ny(i, j) * zz[j];
      if (++guard > K + 2) return false;
    } while (rel_error > 1e-6);
    return true;
  }
};

// modelplayer::prepare_per_traj_dyn + measure_cot (player.cpp:259-285)
struct CotResult { real cot, work, min_cfz, max_mu; int status; };
inline CotResult measure_cot(Model& model, GaitSetup& g, int n_t, real* traj_out = 0, real* x_dump = 0,
                             real* z_dump = 0, real* tau_dump = 0) {
  CotResult r; r.cot = r.work = r.min_cfz = r.max_mu = 0; r.status = 0;
  GaitEvaluator ev(&model);
  if (!ev.record_trajectory(&g, n_t)) { r.status = 1; return r; }  // IK target unreachable
  if (traj_out) for (int i = 0; i < ev.traj_size; i++) for (int j = 0; j < ev.config_dim; j++) traj_out[(size_t)i * ev.config_dim + j] = ev.traj[i][j];
  ev.compute_dynrecs();
  ev.compute_dynrec_ders();
  ev.switch_torso_penalty(true, true);
  if (!ev.work_over_period(r.work, x_dump, z_dump)) { r.status = 2; return r; }
  r.cot = r.work / (ev.total_mass() * g.pattern.step_length);
  r.min_cfz = ev.min_cfz; r.max_mu = ev.max_mu;
  if (tau_dump) for (int i = 2; i < n_t + 2; i++) for (int j = 0; j < ev.nmj; j++) tau_dump[(size_t)(i - 2) * ev.nmj + j] = ev.torques[i % n_t][j];
  return r;
}

}  // namespace orc
