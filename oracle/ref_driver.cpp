// TEST INFRASTRUCTURE ONLY -- C entry points that drive the reference's OWN classes, compiled unmodified from
// /root/reference into oracle/_ref/libhslref.so (oracle/Makefile, target `ref`).  Users: tests/ (pinning the oracle
// restatement and the CUDA path), bench.py's cpu_baseline / --impl reference legs.  Nothing under hslabs_b200/ may
// include, link or call this file.
//
// What runs here is the reference's code: kinematicmodel / liksolver / pergensetup / pgssweeper / periodic /
// dynrecord / forcetorquesolver / modelplayer (matrix.cpp, core.cpp, model.cpp, visualization.cpp, geom.cpp, lik.cpp,
// pergen.cpp, dynrec.cpp, ftsolver.cpp, periodic.cpp, player.cpp, playerexperim.cpp, odestate.cpp, ghost.cpp, cpc.cpp),
// against shim headers for the absent third-party libraries (oracle/shim: ODE, drawstuff, rapidxml, Eigen facade).
// This file only sequences calls the way player.cpp does (measure_cot player.cpp:269-285, prepare_per_traj_dyn
// :259-264, measure_cot_sweep :311-321, make_pergensu :147-166, test_dynamics playerexperim.cpp:95-121) and copies
// results out; it is built with -fno-access-control so it can read private members (dynrecs, traj, ftsolver).
//
// Process isolation: the reference reports errors by printing "ERROR ..." and calling exit(1) (lik.cpp:163,
// ftsolver.cpp:245, ...), and Eigen-style assertions abort.  Every entry that can reach such a path runs in a
// fork()ed child that inherits the loaded model; results come back through an anonymous shared mapping and the
// child's exit status becomes the return code (1 = the reference called exit(1), 2 = it aborted / crashed).
//
// Candidate parameter vector: 13 doubles, same layout as oracle/orc_capi.cpp and include/hsl.h.
#include <signal.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <sstream>
#include <string>
#include <vector>

#include "lik.h"
#include "pergen.h"
#include "player.h"
#include "periodic.h"
#include "dynrec.h"
#include "ftsolver.h"

extern long ds_shim_max_steps;
void dShimSeedRandom(unsigned long s);

namespace {

struct Handle {
  std::string dir, xml;
  modelplayer* mp;
  int n, nf, nmj, cd;
};

class NullBuf : public std::streambuf {
 protected:
  int overflow(int c) { return c == EOF ? 0 : c; }
  std::streamsize xsputn(const char*, std::streamsize n) { return n; }
};
NullBuf g_nullbuf;
int g_models_loaded = 0;  // the reference keeps ONE static ODE world (visualization.cpp:165): simulation needs a process with one model
bool g_quiet = true;   // discard the reference's stdout chatter (warnings, "sweeping over ...")
bool g_isolate = true;

struct Quiet {  // RAII: route std::cout to nowhere (or into a string) while reference code runs
  std::streambuf* old;
  explicit Quiet(std::streambuf* to = 0) : old(0) {
    if (to) old = std::cout.rdbuf(to);
    else if (g_quiet) old = std::cout.rdbuf(&g_nullbuf);
  }
  ~Quiet() { if (old) std::cout.rdbuf(old); }
};

struct Shm {
  void* p;
  size_t bytes;
  explicit Shm(size_t b) : bytes(b ? b : 8) {
    p = mmap(0, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) p = 0;
  }
  ~Shm() { if (p) munmap(p, bytes); }
};

void child_signals() {  // a host process (pytest's faulthandler) may have installed dump handlers; the child dies quietly
  const int sigs[] = {SIGABRT, SIGSEGV, SIGFPE, SIGBUS, SIGILL};
  for (size_t i = 0; i < sizeof sigs / sizeof sigs[0]; i++) signal(sigs[i], SIG_DFL);
}

// Runs body(shared) in a forked child; returns 0, or 1 (reference exit(1)), 2 (abort / signal), -1 (fork failed).
int run_isolated(size_t bytes, const std::function<void(char*)>& body, const std::function<void(const char*)>& collect) {
  if (!g_isolate) {
    std::vector<char> buf(bytes ? bytes : 8, 0);
    body(&buf[0]);
    collect(&buf[0]);
    return 0;
  }
  Shm shm(bytes);
  if (!shm.p) return -1;
  std::memset(shm.p, 0, shm.bytes);
  fflush(0);
  pid_t pid = fork();
  if (pid < 0) return -1;
  if (pid == 0) {
    child_signals();
    body((char*)shm.p);
    _exit(0);
  }
  int st = 0;
  if (waitpid(pid, &st, 0) < 0) return -1;
  if (WIFEXITED(st) && WEXITSTATUS(st) == 0) { collect((const char*)shm.p); return 0; }
  if (WIFEXITED(st)) return 1;
  return 2;
}

void fill_pcp(pgsconfigparams& pcp, const Handle* h, const double* p) {  // what get_pgs_config_params builds (player.cpp:170-208)
  pcp.fname = h->xml;
  pcp.orientation[0].set(p[0], p[1], p[2]);
  pcp.orientation[1].set(p[3], p[4], p[5]);
  pcp.step_duration = p[6];
  pcp.set_TLh(p[7], p[8], p[9]);
  pcp.curvature = p[10];
  pcp.foot_shift = pair<int, double>((int)p[11], p[12]);
}

pergensetup* make_pgs(Handle* h, const double* params) {  // make_pergensu without the file (player.cpp:161-165)
  pgsconfigparams pcp;
  fill_pcp(pcp, h, params);
  int n = h->mp->model->get_lik()->get_number_of_limbs();
  pergensetup* pgs = new pergensetup(n);
  h->mp->setup_pergen(*pgs, pcp);
  return pgs;
}

struct Sizes { size_t traj, x, z, tau; };
Sizes sizes(const Handle* h, int n_t) {
  Sizes s;
  s.traj = (size_t)(n_t + 5) * h->cd;
  s.x = (size_t)n_t * 6 * h->n;
  s.z = (size_t)n_t * 3 * h->nf;
  s.tau = (size_t)n_t * h->nmj;
  return s;
}

// measure_cot (player.cpp:269-285) with every intermediate copied out.  buf: out4 | traj | x | z | tau | complete
void eval_candidate(Handle* h, pergensetup* pgs, int n_t, bool detail, double* buf) {
  modelplayer* mp = h->mp;
  // the body of modelplayer::measure_cot (player.cpp:269-275), kept open so that work and the contact statistics of
  // the same pass can be returned; in detail mode the reference's own measure_cot is called as well
  periodic per(mp->model);
  mp->prepare_per_traj_dyn(per, pgs, n_t);
  double work = per.work_over_period();
  double weight = per.get_total_mass();
  double step_length = pgs->get_pergen()->get_step_length();
  buf[0] = work / (weight * step_length);
  buf[1] = work;
  double stat[2];
  per.get_contforce_stat(stat);
  buf[2] = stat[0];
  buf[3] = stat[1];
  if (!detail) return;
  buf[0] = mp->measure_cot(pgs, n_t);
  Sizes s = sizes(h, n_t);
  double* traj = buf + 4;
  double* x = traj + s.traj;
  double* z = x + s.x;
  double* tau = z + s.z;
  double* complete = tau + s.tau;  // [n_t][2*cd+nmj], the traj.txt record (periodic.cpp:408-426)
  for (int i = 0; i < n_t + 5; i++) std::memcpy(traj + (size_t)i * h->cd, per.traj[i], sizeof(double) * h->cd);
  for (int i = 2; i < n_t + 2; i++) {
    VectorXd xv, zv;
    per.ftsolver->solve_forcetorques(per.dynrecs[i], xv, zv);
    for (int k = 0; k < 6 * h->n; k++) x[(size_t)(i - 2) * 6 * h->n + k] = xv(k);
    for (int k = 0; k < 3 * h->nf; k++) z[(size_t)(i - 2) * 3 * h->nf + k] = zv(k);
    per.get_motor_torques(tau + (size_t)(i - 2) * h->nmj);
  }
  int w = 2 * h->cd + h->nmj;
  double** ct = new_2d_array(n_t, w);
  per.get_complete_traj(ct);
  for (int i = 0; i < n_t; i++) std::memcpy(complete + (size_t)i * w, ct[i], sizeof(double) * w);
  delete_2d_array(ct, n_t);
}

size_t detail_doubles(const Handle* h, int n_t, bool detail) {
  if (!detail) return 4;
  Sizes s = sizes(h, n_t);
  return 4 + s.traj + s.x + s.z + s.tau + (size_t)n_t * (2 * h->cd + h->nmj);
}

int eval_common(Handle* h, const double* params, const double* rec_transl, const double* rec_eas, int n_t, double* out4,
                double* traj, double* x, double* z, double* tau, double* complete) {
  bool detail = traj || x || z || tau || complete;
  size_t nd = detail_doubles(h, n_t, detail);
  return run_isolated(
      nd * sizeof(double),
      [&](char* shm) {
        Quiet q;
        pergensetup* pgs = make_pgs(h, params);
        if (rec_transl) {
          extvec tr(rec_transl[0], rec_transl[1], rec_transl[2]), ea(rec_eas[0], rec_eas[1], rec_eas[2]);
          pgs->set_rec_transform(tr, ea);
        }
        eval_candidate(h, pgs, n_t, detail, (double*)shm);
        delete pgs;
      },
      [&](const char* shm) {
        const double* b = (const double*)shm;
        std::memcpy(out4, b, 4 * sizeof(double));
        if (!detail) return;
        Sizes s = sizes(h, n_t);
        const double* p = b + 4;
        if (traj) std::memcpy(traj, p, s.traj * sizeof(double));
        p += s.traj;
        if (x) std::memcpy(x, p, s.x * sizeof(double));
        p += s.x;
        if (z) std::memcpy(z, p, s.z * sizeof(double));
        p += s.z;
        if (tau) std::memcpy(tau, p, s.tau * sizeof(double));
        p += s.tau;
        if (complete) std::memcpy(complete, p, (size_t)n_t * (2 * h->cd + h->nmj) * sizeof(double));
      });
}

}  // namespace

extern "C" {

void ref_set_quiet(int q) { g_quiet = (q != 0); }
void ref_set_isolation(int on) { g_isolate = (on != 0); }

// dir: directory that holds the model XML and (optionally) pgs_config.txt; xml: bare file name, because the reference
// selects its limb solver by comparing the file name string (lik.cpp:8-11).  The working directory is switched
// to `dir` for the load only.
void* ref_model_load(const char* dir, const char* xml) {
  char cwd[4096];
  if (!getcwd(cwd, sizeof cwd)) return 0;
  if (chdir(dir) != 0) return 0;
  Handle* h = 0;
  {
    Quiet q;
    if (FILE* f = std::fopen(xml, "r")) {
      std::fclose(f);
      h = new Handle;
      h->dir = dir;
      h->xml = xml;
      h->mp = new modelplayer;
      h->mp->load_model(xml);
      g_models_loaded++;
      h->n = (int)h->mp->model->get_odeparts()->size();
      h->nf = h->mp->model->get_lik()->get_number_of_limbs();
      h->nmj = h->mp->model->number_of_motor_joints();
      h->cd = h->mp->model->get_config_dim();
    }
  }
  if (chdir(cwd) != 0) { /* nothing sensible to do */ }
  return h;
}
void ref_model_free(void* hv) {
  Handle* h = (Handle*)hv;
  if (!h) return;
  // the reference's visualizer is a process-wide static that keeps pointers into every loaded model
  // (visualization.cpp:165,305); models are therefore left alive for the life of the process.
  delete h;
}
void ref_model_dims(void* hv, int* out) {
  Handle* h = (Handle*)hv;
  out[0] = h->n; out[1] = h->nf; out[2] = h->nmj; out[3] = h->cd;
}
double ref_model_rcap(void* hv) { return ((Handle*)hv)->mp->model->get_lik()->get_rcap(); }
void ref_set_ignore_reach(void* hv, int flag) { ((Handle*)hv)->mp->model->get_lik()->set_ignore_reach_flag(flag != 0); }

// Load-time constants as the reference's own loader produced them (same arrays as orc_model_constants).
void ref_model_constants(void* hv, int* parent, int* jkind, double* A_pj_body, double* J_A_parent, double* A_body_geom,
                         double* capsule_to_pos, int* limb_top, int* limb_foot) {
  Handle* h = (Handle*)hv;
  kinematicmodel* m = h->mp->model;
  std::map<const modelnode*, int> ids;
  for (int i = 0; i < h->n; i++) ids[m->get_mnode(i)] = i;
  for (int i = 0; i < h->n; i++) {
    const modelnode* node = m->get_mnode(i);
    parent[i] = node->get_parent() ? ids[node->get_parent()] : -1;
    modeljoint* j = node->get_joint();
    jkind[i] = j ? (j->get_type() == free6 ? 2 : 1) : 0;
    std::memcpy(A_pj_body + 16 * i, const_cast<modelnode*>(node)->get_A_pj_body()->get_data(), 16 * sizeof(double));
    if (j) std::memcpy(J_A_parent + 16 * i, j->get_A_parent()->get_data(), 16 * sizeof(double));
    else std::memset(J_A_parent + 16 * i, 0, 16 * sizeof(double));
    const odepart* op = m->get_odepart(i);
    std::memcpy(A_body_geom + 16 * i, op->A_body_geom.get_data(), 16 * sizeof(double));
    for (int k = 0; k < 3; k++) capsule_to_pos[3 * i + k] = op->capsule_to_pos.get_v(k);
  }
  const vector<liklimb*>* limbs = m->get_lik()->get_limbs();
  for (int l = 0; l < h->nf; l++) {
    limb_top[l] = ids[(*limbs)[l]->child];
    limb_foot[l] = ids[(*limbs)[l]->get_foot()];
  }
}

// FK (a5): joint values -> A_ground of every body and of every joint frame (column-major 4x4 each).
void ref_fk(void* hv, const double* q, double* A_ground, double* J_A_ground) {
  Handle* h = (Handle*)hv;
  kinematicmodel* m = h->mp->model;
  m->set_jvalues(q);
  m->recompute_modelnodes();
  for (int i = 0; i < h->n; i++) {
    const modelnode* node = m->get_mnode(i);
    std::memcpy(A_ground + 16 * i, node->get_A_ground()->get_data(), 16 * sizeof(double));
    modeljoint* j = node->get_joint();
    if (j) std::memcpy(J_A_ground + 16 * i, j->get_A_ground()->get_data(), 16 * sizeof(double));
    else std::memset(J_A_ground + 16 * i, 0, 16 * sizeof(double));
  }
}

// IK (a3): rec[6+3nf] -> joint values.  1 = the reference exit(1)ed (unreachable target).
int ref_ik(void* hv, const double* rec, double* q) {
  Handle* h = (Handle*)hv;
  return run_isolated(
      h->cd * sizeof(double),
      [&](char* shm) {
        Quiet qt;
        h->mp->model->set_jvalues_with_lik(rec);
        h->mp->model->get_jvalues((double*)shm);
      },
      [&](const char* shm) { std::memcpy(q, shm, h->cd * sizeof(double)); });
}

// Candidate construction (a1): pos0 in pergen order, lift-off tables, (t_step, v, max_radius).
int ref_gait_setup(void* hv, const double* params, double* pos0, double* ts, double* xs, double* scal) {
  Handle* h = (Handle*)hv;
  int nf = h->nf;
  return run_isolated(
      (5 * nf + 3) * sizeof(double),
      [&](char* shm) {
        Quiet q;
        double* b = (double*)shm;
        pergensetup* pgs = make_pgs(h, params);
        periodicgenerator* pg = pgs->get_pergen();
        for (int i = 0; i < nf; i++) {
          for (int k = 0; k < 3; k++) b[3 * i + k] = pg->limb_pos0s[i].get_v(k);
          b[3 * nf + i] = pg->ts[i];
          b[4 * nf + i] = pg->xs[i];
        }
        b[5 * nf] = pg->t_step;
        b[5 * nf + 1] = pgs->v;
        b[5 * nf + 2] = pg->max_radius;
        delete pgs;
      },
      [&](const char* shm) {
        const double* b = (const double*)shm;
        std::memcpy(pos0, b, 3 * nf * sizeof(double));
        std::memcpy(ts, b + 3 * nf, nf * sizeof(double));
        std::memcpy(xs, b + 4 * nf, nf * sizeof(double));
        std::memcpy(scal, b + 5 * nf, 3 * sizeof(double));
      });
}

// Frame record (a2) at time t.
int ref_gait_rec(void* hv, const double* params, double t, double* rec) {
  Handle* h = (Handle*)hv;
  int w = 6 + 3 * h->nf;
  return run_isolated(
      w * sizeof(double),
      [&](char* shm) {
        Quiet q;
        pergensetup* pgs = make_pgs(h, params);
        pgs->set_rec((double*)shm, t);
        delete pgs;
      },
      [&](const char* shm) { std::memcpy(rec, shm, w * sizeof(double)); });
}

// measure_cot for one candidate; out4 = cot, work, min_cfz, max_mu; optional dumps as in orc_measure_cot plus
// complete[n_t][2*config_dim+nmj] = periodic::get_complete_traj (the traj.txt rows).
int ref_measure_cot(void* hv, const double* params, int n_t, double* out4, double* traj, double* x, double* z, double* tau,
                    double* complete) {
  return eval_common((Handle*)hv, params, 0, 0, n_t, out4, traj, x, z, tau, complete);
}
int ref_measure_cot_rect(void* hv, const double* params, const double* rec_transl, const double* rec_eas, int n_t, double* out4,
                         double* traj, double* x, double* z, double* tau) {
  return eval_common((Handle*)hv, params, rec_transl, rec_eas, n_t, out4, traj, x, z, tau, 0);
}

// Per-frame dynrecord fields of the solved frames 2..n_t+1 (layout as orc_frame_fields).
int ref_frame_fields(void* hv, const double* params, int n_t, double* pos, double* jpos, double* jzaxis, double* mom_rate,
                     double* ang_mom_rate, double* fpos, unsigned char* contacts) {
  Handle* h = (Handle*)hv;
  const int n = h->n, nf = h->nf;
  size_t per_body = (size_t)n_t * n * 3, per_foot = (size_t)n_t * nf * 3;
  size_t nd = 5 * per_body + per_foot + (size_t)n_t * nf;  // contacts stored as doubles in the transfer buffer
  return run_isolated(
      nd * sizeof(double),
      [&](char* shm) {
        Quiet q;
        double* b = (double*)shm;
        pergensetup* pgs = make_pgs(h, params);
        periodic per(h->mp->model);
        h->mp->prepare_per_traj_dyn(per, pgs, n_t);
        for (int t = 0; t < n_t; t++) {
          dynrecord* r = per.dynrecs[t + 2];
          for (int i = 0; i < n; i++)
            for (int k = 0; k < 3; k++) {
              size_t o = ((size_t)t * n + i) * 3 + k;
              b[o] = r->pos[i].get_v(k);
              b[per_body + o] = r->jpos[i].get_v(k);
              b[2 * per_body + o] = r->jzaxis[i].get_v(k);
              b[3 * per_body + o] = r->mom_rate[i].get_v(k);
              b[4 * per_body + o] = r->ang_mom_rate[i].get_v(k);
            }
          for (int fi = 0; fi < nf; fi++) {
            for (int k = 0; k < 3; k++) b[5 * per_body + ((size_t)t * nf + fi) * 3 + k] = r->fpos[fi].get_v(k);
            b[5 * per_body + per_foot + (size_t)t * nf + fi] = r->contacts[fi] ? 1.0 : 0.0;
          }
        }
        delete pgs;
      },
      [&](const char* shm) {
        const double* b = (const double*)shm;
        std::memcpy(pos, b, per_body * sizeof(double));
        std::memcpy(jpos, b + per_body, per_body * sizeof(double));
        std::memcpy(jzaxis, b + 2 * per_body, per_body * sizeof(double));
        std::memcpy(mom_rate, b + 3 * per_body, per_body * sizeof(double));
        std::memcpy(ang_mom_rate, b + 4 * per_body, per_body * sizeof(double));
        std::memcpy(fpos, b + 5 * per_body, per_foot * sizeof(double));
        for (size_t i = 0; i < (size_t)n_t * nf; i++) contacts[i] = (unsigned char)(b[5 * per_body + per_foot + i] != 0);
      });
}

// Evaluate an externally supplied joint trajectory q[(n_t+5)][config_dim] with time step dt through
// periodic::compute_dynrecs / compute_dynrec_ders / work_over_period (the state record_trajectory would leave
// behind, periodic.cpp:77-96, is written directly).
int ref_eval_trajectory(void* hv, const double* q, int n_t, double dt, double* out3, double* x, double* z, double* tau) {
  Handle* h = (Handle*)hv;
  Sizes s = sizes(h, n_t);
  size_t nd = 3 + s.x + s.z + s.tau;
  return run_isolated(
      nd * sizeof(double),
      [&](char* shm) {
        Quiet qt;
        double* b = (double*)shm;
        periodic per(h->mp->model);
        per.n_t = n_t;
        per.traj_size = n_t + 5;
        per.config_dim = h->cd;
        per.nmj = h->nmj;
        per.traj = new_2d_array(per.traj_size, h->cd);
        for (int i = 0; i < per.traj_size; i++) std::memcpy(per.traj[i], q + (size_t)i * h->cd, sizeof(double) * h->cd);
        per.dt_traj = dt;
        per.rcap = h->mp->model->get_lik()->get_rcap();
        per.compute_dynrecs();
        per.compute_dynrec_ders();
        per.switch_torso_penalty(1, 1);
        b[0] = per.work_over_period();
        double stat[2];
        per.get_contforce_stat(stat);
        b[1] = stat[0];
        b[2] = stat[1];
        double *xo = b + 3, *zo = xo + s.x, *to = zo + s.z;
        for (int i = 2; i < n_t + 2; i++) {
          VectorXd xv, zv;
          per.ftsolver->solve_forcetorques(per.dynrecs[i], xv, zv);
          for (int k = 0; k < 6 * h->n; k++) xo[(size_t)(i - 2) * 6 * h->n + k] = xv(k);
          for (int k = 0; k < 3 * h->nf; k++) zo[(size_t)(i - 2) * 3 * h->nf + k] = zv(k);
          per.get_motor_torques(to + (size_t)(i - 2) * h->nmj);
        }
      },
      [&](const char* shm) {
        const double* b = (const double*)shm;
        std::memcpy(out3, b, 3 * sizeof(double));
        if (x) std::memcpy(x, b + 3, s.x * sizeof(double));
        if (z) std::memcpy(z, b + 3 + s.x, s.z * sizeof(double));
        if (tau) std::memcpy(tau, b + 3 + s.x + s.z, s.tau * sizeof(double));
      });
}

// The sequence of modelplayer::test_dynamics (playerexperim.cpp:95-121) at an arbitrary n_t / frame.
int ref_test_dynamics(void* hv, const double* params, int n_t, int frame, double* cf, double* cf1, double* tau) {
  Handle* h = (Handle*)hv;
  size_t nd = 6 * h->nf + h->nmj;
  return run_isolated(
      nd * sizeof(double),
      [&](char* shm) {
        Quiet q;
        double* b = (double*)shm;
        pergensetup* pgs = make_pgs(h, params);
        periodic per(h->mp->model);
        h->mp->prepare_per_traj_dyn(per, pgs, n_t);
        per.solve_torques_contforces(frame, b + 6 * h->nf, b);
        per.solve_contforces_given_torques(frame, b + 3 * h->nf, b + 6 * h->nf);
        delete pgs;
      },
      [&](const char* shm) {
        const double* b = (const double*)shm;
        std::memcpy(cf, b, 3 * h->nf * sizeof(double));
        std::memcpy(cf1, b + 3 * h->nf, 3 * h->nf * sizeof(double));
        std::memcpy(tau, b + 6 * h->nf, h->nmj * sizeof(double));
      });
}

// periodic::solve_contforces_given_torques on every solved frame: tau [n_t][nmj] -> cf [n_t][3nf].
int ref_solve_forces_frames(void* hv, const double* params, int n_t, const double* tau, double* cf) {
  Handle* h = (Handle*)hv;
  size_t nd = (size_t)n_t * 3 * h->nf;
  return run_isolated(
      nd * sizeof(double),
      [&](char* shm) {
        Quiet q;
        double* b = (double*)shm;
        pergensetup* pgs = make_pgs(h, params);
        periodic per(h->mp->model);
        h->mp->prepare_per_traj_dyn(per, pgs, n_t);
        std::vector<double> t(h->nmj);
        for (int i = 0; i < n_t; i++) {
          std::memcpy(&t[0], tau + (size_t)i * h->nmj, h->nmj * sizeof(double));
          per.solve_contforces_given_torques(i + 2, b + (size_t)i * 3 * h->nf, &t[0]);
        }
        delete pgs;
      },
      [&](const char* shm) { std::memcpy(cf, shm, nd * sizeof(double)); });
}

// modelplayer::measure_cot_sweep itself (player.cpp:311-321): its only output channel is stdout
// ("val = <v> COT = <c>"), which is captured and parsed.  vals / cots: [n_val+1].
int ref_measure_cot_sweep(void* hv, const double* params, int n_t, const char* name, double v0, double v1, int n_val,
                          double* vals, double* cots) {
  Handle* h = (Handle*)hv;
  size_t nd = 2 * (size_t)(n_val + 1) + 1;
  return run_isolated(
      nd * sizeof(double),
      [&](char* shm) {
        double* b = (double*)shm;
        std::stringbuf cap;
        pergensetup* pgs;
        {
          Quiet q;
          pgs = make_pgs(h, params);
        }
        {
          Quiet q(&cap);
          std::cout.precision(17);
          h->mp->measure_cot_sweep(pgs, n_t, name, v0, v1, n_val);
        }
        std::istringstream is(cap.str());
        std::string tok;
        int i = 0;
        while (is >> tok) {
          if (tok != "val") continue;
          std::string eq, cot_s, eq2;
          double v, c;
          if (!(is >> eq >> v >> cot_s >> eq2)) break;
          std::string cs;
          is >> cs;
          c = (cs == "nan" || cs == "-nan") ? NAN : std::atof(cs.c_str());
          if (i <= n_val) { b[i] = v; b[n_val + 1 + i] = c; }
          i++;
        }
        b[2 * (n_val + 1)] = i;
        delete pgs;
      },
      [&](const char* shm) {
        const double* b = (const double*)shm;
        std::memcpy(vals, b, (n_val + 1) * sizeof(double));
        std::memcpy(cots, b + n_val + 1, (n_val + 1) * sizeof(double));
      });
}

// Batch of candidates params[C][13] on `nprocs` forked workers (the reference is single-threaded and mutates
// global state; processes keep it that way).  A worker that dies on a candidate (the reference's exit(1)) is
// replaced by one that resumes after it.  status: 0 ok, 1 reference exit(1), 2 abort / crash.
int ref_eval_batch(void* hv, long n_cand, int n_t, const double* params, double* cot, double* work, double* min_cfz,
                   double* max_mu, int* status, int nprocs) {
  Handle* h = (Handle*)hv;
  if (nprocs < 1) nprocs = 1;
  if (n_cand <= 0) return 0;
  Shm shm((size_t)n_cand * 5 * sizeof(double));
  if (!shm.p) return -1;
  double* b = (double*)shm.p;  // [C][5]: cot, work, min_cfz, max_mu, done-flag
  for (long c = 0; c < n_cand; c++) { b[5 * c] = b[5 * c + 1] = b[5 * c + 2] = b[5 * c + 3] = NAN; b[5 * c + 4] = 0; }
  std::vector<long> next(nprocs);
  std::vector<pid_t> pids(nprocs, -1);
  for (int w = 0; w < nprocs; w++) next[w] = w;
  std::vector<int> st(n_cand, 0);
  fflush(0);
  int live = 0;
  auto spawn = [&](int w) {
    if (next[w] >= n_cand) { pids[w] = -1; return; }
    pid_t pid = fork();
    if (pid == 0) {
      child_signals();
      Quiet q;
      for (long c = next[w]; c < n_cand; c += nprocs) {
        b[5 * c + 4] = 1;  // started
        pergensetup* pgs = make_pgs(h, params + 13 * c);
        eval_candidate(h, pgs, n_t, false, b + 5 * c);
        delete pgs;
        b[5 * c + 4] = 2;  // finished
      }
      _exit(0);
    }
    pids[w] = pid;
    if (pid > 0) live++;
  };
  for (int w = 0; w < nprocs; w++) spawn(w);
  while (live > 0) {
    int s = 0;
    pid_t pid = wait(&s);
    if (pid < 0) break;
    int w = -1;
    for (int i = 0; i < nprocs; i++) if (pids[i] == pid) w = i;
    if (w < 0) continue;
    live--;
    if (WIFEXITED(s) && WEXITSTATUS(s) == 0) { pids[w] = -1; continue; }
    long c = next[w];  // find the candidate it died on
    while (c < n_cand && b[5 * c + 4] == 2) c += nprocs;
    if (c < n_cand) {
      st[c] = WIFEXITED(s) ? 1 : 2;
      b[5 * c] = b[5 * c + 1] = b[5 * c + 2] = b[5 * c + 3] = NAN;
      next[w] = c + nprocs;
      spawn(w);
    } else {
      pids[w] = -1;
    }
  }
  for (long c = 0; c < n_cand; c++) {
    cot[c] = b[5 * c]; work[c] = b[5 * c + 1]; min_cfz[c] = b[5 * c + 2]; max_mu[c] = b[5 * c + 3];
    status[c] = st[c];
  }
  return 0;
}

// Preset row through the reference's own parser (get_rec_str + get_pgs_config_params, player.cpp:170-208,230-244).
// 0 ok, 1 = the reference exit(1)ed (no such id / unknown key).
int ref_load_preset(void* hv, const char* file, int id, double* params, char* xml_name) {
  Handle* h = (Handle*)hv;
  return run_isolated(
      13 * sizeof(double) + 64,
      [&](char* shm) {
        Quiet q;
        double* p = (double*)shm;
        string rec;
        h->mp->get_rec_str(rec, file, id);
        pgsconfigparams pcp;
        h->mp->get_pgs_config_params(rec, pcp);
        for (int k = 0; k < 3; k++) { p[k] = pcp.orientation[0].get_v(k); p[3 + k] = pcp.orientation[1].get_v(k); }
        p[6] = pcp.step_duration; p[7] = pcp.TLh[0]; p[8] = pcp.TLh[1]; p[9] = pcp.TLh[2];
        p[10] = pcp.curvature; p[11] = pcp.foot_shift.first; p[12] = pcp.foot_shift.second;
        std::strncpy(shm + 13 * sizeof(double), pcp.fname.c_str(), 63);
      },
      [&](const char* shm) {
        std::memcpy(params, shm, 13 * sizeof(double));
        std::memcpy(xml_name, shm + 13 * sizeof(double), 64);
        xml_name[63] = 0;
      });
}

// Fall / perturbation run (BASELINE configs[4]): the reference's own closed loop
//   position_control_test -> setup_per_controller -> step() -> simulate_ode   (player.cpp:358-382, 326-340, 393-432)
// on the ODE shim's world stepper (oracle/shim/ode_step.cpp), driven step by step instead of through drawstuff's loop
// so that the kick (kick_torso, player.cpp:585-605: a force dv/dt on the torso for one step) and the fall check
// (fall_check, player.cpp:669-681: torso z < hc once play_t >= tmin) take their parameters from the caller.
}  // extern "C" (reopened below)

// params[13] as everywhere; kick_dv[3] = velocity change of the kick applied at step kick_step (< 0: never).
// out[4] = fell (0/1), time of the fall (or end time), final torso z, steps simulated; traj (optional)
// [n_steps][3] torso position after every step.  Needs a process in which only this model was loaded.
namespace {
// position_control_test up to the first step (player.cpp:358-360, 370-382): evaluates the gait, poses the robot at t0
void fall_setup(Handle* h, const double* params, double play_dt, double t0) {
  modelplayer* mp = h->mp;
  mp->set_play_dt(play_dt);
  pergensetup* pgs = make_pgs(h, params);
  mp->ignore_reach();                          // set_fall_test does (player.cpp:785)
  mp->set_flag("position_control", true);
  mp->setup_per_controller(pgs, t0);
}
// the loop of one world from that state; b[4 + 3 n_steps]
void fall_loop(Handle* h, double play_dt, int n_steps, int kick_step, const double* kick_dv, double hc, double tmin, double* b, bool want_traj) {
  modelplayer* mp = h->mp;
  dShimSeedRandom(0);                           // a fresh ODE process
  dBodyID torso = mp->get_torso_odebody();
  int fell = 0, step = 0;
  double t_fall = 0;
  for (; step < n_steps; step++) {
    const dReal* pos = dBodyGetPosition(torso);
    if (mp->play_t >= tmin && pos[2] < hc) { fell = 1; t_fall = mp->play_t; break; }   // fall_check
    if (step == kick_step) {
      double f[3] = {kick_dv[0] / play_dt, kick_dv[1] / play_dt, kick_dv[2] / play_dt};
      mp->get_vis()->add_force(torso, f);
    }
    mp->step();                                 // case 6: simulate_ode()
    if (want_traj) {
      pos = dBodyGetPosition(torso);
      for (int k = 0; k < 3; k++) b[4 + 3 * step + k] = pos[k];
    }
  }
  b[0] = fell;
  b[1] = fell ? t_fall : mp->play_t;
  b[2] = dBodyGetPosition(torso)[2];
  b[3] = step;
}
}  // namespace

extern "C" {
int ref_fall_run(void* hv, const double* params, double play_dt, double t0, int n_steps, int kick_step, const double* kick_dv, double hc,
                 double tmin, double* out, double* traj) {
  Handle* h = (Handle*)hv;
  if (g_models_loaded != 1) return -3;
  const size_t nd = 4 + 3 * (size_t)n_steps;
  return run_isolated(
      nd * sizeof(double),
      [&](char* shm) {
        Quiet q;
        fall_setup(h, params, play_dt, t0);
        fall_loop(h, play_dt, n_steps, kick_step, kick_dv, hc, tmin, (double*)shm, true);
      },
      [&](const char* shm) {
        const double* b = (const double*)shm;
        std::memcpy(out, b, 4 * sizeof(double));
        if (traj) std::memcpy(traj, b + 4, 3 * (size_t)n_steps * sizeof(double));
      });
}

// W fall runs on `nprocs` forked workers.  A worker evaluates the gait and poses the robot once (the part of
// position_control_test every world shares), then forks per world: each world steps its own copy of that pristine state,
// as the separate runs of run_fall_test.sh's loop would.  fell / t_end / final_z: [W].
int ref_fall_batch(void* hv, const double* params, long W, double play_dt, double t0, int n_steps, const int* kick_step, const double* kick_dv,
                   double hc, double tmin, unsigned char* fell, double* t_end, double* final_z, int nprocs) {
  Handle* h = (Handle*)hv;
  if (g_models_loaded != 1) return -3;
  if (nprocs < 1) nprocs = 1;
  Shm shm((size_t)W * 4 * sizeof(double));
  if (!shm.p) return -1;
  double* b = (double*)shm.p;
  for (long i = 0; i < 4 * W; i++) b[i] = NAN;
  fflush(0);
  std::vector<pid_t> pids;
  for (int w = 0; w < nprocs; w++) {
    pid_t pid = fork();
    if (pid == 0) {
      child_signals();
      Quiet q;
      fall_setup(h, params, play_dt, t0);
      const double zero[3] = {0, 0, 0};
      for (long i = w; i < W; i += nprocs) {
        pid_t p2 = fork();
        if (p2 == 0) {
          double out[4];
          fall_loop(h, play_dt, n_steps, kick_step ? kick_step[i] : -1, kick_dv ? kick_dv + 3 * i : zero, hc, tmin, out, false);
          for (int k = 0; k < 4; k++) b[4 * i + k] = out[k];
          _exit(0);
        }
        int st = 0;
        if (p2 > 0) waitpid(p2, &st, 0);
      }
      _exit(0);
    }
    if (pid > 0) pids.push_back(pid);
  }
  for (size_t i = 0; i < pids.size(); i++) { int st = 0; waitpid(pids[i], &st, 0); }
  for (long i = 0; i < W; i++) { fell[i] = (unsigned char)(b[4 * i] == 1.0); t_end[i] = b[4 * i + 1]; final_z[i] = b[4 * i + 2]; }
  return 0;
}

// liksolver::solver_test (lik.cpp:123-128): the reference's own IK round-trip self-check; exit(1) on failure.
int ref_lik_solver_test(void* hv, int n) {
  Handle* h = (Handle*)hv;
  return run_isolated(8, [&](char*) { Quiet q; h->mp->model->get_lik()->solver_test(n); }, [](const char*) {});
}

}  // extern "C"
