// hsl_host.hpp -- C++ host classes that keep the reference's call signatures for the gait-evaluation path and
// forward to the C ABI of include/hsl.h (header only; link with -lhsl_b200).
//
// The reference classes these mirror: kinematicmodel (model.h:100-137), pgsconfigparams / pergensetup / pgssweeper
// (pergen.h:68-146), periodic (periodic.h:27-87), modelplayer's evaluation entry points (player.h:71-74).  The one
// deliberate signature change is the absence of Eigen types (Eigen is not available): vectors are double* /
// std::vector<double>.  Where the reference prints "ERROR ..." and exit(1)s, these classes throw hsl::error.
//
// Evaluation is lazy: periodic::record_trajectory only records the candidate; the GPU is invoked once by the first
// call that needs results (work_over_period, compute_torques_over_period, solve_torques_contforces, ...).
#pragma once
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "hsl.h"

namespace hsl {

struct error : std::runtime_error {
  explicit error(const std::string& s) : std::runtime_error(s) {}
};
inline void check(int rc) {
  if (rc != HSL_OK) throw error(std::string("hsl: ") + hsl_last_error());
}

// pgsconfigparams (pergen.h:137-146)
struct pgsconfigparams {
  std::string fname;
  double orientation[2][3];
  double step_duration;
  double TLh[3];
  double curvature;
  std::pair<int, double> foot_shift;  // 0 lateral shift, 1 radial shift, -1 none
  pgsconfigparams() : step_duration(0), curvature(0), foot_shift(-1, 0.0) {
    for (int i = 0; i < 2; i++) for (int k = 0; k < 3; k++) orientation[i][k] = 0;
    TLh[0] = TLh[1] = TLh[2] = 0;
  }
  void set_TLh(double period, double step_length, double step_height) { TLh[0] = period; TLh[1] = step_length; TLh[2] = step_height; }
  void to_row(double* p) const {  // the 13-scalar candidate row of hsl.h
    for (int k = 0; k < 3; k++) { p[k] = orientation[0][k]; p[3 + k] = orientation[1][k]; }
    p[6] = step_duration; p[7] = TLh[0]; p[8] = TLh[1]; p[9] = TLh[2]; p[10] = curvature;
    p[11] = foot_shift.first; p[12] = foot_shift.second;
  }
};

// kinematicmodel (model.h:100-137): only what the evaluation path needs
class kinematicmodel {
  HslModel* h_;
  std::string xmlfname_;
  int32_t dims_[6];
 public:
  kinematicmodel() : h_(nullptr) { for (int i = 0; i < 6; i++) dims_[i] = 0; }
  ~kinematicmodel() { if (h_) hsl_model_free(h_); }
  kinematicmodel(const kinematicmodel&) = delete;
  kinematicmodel& operator=(const kinematicmodel&) = delete;
  void load_fromxml(const std::string& fname) {
    if (h_) { hsl_model_free(h_); h_ = nullptr; }
    check(hsl_model_load_xml(fname.c_str(), &h_));
    check(hsl_model_dims(h_, dims_));
    xmlfname_ = fname;
  }
  bool if_loaded() const { return h_ != nullptr; }
  std::string get_xmlfname() const { return xmlfname_; }
  int get_config_dim() const { return dims_[3]; }
  int number_of_motor_joints() const { return dims_[2]; }
  int number_of_parts() const { return dims_[0]; }
  int number_of_limbs() const { return dims_[1]; }
  double get_rcap() const { return hsl_model_rcap(h_); }  // liksolver::get_rcap
  HslModel* handle() const { return h_; }
  // model.cpp:354-372: joint values from a record (closed-form limb IK on the GPU), plain setter / getter
  void set_jvalues_with_lik(const double* rec) {
    jvalues_.assign(dims_[3], 0.0);
    int32_t status = 0;
    check(hsl_ik_records_host(h_, 1, rec, ignore_reach_ ? HSL_FLAG_IGNORE_REACH : 0, jvalues_.data(), &status));
    if (status & HSL_ST_UNREACHABLE) throw error("LIK ERROR: limb position is unreachable");  // lik.cpp:161-164
  }
  void set_jvalues(const double* values) { jvalues_.assign(values, values + dims_[3]); }
  void get_jvalues(double* values) const { for (size_t i = 0; i < jvalues_.size(); i++) values[i] = jvalues_[i]; }
  void set_ignore_reach_flag(bool v) { ignore_reach_ = v; }  // liksolver::set_ignore_reach_flag, lik.cpp:142-146
 private:
  std::vector<double> jvalues_;
  bool ignore_reach_ = false;
};

// periodicgenerator accessors used by callers (pergen.h:38-41)
class periodicgenerator_view {
  const pgsconfigparams* p_;
 public:
  explicit periodicgenerator_view(const pgsconfigparams* p) : p_(p) {}
  double get_period() const { return p_->TLh[0]; }
  double get_step_length() const { return p_->TLh[1]; }
  double get_step_duration() const { return p_->step_duration; }
  double get_curvature() const { return p_->curvature; }
};

// pergensetup (pergen.h:68-112): a candidate gait
// rec_transform of a pergensetup (pergen.h:75-76): Euler angles + translation of the rigid map, and its flag
struct rectransform {
  bool flag;
  double transl[3], eas[3];
  rectransform() : flag(false) { for (int k = 0; k < 3; k++) transl[k] = eas[k] = 0; }
  // hands the map to the library handle (hsl_set_rec_transform); every evaluation call of the mirrors does this first
  void apply(HslModel* h) const { check(flag ? hsl_set_rec_transform(h, transl, eas) : hsl_set_rec_transform(h, nullptr, nullptr)); }
};

class pergensetup {
  int n_;
  pgsconfigparams pcp_;
  periodicgenerator_view view_;
  rectransform rec_;
  const kinematicmodel* model_;  // set by make_pergensu / pgssweeper: the reference bakes the model's default foot
                                 // positions into the pergensetup at that point (pergen.cpp:453-507)
 public:
  explicit pergensetup(int n, const kinematicmodel* model = nullptr) : n_(n), view_(&pcp_), model_(model) {}
  pergensetup(const pergensetup& o) : n_(o.n_), pcp_(o.pcp_), view_(&pcp_), rec_(o.rec_), model_(o.model_) {}
  const kinematicmodel* get_model() const { return model_; }
  // pergen.cpp:225-239: frame record (torso position, Euler angles, foot targets in LIK order) at time t
  void set_rec(double* rec, double t) const {
    if (!model_) throw error("pergensetup::set_rec needs the model the setup was made for");
    double row[HSL_NPARAM];
    pcp_.to_row(row);
    int32_t status = 0;
    rec_.apply(model_->handle());
    check(hsl_gait_records_host(model_->handle(), 1, row, 1, &t, 0, rec, &status));
    if (status & HSL_ST_BAD_PARAMS) throw error("ERROR: step_duration out of bounds");
  }
  // pergen.cpp:309-313: keeps the translation, replaces the rotation (eas = phi, theta, psi)
  void set_rec_rotation(const double rec_eas[3]) { for (int k = 0; k < 3; k++) rec_.eas[k] = rec_eas[k]; rec_.flag = true; }
  void copy_rec_transform(const pergensetup* pgs) { rec_ = pgs->rec_; }  // pergen.cpp:338-342
  const rectransform& rec_transform() const { return rec_; }
  const periodicgenerator_view* get_pergen() const { return &view_; }
  int get_limb_number() const { return n_; }
  int get_config_dim() const { return 6 + 3 * n_; }
  double get_period() const { return pcp_.TLh[0]; }
  void set_TLh(double T, double L, double h) { pcp_.set_TLh(T, L, h); }
  void set_TLh(const double TLh[3]) { set_TLh(TLh[0], TLh[1], TLh[2]); }
  void set_curvature(double c) { pcp_.curvature = c; }
  void set_foot_shift(const std::pair<int, double>& fs) { pcp_.foot_shift = fs; }
  void set_config_params(const pgsconfigparams& pcp) { pcp_ = pcp; }
  void get_config_params(pgsconfigparams* pcp) const { *pcp = pcp_; }
  const pgsconfigparams& params() const { return pcp_; }
};

// pgssweeper (pergen.h:116-134, pergen.cpp:417-449)
class pgssweeper {
  const pergensetup* pgs0_;
  pergensetup* pgs_;
  int parami_, n_val_, vali_;
  double val0_, delval_, val_;
 public:
  pgssweeper(const pergensetup* pgs, const kinematicmodel*) : pgs0_(pgs), pgs_(nullptr), parami_(-1), n_val_(0), vali_(0), val0_(0), delval_(0), val_(0) {}
  ~pgssweeper() { delete pgs_; }
  pergensetup* get_pgs() const { return pgs_; }
  double get_val() const { return val_; }
  void sweep(const std::string& param_name, double val0, double val1, int n_val) {
    val0_ = val0; n_val_ = n_val; delval_ = (val1 - val0) / n_val; vali_ = 0;
    static const char* names[] = {"step_duration", "period", "step_length", "step_height"};
    parami_ = -1;
    for (int i = 0; i < 4; i++) if (param_name == names[i]) parami_ = i;
    if (parami_ < 0) throw error("ERROR: cannot sweep over " + param_name);
    std::cout << "sweeping over " << param_name << ":" << std::endl;
  }
  bool next() {
    if (vali_ > n_val_) { vali_ = 0; return false; }
    val_ = val0_ + vali_ * delval_;
    vali_++;
    delete pgs_;
    pgsconfigparams pcp;
    pgs0_->get_config_params(&pcp);
    if (parami_ == 0) pcp.step_duration = val_; else pcp.TLh[parami_ - 1] = val_;
    pgs_ = new pergensetup(pgs0_->get_limb_number(), pgs0_->get_model());
    pgs_->set_config_params(pcp);
    pgs_->copy_rec_transform(pgs0_);  // pergen.cpp:446
    return true;
  }
  int number_of_values() const { return n_val_ + 1; }
};

// dynrecord (dynrec.h:76-106): the per-frame quantities the force-torque system is assembled from, as plain arrays
// ([body][3] for pos, jpos, jzaxis, mom_rate, ang_mom_rate; [foot][3] for fpos; [foot] for contacts).
struct dynrecord {
  int n, nf;
  std::vector<double> pos, jpos, jzaxis, mom_rate, ang_mom_rate, fpos;
  std::vector<uint8_t> contacts;
  dynrecord(int n_, int nf_) : n(n_), nf(nf_), pos(3 * n_), jpos(3 * n_), jzaxis(3 * n_), mom_rate(3 * n_), ang_mom_rate(3 * n_),
                               fpos(3 * nf_), contacts(nf_) {}
  int get_ncontacts() const { int s = 0; for (int i = 0; i < nf; i++) s += contacts[i]; return s; }
};

// forcetorquesolver (ftsolver.h:37-69): solve_forcetorques on one populated dynrecord.  x = joint forces then joint
// torques (6n), z = contact forces of all feet in LIK order (3 nf); get_fts / motor torques as in the reference.
class forcetorquesolver {
  const kinematicmodel* model_;
  std::vector<double> fts_, tau_;
  bool pen_force_, pen_torque_;
 public:
  explicit forcetorquesolver(const kinematicmodel* model) : model_(model), pen_force_(false), pen_torque_(false) {}
  void switch_torso_penalty(bool force_flag, bool torque_flag) { pen_force_ = force_flag; pen_torque_ = torque_flag; }
  void solve_forcetorques(const dynrecord* rec, std::vector<double>& x, std::vector<double>& z) {
    if (!(pen_force_ && pen_torque_)) {
      if (!pen_force_ && !pen_torque_) throw error("ERROR: mask0 not set");  // ftsolver.cpp:245
      throw error("only switch_torso_penalty(1,1) is supported on the GPU path");
    }
    const int n = model_->number_of_parts(), nf = model_->number_of_limbs(), nmj = model_->number_of_motor_joints();
    x.assign(6 * n, 0.0); z.assign(3 * nf, 0.0); tau_.assign(nmj, 0.0);
    int32_t status = 0;
    check(hsl_solve_frames_host(model_->handle(), 1, rec->pos.data(), rec->jpos.data(), rec->jzaxis.data(), rec->mom_rate.data(),
                                rec->ang_mom_rate.data(), rec->fpos.data(), rec->contacts.data(), x.data(), z.data(), tau_.data(), &status));
    fts_ = x;
  }
  // ftsolver.cpp:331-378: z = given motor torques (nmj), y = contact forces of all feet (3 nf); the reference's argument names
  void solve_forces(const dynrecord* rec, const std::vector<double>& z, std::vector<double>& y) {
    const int nf = model_->number_of_limbs(), nmj = model_->number_of_motor_joints();
    if ((int)z.size() != nmj) throw error("solve_forces: one torque per motor joint expected");
    y.assign(3 * nf, 0.0);
    int32_t status = 0;
    check(hsl_solve_forces_host(model_->handle(), 1, rec->pos.data(), rec->jpos.data(), rec->jzaxis.data(), rec->mom_rate.data(),
                                rec->ang_mom_rate.data(), rec->fpos.data(), z.data(), y.data(), &status));
    if (status & HSL_ST_SOLVER) std::cout << "WARNING: singular limb in solve_forces" << std::endl;
  }
  const std::vector<double>* get_fts() const { return &fts_; }
  const std::vector<double>& get_motor_torques() const { return tau_; }  // periodic::get_motor_torques
};

// periodic (periodic.h:27-87)
class periodic {
  const kinematicmodel* model_;
  int n_t_, config_dim_, nmj_, n_, nf_;
  pgsconfigparams pcp_;
  rectransform rec_;
  bool have_, pen_force_, pen_torque_;
  int flags_;
  double work_, cot_, min_cfz_, max_mu_;
  std::vector<double> traj_, x_, z_, tau_, vel_;
  void evaluate() {
    if (have_) return;
    if (n_t_ <= 0) throw error("ERROR: no data");
    if (!(pen_force_ && pen_torque_)) throw error("only switch_torso_penalty(1,1) is supported on the GPU path");
    double row[HSL_NPARAM];
    pcp_.to_row(row);
    traj_.assign((size_t)(n_t_ + 4) * config_dim_, 0.0);
    x_.assign((size_t)n_t_ * 6 * n_, 0.0);
    z_.assign((size_t)n_t_ * 3 * nf_, 0.0);
    tau_.assign((size_t)n_t_ * nmj_, 0.0);
    int32_t status = 0;
    rec_.apply(model_->handle());
    check(hsl_eval_gaits_detail_host(model_->handle(), 1, n_t_, row, flags_, &cot_, &work_, &min_cfz_, &max_mu_, &status, traj_.data(),
                                     x_.data(), z_.data(), tau_.data(), nullptr));
    if (status & HSL_ST_UNREACHABLE) throw error("LIK ERROR: limb position is unreachable");
    if (status & HSL_ST_BAD_PARAMS) throw error("ERROR: step_duration out of bounds");
    // the reference prints this when its retry loop lowers the rank (ftsolver.cpp:219); here it marks the same regime
    if (status & HSL_ST_ILLCOND) std::cout << "WARNING: decomposition threshold increased" << std::endl;
    // periodic::compute_vel_traj (periodic.cpp:261-282) on the host, frames 1..n_t+2
    const double dt = pcp_.TLh[0] / n_t_;
    vel_.assign((size_t)(n_t_ + 4) * config_dim_, 0.0);
    for (int i = 2; i < n_t_ + 4; i++)
      for (int j = 0; j < config_dim_; j++) {
        double d = traj_[(size_t)i * config_dim_ + j] - traj_[(size_t)(i - 2) * config_dim_ + j];
        if (d > M_PI) d -= 2 * M_PI; else if (d < -M_PI) d += 2 * M_PI;
        vel_[(size_t)(i - 1) * config_dim_ + j] = d / (2 * dt);
      }
    have_ = true;
  }
  int row_of(int i) const {  // reference stores solved frame i (2..n_t+1) at index i % n_t; the dumps are in solve order
    int k = i % n_t_;
    int frame = (k < 2) ? k + n_t_ : k;
    return frame - 2;
  }
 public:
  explicit periodic(const kinematicmodel* model) : model_(model), n_t_(0), have_(false), pen_force_(false), pen_torque_(false), flags_(0),
                                                   work_(0), cot_(0), min_cfz_(0), max_mu_(0) {
    config_dim_ = model->get_config_dim(); nmj_ = model->number_of_motor_joints(); n_ = model->number_of_parts(); nf_ = model->number_of_limbs();
  }
  int get_nt() const { return n_t_; }
  int get_nfeet() const { return nf_; }
  int get_number_of_dynparts() const { return n_; }
  void set_ignore_reach_flag(bool v) { flags_ = v ? HSL_FLAG_IGNORE_REACH : 0; have_ = false; }
  void record_trajectory(const pergensetup* pgs, int n_t) { pcp_ = pgs->params(); rec_ = pgs->rec_transform(); n_t_ = n_t; have_ = false; }
  void compute_dynrecs() {}
  void compute_dynrec_ders() {}
  void switch_torso_penalty(bool force, bool torque) { pen_force_ = force; pen_torque_ = torque; }
  double get_total_mass() const { return (double)n_; }  // unit masses (dynrec.cpp:62-68)
  void compute_torques_over_period() { evaluate(); }
  double work_over_period() { evaluate(); return work_; }
  double cost_of_transport() { evaluate(); return cot_; }
  void get_contforce_stat(double* stat) { evaluate(); stat[0] = min_cfz_; stat[1] = max_mu_; }
  const double* get_computed_torques(int i) { evaluate(); return &tau_[(size_t)row_of(i) * nmj_]; }
  // forcetorquesolver::solve_forcetorques results of frame i (2 <= i <= n_t+1)
  void solve_torques_contforces(int i, double* torques, double* contforces) {
    evaluate();
    if (i < 2 || i > n_t_ + 1) throw error("frame index out of the solved range");
    for (int j = 0; j < nmj_; j++) torques[j] = tau_[(size_t)(i - 2) * nmj_ + j];
    for (int j = 0; j < 3 * nf_; j++) contforces[j] = z_[(size_t)(i - 2) * 3 * nf_ + j];
  }
  // periodic.cpp:369-374: contact forces of frame i (2 <= i <= n_t+1) for the given motor torques
  void solve_contforces_given_torques(int i, double* contforces, double* torques) {
    if (n_t_ <= 0) throw error("ERROR: no data");
    if (i < 2 || i > n_t_ + 1) throw error("frame index out of the solved range");
    double row[HSL_NPARAM];
    pcp_.to_row(row);
    // the entry solves every frame of the gait; the other frames get the torques of the forward solve
    evaluate();
    std::vector<double> tq(tau_), zz((size_t)n_t_ * 3 * nf_);
    for (int j = 0; j < nmj_; j++) tq[(size_t)(i - 2) * nmj_ + j] = torques[j];
    int32_t status = 0;
    rec_.apply(model_->handle());
    check(hsl_solve_forces_gait_host(model_->handle(), 1, n_t_, row, flags_, tq.data(), zz.data(), &status));
    for (int j = 0; j < 3 * nf_; j++) contforces[j] = zz[(size_t)(i - 2) * 3 * nf_ + j];
  }
  const double* get_fts(int i) { evaluate(); return &x_[(size_t)(i - 2) * 6 * n_]; }  // forcetorquesolver::get_fts
  const double* get_traj(int i) { evaluate(); return &traj_[(size_t)i * config_dim_]; }
  void get_motor_adas(int tsi, double* as, double* das) {  // periodic.cpp:394-404
    evaluate();
    tsi %= n_t_;
    if (tsi < 2) tsi += n_t_;
    for (int j = 6; j < config_dim_; j++) { *as++ = traj_[(size_t)tsi * config_dim_ + j]; *das++ = vel_[(size_t)tsi * config_dim_ + j]; }
  }
  void get_complete_traj_rec(int tsi, double* rec) {  // periodic.cpp:408-418: [q, qdot, tau]
    evaluate();
    if (tsi >= n_t_) throw error("ERROR: time step must be < n_t");
    const int t2 = (tsi < 2) ? tsi + n_t_ : tsi;
    for (int j = 0; j < config_dim_; j++) { rec[j] = traj_[(size_t)t2 * config_dim_ + j]; rec[config_dim_ + j] = vel_[(size_t)t2 * config_dim_ + j]; }
    const double* tq = get_computed_torques(t2);
    for (int j = 0; j < nmj_; j++) rec[2 * config_dim_ + j] = tq[j];
  }
  void get_complete_traj(double** complete_traj) { for (int i = 0; i < n_t_; i++) get_complete_traj_rec(i, complete_traj[i]); }
};

// save_2d_array (core.cpp:47-61): the traj.txt wire format
inline void save_2d_array(double** array, int n, int m, const std::string& fname, bool append_flag) {
  std::ofstream file;
  if (append_flag) file.open(fname.c_str(), std::ios_base::app); else file.open(fname.c_str());
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < m; j++) { if (j) file << " "; file << array[i][j]; }
    file << std::endl;
  }
}

// modelplayer: the evaluation entry points (player.cpp:147-208, 230-285, 311-321, 619-655)
class modelplayer {
  kinematicmodel model_;
  bool contact_force_flag_;
  double play_dt_;
  int flags_;
 public:
  modelplayer() : contact_force_flag_(false), play_dt_(.01), flags_(0) {}
  kinematicmodel* get_model() { return &model_; }
  void set_play_dt(double dt) { play_dt_ = dt; }
  void set_flag(const std::string& name, bool value) { if (name == "contact_force") contact_force_flag_ = value; }
  void ignore_reach() { flags_ = HSL_FLAG_IGNORE_REACH; }
  void load_model(const std::string& fname) { model_.load_fromxml(fname); }
  static void get_pgs_config_params(const std::string& rec_str, pgsconfigparams& pcp) {  // player.cpp:170-208
    std::stringstream ss(rec_str);
    std::string key;
    double period = 0, sl = 0, sh = 0;
    while (ss >> key) {
      if (key == "xml_file") ss >> pcp.fname;
      else if (key == "torso_pos") ss >> pcp.orientation[0][0] >> pcp.orientation[0][1] >> pcp.orientation[0][2];
      else if (key == "torso_angles") ss >> pcp.orientation[1][0] >> pcp.orientation[1][1] >> pcp.orientation[1][2];
      else if (key == "step_duration") ss >> pcp.step_duration;
      else if (key == "period") ss >> period;
      else if (key == "step_length") ss >> sl;
      else if (key == "step_height") ss >> sh;
      else if (key == "curvature") ss >> pcp.curvature;
      else if (key == "lateral_foot_shift") { double s; ss >> s; pcp.foot_shift = std::make_pair(0, s); }
      else if (key == "radial_foot_shift") { double s; ss >> s; pcp.foot_shift = std::make_pair(1, s); }
      else throw error("ERROR: unknown key " + key);
    }
    pcp.set_TLh(period, sl, sh);
  }
  static void get_rec_str(std::string& rec_str, const std::string& fname, int rec_id) {  // player.cpp:230-244
    std::ifstream file(fname.c_str());
    std::string str;
    while (std::getline(file, str)) {
      std::stringstream ss(str);
      int id;
      if ((ss >> id) && id == rec_id) { rec_str = str.substr(str.find_first_of(" \t") + 1); return; }
    }
    throw error("ERROR: no string with rec_id");
  }
  // model_dir: where the xml named in the preset lives (the reference resolves it against the working directory)
  pergensetup* make_pergensu(const std::string& config_fname, int setup_id, const std::string& model_dir = ".") {
    std::string rec;
    get_rec_str(rec, config_fname, setup_id);
    pgsconfigparams pcp;
    get_pgs_config_params(rec, pcp);
    const std::string path = model_dir + "/" + pcp.fname;
    if (!model_.if_loaded()) load_model(path);
    else if (model_.get_xmlfname() != path) throw error("ERROR: model not from " + pcp.fname);
    pergensetup* pgs = new pergensetup(model_.number_of_limbs(), &model_);
    pgs->set_config_params(pcp);
    return pgs;
  }
  void prepare_per_traj_dyn(periodic& per, pergensetup* pgs, int n_t) {  // player.cpp:259-264
    per.record_trajectory(pgs, n_t);
    per.compute_dynrecs();
    per.compute_dynrec_ders();
    per.switch_torso_penalty(1, 1);
  }
  double measure_cot(pergensetup* pgs, int n_t) {  // player.cpp:269-285
    periodic per(&model_);
    per.set_ignore_reach_flag(flags_ != 0);
    prepare_per_traj_dyn(per, pgs, n_t);
    double work = per.work_over_period();
    double cot = work / (per.get_total_mass() * pgs->get_pergen()->get_step_length());
    if (contact_force_flag_) {
      double stat[2];
      per.get_contforce_stat(stat);
      std::cout << "min cfz = " << stat[0] << ", max mu = " << stat[1] << std::endl;
    }
    return cot;
  }
  // player.cpp:311-321, but all n_val+1 candidates go to the GPU as one batch
  void measure_cot_sweep(pergensetup* pgs, int n_t, const std::string& param_name, double val0, double val1, int n_val,
                         std::vector<double>* vals_out = nullptr, std::vector<double>* cots_out = nullptr) {
    pgssweeper sweeper(pgs, &model_);
    sweeper.sweep(param_name, val0, val1, n_val);
    std::vector<double> rows, vals;
    while (sweeper.next()) {
      double row[HSL_NPARAM];
      sweeper.get_pgs()->params().to_row(row);
      rows.insert(rows.end(), row, row + HSL_NPARAM);
      vals.push_back(sweeper.get_val());
    }
    std::vector<double> cot(vals.size());
    std::vector<int32_t> status(vals.size());
    pgs->rec_transform().apply(model_.handle());
    check(hsl_eval_gaits_host(model_.handle(), (int64_t)vals.size(), n_t, rows.data(), flags_, cot.data(), nullptr, nullptr, nullptr, status.data()));
    for (size_t i = 0; i < vals.size(); i++) std::cout << "val = " << vals[i] << " COT = " << cot[i] << std::endl;
    if (vals_out) *vals_out = vals;
    if (cots_out) *cots_out = cot;
  }
  void record_per_traj(pergensetup* pgs, const std::string& fname = "traj.txt") {  // player.cpp:619-631
    const int n_t = int(pgs->get_period() / play_dt_ + .5);
    const int rec_len = 2 * model_.get_config_dim() + model_.number_of_motor_joints();
    std::vector<std::vector<double> > buf(n_t, std::vector<double>(rec_len));
    std::vector<double*> rows(n_t);
    for (int i = 0; i < n_t; i++) rows[i] = buf[i].data();
    periodic per(&model_);
    per.set_ignore_reach_flag(flags_ != 0);
    prepare_per_traj_dyn(per, pgs, n_t);
    per.get_complete_traj(rows.data());
    save_2d_array(rows.data(), n_t, rec_len, fname, false);
  }
};

}  // namespace hsl
