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
#include <algorithm>
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
  void get_jvalues(double* values) const {
    for (int i = 0; i < dims_[3]; i++) values[i] = (i < (int)jvalues_.size()) ? jvalues_[i] : 0.0;
  }
  void set_ignore_reach_flag(bool v) { ignore_reach_ = v; }  // liksolver::set_ignore_reach_flag, lik.cpp:142-146
  bool get_ignore_reach_flag() const { return ignore_reach_; }
  // model.cpp:314-318: FK of the current joint values (one GPU launch); the frames are then read with get_A_ground /
  // get_joint_A_ground (modelnode::get_A_ground, modeljoint::get_A_ground: column-major 4x4 like `affine`)
  void recompute_modelnodes() {
    if ((int)jvalues_.size() != dims_[3]) jvalues_.assign(dims_[3], 0.0);
    A_.assign((size_t)16 * dims_[0], 0.0);
    J_.assign((size_t)16 * dims_[0], 0.0);
    check(hsl_fk_records_host(h_, 1, jvalues_.data(), A_.data(), J_.data()));
  }
  // model.cpp:403-409: new torso position / Euler angles, then FK
  void orient_torso(const double orientation[2][3]) {
    if ((int)jvalues_.size() != dims_[3]) jvalues_.assign(dims_[3], 0.0);
    for (int i = 0; i < 2; i++) for (int j = 0; j < 3; j++) jvalues_[j + 3 * i] = orientation[i][j];
    recompute_modelnodes();
  }
  const double* get_A_ground(int body) const { need_frames(); return &A_[(size_t)16 * body]; }        // get_mnode(i)->get_A_ground()
  const double* get_joint_A_ground(int body) const { need_frames(); return &J_[(size_t)16 * body]; }  // ...->get_joint()->get_A_ground()
  // periodic::set_dynparts tables (periodic.cpp:34-58)
  struct tables_t {
    std::vector<int32_t> parent, footis, limb_top;
    std::vector<double> masses, com_offset, foot_offset;
  };
  const tables_t& tables() const {
    if (tab_.parent.empty()) {
      tab_.parent.resize(dims_[0]); tab_.footis.resize(dims_[1]); tab_.limb_top.resize(dims_[1]); tab_.masses.resize(dims_[0]);
      tab_.com_offset.resize((size_t)3 * dims_[0]); tab_.foot_offset.resize((size_t)3 * dims_[1]);
      check(hsl_model_tables(h_, tab_.parent.data(), tab_.footis.data(), tab_.limb_top.data(), tab_.masses.data(),
                             tab_.com_offset.data(), tab_.foot_offset.data()));
    }
    return tab_;
  }
  // y = A_ground(body) * [v; 1]
  void to_ground(int body, const double* v, double* y) const {
    const double* A = get_A_ground(body);
    for (int i = 0; i < 3; i++) y[i] = A[i] * v[0] + A[4 + i] * v[1] + A[8 + i] * v[2] + A[12 + i];
  }
 private:
  void need_frames() const { if (A_.empty()) throw error("ERROR: recompute_modelnodes() has not been called"); }
  std::vector<double> jvalues_, A_, J_;
  mutable tables_t tab_;
  bool ignore_reach_ = false;
  friend class liksolver;
};

// liksolver (lik.h:40-56): limb inverse kinematics on the model's current configuration.  Each call is one small GPU
// launch; batches of records go through hsl_ik_records_host / hsl_gait_records_host directly.
class liksolver {
  kinematicmodel* model_;
 public:
  explicit liksolver(kinematicmodel* model) : model_(model) {}
  int get_number_of_limbs() const { return model_->number_of_limbs(); }
  double get_rcap() const { return model_->get_rcap(); }
  void set_ignore_reach_flag(bool value) const { model_->set_ignore_reach_flag(value); }
  // lik.cpp:89-99: rec holds the ground positions of all feet (the part of a record after the six torso values);
  // the torso stays where the model's joint values put it
  void place_limbs(const double* rec) const {
    const int nf = model_->number_of_limbs();
    std::vector<double> full(6 + 3 * nf);
    torso_values(full.data());
    for (int i = 0; i < 3 * nf; i++) full[6 + i] = rec[i];
    model_->set_jvalues_with_lik(full.data());
  }
  // lik.cpp:82-85: one limb; the other limbs keep their joint values
  void place_limb(int limbi, double x, double y, double z) const {
    const int nf = model_->number_of_limbs(), cd = model_->get_config_dim();
    std::vector<double> keep(cd), full(6 + 3 * nf);
    model_->get_jvalues(keep.data());
    model_->recompute_modelnodes();
    torso_values(full.data());
    const kinematicmodel::tables_t& t = model_->tables();
    for (int l = 0; l < nf; l++) model_->to_ground(t.footis[l], &t.foot_offset[3 * l], &full[6 + 3 * l]);  // current foot points: reachable
    full[6 + 3 * limbi] = x; full[7 + 3 * limbi] = y; full[8 + 3 * limbi] = z;
    model_->set_jvalues_with_lik(full.data());
    std::vector<double> q(cd);
    model_->get_jvalues(q.data());
    for (int k = 0; k < 3; k++) keep[6 + 3 * limbi + k] = q[6 + 3 * limbi + k];
    model_->set_jvalues(keep.data());
  }
  // lik.cpp:104-106: body position of the limb's top link (coincides with the hip joint in the reference models)
  void get_limb_hip_pos(int limbi, double pos[3]) const {
    model_->recompute_modelnodes();
    const double* A = model_->get_A_ground(model_->tables().limb_top[limbi]);
    for (int k = 0; k < 3; k++) pos[k] = A[12 + k];
  }
 private:
  void torso_values(double* six) const {
    std::vector<double> q(model_->get_config_dim());
    model_->get_jvalues(q.data());
    for (int k = 0; k < 6; k++) six[k] = q[k];
  }
};

// dynpart (dynrec.h:29-72): static properties and current positions of one body; positions follow the model's last
// recompute_modelnodes() (dynpart::recompute, dynrec.cpp:47-51)
class dynpart {
  const kinematicmodel* model_;
  int id_;
 public:
  dynpart(const kinematicmodel* model, int id) : model_(model), id_(id) {}
  int get_id() const { return id_; }
  int get_parent_id() const { return model_->tables().parent[id_]; }
  double get_mass() const { return model_->tables().masses[id_]; }
  bool if_foot() const { for (int f : model_->tables().footis) if (f == id_) return true; return false; }
  void get_com_pos(double pos[3]) const { model_->to_ground(id_, &model_->tables().com_offset[3 * id_], pos); }  // odepart::get_com_pos
  void get_joint_pos(double pos[3]) const {  // dynrec.cpp:31-39: joint frame origin, or the body origin without a joint
    const double* J = model_->get_joint_A_ground(id_);
    const double* A = (J[15] != 0) ? J : model_->get_A_ground(id_);
    for (int k = 0; k < 3; k++) pos[k] = A[12 + k];
  }
  void get_foot_pos(double pos[3]) const {   // odepart::get_foot_pos
    const kinematicmodel::tables_t& t = model_->tables();
    for (size_t l = 0; l < t.footis.size(); l++) if (t.footis[l] == id_) { model_->to_ground(id_, &t.foot_offset[3 * l], pos); return; }
    throw error("dynpart::get_foot_pos: not a foot");
  }
  void get_joint_zaxis(double axis[3]) const {  // dynrec.cpp:84-93
    const double* J = model_->get_joint_A_ground(id_);
    for (int k = 0; k < 3; k++) axis[k] = (J[15] != 0) ? J[8 + k] : 0.0;
  }
  const double* get_A_ground() const { return model_->get_A_ground(id_); }
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
  // periodic.h:44-48: what forcetorquesolver and callers read from periodic
  dynpart get_dynpart(int i) const { return dynpart(model_, i); }
  const int32_t* get_parentis() const { return model_->tables().parent.data(); }
  const int32_t* get_footis() const { return model_->tables().footis.data(); }
  const double* get_masses() const { return model_->tables().masses.data(); }
  void set_ignore_reach_flag(bool v) { flags_ = v ? HSL_FLAG_IGNORE_REACH : 0; have_ = false; }
  void record_trajectory(const pergensetup* pgs, int n_t) { pcp_ = pgs->params(); rec_ = pgs->rec_transform(); n_t_ = n_t; have_ = false; }
  void compute_dynrecs() {}
  void compute_dynrec_ders() {}
  void switch_torso_penalty(bool force, bool torque) { pen_force_ = force; pen_torque_ = torque; }
  double get_total_mass() const { double s = 0; for (double m : model_->tables().masses) s += m; return s; }  // periodic.cpp:320-325
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
// One process per GPU: rank / world and the NCCL communicator the costs are all-gathered over (hsl_nccl_comm_init or
// the host application's own ncclComm_t).  world == 1: single GPU, no collective.
// gather (optional): a connected HslGather of this job (hsl_gather_create / hsl_gather_connect) -- the costs then travel
// over NVLink peer memory inside the evaluation instead of through NCCL after it.
struct shard {
  int rank, world;
  void* nccl_comm;
  HslGather* gather;
  shard() : rank(0), world(1), nccl_comm(nullptr), gather(nullptr) {}
};

class modelplayer {
  kinematicmodel model_;
  bool contact_force_flag_;
  double play_dt_;
  int flags_;
  shard shard_;
 public:
  modelplayer() : contact_force_flag_(false), play_dt_(.01), flags_(0) {}
  void set_shard(const shard& s) { shard_ = s; }
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
    pgs->rec_transform().apply(model_.handle());
    if (shard_.world <= 1) {
      std::vector<int32_t> status(vals.size());
      check(hsl_eval_gaits_host(model_.handle(), (int64_t)vals.size(), n_t, rows.data(), flags_, cot.data(), nullptr, nullptr, nullptr, status.data()));
    } else {
      // candidates shard over the ranks in contiguous blocks of ceil(C/G) (never across GPUs); the costs are all-gathered
      // over NCCL so that every rank holds the whole sweep
      const int64_t c = (int64_t)vals.size(), per = (c + shard_.world - 1) / shard_.world;
      const int64_t lo = std::min<int64_t>(shard_.rank * per, c), hi = std::min<int64_t>(lo + per, c);
      std::vector<double> local(per, std::nan("")), all((size_t)per * shard_.world);
      if (shard_.gather) {   // per must equal the gather object's n_per_rank
        if (hsl_gather_size(shard_.gather) != per * shard_.world) throw error("ERROR: the gather object is sized for another sweep");
        check(hsl_eval_gaits_gather_host(model_.handle(), shard_.gather, hi - lo, n_t, hi > lo ? &rows[(size_t)lo * HSL_NPARAM] : nullptr, flags_,
                                         all.data(), nullptr));
      } else {
      if (hi > lo) {
        std::vector<int32_t> status(hi - lo);
        check(hsl_eval_gaits_host(model_.handle(), hi - lo, n_t, &rows[(size_t)lo * HSL_NPARAM], flags_, local.data(), nullptr, nullptr, nullptr, status.data()));
      }
      check(hsl_allgather_costs_host(shard_.nccl_comm, shard_.world, local.data(), per, all.data()));
      }
      for (int64_t i = 0; i < c; i++) cot[i] = all[i];  // rank r's block starts at r*per: contiguous for all but the padding
    }
    if (shard_.rank == 0)
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
