// TEST INFRASTRUCTURE ONLY.  Serial CPU emulation of hsl_frames_kernel: the per-thread phase functions
// of hslabs_b200/csrc/hsl_frame.h are compiled for the host and run slot by slot, role by role, with a
// plain array standing in for shared memory and a loop boundary standing in for each block barrier.
// This lets the CPU-only test tier check the exact arithmetic the CUDA kernels run (same source) against
// the oracle.  It is never used by the product: hslabs_b200/ has no CPU path.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../hslabs_b200/csrc/hsl_frame.h"
#include "../../hslabs_b200/csrc/hsl_forces.h"
#include "../../hslabs_b200/csrc/hsl_model.h"
#include "../../hslabs_b200/csrc/hsl_pipe.h"
#include "../../hslabs_b200/csrc/hsl_fall_world.h"

int hsl_build_model_pod(const char* xml_path, HslModelPod* pod, char* err, int errlen);
int hsl_build_sim_pod(const char* xml_path, const HslModelPod* pod, HslSimPod* sim, char* err, int errlen);

namespace {
bool g_axis_special = true;  // emulate the kernels specialised for the model's hinge-axis pattern when there is one
template <int NF, int FB, int MODE, bool DUMP, int AXP = HSL_AXP_GENERIC>
void emulate(const HslModelPod& M, const HslFrameArgs& A) {
  const int roles = NF + 1;
  int64_t blocks;
  if (MODE == HSL_MODE_FIELDS) blocks = (A.n_frames + FB - 1) / FB;
  else {
    const int64_t slots = A.n_cand * (A.n_t + 4);
    blocks = (slots - 4 + (FB - 4) - 1) / (FB - 4);
    if (blocks < 1) blocks = 1;
  }
  std::vector<double> smem((size_t)HslSmem<NF, FB>::doubles_per_slot(M.ntrunk) * FB);
  std::vector<HslLegState<DUMP> > lst((size_t)NF * FB);
  std::vector<HslTrunkState> tst(FB);
  std::vector<HslSlot> sls(FB);
  std::vector<int> bad((size_t)roles * FB);
  for (int64_t b = 0; b < blocks; b++) {
    HslSmem<NF, FB> sm;
    sm.carve(smem.data(), M.ntrunk);
    std::fill(smem.begin(), smem.end(), NAN);
    std::fill(bad.begin(), bad.end(), 0);
    for (int s = 0; s < FB; s++) {
      HslSlot& sl = sls[s];
      sl.s = s;
      if (MODE == HSL_MODE_FIELDS) {
        const int64_t g = b * FB + s;
        sl.valid = g < A.n_frames;
        sl.i = (int32_t)(sl.valid ? g : A.n_frames - 1);
        sl.c = sl.i; sl.fo = sl.i; sl.interior = sl.valid;
      } else {
        const int per = A.n_t + 4;
        const int64_t g = b * (FB - 4) + s;
        sl.c = g / per;
        sl.i = (int32_t)(g - sl.c * per);
        sl.valid = sl.c < A.n_cand;
        if (!sl.valid) { sl.c = A.n_cand - 1; sl.i = 0; }
        sl.interior = sl.valid && s >= 2 && s < FB - 2 && sl.i >= 2 && sl.i <= A.n_t + 1;
        sl.fo = sl.c * A.n_t + (sl.i - 2);
      }
    }
    for (int r = 0; r < roles; r++)
      for (int s = 0; s < FB; s++) {
        if (r < NF) { phase_a_leg<NF, FB, MODE, DUMP, AXP>(M, A, sm, sls[s], r, lst[r * FB + s]); bad[r * FB + s] = lst[r * FB + s].bad; }
        else phase_a_trunk<NF, FB, MODE>(M, A, sm, sls[s], tst[s]);
      }
    for (int r = 0; r < roles; r++)
      for (int s = 0; s < FB; s++) {
        if (!sls[s].interior) continue;
        if (r < NF) { phase_b_leg<NF, FB, MODE, DUMP>(M, A, sm, sls[s], r, lst[r * FB + s]); bad[r * FB + s] |= lst[r * FB + s].bad; }
        else phase_b_trunk<NF, FB, MODE>(M, A, sm, sls[s], tst[s]);
      }
    for (int s = 0; s < FB; s++)
      if (sls[s].interior) bad[NF * FB + s] |= phase_c_trunk<NF, FB, MODE, DUMP>(M, A, sm, sls[s], tst[s]);
    for (int r = 0; r < NF; r++)
      for (int s = 0; s < FB; s++)
        if (sls[s].interior) phase_d_leg<NF, FB, MODE, DUMP>(M, A, sm, sls[s], r, lst[r * FB + s]);
    for (int s = 0; s < FB; s++)
      if (sls[s].interior) phase_e_trunk<NF, FB>(A, sm, sls[s]);
    for (int r = 0; r < roles; r++)
      for (int s = 0; s < FB; s++)
        if (bad[r * FB + s] && sls[s].valid && A.status) A.status[sls[s].c] |= bad[r * FB + s];
  }
}
// Serial emulation of hsl_forces_kernel (hsl_forces.h): phases A | B | F1 | F2 | F3.
template <int NF, int MODE>
void emulate_forces(const HslModelPod& M, const HslFrameArgs& A) {
  constexpr int FB = 32;
  typedef HslSmem<NF, FB, HSL_FORCES_PART> SM;
  int64_t blocks;
  if (MODE == HSL_MODE_FIELDS) blocks = (A.n_frames + FB - 1) / FB;
  else {
    const int64_t slots = A.n_cand * (A.n_t + 4);
    blocks = (slots - 4 + (FB - 4) - 1) / (FB - 4);
    if (blocks < 1) blocks = 1;
  }
  std::vector<double> smem((size_t)SM::doubles_per_slot(M.ntrunk) * FB);
  std::vector<HslLegState<true> > lst((size_t)NF * FB);
  std::vector<HslForcesLeg> fst((size_t)NF * FB);
  std::vector<HslTrunkState> tst(FB);
  std::vector<HslSlot> sls(FB);
  for (int64_t b = 0; b < blocks; b++) {
    SM sm;
    sm.carve(smem.data(), M.ntrunk);
    std::fill(smem.begin(), smem.end(), NAN);
    for (int s = 0; s < FB; s++) {
      HslSlot& sl = sls[s];
      sl.s = s;
      if (MODE == HSL_MODE_FIELDS) {
        const int64_t g = b * FB + s;
        sl.valid = g < A.n_frames;
        sl.i = (int32_t)(sl.valid ? g : A.n_frames - 1);
        sl.c = sl.i; sl.fo = sl.i; sl.interior = sl.valid;
      } else {
        const int per = A.n_t + 4;
        const int64_t g = b * (FB - 4) + s;
        sl.c = g / per;
        sl.i = (int32_t)(g - sl.c * per);
        sl.valid = sl.c < A.n_cand;
        if (!sl.valid) { sl.c = A.n_cand - 1; sl.i = 0; }
        sl.interior = sl.valid && s >= 2 && s < FB - 2 && sl.i >= 2 && sl.i <= A.n_t + 1;
        sl.fo = sl.c * A.n_t + (sl.i - 2);
      }
    }
    auto flag = [&](const HslSlot& sl, int bad) { if (bad && sl.valid && A.status) A.status[sl.c] |= bad; };
    for (int r = 0; r <= NF; r++)
      for (int s = 0; s < FB; s++) {
        if (r < NF) { phase_a_leg<NF, FB, MODE, true>(M, A, sm, sls[s], r, lst[r * FB + s]); flag(sls[s], lst[r * FB + s].bad); }
        else phase_a_trunk<NF, FB, MODE>(M, A, sm, sls[s], tst[s]);
      }
    for (int r = 0; r <= NF; r++)
      for (int s = 0; s < FB; s++) {
        if (!sls[s].interior) continue;
        if (r < NF) phase_b_leg<NF, FB, MODE, true>(M, A, sm, sls[s], r, lst[r * FB + s]);
        else phase_b_trunk<NF, FB, MODE>(M, A, sm, sls[s], tst[s]);
      }
    for (int r = 0; r < NF; r++)
      for (int s = 0; s < FB; s++)
        if (sls[s].interior) { lst[r * FB + s].bad = 0; forces_f1_leg<NF, FB, MODE, true>(M, A, sm, sls[s], r, lst[r * FB + s], fst[r * FB + s]); flag(sls[s], lst[r * FB + s].bad); }
    for (int s = 0; s < FB; s++)
      if (sls[s].interior) flag(sls[s], forces_f2_trunk<NF, FB, MODE>(M, A, sm, sls[s], tst[s]));
    for (int r = 0; r < NF; r++)
      for (int s = 0; s < FB; s++)
        if (sls[s].interior) forces_f3_leg<NF, FB, MODE, true>(M, A, sm, sls[s], r, lst[r * FB + s], fst[r * FB + s]);
  }
}

// Serial emulation of hsl_gait_pipe_kernel with `grid` persistent blocks: the two halves of every iteration are run
// role by role exactly in the order the device schedule allows (see hsl_pipe.h), so a wrong buffer hand-off shows up
// here as a wrong result.
template <int NF, int FB, int AXP = HSL_AXP_GENERIC>
void emulate_pipe(const HslModelPod& M, const HslFrameArgs& A, int grid) {
  const int64_t slots = A.n_cand * (A.n_t + 4);
  int64_t n_tiles = (slots - 4 + (FB - 4) - 1) / (FB - 4);
  if (n_tiles < 1) n_tiles = 1;
  const int per = A.n_t + 4;
  std::vector<double> smem((size_t)HslPipeSmem<NF, FB>::doubles_per_slot(M.ntrunk) * FB);
  std::vector<HslLegState<false> > lst((size_t)NF * FB);
  std::vector<HslTrunkState> tst(FB);
  for (int blk = 0; blk < grid; blk++) {
    HslPipeSmem<NF, FB> sm;
    sm.carve(smem.data(), M.ntrunk);
    std::fill(smem.begin(), smem.end(), NAN);
    std::vector<HslSlot> cur(FB), p1(FB), p2(FB);
    for (int s = 0; s < FB; s++) { p1[s].interior = p2[s].interior = false; p1[s].s = p2[s].s = s; }
    auto trunk_c = [&](std::vector<HslSlot>& ps) {  // solver role: own trunk state, wrench from `twr`
      for (int s = 0; s < FB; s++)
        if (ps[s].interior) {
          HslTrunkState sv;
          for (int k = 0; k < 3; k++) { sv.F0[k] = sm.twr[k * FB + s]; sv.T0[k] = sm.twr[(3 + k) * FB + s]; }
          int tb = phase_c_trunk<NF, FB, HSL_MODE_GAIT, false>(M, A, sm, ps[s], sv);
          if (tb && A.status) A.status[ps[s].c] |= tb;
        }
    };
    auto trunk_e = [&](std::vector<HslSlot>& ps) {
      for (int s = 0; s < FB; s++) if (ps[s].interior) pipe_e_trunk<NF, FB>(A, sm, s, ps[s].fo);
    };
    auto legs_d = [&](std::vector<HslSlot>& ps) {
      for (int r = 0; r < NF; r++) for (int s = 0; s < FB; s++) if (ps[s].interior) pipe_d_leg<NF, FB>(sm, s, r);
    };
    for (int64_t tile = blk; tile < n_tiles; tile += grid) {
      for (int s = 0; s < FB; s++) {
        HslSlot& sl = cur[s];
        sl.s = s;
        const int64_t g = tile * (FB - 4) + s;
        sl.c = g / per; sl.i = (int32_t)(g - sl.c * per);
        sl.valid = sl.c < A.n_cand;
        if (!sl.valid) { sl.c = A.n_cand - 1; sl.i = 0; }
        sl.interior = sl.valid && s >= 2 && s < FB - 2 && sl.i >= 2 && sl.i <= A.n_t + 1;
        sl.fo = sl.c * A.n_t + (sl.i - 2);
      }
      // first half: trunk E(t-2), C(t-1), A'(t) ; limbs A(t)   (any interleaving is legal on the device; limbs first here)
      for (int r = 0; r < NF; r++)
        for (int s = 0; s < FB; s++) phase_a_leg<NF, FB, HSL_MODE_GAIT, false, AXP>(M, A, sm, cur[s], r, lst[r * FB + s]);
      trunk_e(p2);
      trunk_c(p1);
      for (int s = 0; s < FB; s++) phase_a_trunk<NF, FB, HSL_MODE_GAIT>(M, A, sm, cur[s], tst[s]);
      // second half: limbs D(t-1), B(t) ; trunk B'(t)
      legs_d(p1);
      for (int r = 0; r < NF; r++)
        for (int s = 0; s < FB; s++) {
          HslLegState<false>& st = lst[r * FB + s];
          if (cur[s].interior) {
            phase_b_leg<NF, FB, HSL_MODE_GAIT, false>(M, A, sm, cur[s], r, st);
            pipe_store_dstate<NF, FB>(sm, s, r, st);
          }
          if (st.bad && cur[s].valid && A.status) A.status[cur[s].c] |= st.bad;
        }
      for (int s = 0; s < FB; s++)
        if (cur[s].interior) {
          phase_b_trunk<NF, FB, HSL_MODE_GAIT>(M, A, sm, cur[s], tst[s]);
          for (int k = 0; k < 3; k++) { sm.twr[k * FB + s] = tst[s].F0[k]; sm.twr[(3 + k) * FB + s] = tst[s].T0[k]; }
          tst[s].F0[0] = tst[s].T0[0] = NAN;  // the solver role must not see the trunk role's registers
        }
      p2 = p1;
      p1 = cur;
    }
    trunk_e(p2);
    trunk_c(p1);
    legs_d(p1);
    trunk_e(p1);
  }
}

template <int NF>
void run(const HslModelPod& M, const HslFrameArgs& A, int mode) {
  const int axp = g_axis_special ? hsl_axis_pattern(M) : HSL_AXP_GENERIC;
  if (mode == HSL_MODE_GAIT && axp == HSL_AXP_YXX) emulate<NF, 32, HSL_MODE_GAIT, true, HSL_AXP_YXX>(M, A);
  else if (mode == HSL_MODE_GAIT && axp == HSL_AXP_ZXX) emulate<NF, 32, HSL_MODE_GAIT, true, HSL_AXP_ZXX>(M, A);
  else if (mode == HSL_MODE_GAIT) emulate<NF, 32, HSL_MODE_GAIT, true>(M, A);
  else if (mode == HSL_MODE_TRAJ) emulate<NF, 32, HSL_MODE_TRAJ, true>(M, A);
  else emulate<NF, 32, HSL_MODE_FIELDS, true>(M, A);
}
void run_any(const HslModelPod& M, const HslFrameArgs& A, int mode) {
  if (M.nf == 6) run<6>(M, A, mode); else run<4>(M, A, mode);
}
void transpose_out(const std::vector<double>& src, int comps, int64_t nfr, double* dst) {
  if (!dst) return;
  for (int c = 0; c < comps; c++)
    for (int64_t f = 0; f < nfr; f++) dst[f * comps + c] = src[(size_t)c * nfr + f];
}
void finish(int64_t C, int n_t, double total_mass, const HslCand* cand, const double* dt_in, const std::vector<double>& wf,
            const std::vector<double>& fmn, const std::vector<double>& fmx, const int32_t* status, double* cot, double* work,
            double* min_cfz, double* max_mu) {
  for (int64_t c = 0; c < C; c++) {
    const double dt = cand ? cand[c].dt : dt_in[c];
    double w = 0, mn = 1e10, mx = -1e10;
    for (int f = 0; f < n_t; f++) { w += wf[c * n_t + f] * dt; mn = std::fmin(mn, fmn[c * n_t + f]); mx = std::fmax(mx, fmx[c * n_t + f]); }
    const bool fatal = status && (status[c] & (HSL_ST_BAD_PARAMS | HSL_ST_UNREACHABLE));
    if (work) work[c] = fatal ? NAN : w;
    if (cot) cot[c] = (fatal || !cand) ? NAN : w / (total_mass * cand[c].step_length);
    if (min_cfz) min_cfz[c] = fatal ? NAN : mn;
    if (max_mu) max_mu[c] = fatal ? NAN : mx;
  }
}
}  // namespace

extern "C" {

int hc_model_pod(const char* xml, HslModelPod* pod) {
  char err[256];
  return hsl_build_model_pod(xml, pod, err, sizeof err);
}

int hc_setup_candidate(const char* xml, const double* params, int n_t, HslCand* cd, double* ttab) {
  HslModelPod M;
  char err[256];
  int rc = hsl_build_model_pod(xml, &M, err, sizeof err);
  if (rc) return rc;
  setup_candidate(M, params, n_t, *cd, ttab);
  return 0;
}

// 1 (default): emulate the kernels specialised for the model's hinge-axis pattern; 0: the generic ones
void hc_set_axis_specialisation(int on) { g_axis_special = (on != 0); }
int hc_axis_pattern(const char* xml) {
  HslModelPod M;
  char err[256];
  if (hsl_build_model_pod(xml, &M, err, sizeof err)) return -1;
  return hsl_axis_pattern(M);
}

// test-side mirror of hsl_set_rec_transform (process-wide here; the library keeps it per handle)
static bool g_rec_on = false;
static double g_rec_R[9], g_rec_t[3];
void hc_set_rec_transform(const double* transl, const double* eas) {
  g_rec_on = (transl != nullptr) || (eas != nullptr);
  if (!g_rec_on) return;
  const double z3[3] = {0, 0, 0};
  const double* e = eas ? eas : z3;
  const double* t = transl ? transl : z3;
  euler_to_R(e[0], e[1], e[2], g_rec_R);
  for (int k = 0; k < 3; k++) g_rec_t[k] = t[k];
}
static void apply_rec(HslFrameArgs& A) {
  if (!g_rec_on) return;
  A.flags |= HSL_FLAG_REC_TRANSFORM;
  memcpy(A.rec_R, g_rec_R, sizeof g_rec_R);
  memcpy(A.rec_t, g_rec_t, sizeof g_rec_t);
}

int hc_eval_gaits(const char* xml, int64_t C, int n_t, const double* params, int flags, double* cot, double* work, double* min_cfz,
                  double* max_mu, int32_t* status, double* traj, double* x, double* z, double* tau, uint8_t* contacts) {
  HslModelPod M;
  char err[256];
  int rc = hsl_build_model_pod(xml, &M, err, sizeof err);
  if (rc) return rc;
  const int64_t nfr = C * n_t;
  std::vector<HslCand> cand(C);
  std::vector<double> ttab((size_t)C * (n_t + 4)), wf(nfr), fmn(nfr), fmx(nfr);
  std::vector<double> dx((size_t)6 * M.n * nfr), dz((size_t)3 * M.nf * nfr), dtau((size_t)M.nmj * nfr), dq((size_t)M.config_dim * C * (n_t + 4));
  std::vector<uint8_t> dc((size_t)M.nf * nfr);
  std::vector<int32_t> st(C, 0);
  for (int64_t c = 0; c < C; c++) { setup_candidate(M, params + HSL_NPARAM * c, n_t, cand[c], &ttab[c * (n_t + 4)]); st[c] = cand[c].status; }
  HslFrameArgs A;
  memset(&A, 0, sizeof A);
  A.n_cand = C; A.n_t = n_t; A.flags = flags; A.n_frames = nfr;
  A.cand = cand.data(); A.ttab = ttab.data();
  A.wframe = wf.data(); A.fmin_cfz = fmn.data(); A.fmax_mu = fmx.data(); A.status = st.data();
  A.x = dx.data(); A.z = dz.data(); A.tau = dtau.data(); A.q_out = dq.data(); A.contacts = dc.data();
  apply_rec(A);
  run_any(M, A, HSL_MODE_GAIT);
  double tm = 0;
  for (int i = 0; i < M.n; i++) tm += 1.0;
  finish(C, n_t, tm, cand.data(), nullptr, wf, fmn, fmx, st.data(), cot, work, min_cfz, max_mu);
  if (status) memcpy(status, st.data(), sizeof(int32_t) * C);
  transpose_out(dx, 6 * M.n, nfr, x);
  transpose_out(dz, 3 * M.nf, nfr, z);
  transpose_out(dtau, M.nmj, nfr, tau);
  transpose_out(dq, M.config_dim, C * (n_t + 4), traj);
  if (contacts)
    for (int l = 0; l < M.nf; l++)
      for (int64_t f = 0; f < nfr; f++) contacts[f * M.nf + l] = dc[(size_t)l * nfr + f];
  return 0;
}

// The fall / perturbation sweep (hsl_fall_sweep_host) with the kernel's per-world code run serially on the host: same
// sequence as hsl_capi.cu (gait evaluation with ignore_reach -> control table -> initial body poses -> worlds).
int hc_fall_sweep(const char* xml, int64_t W, const double* params, double play_dt, double t0, int n_steps, const int32_t* kick_step,
                  const double* kick_dv, double hc, double tmin, uint8_t* fell, double* t_end, double* final_z, int32_t* status, double* traj) {
  HslModelPod M;
  static HslSimPod S;
  char err[256];
  int rc = hsl_build_model_pod(xml, &M, err, sizeof err);
  if (rc) return rc;
  if ((rc = hsl_build_sim_pod(xml, &M, &S, err, sizeof err))) return rc;
  const int n_t = (int)(params[7] / play_dt + .5);
  std::vector<double> cot(1), work(1), mn(1), mx(1), q((size_t)(n_t + 4) * M.config_dim), x((size_t)n_t * 6 * M.n), z((size_t)n_t * 3 * M.nf),
      tau((size_t)n_t * M.nmj);
  int32_t gst = 0;
  rc = hc_eval_gaits(xml, 1, n_t, params, HSL_FLAG_IGNORE_REACH, cot.data(), work.data(), mn.data(), mx.data(), &gst, q.data(), x.data(), z.data(),
                     tau.data(), nullptr);
  if (rc) return rc;
  std::vector<double> ctrl((size_t)n_t * 3 * M.nmj);
  for (int tm = 0; tm < n_t; tm++)
    for (int j = 0; j < M.nmj; j++)
      hsl_fall::fall_ctrl_entry(n_t, M.nmj, params[7] / n_t, q.data(), M.config_dim, 1, tau.data(), M.nmj, 1, tm, j, ctrl.data());
  const double play_t0 = (double)(int)(t0 / play_dt + .5) * play_dt;
  const int f0 = (int)(play_t0 / (params[7] / n_t) + .5);
  std::vector<double> A0((size_t)16 * M.n);
  for (int role = 0; role <= M.nf; role++) fk_record(M, role, &q[(size_t)f0 * M.config_dim], A0.data(), nullptr);
  static HslFallArgs A;
  memset(&A, 0, sizeof A);
  hsl_sim_state_from_frames(&S, A0.data(), A.pos0, A.quat0);
  A.n_worlds = W; A.n_steps = n_steps; A.n_t = n_t; A.iterations = 20;
  A.play_dt = play_dt; A.play_t0 = play_t0; A.hc = hc; A.tmin = tmin;
  A.erp = 0.8; A.cfm = 1e-10; A.soft_cfm = 1e-3; A.bounce = 0.5; A.bounce_vel = 0.1; A.gravity = 1.0; A.kp = 100.0;
  A.ctrl = ctrl.data(); A.kick_step = kick_step; A.kick_dv = kick_dv;
  A.fell = fell; A.t_end = t_end; A.final_z = final_z; A.status = status; A.traj = traj;
  if (const char* dbg = getenv("HSL_FALL_DEBUG")) {
    if (FILE* f = fopen(dbg, "wb")) {
      fwrite(ctrl.data(), sizeof(double), ctrl.size(), f);
      fwrite(A.pos0, sizeof(double), 3 * HSL_MAX_BODIES, f);
      fwrite(A.quat0, sizeof(double), 4 * HSL_MAX_BODIES, f);
      fwrite(&S, sizeof(HslSimPod), 1, f);
      fclose(f);
    }
  }
  hsl_fall::World* w = new hsl_fall::World;
  for (int64_t wi = 0; wi < W; wi++) hsl_fall::fall_world(S, A, wi, *w);
  delete w;
  return 0;
}

// forces from torques along a generated gait: tau [C][n_t][nmj] -> z [C][n_t][3nf]
int hc_solve_forces_gait(const char* xml, int64_t C, int n_t, const double* params, int flags, const double* tau, double* z, int32_t* status) {
  HslModelPod M;
  char err[256];
  int rc = hsl_build_model_pod(xml, &M, err, sizeof err);
  if (rc) return rc;
  const int64_t nfr = C * n_t;
  std::vector<HslCand> cand(C);
  std::vector<double> ttab((size_t)C * (n_t + 4)), dz((size_t)3 * M.nf * nfr, NAN);
  std::vector<int32_t> st(C, 0);
  for (int64_t c = 0; c < C; c++) { setup_candidate(M, params + HSL_NPARAM * c, n_t, cand[c], &ttab[c * (n_t + 4)]); st[c] = cand[c].status; }
  HslFrameArgs A;
  memset(&A, 0, sizeof A);
  A.n_cand = C; A.n_t = n_t; A.flags = flags; A.n_frames = nfr;
  A.cand = cand.data(); A.ttab = ttab.data(); A.status = st.data();
  A.tau_in = tau; A.z = dz.data();
  apply_rec(A);
  if (M.nf == 6) emulate_forces<6, HSL_MODE_GAIT>(M, A); else emulate_forces<4, HSL_MODE_GAIT>(M, A);
  if (status) memcpy(status, st.data(), sizeof(int32_t) * C);
  transpose_out(dz, 3 * M.nf, nfr, z);
  return 0;
}

int hc_eval_gaits_pipe(const char* xml, int64_t C, int n_t, const double* params, int flags, int fb, int grid, double* cot, double* work,
                       double* min_cfz, double* max_mu, int32_t* status) {
  HslModelPod M;
  char err[256];
  int rc = hsl_build_model_pod(xml, &M, err, sizeof err);
  if (rc) return rc;
  const int64_t nfr = C * n_t;
  std::vector<HslCand> cand(C);
  std::vector<double> ttab((size_t)C * (n_t + 4)), wf(nfr, NAN), fmn(nfr, NAN), fmx(nfr, NAN);
  std::vector<int32_t> st(C, 0);
  for (int64_t c = 0; c < C; c++) { setup_candidate(M, params + HSL_NPARAM * c, n_t, cand[c], &ttab[c * (n_t + 4)]); st[c] = cand[c].status; }
  HslFrameArgs A;
  memset(&A, 0, sizeof A);
  A.n_cand = C; A.n_t = n_t; A.flags = flags; A.n_frames = nfr;
  A.cand = cand.data(); A.ttab = ttab.data();
  A.wframe = wf.data(); A.fmin_cfz = fmn.data(); A.fmax_mu = fmx.data(); A.status = st.data();
  apply_rec(A);
  const int axp = g_axis_special ? hsl_axis_pattern(M) : HSL_AXP_GENERIC;
  if (M.nf == 6 && fb == 64 && axp == HSL_AXP_YXX) emulate_pipe<6, 64, HSL_AXP_YXX>(M, A, grid);
  else if (M.nf == 6 && fb == 64 && axp == HSL_AXP_ZXX) emulate_pipe<6, 64, HSL_AXP_ZXX>(M, A, grid);
  else if (M.nf == 6) { if (fb == 64) emulate_pipe<6, 64>(M, A, grid); else emulate_pipe<6, 32>(M, A, grid); }
  else { if (fb == 64) emulate_pipe<4, 64>(M, A, grid); else emulate_pipe<4, 32>(M, A, grid); }
  finish(C, n_t, (double)M.n, cand.data(), nullptr, wf, fmn, fmx, st.data(), cot, work, min_cfz, max_mu);
  if (status) memcpy(status, st.data(), sizeof(int32_t) * C);
  return 0;
}

int hc_eval_trajectories(const char* xml, int64_t C, int n_t, const double* traj, const double* dt, double* work, double* min_cfz,
                         double* max_mu, int32_t* status, double* x, double* z, double* tau) {
  HslModelPod M;
  char err[256];
  int rc = hsl_build_model_pod(xml, &M, err, sizeof err);
  if (rc) return rc;
  const int64_t nfr = C * n_t;
  std::vector<double> wf(nfr), fmn(nfr), fmx(nfr), dx((size_t)6 * M.n * nfr), dz((size_t)3 * M.nf * nfr), dtau((size_t)M.nmj * nfr);
  std::vector<int32_t> st(C, 0);
  HslFrameArgs A;
  memset(&A, 0, sizeof A);
  A.n_cand = C; A.n_t = n_t; A.n_frames = nfr; A.traj = traj; A.dt_in = dt;
  A.wframe = wf.data(); A.fmin_cfz = fmn.data(); A.fmax_mu = fmx.data(); A.status = st.data();
  A.x = dx.data(); A.z = dz.data(); A.tau = dtau.data();
  run_any(M, A, HSL_MODE_TRAJ);
  finish(C, n_t, (double)M.n, nullptr, dt, wf, fmn, fmx, st.data(), nullptr, work, min_cfz, max_mu);
  if (status) memcpy(status, st.data(), sizeof(int32_t) * C);
  transpose_out(dx, 6 * M.n, nfr, x);
  transpose_out(dz, 3 * M.nf, nfr, z);
  transpose_out(dtau, M.nmj, nfr, tau);
  return 0;
}

int hc_solve_frames(const char* xml, int64_t F, const double* pos, const double* jpos, const double* jz, const double* mom_rate,
                    const double* ang_rate, const double* fpos, const uint8_t* contacts, double* x, double* z, double* tau,
                    int32_t* status) {
  HslModelPod M;
  char err[256];
  int rc = hsl_build_model_pod(xml, &M, err, sizeof err);
  if (rc) return rc;
  std::vector<double> dx((size_t)6 * M.n * F), dz((size_t)3 * M.nf * F), dtau((size_t)M.nmj * F);
  std::vector<int32_t> st(F, 0);
  HslFrameArgs A;
  memset(&A, 0, sizeof A);
  A.n_cand = F; A.n_t = 1; A.n_frames = F;
  A.f_pos = pos; A.f_jpos = jpos; A.f_jz = jz; A.f_momrate = mom_rate; A.f_angrate = ang_rate; A.f_fpos = fpos; A.f_contacts = contacts;
  A.status = st.data(); A.x = dx.data(); A.z = dz.data(); A.tau = dtau.data();
  run_any(M, A, HSL_MODE_FIELDS);
  if (status) memcpy(status, st.data(), sizeof(int32_t) * F);
  transpose_out(dx, 6 * M.n, F, x);
  transpose_out(dz, 3 * M.nf, F, z);
  transpose_out(dtau, M.nmj, F, tau);
  return 0;
}

// forcetorquesolver::solve_forces on populated dynrecords: tau [F][nmj] -> z [F][3nf]
int hc_solve_forces_fields(const char* xml, int64_t F, const double* pos, const double* jpos, const double* jz, const double* mom_rate,
                           const double* ang_rate, const double* fpos, const double* tau, double* z, int32_t* status) {
  HslModelPod M;
  char err[256];
  int rc = hsl_build_model_pod(xml, &M, err, sizeof err);
  if (rc) return rc;
  std::vector<double> dz((size_t)3 * M.nf * F, NAN);
  std::vector<uint8_t> con((size_t)M.nf * F, 1);
  std::vector<int32_t> st(F, 0);
  HslFrameArgs A;
  memset(&A, 0, sizeof A);
  A.n_cand = F; A.n_t = 1; A.n_frames = F;
  A.f_pos = pos; A.f_jpos = jpos; A.f_jz = jz; A.f_momrate = mom_rate; A.f_angrate = ang_rate; A.f_fpos = fpos; A.f_contacts = con.data();
  A.status = st.data(); A.z = dz.data(); A.tau_in = tau;
  if (M.nf == 6) emulate_forces<6, HSL_MODE_FIELDS>(M, A); else emulate_forces<4, HSL_MODE_FIELDS>(M, A);
  if (status) memcpy(status, st.data(), sizeof(int32_t) * F);
  transpose_out(dz, 3 * M.nf, F, z);
  return 0;
}

// record-level entries (serial mirror of hsl_gait_records_kernel / hsl_ik_records_kernel)
int hc_gait_records(const char* xml, int64_t C, const double* params, int n_times, const double* times, int flags, double* rec,
                    int32_t* status) {
  HslModelPod M;
  char err[256];
  int rc = hsl_build_model_pod(xml, &M, err, sizeof err);
  if (rc) return rc;
  const int rl = 6 + 3 * M.nf;
  HslFrameArgs A;
  memset(&A, 0, sizeof A);
  A.n_cand = C; A.n_t = 1; A.flags = flags;
  apply_rec(A);
  for (int64_t c = 0; c < C; c++) {
    HslCand cd;
    double ttab[5];
    setup_candidate(M, params + HSL_NPARAM * c, 1, cd, ttab);
    if (status) status[c] = cd.status;
    for (int k = 0; k < n_times; k++)
      for (int role = 0; role <= M.nf; role++) gait_record(A, cd, M.nf, role, times[k], rec + (c * n_times + k) * rl);
  }
  return 0;
}
int hc_ik_records(const char* xml, int64_t n, const double* rec, int flags, double* q, int32_t* status) {
  HslModelPod M;
  char err[256];
  int rc = hsl_build_model_pod(xml, &M, err, sizeof err);
  if (rc) return rc;
  const int rl = 6 + 3 * M.nf;
  for (int64_t r = 0; r < n; r++) {
    int st = 0;
    for (int role = 0; role <= M.nf; role++)
      if (!ik_record(M, role, rec + r * rl, (flags & HSL_FLAG_IGNORE_REACH) != 0, q + r * M.config_dim)) st |= HSL_ST_UNREACHABLE;
    if (status) status[r] = st;
  }
  return 0;
}
}
