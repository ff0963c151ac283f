// Drives the C++ host mirror (include/hsl_host.hpp) the way the reference's main.cpp drives modelplayer
// (main.cpp:35,67-89): make_pergensu -> measure_cot_sweep / measure_cot / periodic, plus the first half of
// test_dynamics (playerexperim.cpp:95-121).  Output is parsed by tests/test_gpu_host_mirror.py.
#include <cstdio>
#include <cstdlib>

#include "hsl_host.hpp"

int main(int argc, char** argv) {
  if (argc < 4) { std::fprintf(stderr, "usage: %s <presets> <model_dir> <preset_id> [traj_out]\n", argv[0]); return 2; }
  try {
    hsl::modelplayer player0;
    player0.set_play_dt(.02);
    hsl::pergensetup* pgs = player0.make_pergensu(argv[1], std::atoi(argv[3]), argv[2]);
    player0.set_flag("contact_force", true);
    player0.measure_cot_sweep(pgs, 20, "period", 3, 18, 15);
    std::printf("COT = %.17g\n", player0.measure_cot(pgs, 20));
    hsl::periodic per(player0.get_model());
    per.record_trajectory(pgs, 20);
    per.compute_dynrecs();
    per.compute_dynrec_ders();
    per.switch_torso_penalty(1, 1);
    std::printf("work = %.17g\n", per.work_over_period());
    const int nmj = player0.get_model()->number_of_motor_joints(), nf = per.get_nfeet();
    std::vector<double> torques(nmj), cf(3 * nf);
    per.solve_torques_contforces(2, torques.data(), cf.data());
    std::printf("cfz:");
    for (int i = 0; i < nf; i++) std::printf(" %.17g", cf[3 * i + 2]);
    std::printf("\ntorques:");
    for (int i = 0; i < nmj; i++) std::printf(" %.17g", torques[i]);
    std::printf("\n");
    // second half of test_dynamics (playerexperim.cpp:112-117): contact forces back from the motor torques
    std::vector<double> cf1(3 * nf);
    per.solve_contforces_given_torques(2, cf1.data(), torques.data());
    double s = 0;
    for (int i = 0; i < 3 * nf; i++) { const double d = cf[i] - cf1[i]; s += d * d; }
    std::printf("s = %.3e\ncf1:", std::sqrt(s));
    for (int i = 0; i < 3 * nf; i++) std::printf(" %.17g", cf1[i]);
    std::printf("\n");
    if (argc > 4) player0.record_per_traj(pgs, argv[4]);
    // the two kinematic seams on their own: pergensetup::set_rec and kinematicmodel::set_jvalues_with_lik
    {
      const int cd = player0.get_model()->get_config_dim();
      std::vector<double> rec(6 + 3 * nf), q(cd);
      pgs->set_rec(rec.data(), 0.7);
      hsl::kinematicmodel* km = const_cast<hsl::kinematicmodel*>(player0.get_model());
      km->set_jvalues_with_lik(rec.data());
      km->get_jvalues(q.data());
      std::printf("rec:");
      for (size_t i = 0; i < rec.size(); i++) std::printf(" %.17g", rec[i]);
      std::printf("\njvalues:");
      for (int i = 0; i < cd; i++) std::printf(" %.17g", q[i]);
      std::printf("\n");
    }
    // main.cpp:38: extvec rec_eas (0,0,-1.571); pgs->set_rec_rotation(rec_eas);  -- sweep candidates inherit it (pergen.cpp:446)
    const double rec_eas[3] = {0, 0, -1.571};
    pgs->set_rec_rotation(rec_eas);
    std::printf("rotated COT = %.17g\n", player0.measure_cot(pgs, 20));
    std::vector<double> vals, cots;
    player0.measure_cot_sweep(pgs, 20, "step_length", 0.3, 0.5, 2, &vals, &cots);
    std::printf("rotated sweep:");
    for (size_t i = 0; i < cots.size(); i++) std::printf(" %.17g", cots[i]);
    std::printf("\n");
    delete pgs;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}
