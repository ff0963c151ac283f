// Drives the C++ host mirror (include/hsl_host.hpp) the way the reference's main.cpp drives modelplayer
// (main.cpp:35,67-89): make_pergensu -> measure_cot_sweep / measure_cot / periodic, plus the first half of
// test_dynamics (playerexperim.cpp:95-121).  Output is parsed by tests/test_gpu_host_mirror.py.
#include <cstdio>
#include <cstdlib>
#include <ctime>

#include "hsl_host.hpp"

int main(int argc, char** argv) {
  if (argc < 4) { std::fprintf(stderr, "usage: %s <presets> <model_dir> <preset_id> [traj_out]\n", argv[0]); return 2; }
  try {
    if (const char* d = std::getenv("HSL_DEVICE")) hsl::check(hsl_set_device(std::atoi(d)));   // one process per GPU, all GPUs visible
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
    // liksolver / kinematicmodel::orient_torso / dynpart seams (lik.h:46-56, model.h:122-130, dynrec.h:29-72), the way
    // pgssweeper::partial_setup_pergen uses them (pergen.cpp:454-467): orient the torso, read the hip positions
    {
      const double no_rot[3] = {0, 0, 0};
      pgs->set_rec_rotation(no_rot);  // back to the identity map for the records below
      hsl::kinematicmodel* km = player0.get_model();
      hsl::liksolver lik(km);
      hsl::pgsconfigparams pcp;
      pgs->get_config_params(&pcp);
      km->orient_torso(pcp.orientation);
      std::printf("rcap = %.17g nlimbs = %d\n", lik.get_rcap(), lik.get_number_of_limbs());
      std::printf("hips:");
      for (int l = 0; l < nf; l++) { double h[3]; lik.get_limb_hip_pos(l, h); std::printf(" %.17g %.17g %.17g", h[0], h[1], h[2]); }
      std::printf("\n");
      std::vector<double> rec(6 + 3 * nf), q(km->get_config_dim());
      pgs->set_rec(rec.data(), 0.4);
      double six[2][3] = {{rec[0], rec[1], rec[2]}, {rec[3], rec[4], rec[5]}};
      km->orient_torso(six);
      lik.place_limbs(rec.data() + 6);
      km->get_jvalues(q.data());
      std::printf("place_limbs:");
      for (size_t i = 0; i < q.size(); i++) std::printf(" %.17g", q[i]);
      std::printf("\n");
      lik.place_limb(1, rec[9] + 0.05, rec[10], rec[11] + 0.03);  // move one foot, the others keep their joint values
      km->get_jvalues(q.data());
      std::printf("place_limb:");
      for (size_t i = 0; i < q.size(); i++) std::printf(" %.17g", q[i]);
      std::printf("\n");
      km->recompute_modelnodes();
      hsl::periodic per2(km);
      std::printf("dynparts:");
      for (int i = 0; i < per2.get_number_of_dynparts(); i++) {
        hsl::dynpart dp = per2.get_dynpart(i);
        double c[3], j[3], a[3];
        dp.get_com_pos(c); dp.get_joint_pos(j); dp.get_joint_zaxis(a);
        std::printf(" %d %d %g %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g", dp.get_id(), dp.get_parent_id(), dp.get_mass(),
                    (int)dp.if_foot(), c[0], c[1], c[2], j[0], j[1], j[2], a[0], a[1], a[2]);
      }
      std::printf("\nfeet:");
      for (int l = 0; l < nf; l++) { double f[3]; per2.get_dynpart(per2.get_footis()[l]).get_foot_pos(f); std::printf(" %.17g %.17g %.17g", f[0], f[1], f[2]); }
      std::printf("\ntotal mass = %g\n", per2.get_total_mass());
    }
    // sharded sweep: one process per GPU, costs all-gathered through the C ABI's NCCL entries.
    // HSL_RANK / HSL_WORLD / HSL_NCCL_ID_FILE come from the launcher (tests/test_gpu_multirank.py).
    if (const char* w = std::getenv("HSL_WORLD")) {
      hsl::shard sh;
      sh.world = std::atoi(w);
      sh.rank = std::atoi(std::getenv("HSL_RANK"));
      const std::string idf = std::getenv("HSL_NCCL_ID_FILE");
      HslNcclId id;
      if (sh.rank == 0) {
        hsl::check(hsl_nccl_unique_id(&id));
        FILE* f = std::fopen((idf + ".tmp").c_str(), "wb");
        std::fwrite(&id, sizeof id, 1, f);
        std::fclose(f);
        std::rename((idf + ".tmp").c_str(), idf.c_str());
      } else {
        FILE* f = nullptr;
        for (int tries = 0; tries < 600 && !(f = std::fopen(idf.c_str(), "rb")); tries++) { struct timespec ts = {0, 100000000}; nanosleep(&ts, nullptr); }
        if (!f || std::fread(&id, sizeof id, 1, f) != 1) throw hsl::error("no NCCL id from rank 0");
        std::fclose(f);
      }
      hsl::check(hsl_nccl_comm_init(&sh.nccl_comm, sh.world, &id, sh.rank));
      player0.set_shard(sh);
      const double zero[3] = {0, 0, 0};
      pgs->set_rec_rotation(zero);
      std::vector<double> v2, c2;
      player0.measure_cot_sweep(pgs, 20, "period", 3, 18, 15, &v2, &c2);
      std::printf("sharded sweep rank %d:", sh.rank);
      for (size_t i = 0; i < c2.size(); i++) std::printf(" %.17g", c2[i]);
      std::printf("\n");
      hsl::check(hsl_nccl_comm_destroy(sh.nccl_comm));
      // the same sweep with the costs gathered over NVLink peer memory by the evaluation's own finish kernel: every rank
      // publishes the IPC handle of its gather buffer as a file next to the NCCL id and maps the others'
      if (std::getenv("HSL_GATHER_FILES")) {
        const int64_t per = (16 + sh.world - 1) / sh.world;   // the sweep below has n_val + 1 = 16 points
        HslIpcHandle mine;
        std::vector<HslIpcHandle> all(sh.world);
        hsl::check(hsl_gather_create(sh.world, sh.rank, per, &sh.gather, &mine));
        {
          const std::string fn = idf + ".ipc" + std::to_string(sh.rank);
          FILE* f = std::fopen((fn + ".tmp").c_str(), "wb");
          std::fwrite(&mine, sizeof mine, 1, f);
          std::fclose(f);
          std::rename((fn + ".tmp").c_str(), fn.c_str());
        }
        for (int r = 0; r < sh.world; r++) {
          const std::string fn = idf + ".ipc" + std::to_string(r);
          FILE* f = nullptr;
          for (int tries = 0; tries < 600 && !(f = std::fopen(fn.c_str(), "rb")); tries++) { struct timespec ts = {0, 100000000}; nanosleep(&ts, nullptr); }
          if (!f || std::fread(&all[r], sizeof(HslIpcHandle), 1, f) != 1) throw hsl::error("no gather handle from a rank");
          std::fclose(f);
        }
        hsl::check(hsl_gather_connect(sh.gather, all.data()));
        player0.set_shard(sh);
        std::vector<double> v3, c3;
        player0.measure_cot_sweep(pgs, 20, "period", 3, 18, 15, &v3, &c3);
        std::printf("peer-gathered sweep rank %d:", sh.rank);
        for (size_t i = 0; i < c3.size(); i++) std::printf(" %.17g", c3[i]);
        std::printf("\n");
        hsl::check(hsl_gather_free(sh.gather));   // safe: this rank has waited for every rank's flags of the last scatter
      }
    }
    delete pgs;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}
