// TEST INFRASTRUCTURE ONLY -- quad-precision build of the CPU oracle (liborcq.so).
// The same headers compiled with real = __float128: every formula of the reference path is evaluated with ~34
// significant digits on the same double inputs, with the same double-valued constants and rank thresholds.  The
// difference to the double oracle bounds the round-off of the FP64 reference algorithm itself, which is the natural
// yardstick for the GPU-vs-oracle tolerance (pin 8).  Nothing under hslabs_b200/ may use this file.
#define ORC_QUAD 1
#include "orc_dynamics.hpp"

using namespace orc;

namespace {
struct Handle { Model* model; };
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
void narrow(const std::vector<real>& src, double* dst) {
  if (dst) for (size_t i = 0; i < src.size(); i++) dst[i] = (double)src[i];
}
}  // namespace

extern "C" {

void* orcq_model_load(const char* path) {
  try {
    Handle* h = new Handle;
    h->model = new Model(path);
    return h;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "orcq_model_load: %s\n", e.what());
    return 0;
  }
}
void orcq_model_free(void* hv) {
  Handle* h = (Handle*)hv;
  if (!h) return;
  delete h->model;
  delete h;
}

// Same contract as orc_measure_cot; results rounded to double on the way out.
int orcq_measure_cot(void* hv, const double* params, int n_t, double* out4, double* traj, double* x, double* z, double* tau) {
  Model& m = *((Handle*)hv)->model;
  try {
    GaitSetup g(m.nlimbs());
    setup_gait(g, m, params_from(params));
    std::vector<real> qtraj((size_t)(n_t + 5) * m.config_dim()), qx((size_t)n_t * 6 * m.n()), qz((size_t)n_t * 3 * m.nlimbs()),
        qtau((size_t)n_t * m.nmj());
    CotResult r = measure_cot(m, g, n_t, qtraj.data(), qx.data(), qz.data(), qtau.data());
    out4[0] = (double)r.cot; out4[1] = (double)r.work; out4[2] = (double)r.min_cfz; out4[3] = (double)r.max_mu;
    narrow(qtraj, traj); narrow(qx, x); narrow(qz, z); narrow(qtau, tau);
    return r.status;
  } catch (const std::exception&) { return -1; }
}

}  // extern "C"
