// C ABI of libhsl_b200.so (include/hsl.h): model handle, device workspace, host/device entry points.
// No CPU fallback exists: every compute entry needs a CUDA device and fails with HSL_ERR_CUDA otherwise.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hsl.h"
#include "hsl_fall.h"
#include "hsl_internal.h"

namespace {
thread_local char g_err[512] = "";

int set_err(int code, const char* fmt, const char* a = "") {
  snprintf(g_err, sizeof g_err, fmt, a);
  return code;
}
#define HSL_CUDA(call)                                                            \
  do {                                                                            \
    cudaError_t e_ = (call);                                                      \
    if (e_ != cudaSuccess) return set_err(HSL_ERR_CUDA, "CUDA: %s", cudaGetErrorString(e_)); \
  } while (0)

struct DevBuf {  // grow-only device buffer
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t need(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {  // grow-only pinned host buffer
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t need(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMallocHost(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};
}  // namespace

struct HslModel {
  HslModel() { peers.n = 0; }
  HslModelPod pod;
  HslSimPod sim;        // constants of the fall sweep (hsl_fall.cu)
  bool sim_ok = false;
  double total_mass;
  int fb = 64, maxreg = 128;
  HslPeerOut peers;      // set for the duration of hsl_eval_gaits_gather: where the finish kernel also stores the costs
  int fall_variant = 1;  // fall sweep: 1 = a warp per world (hsl_fall_warp.cuh), 0 = a thread per world (hsl_fall_world.h)
  int64_t max_slots = (int64_t)1 << 26;  // frame slots per launch (hsl_set_max_slots)
  int64_t launches = 0;
  bool rec_on = false;           // pergensetup::rec_transform_flag
  double rec_R[9], rec_t[3];
  // workspace
  DevBuf cand, ttab, wframe, fmin, fmax, status, params, out4, dump_x, dump_z, dump_tau, dump_q, dump_c, in_a, in_b, stage;
  PinBuf pin_in, pin_out;
  cudaStream_t stream = nullptr;
  // hsl_set_kernel_timing: CUDA events around the three kernels of the last gait-evaluation chunk (on its stream)
  bool timing = false;
  cudaEvent_t tev[4] = {nullptr, nullptr, nullptr, nullptr};
  bool tev_valid = false;
  ~HslModel() {
    for (cudaEvent_t e : tev) if (e) cudaEventDestroy(e);
    DevBuf* all[] = {&cand, &ttab, &wframe, &fmin, &fmax, &status, &params, &out4, &dump_x, &dump_z, &dump_tau, &dump_q, &dump_c, &in_a, &in_b, &stage};
    for (DevBuf* b : all) b->release();
    pin_in.release();
    pin_out.release();
    if (stream) cudaStreamDestroy(stream);
  }
};

extern "C" {

const char* hsl_last_error(void) { return g_err; }

int hsl_set_device(int device) {
  HSL_CUDA(cudaSetDevice(device));
  return HSL_OK;
}
int hsl_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

// Page-locked host memory for the callers' input / output arrays: the host entries copy straight from / into the
// caller's buffers, which runs at the full PCIe rate only when they are pinned.
void* hsl_pinned_alloc(size_t bytes) {
  void* p = nullptr;
  if (bytes == 0 || cudaMallocHost(&p, bytes) != cudaSuccess) {
    set_err(HSL_ERR_CUDA, "CUDA: %s", bytes ? "cudaMallocHost failed" : "zero-size pinned allocation");
    return nullptr;
  }
  return p;
}
void hsl_pinned_free(void* p) { if (p) cudaFreeHost(p); }

int hsl_model_load_xml(const char* xml_path, HslModel** out) {
  if (!xml_path || !out) return set_err(HSL_ERR_ARG, "null argument");
  HslModel* m = new HslModel;
  char err[400] = "";
  int rc = hsl_build_model_pod(xml_path, &m->pod, err, sizeof err);
  if (rc != 0) {
    delete m;
    return set_err(rc, "%s", err);
  }
  m->sim_ok = (hsl_build_sim_pod(xml_path, &m->pod, &m->sim, err, sizeof err) == 0);
  m->total_mass = 0;  // periodic::get_total_mass, periodic.cpp:320-325 (DFS order)
  std::vector<double> mass(m->pod.n, 0.0);
  for (int t = 0; t < m->pod.ntrunk; t++) mass[m->pod.trunk[t].body] = m->pod.trunk[t].mass;
  for (int l = 0; l < m->pod.nf; l++)
    for (int h = 0; h < 3; h++) mass[m->pod.limb[l].h[h].body] = m->pod.limb[l].h[h].mass;
  for (int i = 0; i < m->pod.n; i++) m->total_mass += mass[i];
  // default kernel variant (tools/tune.py, profiles/r01_optimisation_log.md): four-limbed models fit three 32-slot
  // blocks per SM with the plain kernel; six-limbed ones run best with the persistent pipelined kernel, one 64-slot
  // block per SM
  m->fb = (m->pod.nf <= 4) ? 32 : 64;
  m->maxreg = (m->pod.nf <= 4) ? 128 : 1;
  *out = m;
  return HSL_OK;
}
void hsl_model_free(HslModel* m) { delete m; }
int hsl_model_dims(const HslModel* m, int32_t d[6]) {
  if (!m) return set_err(HSL_ERR_ARG, "null model");
  d[0] = m->pod.n; d[1] = m->pod.nf; d[2] = m->pod.nmj; d[3] = m->pod.config_dim; d[4] = m->pod.ntrunk; d[5] = m->pod.lik_index;
  return HSL_OK;
}
double hsl_model_rcap(const HslModel* m) { return m ? m->pod.rcap : 0.0; }
size_t hsl_model_pod(const HslModel* m, void* dst, size_t cap) {
  if (m && dst) memcpy(dst, &m->pod, cap < sizeof(HslModelPod) ? cap : sizeof(HslModelPod));
  return sizeof(HslModelPod);
}
int hsl_model_tables(const HslModel* m, int32_t* parent, int32_t* footis, int32_t* limb_top, double* masses, double* com_offset,
                     double* foot_offset) {
  if (!m) return set_err(HSL_ERR_ARG, "null model");
  const HslModelPod& P = m->pod;
  for (int i = 0; i < P.n; i++) {
    if (parent) parent[i] = P.parent[i];
    if (masses) masses[i] = 0;
  }
  for (int t = 0; t < P.ntrunk; t++) {
    const int b = P.trunk[t].body;
    if (masses) masses[b] = P.trunk[t].mass;
    if (com_offset) for (int k = 0; k < 3; k++) com_offset[3 * b + k] = P.trunk[t].com[k];
  }
  for (int l = 0; l < P.nf; l++) {
    for (int h = 0; h < 3; h++) {
      const int b = P.limb[l].h[h].body;
      if (masses) masses[b] = P.limb[l].h[h].mass;
      if (com_offset) for (int k = 0; k < 3; k++) com_offset[3 * b + k] = P.limb[l].h[h].com[k];
    }
    if (footis) footis[l] = P.limb[l].h[2].body;
    if (limb_top) limb_top[l] = P.limb[l].h[0].body;
    if (foot_offset) for (int k = 0; k < 3; k++) foot_offset[3 * l + k] = P.limb[l].foot[k];
  }
  return HSL_OK;
}
int hsl_set_tuning(HslModel* m, int fb, int maxreg) {
  if (!m || (fb != 32 && fb != 64) || (maxreg != 1 && (maxreg < 64 || maxreg > 255))) return set_err(HSL_ERR_ARG, "frame slots per block must be 32 or 64; register cap 64..255, or 1 for the pipelined kernel");
  m->fb = fb;
  m->maxreg = maxreg;
  return HSL_OK;
}
int hsl_get_tuning(const HslModel* m, int* fb, int* maxreg) {
  if (!m) return set_err(HSL_ERR_ARG, "null model");
  if (fb) *fb = m->fb;
  if (maxreg) *maxreg = m->maxreg;
  return HSL_OK;
}
int hsl_set_fall_variant(HslModel* m, int variant) {
  if (!m || (variant != 0 && variant != 1)) return set_err(HSL_ERR_ARG, "fall sweep kernel: 0 (a thread per world) or 1 (a warp per world)");
  m->fall_variant = variant;
  return HSL_OK;
}
int hsl_set_max_slots(HslModel* m, int64_t max_slots) {
  if (!m || max_slots < 5 || max_slots > 0x7fffffff) return set_err(HSL_ERR_ARG, "frame slots per launch must be in 5 .. 2^31 - 1");
  m->max_slots = max_slots;
  return HSL_OK;
}
int64_t hsl_launch_count(const HslModel* m) { return m ? m->launches : 0; }
int hsl_set_kernel_timing(HslModel* m, int on) {
  if (!m) return set_err(HSL_ERR_ARG, "null model");
  m->timing = (on != 0);
  m->tev_valid = false;
  return HSL_OK;
}
int hsl_last_kernel_ms(HslModel* m, float ms[3]) {
  if (!m || !ms) return set_err(HSL_ERR_ARG, "null argument");
  if (!m->tev_valid) return set_err(HSL_ERR_ARG, "no timed gait evaluation on this handle (hsl_set_kernel_timing)");
  HSL_CUDA(cudaEventSynchronize(m->tev[3]));
  for (int k = 0; k < 3; k++) HSL_CUDA(cudaEventElapsedTime(&ms[k], m->tev[k], m->tev[k + 1]));
  return HSL_OK;
}
int hsl_set_rec_transform(HslModel* m, const double* transl, const double* eas) {
  if (!m) return set_err(HSL_ERR_ARG, "null model");
  if (!transl && !eas) { m->rec_on = false; return HSL_OK; }
  const double z3[3] = {0, 0, 0};
  const double* e = eas ? eas : z3;
  const double* t = transl ? transl : z3;
  euler_to_R(e[0], e[1], e[2], m->rec_R);  // affine_from_orientation, pergen.cpp:316-319
  for (int k = 0; k < 3; k++) m->rec_t[k] = t[k];
  m->rec_on = true;
  return HSL_OK;
}

// ---------------------------------------------------------------- device-pointer entry
// One launch sequence (setup, frames, finish) over C candidates whose C * (n_t + 4) frame slots fit the kernels'
// 32-bit slot arithmetic and the workspace bound; eval_gaits_dev below splits larger batches.
static int eval_gaits_chunk(HslModel* m, int64_t C, int n_t, const double* d_params, int flags, double* d_cot, double* d_work,
                            double* d_min, double* d_max, int32_t* d_status, bool dump, cudaStream_t st) {
  const int64_t nfr = C * n_t;
  HSL_CUDA(m->cand.need(sizeof(HslCand) * C));
  HSL_CUDA(m->ttab.need(sizeof(double) * C * (n_t + 4)));
  HSL_CUDA(m->wframe.need(sizeof(double) * nfr));
  HSL_CUDA(m->fmin.need(sizeof(double) * nfr));
  HSL_CUDA(m->fmax.need(sizeof(double) * nfr));
  int32_t* st_buf = d_status;
  if (!st_buf) { HSL_CUDA(m->status.need(sizeof(int32_t) * C)); st_buf = (int32_t*)m->status.p; }
  HslFrameArgs A;
  memset(&A, 0, sizeof A);
  A.n_cand = C; A.n_t = n_t; A.flags = flags; A.n_frames = nfr;
  A.cand = (const HslCand*)m->cand.p; A.ttab = (const double*)m->ttab.p;
  A.wframe = (double*)m->wframe.p; A.fmin_cfz = (double*)m->fmin.p; A.fmax_mu = (double*)m->fmax.p;
  A.status = st_buf;
  if (m->rec_on) {
    A.flags |= HSL_FLAG_REC_TRANSFORM;
    memcpy(A.rec_R, m->rec_R, sizeof A.rec_R);
    memcpy(A.rec_t, m->rec_t, sizeof A.rec_t);
  } else {
    A.flags &= ~HSL_FLAG_REC_TRANSFORM;
  }
  if (dump) {
    const HslModelPod& P = m->pod;
    HSL_CUDA(m->dump_x.need(sizeof(double) * 6 * P.n * nfr));
    HSL_CUDA(m->dump_z.need(sizeof(double) * 3 * P.nf * nfr));
    HSL_CUDA(m->dump_tau.need(sizeof(double) * P.nmj * nfr));
    HSL_CUDA(m->dump_q.need(sizeof(double) * P.config_dim * C * (n_t + 4)));
    HSL_CUDA(m->dump_c.need((size_t)P.nf * nfr));
    A.x = (double*)m->dump_x.p; A.z = (double*)m->dump_z.p; A.tau = (double*)m->dump_tau.p;
    A.q_out = (double*)m->dump_q.p; A.contacts = (uint8_t*)m->dump_c.p;
  }
#ifdef HSL_PHASE_CLOCKS
  const int fbp = (dump ? 32 : m->fb), nroles = m->pod.nf + ((m->maxreg == 1 && !dump) ? 2 : 1), warps = nroles * fbp / 32;
  const int64_t nblk = (C * (n_t + 4) - 4 + (fbp - 4) - 1) / (fbp - 4);
  static DevBuf clkbuf;
  HSL_CUDA(clkbuf.need(sizeof(long long) * nblk * warps * 8));
  A.phase_clk = (long long*)clkbuf.p;
#endif
  if (m->timing) {
    for (cudaEvent_t& e : m->tev) if (!e) HSL_CUDA(cudaEventCreate(&e));
    HSL_CUDA(cudaEventRecord(m->tev[0], st));
  }
  HSL_CUDA(hsl_launch_setup(m->pod, C, n_t, d_params, (HslCand*)m->cand.p, (double*)m->ttab.p, st_buf, st));
  if (m->timing) HSL_CUDA(cudaEventRecord(m->tev[1], st));
  HSL_CUDA(hsl_launch_frames(m->pod, A, HSL_MODE_GAIT, dump, m->fb, m->maxreg, st));
  if (m->timing) HSL_CUDA(cudaEventRecord(m->tev[2], st));
#ifdef HSL_PHASE_CLOCKS
  {
    HSL_CUDA(cudaStreamSynchronize(st));
    std::vector<long long> h((size_t)nblk * warps * 8);
    HSL_CUDA(cudaMemcpy(h.data(), clkbuf.p, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    const int roles = nroles, wpr = fbp / 32;
    if (m->maxreg == 1 && !dump) {  // pipelined kernel: per-tile averages of [work before barrier 1 | wait | phase B | wait]
      const int64_t grid = nblk < 148 * (fbp == 64 ? 1 : 2) ? nblk : 148 * (fbp == 64 ? 1 : 2);
      fprintf(stderr, "[pipe clocks] tiles=%lld grid=%lld\n", (long long)nblk, (long long)grid);
      for (int r = 0; r < roles; r++) {
        double seg[4] = {0, 0, 0, 0}, tiles = 0;
        for (int64_t b = 0; b < grid; b++)
          for (int w = 0; w < wpr; w++) {
            const long long* c = &h[((size_t)b * warps + r * wpr + w) * 8];
            for (int k = 0; k < 4; k++) seg[k] += (double)c[k];
            tiles += (double)c[4];
          }
        fprintf(stderr, "  role %d per tile: pre-B1 %.0f | wait %.0f | B %.0f | wait %.0f\n", r, seg[0] / tiles, seg[1] / tiles, seg[2] / tiles, seg[3] / tiles);
      }
    } else {
      fprintf(stderr, "[phase clocks] blocks=%lld warps/block=%d\n", (long long)nblk, warps);
      for (int r = 0; r < roles; r++) {
        double seg[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int64_t b = 0; b < nblk; b++)
          for (int w = 0; w < wpr; w++) {
            const long long* c = &h[((size_t)b * warps + r * wpr + w) * 8];
            for (int k = 0; k < 7; k++) seg[k] += (double)(c[k + 1] - c[k]);
          }
        fprintf(stderr, "  role %d: A %.0f | w %.0f | B %.0f | w %.0f | C %.0f | w %.0f | DE %.0f  (cycles)\n", r, seg[0] / (nblk * wpr),
                seg[1] / (nblk * wpr), seg[2] / (nblk * wpr), seg[3] / (nblk * wpr), seg[4] / (nblk * wpr), seg[5] / (nblk * wpr), seg[6] / (nblk * wpr));
      }
    }
  }
#endif
  HSL_CUDA(hsl_launch_finish(C, n_t, m->total_mass, (const HslCand*)m->cand.p, nullptr, A.wframe, A.fmin_cfz, A.fmax_mu, st_buf,
                             d_cot, d_work, d_min, d_max, st, m->peers.n ? &m->peers : nullptr));
  if (m->timing) { HSL_CUDA(cudaEventRecord(m->tev[3], st)); m->tev_valid = true; }
  m->launches += 3;
  return HSL_OK;
}

// Candidates are independent, so a batch of any size is evaluated as consecutive chunks of at most max_slots frame
// slots: the per-frame workspace (3 doubles per solved frame + the time table) stays bounded however many candidates
// the caller hands over, and slot indices stay below 2^31.  Per-frame dumps are sized by the caller's own output
// arrays and are not split.
// Candidates per launch; argument errors only (no CUDA call), so the host entries can run it first.
static int gait_chunk_candidates(const HslModel* m, int64_t C, int n_t, bool dump, int64_t* cmax) {
  if (!m || C < 1 || n_t < 1) return set_err(HSL_ERR_ARG, "bad argument");
  const int64_t per_cand = (int64_t)n_t + 4, limit = dump ? (int64_t)0x7fffffff : m->max_slots;
  *cmax = limit / per_cand;
  if (*cmax < 1) return set_err(HSL_ERR_ARG, "n_t + 4 frame slots of one candidate exceed the launch bound (hsl_set_max_slots)");
  if (dump && C > *cmax) return set_err(HSL_ERR_ARG, "per-frame dump of more than 2^31 frame slots: split the batch");
  return HSL_OK;
}

static int eval_gaits_dev(HslModel* m, int64_t C, int n_t, const double* d_params, int flags, double* d_cot, double* d_work,
                          double* d_min, double* d_max, int32_t* d_status, bool dump, cudaStream_t st) {
  if (!d_params) return set_err(HSL_ERR_ARG, "bad argument");
  int64_t cmax = 0;
  const int rcs = gait_chunk_candidates(m, C, n_t, dump, &cmax);
  if (rcs) return rcs;
  if (!d_status && C > cmax) {  // the shared status workspace must cover the whole batch before chunks index into it
    HSL_CUDA(m->status.need(sizeof(int32_t) * C));
    d_status = (int32_t*)m->status.p;
  }
  for (int64_t c0 = 0; c0 < C; c0 += cmax) {
    const int64_t cc = (C - c0 < cmax) ? C - c0 : cmax;
    m->peers.signal = (c0 + cc >= C);   // the finish kernel of the last chunk completes this rank's segment of the gather
    const int rc = eval_gaits_chunk(m, cc, n_t, d_params + c0 * HSL_NPARAM, flags, d_cot ? d_cot + c0 : nullptr, d_work ? d_work + c0 : nullptr,
                                    d_min ? d_min + c0 : nullptr, d_max ? d_max + c0 : nullptr, d_status ? d_status + c0 : nullptr, dump, st);
    if (rc) return rc;
    if (m->peers.n) {  // the next chunk's place in the gather buffers
      for (int r = 0; r < m->peers.n; r++) { m->peers.cot[r] += cc; m->peers.status[r] += cc; }
      m->peers.pad_lo -= cc; m->peers.pad_hi -= cc;
    }
  }
  return HSL_OK;
}

int hsl_eval_gaits(HslModel* m, int64_t n_cand, int n_t, const double* d_params, int flags, double* d_cot, double* d_work,
                   double* d_min_cfz, double* d_max_mu, int32_t* d_status, void* stream) {
  return eval_gaits_dev(m, n_cand, n_t, d_params, flags, d_cot, d_work, d_min_cfz, d_max_mu, d_status, false, (cudaStream_t)stream);
}

int hsl_select_best(const double* d_cost, int64_t n, int64_t* d_index, double* d_value, void* stream) {
  if (!d_cost || n < 1) return set_err(HSL_ERR_ARG, "bad argument");
  HSL_CUDA(hsl_launch_argmin(d_cost, n, d_index, d_value, (cudaStream_t)stream));
  return HSL_OK;
}

int hsl_select_topk(const double* d_cost, int64_t n, int k, int64_t* d_index, double* d_value, void* stream) {
  if (!d_cost || n < 1 || n > 0x7fffffff || k < 1 || k > n || (!d_index && !d_value)) return set_err(HSL_ERR_ARG, "bad argument (1 <= k <= n < 2^31)");
  HSL_CUDA(hsl_launch_topk(d_cost, n, k, d_index, d_value, (cudaStream_t)stream));
  return HSL_OK;
}

// ---------------------------------------------------------------- peer-memory all-gather of the costs (hsl_gather.cu)
struct HslGather {
  int nranks = 0, rank = 0;
  int64_t per = 0;
  size_t off_status = 0, off_flags = 0, bytes = 0;
  char* local = nullptr;
  char* peer[HSL_MAX_PEERS] = {};   // peer[rank] == local
  bool connected = false;
  unsigned long long epoch = 0;
};
static size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

int hsl_gather_create(int nranks, int rank, int64_t n_per_rank, HslGather** out, HslIpcHandle* mine) {
  if (!out || !mine || nranks < 1 || nranks > HSL_MAX_PEERS || rank < 0 || rank >= nranks || n_per_rank < 1)
    return set_err(HSL_ERR_ARG, "bad argument (1 <= nranks <= 16, 0 <= rank < nranks, n_per_rank >= 1)");
  static_assert(sizeof(HslIpcHandle) == sizeof(cudaIpcMemHandle_t), "HslIpcHandle carries a cudaIpcMemHandle_t");
  HslGather* g = new HslGather;
  g->nranks = nranks; g->rank = rank; g->per = n_per_rank;
  const size_t n = (size_t)nranks * n_per_rank;
  g->off_status = up256(2 * n * sizeof(double));
  g->off_flags = g->off_status + up256(2 * n * sizeof(int32_t));
  g->bytes = g->off_flags + up256((HSL_MAX_PEERS + 2) * sizeof(unsigned long long));   // flags [16], the finish kernel's block counter, the time-out word
  cudaError_t e = cudaMalloc((void**)&g->local, g->bytes);   // plain cudaMalloc: stream-ordered pool memory cannot be exported
  if (e == cudaSuccess) e = cudaMemset(g->local, 0xff, g->off_status);                       // costs: NaN
  if (e == cudaSuccess) e = cudaMemset(g->local + g->off_status, 0, g->bytes - g->off_status);  // status and flags: 0
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess && nranks > 1) e = cudaIpcGetMemHandle((cudaIpcMemHandle_t*)mine, g->local);
  if (e != cudaSuccess) { if (g->local) cudaFree(g->local); delete g; return set_err(HSL_ERR_CUDA, "CUDA: %s", cudaGetErrorString(e)); }
  g->peer[rank] = g->local;
  g->connected = (nranks == 1);
  *out = g;
  return HSL_OK;
}
int hsl_gather_connect(HslGather* g, const HslIpcHandle* all) {
  if (!g || (!all && g->nranks > 1)) return set_err(HSL_ERR_ARG, "bad argument");
  if (g->connected) return HSL_OK;
  for (int r = 0; r < g->nranks; r++) {
    if (r == g->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, &all[r], sizeof h);
    cudaError_t e = cudaIpcOpenMemHandle((void**)&g->peer[r], h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return set_err(HSL_ERR_CUDA, ("cudaIpcOpenMemHandle of rank " + std::to_string(r) + "'s buffer: %s").c_str(), cudaGetErrorString(e));
  }
  g->connected = true;
  return HSL_OK;
}
int hsl_gather_check(HslGather* g) {
  if (!g) return set_err(HSL_ERR_ARG, "null gather object");
  unsigned long long t = 0;
  HSL_CUDA(cudaDeviceSynchronize());
  HSL_CUDA(cudaMemcpy(&t, g->local + g->off_flags + (HSL_MAX_PEERS + 1) * sizeof(unsigned long long), sizeof t, cudaMemcpyDeviceToHost));
  if (t) return set_err(HSL_ERR_CUDA, ("a rank's costs did not arrive within the time-out (30 s by default) at gather call " + std::to_string(t) + ": %s").c_str(),
                        "the ranks of the job must make the same sequence of scatter calls");
  return HSL_OK;
}
int64_t hsl_gather_size(const HslGather* g) { return g ? (int64_t)g->nranks * g->per : 0; }
int hsl_gather_free(HslGather* g) {
  if (!g) return HSL_OK;
  cudaDeviceSynchronize();
  for (int r = 0; r < g->nranks; r++) if (r != g->rank && g->peer[r]) cudaIpcCloseMemHandle(g->peer[r]);
  if (g->local) cudaFree(g->local);
  delete g;
  return HSL_OK;
}
// evaluation whose finish kernel stores costs and status into every rank's gather buffer and raises this rank's flags
int hsl_eval_gaits_scatter(HslModel* m, HslGather* g, int64_t n_cand, int n_t, const double* d_params, int flags, double* d_cot, double* d_work,
                           double* d_min_cfz, double* d_max_mu, int32_t* d_status, void* stream) {
  if (!m || !g || !g->connected || n_cand < 0 || n_cand > g->per || n_t < 1 || (n_cand > 0 && !d_params))
    return set_err(HSL_ERR_ARG, "bad argument (connected gather object, 0 <= n_cand <= n_per_rank)");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned long long epoch = ++g->epoch;
  const size_t n = (size_t)g->nranks * g->per, par = (size_t)(epoch & 1);
  HslPeerOut po;
  memset(&po, 0, sizeof po);
  po.n = g->nranks;
  po.epoch = epoch;
  po.ticket = (unsigned int*)(g->local + g->off_flags + HSL_MAX_PEERS * sizeof(unsigned long long));
  for (int r = 0; r < g->nranks; r++) {
    po.cot[r] = (double*)g->peer[r] + par * n + (size_t)g->rank * g->per;
    po.status[r] = (int32_t*)(g->peer[r] + g->off_status) + par * n + (size_t)g->rank * g->per;
    po.flag[r] = (unsigned long long*)(g->peer[r] + g->off_flags) + g->rank;
  }
  po.pad_lo = n_cand; po.pad_hi = g->per;   // eval_gaits_dev rebases them per chunk and sets `signal` on the last one
  if (n_cand == 0) {
    po.signal = 1;
    HSL_CUDA(hsl_launch_gather_signal(po, st));
    m->launches += 1;
    return HSL_OK;
  }
  m->peers = po;
  const int rc = eval_gaits_dev(m, n_cand, n_t, d_params, flags, d_cot, d_work, d_min_cfz, d_max_mu, d_status, false, st);
  m->peers.n = 0;
  return rc;
}
static void gather_views(const HslGather* g, const double** d_all_cot, const int32_t** d_all_status) {
  const size_t n = (size_t)g->nranks * g->per, par = (size_t)(g->epoch & 1);
  if (d_all_cot) *d_all_cot = (const double*)g->local + par * n;
  if (d_all_status) *d_all_status = (const int32_t*)(g->local + g->off_status) + par * n;
}
int hsl_gather_wait(HslGather* g, const double** d_all_cot, const int32_t** d_all_status, void* stream) {
  if (!g || !g->connected || g->epoch == 0) return set_err(HSL_ERR_ARG, "no scatter on this gather object yet");
  HSL_CUDA(hsl_launch_gather_wait((const unsigned long long*)(g->local + g->off_flags), g->nranks, g->epoch, (cudaStream_t)stream));
  gather_views(g, d_all_cot, d_all_status);
  return HSL_OK;
}
int hsl_gather_select_best(HslGather* g, int64_t* d_index, double* d_value, void* stream) {
  if (!g || !g->connected || g->epoch == 0 || (!d_index && !d_value)) return set_err(HSL_ERR_ARG, "no scatter on this gather object yet, or no output");
  const double* all = nullptr;
  gather_views(g, &all, nullptr);
  HSL_CUDA(hsl_launch_argmin_gathered(all, (int64_t)g->nranks * g->per, d_index, d_value, (const unsigned long long*)(g->local + g->off_flags),
                                      g->nranks, g->epoch, (cudaStream_t)stream));
  return HSL_OK;
}
int hsl_eval_gaits_gather(HslModel* m, HslGather* g, int64_t n_cand, int n_t, const double* d_params, int flags, double* d_cot, double* d_work,
                          double* d_min_cfz, double* d_max_mu, int32_t* d_status, const double** d_all_cot, const int32_t** d_all_status,
                          void* stream) {
  const int rc = hsl_eval_gaits_scatter(m, g, n_cand, n_t, d_params, flags, d_cot, d_work, d_min_cfz, d_max_mu, d_status, stream);
  if (rc) return rc;
  if (m) m->launches += 1;
  return hsl_gather_wait(g, d_all_cot, d_all_status, stream);
}

static int ensure_stream(HslModel* m);

static int ensure_stream(HslModel* m) {
  if (!m->stream) HSL_CUDA(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  return HSL_OK;
}

// Component-major device dump [comp][nfr] -> the caller's row-major [nfr][comp] host array: transposed on the device
// (hsl_transpose_kernel), then one D2H copy straight into the caller's buffer (full PCIe rate when it is pinned).
static int fetch_transposed(HslModel* m, const void* dsrc, int comps, int64_t nfr, void* dst, cudaStream_t st, int elem_size = 8) {
  if (!dst) return HSL_OK;
  const size_t bytes = (size_t)elem_size * comps * nfr;
  HSL_CUDA(m->stage.need(bytes));
  HSL_CUDA(hsl_launch_transpose(dsrc, m->stage.p, comps, nfr, elem_size, st));
  m->launches += 1;
  HSL_CUDA(cudaMemcpyAsync(dst, m->stage.p, bytes, cudaMemcpyDeviceToHost, st));
  HSL_CUDA(cudaStreamSynchronize(st));
  return HSL_OK;
}

static bool host_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

static int eval_gaits_host_impl(HslModel* m, int64_t C, int n_t, const double* params, int flags, double* cot, double* work,
                                double* min_cfz, double* max_mu, int32_t* status, bool dump, double* traj, double* x, double* z,
                                double* tau, uint8_t* contacts) {
  if (!m || C < 1 || n_t < 1 || !params) return set_err(HSL_ERR_ARG, "bad argument");
  int64_t cmax = 0;
  int rc = gait_chunk_candidates(m, C, n_t, dump, &cmax);
  if (rc) return rc;
  if ((rc = ensure_stream(m))) return rc;
  cudaStream_t st = m->stream;
  const size_t pbytes = sizeof(double) * HSL_NPARAM * C;
  HSL_CUDA(m->params.need(pbytes));
  HSL_CUDA(m->out4.need(sizeof(double) * 4 * C + sizeof(int32_t) * C));
  const void* src = params;
  if (!host_pinned(params)) {   // pageable input: stage it (a page-locked array, e.g. from hsl_pinned_alloc, is copied from where it lies)
    HSL_CUDA(m->pin_in.need(pbytes));
    memcpy(m->pin_in.p, params, pbytes);
    src = m->pin_in.p;
  }
  HSL_CUDA(cudaMemcpyAsync(m->params.p, src, pbytes, cudaMemcpyHostToDevice, st));
  double* d4 = (double*)m->out4.p;
  int32_t* dst = (int32_t*)(d4 + 4 * C);
  rc = eval_gaits_dev(m, C, n_t, (const double*)m->params.p, flags, d4, d4 + C, d4 + 2 * C, d4 + 3 * C, dst, dump, st);
  if (rc) return rc;
  const size_t obytes = sizeof(double) * 4 * C + sizeof(int32_t) * C;
  HSL_CUDA(m->pin_out.need(obytes));
  HSL_CUDA(cudaMemcpyAsync(m->pin_out.p, d4, obytes, cudaMemcpyDeviceToHost, st));
  HSL_CUDA(cudaStreamSynchronize(st));
  const double* h4 = (const double*)m->pin_out.p;
  if (cot) memcpy(cot, h4, sizeof(double) * C);
  if (work) memcpy(work, h4 + C, sizeof(double) * C);
  if (min_cfz) memcpy(min_cfz, h4 + 2 * C, sizeof(double) * C);
  if (max_mu) memcpy(max_mu, h4 + 3 * C, sizeof(double) * C);
  if (status) memcpy(status, h4 + 4 * C, sizeof(int32_t) * C);
  if (dump) {
    const HslModelPod& P = m->pod;
    const int64_t nfr = C * n_t;
    if ((rc = fetch_transposed(m, m->dump_x.p, 6 * P.n, nfr, x, st))) return rc;
    if ((rc = fetch_transposed(m, m->dump_z.p, 3 * P.nf, nfr, z, st))) return rc;
    if ((rc = fetch_transposed(m, m->dump_tau.p, P.nmj, nfr, tau, st))) return rc;
    if ((rc = fetch_transposed(m, m->dump_q.p, P.config_dim, C * (n_t + 4), traj, st))) return rc;
    if ((rc = fetch_transposed(m, m->dump_c.p, P.nf, nfr, contacts, st, 1))) return rc;
  }
  return HSL_OK;
}

// HOST arrays in, the gathered HOST arrays out: for hosts that keep no device memory of their own (the C++ mirror's sharded
// measure_cot_sweep).  Synchronous.
int hsl_eval_gaits_gather_host(HslModel* m, HslGather* g, int64_t n_cand, int n_t, const double* params, int flags, double* all_cot,
                               int32_t* all_status) {
  if (!m || !g || n_cand < 0 || (n_cand > 0 && !params)) return set_err(HSL_ERR_ARG, "bad argument");
  int rc = ensure_stream(m);
  if (rc) return rc;
  cudaStream_t st = m->stream;
  if (n_cand > 0) {
    const size_t pbytes = sizeof(double) * HSL_NPARAM * n_cand;
    HSL_CUDA(m->params.need(pbytes));
    HSL_CUDA(cudaMemcpyAsync(m->params.p, params, pbytes, cudaMemcpyHostToDevice, st));
  }
  const double* d_all = nullptr;
  const int32_t* d_st = nullptr;
  rc = hsl_eval_gaits_gather(m, g, n_cand, n_t, (const double*)m->params.p, flags, nullptr, nullptr, nullptr, nullptr, nullptr, &d_all, &d_st, st);
  if (rc) return rc;
  const size_t n = (size_t)hsl_gather_size(g);
  if (all_cot) HSL_CUDA(cudaMemcpyAsync(all_cot, d_all, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  if (all_status) HSL_CUDA(cudaMemcpyAsync(all_status, d_st, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
  HSL_CUDA(cudaStreamSynchronize(st));
  return HSL_OK;
}

int hsl_eval_gaits_host(HslModel* m, int64_t n_cand, int n_t, const double* params, int flags, double* cot, double* work,
                        double* min_cfz, double* max_mu, int32_t* status) {
  return eval_gaits_host_impl(m, n_cand, n_t, params, flags, cot, work, min_cfz, max_mu, status, false, nullptr, nullptr, nullptr,
                              nullptr, nullptr);
}
int hsl_eval_gaits_detail_host(HslModel* m, int64_t n_cand, int n_t, const double* params, int flags, double* cot, double* work,
                               double* min_cfz, double* max_mu, int32_t* status, double* traj, double* x, double* z, double* tau,
                               uint8_t* contacts) {
  return eval_gaits_host_impl(m, n_cand, n_t, params, flags, cot, work, min_cfz, max_mu, status, true, traj, x, z, tau, contacts);
}

// L2: supplied joint trajectories.  Launch part shared by the device-pointer and the host entry: per-frame dumps go to
// the component-major workspace (dump_x / dump_z / dump_tau), the per-candidate results to out4 = [cot | work | min | max].
static int eval_trajectories_launch(HslModel* m, int64_t C, int n_t, const double* d_traj, const double* d_dt, int32_t* d_status,
                                    bool dumps, cudaStream_t st) {
  const HslModelPod& P = m->pod;
  const int64_t nfr = C * n_t;
  HSL_CUDA(m->wframe.need(sizeof(double) * nfr));
  HSL_CUDA(m->fmin.need(sizeof(double) * nfr));
  HSL_CUDA(m->fmax.need(sizeof(double) * nfr));
  HSL_CUDA(cudaMemsetAsync(d_status, 0, sizeof(int32_t) * C, st));
  HSL_CUDA(m->out4.need(sizeof(double) * 4 * C));
  HslFrameArgs A;
  memset(&A, 0, sizeof A);
  A.n_cand = C; A.n_t = n_t; A.n_frames = nfr;
  A.traj = d_traj; A.dt_in = d_dt;
  A.wframe = (double*)m->wframe.p; A.fmin_cfz = (double*)m->fmin.p; A.fmax_mu = (double*)m->fmax.p;
  A.status = d_status;
  if (dumps) {
    HSL_CUDA(m->dump_x.need(sizeof(double) * 6 * P.n * nfr));
    HSL_CUDA(m->dump_z.need(sizeof(double) * 3 * P.nf * nfr));
    HSL_CUDA(m->dump_tau.need(sizeof(double) * P.nmj * nfr));
    A.x = (double*)m->dump_x.p; A.z = (double*)m->dump_z.p; A.tau = (double*)m->dump_tau.p;
  }
  HSL_CUDA(hsl_launch_frames(P, A, HSL_MODE_TRAJ, true, 32, 1, st));
  double* d4 = (double*)m->out4.p;
  HSL_CUDA(hsl_launch_finish(C, n_t, m->total_mass, nullptr, A.dt_in, A.wframe, A.fmin_cfz, A.fmax_mu, A.status, nullptr, d4 + C,
                             d4 + 2 * C, d4 + 3 * C, st));
  m->launches += 2;
  return HSL_OK;
}

// [comps][nfr] workspace -> the caller's row-major [nfr][comps] DEVICE array
static int transpose_to(HslModel* m, const void* dsrc, int comps, int64_t nfr, void* d_dst, cudaStream_t st) {
  if (!d_dst) return HSL_OK;
  HSL_CUDA(hsl_launch_transpose(dsrc, d_dst, comps, nfr, 8, st));
  m->launches += 1;
  return HSL_OK;
}

int hsl_eval_trajectories(HslModel* m, int64_t C, int n_t, const double* d_traj, const double* d_dt, double* d_work, double* d_min_cfz,
                          double* d_max_mu, int32_t* d_status, double* d_x, double* d_z, double* d_tau, void* stream) {
  if (!m || C < 1 || n_t < 1 || !d_traj || !d_dt) return set_err(HSL_ERR_ARG, "bad argument");
  if (C * ((int64_t)n_t + 5) > 0x7fffffff) return set_err(HSL_ERR_ARG, "more than 2^31 frame slots in one call: split the batch");
  cudaStream_t st = (cudaStream_t)stream;
  const HslModelPod& P = m->pod;
  const int64_t nfr = C * n_t;
  int32_t* stp = d_status;
  if (!stp) { HSL_CUDA(m->status.need(sizeof(int32_t) * C)); stp = (int32_t*)m->status.p; }
  int rc = eval_trajectories_launch(m, C, n_t, d_traj, d_dt, stp, d_x || d_z || d_tau, st);
  if (rc) return rc;
  const double* d4 = (const double*)m->out4.p;
  if (d_work) HSL_CUDA(cudaMemcpyAsync(d_work, d4 + C, sizeof(double) * C, cudaMemcpyDeviceToDevice, st));
  if (d_min_cfz) HSL_CUDA(cudaMemcpyAsync(d_min_cfz, d4 + 2 * C, sizeof(double) * C, cudaMemcpyDeviceToDevice, st));
  if (d_max_mu) HSL_CUDA(cudaMemcpyAsync(d_max_mu, d4 + 3 * C, sizeof(double) * C, cudaMemcpyDeviceToDevice, st));
  if ((rc = transpose_to(m, m->dump_x.p, 6 * P.n, nfr, d_x, st))) return rc;
  if ((rc = transpose_to(m, m->dump_z.p, 3 * P.nf, nfr, d_z, st))) return rc;
  if ((rc = transpose_to(m, m->dump_tau.p, P.nmj, nfr, d_tau, st))) return rc;
  return HSL_OK;
}

int hsl_eval_trajectories_host(HslModel* m, int64_t C, int n_t, const double* traj, const double* dt, double* work, double* min_cfz,
                               double* max_mu, int32_t* status, double* x, double* z, double* tau) {
  if (!m || C < 1 || n_t < 1 || !traj || !dt) return set_err(HSL_ERR_ARG, "bad argument");
  if (C * ((int64_t)n_t + 5) > 0x7fffffff) return set_err(HSL_ERR_ARG, "more than 2^31 frame slots in one call: split the batch");
  int rc = ensure_stream(m);
  if (rc) return rc;
  cudaStream_t st = m->stream;
  const HslModelPod& P = m->pod;
  const int64_t nfr = C * n_t;
  const size_t tbytes = sizeof(double) * C * (n_t + 5) * P.config_dim;
  HSL_CUDA(m->in_a.need(tbytes));
  HSL_CUDA(m->in_b.need(sizeof(double) * C));
  HSL_CUDA(cudaMemcpyAsync(m->in_a.p, traj, tbytes, cudaMemcpyHostToDevice, st));
  HSL_CUDA(cudaMemcpyAsync(m->in_b.p, dt, sizeof(double) * C, cudaMemcpyHostToDevice, st));
  HSL_CUDA(m->status.need(sizeof(int32_t) * C));
  if ((rc = eval_trajectories_launch(m, C, n_t, (const double*)m->in_a.p, (const double*)m->in_b.p, (int32_t*)m->status.p, true, st))) return rc;
  double* d4 = (double*)m->out4.p;
  HSL_CUDA(m->pin_in.need(sizeof(double) * 4 * C + sizeof(int32_t) * C));
  double* h4 = (double*)m->pin_in.p;
  HSL_CUDA(cudaMemcpyAsync(h4, d4, sizeof(double) * 4 * C, cudaMemcpyDeviceToHost, st));
  HSL_CUDA(cudaMemcpyAsync(h4 + 4 * C, m->status.p, sizeof(int32_t) * C, cudaMemcpyDeviceToHost, st));
  HSL_CUDA(cudaStreamSynchronize(st));
  if (work) memcpy(work, h4 + C, sizeof(double) * C);
  if (min_cfz) memcpy(min_cfz, h4 + 2 * C, sizeof(double) * C);
  if (max_mu) memcpy(max_mu, h4 + 3 * C, sizeof(double) * C);
  if (status) memcpy(status, h4 + 4 * C, sizeof(int32_t) * C);
  if ((rc = fetch_transposed(m, m->dump_x.p, 6 * P.n, nfr, x, st))) return rc;
  if ((rc = fetch_transposed(m, m->dump_z.p, 3 * P.nf, nfr, z, st))) return rc;
  if ((rc = fetch_transposed(m, m->dump_tau.p, P.nmj, nfr, tau, st))) return rc;
  return HSL_OK;
}

// L1: populated dynrecords.  Launch part shared by the device-pointer and the host entry.
static int solve_frames_launch(HslModel* m, int64_t F, const double* d_pos, const double* d_jpos, const double* d_jz,
                               const double* d_momrate, const double* d_angrate, const double* d_fpos, const uint8_t* d_contacts,
                               int32_t* d_status, cudaStream_t st) {
  const HslModelPod& P = m->pod;
  HSL_CUDA(cudaMemsetAsync(d_status, 0, sizeof(int32_t) * F, st));
  HSL_CUDA(m->dump_x.need(sizeof(double) * 6 * P.n * F));
  HSL_CUDA(m->dump_z.need(sizeof(double) * 3 * P.nf * F));
  HSL_CUDA(m->dump_tau.need(sizeof(double) * P.nmj * F));
  HslFrameArgs A;
  memset(&A, 0, sizeof A);
  A.n_cand = F;  // status is per frame in this mode (slot.c is forced to the frame below)
  A.n_t = 1; A.n_frames = F;
  A.f_pos = d_pos; A.f_jpos = d_jpos; A.f_jz = d_jz; A.f_momrate = d_momrate; A.f_angrate = d_angrate;
  A.f_fpos = d_fpos; A.f_contacts = d_contacts;
  A.status = d_status;
  A.x = (double*)m->dump_x.p; A.z = (double*)m->dump_z.p; A.tau = (double*)m->dump_tau.p;
  HSL_CUDA(hsl_launch_frames(P, A, HSL_MODE_FIELDS, true, 32, 1, st));
  m->launches += 1;
  return HSL_OK;
}

int hsl_solve_frames(HslModel* m, int64_t F, const double* d_pos, const double* d_jpos, const double* d_jzaxis, const double* d_mom_rate,
                     const double* d_ang_mom_rate, const double* d_fpos, const uint8_t* d_contacts, double* d_x, double* d_z, double* d_tau,
                     int32_t* d_status, void* stream) {
  if (!m || F < 1 || !d_pos || !d_jpos || !d_jzaxis || !d_mom_rate || !d_ang_mom_rate || !d_fpos || !d_contacts)
    return set_err(HSL_ERR_ARG, "bad argument");
  if (F > 0x7fffffff) return set_err(HSL_ERR_ARG, "more than 2^31 frames in one call: split the batch");
  cudaStream_t st = (cudaStream_t)stream;
  const HslModelPod& P = m->pod;
  int32_t* stp = d_status;
  if (!stp) { HSL_CUDA(m->status.need(sizeof(int32_t) * F)); stp = (int32_t*)m->status.p; }
  int rc = solve_frames_launch(m, F, d_pos, d_jpos, d_jzaxis, d_mom_rate, d_ang_mom_rate, d_fpos, d_contacts, stp, st);
  if (rc) return rc;
  if ((rc = transpose_to(m, m->dump_x.p, 6 * P.n, F, d_x, st))) return rc;
  if ((rc = transpose_to(m, m->dump_z.p, 3 * P.nf, F, d_z, st))) return rc;
  if ((rc = transpose_to(m, m->dump_tau.p, P.nmj, F, d_tau, st))) return rc;
  return HSL_OK;
}

int hsl_solve_frames_host(HslModel* m, int64_t F, const double* pos, const double* jpos, const double* jzaxis, const double* mom_rate,
                          const double* ang_mom_rate, const double* fpos, const uint8_t* contacts, double* x, double* z, double* tau,
                          int32_t* status) {
  if (!m || F < 1 || !pos || !jpos || !jzaxis || !mom_rate || !ang_mom_rate || !fpos || !contacts)
    return set_err(HSL_ERR_ARG, "bad argument");
  if (F > 0x7fffffff) return set_err(HSL_ERR_ARG, "more than 2^31 frames in one call: split the batch");
  int rc = ensure_stream(m);
  if (rc) return rc;
  cudaStream_t st = m->stream;
  const HslModelPod& P = m->pod;
  const size_t nb = sizeof(double) * F * P.n * 3, fb = sizeof(double) * F * P.nf * 3, cb = (size_t)F * P.nf;
  HSL_CUDA(m->in_a.need(5 * nb + fb + cb));
  char* base = (char*)m->in_a.p;
  const void* srcs[5] = {pos, jpos, jzaxis, mom_rate, ang_mom_rate};
  for (int k = 0; k < 5; k++) HSL_CUDA(cudaMemcpyAsync(base + k * nb, srcs[k], nb, cudaMemcpyHostToDevice, st));
  HSL_CUDA(cudaMemcpyAsync(base + 5 * nb, fpos, fb, cudaMemcpyHostToDevice, st));
  HSL_CUDA(cudaMemcpyAsync(base + 5 * nb + fb, contacts, cb, cudaMemcpyHostToDevice, st));
  HSL_CUDA(m->status.need(sizeof(int32_t) * F));
  if ((rc = solve_frames_launch(m, F, (const double*)base, (const double*)(base + nb), (const double*)(base + 2 * nb),
                                (const double*)(base + 3 * nb), (const double*)(base + 4 * nb), (const double*)(base + 5 * nb),
                                (const uint8_t*)(base + 5 * nb + fb), (int32_t*)m->status.p, st))) return rc;
  if (status) {
    HSL_CUDA(cudaMemcpyAsync(status, m->status.p, sizeof(int32_t) * F, cudaMemcpyDeviceToHost, st));
    HSL_CUDA(cudaStreamSynchronize(st));
  }
  if ((rc = fetch_transposed(m, m->dump_x.p, 6 * P.n, F, x, st))) return rc;
  if ((rc = fetch_transposed(m, m->dump_z.p, 3 * P.nf, F, z, st))) return rc;
  if ((rc = fetch_transposed(m, m->dump_tau.p, P.nmj, F, tau, st))) return rc;
  return HSL_OK;
}

// ---------------------------------------------------------------- record-level entries (a2, a3)
int hsl_gait_records_host(HslModel* m, int64_t C, const double* params, int n_times, const double* times, int flags, double* rec,
                          int32_t* status) {
  if (!m || C < 1 || n_times < 1 || !params || !times || !rec) return set_err(HSL_ERR_ARG, "bad argument");
  int rc = ensure_stream(m);
  if (rc) return rc;
  cudaStream_t st = m->stream;
  const HslModelPod& P = m->pod;
  const int rl = 6 + 3 * P.nf;
  const size_t pbytes = sizeof(double) * HSL_NPARAM * C, tbytes = sizeof(double) * n_times, rbytes = sizeof(double) * rl * C * n_times;
  HSL_CUDA(m->params.need(pbytes));
  HSL_CUDA(m->in_b.need(tbytes));
  HSL_CUDA(m->in_a.need(rbytes));
  HSL_CUDA(m->cand.need(sizeof(HslCand) * C));
  HSL_CUDA(m->ttab.need(sizeof(double) * C * 5));
  HSL_CUDA(m->status.need(sizeof(int32_t) * C));
  HSL_CUDA(cudaMemcpyAsync(m->params.p, params, pbytes, cudaMemcpyHostToDevice, st));
  HSL_CUDA(cudaMemcpyAsync(m->in_b.p, times, tbytes, cudaMemcpyHostToDevice, st));
  HslFrameArgs A;
  memset(&A, 0, sizeof A);
  A.n_cand = C; A.n_t = 1; A.flags = flags & ~HSL_FLAG_REC_TRANSFORM;
  A.cand = (const HslCand*)m->cand.p;
  if (m->rec_on) {
    A.flags |= HSL_FLAG_REC_TRANSFORM;
    memcpy(A.rec_R, m->rec_R, sizeof A.rec_R);
    memcpy(A.rec_t, m->rec_t, sizeof A.rec_t);
  }
  // the candidate constants do not depend on n_t except for dt / the time table, which this entry does not use
  HSL_CUDA(hsl_launch_setup(P, C, 1, (const double*)m->params.p, (HslCand*)m->cand.p, (double*)m->ttab.p, (int32_t*)m->status.p, st));
  HSL_CUDA(hsl_launch_gait_records(P, A, n_times, (const double*)m->in_b.p, (double*)m->in_a.p, st));
  m->launches += 2;
  HSL_CUDA(cudaMemcpyAsync(rec, m->in_a.p, rbytes, cudaMemcpyDeviceToHost, st));
  if (status) HSL_CUDA(cudaMemcpyAsync(status, m->status.p, sizeof(int32_t) * C, cudaMemcpyDeviceToHost, st));
  HSL_CUDA(cudaStreamSynchronize(st));
  return HSL_OK;
}

int hsl_ik_records_host(HslModel* m, int64_t n, const double* rec, int flags, double* q, int32_t* status) {
  if (!m || n < 1 || !rec || !q) return set_err(HSL_ERR_ARG, "bad argument");
  int rc = ensure_stream(m);
  if (rc) return rc;
  cudaStream_t st = m->stream;
  const HslModelPod& P = m->pod;
  const size_t rbytes = sizeof(double) * (6 + 3 * P.nf) * n, qbytes = sizeof(double) * P.config_dim * n;
  HSL_CUDA(m->in_a.need(rbytes));
  HSL_CUDA(m->in_b.need(qbytes));
  HSL_CUDA(m->status.need(sizeof(int32_t) * n));
  HSL_CUDA(cudaMemcpyAsync(m->in_a.p, rec, rbytes, cudaMemcpyHostToDevice, st));
  HSL_CUDA(cudaMemsetAsync(m->status.p, 0, sizeof(int32_t) * n, st));
  HSL_CUDA(hsl_launch_ik_records(P, n, flags, (const double*)m->in_a.p, (double*)m->in_b.p, (int32_t*)m->status.p, st));
  m->launches += 1;
  HSL_CUDA(cudaMemcpyAsync(q, m->in_b.p, qbytes, cudaMemcpyDeviceToHost, st));
  if (status) HSL_CUDA(cudaMemcpyAsync(status, m->status.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
  HSL_CUDA(cudaStreamSynchronize(st));
  return HSL_OK;
}

int hsl_fk_records_host(HslModel* m, int64_t n, const double* q, double* A_ground, double* J_A_ground) {
  if (!m || n < 1 || !q || (!A_ground && !J_A_ground)) return set_err(HSL_ERR_ARG, "bad argument");
  int rc = ensure_stream(m);
  if (rc) return rc;
  cudaStream_t st = m->stream;
  const HslModelPod& P = m->pod;
  const size_t qbytes = sizeof(double) * P.config_dim * n, abytes = sizeof(double) * 16 * P.n * n;
  HSL_CUDA(m->in_b.need(qbytes));
  HSL_CUDA(m->in_a.need(2 * abytes));
  HSL_CUDA(cudaMemcpyAsync(m->in_b.p, q, qbytes, cudaMemcpyHostToDevice, st));
  double* dA = (double*)m->in_a.p;
  double* dJ = (double*)((char*)m->in_a.p + abytes);
  HSL_CUDA(hsl_launch_fk_records(P, n, (const double*)m->in_b.p, A_ground ? dA : nullptr, J_A_ground ? dJ : nullptr, st));
  m->launches += 1;
  if (A_ground) HSL_CUDA(cudaMemcpyAsync(A_ground, dA, abytes, cudaMemcpyDeviceToHost, st));
  if (J_A_ground) HSL_CUDA(cudaMemcpyAsync(J_A_ground, dJ, abytes, cudaMemcpyDeviceToHost, st));
  HSL_CUDA(cudaStreamSynchronize(st));
  return HSL_OK;
}

// ---------------------------------------------------------------- forces from torques (hsl_forces.h)
int hsl_solve_forces_host(HslModel* m, int64_t F, const double* pos, const double* jpos, const double* jzaxis, const double* mom_rate,
                          const double* ang_mom_rate, const double* fpos, const double* torques, double* z, int32_t* status) {
  if (!m || F < 1 || !pos || !jpos || !jzaxis || !mom_rate || !ang_mom_rate || !fpos || !torques || !z)
    return set_err(HSL_ERR_ARG, "bad argument");
  if (F > 0x7fffffff) return set_err(HSL_ERR_ARG, "more than 2^31 frames in one call: split the batch");
  int rc = ensure_stream(m);
  if (rc) return rc;
  cudaStream_t st = m->stream;
  const HslModelPod& P = m->pod;
  const size_t nb = sizeof(double) * F * P.n * 3, fb = sizeof(double) * F * P.nf * 3, tb = sizeof(double) * F * P.nmj, cb = (size_t)F * P.nf;
  HSL_CUDA(m->in_a.need(5 * nb + fb + tb + cb));
  char* base = (char*)m->in_a.p;
  const void* srcs[5] = {pos, jpos, jzaxis, mom_rate, ang_mom_rate};
  for (int k = 0; k < 5; k++) HSL_CUDA(cudaMemcpyAsync(base + k * nb, srcs[k], nb, cudaMemcpyHostToDevice, st));
  HSL_CUDA(cudaMemcpyAsync(base + 5 * nb, fpos, fb, cudaMemcpyHostToDevice, st));
  HSL_CUDA(cudaMemcpyAsync(base + 5 * nb + fb, torques, tb, cudaMemcpyHostToDevice, st));
  HSL_CUDA(cudaMemsetAsync(base + 5 * nb + fb + tb, 1, cb, st));  // every foot takes part (ftsolver.cpp:343)
  HSL_CUDA(m->status.need(sizeof(int32_t) * F));
  HSL_CUDA(cudaMemsetAsync(m->status.p, 0, sizeof(int32_t) * F, st));
  HSL_CUDA(m->dump_z.need(sizeof(double) * 3 * P.nf * F));
  HslFrameArgs A;
  memset(&A, 0, sizeof A);
  A.n_cand = F; A.n_t = 1; A.n_frames = F;
  A.f_pos = (const double*)(base); A.f_jpos = (const double*)(base + nb); A.f_jz = (const double*)(base + 2 * nb);
  A.f_momrate = (const double*)(base + 3 * nb); A.f_angrate = (const double*)(base + 4 * nb);
  A.f_fpos = (const double*)(base + 5 * nb); A.tau_in = (const double*)(base + 5 * nb + fb);
  A.f_contacts = (const uint8_t*)(base + 5 * nb + fb + tb);
  A.status = (int32_t*)m->status.p;
  A.z = (double*)m->dump_z.p;
  HSL_CUDA(hsl_launch_forces(P, A, HSL_MODE_FIELDS, st));
  m->launches += 1;
  if (status) {
    HSL_CUDA(cudaMemcpyAsync(status, m->status.p, sizeof(int32_t) * F, cudaMemcpyDeviceToHost, st));
    HSL_CUDA(cudaStreamSynchronize(st));
  }
  return fetch_transposed(m, m->dump_z.p, 3 * P.nf, F, z, st);
}

int hsl_solve_forces_gait_host(HslModel* m, int64_t C, int n_t, const double* params, int flags, const double* torques, double* z,
                               int32_t* status) {
  if (!m || C < 1 || n_t < 1 || !params || !torques || !z) return set_err(HSL_ERR_ARG, "bad argument");
  if (C * ((int64_t)n_t + 5) > 0x7fffffff) return set_err(HSL_ERR_ARG, "more than 2^31 frame slots in one call: split the batch");
  int rc = ensure_stream(m);
  if (rc) return rc;
  cudaStream_t st = m->stream;
  const HslModelPod& P = m->pod;
  const int64_t nfr = C * n_t;
  const size_t pbytes = sizeof(double) * HSL_NPARAM * C, tb = sizeof(double) * nfr * P.nmj;
  HSL_CUDA(m->params.need(pbytes));
  HSL_CUDA(m->in_b.need(tb));
  HSL_CUDA(cudaMemcpyAsync(m->params.p, params, pbytes, cudaMemcpyHostToDevice, st));
  HSL_CUDA(cudaMemcpyAsync(m->in_b.p, torques, tb, cudaMemcpyHostToDevice, st));
  HSL_CUDA(m->cand.need(sizeof(HslCand) * C));
  HSL_CUDA(m->ttab.need(sizeof(double) * C * (n_t + 4)));
  HSL_CUDA(m->status.need(sizeof(int32_t) * C));
  HSL_CUDA(m->dump_z.need(sizeof(double) * 3 * P.nf * nfr));
  HslFrameArgs A;
  memset(&A, 0, sizeof A);
  A.n_cand = C; A.n_t = n_t; A.flags = flags & ~HSL_FLAG_REC_TRANSFORM; A.n_frames = nfr;
  A.cand = (const HslCand*)m->cand.p; A.ttab = (const double*)m->ttab.p;
  A.status = (int32_t*)m->status.p;
  A.tau_in = (const double*)m->in_b.p;
  A.z = (double*)m->dump_z.p;
  if (m->rec_on) {
    A.flags |= HSL_FLAG_REC_TRANSFORM;
    memcpy(A.rec_R, m->rec_R, sizeof A.rec_R);
    memcpy(A.rec_t, m->rec_t, sizeof A.rec_t);
  }
  HSL_CUDA(hsl_launch_setup(P, C, n_t, (const double*)m->params.p, (HslCand*)m->cand.p, (double*)m->ttab.p, A.status, st));
  HSL_CUDA(hsl_launch_forces(P, A, HSL_MODE_GAIT, st));
  m->launches += 2;
  if (status) {
    HSL_CUDA(cudaMemcpyAsync(status, m->status.p, sizeof(int32_t) * C, cudaMemcpyDeviceToHost, st));
    HSL_CUDA(cudaStreamSynchronize(st));
  }
  return fetch_transposed(m, m->dump_z.p, 3 * P.nf, nfr, z, st);
}

// out: [10][n] = hsl_div, a/b, hsl_sqrt(|a|), sqrt(|a|), hsl_atan2(a,b), atan2(a,b), then sin, sin_ref, cos, cos_ref of |a| (|a| <= pi)
int hsl_math_selftest(int n, const double* a, const double* b, double* out) {
  if (n < 1 || !a || !b || !out) return set_err(HSL_ERR_ARG, "bad argument");
  DevBuf da, db, dout;  // released on every path
  cudaError_t e = da.need(sizeof(double) * n);
  if (e == cudaSuccess) e = db.need(sizeof(double) * n);
  if (e == cudaSuccess) e = dout.need(sizeof(double) * 10 * (size_t)n);
  if (e == cudaSuccess) e = cudaMemcpy(da.p, a, sizeof(double) * n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(db.p, b, sizeof(double) * n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = hsl_launch_math_selftest(n, (const double*)da.p, (const double*)db.p, (double*)dout.p, nullptr);
  if (e == cudaSuccess) e = cudaMemcpy(out, dout.p, sizeof(double) * 10 * (size_t)n, cudaMemcpyDeviceToHost);
  da.release(); db.release(); dout.release();
  if (e != cudaSuccess) return set_err(HSL_ERR_CUDA, "CUDA: %s", cudaGetErrorString(e));
  return HSL_OK;
}

// ---------------------------------------------------------------- fall / perturbation sweep
int hsl_fall_sweep_host(HslModel* m, int64_t n_worlds, const double* params, double play_dt, double t0, int n_steps, const int32_t* kick_step,
                        const double* kick_dv, double hc, double tmin, uint8_t* fell, double* t_end, double* final_z, int32_t* status,
                        double* traj, float* kernel_ms) {
  if (!m || n_worlds < 1 || !params || !(play_dt > 0) || n_steps < 1) return set_err(HSL_ERR_ARG, "bad argument");
  if (!m->sim_ok) return set_err(HSL_ERR_UNSUPPORTED, "no simulation constants for this model");
  int rc = ensure_stream(m);
  if (rc) return rc;
  cudaStream_t st = m->stream;
  const HslModelPod& P = m->pod;
  // setup_per_controller (player.cpp:370-382): the gait at n_t = int(T / play_dt + .5) frames, evaluated by the hot path
  const int n_t = (int)(params[7] / play_dt + .5);
  if (n_t < 1) return set_err(HSL_ERR_ARG, "period shorter than play_dt");
  const size_t pbytes = sizeof(double) * HSL_NPARAM;
  HSL_CUDA(m->params.need(pbytes));
  HSL_CUDA(cudaMemcpyAsync(m->params.p, params, pbytes, cudaMemcpyHostToDevice, st));
  HSL_CUDA(m->out4.need(sizeof(double) * 4 + sizeof(int32_t)));
  double* d4 = (double*)m->out4.p;
  int32_t* dst = (int32_t*)(d4 + 4);
  // ignore_reach: set_fall_test switches it on (player.cpp:785)
  rc = eval_gaits_dev(m, 1, n_t, (const double*)m->params.p, HSL_FLAG_IGNORE_REACH, d4, d4 + 1, d4 + 2, d4 + 3, dst, true, st);
  if (rc) return rc;
  int32_t gait_status = 0;
  HSL_CUDA(cudaMemcpyAsync(&gait_status, dst, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  DevBuf ctrl, dks, dkv, dfell, dtend, dz, dstat, dsteps, dtraj, dq1, dA;
  auto cleanup = [&]() { DevBuf* all[] = {&ctrl, &dks, &dkv, &dfell, &dtend, &dz, &dstat, &dsteps, &dtraj, &dq1, &dA}; for (DevBuf* b : all) b->release(); };
#define HSL_CUDA_F(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return set_err(HSL_ERR_CUDA, "CUDA: %s", cudaGetErrorString(e_)); } } while (0)
  HSL_CUDA_F(ctrl.need(sizeof(double) * (size_t)n_t * 3 * P.nmj));
  HSL_CUDA_F(hsl_launch_fall_ctrl(n_t, P.nmj, params[7] / n_t, (const double*)m->dump_q.p, (const double*)m->dump_tau.p, (double*)ctrl.p, st));
  // init_play_config (player.cpp:349-355): joint values of the gait at play_t (frame index of the trajectory dump), FK, body poses
  const double play_t0 = (double)(int)(t0 / play_dt + .5) * play_dt;
  const int f0 = (int)(play_t0 / (params[7] / n_t) + .5);
  std::vector<double> q0(P.config_dim), A0((size_t)16 * P.n);
  if (f0 > n_t + 3) { cleanup(); return set_err(HSL_ERR_ARG, "t0 beyond the recorded trajectory (t0 <= period)"); }
  HSL_CUDA_F(dq1.need(sizeof(double) * P.config_dim));
  HSL_CUDA_F(dA.need(sizeof(double) * 16 * P.n));
  HSL_CUDA_F(cudaMemcpy2DAsync(dq1.p, sizeof(double), (const double*)m->dump_q.p + f0, sizeof(double) * (n_t + 4), sizeof(double), P.config_dim,
                               cudaMemcpyDeviceToDevice, st));  // column f0 of the component-major trajectory
  HSL_CUDA_F(hsl_launch_fk_records(P, 1, (const double*)dq1.p, (double*)dA.p, nullptr, st));
  HSL_CUDA_F(cudaMemcpyAsync(A0.data(), dA.p, sizeof(double) * 16 * P.n, cudaMemcpyDeviceToHost, st));
  HSL_CUDA_F(cudaStreamSynchronize(st));
  if (gait_status & (HSL_ST_BAD_PARAMS | HSL_ST_SOLVER | HSL_ST_FEW_CONTACTS)) { cleanup(); return set_err(HSL_ERR_ARG, "the gait itself does not evaluate (status bits set)"); }
  HslFallArgs A;
  memset(&A, 0, sizeof A);
  hsl_sim_state_from_frames(&m->sim, A0.data(), A.pos0, A.quat0);
  A.n_worlds = n_worlds; A.n_steps = n_steps; A.n_t = n_t; A.iterations = 20;
  A.play_dt = play_dt; A.play_t0 = play_t0; A.hc = hc; A.tmin = tmin;
  A.erp = 0.8; A.cfm = 1e-10; A.soft_cfm = 1e-3; A.bounce = 0.5; A.bounce_vel = 0.1; A.gravity = 1.0; A.kp = 100.0;
  A.ctrl = (const double*)ctrl.p;
  if (kick_step) {
    HSL_CUDA_F(dks.need(sizeof(int32_t) * n_worlds));
    HSL_CUDA_F(cudaMemcpyAsync(dks.p, kick_step, sizeof(int32_t) * n_worlds, cudaMemcpyHostToDevice, st));
    A.kick_step = (const int32_t*)dks.p;
  }
  if (kick_dv) {
    HSL_CUDA_F(dkv.need(sizeof(double) * 3 * n_worlds));
    HSL_CUDA_F(cudaMemcpyAsync(dkv.p, kick_dv, sizeof(double) * 3 * n_worlds, cudaMemcpyHostToDevice, st));
    A.kick_dv = (const double*)dkv.p;
  }
  HSL_CUDA_F(dfell.need(n_worlds)); HSL_CUDA_F(dtend.need(sizeof(double) * n_worlds)); HSL_CUDA_F(dz.need(sizeof(double) * n_worlds));
  HSL_CUDA_F(dstat.need(sizeof(int32_t) * n_worlds)); HSL_CUDA_F(dsteps.need(sizeof(int32_t) * n_worlds));
  A.fell = (uint8_t*)dfell.p; A.t_end = (double*)dtend.p; A.final_z = (double*)dz.p; A.status = (int32_t*)dstat.p; A.steps_done = (int32_t*)dsteps.p;
  if (traj) { HSL_CUDA_F(dtraj.need(sizeof(double) * 3 * (size_t)n_worlds * n_steps)); HSL_CUDA_F(cudaMemsetAsync(dtraj.p, 0, sizeof(double) * 3 * (size_t)n_worlds * n_steps, st)); A.traj = (double*)dtraj.p; }
  if (const char* dbg = getenv("HSL_FALL_DEBUG")) {  // dump the kernel's inputs (control table, initial poses) for inspection
    std::vector<double> hc((size_t)n_t * 3 * P.nmj);
    HSL_CUDA_F(cudaMemcpy(hc.data(), ctrl.p, sizeof(double) * hc.size(), cudaMemcpyDeviceToHost));
    if (FILE* f = fopen(dbg, "wb")) {
      fwrite(hc.data(), sizeof(double), hc.size(), f);
      fwrite(A.pos0, sizeof(double), 3 * HSL_MAX_BODIES, f);
      fwrite(A.quat0, sizeof(double), 4 * HSL_MAX_BODIES, f);
      fwrite(&m->sim, sizeof(HslSimPod), 1, f);
      fclose(f);
    }
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  HSL_CUDA_F(cudaEventCreate(&e0)); HSL_CUDA_F(cudaEventCreate(&e1));
  HSL_CUDA_F(cudaEventRecord(e0, st));
  cudaError_t le = hsl_launch_fall(m->sim, A, m->fall_variant, st);
  cudaEventRecord(e1, st);
  m->launches += 3;
  if (le == cudaSuccess) le = cudaStreamSynchronize(st);
  float ms = 0;
  if (le == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (le != cudaSuccess) { cleanup(); return set_err(HSL_ERR_CUDA, "CUDA: %s", cudaGetErrorString(le)); }
  if (kernel_ms) *kernel_ms = ms;
  if (fell) HSL_CUDA_F(cudaMemcpy(fell, dfell.p, n_worlds, cudaMemcpyDeviceToHost));
  if (t_end) HSL_CUDA_F(cudaMemcpy(t_end, dtend.p, sizeof(double) * n_worlds, cudaMemcpyDeviceToHost));
  if (final_z) HSL_CUDA_F(cudaMemcpy(final_z, dz.p, sizeof(double) * n_worlds, cudaMemcpyDeviceToHost));
  if (status) HSL_CUDA_F(cudaMemcpy(status, dstat.p, sizeof(int32_t) * n_worlds, cudaMemcpyDeviceToHost));
  if (traj) HSL_CUDA_F(cudaMemcpy(traj, dtraj.p, sizeof(double) * 3 * (size_t)n_worlds * n_steps, cudaMemcpyDeviceToHost));
#undef HSL_CUDA_F
  cleanup();
  return HSL_OK;
}

// ---------------------------------------------------------------- multi-GPU: the one collective of the path
// Candidates shard over ranks and never span GPUs; the only exchange is an all-gather of the per-candidate costs
// (8 B each) before selection.  NCCL is bound at run time (dlsym on the process image first, so a host
// that already carries an NCCL -- e.g. torch's bundled one -- is reused; else dlopen of libnccl.so.2): the library has
// no link-time NCCL dependency and single-GPU users never load it.
namespace {
struct NcclApi {
  bool ok = false;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, HslNcclId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi& nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  void* h = dlsym(RTLD_DEFAULT, "ncclAllGather") ? RTLD_DEFAULT : dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return api;
  api.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
  api.CommInitRank = (int (*)(void**, int, HslNcclId, int))dlsym(h, "ncclCommInitRank");
  api.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
  api.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
  api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather;
  return api;
}
int nccl_check(int rc, const char* what) {
  if (rc == 0) return HSL_OK;
  NcclApi& a = nccl_api();
  const std::string msg = std::string(what) + " failed: " + (a.GetErrorString ? a.GetErrorString(rc) : "?");
  return set_err(HSL_ERR_CUDA, "NCCL: %s", msg.c_str());
}
}  // namespace

int hsl_nccl_unique_id(HslNcclId* id) {
  NcclApi& a = nccl_api();
  if (!a.ok) return set_err(HSL_ERR_UNSUPPORTED, "NCCL is not available in this process (libnccl.so.2 not found)");
  if (!id) return set_err(HSL_ERR_ARG, "null argument");
  return nccl_check(a.GetUniqueId(id), "ncclGetUniqueId");
}
int hsl_nccl_comm_init(void** comm, int nranks, const HslNcclId* id, int rank) {
  NcclApi& a = nccl_api();
  if (!a.ok) return set_err(HSL_ERR_UNSUPPORTED, "NCCL is not available in this process (libnccl.so.2 not found)");
  if (!comm || !id || nranks < 1 || rank < 0 || rank >= nranks) return set_err(HSL_ERR_ARG, "bad argument");
  return nccl_check(a.CommInitRank(comm, nranks, *id, rank), "ncclCommInitRank");
}
int hsl_nccl_comm_destroy(void* comm) {
  NcclApi& a = nccl_api();
  if (!a.ok || !comm) return set_err(HSL_ERR_ARG, "bad argument");
  return nccl_check(a.CommDestroy(comm), "ncclCommDestroy");
}
int hsl_allgather_costs(void* nccl_comm, const double* d_local, int64_t n_per_rank, double* d_all, void* stream) {
  NcclApi& a = nccl_api();
  if (!a.ok) return set_err(HSL_ERR_UNSUPPORTED, "NCCL is not available in this process (libnccl.so.2 not found)");
  if (!nccl_comm || !d_local || !d_all || n_per_rank < 1) return set_err(HSL_ERR_ARG, "bad argument");
  return nccl_check(a.AllGather(d_local, d_all, (size_t)n_per_rank, 8 /* ncclFloat64 */, nccl_comm, (cudaStream_t)stream), "ncclAllGather");
}

int hsl_allgather_costs_host(void* nccl_comm, int nranks, const double* local, int64_t n_per_rank, double* all) {
  if (!nccl_comm || !local || !all || n_per_rank < 1 || nranks < 1) return set_err(HSL_ERR_ARG, "bad argument");
  DevBuf a, b;
  cudaStream_t st = nullptr;
  int rc = HSL_OK;
  do {
    if (a.need(sizeof(double) * n_per_rank) != cudaSuccess || b.need(sizeof(double) * n_per_rank * nranks) != cudaSuccess ||
        cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { rc = set_err(HSL_ERR_CUDA, "CUDA: %s", "allocation failed"); break; }
    if (cudaMemcpyAsync(a.p, local, sizeof(double) * n_per_rank, cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = set_err(HSL_ERR_CUDA, "CUDA: %s", "H2D copy failed"); break; }
    if ((rc = hsl_allgather_costs(nccl_comm, (const double*)a.p, n_per_rank, (double*)b.p, st))) break;
    if (cudaMemcpyAsync(all, b.p, sizeof(double) * n_per_rank * nranks, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) { rc = set_err(HSL_ERR_CUDA, "CUDA: %s", "D2H copy failed"); break; }
  } while (0);
  a.release();
  b.release();
  if (st) cudaStreamDestroy(st);
  return rc;
}

int hsl_dfma_probe(int blocks, int threads, int iters, double* tflops, float* ms) {
  if (blocks < 1 || threads < 1 || threads > 1024 || iters < 1 || (size_t)blocks * (size_t)threads > ((size_t)1 << 28))
    return set_err(HSL_ERR_ARG, "bad argument (blocks, iters >= 1, 1 <= threads <= 1024, blocks * threads <= 2^28)");
  DevBuf d;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  float best = 1e30f;
  cudaError_t e = d.need(sizeof(double) * (size_t)blocks * (size_t)threads);
  if (e == cudaSuccess) e = cudaEventCreate(&e0);
  if (e == cudaSuccess) e = cudaEventCreate(&e1);
  if (e == cudaSuccess) e = hsl_launch_dfma_probe((double*)d.p, blocks, threads, iters, nullptr);  // warm-up
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  for (int rep = 0; rep < 5 && e == cudaSuccess; rep++) {
    e = cudaEventRecord(e0, nullptr);
    if (e == cudaSuccess) e = hsl_launch_dfma_probe((double*)d.p, blocks, threads, iters, nullptr);
    if (e == cudaSuccess) e = cudaEventRecord(e1, nullptr);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    float t = 0;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&t, e0, e1);
    if (e == cudaSuccess && t < best) best = t;
  }
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  d.release();
  if (e != cudaSuccess) return set_err(HSL_ERR_CUDA, "CUDA: %s", cudaGetErrorString(e));
  if (ms) *ms = best;
  if (tflops) *tflops = 2.0 * 8.0 * (double)iters * blocks * threads / (best * 1e-3) / 1e12;
  return HSL_OK;
}

}  // extern "C"
