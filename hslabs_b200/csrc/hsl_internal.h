// Internal launcher interface between the C-ABI layer and the kernels.
#pragma once
#include <cuda_runtime.h>

#include "hsl_frame.h"
#include "hsl_model.h"

cudaError_t hsl_launch_frames(const HslModelPod& M, const HslFrameArgs& A, int mode, bool dump, int fb, int maxreg, cudaStream_t st);
cudaError_t hsl_launch_forces(const HslModelPod& M, const HslFrameArgs& A, int mode, cudaStream_t st);
cudaError_t hsl_launch_gait_records(const HslModelPod& M, const HslFrameArgs& A, int n_times, const double* times, double* rec,
                                    cudaStream_t st);
cudaError_t hsl_launch_fk_records(const HslModelPod& M, int64_t n, const double* q, double* Aout, double* Jout, cudaStream_t st);
cudaError_t hsl_launch_ik_records(const HslModelPod& M, int64_t n, int flags, const double* rec, double* q, int32_t* status,
                                  cudaStream_t st);
cudaError_t hsl_launch_setup(const HslModelPod& M, int64_t n_cand, int n_t, const double* params, HslCand* cand, double* ttab,
                             int32_t* status, cudaStream_t st);
// Where the finish kernel also stores a candidate's cost and status: one gather buffer per rank of the job, reached over
// NVLink peer mappings (hsl_gather.cu).  n = 0: nowhere.
#define HSL_MAX_PEERS 16
struct HslPeerOut {
  int32_t n, signal;               // signal: this launch completes the rank's segment -> pad its tail and raise the flags
  double* cot[HSL_MAX_PEERS];      // each already offset to this rank's segment (+ the chunk's first candidate)
  int32_t* status[HSL_MAX_PEERS];
  unsigned long long* flag[HSL_MAX_PEERS];  // this rank's flag in every rank's buffer
  unsigned long long epoch;        // the call number the flags are raised to
  unsigned int* ticket;            // block counter of the finish kernel (this rank's buffer; 0 between launches)
  int64_t pad_lo, pad_hi;          // entries [pad_lo, pad_hi) behind the pointers above do not exist: NaN / 0
};
cudaError_t hsl_launch_finish(int64_t n_cand, int n_t, double total_mass, const HslCand* cand, const double* dt_in,
                              const double* wframe, const double* fmin, const double* fmax, const int32_t* status, double* cot,
                              double* work, double* min_cfz, double* max_mu, cudaStream_t st, const HslPeerOut* peers = nullptr);
// hsl_gather.cu: pad the unused tail of this rank's segment with NaN at every peer, then raise this rank's flag there;
// wait until every rank's flag of this epoch has arrived here
cudaError_t hsl_launch_gather_signal(const HslPeerOut& peers, cudaStream_t st);   // for a rank without candidates (no finish kernel)
cudaError_t hsl_launch_gather_wait(const unsigned long long* flags, int nranks, unsigned long long epoch, cudaStream_t st);
// argmin whose first instructions wait for the ranks' flags (the acquire side of the gather)
cudaError_t hsl_launch_argmin_gathered(const double* cost, int64_t n, int64_t* out_index, double* out_value, const unsigned long long* flags,
                                       int nranks, unsigned long long epoch, cudaStream_t st);

// Launch configuration with programmatic dependent launch allowed (sm_90+): the kernel's blocks may become resident while the
// kernel in front of it on the stream is still draining; the kernel itself must execute griddepcontrol.wait before it
// touches anything that kernel wrote (every kernel launched through this helper does, as its first instruction).
struct HslPdlConfig {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  HslPdlConfig(dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
    cfg = cudaLaunchConfig_t();
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  }
};
#define HSL_GRID_DEP_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define HSL_GRID_DEP_LAUNCH() asm volatile("griddepcontrol.launch_dependents;")

cudaError_t hsl_launch_math_selftest(int n, const double* a, const double* b, double* out, cudaStream_t st);
cudaError_t hsl_launch_topk(const double* cost, int64_t n, int k, int64_t* out_index, double* out_value, cudaStream_t st);
int hsl_topk_launches(int64_t n);  // kernels hsl_launch_topk issues for n costs
cudaError_t hsl_launch_argmin(const double* cost, int64_t n, int64_t* out_index, double* out_value, cudaStream_t st);
// [comps][nfr] -> [nfr][comps] on the device (elem_size 8: double, 1: uint8_t)
cudaError_t hsl_launch_transpose(const void* src, void* dst, int comps, int64_t nfr, int elem_size, cudaStream_t st);
cudaError_t hsl_launch_dfma_probe(double* out, int blocks, int threads, int iters, cudaStream_t st);

// hsl_model_load.cpp
int hsl_build_model_pod(const char* xml_path, HslModelPod* pod, char* err, int errlen);
int hsl_build_sim_pod(const char* xml_path, const HslModelPod* pod, HslSimPod* sim, char* err, int errlen);
