/* hsl.h -- C ABI of the B200 gait-evaluation library (libhsl_b200.so).
 *
 * The reference (underactuated/HSLabs) has no plugin / FFI boundary: its evaluation path is a set of
 * C++ member functions linked into one binary.  These entry points are what a cgo/ctypes/
 * C++ binding of that path would bind; each one names the reference interface it replaces.  All
 * pointers are plain; "_host" entries take host buffers and do the host<->device copies themselves,
 * the others take device pointers and a cudaStream_t (passed as void*).  Every function returns 0 on
 * success or a negative HSL_ERR_* code; hsl_last_error() gives the message.  Per-candidate failures the
 * reference turns into exit(1) are reported in a status array (HSL_ST_* bits) and NaN results.
 *
 * There is no CPU fallback: without a CUDA device every compute entry fails with HSL_ERR_CUDA.
 */
#ifndef HSL_H
#define HSL_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HSL_NPARAM 13
/* candidate row (doubles), the scalars of pgsconfigparams (pergen.h:137-146):
 *  [0..2] torso_pos  [3..5] torso_angles  [6] step_duration  [7] period  [8] step_length
 *  [9] step_height  [10] curvature  [11] foot shift type (-1 none, 0 lateral, 1 radial)  [12] shift value */

#define HSL_OK 0
#define HSL_ERR_ARG -1
#define HSL_ERR_XML -2          /* unreadable / not a <mujoco> file (model.cpp:230-231) */
#define HSL_ERR_UNSUPPORTED -3  /* topology / model outside what the path handles (lik.cpp:12-16) */
#define HSL_ERR_CUDA -4
#define HSL_ERR_NOMEM -5

#define HSL_ST_BAD_PARAMS 1   /* step_duration outside [0,1] (pergen.cpp:31) */
#define HSL_ST_UNREACHABLE 2  /* IK target out of reach (lik.cpp:161-164) */
#define HSL_ST_SOLVER 4       /* contact system not positive definite (ftsolver.cpp:208-232) */
#define HSL_ST_FEW_CONTACTS 8 /* fewer than two feet on the ground */
#define HSL_ST_ILLCOND 16     /* informational: in some frame the feet on the ground are almost collinear (an LDL^T pivot
                                 of the 6x6 level-0 matrix below 1e-4 of its trace).  The reference's FP64 result then
                                 depends on its rank-threshold retry loop (ftsolver.cpp:208-232) and is not
                                 reproducible; the value returned here is the one its algorithm gives in exact
                                 arithmetic.  Results are still written. */

#define HSL_FLAG_IGNORE_REACH 1 /* liksolver::set_ignore_reach_flag(true) (lik.cpp:142-146) */

typedef struct HslModel HslModel;

/* kinematicmodel::load_fromxml + liksolver + periodic::set_dynparts (model.cpp:224-242, lik.cpp:7-78,
 * periodic.cpp:34-58): parse the MuJoCo-style XML, flatten it and keep it resident for the kernels. */
int hsl_model_load_xml(const char* xml_path, HslModel** out);
void hsl_model_free(HslModel* m);
/* dims[6] = n bodies, nf feet, nmj motor joints, config_dim, trunk bodies, lik index (0 myant,1 hexapod,2 spider) */
int hsl_model_dims(const HslModel* m, int32_t dims[6]);
double hsl_model_rcap(const HslModel* m); /* liksolver::get_rcap (lik.h:51) */
/* What periodic::set_dynparts / dynpart::setup hand out (periodic.cpp:34-58, dynrec.cpp:5-16, periodic.h:44-48): parent
 * ids [n] (-1 for the torso), foot body ids [nf] and limb top-link body ids [nf] in LIK order, masses [n], COM offset of
 * every body in its body frame [n][3] (odepart::A_body_geom translation), foot point in the foot body's frame [nf][3]
 * (odepart::capsule_to_pos).  Any pointer may be NULL. */
int hsl_model_tables(const HslModel* m, int32_t* parent, int32_t* footis, int32_t* limb_top, double* masses, double* com_offset,
                     double* foot_offset);
/* packed model block (HslModelPod of hslabs_b200/csrc/hsl_model.h); returns its size, copies min(size,cap) bytes */
size_t hsl_model_pod(const HslModel* m, void* dst, size_t cap);
const char* hsl_last_error(void);
int hsl_device_count(void);
int hsl_set_device(int device);   /* cudaSetDevice for hosts that do not link the CUDA runtime themselves: handles, gather buffers and
                                  * launches of the calling thread go to this device (one process per GPU: call it first) */

/* Two deliberate differences from the reference's outputs, valid for every entry below:
 *  - min_cfz / max_mu are taken over the feet ON THE GROUND.  periodic::analyze_contforces (periodic.cpp:347-357) loops
 *    over all nfeet and so folds in the swing feet's contact-force slots, which hold round-off of the null-space basis
 *    (1e-17): preset 8 reports max_mu = 350 in FP64 and 19.9 when the same code runs in __float128.
 *  - z (contact forces) is always [3 nf], in LIK limb order, with exact zeros for feet in the air; the reference's
 *    solve_forcetorques returns the same layout (extract_N_contact, ftsolver.cpp:276-284) but with that round-off
 *    in the swing-foot slots. */

/* modelplayer::measure_cot over a batch (player.cpp:269-285 inside the sweep of player.cpp:311-321):
 * candidates -> cost of transport, work per period, min contact z-force, max friction ratio.
 * d_params [n_cand][13], outputs [n_cand] (any output may be NULL); all DEVICE pointers. */
int hsl_eval_gaits(HslModel* m, int64_t n_cand, int n_t, const double* d_params, int flags, double* d_cot, double* d_work,
                   double* d_min_cfz, double* d_max_mu, int32_t* d_status, void* stream);
/* Selection after the (all-gathered) costs: index and value of the cheapest valid candidate; NaN costs (failed
 * candidates) are never selected, ties go to the lowest index; index -1 / value NaN when nothing is valid.
 * DEVICE pointers. */
int hsl_select_best(const double* d_cost, int64_t n, int64_t* d_index, double* d_value, void* stream);
/* The k cheapest valid candidates (1 <= k <= n) in ascending (cost, index) order (what a stable sort of the costs gives;
 * NaN never selected): d_index [k], d_value [k] (either may be NULL); entries past the number of valid candidates get
 * -1 / NaN.  A bitonic sort of (cost, index) pairs over all n costs in a stream-ordered workspace (tiles of 4096 sorted in
 * shared memory); DEVICE pointers, queued on `stream`. */
int hsl_select_topk(const double* d_cost, int64_t n, int k, int64_t* d_index, double* d_value, void* stream);
/* same, HOST pointers (pinned staging inside). */
int hsl_eval_gaits_host(HslModel* m, int64_t n_cand, int n_t, const double* params, int flags, double* cot, double* work,
                        double* min_cfz, double* max_mu, int32_t* status);
/* periodic::record_trajectory + compute_torques_over_period with everything kept (periodic.cpp:77-96,
 * 377-391; forcetorquesolver::solve_forcetorques x/z, ftsolver.cpp:78-102).  HOST pointers, row-major:
 * traj [n_cand][n_t+4][config_dim] (frames 0..n_t+3), x [n_cand][n_t][6n], z [n_cand][n_t][3nf],
 * tau [n_cand][n_t][nmj], contacts [n_cand][n_t][nf] -- solved frames 2..n_t+1 in order.  Any may be NULL. */
int hsl_eval_gaits_detail_host(HslModel* m, int64_t n_cand, int n_t, const double* params, int flags, double* cot, double* work,
                               double* min_cfz, double* max_mu, int32_t* status, double* traj, double* x, double* z,
                               double* tau, uint8_t* contacts);
/* periodic::compute_dynrecs .. work_over_period on supplied joint trajectories (periodic.cpp:149-160,
 * 192-202, 285-307): traj [n_cand][n_t+5][config_dim], dt [n_cand]; outputs as above (cot is not defined). */
int hsl_eval_trajectories_host(HslModel* m, int64_t n_cand, int n_t, const double* traj, const double* dt, double* work,
                               double* min_cfz, double* max_mu, int32_t* status, double* x, double* z, double* tau);
/* forcetorquesolver::solve_forcetorques + periodic::get_motor_torques on populated dynrecords
 * (ftsolver.cpp:78-102, periodic.cpp:328-343, dynrec.h:76-106), both torso penalties on.
 * Inputs [n_frames][n][3] (pos, jpos, jzaxis, mom_rate, ang_mom_rate), fpos [n_frames][nf][3],
 * contacts [n_frames][nf]; outputs x [n_frames][6n], z [n_frames][3nf], tau [n_frames][nmj],
 * status [n_frames].  HOST pointers. */
int hsl_solve_frames_host(HslModel* m, int64_t n_frames, const double* pos, const double* jpos, const double* jzaxis,
                          const double* mom_rate, const double* ang_mom_rate, const double* fpos, const uint8_t* contacts,
                          double* x, double* z, double* tau, int32_t* status);

/* Device-pointer forms of the two per-frame entries (L2 and L1), for callers whose dynrecords /
 * trajectories already live in HBM: same layouts as the *_host forms below / above, every pointer a DEVICE pointer, work
 * queued on `stream` (no synchronisation; outputs may be NULL).  status is zeroed by the call. */
int hsl_eval_trajectories(HslModel* m, int64_t n_cand, int n_t, const double* d_traj, const double* d_dt, double* d_work,
                          double* d_min_cfz, double* d_max_mu, int32_t* d_status, double* d_x, double* d_z, double* d_tau,
                          void* stream);
int hsl_solve_frames(HslModel* m, int64_t n_frames, const double* d_pos, const double* d_jpos, const double* d_jzaxis,
                     const double* d_mom_rate, const double* d_ang_mom_rate, const double* d_fpos, const uint8_t* d_contacts,
                     double* d_x, double* d_z, double* d_tau, int32_t* d_status, void* stream);

/* kinematicmodel::set_jvalues + recompute_modelnodes + get_mnode(i)->get_A_ground() / get_joint()->get_A_ground()
 * (model.cpp:183-201,314-318,362-366; what liksolver::get_limb_hip_pos and kinematicmodel::orient_torso read):
 * joint values q [n][config_dim] -> A_ground [n][bodies][16], J_A_ground [n][bodies][16], column-major 4x4 like the
 * reference's `affine` (matrix.cpp:138-146); joint frames of bodies without a joint are zero.  HOST pointers, either
 * output may be NULL. */
int hsl_fk_records_host(HslModel* m, int64_t n, const double* q, double* A_ground, double* J_A_ground);

/* pergensetup::set_rec (pergen.cpp:225-239): frame records rec [n_cand][n_times][6+3nf] (torso position, Euler angles,
 * foot targets in LIK order) of the candidates params [n_cand][13] at the times [n_times]; honours
 * hsl_set_rec_transform.  status [n_cand] (HSL_ST_BAD_PARAMS).  HOST pointers. */
int hsl_gait_records_host(HslModel* m, int64_t n_cand, const double* params, int n_times, const double* times, int flags,
                          double* rec, int32_t* status);
/* kinematicmodel::set_jvalues_with_lik + get_jvalues / liksolver::place_limbs (model.cpp:354-359,369-372,
 * lik.cpp:89-99,316-354): joint values q [n][config_dim] of the records rec [n][6+3nf]; status [n] gets
 * HSL_ST_UNREACHABLE where the reference prints "LIK ERROR" and exits (unless HSL_FLAG_IGNORE_REACH).  HOST pointers. */
int hsl_ik_records_host(HslModel* m, int64_t n, const double* rec, int flags, double* q, int32_t* status);

/* forcetorquesolver::solve_forces (ftsolver.cpp:331-378): least-squares contact forces of ALL feet for given motor
 * torques, torso joint force/torque forced to zero, on populated dynrecords (layouts as hsl_solve_frames_host;
 * torques [n_frames][nmj]; output z [n_frames][3nf]; status [n_frames], HSL_ST_SOLVER when a limb is singular). */
int hsl_solve_forces_host(HslModel* m, int64_t n_frames, const double* pos, const double* jpos, const double* jzaxis,
                          const double* mom_rate, const double* ang_mom_rate, const double* fpos, const double* torques,
                          double* z, int32_t* status);
/* periodic::solve_contforces_given_torques (periodic.cpp:369-374) for every solved frame of generated gaits:
 * params [n_cand][13], torques [n_cand][n_t][nmj] (solve order, frames 2..n_t+1), z [n_cand][n_t][3nf],
 * status [n_cand].  HOST pointers. */
int hsl_solve_forces_gait_host(HslModel* m, int64_t n_cand, int n_t, const double* params, int flags, const double* torques,
                               double* z, int32_t* status);

/* pergensetup::set_rec_rotation / set_rec_transform (pergen.cpp:309-320): a rigid map [Rz(psi)Ry(theta)Rx(phi) | transl]
 * applied to every generated frame record (torso pose and foot targets, pergen.cpp:325-335) of the following
 * hsl_eval_gaits* calls on this handle -- the reference copies one rec_transform to every candidate of a sweep
 * (pergen.cpp:446).  eas = (phi, theta, psi); either pointer may be NULL (= zero); both NULL switches it off. */
int hsl_set_rec_transform(HslModel* m, const double* transl /*[3]*/, const double* eas /*[3]*/);

/* Fall / perturbation sweep (BASELINE configs[4]; run_fall_test.sh, main.cpp:48-50,64): n_worlds copies of the reference's
 * closed loop -- position_control_test(pgs, t0) with set_fall_test(hc, tmin, .) and torso kicks (player.cpp:358-382,
 * 326-340, 393-432, 585-605, 669-681) -- each with its own kick: kick_step [W] (step index, < 0 never) and kick_dv [W][3]
 * (velocity change, applied as the force dv / play_dt for one step), either may be NULL.  params [13]: ONE gait (the
 * controller's targets and feed-forward torques come from its evaluation at n_t = int(period / play_dt + .5)).
 * Outputs [W] (any may be NULL): fell, t_end (time of the fall or end time), final torso COM height, status
 * (bit 0: more than 16 bodies touched the ground); traj optional [W][n_steps][3] torso COM after every step;
 * kernel_ms optional: duration of the sweep kernel.  The rigid-body stepper restates what the reference asks of ODE
 * (dWorldQuickStep: DESIGN.md section 9); agreement with a real ODE build is statistical.  HOST pointers. */
int hsl_fall_sweep_host(HslModel* m, int64_t n_worlds, const double* params, double play_dt, double t0, int n_steps,
                        const int32_t* kick_step, const double* kick_dv, double hc, double tmin, uint8_t* fell, double* t_end,
                        double* final_z, int32_t* status, double* traj, float* kernel_ms);

/* Multi-GPU: candidates shard over one process per GPU and never span GPUs; the path's one collective is
 * the all-gather of the per-candidate costs before selection.  d_local [n_per_rank] -> d_all [nranks][n_per_rank] on
 * every rank, DEVICE pointers, queued on `stream`; nccl_comm is an ncclComm_t (any communicator of the process: the
 * host's own, or one made with the two helpers below from an id that rank 0 creates and hands to the other ranks by
 * whatever means the launcher has).  NCCL is bound at run time; HSL_ERR_UNSUPPORTED when the process has none. */
typedef struct HslNcclId { char internal[128]; } HslNcclId; /* ncclUniqueId */
int hsl_nccl_unique_id(HslNcclId* id);
int hsl_nccl_comm_init(void** nccl_comm, int nranks, const HslNcclId* id, int rank);
int hsl_nccl_comm_destroy(void* nccl_comm);
int hsl_allgather_costs(void* nccl_comm, const double* d_local, int64_t n_per_rank, double* d_all, void* stream);
/* same with HOST arrays (staged through the device, synchronous): for hosts that keep no device memory of their own,
 * e.g. the C++ mirror's sharded measure_cot_sweep. */
int hsl_allgather_costs_host(void* nccl_comm, int nranks, const double* local, int64_t n_per_rank, double* all);

/* The same collective over NVLink peer memory, fused into the evaluation (hsl_gather.cu): every rank owns a gather buffer
 * that all ranks of the job map through CUDA IPC (one process per GPU, one node); the finish kernel of the gait evaluation
 * stores each candidate's cost and status straight into all of them, and a flag per rank (release after the stores,
 * acquire before the selection reads) is all that is left of the all-gather.
 *   hsl_gather_create   allocates this rank's buffer for n_per_rank candidates per rank and returns its IPC handle;
 *   hsl_gather_connect  maps the other ranks' buffers from their handles (all[nranks], exchanged by whatever means the
 *                       launcher has); every rank must have created its buffer before any rank connects;
 *   hsl_eval_gaits_scatter = hsl_eval_gaits of this rank's n_cand <= n_per_rank candidates (0 allowed: an empty shard)
 *                       whose finish kernel also stores cost and status into every rank's buffer; its last block
 *                       raises this rank's flag everywhere.  No extra launch.  Every rank of the job must make the same
 *                       sequence of scatter calls.  The reading side, one of:
 *   hsl_gather_wait     one small kernel that waits for all ranks' flags of the latest scatter; *d_all_cot /
 *                       *d_all_status point at [nranks][n_per_rank] arrays in this rank's buffer (unused tail entries
 *                       NaN / 0), complete for work queued on `stream` afterwards and valid until the second-next
 *                       scatter on this object (two buffers alternate);
 *   hsl_gather_select_best  hsl_select_best over the gathered costs with the wait as the kernel's first instructions;
 *   hsl_eval_gaits_gather = scatter + wait.
 *   hsl_gather_free     unmaps and frees.  Safe once this rank has completed the reading side (wait / select_best /
 *                       gather, stream synchronised) of the last scatter: by then every rank's stores into this rank's
 *                       buffer have landed, and a rank only ever reads its own buffer.  A rank that scattered without
 *                       reading must not free before the other ranks have read (barrier first). */
typedef struct HslGather HslGather;
typedef struct HslIpcHandle { char internal[64]; } HslIpcHandle; /* cudaIpcMemHandle_t */
int hsl_gather_create(int nranks, int rank, int64_t n_per_rank, HslGather** g, HslIpcHandle* mine);
int hsl_gather_connect(HslGather* g, const HslIpcHandle* all);
int hsl_gather_free(HslGather* g);
int hsl_eval_gaits_scatter(HslModel* m, HslGather* g, int64_t n_cand, int n_t, const double* d_params, int flags, double* d_cot,
                           double* d_work, double* d_min_cfz, double* d_max_mu, int32_t* d_status, void* stream);
int hsl_gather_wait(HslGather* g, const double** d_all_cot, const int32_t** d_all_status, void* stream);
int hsl_gather_select_best(HslGather* g, int64_t* d_index, double* d_value, void* stream);
int hsl_eval_gaits_gather(HslModel* m, HslGather* g, int64_t n_cand, int n_t, const double* d_params, int flags, double* d_cot,
                          double* d_work, double* d_min_cfz, double* d_max_mu, int32_t* d_status, const double** d_all_cot,
                          const int32_t** d_all_status, void* stream);
int64_t hsl_gather_size(const HslGather* g);   /* nranks * n_per_rank: entries of the gathered arrays */
/* The waits on the device are bounded (30 s): a rank that never arrives must not hang the GPU.  hsl_gather_check synchronises
 * the device and returns an error if any wait on this object gave up (its results are then incomplete). */
int hsl_gather_check(HslGather* g);
/* the same with HOST arrays (params [n_cand][13] in; all_cot / all_status [nranks * n_per_rank] out, either may be NULL);
 * synchronous -- for hosts that keep no device memory of their own, e.g. the C++ mirror's sharded measure_cot_sweep */
int hsl_eval_gaits_gather_host(HslModel* m, HslGather* g, int64_t n_cand, int n_t, const double* params, int flags, double* all_cot,
                               int32_t* all_status);

/* Page-locked host memory for input / output arrays of the *_host entries (they copy straight from / into the caller's
 * buffers; pageable memory works too, at the driver's staged-copy rate).  NULL on failure. */
void* hsl_pinned_alloc(size_t bytes);
void hsl_pinned_free(void* p);

/* tuning / measurement helpers */
int hsl_set_tuning(HslModel* m, int fb, int maxreg);          /* cost-only kernel variant: frame slots per block (32|64), register cap per thread (64..255; 1 = persistent pipelined kernel) */
int hsl_get_tuning(const HslModel* m, int* fb, int* maxreg);  /* the variant in use (the per-model default unless hsl_set_tuning was called) */
/* hsl_eval_gaits / hsl_eval_gaits_host evaluate a batch of any size as consecutive chunks of at most max_slots frame
 * slots (a candidate takes n_t + 4); default 2^26 (~2 GB of per-frame workspace), bounds 5 .. 2^31 - 1.  Results do
 * not depend on the chunking. */
int hsl_set_max_slots(HslModel* m, int64_t max_slots);
/* hsl_fall_sweep_host kernel: 1 (default) a warp per world with the world in registers, 0 a thread per world (the
 * host-emulatable form tests/hostcheck runs on the CPU); same arithmetic and row order, results agree world by world. */
int hsl_set_fall_variant(HslModel* m, int variant);
int64_t hsl_launch_count(const HslModel* m);                  /* kernels launched through this handle so far */
/* Measurement: with timing on, hsl_eval_gaits* brackets its three kernels (candidate setup, per-frame kernel, per-candidate
 * finish) with CUDA events on the launching stream; hsl_last_kernel_ms waits for the last chunk evaluated on this handle and
 * returns their durations.  bench.py takes the roofline's kernel time from ms[1]. */
int hsl_set_kernel_timing(HslModel* m, int on);
int hsl_last_kernel_ms(HslModel* m, float ms[3]);
int hsl_dfma_probe(int blocks, int threads, int iters, double* tflops, float* ms); /* FP64 FMA throughput of the device */
int hsl_math_selftest(int n, const double* a, const double* b, double* out /*[10][n]*/); /* accuracy of the kernels' branch-free div/sqrt/atan2/sincos vs the library ones (HOST pointers) */

#ifdef __cplusplus
}
#endif
#endif
