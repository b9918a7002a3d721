/* islam_pvgo.h — C ABI of the B200-native PVGO back-end (libislam_pvgo.so).
 *
 * The reference (sair-lab/iSLAM) has no FFI: its back-end is Python calling PyPose.  The entry points below
 * are what a maintainer binds (ctypes) to replace, one for one, the reference calls
 *
 *   pvgo.py:168        PoseVelGraph(init_nodes, init_vels)            -> islam_pvgo_create / _set_state
 *   pvgo.py:125-165    information matrices + input staging            -> islam_pvgo_set_problem (4 scalars)
 *   pvgo.py:26-64      PoseVelGraph.forward (4 residual groups)        -> islam_pvgo_linearize / _get_residuals
 *   pvgo.py:169-180    pp.optim.LM(...).step + StopOnPlateau loop      -> islam_pvgo_lm_reset / _lm_try / _lm_run
 *   pvgo.py:67-78,95-111  vo_loss / imu_loss (+ autograd to vo_motions) -> islam_pvgo_vo_loss / _imu_loss
 *   pvgo.py:114-119    align_to                                        -> islam_pvgo_align
 *   imu_integrator.py:69-164 + pp.module.IMUPreintegrator.forward      -> islam_imu_preintegrate
 *   PyPose LieTensor Exp/Log/Inv/Mul/Act (+ left-tangent backward)     -> islam_lie_* (elementwise)
 *   dense_ba.py:88-176 scale_from_disp_flow (TartanVO.py:159-171)      -> islam_scale_from_disp_flow (batched)
 *
 * Conventions
 *   - All tensor arguments are raw DEVICE pointers to contiguous row-major float32 (or as stated) buffers
 *     allocated by the caller (PyTorch).  The library owns only its handle and workspace.
 *   - Every call takes the cudaStream_t to enqueue on (as void*), is asynchronous unless stated, and is
 *     CUDA-graph capturable except the functions marked "synchronises".
 *   - Return value: 0 ok, <0 invalid argument / unsupported, >0 cudaError_t.  islam_pvgo_create: -2 bad edge list,
 *     -5 boundary too wide for the back-substitution kernel, -6 a single-GPU-only call on an n_parts > 1 handle
 *     (islam_pvgo_lm_try / _lm_run / _solve / _profile_try), -7 more than 128 GB of factor panels, -8 graph too large for the 29-bit block offsets.
 *     islam_pvgo_lm_step / _lm_run: -9 the step / loop did not close within its worst-case try budget.
 *     LM state `info`: 1 Cholesky failed (PyPose's "Linear solver failed"), 2 a multi-GPU peer never answered, 3 a device-side
 *     wait inside the back-substitution timed out (wedged device); 2 and 3 also clear `continual`.
 *   - Numerical failure of the Cholesky (non-positive pivot / NaN) does not abort: it raises `info` in the
 *     LM state, and the step is abandoned exactly as PyPose's "Linear solver failed. Breaking..." path.
 *   - A handle is not thread-safe; use one per host thread / stream.
 *   - SE3 storage [tx,ty,tz,qx,qy,qz,qw]; tangent order [tau(3), phi(3)]; unknown block per node
 *     [tau, phi, v] (9); left perturbation X <- Exp(d) X.
 */
#ifndef ISLAM_PVGO_H_
#define ISLAM_PVGO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct islam_pvgo islam_pvgo;

typedef struct islam_pvgo_opts {
    int32_t band_max;    /* edges with |i-j| above this are loop closures (root separator); default 16 */
    int32_t leaf_max;    /* a leaf front holds up to 3*leaf_max 3-dof variables (tau, phi, v of a pose); default 8 */
    int32_t pivot_max;   /* a front eliminates up to 3*pivot_max variables = 9*pivot_max columns; default 8, max 21 */
    int32_t n_parts;     /* contiguous pose windows (multi-GPU sharding); default 1 */
    int32_t part;        /* this rank's window in [0, n_parts); default 0 */
    int32_t reserved[3];
} islam_pvgo_opts;

typedef struct islam_pvgo_dims {
    int32_t N, E, M;         /* poses, VO/loop-closure edges, IMU pairs (N-1) */
    int32_t P;               /* unique off-diagonal 9x9 blocks of J^T W J */
    int32_t F, levels;       /* fronts, elimination-tree height */
    int32_t band, root_pivots;   /* longest short edge; loop-closure poses promoted to the root */
    int32_t max_rows, max_cols;
    int32_t n_shared_fronts; /* fronts factored redundantly on every rank (multi-GPU) */
    int32_t bs_launches;     /* kernel launches of one back-substitution (top levels are chained inside one launch) */
    int64_t L_doubles, U_doubles;
    int64_t shared_doubles;  /* length of the per-try all-reduce buffer (multi-GPU) */
    double factor_flops;
} islam_pvgo_dims;

/* LM / trust-region / scheduler state, resident on the device; mirrors the attributes PyPose keeps on
 * pp.optim.LM (loss, last, reject_count), strategy.TrustRegion (radius, down, damping) and
 * scheduler.StopOnPlateau (steps, patience_count, continual).  SURVEY.md A.4. */
typedef struct islam_lm_state {
    double loss, last, loss_trial;
    double damping, radius, down;
    double diag_scale, quality, denom;
    double lin_loss;          /* sum r^2 at the last linearisation point */
    int32_t reject_count, steps_done, tries_total, accepted_last;
    int32_t need_linearize, continual, patience_count, info;
    int32_t cur, active, do_lin, chol_fail;
    int32_t loss_valid, pad0, pad1, pad2;
} islam_lm_state;

typedef struct islam_lm_params {
    double radius;            /* TrustRegion(radius=...)  pvgo.py:170 */
    double lm_min, lm_max;    /* LM(min=1e-4, max=1e32)   pvgo.py:171 */
    double high, low, up, down, factor, tr_min, tr_max;   /* TrustRegion defaults */
    int32_t reject;           /* LM(reject=16) */
    int32_t max_steps;        /* StopOnPlateau(steps=10) */
    int32_t patience;         /* StopOnPlateau(patience=3) */
    int32_t use_scheduler;    /* 0: exactly max_steps optimizer.step calls */
    double decreasing;        /* StopOnPlateau(decreasing=1e-3) */
} islam_lm_params;

/* ---- lifetime (synchronises; host-side symbolic analysis of the fixed graph structure) ------------- */
int islam_pvgo_create(islam_pvgo** out, int32_t N, int32_t E, const int64_t* links_host /* E x 2 */,
                      const islam_pvgo_opts* opts /* may be NULL */);
void islam_pvgo_destroy(islam_pvgo* h);
int islam_pvgo_get_dims(const islam_pvgo* h, islam_pvgo_dims* out);
void islam_lm_default_params(islam_lm_params* p);

/* ---- problem data (device pointers, copied into the handle) ---------------------------------------- */
int islam_pvgo_set_problem(islam_pvgo* h, const float* vo_motions /* E x 7 */, const float* imu_drots /* M x 4 */,
                           const float* imu_dtrans /* M x 3 */, const float* imu_dvels /* M x 3 */,
                           const float* dts /* M */, const double info_w[4] /* host: w0^2,w1^2,w2^2,w3^2 */,
                           void* stream);
/* optional 5th residual group, the sparse reprojection factor (pvgo.py:53-61,130-165 with dense_ba.py:276-305
 * SparseReprojectionLoss): per consecutive pair i, n_points camera-frame points of pose i (point3d, M x n_points x 3) and
 * their target pixels in the camera at pose i+1 (M x n_points x 2), both device pointers (copied); intrinsics fx, fy, cx, cy
 * and the rgb2imu pose (7) are host arrays; info_w = (loss_weight[4] / n_points)^2.  motion[0] is the constant 0.1 as at
 * pvgo.py:57 (zero Jacobian for pair 0).  n_points = 0 removes the factor.  Multi-GPU: every rank stages all pairs and evaluates
 * the ones it owns, like the IMU factors. */
int islam_pvgo_set_reproj(islam_pvgo* h, const float* point3d, const float* target, int32_t n_points,
                          const float intrinsics[4], const float rgb2imu[7], double info_w, void* stream);
int islam_pvgo_get_reproj_residuals(islam_pvgo* h, float* reprojerr /* M x 2 n_points */, void* stream);
int islam_pvgo_set_state(islam_pvgo* h, const float* nodes /* N x 7 */, const float* vels /* N x 3 */, void* stream);
int islam_pvgo_get_state(islam_pvgo* h, float* nodes, float* vels, void* stream);

/* ---- kernel family 1: residuals, Jacobian blocks, J^T W J / J^T W r ------------------------------- */
int islam_pvgo_linearize(islam_pvgo* h, void* stream);
/* residual groups in the reference's return order (pvgo.py:64); any pointer may be NULL */
int islam_pvgo_get_residuals(islam_pvgo* h, float* pgerr /* E x 6 */, float* adjvelerr /* M x 3 */,
                             float* imuroterr /* M x 3 */, float* transvelerr /* M x 3 */, void* stream);
/* block-sparse normal equations (float64): Hd N x 81, Ho P x 81 (rows = lower-indexed node), g N x 9,
 * pairs P x 2 (int32 lo,hi).  Any pointer may be NULL. */
int islam_pvgo_get_normal_eq(islam_pvgo* h, double* Hd, double* Ho, double* g, int32_t* pairs, void* stream);

/* ---- kernel family 2: damped multifrontal Cholesky solve (testing hook: one solve at a given scale) - */
int islam_pvgo_solve(islam_pvgo* h, double diag_scale, double lm_min, double lm_max, double* D /* N x 9 */,
                     int32_t* info_host /* may be NULL; non-NULL synchronises */, void* stream);

/* ---- LM driver ---------------------------------------------------------------------------------------- */
int islam_pvgo_lm_reset(islam_pvgo* h, const islam_lm_params* p, void* stream);
/* one try of optimizer.step (linearise if a new step starts, damp, factor, solve, retract, trial loss,
 * trust-region update, accept / roll back); no host synchronisation */
int islam_pvgo_lm_try(islam_pvgo* h, void* stream);
/* optimizer.step: tries until accepted (synchronises once per try to read the verdict) */
int islam_pvgo_lm_step(islam_pvgo* h, islam_lm_state* out /* may be NULL */, void* stream);
/* the whole `while scheduler.continual()` loop of pvgo.py:177-180, control flow on the device;
 * synchronises once at the end (again only if rejected tries exhausted the speculative budget) */
int islam_pvgo_lm_run(islam_pvgo* h, islam_lm_state* out /* may be NULL */, void* stream);
int islam_pvgo_get_lm_state(islam_pvgo* h, islam_lm_state* out, void* stream); /* synchronises */
/* one try timed phase by phase with CUDA events on `stream` (ms[5]: linearise, factor, back-substitution,
 * retract + trial loss + control, total); synchronises — measurement aid for bench.py's roofline object */
int islam_pvgo_profile_try(islam_pvgo* h, float* ms, void* stream);
/* multi-GPU (n_parts > 1): one try split around the collectives.  Every rank holds the whole (small) state but
 * linearises only the factors it owns and factors only its window's fronts:
 *   try_begin : linearise owned factors, factor the private fronts, write the partial panels of the shared
 *               (separator) fronts + the partial linearisation loss into the shared buffer
 *   -> all-reduce(SUM) of islam_pvgo_shared_buffer   (the per-try exchange of separator J^T W J / J^T W r blocks)
 *   try_mid   : factor the shared fronts (redundantly, identical on every rank), back-substitute, retract, evaluate
 *               the trial residuals of the owned factors -> 2 partial sums
 *   try_end   : exchange of the 2 trial sums (sum r^2, quality term) + trust-region update + accept / roll back
 *               (identical decision on every rank).  With peer mailboxes connected (islam_pvgo_mailbox_*: CUDA IPC, ranks
 *               on one node) the exchange happens INSIDE the try_end kernel by direct NVLink stores into the peers'
 *               memory: the all-reduce above is the only collective of the try.  Without them the caller all-reduces
 *               islam_pvgo_sums_buffer between try_mid and try_end.
 * With n_parts == 1 the three calls in sequence are exactly islam_pvgo_lm_try. */
int islam_pvgo_lm_try_begin(islam_pvgo* h, void* stream);
int islam_pvgo_shared_buffer(islam_pvgo* h, double** dev_ptr, int64_t* n_doubles);
int islam_pvgo_lm_try_mid(islam_pvgo* h, void* stream);
int islam_pvgo_sums_buffer(islam_pvgo* h, double** dev_ptr, int64_t* n_doubles);
int islam_pvgo_lm_try_end(islam_pvgo* h, void* stream);
/* multi-GPU with a DENSE root (BASELINE config 4: thousands of loop closures; csrc/dense_root.cuh): the root is factored
 * by all ranks together, 1-D block-column-cyclic over 128-column tile columns.  The try becomes
 *   try_begin : ... as above, plus this rank's share of the root (its factors' blocks, its subtrees' update matrices)
 *   -> all-reduce(SUM) of the shared buffer, of R (ld * n doubles) and of diag (n doubles: the original diagonal, whose
 *      clamp is not linear, travels apart)
 *   try_mid   : shared fronts, then the root's clamped + damped diagonal; returns BEFORE the root is factored
 *   for k0 = 0, block, 2 block ... < n:
 *       islam_pvgo_root_panel(k0)                       (factors the block's columns; a no-op except on
 *                                                        islam_pvgo_root_owner(k0))
 *       -> broadcast of R[k0 * ld, (k0 + min(block, n - k0)) * ld) from that owner (the factored block column)
 *       islam_pvgo_root_update(k0)                      (trailing update of this rank's tile columns; block = 128)
 *   try_mid2  : root back-substitution (replicated), the window's back-substitution, retract, trial residuals
 *   try_end   : as above.
 * islam_pvgo_root_buffers returns n == 0 when the graph has no dense root or n_parts == 1 (then try_mid does it all and
 * try_mid2 / root_panel / root_update return -1). */
int islam_pvgo_root_buffers(islam_pvgo* h, double** R, int64_t* n, int64_t* ld, double** diag, int32_t* block);
int islam_pvgo_root_owner(const islam_pvgo* h, int64_t k0);
int islam_pvgo_root_panel(islam_pvgo* h, int64_t k0, void* stream);
int islam_pvgo_root_update(islam_pvgo* h, int64_t k0, void* stream);
/* look-ahead: root_update split in two — which = 1: only the NEXT block's tile column (work for islam_pvgo_root_owner(k0 +
 * block) alone), which = 2: everything right of it.  After part 1 the next block can be factored and broadcast on a second
 * stream while part 2 of this block still runs (islam_b200/dist.py). */
int islam_pvgo_root_update_part(islam_pvgo* h, int64_t k0, int32_t which, void* stream);
/* optional, after try_mid: zero this rank's copy of the tile columns it does not own.  A factored block can then travel as
 * all-reduce(SUM) of R[k0 * ld, ...) instead of a broadcast (the owner contributes the block, everybody else zeros): on
 * NVSwitch the reduction + multicast happen in the fabric. */
int islam_pvgo_root_zero_foreign(islam_pvgo* h, void* stream);
int islam_pvgo_lm_try_mid2(islam_pvgo* h, void* stream);
/* peer mailboxes: every rank exports the IPC handle of its mailbox (64 bytes), the caller all-gathers them (rank order)
 * and hands the table to every rank.  LM state info = 2 reports a peer that never answered (2 s timeout). */
int islam_pvgo_mailbox_export(islam_pvgo* h, void* handle_out /* 64 bytes */);
int islam_pvgo_mailbox_connect(islam_pvgo* h, const void* handles /* n_parts x 64 bytes */);
/* owner window of every 3-dof variable (host array of 3N int32, [tau, phi, v] per pose): >= 0 private to that rank,
 * -1 shared / replicated (solved redundantly on every rank) */
int islam_pvgo_var_parts(const islam_pvgo* h, int32_t* out_host);

/* ---- small-graph fast path: the whole run_pvgo of a window (pvgo.py:122-205) in ONE launch, one CTA per window --------
 * For the window sizes train.py actually uses (run_kitti.sh:8: batch_size 8 => N = 9 poses): linearise, damp, dense banded
 * float64 Cholesky in shared memory, solve, retract, trial loss, trust region, accept / roll back, StopOnPlateau, align_to
 * and vo_loss (+ gradient) without leaving the SM.  B windows of identical structure (same N and edge list) per call.
 * links_dev: E x 2 int32 on the device; links_host: the same as int64 on the host (validation, bandwidth).  All other
 * pointers are device arrays, window-major: nodes0 B x N x 7, vels0 B x N x 3, vo_motions B x E x 7, imu_* B x (N-1) x ..,
 * dts B x (N-1).  info_w: host, loss_weight^2 as islam_pvgo_set_problem.  Outputs: nodes / vels ALIGNED to each window's
 * first initial pose (pvgo.py:195), one islam_lm_state per window, and — if trans_loss is given — vo_loss of `vo_P` (NULL:
 * vo_motions itself) with its left-tangent gradients (nullable).  Supported: 2 <= N <= 16, E <= 128 (islam_pvgo_small_supported). */
int islam_pvgo_small_supported(int32_t N, int32_t E);
int islam_pvgo_small_run(int32_t B, int32_t N, int32_t E, const int32_t* links_dev, const int64_t* links_host,
                         const float* nodes0, const float* vels0, const float* vo_motions, const float* imu_drots,
                         const float* imu_dtrans, const float* imu_dvels, const float* dts, const double info_w[4],
                         const islam_lm_params* params, float* nodes_out, float* vels_out, islam_lm_state* state_out /* device, B */,
                         const float* vo_P, float* trans_loss /* B x E */, float* rot_loss, float* grad_trans /* B x E x 6 */,
                         float* grad_rot, void* stream);

/* ---- outer losses and gauge alignment ------------------------------------------------------------------ */
/* vo_loss (pvgo.py:67-78) at the current nodes (detached) for arbitrary vo_motions P (E x 7):
 * trans_loss/rot_loss (E); if grad_* given: d loss_e / d(left tangent of P_e) (E x 6 each) */
int islam_pvgo_vo_loss(islam_pvgo* h, const float* P, float* trans_loss, float* rot_loss,
                       float* grad_trans /* nullable */, float* grad_rot /* nullable */, void* stream);
/* imu_loss (pvgo.py:95-111) at the current nodes / velocities for the given IMU measurements (NULL: the ones stored by
 * set_problem): trans_loss = |dv - diff(v)|^2, rot_loss = |Log(dR^-1 R_i^-1 R_{i+1})|^2 per pair (M).  grad_*: nullable,
 * M x 3 each: d rot_loss / d(left tangent of dR), d trans_loss / d(dv)  (pvgo.py:149-150,188-189: the loss back-propagates
 * into the IMU model through imu_drots / imu_dvels) */
int islam_pvgo_imu_loss(islam_pvgo* h, const float* imu_drots /* M x 4, nullable */, const float* imu_dvels /* M x 3, nullable */,
                        float* trans_loss /* M */, float* rot_loss /* M */, float* grad_drots /* nullable */,
                        float* grad_dvels /* nullable */, void* stream);
/* align_to (pvgo.py:114-119): nodes <- T X0^-1 nodes, vels <- R_T R0^-1 vels; target: 7 floats on device */
int islam_pvgo_align(islam_pvgo* h, const float* target, float* nodes_out, float* vels_out, void* stream);

/* ---- IMU pre-integration (imu_integrator.py:69-164 fused with pp.module.IMUPreintegrator.forward) ----- */
/* S samples; K frames; frame f integrates samples [offsets[f], offsets[f+1]) (offsets: K+1 int32, device).
 * init: pos(3), rot(4 xyzw), vel(3) on device.  motion_mode 0: world-frame chain; 1: relative deltas.
 * outputs K x 3 / K x 4 / K x 3 (the caller prepends init in world mode, as imu_integrator.py:86-89). */
int islam_imu_preintegrate(const float* acc /* S x 3 */, const float* gyro /* S x 3 */, const float* dt /* S */,
                           int32_t S, const int32_t* offsets, int32_t K, const float* init /* 10 */,
                           float gravity, int32_t motion_mode, float* pos, float* rot, float* vel,
                           void* workspace /* islam_imu_workspace_bytes(S,K) */, void* stream);
int64_t islam_imu_workspace_bytes(int32_t S, int32_t K);

/* ---- metric scale of the VO translation (dense_ba.py:88-176 scale_from_disp_flow; call site TartanVO.py:159-171) ---- */
/* One fused pass over a batch of B samples (the reference loops over samples in Python).  disp B x H x W (or NULL when
 * depth is given), flow B x 2 x H x W, motion B x 7 (SE3, frame k -> k+1 as TartanVO reports it), intr B x 4 (fx, fy, cx, cy),
 * baseline B, depth B x H x W or NULL, mask_in B x H x W bytes or NULL (the Canny edge mask), disp_th B.
 * Outputs: scale B, z B x H x W, mask / depth_mask B x H x W bytes, mask_count B (nullable; the reference warns below 500).
 * grad_sums (nullable, B x 11 float64): what the backward pass of s = num / den into `motion` needs, reduced in the same pass:
 * {num, den, d num / d a (3), d den / d a (3), d num / d(left tangent of R = T.Inv().rotation()) (3)}, a = K normalize(t_inv). */
int islam_scale_from_disp_flow(const float* disp, const float* flow, const float* motion, const float* intr,
                               const float* baseline, const float* depth, const uint8_t* mask_in, const float* disp_th,
                               int32_t B, int32_t H, int32_t W, float* scale, float* z, uint8_t* mask, uint8_t* depth_mask,
                               int32_t* mask_count, double* grad_sums, void* workspace /* islam_scale_workspace_bytes(B,H,W) */,
                               void* stream);
int64_t islam_scale_workspace_bytes(int32_t B, int32_t H, int32_t W);

/* ---- elementwise LieTensor maps (forward + left-tangent backward), n elements ------------------------- */
enum { ISLAM_SE3 = 0, ISLAM_SO3 = 1 };
int islam_lie_exp(int32_t group, const float* x, float* y, int64_t n, void* stream);          /* algebra -> group */
int islam_lie_log(int32_t group, const float* x, float* y, int64_t n, void* stream);          /* group -> algebra */
int islam_lie_inv(int32_t group, const float* x, float* y, int64_t n, void* stream);
int islam_lie_mul(int32_t group, const float* a, const float* b, float* y, int64_t n, void* stream);
int islam_lie_act(int32_t group, const float* x, const float* p, float* y, int64_t n, void* stream);
/* backward: given dL/dy (embedding-sized rows, tangent in the leading slots) produce dL/dx likewise */
int islam_lie_exp_bwd(int32_t group, const float* x, const float* gy, float* gx, int64_t n, void* stream);
int islam_lie_log_bwd(int32_t group, const float* y, const float* gy, float* gx, int64_t n, void* stream);
int islam_lie_inv_bwd(int32_t group, const float* y, const float* gy, float* gx, int64_t n, void* stream);
int islam_lie_mul_bwd(int32_t group, const float* a, const float* gy, float* ga, float* gb, int64_t n, void* stream);
int islam_lie_act_bwd(int32_t group, const float* x, const float* p, const float* gy, float* gx, float* gp,
                      int64_t n, void* stream);

/* ---- host-only introspection of the symbolic analysis (needs no GPU; arrays are int32 except f_Loff/f_Uoff int64) -- */
typedef struct islam_plan islam_plan;
int islam_plan_build(islam_plan** out, int32_t N, int32_t E, const int64_t* links_host, const islam_pvgo_opts* opts);
void islam_plan_free(islam_plan* p);
int64_t islam_plan_array(const islam_plan* p, const char* name, const void** ptr); /* returns length or -1 */

/* ordered prefix product over n elements (pp.cumprod; Datasets/transformation.py:100-113 motion2pose_pypose is
 * left = 0 with the start pose prepended): left = 1: y_i = x_i * y_{i-1};  left = 0: y_i = y_{i-1} * x_i */
int islam_lie_cumprod(int32_t group, const float* x, float* y, int64_t n, int32_t left, void* stream);

const char* islam_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ISLAM_PVGO_H_ */
