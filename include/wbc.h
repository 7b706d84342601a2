/* wbc.h — C ABI of the batched whole-body controller (libwbc_b200.so).
 *
 * Drop-in boundary for the per-control-step path of vincekurtz/quadruped_drake.
 * Every entry point replaces a block of pydrake calls made by the reference
 * controllers; the citation after each declaration is the reference code it
 * stands in for (paths relative to the reference tree).
 *
 * Conventions
 *   - plain C, no exceptions; every function returns an int status, 0 = WBC_OK;
 *     wbc_last_error() gives the text for the last non-zero return on a handle;
 *   - all arithmetic is FP64; arrays are instance-major (row i = instance i);
 *   - q  = [qw qx qy qz  px py pz  theta(12)]            (19)  Drake position order
 *     v  = [omega_W(3)  pdot_W(3)  thetadot(12)]         (18)  Drake velocity order
 *     joint/velocity order inside theta is given by wbc_model.v_index;
 *   - traj[54] follows lcm_types/trunk_state_t.lcm:10-35 field order:
 *       base p, pd, pdd, rpy, rpyd, rpydd (18),
 *       foot p   LF RF LH RH (12), foot pd (12), foot pdd (12);
 *     contact[4] = planned stance flags LF RF LH RH (trunk_state_t.lcm:38-41,
 *     planners/simple.py:67);
 *   - tau[12] is in actuator (URDF <transmission>) order, like the vector the
 *     reference writes to its `quad_torques` port (basic_controller.py:320);
 *   - metrics[4] = [V, err, res, Vdot] (basic_controller.py:271-283);
 *   - per-instance failures go to status[i] (bit mask below), never abort the batch
 *     (the reference asserts instead: inverse_dynamics_controller.py:224). An instance with ANY status bit set
 *     returns tau = 0 (and f = vd = lam = 0 where requested): an unfinished active-set iterate, a substituted
 *     orientation (bad quaternion, gimbal lock) or a rank-deficient problem never leaves the library as torques.
 *     metrics[1] (err) is still the tracking error of the state; callers must mask on status;
 *   - "device" entry points take device pointers and a cudaStream_t passed as void*;
 *     "_host" entry points take host pointers and do the copies themselves.
 *   - one handle per device; calls on one handle must be serialised by the caller.
 */
#ifndef WBC_B200_H
#define WBC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WBC_NQ 19
#define WBC_NV 18
#define WBC_NU 12
#define WBC_NLEG 4
#define WBC_NBODY 13      /* floating base + 4 x (hip/abduct, thigh, shank) after welding */
#define WBC_NTRAJ 54
#define WBC_NMETRIC 4
#define WBC_NLAM 42       /* inequality multipliers: 16 friction rows + 2 controller rows + 24 torque-box rows */

/* return codes */
#define WBC_OK 0
#define WBC_ERR_ARG 1
#define WBC_ERR_CUDA 2
#define WBC_ERR_NOMEM 3

/* per-instance status bits */
#define WBC_ST_OK 0
#define WBC_ST_MAXITER 1       /* active-set iteration cap reached                       */
#define WBC_ST_INFEASIBLE 2    /* QP constraints inconsistent                            */
#define WBC_ST_RANKDEF 4       /* equality constraints rank deficient (e.g. singular leg) */
#define WBC_ST_GIMBAL 8        /* |cos(pitch)| < 1e-6: rpy rates undefined   */
#define WBC_ST_NOTPD 16        /* reduced Hessian not positive definite                   */
#define WBC_ST_BADQUAT 32      /* zero / non-finite quaternion                            */
#define WBC_ST_UNSUPPORTED 64  /* PC controller with no stance foot (the reference raises too) */
#define WBC_ST_DIVERGED 128    /* rollout only: non-finite / runaway velocity, instance frozen   */

/* controller kinds: reference controllers/__init__.py:1-5 */
#define WBC_CTRL_ID 0          /* controllers/inverse_dynamics_controller.py */
#define WBC_CTRL_CLF 1         /* controllers/clf_controller.py              */
#define WBC_CTRL_PC 2          /* controllers/pc_controller.py               */
#define WBC_CTRL_MPTC 3        /* controllers/mptc_controller.py (PC without the passivity rows) */
#define WBC_CTRL_PD 4          /* BasicController.ControlLaw, basic_controller.py:322-352 (joint PD) */

/* Flattened robot: what Drake's Parser + MultibodyPlant::Finalize hold after
 * simulate.py:37-64. Body 0 = floating base, body 1+3*leg+j = link j of leg
 * (legs in foot order LF RF LH RH, basic_controller.py:67-70). Welded links are
 * already merged into their parent. Joint k = 3*leg+j connects body k+1 to its
 * parent (base for j = 0, body k otherwise); all joint frames are unrotated
 * (every moving-joint rpy is 0 in both reference URDFs). */
typedef struct wbc_model {
  double mass[WBC_NBODY];
  double com[WBC_NBODY][3];          /* CoM in the body frame                                  */
  double inertia_com[WBC_NBODY][6];  /* about the CoM, body axes: xx yy zz xy xz yz            */
  double joint_xyz[WBC_NU][3];       /* joint origin in the parent body frame                  */
  double joint_axis[WBC_NU][3];      /* unit axis (same in parent and child frame)             */
  double foot_xyz[WBC_NLEG][3];      /* foot frame origin in the shank frame                   */
  double effort[WBC_NU];             /* URDF effort limit of joint k                           */
  double gravity[3];                 /* world gravity vector, (0,0,-9.81)                      */
  int32_t v_index[WBC_NU];           /* index of joint k in Drake's v (6..17); q index = +1    */
  int32_t act_index[WBC_NU];         /* actuator (output tau) index of joint k                 */
} wbc_model;

/* Gains and weights: the literals inside each reference ControlLaw. */
typedef struct wbc_params {
  /* ID  (inverse_dynamics_controller.py:116-128) */
  double id_kp_body_p, id_kd_body_p, id_kp_body_rpy, id_kd_body_rpy;
  double id_kp_foot, id_kd_foot, id_w_body, id_w_foot;
  /* CLF (clf_controller.py:65-78) */
  double clf_q_body_p, clf_q_body_pd, clf_q_body_rpy, clf_q_body_rpyd;
  double clf_q_foot_p, clf_q_foot_pd, clf_r, clf_w_delta;
  /* PC  (pc_controller.py:63-76) */
  double pc_kp_body_p, pc_kd_body_p, pc_kp_body_rpy, pc_kd_body_rpy;
  double pc_kp_foot, pc_kd_foot, pc_w_body, pc_w_foot;
  /* shared (inverse_dynamics_controller.py:19,94) */
  double mu;                 /* friction coefficient, 0.7                      */
  double contact_damping;    /* Kd of the no-slip constraint, 100              */
  /* declared tie-break: + reg_f/2 |f|^2 + reg_tau/2 |tau|^2
   * + reg_vd/2 |vd|^2 added to the reference cost so the optimum is unique.     */
  double reg_f, reg_tau, reg_vd;
  int32_t torque_limits;     /* 0 = reference QP; 1 = add |tau_k| <= effort_k  */
  int32_t max_iter;          /* active-set iteration cap                       */
  /* Basic PD law (basic_controller.py:331-350): tau = -kp (theta - theta_nom) - kd thetadot, clipped */
  double pd_kp, pd_kd, pd_clip;
  double pd_q_nom[WBC_NU];   /* nominal joint angles in the caller's joint order (q[7:19])       */
} wbc_params;

typedef struct wbc_handle wbc_handle;

/* Optional per-step buffers beyond tau/metrics/status; any pointer may be NULL. */
typedef struct wbc_io {
  const double* q;        /* [N][19] */
  const double* v;        /* [N][18] */
  const double* traj;     /* [N][54] */
  const uint8_t* contact; /* [N][4]  */
  double* tau;            /* [N][12] */
  double* metrics;        /* [N][4]  */
  int32_t* status;        /* [N]     */
  double* vd;             /* [N][18] QP accelerations, Drake velocity order (optional)  */
  double* f;              /* [N][12] contact forces LF RF LH RH, 0 for swing (optional) */
  double* qp_info;        /* [N][4]  objective, max primal violation, delta, #iters (optional) */
  double* lam;            /* [N][42] multipliers (>= 0) of the inequality rows at the returned optimum (optional):
                           *   [4 foot + t]  friction pyramid of foot LF RF LH RH, rows t = +x, -x, +y, -y
                           *                 (inverse_dynamics_controller.py:66-86, A_i row order)
                           *   [16], [17]    controller rows: CLF Vdot row (clf_controller.py:27-45) / PC passivity row
                           *                 (pc_controller.py:14-40, delta eliminated); [17] unused
                           *   [18 + a], [30 + a]   torque box  +tau_a <= effort_a,  -tau_a <= effort_a  (actuator order a)
                           * with these, x = [vd; tau; f] satisfies the KKT conditions of the reference QP (+ tie-break) */
} wbc_io;

/* Fill *p with the reference's constants. */
int wbc_default_params(wbc_params* p);

/* BasicController.__init__ (basic_controller.py:21-77): create the plant context
 * equivalent — uploads the model tables and gains to `device`. */
int wbc_create(const wbc_model* model, const wbc_params* params, int device, wbc_handle** out);
int wbc_destroy(wbc_handle* h);
const char* wbc_last_error(const wbc_handle* h);

/* BasicController.CalcDynamics (basic_controller.py:101-115),
 * CalcFramePositionQuantities x4 (:173-196) on device buffers. Outputs (any may be NULL):
 *   M [N][18][18], Cv [N][18], taug [N][18] (= -CalcGravityGeneralizedForces, :112),
 *   Jfeet [N][4][3][18], Jdv [N][4][3], pfeet [N][4][3]; all in Drake velocity order. */
int wbc_dynamics(wbc_handle* h, int64_t n, const double* q, const double* v,
                 double* M, double* Cv, double* taug, double* Jfeet, double* Jdv,
                 double* pfeet, void* stream);

/* BasicController.CalcCoriolisMatrix (basic_controller.py:117-132) and
 * CalcFrameJacobianDot x4 (:198-220): C [N][18][18] with C v = Cv, Jd [N][4][3][18]. */
int wbc_coriolis(wbc_handle* h, int64_t n, const double* q, const double* v,
                 double* C, double* Jd, void* stream);

int wbc_coriolis_host(wbc_handle* h, int64_t n, const double* q, const double* v, double* C, double* Jd);

/* One control step for n instances: DoSetControlTorques -> ControlLaw
 * (basic_controller.py:286-320; inverse_dynamics_controller.py:103-234;
 * clf_controller.py:48-234; pc_controller.py:43-255). Device pointers (or page-locked host memory mapped into the
 * device address space). The step uses per-handle scratch (4.5 KB per instance, at most 262144 instances at a time), so steps
 * of one handle are serialised on the device: a call on a different stream than the previous call of the handle first makes
 * its stream wait for the work submitted to the previous one (an event dependency, no host synchronisation). Host threads must
 * still serialise their calls on one handle; different handles are independent. Batches of 4096 - 65536 instances are issued as
 * two to four chunks, all but the first on library-owned streams that fork from and join `stream` through events: to the caller
 * the step remains one stream-ordered operation on `stream` (not inside a stream capture, where it stays a single chain). */
int wbc_step(wbc_handle* h, int kind, int64_t n, const wbc_io* io, void* stream);
int wbc_step_id(wbc_handle* h, int64_t n, const double* q, const double* v, const double* traj,
                const uint8_t* contact, double* tau, double* metrics, int32_t* status, void* stream);
int wbc_step_clf(wbc_handle* h, int64_t n, const double* q, const double* v, const double* traj,
                 const uint8_t* contact, double* tau, double* metrics, int32_t* status, void* stream);
int wbc_step_pc(wbc_handle* h, int64_t n, const double* q, const double* v, const double* traj,
                const uint8_t* contact, double* tau, double* metrics, int32_t* status, void* stream);
int wbc_step_mptc(wbc_handle* h, int64_t n, const double* q, const double* v, const double* traj,
                  const uint8_t* contact, double* tau, double* metrics, int32_t* status, void* stream);
/* BasicController.ControlLaw (basic_controller.py:322-352); traj / contact are ignored and may be NULL. */
int wbc_step_pd(wbc_handle* h, int64_t n, const double* q, const double* v, double* tau, void* stream);

/* Same step with HOST buffers (what the Python LeafSystem shim calls); returns after tau / metrics / status (and any
 * optional outputs that are non-NULL) are in the caller's buffers. Page-locked buffers (wbc_host_alloc /
 * cudaHostRegister): the kernels write the outputs straight into them; the inputs are read over the host link by the kernels
 * (zero-copy) below 24576 instances (131072 for PC / MPTC) and go through the copy engine into device staging in chunks above.
 * Pageable buffers go through a chunked two-stream copy / compute pipeline in both directions. */
int wbc_step_host(wbc_handle* h, int kind, int64_t n, const wbc_io* host_io);

/* All GPUs of the box from one call (north star: "instances shard trivially across the 8 GPUs of one box, with no NCCL
 * beyond an optional host gather"). wbc_multi_create makes one handle per listed device (devices == NULL: 0 ..
 * n_devices-1; a device may be listed more than once). wbc_multi_step_host splits [0, n) into contiguous shards whose sizes
 * differ by at most one, runs every shard on its device side by side (one host thread, asynchronous enqueue on all devices,
 * then one wait per device) and returns when all outputs are in the caller's host arrays - which is the host gather. Same
 * buffer rules as wbc_step_host; page-locked buffers (wbc_host_alloc) are mapped into every device. */
typedef struct wbc_multi wbc_multi;
int wbc_multi_create(const wbc_model* model, const wbc_params* params, int n_devices, const int* devices, wbc_multi** out);
int wbc_multi_destroy(wbc_multi* m);
const char* wbc_multi_last_error(const wbc_multi* m);
int wbc_multi_device_count(const wbc_multi* m);
int64_t wbc_multi_launch_count(const wbc_multi* m);
int wbc_multi_step_host(wbc_multi* m, int kind, int64_t n, const wbc_io* host_io);

/* Host-buffer version of wbc_dynamics (allocates device scratch per call; a test/debug entry). */
int wbc_dynamics_host(wbc_handle* h, int64_t n, const double* q, const double* v,
                      double* M, double* Cv, double* taug, double* Jfeet, double* Jdv, double* pfeet);

/* Pinned (page-locked) host memory for wbc_step_host buffers. */
void* wbc_host_alloc(size_t bytes);
void wbc_host_free(void* p);

/* Device-side timing helper for benchmarks: runs `reps` back-to-back wbc_step
 * launches on `stream` and returns the mean device time per launch (CUDA events
 * recorded on that same stream). */
int wbc_time_step(wbc_handle* h, int kind, int64_t n, const wbc_io* io, int reps, void* stream,
                  double* ms_per_launch);

/* Per-kernel share of a control step. A step is two stream-ordered kernels: the reduce kernel (state -> dynamics ->
 * reduced QP, one hand-over record per instance in library-owned device scratch) and the solve kernel (active-set QP ->
 * tau / metrics / status). Events are recorded before, between and after them; means over `reps` steps. n <= 262144. */
int wbc_profile_step(wbc_handle* h, int kind, int64_t n, const wbc_io* io, int reps, void* stream,
                     double* ms_reduce, double* ms_solve);

/* Measured-peak helper: FP64 FMA throughput of this device in TFLOP/s (a register
 * resident DFMA loop over all SMs), used as the FP64 roofline denominator. */
int wbc_measure_fp64_peak(int device, double* tflops);

/* ---- LCM wire codecs: the byte formats either side of the control step. Messages are packed back
 * to back, `msgs` must be 16-byte aligned; any optional pointer may be NULL. Per-message problems go to status[i]. */
#define WBC_LCM_TRUNK_STATE_BYTES 549   /* lcm_types/trunk_state_t.lcm:1-50: 8 fingerprint + 8 + 1 + 18*24 + 4 + 4*24 */
#define WBC_LCM_ROBOT_STATE_BYTES 204   /* lcm_types/robot_state_control_lcmt.lcm:1-7: 8 fingerprint + 4*(19+18+12)   */
#define WBC_WIRE_BADFINGERPRINT 1       /* the generated decoders raise ValueError("Decode error")                    */
#define WBC_WIRE_OVERFLOW 2             /* finite double beyond float32 range: struct.pack('>f') raises OverflowError */

/* trunk_state_t.decode (lcm_types/trunklcm/trunk_state_t.py:79-120) + the field copy of planners/towr.py:112-145:
 * n messages -> traj [N][54], contact [N][4] (wbc.h layout), optional timestamp [N], finished [N], planned forces
 * f_plan [N][12] (LF RF LH RH, unused by every controller), status [N]. Device pointers. */
int wbc_lcm_decode_trunk_state(wbc_handle* h, int64_t n, const uint8_t* msgs, double* timestamp, uint8_t* finished,
                               double* traj, uint8_t* contact, double* f_plan, int32_t* status, void* stream);
/* trunk_state_t.encode (trunk_state_t.py:54-77), what towr/trunk_mpc.cpp:19-68 publishes per sample. */
int wbc_lcm_encode_trunk_state(wbc_handle* h, int64_t n, const double* timestamp, const uint8_t* finished,
                               const double* traj, const uint8_t* contact, const double* f_plan, uint8_t* msgs, void* stream);
/* robot_state_control_lcmt.decode (lcm_types/cheetahlcm/robot_state_control_lcmt.py:35-47) as used by
 * BasicController.lcm_callback (basic_controller.py:79-87): float32 wire values widened to q [N][19], v [N][18],
 * optional tau [N][12]. */
int wbc_lcm_decode_robot_state(wbc_handle* h, int64_t n, const uint8_t* msgs, double* q, double* v, double* tau,
                               int32_t* status, void* stream);
/* robot_state_control_lcmt.encode (:24-33). q / v may be NULL (zeros: the controller publishes a fresh message that
 * carries only the torques, basic_controller.py:309-314). tau_in_actuator_order != 0: `tau` is the step output of
 * this library (actuator order) and is sent in velocity order like the reference's (S.T @ u)[-12:]; 0: copied as is. */
int wbc_lcm_encode_robot_state(wbc_handle* h, int64_t n, const double* q, const double* v, const double* tau,
                               int tau_in_actuator_order, uint8_t* msgs, int32_t* status, void* stream);
/* The same four with HOST buffers (copies inside, synchronous). */
int wbc_lcm_decode_trunk_state_host(wbc_handle* h, int64_t n, const uint8_t* msgs, double* timestamp, uint8_t* finished,
                                    double* traj, uint8_t* contact, double* f_plan, int32_t* status);
int wbc_lcm_encode_trunk_state_host(wbc_handle* h, int64_t n, const double* timestamp, const uint8_t* finished,
                                    const double* traj, const uint8_t* contact, const double* f_plan, uint8_t* msgs);
int wbc_lcm_decode_robot_state_host(wbc_handle* h, int64_t n, const uint8_t* msgs, double* q, double* v, double* tau,
                                    int32_t* status);
int wbc_lcm_encode_robot_state_host(wbc_handle* h, int64_t n, const double* q, const double* v, const double* tau,
                                    int tau_in_actuator_order, uint8_t* msgs, int32_t* status);

/* ---- Trunk-trajectory sampler: TOWR spline solution -> controller input on the device.
 * A plan is what towr's SplineHolder holds after the solve (towr/trunk_mpc.cpp:151-156): cubic Hermite node splines for
 * base position, base Euler angles (rpy), 4 foot positions, 4 foot forces, and the per-foot phase durations. */
#define WBC_TRAJ_CLAMPED 1   /* t outside [0, total]: evaluated at the nearest end (the reference asserts / runs off) */
#define WBC_TRAJ_BADPLAN 2   /* plan index out of range: plan 0 used                                                  */

typedef struct wbc_spline_desc {
  int32_t n_poly;            /* number of cubic polynomials                                                        */
  const double* durations;   /* [n_poly]           polynomial durations (Spline::GetPolyDurations)                 */
  const double* nodes;       /* [n_poly + 1][6]    node position (3) and velocity (3), towr Node                   */
} wbc_spline_desc;

typedef struct wbc_plan_desc {
  wbc_spline_desc base_linear, base_angular;   /* SplineHolder::base_linear_, base_angular_                        */
  wbc_spline_desc ee_motion[WBC_NLEG];         /* SplineHolder::ee_motion_ (LF RF LH RH)                           */
  wbc_spline_desc ee_force[WBC_NLEG];          /* SplineHolder::ee_force_                                          */
  int32_t n_phase[WBC_NLEG];                   /* PhaseDurations::GetPhaseDurations().size()                       */
  const double* phase_durations[WBC_NLEG];     /* [n_phase] alternating contact / swing phases of the foot         */
  uint8_t contact_at_start[WBC_NLEG];          /* GaitGenerator::IsInContactAtStart                                */
  /* Planner semantics of planners/towr.py:92-148 (optional, n_grid = 0 -> evaluate the splines at t itself):       */
  int32_t n_grid;                              /* number of stored samples (5001 for trunk_mpc's 1 kHz x 5 s)      */
  const double* grid_timestamps;               /* [n_grid] increasing timestamps of the stored samples             */
  double wait_time;                            /* stand this long before the motion starts (planners/towr.py:35)   */
  double standing[WBC_NTRAJ];                  /* SimpleStanding reference (planners/simple.py:39-85) while waiting */
} wbc_plan_desc;

typedef struct wbc_plan wbc_plan;

/* Upload n_plans plans (host descriptors) and build their device tables; Hermite coefficients
 * (towr/src/polynomial.cc:98-104) are computed on the device. */
int wbc_plan_create(wbc_handle* h, int32_t n_plans, const wbc_plan_desc* plans, wbc_plan** out);
int wbc_plan_destroy(wbc_plan* plan);
/* publish_trunk_state (towr/trunk_mpc.cpp:19-68) + TowrTrunkPlanner.SetTrunkOutputs (planners/towr.py:92-148) for n
 * (plan, time) pairs: traj [N][54], contact [N][4]; optional plan_index [N] (NULL = plan 0), f_plan [N][12],
 * t_eval [N] (the spline time actually evaluated, -1 while standing), status [N]. Device pointers. */
int wbc_sample_trajectory(wbc_handle* h, const wbc_plan* plan, int64_t n, const int32_t* plan_index, const double* t,
                          double* traj, uint8_t* contact, double* f_plan, double* t_eval, int32_t* status, void* stream);
int wbc_sample_trajectory_host(wbc_handle* h, const wbc_plan* plan, int64_t n, const int32_t* plan_index, const double* t,
                               double* traj, uint8_t* contact, double* f_plan, double* t_eval, int32_t* status);

/* Control step with the trunk targets taken from a device-resident plan (TowrTrunkPlanner.SetTrunkOutputs feeding
 * DoSetControlTorques, planners/towr.py:92-148 -> basic_controller.py:286-320) on HOST state buffers: the host sends q, v
 * and the plan time t [N] (+ optional plan_index [N]) - 308 B per instance instead of 736 B - the trajectory rows are sampled on
 * the device into library scratch. Outputs as wbc_step_host. Page-locked buffers are read / written by the kernels directly. */
int wbc_step_plan_host(wbc_handle* h, int kind, const wbc_plan* plan, int64_t n, const double* q, const double* v,
                       const double* t, const int32_t* plan_index, double* tau, double* metrics, int32_t* status);

/* ---- Closed-loop batched rollout: the caller of the control step, simulate.py:160-182.
 * wbc_integrate: semi-implicit Euler step of n states with the accelerations vd returned by wbc_step
 * (v += dt vd; q += dt N(q) v, quaternion renormalised); t [N] (optional) is advanced by dt. Device pointers. */
int wbc_integrate(wbc_handle* h, int64_t n, double dt, double* q, double* v, const double* vd, double* t, void* stream);

typedef struct wbc_rollout_io {
  double* q;              /* [N][19] in: initial state (simulate.py:171-179), out: final state                  */
  double* v;              /* [N][18]                                                                             */
  double* t;              /* [N]     in: start time of each instance, out: end time                              */
  const int32_t* plan_index; /* [N] or NULL (plan 0)                                                             */
  double* tau;            /* [N][12] torques of the last step                                                    */
  double* metrics;        /* [N][4]  metrics of the last step                                                    */
  int32_t* status_or;     /* [N]     OR of the per-step status words (optional)                                  */
  double* err_max;        /* [N]     max over the steps of the tracking-error metric (optional)                  */
  double* metrics_log;    /* [n_steps][N][4] every step's output_metrics, what simulate.py:142 logs (optional)   */
} wbc_rollout_io;

/* One time step of n simulated robots on flat ground (the plant half of simulate.py:36-57,160-182): forward dynamics from
 * the APPLIED torques, velocity-level ground contact at the four feet (no penetration, Coulomb pyramid with `mu`; projected
 * Gauss-Seidel, `iters` sweeps from zero; `erp` = fraction of a penetration pushed out per step), semi-implicit Euler. The
 * scheme is defined in csrc/wbc_plant.cuh - Drake's own contact solver is third party and not restatable - and pinned by
 * invariants and by oracle/rollout.py. tau [N][12] in actuator order; q, v updated in place; t (optional) advanced by dt;
 * ctrl_status (optional): robots with a non-zero controller status are frozen; status_or (optional) accumulates;
 * f_contact (optional) [N][12]: ground forces on LF RF LH RH in world axes (impulse / dt). Device pointers. */
typedef struct wbc_plant_opts {
  double mu;        /* ground friction, 1.0 (simulate.py:44-46: static = dynamic = 1.0) */
  double erp;       /* penetration correction per step, 0.2                            */
  int32_t iters;    /* Gauss-Seidel sweeps, 30                                         */
  int32_t reserved;
} wbc_plant_opts;
int wbc_default_plant_opts(wbc_plant_opts* o);
int wbc_plant_step(wbc_handle* h, int64_t n, double dt, const wbc_plant_opts* opts, double* q, double* v, const double* tau,
                   double* t, const int32_t* ctrl_status, int32_t* status_or, double* f_contact, void* stream);
int wbc_plant_step_host(wbc_handle* h, int64_t n, double dt, const wbc_plant_opts* opts, double* q, double* v, const double* tau,
                        double* t, const int32_t* ctrl_status, int32_t* status_or, double* f_contact);

/* n_steps control steps of n robots, entirely on the device: sample the plan at t (wbc_sample_trajectory), run the
 * controller `kind` (wbc_step), integrate (wbc_integrate). No host synchronisation inside; with use_graph != 0 the
 * per-step launches are captured once into a CUDA graph and replayed. Device pointers. */
int wbc_rollout(wbc_handle* h, int kind, const wbc_plan* plan, int64_t n, int32_t n_steps, double dt,
                const wbc_rollout_io* io, int use_graph, void* stream);
int wbc_rollout_host(wbc_handle* h, int kind, const wbc_plan* plan, int64_t n, int32_t n_steps, double dt,
                     const wbc_rollout_io* host_io, int use_graph);
/* The same loop with a choice of plant: plant == 0 integrates the QP's own contact-consistent accelerations ("planned contacts
 * hold", wbc_rollout); plant == 1 applies the controller's TORQUES to the simulated robot on the ground (wbc_plant_step), so a
 * controller can fail physically: slip, land, fall. f_contact (optional, device) receives the last step's ground forces. */
typedef struct wbc_rollout_opts {
  int32_t use_graph;
  int32_t plant;
  wbc_plant_opts plant_opts;
  double* f_contact;      /* [N][12] or NULL */
} wbc_rollout_opts;
int wbc_rollout_ex(wbc_handle* h, int kind, const wbc_plan* plan, int64_t n, int32_t n_steps, double dt,
                   const wbc_rollout_io* io, const wbc_rollout_opts* opts, void* stream);
int wbc_rollout_ex_host(wbc_handle* h, int kind, const wbc_plan* plan, int64_t n, int32_t n_steps, double dt,
                        const wbc_rollout_io* host_io, const wbc_rollout_opts* opts);

/* Number of kernel launches issued through this handle since creation. */
int64_t wbc_launch_count(const wbc_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* WBC_B200_H */
