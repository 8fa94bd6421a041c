/*
 * upright_b200 — C ABI of the batched waiter's-problem MPC solve on B200.
 *
 * This header is the drop-in boundary for the ONE hot path this repository
 * accelerates: `ControllerInterface::advanceMpc()` of utiasDSL/upright
 * (upright_control/src/pybindings.cpp:364-427 binds it; the OCP it solves is
 * assembled in upright_control/src/controller_interface.cpp:103-393), batched
 * over independent MPC instances.  Plain C types only; no torch / CUDA types.
 *
 * Conventions
 *   - all matrices row-major; all host-facing numbers are IEEE double like the
 *     reference (`ocs2::scalar_t`, upright_control/include/upright_control/types.h:16-35);
 *   - state  x = [q, v, a]            (nx = 3*nq; dimensions.h:10-46)
 *     input  u = [jerk, f_1 .. f_nc]  (nu = nq + nf*nc)
 *   - every function returns 0 on success or a negative UB_E_* code;
 *     `ub_last_error()` gives the message of the last failure on this thread.
 */
#ifndef UPRIGHT_B200_H
#define UPRIGHT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UB_MAX_JOINTS 9
#define UB_MAX_BODIES 8
#define UB_MAX_CONTACTS 32
#define UB_MAX_SPHERES 16
#define UB_MAX_PAIRS 32
#define UB_MAX_DYNAMIC_OBSTACLES 4
#define UB_MAX_PROJECTILE_LINKS 8
#define UB_MAX_NX (3 * UB_MAX_JOINTS)
#define UB_BODY_PARAMS 10 /* m, m*com(3), vech(I)(6): rigid_body.h:36-54 */
#define UB_MAX_GATHER 8   /* peer GPUs whose gathered buffers a solve writes into (ub_set_gather_targets) */
#define UB_STATS 8

enum {
    UB_OK = 0,
    UB_E_INVALID = -1,   /* bad argument / unsupported setting            */
    UB_E_CUDA = -2,      /* CUDA runtime error (message in ub_last_error) */
    UB_E_NO_DEVICE = -3, /* no CUDA device: there is NO CPU fallback       */
    UB_E_ALLOC = -4
};

/* per-instance solve status written to `status[b]` */
enum {
    UB_STATUS_CONVERGED = 0,   /* QP converged, line search accepted a step */
    UB_STATUS_QP_MAXITER = 1,  /* QP iteration cap hit (step still taken)   */
    UB_STATUS_LS_FAILED = 2,   /* no step size accepted: iterate unchanged  */
    UB_STATUS_NAN = 3          /* non-finite value encountered              */
};

enum { UB_JOINT_REVOLUTE = 0, UB_JOINT_PRISMATIC = 1 };

/* One single-dof joint of the serial chain.  Link i is the frame after joint
 * i.  The Pinocchio model it replaces: composite PX,PY,RZ root + URDF chain
 * (upright_control/include/upright_control/util.h:27-63). */
typedef struct ub_joint {
    int32_t type; /* UB_JOINT_* */
    int32_t reserved;
    double R[9];    /* fixed rotation parent link -> joint frame */
    double p[3];    /* fixed translation, parent-link coordinates */
    double axis[3]; /* unit motion axis, joint-frame coordinates  */
} ub_joint_t;

/* upright::ContactPoint (upright_core/include/upright_core/contact.h:10-48).
 * body index -1 = a fixture (the tray/EE) that gets no dynamics rows. */
typedef struct ub_contact {
    int32_t body1;
    int32_t body2;
    double mu;
    double r_co_o1[3];
    double r_co_o2[3];
    double normal[3]; /* points into object 1 */
    double span[6];   /* 2x3, rows orthogonal to normal */
} ub_contact_t;

/* Collision sphere rigidly attached to link `link` (0..nq-1), to the tool
 * frame (link == nq), to the world (link == -1), or riding on dynamic obstacle j
 * (link == -2 - j; its centre is the position block of that obstacle's state).  All collision geometry of
 * the reference is spheres (upright_assets/thing/xacro/collision_links.urdf.xacro:31-184,
 * obstacles/simple.urdf.xacro:40-102). */
typedef struct ub_sphere {
    int32_t link;
    int32_t shape;  /* UB_SHAPE_SPHERE, or UB_SHAPE_HALFSPACE: the `ground` object the reference adds to every collision
                       model (add_ground_plane, controller_interface.cpp:93-101,189: hpp::fcl::Halfspace(UnitZ, 0)) —
                       link = -1, offset = unit normal n, radius = plane offset d; the solid is {p : n.p <= d} and the
                       row of a pair (sphere a, half-space) is n.c_a - d - r_a - minimum_distance >= 0 */
    double radius;
    double offset[3];
} ub_sphere_t;
enum { UB_SHAPE_SPHERE = 0, UB_SHAPE_HALFSPACE = 1 };

typedef struct ub_pair {
    int32_t a;
    int32_t b;
} ub_pair_t;

/* HPIPM slack settings surface (upright_control/src/upright_control/wrappers.py:121-143) */
typedef struct ub_slack_settings {
    int32_t enabled;
    int32_t input_box;
    int32_t state_box;
    int32_t poly_ineq;
    double upper_L2_penalty;
    double lower_L2_penalty;
} ub_slack_settings_t;

/* Immutable per-configuration data: what ControllerSettings carries into
 * ControllerInterface (upright_control/include/upright_control/controller_settings.h:47-119). */
typedef struct ub_problem_desc {
    int32_t nq;        /* robot joints (6 fixed-base UR10, 9 Thing) */
    int32_t nb;        /* balanced bodies                            */
    int32_t nc;        /* contact points                             */
    int32_t nf;        /* force dimension per contact: 1 or 3        */
    int32_t N;         /* knots = time_horizon / sqp.dt              */
    int32_t n_spheres;
    int32_t n_pairs;
    int32_t sqp_iteration; /* SQP iterations per solve (controller.yaml:56)  */
    int32_t qp_iter_max;   /* sqp.hpipm.iter_max: Riccati solves per QP      */
    int32_t balancing_enabled;
    int32_t obstacles_enabled;
    int32_t qp_method;     /* 0 = interior point (product), 1 = semismooth Newton (oracle cross-check only) */
    double dt;

    ub_joint_t joints[UB_MAX_JOINTS];
    double tool_R[9]; /* last link -> end_effector_link_name frame */
    double tool_p[3];

    double gravity[3];
    double state_weight[UB_MAX_NX];  /* diag(Q)  (controller_interface.cpp:400-420) */
    double input_weight[UB_MAX_JOINTS]; /* diag(R) on the jerk                      */
    double ee_weight[6];             /* diag(W) of the end-effector cost: position (3), orientation error (3;
                                        end_effector_cost.h:61-81, zero in every shipped configuration) */
    double force_weight;             /* balancing.force_weight                   */
    double xd[UB_MAX_NX];            /* desired joint state                      */

    double state_lb[UB_MAX_NX], state_ub[UB_MAX_NX];
    double input_lb[UB_MAX_JOINTS], input_ub[UB_MAX_JOINTS];
    double force_lb, force_ub; /* controller_interface.cpp:330-356 */

    double body_params[UB_MAX_BODIES][UB_BODY_PARAMS]; /* map-key order   */
    ub_contact_t contacts[UB_MAX_CONTACTS];

    ub_sphere_t spheres[UB_MAX_SPHERES];
    ub_pair_t pairs[UB_MAX_PAIRS];
    double minimum_distance;

    ub_slack_settings_t slacks;

    /* QP / SQP numerics (documented choices) */
    double rho_hard;     /* proximal/AL penalty of hard equality rows (unit-normalised rows) */
    double qp_mu0;       /* initial complementarity t*lambda                */
    double qp_thr0;      /* floor of the initial inequality slacks t        */
    double qp_mu_target; /* central-path point the QP is solved to          */
    double qp_tol;       /* residual tolerance (inequality / equality rows) */
    double reg_input;    /* Levenberg term on the input Hessian             */
    /* filter line search (ocs2_sqp defaults) */
    double alpha_decay, alpha_min, g_max, g_min, gamma_c, armijo_factor;
    double delta_tol, cost_tol; /* controller.yaml:58-59 */

    /* EndEffectorBoxConstraint (constraint/end_effector_box_constraint.h:12-88; controller.yaml:92-95):
     * xyz_lower <= r(x) - r_d(t) <= xyz_upper at every intermediate knot, an inequality of the
     * "poly_ineq" family (softened with it) */
    int32_t ee_box_enabled;
    int32_t reserved0;
    double ee_box_lower[3], ee_box_upper[3];

    /* InertialAlignmentCostGaussNewton (inertial_alignment.h:146-204, src/inertial_alignment.cpp:90-163;
     * added to the problem at controller_interface.cpp:296-305): intermediate cost
     * 1/2 w |S C_we' (a - g) / |g||^2 with the Gauss-Newton Hessian w J'J; S = contact_plane_span (2x3) */
    int32_t ia_cost_enabled;
    int32_t reserved1;
    double ia_cost_weight;
    double ia_span[6];
    /* InertialAlignmentConstraint (inertial_alignment.h:68-110, src/inertial_alignment.cpp:7-53;
     * controller_interface.cpp:306-315): five inequality rows of the "poly_ineq" family at the intermediate
     * knots, a = C_we'(acc - g) [+ ddC_we com | = C_we' n], h = [a_n, alpha a_n -+ a_t0 -+ a_t1] >= 0 */
    int32_t ia_constraint_enabled;
    int32_t ia_use_angular_acceleration;
    int32_t ia_align_with_fixed_vector;
    int32_t reserved2;
    double ia_alpha;
    double ia_normal[3];
    double ia_com[3];

    /* Dynamic obstacles (constraint/obstacle_constraint.h:8-43, controller_interface.cpp:54-82,194-210;
     * dynamics/system_dynamics.h:28-38): each appends [p, v, a] (9 values) to the state, x = [x_robot, x_obs...],
     * with the uncontrolled constant-acceleration model p' = v, v' = a, a' = 0.  State dimension as seen through
     * the ABI = 3 nq + 9 n_dynamic_obstacles (ub_problem_dims out[0]); x0 and X carry that many columns. */
    int32_t n_dynamic_obstacles;
    int32_t reserved3;

    /* ProjectilePathConstraint (constraint/projectile_path_constraint.h:46-156; added to the problem at
     * controller_interface.cpp:272-294, settings wrappers.py:252-265): one inequality row per listed collision
     * link at the intermediate knots, "poly_ineq" family, behind the inertial-alignment rows:
     *     h_i = (scale / d_i) s (|c_i(x) - r_closest| - d_i) >= 0
     * c_i = origin of the link's collision frame (= centre of collision sphere projectile_spheres[i]),
     * r_closest = point of the LAST dynamic obstacle's ballistic path p + t v + t^2 a / 2 nearest to c_i over t >= 0
     * (cubic stationarity condition, <= 10 Newton steps from t = 0, step tolerance 1e-4; :11-44), Jacobian with t
     * held fixed (:120-146).  s = last element of the first target state (wrappers.py:36-42; raised to 1 by
     * mrt_node.cpp:241-252 while the projectile is in flight): rows are identically zero while s = 0 and
     * r_closest = p (t = 0) unless s > 0.5.  Change it between solves with ub_set_option("projectile_active"). */
    int32_t projectile_enabled;
    int32_t n_projectile_links;
    int32_t projectile_spheres[UB_MAX_PROJECTILE_LINKS];
    double projectile_distances[UB_MAX_PROJECTILE_LINKS];
    double projectile_scale;
    double projectile_active;
} ub_problem_desc_t;

typedef struct ub_problem ub_problem_t;

/* Flags for ub_solve_batch */
#define UB_PTRS_DEVICE 0x1u /* all batch pointers are device pointers (f32)  */
#define UB_WARM_START 0x2u  /* X/U hold the previous solution on entry       */
#define UB_COMPUTE_F64 0x4u /* validation build: run the kernels in fp64     */
#define UB_RESCUE_F64 0x8u  /* host mode: instances the fp32 kernels end with UB_STATUS_NAN are solved again by
                               the fp64 kernels inside the same call (from the same starting iterate) */

const char* ub_last_error(void);
int ub_version(void);

/* Replaces: ControllerInterface::ControllerInterface(settings)
 * (upright_control/src/controller_interface.cpp:103-393). Validates the
 * description, uploads constants to the current CUDA device. */
int ub_problem_create(const ub_problem_desc_t* desc, ub_problem_t** out);
void ub_problem_destroy(ub_problem_t* problem);

/* Dimensions helper: nx, nu, rows per knot etc. out[0..7] =
 * {nx, nu, n_eq, n_ineq_poly, n_terminal, N, nb, nc}. */
int ub_problem_dims(const ub_problem_t* problem, int32_t out[8]);

/* Bytes of device workspace ub_solve_batch needs for a batch of B. */
int64_t ub_workspace_bytes(const ub_problem_t* problem, int32_t B, uint32_t flags);

/* Replaces: B independent calls of
 *   mpc.setObservation(t, x, u); mpc.advanceMpc(); mpc.getMpcSolution(...)
 * (upright_control/src/pybindings.cpp:369-377; manager.py:156-170).
 *
 *   x0      [B, nx]          observed state per instance
 *   target  [B, N+1, 3]      desired EE position at each knot time
 *                            (interpolate_end_effector_pose, reference_trajectory.h:18-47); [B, N+1, 7] with the
 *                            desired quaternion [x y z w] behind it when ee_weight[3..5] != 0
 *   body_params [B, nb, 10]  per-instance inertial parameters or NULL (shared)
 *   X  [B, N+1, nx], U [B, N, nu]  solution (in/out when UB_WARM_START)
 *   K  [B, N, nu, 3 nq] Riccati feedback gains or NULL (over the ROBOT state: the dynamic-obstacle states that
 *      ub_problem_dims counts in nx are uncontrolled and carry no gain columns)
 *   status [B] int32, stats [B, UB_STATS] or NULL
 *      stats = {qp_iters, cost, violation, step alpha, qp_residual,
 *               max |object-dynamics eq|, min ineq margin, reserved}
 *
 * Host mode (default): pointers are host `double` arrays; copies to/from the
 * device happen inside the call (synchronous on `stream`).
 * Device mode (UB_PTRS_DEVICE): pointers are device arrays of `float`
 * (`double` with UB_COMPUTE_F64); the call only enqueues work on `stream`.
 * `workspace` may be NULL in host mode (allocated and cached internally).
 */
int ub_solve_batch(ub_problem_t* problem, int32_t B, const void* x0, const void* target,
                   const void* body_params, void* X, void* U, void* K, int32_t* status,
                   void* stats, void* workspace, int64_t workspace_bytes, uint32_t flags,
                   void* cuda_stream);

/* Replaces the named probes getStateInputEqualityConstraintValue("object_dynamics"),
 * getStateInputInequalityConstraintValue("contact_forces" | "obstacle_avoidance")
 * and getCostValue (controller_python_interface.h:31-88), batched: evaluates
 * at M (x,u) pairs on the device.  Host double pointers.
 *   name in {"object_dynamics","contact_forces","obstacle_avoidance",
 *            "end_effector_box_constraint" (needs target),"end_effector_position","cost",
 *            "inertial_alignment_cost","inertial_alignment_constraint","projectile_constraint",
 *            "end_effector_jacobian" (3 x nq, row-major), "object_dynamics_jacobian" (neq x (nx + nu): [dg/dx | dg/du],
 *            what BalancingConstraintWrapper::getLinearApproximation exposes, balancing_constraint_wrapper.h:45-60)}
 *   out [M, rows]; rows returned through *rows_out. */
int ub_eval(ub_problem_t* problem, const char* name, int32_t M, const double* x,
            const double* u, const double* target /*[M,3] or NULL*/,
            const double* body_params /*[M,nb,10] or NULL*/, double* out, int32_t out_capacity,
            int32_t* rows_out);

/* Closed-loop rollout settings: the knobs of the simulation loop
 * upright_cmd/scripts/simulations/mpc_sim.py:118-160 around ControllerManager.step
 * (upright_control/src/upright_control/manager.py:156-176). */
typedef struct ub_closed_loop_params {
    double sim_dt;              /* simulation step (upright_cmd/config/simulation.yaml:7)           */
    double replan_period;       /* tracking.min_policy_update_time (controller.yaml:33)             */
    int32_t n_steps;            /* simulation steps to run                                          */
    int32_t log_stride;         /* keep every log_stride-th step in xs / us (>= 1)                   */
    int32_t use_feedback;       /* sqp.use_feedback_policy (controller.yaml:60)                     */
    int32_t cold_start;         /* mpc.cold_start: no warm start between replans                    */
    int32_t init_sqp_iteration; /* SQP iterations of the first solve (controller.yaml:57)           */
    int32_t sqp_iteration;      /* SQP iterations of every later solve (controller.yaml:56)         */
    double kp, kv, ka;          /* tracking gains Kx = [kp I, kv I, ka I] (mpc_sim.py:101-108)      */
} ub_closed_loop_params_t;

/* Replaces, for B robots at once, the closed loop
 *     for each simulation step:  xd, u = ctrl_manager.step(t, x);  u_cmd = Kx (xd - x) + u;  integrate
 * (mpc_sim.py:118-160): replan gate `t >= last_planning_time + replan_period`, warm start from the previous
 * solution shifted to the new time grid, desired end-effector position interpolated between the waypoints
 * (reference_trajectory.h:18-47), policy evaluation u_ff(t) + K(t)(x - x_nom(t)) with linear interpolation
 * (pybindings.cpp:378-381), and the model's own triple integrator as the plant.  Everything stays on the
 * device between the first upload and the final download.  Host double pointers.
 *   x0 [B, nx]; target_times [M] increasing; target_pos [B, M, 3]; body_params [B, nb, 10] or NULL
 *   (nx as ub_problem_dims reports it: 3 nq + 9 per dynamic obstacle; the obstacle columns of x0 are their states at
 *   the start, and they evolve by ub_closed_loop_set_obstacles)
 *   xs [B, n_log, nx], us [B, n_log, nq] (n_log = ceil(n_steps / log_stride)) or both NULL
 *   x_final [B, nx] or NULL; *n_replans or NULL; status_counts [B, 4] (solves per UB_STATUS_*) or NULL */
int ub_closed_loop(ub_problem_t* problem, int32_t B, const double* x0, const double* target_times,
                   const double* target_pos, int32_t M, const double* body_params,
                   const ub_closed_loop_params_t* params, double* xs, double* us, double* x_final,
                   int32_t* n_replans, int32_t* status_counts, uint32_t flags, void* cuda_stream);

/* Simulated dynamic obstacles of ub_closed_loop (the uncontrolled obstacles of the reference's simulation,
 * upright_sim/src/upright_sim/simulation.py:300-435, configured under `simulation.dynamic_obstacles.obstacles`, e.g.
 * upright_cmd/config/obstacles/dynamic.yaml:38-75): free flight under the current mode's acceleration; at the first
 * simulation step whose start time has reached the next mode's `time` the state is reset to that mode's position
 * (+ the instance's offset, for `relative` obstacles) and velocity.  The obstacle columns of the plant state are what
 * the controller observes at every replan (mpc_sim.py:120-121).  One entry per dynamic obstacle of the problem, in
 * order; modes [n_obstacles][UB_MAX_OBSTACLE_MODES]; offsets [B, n_obstacles, 3] or NULL.  Without a plant the
 * obstacle states of x0 fly freely under their own acceleration.  n_obstacles = 0 removes the plant. */
#define UB_MAX_OBSTACLE_MODES 8
typedef struct ub_obstacle_mode {
    double time;
    double position[3], velocity[3], acceleration[3];
} ub_obstacle_mode_t;
int ub_closed_loop_set_obstacles(ub_problem_t* problem, int32_t n_obstacles, const int32_t* n_modes,
                                 const ub_obstacle_mode_t* modes, int32_t B, const double* offsets);

/* Multi-GPU gather fused into the solve (no counterpart in the reference, which is single-process):
 * after this call every device-mode ub_solve_batch stores the trajectories of instance b not only to its X / U
 * arguments but also to row `row_offset + b` of each of the `n` (<= UB_MAX_GATHER) peer buffers
 *     X_bases[p] : [rows, N+1, nx],   U_bases[p] : [rows, N, nu]      (same element type as X / U)
 * which are device pointers of OTHER GPUs mapped into this process (CUDA IPC; peer access enabled) — the solve
 * kernel's epilogue writes them over NVLink while the rest of the batch is still being solved, so the all-gather
 * of the results needs no collective kernel and no copy.  n = 0 switches it off.  The caller synchronises the
 * ranks (one barrier) before any rank reads its gathered buffer. */
int ub_set_gather_targets(ub_problem_t* problem, int32_t n, void* const* X_bases, void* const* U_bases,
                          int64_t row_offset);

/* Gathered buffers for ub_set_gather_targets, shared between the processes (one per GPU) of a node.  The owner
 * allocates device memory on ITS current device and gets an opaque handle to pass to its peers by any host channel;
 * a peer opens the handle with ITS OWN device current (lazy peer access), which maps the owner's memory for stores from
 * the peer's kernels over NVLink.  Close every opened mapping before the owner frees the buffer. */
#define UB_IPC_HANDLE_BYTES 64
int ub_gather_alloc(int64_t bytes, void** ptr, unsigned char handle[UB_IPC_HANDLE_BYTES]);
int ub_gather_open(const unsigned char handle[UB_IPC_HANDLE_BYTES], void** ptr);
int ub_gather_close(void* ptr);
int ub_gather_free(void* ptr);

/* Runtime options: "sqp_iteration" (init_sqp_iteration vs sqp_iteration,
 * controller.yaml:56-57), "projectile_active" (the target-state flag s of the
 * projectile path constraint, 0 or 1) and the test aid "stop_after" (0 full
 * solve, 1 stop after the first linearisation, 2 after the first QP). */
int ub_set_option(ub_problem_t* problem, const char* key, int value);

/* Per-instance workspace layout for tests that inspect intermediate blocks: offsets in units of the kernels'
 * matrix type (float; double with UB_COMPUTE_F64); blocks holding doubles take `rw` units per element.
 * Field order: upright_b200/engine.py LAYOUT_FIELDS. */
int ub_workspace_layout(const ub_problem_t* problem, uint32_t flags, int32_t out[80]);

/* Device time of the last ub_solve_batch on this problem (CUDA events), ms.
 * Replaces getLastSolveTime() (controller_python_interface.h:27-29). */
float ub_last_solve_ms(const ub_problem_t* problem);

/* Measured FP32 multiply-add throughput of the current device in TFLOP/s (a register-resident FMA loop on every SM):
 * the denominator of the arithmetic roofline bench.py reports for the solve kernel. */
int ub_measure_fma_peak(double* tflops);

/* Number of kernel launches issued by this library since load. */
int64_t ub_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* UPRIGHT_B200_H */
