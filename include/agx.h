/*
 * agx.h — C ABI of the B200-native batched OCP solve path.
 *
 * This is the drop-in boundary for the arithmetic that agimus_controller reaches
 * through Boost.Python today (reference paths relative to /root/reference):
 *
 *   crocoddyl.ShootingProblem(x0, running, terminal)   agimus_controller/agimus_controller/ocp_base_croco.py:55-62
 *   solver.solve(xs, us, max_iters)                    agimus_controller/agimus_controller/ocp_base_croco.py:172
 *   problem.calc / problem.calcDiff                    agimus_controller_ros/agimus_controller_ros/mpc_debugger_node.py:300-301
 *   problem.rollout(us)                                agimus_controller/tests/test_warm_start_shift_previous_reference.py:76
 *   runningModels[0].calc(data, x, u) -> xnext         agimus_controller/agimus_controller/ocp_base_croco.py:184-189
 *   per-tick reference / weight setters                agimus_controller/agimus_controller/ocp/ocp_croco_generic.py:855-892
 *   pin.rnea warm start                                agimus_controller/agimus_controller/warm_start_reference.py:77-87
 *
 * Conventions
 *  - every array is fp64, C order, and lives in DEVICE memory unless the name ends in _host;
 *  - the caller owns every buffer; the library borrows pointers for the duration of the
 *    stream-ordered call and never allocates inside agx_solve;
 *  - every entry point returns 0 on success, a negative AGX_E* code otherwise; no C++
 *    exception crosses the boundary; agx_last_error() gives a message;
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *  - state x = [q (nv); v (nv)], nx = ndx = 2 nv, nu = nv (full actuation, vector-space state);
 *  - policy convention (Crocoddyl): u = us - k - K (x - xs), K is [nu, nx] row-major.
 */
#ifndef AGX_H_
#define AGX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGX_MAX_NV 16
#define AGX_MAX_CAPSULES 4
#define AGX_MAX_COLLISION_PAIRS 2

/* joint types */
#define AGX_JOINT_REVOLUTE 0
#define AGX_JOINT_PRISMATIC 1

/* error codes */
#define AGX_OK 0
#define AGX_EINVAL -1       /* bad argument */
#define AGX_EUNSUPPORTED -2 /* model / problem shape the kernels do not cover */
#define AGX_ECUDA -3        /* CUDA runtime error */
#define AGX_ENOMEM -4

/* per-problem solver status (out_status) */
#define AGX_STATUS_CONVERGED 0 /* stop criterion met with a feasible trajectory */
#define AGX_STATUS_MAXITER 1   /* iteration budget used */
#define AGX_STATUS_REGMAX 2    /* regularisation hit reg_max */
#define AGX_STATUS_NAN 3       /* non-finite value met */
#define AGX_STATUS_LINESEARCH 4 /* reserved (a refused line search raises the regularisation and the solve goes on) */
#define AGX_ACCEPT_CROCODDYL2 0
#define AGX_ACCEPT_CROCODDYL1 1
#define AGX_STATUS_TIMEOUT 5    /* max_solve_time elapsed: the current iterate is returned */

/*
 * Kinematic-tree table: what factory/robot_model.py:88-351 (RobotModels.robot_model, .armature)
 * flattens to.  Joint i moves body i; parent[i] < i, -1 = world.  placement = joint frame in the
 * parent body frame; inertia about the COM in body axes, order xx xy xz yy yz zz.
 * The task frame (ResidualModelFramePlacement id, ocp_croco_generic.py:197-219) is attached to
 * body frame_parent with placement (frame_R, frame_p).
 */
typedef struct agx_model {
  int32_t nv;
  int32_t frame_parent;
  int32_t parent[AGX_MAX_NV];
  int32_t jtype[AGX_MAX_NV];
  double axis[AGX_MAX_NV][3];
  double placement_R[AGX_MAX_NV][9];
  double placement_p[AGX_MAX_NV][3];
  double mass[AGX_MAX_NV];
  double com[AGX_MAX_NV][3];
  double inertia[AGX_MAX_NV][6];
  double armature[AGX_MAX_NV];
  double gravity[3];
  double frame_R[9];
  double frame_p[3];
  /* Collision geometry for ResidualDistanceCollision (ocp_croco_generic.py:495-533; capsules as built by
   * factory/robot_model.py:261-302): a capsule is a segment [a0, a1] with a radius, given in the frame of the body
   * it is attached to (cap_parent >= 0) or in the world frame (cap_parent = -1).  A pair (pair_a[k], pair_b[k])
   * contributes the residual r = distance between the two capsule surfaces, with the activation exp(-r^2 / alpha)
   * (ActivationModelQuadExp, ocp_croco_generic.py:117-143) and a per-node weight (reference record, below). */
  double cap_a0[AGX_MAX_CAPSULES][3];
  double cap_a1[AGX_MAX_CAPSULES][3];
  double cap_radius[AGX_MAX_CAPSULES];
  double col_alpha;
  int32_t n_capsules;
  int32_t cap_parent[AGX_MAX_CAPSULES];
  int32_t n_pairs;
  int32_t pair_a[AGX_MAX_COLLISION_PAIRS];
  int32_t pair_b[AGX_MAX_COLLISION_PAIRS];
  /* Form of the task-frame residual (the "pose" slot of the reference record):
   *   AGX_POSE_PLACEMENT (0): r = log6(Mref^-1 oMf) -- ResidualModelFramePlacement[Static]
   *                           (ocp_croco_generic.py:197-249).  With zero weights on its linear part this is also
   *                           ResidualModelFrameRotation[Static] (:306-357): r_ang = log3(Rref^T oRf), same Jacobian rows;
   *   AGX_POSE_TRANSLATION_WORLD (1): linear part r = p_f - pref in the WORLD frame, Rq = oRf fJf[:3] --
   *                           ResidualModelFrameTranslation[Static] (:252-303); the angular part stays log3(Rref^T oRf), so
   *                           a FrameTranslation cost and a FrameRotation cost on the same frame share the record. */
  int32_t pose_mode;
  int32_t reserved_;
} agx_model;
#define AGX_POSE_PLACEMENT 0
#define AGX_POSE_TRANSLATION_WORLD 1

/*
 * FDDP parameters (Crocoddyl SolverFDDP defaults are what agx_fddp_opts_default fills).
 * fixed_iters != 0: run exactly max_iter iterations, no early exit (benchmark mode).
 * eager_exit != 0 (and not fixed_iters): latency mode of the single MPC tick (B <= 64; ignored for larger batches): no
 * kernel runs past the iteration in which the last problem finishes.
 *   agx_solve (chain and general-tree kernels) and agx_solve_sqp on the 7-joint chain: the whole solve is ONE graph launch -- init, (FDDP: first costs,)
 *   a conditional WHILE node around the iteration whose condition the iteration's last kernel sets on the device (in
 *   agx_solve_sqp the line search is a second WHILE node nested in it, armed per iteration), (SQP: final sweep,)
 *   finalize -- so there is no host round trip between iterations and the call stays stream-ordered (no
 *   synchronisation; ticks may be queued back to back).  The caller's pointers reach the graph through a small device
 *   table refreshed by one copy per call.  AGX_TICK_GRAPH=0 in the environment, a timing run (agx_set_timing) or a
 *   driver that refuses conditional nodes select the stream path below; both give the same bits
 *   (tests/test_gpu_tick_graph.py).
 *   agx_solve_sqp on other trees, and the stream path: after every iteration (and, in agx_solve_sqp, after every
 *   step length of the line search) a completion flag / counter is read back (one small device-to-host copy and a
 *   stream synchronisation each) and the call returns as soon as every problem has finished.
 * agx_solve uses the same graph, at any batch size, for budgets above 32 iterations without fixed_iters (a batch solved
 * to convergence, the controller's first solve with max_iter = 1000): the loop ends with the last problem.
 * In latency mode the FDDP forward pass of the chain runs on a kernel that puts two warps on each problem group and the
 * cost records come from the octet path of the derivative kernel instead of the thread-per-node kernel; the results
 * agree with the throughput kernels' to rounding (1e-15 per operation), not bitwise.
 */
typedef struct agx_fddp_opts {
  double reg_min, reg_max, reg_incfactor, reg_decfactor;
  double th_grad, th_stepdec, th_stepinc, th_acceptstep, th_acceptnegstep, th_stop;
  double reg_init; /* NaN -> reg_min */
  int32_t fixed_iters;
  int32_t n_alphas; /* step lengths 2^-n, n = 0..n_alphas-1 (<= 10) */
  int32_t eager_exit;
  /* Acceptance test of a trial step (SolverFDDP::solve).  Crocoddyl's releases differ in two details and its source is
   * not in the reference tree, so both forms are provided:
   *   AGX_ACCEPT_CROCODDYL2 (0, default; Crocoddyl >= 2.0, what the reference's flake pins):
   *       dVexp >= 0: |d1| < th_grad or dV > th_acceptstep dVexp;   dVexp < 0: NOT feasible and dV > th_acceptnegstep dVexp
   *   AGX_ACCEPT_CROCODDYL1 (1; Crocoddyl 1.x): d1 < th_grad without the absolute value, no feasibility guard.
   * They differ only for d1 < 0 (gap terms dominating) or for dVexp < 0 on a feasible candidate -- which exact
   * arithmetic excludes (feasible: d1 + d2/2 = Qu.k/2 >= 0); tests/test_gpu_parity.py checks that the benchmark batch
   * gives the same iterates under both. */
  int32_t accept_rule;
  /* max_solve_time of the reference's solver (ocp_base_croco.py:70-71, passed when use_iteration_limits_and_timeout,
   * :166-171), in seconds; <= 0: none.  The deadline is kept ON THE DEVICE: the solve's first kernel stamps the device
   * clock, and a problem whose iteration ends later than stamp + max_solve_time stops iterating and returns its current
   * iterate (status AGX_STATUS_TIMEOUT) -- no host synchronisation, the remaining launches find the problem finished.
   * The clock starts when the solve starts EXECUTING on the stream, not when it is queued. */
  double max_solve_time;
} agx_fddp_opts;

/*
 * Parameters of the SQP mode (mim_solvers.SolverCSQP as the reference configures it, ocp_base_croco.py:64-75, with no
 * constraint active): agx_sqp_opts_default fills sigma = 1e-6 (proximal term), reg = 1e-9 (regularisation at its floor
 * reg_min), mu = 10 (weight of the L1 gap norm in the merit function), termination_tolerance = 1e-3
 * (ocp_param_base.py:54-56), n_alphas = 10.
 */
typedef struct agx_sqp_opts {
  double sigma, reg, mu, termination_tolerance;
  int32_t n_alphas;
  int32_t eager_exit; /* as in agx_fddp_opts */
  double max_solve_time; /* as in agx_fddp_opts */
} agx_sqp_opts;

typedef struct agx_handle agx_handle;

/* Size in doubles of one node's reference record:
 *   [xref nx][wx nx][uref nu][wu nu][Rref 9][pref 3][wpose 6][wcol 2]
 * wx/wu/wpose are the activation weights already multiplied by the CostModelSum weight
 * (ocp_croco_generic.py:577-585, :688-691); an inactive cost has zero weights; wcol[k] is the scalar weight of
 * collision pair k (w_collision_avoidance, ocp_croco_generic.py:718-719), 0 when unused.  The terminal
 * node (t = T) ignores uref/wu.  Layout of `refs`: [B][T+1][agx_ref_size(nv)]. */
int agx_ref_size(int nv);

void agx_fddp_opts_default(agx_fddp_opts* opts);
void agx_sqp_opts_default(agx_sqp_opts* opts);

/* Build a handle for B problems with T running nodes on CUDA device `device`.
 * dts_host: T step sizes (host memory; ocp_param_base.py:67-78 timesteps).
 * models_host: 1 model shared by the batch (n_models = 1) or one per problem (n_models = B). */
int agx_create(const agx_model* models_host, int n_models, const double* dts_host, int B, int T,
               int device, agx_handle** out);
int agx_destroy(agx_handle* h);
const char* agx_last_error(const agx_handle* h);

/* Replaces the per-tick reference/weight setter loop (ocp_croco_generic.py:855-892).
 * refs: device [B][T+1][ref_size]; copied into the handle (stream ordered). */
int agx_set_refs(agx_handle* h, const double* refs, void* stream);

/* OCPBaseCroco.update_geometry_placement (ocp_base_croco.py:110-131; the controller calls it for every obstacle pose it
 * receives, agimus_controller_ros/agimus_controller_ros/agimus_controller.py:406): new end points (in the frame of the
 * capsule's parent joint, or of the world for an obstacle) and radius of collision capsule `capsule` of every model of
 * the handle.  a0, a1: HOST pointers to 3 doubles. */
int agx_set_capsule(agx_handle* h, int capsule, const double* a0, const double* a1, double radius, void* stream);

/* Device-side reference stream: instead of rebuilding the [B][T+1][ref_size] table on the host every tick, keep the
 * whole weighted reference trajectory on the device — stream_refs [n_streams][n_points][ref_size], n_streams = 1
 * (shared by the batch) or B — and select the horizon window: node t of problem b reads point
 * start_b + hidx[t], hidx = cumulative step factors dts[i]/dts[0] (TrajectoryBuffer.compute_horizon_indexes,
 * trajectory.py:199-215); indices past the end repeat the last point.  `start` is a device int32 [B] or NULL (then
 * start0 applies to every problem).  Replaces buffer.horizon + set_reference_weighted_trajectory (mpc.py:40-41,
 * ocp_croco_generic.py:855-892) and the rolling-buffer circularAppend. */
int agx_set_refs_window(agx_handle* h, const double* stream_refs, int n_streams, int n_points, const int32_t* start,
                        int start0, void* stream);

/* problem.calc: xs [B][T+1][nx], us [B][T][nu] -> out_cost [B][T+1] (node costs),
 * out_xnext [B][T+1][nx] (terminal row = xs_T). Either output may be NULL. */
int agx_calc(agx_handle* h, const double* xs, const double* us, double* out_cost,
             double* out_xnext, void* stream);

/* problem.calc + problem.calcDiff, dense per-node outputs (row-major):
 *  Fx [B][T+1][nx][nx], Fu [B][T+1][nx][nu], Lx [B][T+1][nx], Lu [B][T+1][nu],
 *  Lxx [B][T+1][nx][nx], Lxu [B][T+1][nx][nu], Luu [B][T+1][nu][nu].
 * Terminal rows hold Fx = I, Fu = 0, Lu = Lxu = Luu = 0.  Any output may be NULL. */
int agx_calc_diff(agx_handle* h, const double* xs, const double* us, double* out_cost,
                  double* out_xnext, double* Fx, double* Fu, double* Lx, double* Lu, double* Lxx,
                  double* Lxu, double* Luu, void* stream);

/* Per-cost evaluation (what mpc_debugger_node.py:294-323 reads off problem.runningDatas): for every node the
 * unscaled value of each named cost of ocp_goal_reaching.yaml and the frame-placement residual,
 * out_terms [B][T+1][AGX_N_COST_TERMS] = [state_reg, control_reg, goal_tracking, r6 (lin 3, ang 3),
 * collision cost of pair 0 / 1, signed distance of pair 0 / 1 (plots/plots_utils.py:160-208)]. */
#define AGX_N_COST_TERMS 13
int agx_cost_terms(agx_handle* h, const double* xs, const double* us, double* out_terms, void* stream);

/* Per-cost gradients (what mpc_debugger_node.py:303-323 reads: w * d.Lx, w * d.Lu of every named cost of
 * runningDatas[i].differential.costs, i.e. UNSCALED by the time step): out_Lx [B][T+1][AGX_N_COSTS][nx],
 * out_Lu [B][T+1][AGX_N_COSTS][nu] (either may be NULL), costs in the order [state_reg, control_reg, task frame
 * (goal_tracking), collision pair 0, collision pair 1], CostModelSum weights included (they are folded into the
 * reference record).  Computed by problem.calcDiff with the other costs' weights zeroed, one pass per cost. */
#define AGX_N_COSTS 5
int agx_cost_derivatives(agx_handle* h, const double* xs, const double* us, double* out_Lx, double* out_Lu, void* stream);

/* WarmStartShiftPreviousSolution.shift (warm_start_shift_previous_solution.py:85-104) for the whole batch:
 * nodes whose time step equals dts[0] take the next node's state and control (the last control is repeated),
 * coarser nodes are re-integrated over dts[0] with their own control; xs[T] is kept.  Out of place. */
int agx_shift_warmstart(agx_handle* h, const double* xs, const double* us, double* out_xs, double* out_us,
                        void* stream);

/* problem.rollout(us): x0 [B][nx], us [B][T][nu] -> out_xs [B][T+1][nx]. */
int agx_rollout(agx_handle* h, const double* x0, const double* us, double* out_xs, void* stream);

/* IntegratedActionModelEuler.calc for n independent (x, u) pairs with step dt
 * (ocp_base_croco.py:184-189; model 0 of the handle). x [n][nx], u [n][nu] -> out_xnext [n][nx]. */
int agx_integrate(agx_handle* h, const double* x, const double* u, double dt, int n,
                  double* out_xnext, void* stream);

/* pin.rnea(model, data, q, v, a) for n independent triples (warm_start_reference.py:77-87). */
int agx_rnea(agx_handle* h, const double* q, const double* v, const double* a, int n,
             double* out_tau, void* stream);

/* solver.solve(xs, us, max_iter) with problem.x0 = x0 (ocp_base_croco.py:158, :172).
 * In:  x0 [B][nx], xs_ws [B][T+1][nx], us_ws [B][T][nu].
 * Out: out_xs [B][T+1][nx], out_us [B][T][nu], out_K [B][T][nu][nx], out_k [B][T][nu] (may be NULL),
 *      out_cost [B], out_iters [B] (int32), out_status [B] (int32), out_stop [B] (may be NULL).
 * A problem whose trial step is rejected takes its next step length in the next round of launches (deferred line
 * search, up to twice per solve; deeper searches run in line), so max_iter + 2 rounds are queued; per problem the
 * iterates are those of the sequential search.
 * Stream-ordered and asynchronous for max_iter <= 32 or opts->fixed_iters; with a larger budget the call
 * synchronises `stream` every 16 iterations to stop as soon as every problem has finished. */
int agx_solve(agx_handle* h, const double* x0, const double* xs_ws, const double* us_ws,
              int max_iter, const agx_fddp_opts* opts, double* out_xs, double* out_us,
              double* out_K, double* out_k, double* out_cost, int32_t* out_iters,
              int32_t* out_status, double* out_stop, void* stream);

/* The solver the reference actually instantiates: mim_solvers.SolverCSQP(problem).solve(xs, us, max_iter)
 * (ocp_base_croco.py:64-75, :172) with no constraint active, i.e. a Gauss-Newton SQP.  Per iteration:
 * calc + calcDiff and the gaps; the equality-constrained QP solved by one Riccati sweep (regularisation `reg`) and a
 * LINEAR rollout du = -k - K dx, dx' = Fx dx + Fu du + fs'; KKT = max(|Lx + Fx^T l' - l|_inf, |Lu + Fu^T l'|_inf,
 * |fs|_inf) with the QP multipliers l; stop when KKT <= termination_tolerance (the iterate is returned as is);
 * otherwise the first step length 2^-n with merit(xs + a dx, us + a du) < merit(xs, us),
 * merit = cost + mu * |gaps|_1.  The regularisation follows SolverDDP's schedule, which the mim_solvers solvers
 * inherit: x10 after a failed factorisation or a step length <= 0.01 (a refused line search included), /10 after a
 * step length > 0.5, floor `reg`, AGX_STATUS_REGMAX at 1e9.  The gains returned are those of the solver's last backward pass: a sweep at the
 * final iterate with sigma + reg on Quu, Qxx, Vxx_T.  (With no constraint the ADMM/proximal inner loop sits at its
 * fixed point, so it is not iterated.)  Replaying the reference's golden test this way (zero warm start,
 * tests/test_ocp_croco_base.py:140-158) stops at the same criterion 6e-5 from the golden states and reproduces the
 * golden gains to 1e-11 at the golden point.
 * Out as agx_solve; out_stop = the KKT norm; out_iters = passes of the solver loop whose line search ran (the
 * solver's iter counter: a pass whose every step length is refused counts too); AGX_STATUS_CONVERGED = KKT met. */
int agx_solve_sqp(agx_handle* h, const double* x0, const double* xs_ws, const double* us_ws,
                  int max_iter, const agx_sqp_opts* opts, double* out_xs, double* out_us,
                  double* out_K, double* out_k, double* out_cost, int32_t* out_iters,
                  int32_t* out_status, double* out_stop, void* stream);

/* One calc + calcDiff at (xs, us) followed by ONE backward Riccati sweep with the fixed regularisation `reg`
 * added to Quu and to the diagonal of every Vxx (no retry on failure).  With reg = 1e-6 this is the backward
 * pass of the solver the reference instantiates (mim_solvers.SolverCSQP, ocp_base_croco.py:64, unconstrained
 * case, proximal sigma = 1e-6) and reproduces the gains of the reference's golden file
 * (agimus_controller/tests/test_ocp_croco_base.py:175-204).  Gaps are x0 - xs_0 and xnext_t - xs_{t+1}.
 * Out: out_K [B][T][nu][nx], out_k [B][T][nu] (may be NULL), out_status [B] (may be NULL;
 * AGX_STATUS_REGMAX = the sweep failed). */
int agx_riccati(agx_handle* h, const double* x0, const double* xs, const double* us, double reg, double* out_K,
                double* out_k, int32_t* out_status, void* stream);

/* Optional device timing of the solve's kernels (bench evidence, off by default): while enabled, agx_solve
 * brackets every launch of an iteration with a CUDA event pair on `stream`.  agx_get_timing synchronises on those
 * events, adds the elapsed milliseconds and launch counts per kernel into out_ms[5] / out_launches[5]
 * (0 = calc_diff, 1 = backward sweep, 2 = alpha-1 rollout, 3 = trial cost records, 4 = accept + line search)
 * and clears the record (at most 4096 launches are kept between two reads). */
int agx_set_timing(agx_handle* h, int enable);
int agx_get_timing(agx_handle* h, double* out_ms, long long* out_launches);

/* Roofline denominator: sustained FP64 FMA throughput of `device` (TFLOP/s, FMA = 2 flops) measured by
 * a register-only probe kernel run for about `seconds`; out_ms (may be NULL) = device time of the run. */
int agx_probe_fp64(int device, double seconds, double* out_tflops, double* out_ms);

/* Number of kernel launches the library issued on this handle since creation (bench evidence).  A tick-graph solve
 * (eager_exit, see agx_fddp_opts) counts the kernels outside the loop plus ONE round: how many further rounds ran is
 * decided on the device and reported per problem in out_iters. */
long long agx_launch_count(const agx_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* AGX_H_ */
