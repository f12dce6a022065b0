"""GPU parity: the sm_100a kernels, called through the C ABI (libagx.so), against the CPU oracle.

Tolerances are the ones BASELINE.json's north_star states: per-node derivatives within 1e-9 relative,
final xs / us / cost after a fixed iteration count within 1e-6 relative.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from agimus_controller_b200 import _abi, panda_table  # noqa: E402
from agimus_controller_b200.workloads import goal_reaching_batch, golden_problem  # noqa: E402

DERIV_RTOL = 1e-9
TRAJ_RTOL = 1e-6


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def node_rel(a, b, floor=1e-9):
    """Worst PER-NODE relative error: every (problem, node) block is normalised by its own largest entry (not by the
    largest entry of the whole batch, behind which one badly scaled node could hide)."""
    a, b = np.asarray(a), np.asarray(b)
    a2, b2 = a.reshape(a.shape[0] * a.shape[1], -1), b.reshape(b.shape[0] * b.shape[1], -1)
    return float((np.abs(a2 - b2).max(axis=1) / np.maximum(np.abs(b2).max(axis=1), floor)).max())


@pytest.fixture(scope="module")
def solver_mod():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from agimus_controller_b200 import solver

    return solver


def _workload(orc, B, T, **kw):
    m = panda_table().to_struct()
    return goal_reaching_batch(B, T=T, rnea=lambda q, v, a: orc.rnea(m, q, v, a), **kw), m


def _problem(solver_mod, w, B):
    p = solver_mod.BatchedShootingProblem(w["table"], w["dts"], B)
    p.set_refs(w["refs"])
    return p


def test_rnea_and_integrate(solver_mod, orc):
    w, m = _workload(orc, 4, 5)
    p = _problem(solver_mod, w, 4)
    rng = np.random.default_rng(0)
    q, v, a = rng.uniform(-2, 2, (3, 1000, 7))
    tau = p.rnea(q, v, a).cpu().numpy()
    assert rel(tau, orc.rnea(m, q, v, a)) < 1e-12
    x = np.concatenate([q, v], 1)
    xn = p.integrate(x, a, 0.01).cpu().numpy()
    assert rel(xn, orc.integrate(m, x, a, 0.01)) < 1e-12


@pytest.mark.parametrize("target_R", ["tool_down", "identity"])
def test_calc_and_calc_diff_per_node(solver_mod, orc, target_R):
    """4096 x 51 random nodes would need 1.1 GB of dense outputs per side; 512 x 51 nodes are checked."""
    B, T = 512, 50
    kw = {} if target_R == "tool_down" else dict(target_R=np.eye(3))
    w, m = _workload(orc, B, T, **kw)
    rng = np.random.default_rng(7)
    xs = w["xs_ws"] + rng.uniform(-0.3, 0.3, w["xs_ws"].shape)
    us = w["us_ws"] + rng.uniform(-5, 5, w["us_ws"].shape)
    p = _problem(solver_mod, w, B)
    cost, xnext = p.calc(xs, us)
    o = orc.calc_diff(m, w["refs"], w["dts"], xs, us)
    assert rel(cost.cpu().numpy(), o["cost"]) < DERIV_RTOL
    assert rel(xnext.cpu().numpy(), o["xnext"]) < DERIV_RTOL
    g = p.calc_diff(xs, us)
    for k in ("cost", "xnext", "Fx", "Fu", "Lx", "Lu", "Lxx", "Luu"):
        assert rel(g[k].cpu().numpy(), o[k]) < DERIV_RTOL, k
    # the 1e-9 gate node by node
    for k in ("xnext", "Fx", "Fu", "Lx", "Lxx"):
        assert node_rel(g[k].cpu().numpy(), o[k]) < DERIV_RTOL, k
    assert float(g["Lxu"].abs().max()) == 0.0


def test_rollout(solver_mod, orc):
    B, T = 64, 50
    w, m = _workload(orc, B, T)
    p = _problem(solver_mod, w, B)
    xs = p.rollout(w["x0"], w["us_ws"]).cpu().numpy()
    assert rel(xs, orc.rollout(m, w["refs"], w["dts"], w["x0"], w["us_ws"])) < 1e-9


@pytest.mark.parametrize("fixed,iters", [(True, 10), (False, 100)])
def test_solve_matches_oracle(solver_mod, orc, fixed, iters):
    """Config 2 shapes (T = 50, dt = 0.01), 256 problems: same iterates as the CPU restatement."""
    B, T = 256, 50
    w, m = _workload(orc, B, T)
    opts = _abi.default_fddp_opts(fixed_iters=fixed)
    o = orc.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], iters, opts)
    p = _problem(solver_mod, w, B)
    g = {k: v.cpu().numpy() for k, v in p.solve(w["x0"], w["xs_ws"], w["us_ws"], iters, opts).items()}
    np.testing.assert_array_equal(g["iters"], o["iters"])
    np.testing.assert_array_equal(g["status"], o["status"])
    for k in ("xs", "us", "cost"):
        assert rel(g[k], o[k]) < TRAJ_RTOL, k
    assert rel(g["K"], o["K"]) < 1e-5
    np.testing.assert_allclose(g["xs"][:, 0], w["x0"], rtol=0, atol=1e-12)


def _near_pi_nodes(table, xs, Rref=np.eye(3), margin=2e-2):
    """[B] bool: some node of the trajectory has its task frame within `margin` rad of a pi rotation from Rref —
    the band where log3 switches to its near-pi formula (theta >= pi - 1e-2) and sqrt((R_ii - cos)/(1 - cos)) of
    near-zero arguments amplifies a rounding difference of 1e-16 to 1e-8."""
    hit = np.zeros(xs.shape[0], dtype=bool)
    for b in range(xs.shape[0]):
        for t in range(xs.shape[1]):
            R, _ = table.frame_placement(xs[b, t, :7])
            c = 0.5 * (np.trace(Rref.T @ R) - 1.0)
            if np.arccos(np.clip(c, -1.0, 1.0)) >= np.pi - margin:
                hit[b] = True
                break
    return hit


def test_solve_identity_target_near_pi(solver_mod, orc):
    """Adversarial variant: Rref = I puts the log map on its theta = pi cut at the nominal posture.  The two
    implementations are compared iteration by iteration until their first different decision: every problem that
    stays clear of the near-pi band agrees to 1e-6 with identical decisions through all 10 iterations, and every
    problem that diverges does so at an iterate that has a node inside the band (where the branch taken is decided by
    a comparison of two numbers that agree to rounding)."""
    B, T, N = 64, 50, 10
    w, m = _workload(orc, B, T, target_R=np.eye(3))
    opts = _abi.default_fddp_opts(fixed_iters=True)
    p = _problem(solver_mod, w, B)
    first_diff = np.full(B, N + 1)
    prev_o = None
    last_common = {}
    for k in range(1, N + 1):
        o = orc.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], k, opts)
        g = {kk: v.cpu().numpy() for kk, v in p.solve(w["x0"], w["xs_ws"], w["us_ws"], k, opts).items()}
        scale = np.abs(o["xs"]).reshape(B, -1).max(1)
        differs = (np.abs(g["xs"] - o["xs"]).reshape(B, -1).max(1) > TRAJ_RTOL * scale) | (g["iters"] != o["iters"])
        for b in np.nonzero(differs & (first_diff > N))[0]:
            first_diff[b] = k
            last_common[b] = w["xs_ws"][b] if prev_o is None else prev_o["xs"][b]
        prev_o = o
    same = first_diff > N
    print("identity-target: problems with identical decisions through 10 iterations:", same.mean())
    assert same.mean() > 0.75, same.mean()
    assert rel(g["xs"][same], o["xs"][same]) < 1e-5
    # problems that never come near the cut must be among the agreeing ones
    clear = ~_near_pi_nodes(w["table"], o["xs"]) & ~_near_pi_nodes(w["table"], w["xs_ws"])
    assert same[clear].all(), "a problem away from the theta = pi band diverged"
    # and each diverging problem was inside the band at the last iterate the two implementations shared
    if last_common:
        lc = np.stack([last_common[b] for b in sorted(last_common)])
        assert _near_pi_nodes(w["table"], lc, margin=5e-2).all()
    # the diverged ones are still finite FDDP iterates
    assert np.isfinite(g["cost"]).all() and np.isfinite(g["xs"]).all() and np.isfinite(g["us"]).all()


def test_acceptance_rule_versions_give_the_same_benchmark_iterates(solver_mod, orc):
    """agx_fddp_opts.accept_rule: Crocoddyl >= 2.0 (|d1|, no negative-expectation step from a feasible candidate) vs
    1.x.  The two rules differ only for d1 < 0 or dVexp < 0 on a feasible candidate; the benchmark batch (cfg 2) never
    gets there: bit-identical iterates and decisions under both, on the GPU and on the oracle."""
    B, T = 512, 50
    w, m = _workload(orc, B, T)
    p = _problem(solver_mod, w, B)
    res = {}
    for rule in (0, 1):
        for fixed, iters in ((True, 10), (False, 60)):
            opts = _abi.default_fddp_opts(fixed_iters=fixed)
            opts.accept_rule = rule
            res[(rule, fixed)] = {k: v.clone() for k, v in p.solve(w["x0"], w["xs_ws"], w["us_ws"], iters, opts).items()}
    for fixed in (True, False):
        for k in ("xs", "us", "cost", "iters", "status"):
            assert torch.equal(res[(0, fixed)][k], res[(1, fixed)][k]), (fixed, k)
    o0 = orc.solve(m, w["refs"][:64], w["dts"], w["x0"][:64], w["xs_ws"][:64], w["us_ws"][:64], 60, _abi.default_fddp_opts())
    o1o = _abi.default_fddp_opts()
    o1o.accept_rule = 1
    o1 = orc.solve(m, w["refs"][:64], w["dts"], w["x0"][:64], w["xs_ws"][:64], w["us_ws"][:64], 60, o1o)
    assert np.array_equal(o0["xs"], o1["xs"]) and np.array_equal(o0["iters"], o1["iters"])


def test_solve_golden_problem(solver_mod, orc):
    """The reference's own golden OCP (tests/test_ocp_croco_base.py): T = 9, dt = 1e-3, zero warm start."""
    w = golden_problem()
    m = w["table"].to_struct()
    opts = _abi.default_fddp_opts(fixed_iters=True)
    o = orc.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 10, opts)
    p = _problem(solver_mod, w, 1)
    g = {k: v.cpu().numpy() for k, v in p.solve(w["x0"], w["xs_ws"], w["us_ws"], 10, opts).items()}
    np.testing.assert_array_equal(g["iters"], o["iters"])
    assert rel(g["cost"], o["cost"]) < TRAJ_RTOL
    assert rel(g["xs"], o["xs"]) < 1e-5
    assert rel(g["us"], o["us"]) < 1e-5


def test_per_problem_models(solver_mod, orc):
    """Config 5 shape: one inertial table per problem (evaluate_model_sensibility.py perturbations)."""
    B, T = 70, 20
    w, m0 = _workload(orc, B, T)
    base = w["table"]
    tables = [base.perturbed(b // 10, b % 10, 0.01) for b in range(B)]
    structs = [t.to_struct() for t in tables]
    opts = _abi.default_fddp_opts(fixed_iters=True)
    o = orc.solve(structs, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 5, opts)
    p = solver_mod.BatchedShootingProblem(tables, w["dts"], B)
    p.set_refs(w["refs"])
    g = {k: v.cpu().numpy() for k, v in p.solve(w["x0"], w["xs_ws"], w["us_ws"], 5, opts).items()}
    assert rel(g["xs"], o["xs"]) < TRAJ_RTOL
    assert rel(g["cost"], o["cost"]) < TRAJ_RTOL


def test_full_size_properties(solver_mod, orc):
    """BASELINE config 2 at full size (B = 4096, T = 50, 10 fixed iterations): size-independent checks.

    * xs[:, 0] == x0; feasible results satisfy xs == rollout(x0, us) (dynamics consistency);
    * the returned cost equals the sum of node costs re-evaluated by problem.calc;
    * the cost never increases w.r.t. the warm start; a 256-problem slab solved alone is bitwise
      identical to the same slab inside the full batch (sharding invariance);
    * all 4096 problems match the oracle (states, controls, cost, gains, iteration counts, statuses).
    """
    B, T = 4096, 50
    w, m = _workload(orc, B, T)
    opts = _abi.default_fddp_opts(fixed_iters=True)
    p = _problem(solver_mod, w, B)
    g = p.solve(w["x0"], w["xs_ws"], w["us_ws"], 10, opts)
    xs, us = g["xs"], g["us"]
    assert bool(torch.isfinite(xs).all()) and bool(torch.isfinite(us).all())
    assert float((xs[:, 0] - torch.as_tensor(w["x0"], device=xs.device)).abs().max()) < 1e-12
    xr = p.rollout(w["x0"], us)
    assert float((xr - xs).abs().max()) < 1e-7
    cost_nodes, _ = p.calc(xs, us)
    assert rel(cost_nodes.sum(1).cpu().numpy(), g["cost"].cpu().numpy()) < 1e-10
    c0, _ = p.calc(w["xs_ws"], w["us_ws"])
    assert bool((g["cost"] <= c0.sum(1) * (1 + 1e-12)).all())
    # sharding invariance
    sl = slice(1024, 1280)
    ps = solver_mod.BatchedShootingProblem(w["table"], w["dts"], 256)
    ps.set_refs(w["refs"][sl])
    gs = ps.solve(w["x0"][sl], w["xs_ws"][sl], w["us_ws"][sl], 10, opts)
    assert torch.equal(gs["xs"], xs[sl]) and torch.equal(gs["us"], us[sl]) and torch.equal(gs["K"], g["K"][sl])
    # every one of the 4096 problems against the oracle (a few seconds of CPU time on the box's host cores)
    o = orc.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 10, opts)
    np.testing.assert_array_equal(g["iters"].cpu().numpy(), o["iters"])
    np.testing.assert_array_equal(g["status"].cpu().numpy(), o["status"])
    per_problem = np.abs(xs.cpu().numpy() - o["xs"]).reshape(B, -1).max(1) / np.abs(o["xs"]).reshape(B, -1).max(1)
    assert per_problem.max() < TRAJ_RTOL, (per_problem.max(), int(per_problem.argmax()))
    assert rel(us.cpu().numpy(), o["us"]) < TRAJ_RTOL
    assert np.abs(g["cost"].cpu().numpy() - o["cost"]).max() / np.abs(o["cost"]).max() < TRAJ_RTOL
    assert rel(g["K"].cpu().numpy(), o["K"]) < 1e-5


def test_reference_golden_file_on_the_gpu(solver_mod, orc, golden):
    """The reference's own golden data (agimus_controller/tests/resources/simple_ocp_croco_results.pkl, compared by
    tests/test_ocp_croco_base.py:175-204) against the sm_100a kernels, no oracle in between:
    KAT-1 golden states follow from golden controls through agx_integrate (symplectic Euler, armature 0.1, dt 1e-3);
    KAT-3 golden Riccati gains = agx_riccati at the golden point with CSQP's diagonal terms (proximal sigma = 1e-6 plus
    the regularisation floor 1e-9), to < 1e-9 relative."""
    w = golden_problem()
    p = _problem(solver_mod, w, 1)
    xs, us, Kg = golden["states"], golden["feed_forward_terms"], golden["ricatti_gains"]
    xn = p.integrate(xs[:9], us, 1e-3).cpu().numpy()
    assert np.abs(xn - xs[1:]).max() < 2e-9
    K, k, status = p.riccati(w["x0"], xs[None], us[None], 1e-6 + 1e-9)
    K = K.cpu().numpy()[0]
    assert int(status[0]) != _abi.AGX_STATUS_REGMAX
    for t in range(9):
        assert np.abs(K[t] - Kg[t]).max() / np.abs(Kg[t]).max() < 1e-9, t
    m = w["table"].to_struct()
    Ko, ko, _ = orc.riccati_sigma(m, w["refs"][0], w["dts"], w["x0"][0], xs, us, 1e-6 + 1e-9)
    assert rel(K, Ko) < 1e-9
    K0, _, _ = p.riccati(w["x0"], xs[None], us[None], 0.0)
    assert np.abs(K0.cpu().numpy()[0, 0] - Kg[0]).max() / np.abs(Kg[0]).max() > 0.5


def test_cfg3_cartesian_sine_tracking(solver_mod, orc):
    """BASELINE config 3: Cartesian sine end-effector tracking (frame-placement residuals, one phase per problem).
    128 problems against the oracle; the full 16384-problem batch through size-independent properties."""
    from agimus_controller_b200.workloads import cartesian_sine_batch

    m = panda_table().to_struct()
    rn = lambda q, v, a: orc.rnea(m, q, v, a)  # noqa: E731
    opts = _abi.default_fddp_opts(fixed_iters=True)
    B, T = 128, 50
    w = cartesian_sine_batch(B, T=T, rnea=rn)
    o = orc.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 10, opts)
    p = _problem(solver_mod, w, B)
    g = {k: v.cpu().numpy() for k, v in p.solve(w["x0"], w["xs_ws"], w["us_ws"], 10, opts).items()}
    np.testing.assert_array_equal(g["iters"], o["iters"])
    for k in ("xs", "us", "cost"):
        assert rel(g[k], o[k]) < TRAJ_RTOL, k
    # the end effector moves towards the moving target along the horizon
    terms = p.cost_terms(g["xs"], g["us"])
    err = terms["r_pose"][..., :3].norm(dim=-1)
    assert float(err[:, -1].mean()) < 0.5 * float(err[:, 0].mean())
    # full size
    Bf = 16384
    wf = cartesian_sine_batch(Bf, T=T, rnea=rn)
    pf = _problem(solver_mod, wf, Bf)
    gf = pf.solve(wf["x0"], wf["xs_ws"], wf["us_ws"], 10, opts)
    assert bool(torch.isfinite(gf["xs"]).all())
    assert float((pf.rollout(wf["x0"], gf["us"]) - gf["xs"]).abs().max()) < 1e-7
    cost_nodes, _ = pf.calc(gf["xs"], gf["us"])
    assert rel(cost_nodes.sum(1).cpu().numpy(), gf["cost"].cpu().numpy()) < 1e-10
    # problems b and b + B/128 * k ... the 128-problem batch is the stride-128 subsample of the phases
    idx = np.arange(0, Bf, Bf // B)
    assert rel(gf["xs"].cpu().numpy()[idx], o["xs"]) < TRAJ_RTOL


def test_cfg5_model_sensibility_ensemble(solver_mod, orc):
    """BASELINE config 5 shape: every problem has its own inertial table (one parameter perturbed by delta * s)."""
    from agimus_controller_b200.workloads import model_sensibility_batch

    m = panda_table().to_struct()
    B, T = 280, 50
    w = model_sensibility_batch(B, T=T, rnea=lambda q, v, a: orc.rnea(m, q, v, a))
    structs = [t.to_struct() for t in w["tables"]]
    opts = _abi.default_fddp_opts(fixed_iters=True)
    o = orc.solve(structs, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 10, opts)
    p = solver_mod.BatchedShootingProblem(w["tables"], w["dts"], B)
    p.set_refs(w["refs"])
    g = {k: v.cpu().numpy() for k, v in p.solve(w["x0"], w["xs_ws"], w["us_ws"], 10, opts).items()}
    for k in ("xs", "us", "cost"):
        assert rel(g[k], o[k]) < TRAJ_RTOL, k
    # the perturbation matters: the same problems with the nominal table give different trajectories
    p0 = _problem(solver_mod, w, B)
    g0 = p0.solve(w["x0"], w["xs_ws"], w["us_ws"], 10, opts)
    assert float((g0["us"] - torch.as_tensor(g["us"], device=g0["us"].device)).abs().max()) > 1e-6


def test_cost_terms_and_shift_on_gpu(solver_mod, orc):
    B, T = 32, 12
    w, m = _workload(orc, B, T)
    dts = np.array([0.01] * 6 + [0.02] * 4 + [0.04] * 2)
    p = solver_mod.BatchedShootingProblem(w["table"], dts, B)
    p.set_refs(w["refs"])
    rng = np.random.default_rng(4)
    us = w["us_ws"] + rng.uniform(-2, 2, w["us_ws"].shape)
    xs = p.rollout(w["x0"], us)
    terms = p.cost_terms(xs, us)
    cost, _ = p.calc(xs, us)
    scale = torch.as_tensor(np.concatenate([dts, [1.0]]), device=cost.device)
    tot = (terms["state_reg"] + terms["control_reg"] + terms["goal_tracking"]) * scale
    assert rel(tot.cpu().numpy(), cost.cpu().numpy()) < 1e-12
    oxs, ous = p.shift_warmstart(xs, us)
    xs_h, oxs_h = xs.cpu().numpy(), oxs.cpu().numpy()
    np.testing.assert_array_equal(oxs_h[:, :6], xs_h[:, 1:7])
    exp = orc.integrate(m, xs_h[:, 6:12].reshape(-1, 14), us[:, 6:12].reshape(-1, 7), 0.01).reshape(B, 6, 14)
    np.testing.assert_allclose(oxs_h[:, 6:12], exp, rtol=0, atol=1e-11)
    np.testing.assert_array_equal(oxs_h[:, 12], xs_h[:, 12])
    np.testing.assert_array_equal(ous.cpu().numpy()[:, :6], us[:, 1:7])
    np.testing.assert_array_equal(ous.cpu().numpy()[:, 6:], us[:, 6:])


def test_reference_window_on_gpu(solver_mod, orc):
    """agx_set_refs_window gathers the same table the host would build from TrajectoryBuffer.horizon."""
    from agimus_controller_b200.workloads import sine_configuration_reference

    m = panda_table().to_struct()
    table, rows, q, v, u = sine_configuration_reference(200, rnea=lambda q_, v_, a_: orc.rnea(m, q_, v_, a_))
    dts = np.array([0.01] * 6 + [0.02] * 4 + [0.04] * 2)
    hidx = np.concatenate([[0], np.cumsum(np.round(dts / dts[0]).astype(int))])
    B, T = 64, 12
    p = solver_mod.BatchedShootingProblem(table, dts, B)
    rng = np.random.default_rng(1)
    xs = rng.uniform(-0.3, 0.3, (B, T + 1, 14)) + np.concatenate([q[0], v[0]])
    us = rng.uniform(-2, 2, (B, T, 7))
    start = rng.integers(0, 190, B)
    p.set_refs_window(torch.as_tensor(rows, device=p.device), torch.as_tensor(start, dtype=torch.int32))
    c_win, _ = p.calc(xs, us)
    refs = np.stack([rows[np.minimum(s0 + hidx, 199)] for s0 in start])
    p.set_refs(refs)
    c_tab, _ = p.calc(xs, us)
    assert torch.equal(c_win, c_tab)
    assert rel(c_win.cpu().numpy(), orc.calc(m, refs, dts, xs, us)[0]) < 1e-12


def test_converged_fddp_on_the_gpu_lands_on_the_golden_solution(solver_mod, golden):
    """KAT-8 on the GPU: the FDDP solve of the reference's golden OCP (tests/test_ocp_croco_base.py), zero warm start,
    run to convergence, lands on the reference's golden states / controls (3e-3 / 0.15: the golden file is an
    unconverged iterate of the reference's CSQP) at cost 202.6215."""
    w = golden_problem()
    p = _problem(solver_mod, w, 1)
    g = {k: v.cpu().numpy() for k, v in p.solve(w["x0"], w["xs_ws"], w["us_ws"], 100, _abi.default_fddp_opts()).items()}
    assert g["status"][0] == _abi.AGX_STATUS_CONVERGED and g["iters"][0] < 100
    assert np.abs(g["xs"][0] - golden["states"]).max() < 3e-3
    assert np.abs(g["us"][0] - golden["feed_forward_terms"]).max() < 0.15
    assert abs(g["cost"][0] - 202.6215) < 1e-3


def test_cfg4_collision_avoidance(solver_mod, orc):
    """BASELINE config 4 shape (T = 100, pick-and-place joint move, capsule-pair distance residuals under QuadExp;
    fingers locked): per-node derivatives and the solve against the CPU restatement, then the full 4096-problem batch
    through size-independent properties.  A10 is "parity unpinned": the restatement is the only oracle."""
    from agimus_controller_b200.workloads import pick_and_place_collision_batch

    m0 = panda_table().to_struct()
    rn = lambda q, v, a: orc.rnea(m0, q, v, a)  # noqa: E731
    B, T = 96, 100
    w = pick_and_place_collision_batch(B, T=T, rnea=rn, alpha=1e-3, w_col=(20.0, 20.0))
    m = w["table"].to_struct()
    p = _problem(solver_mod, w, B)
    rng = np.random.default_rng(5)
    xs = w["xs_ws"] + rng.uniform(-0.1, 0.1, w["xs_ws"].shape)
    us = w["us_ws"] + rng.uniform(-2, 2, w["us_ws"].shape)
    o = orc.calc_diff(m, w["refs"], w["dts"], xs, us)
    g = p.calc_diff(xs, us)
    for k in ("cost", "xnext", "Fx", "Fu", "Lx", "Lu", "Lxx", "Luu"):
        assert rel(g[k].cpu().numpy(), o[k]) < DERIV_RTOL, k
    cost, _ = p.calc(xs, us)
    assert rel(cost.cpu().numpy(), o["cost"]) < DERIV_RTOL
    # per-cost view: distances and activations of both pairs
    terms = p.cost_terms(xs, us)
    for b, t in ((0, 0), (5, 40), (95, 100)):
        for k in range(2):
            d, _, act = orc.collision(m, xs[b, t, :7], k)
            assert abs(float(terms["collision_distance"][b, t, k]) - d) < 1e-12
            assert abs(float(terms["collision"][b, t, k]) - w["refs"][b, t, 60 + k] * act[0]) < 1e-12
    for fixed, iters in ((True, 3), (False, 60)):
        opts = _abi.default_fddp_opts(fixed_iters=fixed)
        so = orc.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], iters, opts)
        sg = {k: v.cpu().numpy() for k, v in p.solve(w["x0"], w["xs_ws"], w["us_ws"], iters, opts).items()}
        np.testing.assert_array_equal(sg["iters"], so["iters"])
        np.testing.assert_array_equal(sg["status"], so["status"])
        for k in ("xs", "us", "cost"):
            assert rel(sg[k], so[k]) < TRAJ_RTOL, k
    # the collision cost moves the arm away from the obstacle: larger clearance than the plan without it
    w0 = dict(w, refs=w["refs"].copy())
    w0["refs"][..., 60:62] = 0.0
    p0 = _problem(solver_mod, w0, B)
    opts = _abi.default_fddp_opts()
    s1 = p.solve(w["x0"], w["xs_ws"], w["us_ws"], 60, opts)
    s0 = p0.solve(w["x0"], w["xs_ws"], w["us_ws"], 60, opts)
    d1 = p.cost_terms(s1["xs"], s1["us"])["collision_distance"][..., 1]
    d0 = p.cost_terms(s0["xs"], s0["us"])["collision_distance"][..., 1]
    assert float(d1[:, 5:30].min(1).values.mean()) > float(d0[:, 5:30].min(1).values.mean()) + 1e-3
    # full size
    Bf = 4096
    wf = pick_and_place_collision_batch(Bf, T=T, rnea=rn, alpha=1e-3, w_col=(20.0, 20.0))
    pf = _problem(solver_mod, wf, Bf)
    gf = pf.solve(wf["x0"], wf["xs_ws"], wf["us_ws"], 3, _abi.default_fddp_opts(fixed_iters=True))
    assert bool(torch.isfinite(gf["xs"]).all())
    assert float((pf.rollout(wf["x0"], gf["us"]) - gf["xs"]).abs().max()) < 1e-7
    cost_nodes, _ = pf.calc(gf["xs"], gf["us"])
    assert rel(cost_nodes.sum(1).cpu().numpy(), gf["cost"].cpu().numpy()) < 1e-10
    np.testing.assert_allclose(gf["xs"].cpu().numpy()[:B], pf.solve(wf["x0"], wf["xs_ws"], wf["us_ws"], 3,
                               _abi.default_fddp_opts(fixed_iters=True))["xs"].cpu().numpy()[:B], rtol=0, atol=0)


def test_sqp_mode_matches_oracle_and_the_golden_file(solver_mod, orc, golden):
    """SQP mode (mim_solvers.SolverCSQP without active constraints, what the reference runs): cfg-2 shapes against the
    CPU restatement with identical per-problem decisions, and the reference's golden file at its own 6 decimals."""
    B, T = 256, 50
    w, m = _workload(orc, B, T)
    p = _problem(solver_mod, w, B)
    for max_iter in (3, 60):
        o = orc.solve_sqp(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], max_iter)
        g = {k: v.cpu().numpy() for k, v in p.solve_sqp(w["x0"], w["xs_ws"], w["us_ws"], max_iter).items()}
        np.testing.assert_array_equal(g["iters"], o["iters"])
        np.testing.assert_array_equal(g["status"], o["status"])
        for k in ("xs", "us", "cost", "stop"):
            assert rel(g[k], o[k]) < TRAJ_RTOL, k
        assert rel(g["K"], o["K"]) < 1e-5
    assert (g["status"] == _abi.AGX_STATUS_CONVERGED).all() and (g["stop"] <= 1e-3).all()
    wg = golden_problem()
    pg = _problem(solver_mod, wg, 1)
    s = {k: v.cpu().numpy() for k, v in pg.solve_sqp(wg["x0"], wg["xs_ws"], wg["us_ws"], 100).items()}
    assert int(s["iters"][0]) == 33 and int(s["status"][0]) == _abi.AGX_STATUS_CONVERGED
    np.testing.assert_array_almost_equal(s["xs"][0], golden["states"], decimal=6)
    np.testing.assert_array_almost_equal(s["K"][0], golden["ricatti_gains"], decimal=6)
    np.testing.assert_array_almost_equal(s["us"][0], golden["feed_forward_terms"], decimal=6)
