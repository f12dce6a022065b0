"""OCPBatchedFDDP behind an MPC.run-shaped driver (the contract agimus_controller/tests/test_mpc_unicycle.py
pins with a fake OCP: any OCPBase works under MPC; res.states[0] == x0 and res.states[1] == integrate(x0, u0))."""
import pathlib
import time

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from agimus_controller_b200 import PANDA_Q_NOMINAL, panda_table  # noqa: E402
from agimus_controller_b200.ocp_interface import (DTFactorsNSeq, OCPBase, OCPParamsBaseCroco, SE3,  # noqa: E402
                                                  TrajectoryPoint, TrajectoryPointWeights, WeightedTrajectoryPoint)

YAML = pathlib.Path(__file__).parent / "golden" / "ocp_goal_reaching.yaml"


def _wpoint(i, q, nv=7):
    return WeightedTrajectoryPoint(
        point=TrajectoryPoint(id=i, time_ns=i * 10_000_000, robot_configuration=q, robot_velocity=np.zeros(nv),
                              robot_acceleration=np.zeros(nv), robot_effort=np.zeros(nv),
                              end_effector_poses={"panda_hand_tcp": SE3(np.diag([1.0, -1.0, -1.0]),
                                                                        np.array([0.5, 0.2, 0.5]))}),
        weights=TrajectoryPointWeights(w_robot_configuration=np.full(nv, 1.0), w_robot_velocity=np.full(nv, 0.1),
                                       w_robot_acceleration=np.zeros(nv), w_robot_effort=np.full(nv, 1e-3),
                                       w_end_effector_poses={"panda_hand_tcp": np.full(6, 0.1)}))


class ShiftWarmStart:
    """Minimal WarmStartBase stand-in: previous solution shifted by one node (warm_start_shift_previous_solution.py)."""

    def __init__(self):
        self.prev = None

    def generate(self, x0, T, nv, u_grav):
        if self.prev is None:
            return x0, [x0] * (T + 1), [u_grav] * T
        xs, us = self.prev.states, self.prev.feed_forward_terms
        return x0, [x0] + list(xs[2:]) + [xs[-1]], list(us[1:]) + [us[-1]]

    def update_previous_solution(self, res):
        self.prev = res


@pytest.mark.parametrize("solver", ["fddp", "csqp"])
def test_ocp_plugs_into_an_mpc_loop(solver):
    """Both solver modes: FDDP (north_star) and the reference's own CSQP in its unconstrained form."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from agimus_controller_b200.ocp_batched import OCPBatchedFDDP

    T = 20
    params = OCPParamsBaseCroco(dt=0.01, solver_iters=10, dt_factor_n_seq=DTFactorsNSeq([1], [T]), horizon_size=T)
    ocp = OCPBatchedFDDP(panda_table(), params, str(YAML), batch_size=1, solver=solver)
    assert isinstance(ocp, OCPBase) and ocp.n_controls == T and ocp.dt == 0.01
    nv = 7
    # sine-wave configuration-space reference (trajectories/sine_wave_configuration_space.py): amplitude 0.2, period 4 s
    buffer = [_wpoint(i, PANDA_Q_NOMINAL + 0.2 * np.sin(2 * np.pi * i * 0.01 / 4.0) * np.ones(nv)) for i in range(T + 60)]
    ws = ShiftWarmStart()
    x = np.concatenate([PANDA_Q_NOMINAL, np.zeros(nv)])
    u_grav = ocp.problem.rnea(PANDA_Q_NOMINAL, np.zeros(nv), np.zeros(nv))[0].cpu().numpy()
    solve_ns = []
    for tick in range(30):
        horizon = buffer[: T + 1]
        ocp.set_reference_weighted_trajectory(horizon)
        x0, x_init, u_init = ws.generate(x, T, nv, u_grav)
        assert len(x_init) == ocp.n_controls + 1 and len(u_init) == ocp.n_controls
        t0 = time.perf_counter_ns()
        ocp.solve(x0, x_init, u_init)          # MPC.run passes exactly three positional arguments (mpc.py:53)
        solve_ns.append(time.perf_counter_ns() - t0)
        res = ocp.ocp_results
        ws.update_previous_solution(res)
        buffer.pop(0)
        assert len(res.states) == T + 1 and len(res.ricatti_gains) == T and len(res.feed_forward_terms) == T
        assert res.ricatti_gains[0].shape == (nv, 2 * nv)
        np.testing.assert_allclose(res.states[0], x0, atol=1e-12)
        nxt = ocp.integrate(x0, res.feed_forward_terms[0])
        # FDDP iterates are rollouts; SQP iterates are feasible up to the KKT tolerance (gaps <= 1e-3)
        np.testing.assert_allclose(res.states[1], nxt, atol=1e-8 if solver == "fddp" else 2e-3)
        assert ocp.debug_data.nb_iter >= (1 if solver == "fddp" else 0) and np.isfinite(ocp.debug_data.kkt_norm)
        x = nxt
    # the tracked joint positions follow the reference (weights 1.0 on q): error stays small
    assert np.abs(x[:nv] - buffer[0].point.robot_configuration).max() < 0.05
    # first solve without iteration limits (agimus_controller.py:376-381)
    ocp.solve(x0, x_init, u_init, use_iteration_limits_and_timeout=False)
    assert ocp.debug_data.problem_solved


def test_batched_overload_matches_single_problem():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from agimus_controller_b200.ocp_batched import OCPBatchedFDDP

    T, B, nv = 20, 8, 7
    params = OCPParamsBaseCroco(dt=0.01, solver_iters=5, dt_factor_n_seq=DTFactorsNSeq([1], [T]), horizon_size=T)
    horizon = [_wpoint(i, PANDA_Q_NOMINAL) for i in range(T + 1)]
    rng = np.random.default_rng(0)
    x0 = np.concatenate([PANDA_Q_NOMINAL + rng.uniform(-0.2, 0.2, (B, nv)), np.zeros((B, nv))], 1)
    ob = OCPBatchedFDDP(panda_table(), params, str(YAML), batch_size=B)
    ob.set_reference_weighted_trajectory(horizon)
    xs = np.repeat(x0[:, None], T + 1, 1)
    us = np.zeros((B, T, nv))
    ob.solve(torch.as_tensor(x0, device="cuda"), torch.as_tensor(xs, device="cuda"), torch.as_tensor(us, device="cuda"))
    rb = ob.ocp_results_batched
    o1 = OCPBatchedFDDP(panda_table(), params, str(YAML), batch_size=1)
    o1.set_reference_weighted_trajectory(horizon)
    for b in (0, 5):
        o1.solve(x0[b], list(xs[b]), list(us[b]))
        # the single-problem form runs in latency mode (eager_exit: two-warp forward pass), which agrees with the
        # throughput kernels to rounding, not bitwise
        np.testing.assert_allclose(np.stack(o1.ocp_results.states), rb["xs"][b].cpu().numpy(), rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(np.stack(o1.ocp_results.ricatti_gains), rb["K"][b].cpu().numpy(), rtol=1e-8, atol=1e-9)


def test_batched_closed_loop_with_device_warm_starts(orc):
    """A batch of closed-loop MPCs that never leaves the device between ticks: reference warm start on the first
    tick (batched RNEA), shifted previous solution afterwards; every tick's warm start equals the reference
    semantics evaluated on the host with the oracle."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from agimus_controller_b200 import _abi
    from agimus_controller_b200.solver import BatchedShootingProblem
    from agimus_controller_b200.warm_start import WarmStartReference, WarmStartShiftPreviousSolution
    from agimus_controller_b200.workloads import goal_reaching_batch

    B, T, nv = 16, 20, 7
    m = panda_table().to_struct()
    w = goal_reaching_batch(B, T=T, rnea=lambda q, v, a: orc.rnea(m, q, v, a))
    p = BatchedShootingProblem(w["table"], w["dts"], B)
    p.set_refs(w["refs"])
    ref_q = np.broadcast_to(PANDA_Q_NOMINAL, (B, T + 1, nv)).copy()
    zeros = np.zeros((B, T + 1, nv))
    x0, xs, us = WarmStartReference(p).generate_batched(w["x0"], ref_q, zeros, zeros)
    # u_init[t] = rnea(q_t, v_t, a_t) exactly as tests/test_warm_start_reference.py:19-75 pins it
    xs_h = xs.cpu().numpy()
    np.testing.assert_allclose(us.cpu().numpy().reshape(-1, nv),
                               orc.rnea(m, xs_h[:, :-1, :nv].reshape(-1, nv), xs_h[:, :-1, nv:].reshape(-1, nv),
                                        np.zeros((B * T, nv))), rtol=0, atol=1e-10)
    ws = WarmStartShiftPreviousSolution(p)
    opts = _abi.default_fddp_opts()
    x = x0
    cost_first = None
    for tick in range(8):
        out = p.solve(x, xs, us, 10, opts)
        if cost_first is None:
            cost_first = out["cost"].clone()
        ws.update_previous_solution({k: v.clone() for k, v in out.items()})
        x = p.integrate(x, out["us"][:, 0], 0.01)          # plant = the OCP's own integrator
        _, xs, us = ws.generate_batched(x)
        torch.testing.assert_close(xs[:, :-1], out["xs"][:, 1:], rtol=0, atol=0)   # constant dt: pure shift
        torch.testing.assert_close(xs[:, 0], x, rtol=0, atol=1e-9)                  # the plant follows the plan
    assert bool((out["cost"] < cost_first).all())


def test_pick_and_place_configuration_through_the_ocp_class(orc):
    """The reference's pick-and-place example configuration (panda_pick_and_place/config: dt 0.01, 60 nodes =
    30 x dt + 20 x 2dt + 10 x 4dt, max_iter 3, control_reg + state_reg only, terminal weight 0) through
    OCPBatchedFDDP, against the oracle fed with the same flattened tables."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import yaml

    from agimus_controller_b200 import _abi
    from agimus_controller_b200.ocp_batched import OCPBatchedFDDP, build_reference_rows, flatten_cost_stack
    from agimus_controller_b200.workloads import quintic

    nv = 7
    params = OCPParamsBaseCroco(dt=0.01, solver_iters=3, dt_factor_n_seq=DTFactorsNSeq([1, 2, 4], [30, 20, 10]),
                                horizon_size=60)
    assert abs(params.total_time - 1.1) < 1e-12
    yml = pathlib.Path(__file__).parent / "golden" / "ocp_pick_and_place.yaml"
    table = panda_table()
    ocp = OCPBatchedFDDP(table, params, str(yml), batch_size=1)
    # quintic joint-space move q_init -> q_nom over 1 s, sampled at the horizon indexes of the trajectory buffer
    q_init = PANDA_Q_NOMINAL + np.array([0.3, -0.2, 0.25, 0.3, -0.3, 0.2, 0.1])
    hidx = np.concatenate([[0], np.cumsum([1] * 30 + [2] * 20 + [4] * 10)])
    horizon = []
    for i in hidx:
        s = float(quintic(0.01 * i, 1.0))
        q = q_init + s * (PANDA_Q_NOMINAL - q_init)
        horizon.append(WeightedTrajectoryPoint(
            point=TrajectoryPoint(id=int(i), robot_configuration=q, robot_velocity=np.zeros(nv),
                                  robot_acceleration=np.zeros(nv), robot_effort=np.zeros(nv),
                                  end_effector_poses={"panda_hand_tcp": SE3()}),
            weights=TrajectoryPointWeights(w_robot_configuration=np.full(nv, 3.0), w_robot_velocity=np.full(nv, 0.12),
                                           w_robot_acceleration=np.zeros(nv), w_robot_effort=np.full(nv, 8e-4),
                                           w_end_effector_poses={"panda_hand_tcp": np.zeros(6)})))
    ocp.set_reference_weighted_trajectory(horizon)
    x0 = np.concatenate([q_init, np.zeros(nv)])
    u_grav = ocp.problem.rnea(q_init, np.zeros(nv), np.zeros(nv))[0].cpu().numpy()
    ocp.solve(x0, [x0] * 61, [u_grav] * 60)
    res = ocp.ocp_results
    # oracle on the same flattened problem
    data = yaml.safe_load(yml.read_text())
    rows = build_reference_rows(table, flatten_cost_stack(data["running_model"], False),
                                flatten_cost_stack(data["terminal_model"], True), horizon)
    assert np.all(rows[-1, 14:28] == 0.0)  # terminal CostModelSum weight 0
    m = table.to_struct()
    o = orc.solve(m, rows[None], np.asarray(params.timesteps), x0[None], np.repeat(x0[None, None], 61, 1),
                  np.repeat(u_grav[None, None], 60, 1), 3, _abi.default_fddp_opts())
    assert np.abs(np.stack(res.states) - o["xs"][0]).max() / np.abs(o["xs"]).max() < 1e-6
    assert np.abs(np.stack(res.feed_forward_terms) - o["us"][0]).max() / np.abs(o["us"]).max() < 1e-6
    assert np.abs(np.stack(res.ricatti_gains) - o["K"][0]).max() / np.abs(o["K"]).max() < 1e-5
    assert ocp.debug_data.nb_iter == int(o["iters"][0])


def test_check_results(golden):
    """The reference's golden test, line for line (agimus_controller/tests/test_ocp_croco_base.py:140-204): an OCP with
    the test's cost stack, nine nodes, Euler step 1e-3, solver_iters = 100, solved from `state_reg` with a zero warm
    start by the solver the reference instantiates (CSQP mode); `ocp_results.states / ricatti_gains /
    feed_forward_terms` must equal the pickle at 6 decimals, as `test_check_results` demands of the reference."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from agimus_controller_b200.ocp_batched import OCPBatchedFDDP

    nv = 7
    # the reference test builds IntegratedActionModelEuler without a step, i.e. Crocoddyl's default 1e-3 (:54-56)
    params = OCPParamsBaseCroco(dt=1e-3, horizon_size=9, dt_factor_n_seq=DTFactorsNSeq(factors=[1], n_steps=[9]),
                                solver_iters=100, callbacks=False)
    yml = pathlib.Path(__file__).parent / "golden" / "ocp_croco_base_test.yaml"
    ocp = OCPBatchedFDDP(panda_table(), params, str(yml), batch_size=1, solver="csqp")
    state_reg = np.concatenate((np.zeros(nv), np.zeros(nv)))      # pin.neutral, zero velocity
    point = WeightedTrajectoryPoint(
        point=TrajectoryPoint(id=0, robot_configuration=np.zeros(nv), robot_velocity=np.zeros(nv),
                              robot_acceleration=np.zeros(nv), robot_effort=np.zeros(nv),
                              end_effector_poses={"panda_hand_tcp": SE3(np.eye(3), np.array([1.0, 1.0, 1.0]))}),
        weights=TrajectoryPointWeights(w_robot_configuration=np.ones(nv), w_robot_velocity=np.ones(nv),
                                       w_robot_acceleration=np.zeros(nv), w_robot_effort=np.ones(nv),
                                       w_end_effector_poses={"panda_hand_tcp": np.ones(6)}))
    ocp.set_reference_weighted_trajectory([point] * (params.n_controls + 1))
    state_warmstart = [np.zeros(2 * nv)] * (params.n_controls + 1)
    control_warmstart = [np.zeros(nv)] * params.n_controls
    ocp.solve(state_reg, state_warmstart, control_warmstart)
    for it, state in enumerate(golden["states"]):
        np.testing.assert_array_almost_equal(state, ocp.ocp_results.states[it], err_msg="States are not equal")
    for it, gain in enumerate(golden["ricatti_gains"]):
        np.testing.assert_array_almost_equal(gain, ocp.ocp_results.ricatti_gains[it],
                                             err_msg="Ricatti gains are not equal")
    for it, term in enumerate(golden["feed_forward_terms"]):
        np.testing.assert_array_almost_equal(term, ocp.ocp_results.feed_forward_terms[it],
                                             err_msg="Feed forward term are not equal")
    assert ocp.debug_data.problem_solved and ocp.debug_data.nb_iter == 33 and ocp.debug_data.kkt_norm <= 1e-3


def test_collision_ocp_and_update_geometry_placement(orc):
    """The collision-avoidance cost stack (ResidualDistanceCollision + QuadExp, tests/golden/ocp_collision_avoidance.yaml)
    through the OCP class, and `update_geometry_placement` (ocp_base_croco.py:110-131; the controller calls it for every
    obstacle pose, agimus_controller.py:406): moving the obstacle changes the device tables exactly as rebuilding the
    model with the obstacle there would."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import yaml

    from agimus_controller_b200 import _abi
    from agimus_controller_b200.ocp_batched import (OCPBatchedFDDP, build_reference_rows, flatten_cost_stack,
                                                    resolve_collision_pairs)
    from agimus_controller_b200.robot_model import PANDA_CAPSULES

    T, nv = 12, 7
    yml = pathlib.Path(__file__).parent / "golden" / "ocp_collision_avoidance.yaml"
    params = OCPParamsBaseCroco(dt=0.01, solver_iters=4, dt_factor_n_seq=DTFactorsNSeq([1], [T]), horizon_size=T)
    table = panda_table().with_capsules(PANDA_CAPSULES, [])
    ocp = OCPBatchedFDDP(table, params, str(yml), batch_size=1)
    horizon = [_wpoint(i, PANDA_Q_NOMINAL) for i in range(T + 1)]
    for pt in horizon:
        pt.weights.w_collision_avoidance = 25.0
    ocp.set_reference_weighted_trajectory(horizon)
    x0 = np.concatenate([PANDA_Q_NOMINAL + 0.1, np.zeros(nv)])
    u_grav = ocp.problem.rnea(x0[:nv], np.zeros(nv), np.zeros(nv))[0].cpu().numpy()
    with pytest.raises(RuntimeError, match="Unknown geometry name 'no_such_obstacle'"):
        ocp.update_geometry_placement("no_such_obstacle", SE3(np.eye(3), np.zeros(3)))
    # the obstacle moves next to the hand: capsule axis = the placement's z axis, same length and radius
    Rz = np.array([[1.0, 0.0, 0.0], [0.0, 0.0, -1.0], [0.0, 1.0, 0.0]])     # z axis -> world y
    centre = np.array([0.40, 0.0, 0.45])
    ocp.update_geometry_placement("obstacle_capsule", SE3(Rz, centre))
    ocp.solve(x0, [x0] * (T + 1), [u_grav] * T)
    res = ocp.ocp_results
    # oracle on a model rebuilt with the obstacle at its new place
    data = yaml.safe_load(yml.read_text())
    run, term = flatten_cost_stack(data["running_model"], False), flatten_cost_stack(data["terminal_model"], True)
    caps = dict(PANDA_CAPSULES)
    half = 0.2
    caps["obstacle_capsule"] = (None, centre - half * Rz[:, 2], centre + half * Rz[:, 2], 0.05)
    moved = resolve_collision_pairs(panda_table().with_capsules(caps, []), run, term)
    rows = build_reference_rows(moved, run, term, horizon)
    o = orc.solve(moved.to_struct(), rows[None], np.asarray(params.timesteps), x0[None], np.repeat(x0[None, None], T + 1, 1),
                  np.repeat(u_grav[None, None], T, 1), 4, _abi.default_fddp_opts())
    assert np.abs(np.stack(res.states) - o["xs"][0]).max() / np.abs(o["xs"]).max() < 1e-6
    assert np.abs(np.stack(res.feed_forward_terms) - o["us"][0]).max() / np.abs(o["us"]).max() < 1e-6
    # and the obstacle matters: the distance read-out of the per-cost view is small enough for the cost to act
    terms = ocp.problem.cost_terms(np.stack(res.states)[None], np.stack(res.feed_forward_terms)[None])
    assert float(terms["collision_distance"][0, :, 0].min()) < 0.3 and float(terms["collision"][0, :, 0].max()) > 1e-6
