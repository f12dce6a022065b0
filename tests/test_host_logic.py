"""Host-side flattening: YAML cost stack -> cost slots, weighted trajectory points -> reference records.

Mirrors the identities the reference pins in agimus_controller/tests/test_ocp_croco_generic.py:30-113
(r == x - xref, cost == sum 1/2 w r^2, before and after an update) on the flattened tables, checked with the
CPU oracle so that no GPU is needed.
"""
import pathlib

import numpy as np
import pytest
import yaml

from agimus_controller_b200 import PANDA_Q_NOMINAL, _abi, panda_table
from agimus_controller_b200.ocp_batched import build_reference_rows, flatten_cost_stack, resolve_collision_pairs
from agimus_controller_b200.robot_model import PANDA_CAPSULES
from agimus_controller_b200.ocp_interface import (DTFactorsNSeq, OCPParamsBaseCroco, SE3, TrajectoryPoint,
                                                  TrajectoryPointWeights, WeightedTrajectoryPoint)

GOAL_REACHING = pathlib.Path(__file__).parent / "golden" / "ocp_goal_reaching.yaml"
COLLISION = pathlib.Path(__file__).parent / "golden" / "ocp_collision_avoidance.yaml"


def _point(nv, q, v, u, R, p, wq, wv, wu, wpose, name="panda_hand_tcp"):
    return WeightedTrajectoryPoint(
        point=TrajectoryPoint(robot_configuration=q, robot_velocity=v, robot_acceleration=np.zeros(nv),
                              robot_effort=u, end_effector_poses={name: SE3(R, p)}),
        weights=TrajectoryPointWeights(w_robot_configuration=np.full(nv, wq), w_robot_velocity=np.full(nv, wv),
                                       w_robot_acceleration=np.zeros(nv), w_robot_effort=np.full(nv, wu),
                                       w_end_effector_poses={name: np.full(6, wpose)}))


def test_params_timesteps_match_reference_semantics():
    p = OCPParamsBaseCroco(dt=0.01, solver_iters=3, dt_factor_n_seq=DTFactorsNSeq([1, 2, 4], [30, 20, 10]),
                           horizon_size=60)
    assert p.n_controls == 60
    assert p.timesteps[:30] == (0.01,) * 30 and p.timesteps[30:50] == (0.02,) * 20 and p.timesteps[50:] == (0.04,) * 10
    assert abs(p.total_time - 1.1) < 1e-12
    with pytest.raises(AssertionError):
        OCPParamsBaseCroco(dt=0.01, solver_iters=3, dt_factor_n_seq=DTFactorsNSeq([1], [5]), horizon_size=4)


def test_flatten_goal_reaching_yaml():
    data = yaml.safe_load(GOAL_REACHING.read_text())
    run = flatten_cost_stack(data["running_model"], terminal=False)
    term = flatten_cost_stack(data["terminal_model"], terminal=True)
    assert run["weights"] == {"state": 1.0, "control": 1.0, "pose": 1.0}
    assert term["weights"] == {"state": 1.0, "control": 0.0, "pose": 1.0}
    assert run["names"] == {"control": "control_reg", "state": "state_reg", "pose": "goal_tracking"}


def test_flatten_refuses_what_the_device_path_does_not_cover():
    data = yaml.safe_load(GOAL_REACHING.read_text())
    bad = yaml.safe_load(GOAL_REACHING.read_text())
    bad["running_model"]["differential"]["costs"][0]["cost"]["residual"]["class"] = "ResidualDistanceCollision"
    with pytest.raises(NotImplementedError):
        flatten_cost_stack(bad["running_model"], terminal=False)
    data["running_model"]["differential"]["constraints"] = [{"name": "c"}]
    with pytest.raises(NotImplementedError):
        flatten_cost_stack(data["running_model"], terminal=False)


def test_reference_rows_reproduce_residual_identities(orc):
    """cost == sum 1/2 w r^2 with r = x - xref, u - uref, log6(Mref^-1 oMf), through the flattened records."""
    table = panda_table()
    m = table.to_struct()
    nv, T = 7, 3
    data = yaml.safe_load(GOAL_REACHING.read_text())
    run = flatten_cost_stack(data["running_model"], terminal=False)
    term = flatten_cost_stack(data["terminal_model"], terminal=True)
    rng = np.random.default_rng(3)
    R = np.diag([1.0, -1.0, -1.0])
    horizon = [_point(nv, PANDA_Q_NOMINAL + 0.1 * rng.standard_normal(nv), 0.1 * rng.standard_normal(nv),
                      rng.standard_normal(nv), R, np.array([0.5, 0.2, 0.5]), 2.0, 0.5, 1e-2, 10.0) for _ in range(T + 1)]
    rows = build_reference_rows(table, run, term, horizon)
    assert rows.shape == (T + 1, _abi.ref_size(nv))
    xs = rng.uniform(-0.5, 0.5, (1, T + 1, 2 * nv)) + np.concatenate([PANDA_Q_NOMINAL, np.zeros(nv)])
    us = rng.uniform(-3, 3, (1, T, nv))
    dts = np.full(T, 0.01)
    cost, _ = orc.calc(m, rows[None], dts, xs, us)
    for t in range(T + 1):
        pt = horizon[t]
        rx = xs[0, t] - pt.point.robot_state
        c = 0.5 * np.sum(pt.weights.w_robot_state * rx ** 2)
        Rf, pf = orc.frame_placement(m, xs[0, t, :nv])
        r6 = orc.log6(R.T @ Rf, R.T @ (pf - np.array([0.5, 0.2, 0.5])))
        c += 0.5 * np.sum(10.0 * r6 ** 2)
        if t < T:
            ru = us[0, t] - pt.point.robot_effort
            c += 0.5 * np.sum(pt.weights.w_robot_effort * ru ** 2)
            c *= dts[t]
        assert abs(cost[0, t] - c) <= 1e-12 * max(1.0, abs(c)), t
    # an update (new references / weights) changes the records and nothing else
    horizon2 = [_point(nv, np.zeros(nv), np.zeros(nv), np.zeros(nv), R, np.zeros(3), 1.0, 1.0, 1.0, 0.0) for _ in range(T + 1)]
    rows2 = build_reference_rows(table, run, term, horizon2)
    cost2, _ = orc.calc(m, rows2[None], dts, xs, us)
    c0 = 0.5 * np.sum(xs[0, 0] ** 2) + 0.5 * np.sum(us[0, 0] ** 2)
    assert abs(cost2[0, 0] - dts[0] * c0) < 1e-12 * c0
    assert abs(cost2[0, T] - 0.5 * np.sum(xs[0, T] ** 2)) < 1e-12


def test_wrong_frame_is_refused():
    table = panda_table()
    data = yaml.safe_load(GOAL_REACHING.read_text())
    run = flatten_cost_stack(data["running_model"], terminal=False)
    term = flatten_cost_stack(data["terminal_model"], terminal=True)
    pt = _point(7, np.zeros(7), np.zeros(7), np.zeros(7), np.eye(3), np.zeros(3), 1, 1, 1, 1, name="other_frame")
    with pytest.raises(NotImplementedError):
        build_reference_rows(table, run, term, [pt, pt])


def test_collision_costs_are_flattened_into_pair_slots(orc):
    """ResidualDistanceCollision + QuadExp costs (ocp_croco_generic.py:119-147, :499-535): pairs are registered on the
    table, `update: true` takes the point's scalar w_collision_avoidance (:714-719), `update: false` keeps the
    CostModelSum weight of the YAML; the terminal stack of the fixture has no collision cost."""
    data = yaml.safe_load(COLLISION.read_text())
    run = flatten_cost_stack(data["running_model"], terminal=False)
    term = flatten_cost_stack(data["terminal_model"], terminal=True)
    assert [c["name"] for c in run["collisions"]] == ["avoid_collision_obstacle", "avoid_collision_self"]
    assert term["collisions"] == []
    table = resolve_collision_pairs(panda_table().with_capsules(PANDA_CAPSULES, []), run, term)
    assert table.collision_pairs == [("link7_capsule", "obstacle_capsule"), ("link7_capsule", "link3_capsule")]
    assert table.collision_alpha == 1e-2
    m = table.to_struct()
    nv, T = 7, 2
    R = np.diag([1.0, -1.0, -1.0])
    horizon = [_point(nv, PANDA_Q_NOMINAL, np.zeros(nv), np.zeros(nv), R, np.array([0.5, 0.2, 0.5]), 1.0, 1.0, 1.0, 1.0)
               for _ in range(T + 1)]
    for k, pt in enumerate(horizon):
        pt.weights.w_collision_avoidance = 4.0 + k
    rows = build_reference_rows(table, run, term, horizon)
    np.testing.assert_array_equal(rows[:, 60:62], [[4.0, 2.5], [5.0, 2.5], [0.0, 0.0]])
    # cost identity through the records: w a(r) on top of the weighted-quad terms
    q = PANDA_Q_NOMINAL + 0.3
    xs = np.tile(np.concatenate([q, np.zeros(nv)]), (1, T + 1, 1))
    us = np.zeros((1, T, nv))
    dts = np.full(T, 0.01)
    with_col, _ = orc.calc(m, rows[None], dts, xs, us)
    rows0 = rows.copy()
    rows0[:, 60:62] = 0.0
    without, _ = orc.calc(m, rows0[None], dts, xs, us)
    a = [np.exp(-orc.collision(m, q, k)[0] ** 2 / 1e-2) for k in range(2)]
    assert abs((with_col - without)[0, 0] - dts[0] * (4.0 * a[0] + 2.5 * a[1])) < 1e-14
    assert (with_col - without)[0, T] == 0.0


def test_collision_flattening_errors():
    data = yaml.safe_load(COLLISION.read_text())
    run = flatten_cost_stack(data["running_model"], terminal=False)
    term = flatten_cost_stack(data["terminal_model"], terminal=True)
    with pytest.raises(ValueError, match="Geometry object 'link7_capsule' not found"):
        resolve_collision_pairs(panda_table(), run, term)  # a table without capsules
    bad = yaml.safe_load(COLLISION.read_text())
    bad["running_model"]["differential"]["costs"][3]["cost"]["activation"] = {"class": "ActivationModelExp", "alpha": 1.0}
    with pytest.raises(NotImplementedError, match="ActivationModelExp"):
        flatten_cost_stack(bad["running_model"], terminal=False)  # exponent 1 is not on the device path
    bad = yaml.safe_load(COLLISION.read_text())
    bad["running_model"]["differential"]["costs"][4]["cost"]["activation"]["alpha"] = 0.5
    with pytest.raises(NotImplementedError, match="different activation alphas"):
        resolve_collision_pairs(panda_table().with_capsules(PANDA_CAPSULES, []),
                                flatten_cost_stack(bad["running_model"], terminal=False), term)
