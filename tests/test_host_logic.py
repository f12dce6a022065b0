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


# ------------------------------------------------------------------ round 2: update flags, Frame* residuals, adapters
def _stack_yaml(running_costs, terminal_costs=None):
    mk = lambda costs: {"class": "IntegratedActionModelEuler",  # noqa: E731
                        "differential": {"class": "DifferentialActionModelFreeFwdDynamics", "costs": costs}}
    return {"running_model": mk(running_costs), "terminal_model": mk(terminal_costs or [])}


def _cost(name, residual, update=True, weight=1.0, weights=None, publish=False):
    act = {"class": "ActivationModelWeightedQuad"}
    if weights is not None:
        act["weights"] = weights
    return {"name": name, "update": update, "weight": weight, "publish_residual": publish,
            "cost": {"class": "CostModelResidual", "activation": act, "residual": residual}}


def _point2(nv=7, pose=None, frame="panda_hand_tcp"):
    from agimus_controller_b200.ocp_interface import SE3, TrajectoryPoint, TrajectoryPointWeights, WeightedTrajectoryPoint

    pose = pose or SE3(np.diag([1.0, -1.0, -1.0]), np.array([0.5, 0.2, 0.5]))
    return WeightedTrajectoryPoint(
        point=TrajectoryPoint(robot_configuration=np.arange(nv) * 0.1, robot_velocity=np.ones(nv), robot_acceleration=np.zeros(nv),
                              robot_effort=np.full(nv, 2.0), end_effector_poses={frame: pose}),
        weights=TrajectoryPointWeights(w_robot_configuration=np.full(nv, 3.0), w_robot_velocity=np.full(nv, 4.0),
                                       w_robot_acceleration=np.zeros(nv), w_robot_effort=np.full(nv, 5.0),
                                       w_end_effector_poses={frame: np.arange(1.0, 7.0)}))


def test_update_false_keeps_the_yaml_reference_and_weights():
    """DifferentialActionModelFreeFwdDynamics.update (ocp_croco_generic.py:712-724) only touches costs with
    `update: true`; the others keep the reference and activation weights the YAML built them with (:97-114, :153-219)."""
    from agimus_controller_b200.ocp_batched import build_reference_rows, flatten_cost_stack

    nv = 7
    xref = list(np.linspace(-1, 1, 2 * nv))
    data = _stack_yaml([
        _cost("state_reg", {"class": "ResidualModelState", "xref": xref}, update=False, weight=2.0, weights=0.5),
        _cost("control_reg", {"class": "ResidualModelControl"}, update=False, weights=list(np.arange(1.0, 8.0))),
        _cost("goal", {"class": "ResidualModelFramePlacement", "id": "panda_hand_tcp",
                       "pref": [0.1, 0.2, 0.3, 0.0, 0.0, 0.0, 1.0]}, update=False, weight=3.0),
    ], [_cost("state_reg", {"class": "ResidualModelState"}, update=True)])
    table = panda_table()
    run = flatten_cost_stack(data["running_model"], False, nv)
    term = flatten_cost_stack(data["terminal_model"], True, nv)
    rows = build_reference_rows(table, run, term, [_point2(), _point2()])
    r0 = rows[0]
    np.testing.assert_allclose(r0[:14], xref)                      # the YAML's xref, not the point's state
    np.testing.assert_allclose(r0[14:28], 2.0 * 0.5)               # CostModelSum weight x scalar activation weight
    np.testing.assert_allclose(r0[28:35], 0.0)                     # ResidualModelControl(state): uref = 0
    np.testing.assert_allclose(r0[35:42], np.arange(1.0, 8.0))
    np.testing.assert_allclose(r0[42:51], np.eye(3).reshape(9))    # pref: identity quaternion
    np.testing.assert_allclose(r0[51:54], [0.1, 0.2, 0.3])
    np.testing.assert_allclose(r0[54:60], 3.0)                     # no weights in the YAML: ones
    # the terminal stack updates: the point's state and weights
    np.testing.assert_allclose(rows[1][:7], np.arange(7) * 0.1)
    np.testing.assert_allclose(rows[1][14:21], 3.0)


def test_frame_translation_rotation_and_static_variants():
    from agimus_controller_b200.ocp_batched import build_reference_rows, flatten_cost_stack, pose_mode_of

    data = _stack_yaml([
        _cost("tr", {"class": "ResidualModelFrameTranslationStatic", "frame_id": "panda_hand_tcp"}, weight=2.0),
        _cost("rot", {"class": "ResidualModelFrameRotation", "id": 0}, weight=10.0),
    ], [_cost("pl", {"class": "ResidualModelFrameRotationStatic", "frame_id": "panda_hand_tcp"})])
    run = flatten_cost_stack(data["running_model"], False)
    term = flatten_cost_stack(data["terminal_model"], True)
    assert pose_mode_of(run, term) == _abi.AGX_POSE_TRANSLATION_WORLD
    rows = build_reference_rows(panda_table(), run, term, [_point2(), _point2()])
    np.testing.assert_allclose(rows[0][54:57], 2.0 * np.array([1.0, 2.0, 3.0]))   # w_end_effector_poses[:3]
    np.testing.assert_allclose(rows[0][57:60], 10.0 * np.array([4.0, 5.0, 6.0]))  # w_end_effector_poses[3:]
    np.testing.assert_allclose(rows[0][51:54], [0.5, 0.2, 0.5])
    np.testing.assert_allclose(rows[1][54:57], 0.0)                               # rotation only on the terminal node
    np.testing.assert_allclose(rows[1][57:60], [4.0, 5.0, 6.0])
    # a placement cost cannot share the record with a translation cost
    bad = _stack_yaml([_cost("tr", {"class": "ResidualModelFrameTranslation", "id": 0})],
                      [_cost("pl", {"class": "ResidualModelFramePlacement", "id": 0})])
    with pytest.raises(NotImplementedError, match="one form"):
        pose_mode_of(flatten_cost_stack(bad["running_model"], False), flatten_cost_stack(bad["terminal_model"], True))
    # a static frame that is not the table's
    other = _stack_yaml([_cost("tr", {"class": "ResidualModelFramePlacementStatic", "frame_id": "panda_link5"})])
    with pytest.raises((NotImplementedError, AssertionError)):
        build_reference_rows(panda_table(), flatten_cost_stack(other["running_model"], False),
                             flatten_cost_stack(other["terminal_model"], True), [_point2(), _point2()])


@pytest.mark.parametrize("cls", ["ResidualModelControlGrav", "ResidualModelFrameVelocity", "ResidualModelFrameVelocityStatic"])
def test_residuals_outside_the_record_structure_are_refused(cls):
    from agimus_controller_b200.ocp_batched import flatten_cost_stack

    data = _stack_yaml([_cost("c", {"class": cls, "id": 0})])
    with pytest.raises(NotImplementedError, match="not supported on the device path"):
        flatten_cost_stack(data["running_model"], False)


def test_visual_servoing_reference_is_the_transform_times_the_target():
    """ResidualModelVisualServoing (ocp_croco_generic.py:434-491): reference = wMo_vision * oMf_target, input key
    `<robot_frame>_vs`, transform requested through input_transforms."""
    from agimus_controller_b200.ocp_batched import flatten_cost_stack, node_references
    from agimus_controller_b200.ocp_interface import SE3

    data = _stack_yaml([_cost("vs", {"class": "ResidualModelVisualServoing", "world_frame": "universe",
                                     "object_frame": "box", "robot_frame": "panda_hand_tcp"})])
    run = flatten_cost_stack(data["running_model"], False)
    d = run["pose"][0]
    assert d["input_key"] == "panda_hand_tcp_vs" and d["transforms_key"] == ("universe", "box")
    target = SE3(np.diag([1.0, -1.0, -1.0]), np.array([0.1, 0.0, 0.2]))
    pt = _point2(pose=target, frame="panda_hand_tcp_vs")
    wMo = SE3(np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]]), np.array([1.0, 2.0, 3.0]))
    r = node_references(panda_table(), run, pt, {("universe", "box"): wMo})
    np.testing.assert_allclose(r["Rref"], wMo.rotation @ target.rotation)
    np.testing.assert_allclose(r["pref"], wMo.translation + wMo.rotation @ target.translation)
    with pytest.raises(AssertionError, match="no transform"):
        node_references(panda_table(), run, pt, {("universe", "box"): None})


class _FakeSE3:
    def __init__(self, R, p):
        self.rotation, self.translation = np.asarray(R, dtype=float), np.asarray(p, dtype=float)


class _FakePinModel:
    """Attribute-for-attribute stand-in of a reduced pinocchio.Model built from the Panda link table."""

    def __init__(self, t):
        class J:
            def __init__(self, short, axis):
                self._s, self.axis = short, axis

            def shortname(self):
                return self._s

        class Y:
            pass

        class F:
            pass

        nv = t.nv
        self.njoints = nv + 1
        self.names = ["universe"] + list(t.joint_names)
        self.parents = [0] + [int(p) + 1 for p in t.parent]
        self.jointPlacements = [_FakeSE3(np.eye(3), np.zeros(3))] + [_FakeSE3(t.placement_R[i], t.placement_p[i]) for i in range(nv)]
        self.joints = [J("JointModelRZ", None)]
        for i in range(nv):
            ax = tuple(t.axis[i])
            if t.jtype[i] == 0:
                self.joints.append(J("JointModelRZ", None) if ax == (0.0, 0.0, 1.0) else J("JointModelRevoluteUnaligned", np.array(ax)))
            else:
                self.joints.append(J("JointModelPY", None) if ax == (0.0, 1.0, 0.0) else J("JointModelPrismaticUnaligned", np.array(ax)))
        self.inertias = [None]
        for i in range(nv):
            y = Y()
            y.mass, y.lever, y.inertia = t.mass[i], t.com[i], t.inertia[i]
            self.inertias.append(y)
        self.frames = []
        for name, (par, R, p) in t.frames.items():
            f = F()
            f.name, f.parentJoint, f.placement = name, par + 1, _FakeSE3(R, p)
            self.frames.append(f)

        class G:
            linear = np.array([0.0, 0.0, -9.81])

        self.gravity = G()


def test_table_from_a_pinocchio_like_model():
    """RobotTable.from_pinocchio_like: the adapter from RobotModels.robot_model / .collision_model / .armature
    (factory/robot_model.py:88-351) to the device table, on a duck-typed model — 9-DoF Panda with its prismatic
    finger joints (one aligned, one unaligned) and a collision model with capsules, a sphere and a box."""
    from agimus_controller_b200.robot_model import RobotTable

    t9 = panda_table(lock_fingers=False, armature=0.1)
    pm = _FakePinModel(t9)

    class Geo:
        pass

    def geom(name, parent_joint, R, p, **shape):
        g = Geo()
        g.name, g.parentJoint, g.placement = name, parent_joint, _FakeSE3(R, p)
        g.geometry = Geo()
        for k, v in shape.items():
            setattr(g.geometry, k, v)
        return g

    class Pair:
        def __init__(self, a, b):
            self.first, self.second = a, b

    class CM:
        geometryObjects = [geom("link3_capsule_0", 3, np.eye(3), [0.0, 0.0, -0.07], radius=0.07, halfLength=0.05),
                           geom("hand_box_0", 7, np.eye(3), [0, 0, 0.1], halfSide=np.ones(3)),
                           geom("link7_capsule_0", 7, np.eye(3), [0.0, 0.0, 0.13], radius=0.06, halfLength=0.07),
                           geom("obstacle_0", 0, np.eye(3), [0.35, 0.0, 0.30], radius=0.05)]
        collisionPairs = [Pair(2, 0), Pair(2, 3), Pair(1, 3)]

    class RM:
        robot_model, collision_model, armature = pm, CM(), np.full(9, 0.1)

    t = RobotTable.from_robot_models(RM(), frame="panda_hand_tcp")
    for f in ("parent", "jtype", "axis", "placement_R", "placement_p", "mass", "com", "inertia", "armature", "gravity"):
        np.testing.assert_allclose(np.asarray(getattr(t, f), dtype=float), np.asarray(getattr(t9, f), dtype=float), atol=0)
    assert t.joint_names == t9.joint_names and t.frame_name == "panda_hand_tcp"
    assert set(t.capsules) == {"link3_capsule_0", "link7_capsule_0", "obstacle_0"}   # the box is skipped
    par, a0, a1, r = t.capsules["link3_capsule_0"]
    assert par == 2 and r == 0.07
    np.testing.assert_allclose(a0, [0, 0, -0.12])
    np.testing.assert_allclose(a1, [0, 0, -0.02])
    par, a0, a1, r = t.capsules["obstacle_0"]                                        # a sphere: zero-length capsule
    assert par == -1 and np.array_equal(a0, a1)
    assert t.collision_pairs == [("link7_capsule_0", "link3_capsule_0"), ("link7_capsule_0", "obstacle_0")]
    m = t.to_struct()
    assert m.nv == 9 and m.n_capsules == 3 and m.n_pairs == 2
    # joints the device state cannot carry are refused
    pm.joints[3] = type(pm.joints[3])("JointModelSpherical", None)
    with pytest.raises(NotImplementedError, match="JointModelSpherical"):
        RobotTable.from_pinocchio_like(pm)


def test_urdf_reader_reproduces_the_panda_table():
    """agimus_controller_b200.urdf.load_urdf: links / joints / inertial origins / locked joints / collision cylinders
    -> capsules named <link>_capsule_<i> (factory/robot_model.py:231-302) on a URDF written from the Panda link table."""
    from agimus_controller_b200.robot_model import PANDA_FRAMES, panda_links
    from agimus_controller_b200.urdf import load_urdf

    out = ['<robot name="panda">', '<link name="world"/>']
    for l in panda_links():
        out.append(f'<link name="{l.name}">')
        if l.mass > 0:
            i = l.inertia
            out.append(f'<inertial><origin xyz="{l.com[0]} {l.com[1]} {l.com[2]}" rpy="0 0 0"/><mass value="{l.mass}"/>'
                       f'<inertia ixx="{i[0]}" ixy="{i[1]}" ixz="{i[2]}" iyy="{i[3]}" iyz="{i[4]}" izz="{i[5]}"/></inertial>')
        if l.name == "panda_link3":
            out.append('<collision><origin xyz="0 0 -0.07" rpy="0 0 0"/><geometry><cylinder radius="0.07" length="0.1"/></geometry></collision>')
            out.append('<collision><origin xyz="0 0 0" rpy="0 0 0"/><geometry><sphere radius="0.07"/></geometry></collision>')
        if l.name == "panda_hand":
            out.append('<collision><origin xyz="0 0 0.02" rpy="0 1.57079632679 0"/><geometry><cylinder radius="0.04" length="0.2"/></geometry></collision>')
        out.append('</link>')
        out.append(f'<joint name="{l.joint_name}" type="{l.joint_type}"><parent link="{l.parent or "world"}"/><child link="{l.name}"/>'
                   f'<origin xyz="{l.xyz[0]} {l.xyz[1]} {l.xyz[2]}" rpy="{l.rpy[0]} {l.rpy[1]} {l.rpy[2]}"/>'
                   f'<axis xyz="{l.axis[0]} {l.axis[1]} {l.axis[2]}"/></joint>')
    for n, (pl, xyz, rpy) in PANDA_FRAMES.items():
        out.append(f'<link name="{n}"/><joint name="{n}_joint" type="fixed"><parent link="{pl}"/><child link="{n}"/>'
                   f'<origin xyz="{xyz[0]} {xyz[1]} {xyz[2]}" rpy="{rpy[0]} {rpy[1]} {rpy[2]}"/></joint>')
    out.append('</robot>')
    xml = "\n".join(out)
    arm = [f"panda_joint{i}" for i in range(1, 8)]
    for moving, ref in ((arm, panda_table(lock_fingers=True)), (None, panda_table(lock_fingers=False))):
        t = load_urdf(xml, moving, frame="panda_hand_tcp", armature=0.1,
                      collision_pairs=[("panda_hand_capsule_0", "panda_link3_capsule_0")])
        for f in ("parent", "jtype", "axis", "placement_R", "placement_p", "mass", "com", "inertia"):
            np.testing.assert_allclose(np.asarray(getattr(t, f), float), np.asarray(getattr(ref, f), float), atol=1e-15)
        np.testing.assert_allclose(t.frames["panda_hand_tcp"][2], ref.frames["panda_hand_tcp"][2], atol=1e-15)
        assert set(t.capsules) == {"panda_link3_capsule_0", "panda_link3_1", "panda_hand_capsule_0"}
        par, a0, a1, r = t.capsules["panda_hand_capsule_0"]     # hand is welded to joint 7: expressed in its frame
        assert par == 6 and r == 0.04 and abs(np.linalg.norm(a1 - a0) - 0.2) < 1e-12
        assert t.to_struct().n_pairs == 1
    with pytest.raises(ValueError, match="not in the model"):
        load_urdf(xml, ["no_such_joint"])


def _random_horizon(n, nv=7, seed=0, frame="panda_hand_tcp", wcol=None):
    from agimus_controller_b200.ocp_interface import SE3, TrajectoryPoint, TrajectoryPointWeights, WeightedTrajectoryPoint

    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        w = TrajectoryPointWeights(w_robot_configuration=rng.uniform(size=nv), w_robot_velocity=rng.uniform(size=nv),
                                   w_robot_acceleration=np.zeros(nv), w_robot_effort=rng.uniform(size=nv),
                                   w_end_effector_poses={frame: rng.uniform(size=6)})
        if wcol is not None:
            w.w_collision_avoidance = float(rng.uniform()) * wcol
        out.append(WeightedTrajectoryPoint(
            point=TrajectoryPoint(id=i, time_ns=i, robot_configuration=rng.normal(size=nv),
                                  robot_velocity=rng.normal(size=nv), robot_acceleration=np.zeros(nv),
                                  robot_effort=rng.normal(size=nv),
                                  end_effector_poses={frame: SE3(R, rng.normal(size=3))}), weights=w))
    return out


def test_vectorised_reference_rows_equal_the_per_node_form():
    """build_reference_rows packs the running nodes with one numpy operation per field (0.09 ms for 20 nodes instead of
    1.5 ms): bit for bit the per-node form, for the goal-reaching stack, the Frame translation / rotation stack and the
    collision stack; stacks with static references take the per-node path either way."""
    from agimus_controller_b200.ocp_batched import build_reference_rows, flatten_cost_stack

    table = panda_table()
    stacks = [yaml.safe_load(GOAL_REACHING.read_text()),
              _stack_yaml([_cost("tr", {"class": "ResidualModelFrameTranslationStatic", "frame_id": "panda_hand_tcp"}, weight=2.0),
                           _cost("rot", {"class": "ResidualModelFrameRotation", "id": 0}, weight=10.0),
                           _cost("x", {"class": "ResidualModelState"}, weight=0.3)],
                          [_cost("pl", {"class": "ResidualModelFrameRotationStatic", "frame_id": "panda_hand_tcp"})]),
              _stack_yaml([_cost("x", {"class": "ResidualModelState", "xref": list(np.zeros(14))}, update=False),
                           _cost("u", {"class": "ResidualModelControl"}, weight=0.5)])]
    for data in stacks:
        run = flatten_cost_stack(data["running_model"], False, 7)
        term = flatten_cost_stack(data["terminal_model"], True, 7)
        for n in (2, 21):
            h = _random_horizon(n, seed=n)
            np.testing.assert_array_equal(build_reference_rows(table, run, term, h),
                                          build_reference_rows(table, run, term, h, vectorised=False))
    data = yaml.safe_load(COLLISION.read_text())
    run = flatten_cost_stack(data["running_model"], False, 7)
    term = flatten_cost_stack(data["terminal_model"], True, 7)
    tcol = resolve_collision_pairs(panda_table().with_capsules(PANDA_CAPSULES, []), run, term)
    h = _random_horizon(11, seed=4, wcol=3.0)
    a = build_reference_rows(tcol, run, term, h)
    np.testing.assert_array_equal(a, build_reference_rows(tcol, run, term, h, vectorised=False))
    assert np.abs(a[:-1, 60:62]).max() > 0
    # the vectorised path keeps the per-node form's refusals
    with pytest.raises(NotImplementedError, match="built for frame"):
        build_reference_rows(table, flatten_cost_stack(stacks[0]["running_model"], False, 7),
                             flatten_cost_stack(stacks[0]["terminal_model"], True, 7), _random_horizon(5, frame="other"))
