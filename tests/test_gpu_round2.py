"""GPU tests of the round-2 boundary features, through the C ABI: task-frame translation / rotation residuals,
per-cost gradients, the device-side max_solve_time deadline, per-problem models in integrate / rnea, debug data."""
import pathlib
import time

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from agimus_controller_b200 import PANDA_Q_NOMINAL, _abi, panda_table  # noqa: E402
from agimus_controller_b200.workloads import goal_reaching_batch, model_sensibility_batch  # noqa: E402

YAML = pathlib.Path(__file__).parent / "golden" / "ocp_goal_reaching.yaml"


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def node_rel(a, b, floor=1e-12):
    a, b = np.asarray(a), np.asarray(b)
    a2, b2 = a.reshape(a.shape[0] * a.shape[1], -1), b.reshape(b.shape[0] * b.shape[1], -1)
    return float((np.abs(a2 - b2).max(axis=1) / np.maximum(np.abs(b2).max(axis=1), floor)).max())


@pytest.fixture(scope="module")
def solver_mod():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from agimus_controller_b200 import solver

    return solver


def test_frame_translation_mode_on_the_gpu(solver_mod, orc):
    """pose_mode = 1 (ResidualModelFrameTranslation + FrameRotation sharing the record): per-node derivatives 1e-9
    (normalised per node), 10 fixed iterations 1e-6 with identical decisions, on 256 problems."""
    t = panda_table().with_pose_mode(_abi.AGX_POSE_TRANSLATION_WORLD)
    m = t.to_struct()
    B, T = 256, 50
    w = goal_reaching_batch(B, T=T, rnea=lambda q, v, a: orc.rnea(m, q, v, a), w_pose=200.0)
    p = solver_mod.BatchedShootingProblem(t, w["dts"], B)
    p.set_refs(w["refs"])
    rng = np.random.default_rng(1)
    xs = w["xs_ws"] + rng.uniform(-0.2, 0.2, w["xs_ws"].shape)
    us = w["us_ws"] + rng.uniform(-2, 2, w["us_ws"].shape)
    o = orc.calc_diff(m, w["refs"], w["dts"], xs, us)
    g = {k: v.cpu().numpy() for k, v in p.calc_diff(xs, us).items()}
    for k in ("Lx", "Lxx", "Fx", "Fu", "xnext"):
        assert node_rel(g[k], o[k]) < 1e-9, k
    assert rel(g["cost"], o["cost"]) < 1e-9
    opts = _abi.default_fddp_opts(fixed_iters=True)
    so = orc.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 10, opts)
    sg = {k: v.cpu().numpy() for k, v in p.solve(w["x0"], w["xs_ws"], w["us_ws"], 10, opts).items()}
    np.testing.assert_array_equal(sg["iters"], so["iters"])
    np.testing.assert_array_equal(sg["status"], so["status"])
    for k in ("xs", "us", "cost"):
        assert rel(sg[k], so[k]) < 1e-6, k


def test_per_cost_gradients_on_the_gpu(solver_mod, orc):
    """agx_cost_derivatives (mpc_debugger_node.py:303-323): per-cost w * Lx, w * Lu sum to the node gradient and each
    equals calcDiff with only that cost's weights — chain kernels and (AGX 9-DoF) tree kernels."""
    for lock in (True, False):
        t = panda_table(lock_fingers=lock)
        m = t.to_struct()
        nv = t.nv
        from agimus_controller_b200.workloads import pick_and_place_collision_batch

        B, T = 32, 20
        w = pick_and_place_collision_batch(B, T=T, rnea=lambda q, v, a: orc.rnea(m, q, v, a), alpha=1e-3, w_col=(20.0, 20.0),
                                           lock_fingers=lock)
        mm = w["table"].to_struct()
        refs = w["refs"].copy()
        o0 = 6 * nv
        refs[..., o0:o0 + 9] = np.diag([1.0, -1.0, -1.0]).reshape(9)
        refs[..., o0 + 9:o0 + 12] = [0.5, 0.2, 0.5]
        refs[..., o0 + 12:o0 + 18] = 10.0
        p = solver_mod.BatchedShootingProblem(w["table"], w["dts"], B)
        p.set_refs(refs)
        rng = np.random.default_rng(2)
        xs = w["xs_ws"] + rng.uniform(-0.05, 0.05, w["xs_ws"].shape)
        us = w["us_ws"] + rng.uniform(-1, 1, w["us_ws"].shape)
        d = p.cost_derivatives(xs, us)
        Lx, Lu = d["Lx"].cpu().numpy(), d["Lu"].cpu().numpy()
        s = np.concatenate([w["dts"], [1.0]])[None, :, None]
        o = orc.calc_diff(mm, refs, w["dts"], xs, us)
        assert rel(Lx.sum(2) * s, o["Lx"]) < 1e-9
        assert rel(Lu[:, :-1].sum(2) * s[:, :-1], o["Lu"][:, :-1]) < 1e-9
        keep = {0: slice(2 * nv, 4 * nv), 1: slice(5 * nv, 6 * nv), 2: slice(o0 + 12, o0 + 18), 3: slice(o0 + 18, o0 + 19),
                4: slice(o0 + 19, o0 + 20)}
        for slot, sl in keep.items():
            r1 = refs.copy()
            for other, so_ in keep.items():
                if other != slot:
                    r1[..., so_] = 0.0
            o1 = orc.calc_diff(mm, r1, w["dts"], xs, us)
            assert rel(Lx[:, :, slot] * s, o1["Lx"]) < 1e-9, (lock, slot)
        # the references of the problem are untouched by the masked passes
        c_after, _ = p.calc(xs, us)
        assert rel(c_after.cpu().numpy(), o["cost"]) < 1e-9


def test_max_solve_time_deadline_on_the_device(solver_mod, orc):
    """max_solve_time (ocp_base_croco.py:70-71, :166-171): a device-clock deadline checked at the end of every
    iteration.  A generous budget changes nothing; a budget shorter than one iteration stops every problem after its
    first iteration with AGX_STATUS_TIMEOUT and returns that iterate."""
    m = panda_table().to_struct()
    B, T = 512, 50
    w = goal_reaching_batch(B, T=T, rnea=lambda q, v, a: orc.rnea(m, q, v, a))
    p = solver_mod.BatchedShootingProblem(w["table"], w["dts"], B)
    p.set_refs(w["refs"])
    base = {k: v.clone() for k, v in p.solve(w["x0"], w["xs_ws"], w["us_ws"], 30, _abi.default_fddp_opts()).items()}
    opts = _abi.default_fddp_opts()
    opts.max_solve_time = 10.0
    same = p.solve(w["x0"], w["xs_ws"], w["us_ws"], 30, opts)
    assert torch.equal(same["xs"], base["xs"]) and torch.equal(same["iters"], base["iters"])
    opts.max_solve_time = 20e-6   # one iteration of 512 problems takes far longer
    cut = {k: v.clone() for k, v in p.solve(w["x0"], w["xs_ws"], w["us_ws"], 30, opts).items()}
    assert (cut["status"] == _abi.AGX_STATUS_TIMEOUT).all()
    assert int(cut["iters"].max()) <= 2 and int(cut["iters"].min()) >= 1
    one = p.solve(w["x0"], w["xs_ws"], w["us_ws"], 1, _abi.default_fddp_opts(fixed_iters=True))
    first = cut["iters"] == 1
    assert bool(first.any())
    assert torch.equal(cut["xs"][first], one["xs"][first])
    # SQP mode honours it too
    so = _abi.default_sqp_opts()
    so.max_solve_time = 20e-6
    cs = p.solve_sqp(w["x0"], w["xs_ws"], w["us_ws"], 30, so)
    assert (cs["status"] == _abi.AGX_STATUS_TIMEOUT).all() and int(cs["iters"].max()) <= 2
    # through the OCP class: OCPParamsBaseCroco.max_solve_time, lifted when use_iteration_limits_and_timeout is False
    from agimus_controller_b200.ocp_batched import OCPBatchedFDDP
    from agimus_controller_b200.ocp_interface import DTFactorsNSeq, OCPParamsBaseCroco

    params = OCPParamsBaseCroco(dt=0.01, solver_iters=30, dt_factor_n_seq=DTFactorsNSeq([1], [T]), horizon_size=T,
                                max_solve_time=20e-6)
    ocp = OCPBatchedFDDP(panda_table(), params, str(YAML), batch_size=B)
    ocp.set_reference_table(torch.as_tensor(w["refs"], device="cuda"))
    x0, xs, us = (torch.as_tensor(w[k], device="cuda") for k in ("x0", "xs_ws", "us_ws"))
    ocp.solve(x0, xs, us)
    assert (ocp.ocp_results_batched["status"] == _abi.AGX_STATUS_TIMEOUT).all()
    ocp.solve(x0, xs, us, use_iteration_limits_and_timeout=False)
    assert not (ocp.ocp_results_batched["status"] == _abi.AGX_STATUS_TIMEOUT).any()


def test_per_problem_models_in_integrate_and_rnea(solver_mod, orc):
    """A handle with one inertial table per problem (cfg 5) integrates / inverts row b with model b."""
    B, T = 24, 4
    w = model_sensibility_batch(B, T=T, rnea=None, delta=0.2)
    tables = w["tables"]
    p = solver_mod.BatchedShootingProblem(tables, w["dts"], B)
    rng = np.random.default_rng(4)
    q = PANDA_Q_NOMINAL + rng.uniform(-0.5, 0.5, (B, 7))
    v, a = rng.uniform(-1, 1, (B, 7)), rng.uniform(-2, 2, (B, 7))
    tau = p.rnea(q, v, a).cpu().numpy()
    xn = p.integrate(np.concatenate([q, v], 1), a, 0.01).cpu().numpy()
    for b in range(B):
        mb = tables[b].to_struct()
        assert rel(tau[b], orc.rnea(mb, q[b:b + 1], v[b:b + 1], a[b:b + 1])[0]) < 1e-12
        assert rel(xn[b], orc.integrate(mb, np.concatenate([q[b], v[b]])[None], a[b:b + 1], 0.01)[0]) < 1e-12
    m0 = tables[0].to_struct()
    assert max(rel(tau[b], orc.rnea(m0, q[b:b + 1], v[b:b + 1], a[b:b + 1])[0]) for b in range(B)) > 1e-6  # really different models
    with pytest.raises(RuntimeError, match="exactly B rows"):
        p.rnea(q[:3], v[:3], a[:3])


def test_debug_data_references_and_residuals(solver_mod):
    """OCPDebugData.references / .residuals (ocp_croco_generic.py:814-853): references of the first running node,
    residual predictions [n_controls, nr] of every cost with publish_residual."""
    import yaml

    from agimus_controller_b200.ocp_batched import OCPBatchedFDDP
    from agimus_controller_b200.ocp_interface import (DTFactorsNSeq, OCPParamsBaseCroco, SE3, TrajectoryPoint,
                                                      TrajectoryPointWeights, WeightedTrajectoryPoint)

    data = yaml.safe_load(YAML.read_text())
    for item in data["running_model"]["differential"]["costs"]:
        item["publish_residual"] = True
    T, nv = 10, 7
    params = OCPParamsBaseCroco(dt=0.01, solver_iters=5, dt_factor_n_seq=DTFactorsNSeq([1], [T]), horizon_size=T)
    ocp = OCPBatchedFDDP(panda_table(), params, data, batch_size=1)
    target = SE3(np.diag([1.0, -1.0, -1.0]), np.array([0.5, 0.2, 0.5]))
    pts = [WeightedTrajectoryPoint(
        point=TrajectoryPoint(id=i, robot_configuration=PANDA_Q_NOMINAL + 0.01 * i, robot_velocity=np.zeros(nv),
                              robot_acceleration=np.zeros(nv), robot_effort=np.full(nv, 0.5),
                              end_effector_poses={"panda_hand_tcp": target}),
        weights=TrajectoryPointWeights(w_robot_configuration=np.ones(nv), w_robot_velocity=np.full(nv, 0.1),
                                       w_robot_acceleration=np.zeros(nv), w_robot_effort=np.full(nv, 1e-3),
                                       w_end_effector_poses={"panda_hand_tcp": np.full(6, 1.0)})) for i in range(T + 1)]
    ocp.set_reference_weighted_trajectory(pts)
    x0 = np.concatenate([PANDA_Q_NOMINAL, np.zeros(nv)])
    ocp.solve(x0, [x0] * (T + 1), [np.zeros(nv)] * T)
    dd = ocp.debug_data
    refs, res = dict(dd.references), dict(dd.residuals)
    np.testing.assert_allclose(refs["state_reg"], np.concatenate([PANDA_Q_NOMINAL, np.zeros(nv)]))
    np.testing.assert_allclose(refs["control_reg"], 0.5)
    np.testing.assert_allclose(refs["goal_tracking"], [0.5, 0.2, 0.5, 1, 0, 0, 0], atol=1e-12)
    xs = np.stack(ocp.ocp_results.states)
    us = np.stack(ocp.ocp_results.feed_forward_terms)
    assert res["state_reg"].shape == (T, 2 * nv) and res["control_reg"].shape == (T, nv) and res["goal_tracking"].shape == (T, 6)
    np.testing.assert_allclose(res["control_reg"], us - 0.5, atol=1e-12)
    np.testing.assert_allclose(res["state_reg"][:, :nv], xs[:T, :nv] - np.stack([p.point.robot_configuration for p in pts[:T]]), atol=1e-12)
    # the pose residual is log6(Mref^-1 oMf): its angular part vanishes when the tool points down like the target
    R, p = ocp._table.frame_placement(xs[3, :nv])
    assert np.abs(res["goal_tracking"][3]).max() > 1e-3
