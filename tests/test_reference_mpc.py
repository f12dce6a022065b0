"""The reference's OWN ``MPC`` class, unmodified, driving ``OCPBatchedFDDP``.

``agimus_controller/agimus_controller/mpc.py`` (and ``trajectory.py``, ``warm_start_base.py``, ``ocp_param_base.py``,
``mpc_data.py``, ``ocp_base.py``) are imported from ``/root/reference`` as they are; the only stand-in is a ``pinocchio``
module with the three type names ``trajectory.py:5`` imports.  The OCP is the product class; its device problem is
swapped for the emulator-backed one (same C ABI, the product's kernels compiled for the CPU) because this container
has no GPU — and the GPU box has no ``/root/reference``, where this test skips.  What is checked is the contract
``agimus_controller/tests/test_mpc_unicycle.py:197-263`` pins with a fake OCP: horizon extraction through the
reference's TrajectoryBuffer, ``res.states[0] == x0``, ``res.states[1] == integrate(x0, u0)``, the four timers.
"""
import pathlib
import sys
import types

import numpy as np
import pytest

REF = pathlib.Path("/root/reference/agimus_controller")
YAML = pathlib.Path(__file__).parent / "golden" / "ocp_goal_reaching.yaml"

pytestmark = pytest.mark.skipif(not (REF / "agimus_controller" / "mpc.py").exists(), reason="reference tree absent")


@pytest.fixture(scope="module")
def ref():
    """The reference modules, imported unmodified behind a three-name pinocchio stand-in."""
    class SE3:
        def __init__(self, rotation=None, translation=None):
            self.rotation = np.eye(3) if rotation is None else np.asarray(rotation, dtype=float)
            self.translation = np.zeros(3) if translation is None else np.asarray(translation, dtype=float)

    pin = types.ModuleType("pinocchio")
    pin.SE3, pin.Force, pin.Motion = SE3, type("Force", (), {}), type("Motion", (), {})
    saved = {k: sys.modules.get(k) for k in ("pinocchio",)}
    sys.modules["pinocchio"] = pin
    sys.path.insert(0, str(REF))
    try:
        import importlib

        mods = {n: importlib.import_module(f"agimus_controller.{n}")
                for n in ("mpc", "trajectory", "warm_start_base", "ocp_param_base", "mpc_data", "ocp_base")}
        mods["SE3"] = SE3
        yield mods
    finally:
        sys.path.remove(str(REF))
        for k in [k for k in sys.modules if k == "agimus_controller" or k.startswith("agimus_controller.")]:
            del sys.modules[k]
        if saved["pinocchio"] is None:
            sys.modules.pop("pinocchio", None)
        else:
            sys.modules["pinocchio"] = saved["pinocchio"]


def test_unmodified_mpc_run_drives_the_device_ocp(ref, monkeypatch):
    from agimus_controller_b200 import PANDA_Q_NOMINAL, ocp_batched, panda_table
    from emul.emu_problem import EmuShootingProblem

    monkeypatch.setattr(ocp_batched, "BatchedShootingProblem", EmuShootingProblem)
    tr, pb = ref["trajectory"], ref["ocp_param_base"]
    nv, n_steps, factors = 7, [3, 2, 1], [1, 2, 4]
    T = sum(n_steps)
    seq = pb.DTFactorsNSeq(factors=factors, n_steps=n_steps)
    params = pb.OCPParamsBaseCroco(dt=0.01, solver_iters=3, dt_factor_n_seq=seq, horizon_size=T)
    assert params.timesteps == (0.01, 0.01, 0.01, 0.02, 0.02, 0.04)
    ocp = ocp_batched.OCPBatchedFDDP(panda_table(), params, str(YAML), batch_size=1)
    # a virtual subclass of the REFERENCE's abstract interface: every abstract member is there
    ref["ocp_base"].OCPBase.register(ocp_batched.OCPBatchedFDDP)
    assert isinstance(ocp, ref["ocp_base"].OCPBase)
    missing = [m for m in ref["ocp_base"].OCPBase.__abstractmethods__ if not hasattr(ocp, m)]
    assert not missing, missing

    class ShiftWarmStart(ref["warm_start_base"].WarmStartBase):
        """Previous solution shifted on the device path (agx_shift_warmstart), reference points on the first tick."""

        def setup(self, problem):
            self._p = problem

        def generate(self, initial_state, reference_trajectory):
            x0 = np.concatenate([initial_state.robot_configuration, initial_state.robot_velocity])
            if self._previous_solution is None:
                xs = [x0] + [np.concatenate([p.robot_configuration, p.robot_velocity]) for p in reference_trajectory[1:]]
                us = [np.asarray(p.robot_effort) for p in reference_trajectory[:-1]]
                return x0, xs, us
            prev = self._previous_solution
            xs, us = self._p.shift_warmstart(np.stack(prev.states)[None], np.stack(prev.feed_forward_terms)[None])
            xs = list(xs[0].numpy())
            xs[0] = x0
            return x0, xs, list(us[0].numpy())

    def wpoint(i):
        q = PANDA_Q_NOMINAL + 0.2 * np.sin(2 * np.pi * i * 0.01 / 4.0) * np.ones(nv)
        pt = tr.TrajectoryPoint(id=i, time_ns=i * 10_000_000, robot_configuration=q, robot_velocity=np.zeros(nv),
                                robot_acceleration=np.zeros(nv), robot_effort=np.zeros(nv),
                                end_effector_poses={"panda_hand_tcp": ref["SE3"](np.diag([1.0, -1.0, -1.0]),
                                                                                 np.array([0.5, 0.2, 0.5]))})
        w = tr.TrajectoryPointWeights(w_robot_configuration=np.full(nv, 1.0), w_robot_velocity=np.full(nv, 0.1),
                                      w_robot_acceleration=np.zeros(nv), w_robot_effort=np.full(nv, 1e-3),
                                      w_end_effector_poses={"panda_hand_tcp": np.full(6, 0.1)})
        return tr.WeightedTrajectoryPoint(point=pt, weights=w)

    ws = ShiftWarmStart()
    ws.setup(ocp.problem)
    mpc = ref["mpc"].MPC()
    mpc.setup(ocp, ws, tr.TrajectoryBuffer(seq))
    state = tr.TrajectoryPoint(time_ns=0, robot_configuration=PANDA_Q_NOMINAL.copy(), robot_velocity=np.zeros(nv),
                               robot_acceleration=np.zeros(nv))
    # not enough points yet: MPC.run returns None (mpc.py:38-39); the horizon spans 1 + 3*1 + 2*2 + 1*4 = 12 points
    # (MPC.append_trajectory_points calls TrajectoryBuffer.extend, which the reference's buffer does not have —
    # mpc.py:92 vs trajectory.py:181-231 — so the points go in one by one, as the ROS node does)
    for i in range(5):
        mpc.append_trajectory_point(wpoint(i))
    assert mpc.run(state, 0) is None
    for i in range(5, 40):
        mpc.append_trajectory_point(wpoint(i))
    n_ticks = 4
    for k in range(n_ticks):
        x0 = state.robot_state.copy()
        res = mpc.run(state, k * 10_000_000)
        assert res is not None and len(res.states) == T + 1 and len(res.feed_forward_terms) == T
        assert len(res.ricatti_gains) == T and res.ricatti_gains[0].shape == (nv, 2 * nv)
        np.testing.assert_allclose(res.states[0], x0, atol=1e-12)
        np.testing.assert_allclose(res.states[1], ocp.integrate(x0, res.feed_forward_terms[0]), atol=1e-8)
        dbg = mpc.mpc_debug_data
        assert dbg.duration_ocp_solve_ns > 0 and dbg.duration_iteration_ns >= dbg.duration_ocp_solve_ns
        assert dbg.reference_id == k            # clear_past dropped one point per tick
        assert dbg.ocp.nb_iter >= 1
        state = mpc.integrate(state, res.feed_forward_terms[0])
    # the reference's buffer picked the horizon points at the cumulative step factors (trajectory.py:199-215)
    assert list(mpc._buffer.horizon_indexes) == [0, 1, 2, 3, 5, 7, 11]
    # debug data: the references of the first running node (ocp_croco_generic.py:827-838), XYZQUAT for the SE3 one
    refs = dict(ocp.debug_data.references)
    assert set(refs) == {"control_reg", "state_reg", "goal_tracking"}
    np.testing.assert_allclose(refs["goal_tracking"], [0.5, 0.2, 0.5, 1.0, 0.0, 0.0, 0.0], atol=1e-12)
    assert refs["state_reg"].shape == (2 * nv,)
