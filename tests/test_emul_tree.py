"""The general-tree kernels (agx_tree.cuh), compiled for the CPU SIMT emulator, against the oracle.

BASELINE config 4 is the 9-DoF Panda: the two prismatic finger joints branch off the hand
(agimus_controller/agimus_controller/factory/robot_model.py:231-259 locks them only when asked to).  The 7-joint chain
is also pushed through the tree kernels (AGX_TREE=1) and compared with the tuned chain kernels: two different
mappings of the same algorithm.  The GPU twin of this file is tests/test_gpu_tree.py.
"""
import numpy as np
import pytest

from agimus_controller_b200 import PANDA_Q_NOMINAL, _abi, panda_table
from agimus_controller_b200.problem import pack_refs
from agimus_controller_b200.robot_model import Link, RobotTable
from agimus_controller_b200.workloads import goal_reaching_batch, pick_and_place_collision_batch
from emul import emu


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300))


def node_rel(a, b, floor=1e-12):
    """Worst PER-NODE relative error: every (problem, node) block is normalised by its own largest entry."""
    a, b = np.asarray(a), np.asarray(b)
    a2, b2 = a.reshape(a.shape[0] * a.shape[1], -1), b.reshape(b.shape[0] * b.shape[1], -1)
    return float((np.abs(a2 - b2).max(axis=1) / np.maximum(np.abs(b2).max(axis=1), floor)).max())


@pytest.fixture(scope="module")
def t9():
    return panda_table(lock_fingers=False, armature=0.1)


def _states9(rng, n):
    q = np.concatenate([PANDA_Q_NOMINAL + rng.uniform(-1, 1, (n, 7)), rng.uniform(0, 0.04, (n, 2))], 1)
    return q, rng.uniform(-1, 1, (n, 9)), rng.uniform(-3, 3, (n, 9))


def _goal9(t9, B, T, rng, orc, **kw):
    nv = 9
    m = t9.to_struct()
    q, v, _ = _states9(rng, B)
    x0 = np.concatenate([q, 0.1 * v], 1)
    xref = np.concatenate([PANDA_Q_NOMINAL, [0.02, 0.02], np.zeros(nv)])
    refs = pack_refs(nv, T, B, xref, np.full(2 * nv, kw.get("w_x", 0.01)), np.zeros(nv), np.full(nv, kw.get("w_u", 1e-4)),
                     np.diag([1.0, -1.0, -1.0]), np.array([0.5, 0.2, 0.5]), np.full(6, kw.get("w_pose", 1e3)))
    z = np.zeros((B, nv))
    us = np.repeat(orc.rnea(m, x0[:, :nv], z, z)[:, None, :], T, 1)
    return dict(m=m, refs=refs, dts=np.full(T, 0.01), x0=x0, xs_ws=np.repeat(x0[:, None, :], T + 1, 1),
                us_ws=np.ascontiguousarray(us))


def test_panda9_rnea_and_integrate(orc, t9):
    rng = np.random.default_rng(0)
    m = t9.to_struct()
    q, v, a = _states9(rng, 7)
    assert rel(emu.rnea(m, q, v, a), orc.rnea(m, q, v, a)) < 1e-13
    x = np.concatenate([q, v], 1)
    assert rel(emu.integrate(m, x, a, 0.01), orc.integrate(m, x, a, 0.01)) < 1e-13


def test_panda9_calc_diff_per_node(orc, t9):
    """Per-node derivatives within 1e-9 relative, normalised node by node."""
    rng = np.random.default_rng(1)
    B, T = 3, 5
    w = _goal9(t9, B, T, rng, orc)
    xs = w["xs_ws"] + rng.uniform(-0.1, 0.1, w["xs_ws"].shape)
    xs[..., 7:9] = np.abs(xs[..., 7:9])
    us = w["us_ws"] + rng.uniform(-2, 2, w["us_ws"].shape)
    c0, xn0 = orc.calc(w["m"], w["refs"], w["dts"], xs, us)
    c1, xn1 = emu.calc(w["m"], w["refs"], w["dts"], xs, us)
    assert rel(c1, c0) < 1e-12 and rel(xn1, xn0) < 1e-12
    o = orc.calc_diff(w["m"], w["refs"], w["dts"], xs, us)
    e = emu.calc_diff(w["m"], w["refs"], w["dts"], xs, us)
    for k in ("xnext", "Fx", "Fu", "Lx", "Lxx"):
        assert node_rel(e[k], o[k]) < 1e-9, k
    for k in ("cost", "Lu", "Luu"):
        assert rel(e[k], o[k]) < 1e-9, k
    assert np.abs(e["Lxu"]).max() == 0.0
    # the finger joints are prismatic: their rows of Fu are not those of a revolute joint
    assert np.abs(o["Fu"][0, 0, 7, 7]) > 0.0


@pytest.mark.parametrize("fixed,iters", [(True, 3), (False, 40)])
def test_panda9_solve_matches_oracle(orc, t9, fixed, iters):
    rng = np.random.default_rng(2)
    w = _goal9(t9, 3, 8, rng, orc)
    opts = _abi.default_fddp_opts(fixed_iters=fixed)
    o = orc.solve(w["m"], w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], iters, opts)
    e = emu.solve(w["m"], w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], iters, opts)
    np.testing.assert_array_equal(e["iters"], o["iters"])
    np.testing.assert_array_equal(e["status"], o["status"])
    for k in ("xs", "us", "cost", "K", "k"):
        assert rel(e[k], o[k]) < 1e-6, k
    if fixed:
        assert e["launches"] == 4 * iters + 2  # init + (cost records, dynamics records, sweep, forward) per iteration + finalize


def test_panda9_ragged_dts_rollout_shift(orc, t9):
    rng = np.random.default_rng(3)
    w = _goal9(t9, 2, 5, rng, orc)
    dts = np.array([0.01, 0.01, 0.02, 0.02, 0.04])
    xs = emu.rollout(w["m"], w["refs"], dts, w["x0"], w["us_ws"])
    assert rel(xs, orc.rollout(w["m"], w["refs"], dts, w["x0"], w["us_ws"])) < 1e-12
    oxs, ous = emu.shift_warmstart(w["m"], w["refs"], dts, xs, w["us_ws"])
    # fine nodes shift, coarse nodes are re-integrated over dt0 with their own control
    assert np.array_equal(oxs[:, 0], xs[:, 1]) and np.array_equal(oxs[:, 1], xs[:, 2])
    for i in (2, 3, 4):
        ref = orc.integrate(w["m"], xs[:, i], w["us_ws"][:, i], 0.01)
        assert rel(oxs[:, i], ref) < 1e-12
    assert np.array_equal(oxs[:, 5], xs[:, 5])


def test_panda9_collision_costs(orc):
    """cfg 4 as BASELINE.json states it: nv = 9 with fingers, two capsule pairs under QuadExp."""
    rn = lambda q, v, a: orc.rnea(panda_table(lock_fingers=False).to_struct(), q, v, a)  # noqa: E731
    w = pick_and_place_collision_batch(2, T=6, rnea=rn, alpha=1e-3, w_col=(20.0, 20.0), lock_fingers=False)
    m = w["table"].to_struct()
    assert m.nv == 9 and m.n_pairs == 2
    rng = np.random.default_rng(4)
    xs = w["xs_ws"] + rng.uniform(-0.05, 0.05, w["xs_ws"].shape)
    us = w["us_ws"]
    o = orc.calc_diff(m, w["refs"], w["dts"], xs, us)
    e = emu.calc_diff(m, w["refs"], w["dts"], xs, us)
    for k in ("cost", "xnext", "Fx", "Fu", "Lx", "Lxx"):
        assert node_rel(e[k], o[k]) < 1e-9, k
    te = emu.cost_terms(m, w["refs"], w["dts"], xs, us)
    assert np.abs(te[..., 9:11]).max() > 0.0  # the pairs are close enough to cost something
    assert rel(te[..., :3].sum(-1) + te[..., 9:11].sum(-1),
               o["cost"] / np.concatenate([w["dts"], [1.0]])[None, :]) < 1e-10
    opts = _abi.default_fddp_opts(fixed_iters=True)
    so = orc.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 3, opts)
    se = emu.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 3, opts)
    np.testing.assert_array_equal(se["status"], so["status"])
    for k in ("xs", "us", "cost"):
        assert rel(se[k], so[k]) < 1e-6, k


def test_chain7_through_the_tree_kernels(orc, monkeypatch):
    """AGX_TREE=1 sends the 7-joint chain through the general-tree kernels: same results as the tuned chain kernels
    (two mappings of one algorithm) and as the oracle, same FDDP decisions."""
    m7 = panda_table().to_struct()
    w = goal_reaching_batch(3, T=8, rnea=lambda q, v, a: orc.rnea(m7, q, v, a))
    opts = _abi.default_fddp_opts(fixed_iters=False)
    chain = emu.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 40, opts)
    monkeypatch.setenv("AGX_TREE", "1")
    tree = emu.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 40, opts)
    monkeypatch.delenv("AGX_TREE")
    o = orc.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 40, opts)
    np.testing.assert_array_equal(tree["iters"], chain["iters"])
    np.testing.assert_array_equal(tree["iters"], o["iters"])
    for k in ("xs", "us", "cost", "K"):
        assert rel(tree[k], chain[k]) < 1e-6, k
        assert rel(tree[k], o[k]) < 1e-6, k
    assert tree["launches"] != chain["launches"]  # really two paths


def test_six_joint_arm_with_mixed_axes(orc):
    """A 6-joint arm whose joints turn about x, y and z and slide along an oblique axis (JointModelRevoluteUnaligned /
    PrismaticUnaligned territory): dynamics and derivatives through the tree kernels."""
    s = 1.0 / np.sqrt(3.0)
    L = Link
    links = [
        L("l1", None, "j1", "revolute", (0, 0, 0.3), (0, 0, 0), (0, 0, 1), 3.0, (0.01, 0.02, -0.05), (0.1, 0.001, 0.002, 0.12, 0.003, 0.05)),
        L("l2", "l1", "j2", "revolute", (0, 0.1, 0.2), (0.3, 0, 0), (0, 1, 0), 2.0, (0.0, 0.1, 0.02), (0.05, 0, 0.001, 0.04, 0, 0.03)),
        L("l3", "l2", "j3", "revolute", (0.3, 0, 0), (0, 0.2, 0), (1, 0, 0), 1.5, (0.1, 0, 0), (0.02, 0, 0, 0.03, 0.001, 0.03)),
        L("l4", "l3", "j4", "prismatic", (0.1, 0, 0.1), (0, 0, 0.5), (s, s, s), 1.0, (0.02, 0.01, 0), (0.01, 0, 0, 0.01, 0, 0.01)),
        L("l5", "l3", "j5", "revolute", (0, 0.2, 0), (0, 0, 0), (s, -s, s), 0.8, (0, 0.05, 0.01), (0.004, 0, 0, 0.005, 0, 0.006)),
        L("l6", "l5", "j6", "revolute", (0, 0, 0.15), (0.1, 0.2, 0.3), (0, 0, 1), 0.5, (0.01, 0, 0.03), (0.002, 0, 0, 0.002, 0, 0.001)),
    ]
    t6 = RobotTable.from_links(links, (), {"tool": ("l6", (0, 0, 0.1), (0, 0, 0))}, armature=0.05).with_frame("tool")
    m = t6.to_struct()
    assert list(t6.parent) == [-1, 0, 1, 2, 2, 4]
    rng = np.random.default_rng(6)
    q, v, a = rng.uniform(-1, 1, (4, 6)), rng.uniform(-1, 1, (4, 6)), rng.uniform(-2, 2, (4, 6))
    assert rel(emu.rnea(m, q, v, a), orc.rnea(m, q, v, a)) < 1e-13
    B, T, nv = 2, 3, 6
    R, p = t6.frame_placement(np.zeros(6))
    refs = pack_refs(nv, T, B, np.zeros(2 * nv), np.full(2 * nv, 0.1), np.zeros(nv), np.full(nv, 1e-3), R,
                     p + np.array([0.1, -0.1, 0.05]), np.full(6, 10.0))
    xs = rng.uniform(-0.5, 0.5, (B, T + 1, 2 * nv))
    us = rng.uniform(-3, 3, (B, T, nv))
    dts = np.full(T, 0.02)
    o = orc.calc_diff(m, refs, dts, xs, us)
    e = emu.calc_diff(m, refs, dts, xs, us)
    for k in ("cost", "xnext", "Fx", "Fu", "Lx", "Lxx"):
        assert node_rel(e[k], o[k]) < 1e-9, k
    # and whole solves (FDDP, fixed and converged; the sweep's nv = 6 tile shapes)
    x0 = xs[:, 0].copy()
    xs_ws = np.repeat(x0[:, None, :], T + 1, 1)
    z = np.zeros((B, nv))
    us_ws = np.repeat(orc.rnea(m, x0[:, :nv], z, z)[:, None, :], T, 1)
    for fixed, iters in ((True, 3), (False, 30)):
        opts = _abi.default_fddp_opts(fixed_iters=fixed)
        so = orc.solve(m, refs, dts, x0, xs_ws, us_ws, iters, opts)
        se = emu.solve(m, refs, dts, x0, xs_ws, us_ws, iters, opts)
        np.testing.assert_array_equal(se["iters"], so["iters"])
        np.testing.assert_array_equal(se["status"], so["status"])
        for k in ("xs", "us", "cost", "K"):
            assert rel(se[k], so[k]) < 1e-6, k


def test_unsupported_tree_sizes_are_refused():
    L = Link
    links = [L(f"l{i}", None if i == 0 else f"l{i-1}", f"j{i}", "revolute", (0, 0, 0.1), (0, 0, 0), (0, 0, 1), 1.0,
               (0, 0, 0.05), (0.01, 0, 0, 0.01, 0, 0.01)) for i in range(5)]
    t5 = RobotTable.from_links(links, (), {"tool": ("l4", (0, 0, 0.1), (0, 0, 0))}).with_frame("tool")
    with pytest.raises(RuntimeError, match="instantiated for nv"):
        emu.Handle(t5.to_struct(), np.full(3, 0.01), 1, 3)


def test_panda9_frame_translation_mode(orc, t9):
    """pose_mode = 1 (ResidualModelFrameTranslation) on the tree kernels."""
    rng = np.random.default_rng(8)
    t = t9.with_pose_mode(_abi.AGX_POSE_TRANSLATION_WORLD)
    w = _goal9(t, 2, 4, rng, orc, w_pose=30.0)
    xs = w["xs_ws"] + rng.uniform(-0.1, 0.1, w["xs_ws"].shape)
    us = w["us_ws"] + rng.uniform(-1, 1, w["us_ws"].shape)
    o = orc.calc_diff(w["m"], w["refs"], w["dts"], xs, us)
    e = emu.calc_diff(w["m"], w["refs"], w["dts"], xs, us)
    for k in ("cost", "Lx", "Lxx"):
        assert rel(e[k], o[k]) < 1e-9, k
    o0 = orc.calc_diff(t9.to_struct(), w["refs"], w["dts"], xs, us)
    assert rel(o["cost"], o0["cost"]) > 1e-3


def test_panda9_sqp_mode_matches_oracle(orc, t9):
    """The reference's own solver (mim_solvers.SolverCSQP without active constraints, agx_solve_sqp) on the 9-DoF tree:
    same iterates, KKT norms and decisions as the CPU restatement."""
    rng = np.random.default_rng(12)
    w = _goal9(t9, 2, 6, rng, orc)
    for max_iter in (2, 14):
        o = orc.solve_sqp(w["m"], w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], max_iter)
        e = emu.solve_sqp(w["m"], w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], max_iter)
        np.testing.assert_array_equal(e["iters"], o["iters"])
        np.testing.assert_array_equal(e["status"], o["status"])
        for k in ("xs", "us", "cost", "stop"):
            assert rel(e[k], o[k]) < 1e-6, k
        assert rel(e["K"], o["K"]) < 1e-5
    assert (e["iters"] > 0).all() and np.isfinite(e["stop"]).all()
