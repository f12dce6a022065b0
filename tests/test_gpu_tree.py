"""GPU parity of the general-tree kernels (agx_tree.cuh) through the C ABI: the 9-DoF Panda with its finger joints
(BASELINE config 4 as stated: nv = 9, T = 100, two capsule-pair collision costs) against the CPU oracle.

Tolerances of BASELINE.json's north_star: per-node derivatives within 1e-9 relative (normalised PER NODE),
xs / us / cost after a fixed iteration count within 1e-6 relative, identical per-problem decisions.
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from agimus_controller_b200 import PANDA_Q_NOMINAL, _abi, panda_table  # noqa: E402
from agimus_controller_b200.problem import pack_refs  # noqa: E402
from agimus_controller_b200.workloads import goal_reaching_batch, pick_and_place_collision_batch  # noqa: E402

DERIV_RTOL = 1e-9
TRAJ_RTOL = 1e-6


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


@pytest.fixture(scope="module")
def t9():
    return panda_table(lock_fingers=False, armature=0.1)


def _problem(solver_mod, table, dts, refs, B):
    p = solver_mod.BatchedShootingProblem(table, dts, B)
    p.set_refs(refs)
    return p


def _goal9(t9, B, T, seed, orc):
    nv = 9
    m = t9.to_struct()
    rng = np.random.default_rng(seed)
    q = np.concatenate([PANDA_Q_NOMINAL + rng.uniform(-0.3, 0.3, (B, 7)), rng.uniform(0, 0.04, (B, 2))], 1)
    v = rng.uniform(-0.1, 0.1, (B, nv))
    x0 = np.concatenate([q, v], 1)
    xref = np.concatenate([PANDA_Q_NOMINAL, [0.02, 0.02], np.zeros(nv)])
    refs = pack_refs(nv, T, B, xref, np.full(2 * nv, 0.01), np.zeros(nv), np.full(nv, 1e-4),
                     np.diag([1.0, -1.0, -1.0]), np.array([0.5, 0.2, 0.5]), np.full(6, 1e3))
    z = np.zeros((B, nv))
    us = np.repeat(orc.rnea(m, x0[:, :nv], z, z)[:, None, :], T, 1)
    return dict(m=m, refs=refs, dts=np.full(T, 0.01), x0=x0, xs_ws=np.repeat(x0[:, None, :], T + 1, 1),
                us_ws=np.ascontiguousarray(us))


def test_panda9_rnea_integrate(solver_mod, orc, t9):
    m = t9.to_struct()
    p = solver_mod.BatchedShootingProblem(t9, np.full(3, 0.01), 2)
    rng = np.random.default_rng(0)
    q, v, a = rng.uniform(-2, 2, (3, 1000, 9))
    q[:, 7:] = np.abs(q[:, 7:]) * 0.02
    assert rel(p.rnea(q, v, a).cpu().numpy(), orc.rnea(m, q, v, a)) < 1e-12
    x = np.concatenate([q, v], 1)
    assert rel(p.integrate(x, a, 0.01).cpu().numpy(), orc.integrate(m, x, a, 0.01)) < 1e-12


def test_panda9_calc_diff_per_node(solver_mod, orc, t9):
    B, T = 256, 50
    w = _goal9(t9, B, T, 1, orc)
    p = _problem(solver_mod, t9, w["dts"], w["refs"], B)
    rng = np.random.default_rng(11)
    xs = w["xs_ws"] + rng.uniform(-0.2, 0.2, w["xs_ws"].shape)
    us = w["us_ws"] + rng.uniform(-2, 2, w["us_ws"].shape)
    o = orc.calc_diff(w["m"], w["refs"], w["dts"], xs, us)
    g = {k: v.cpu().numpy() for k, v in p.calc_diff(xs, us).items()}
    for k in ("xnext", "Fx", "Fu", "Lx", "Lxx"):
        assert node_rel(g[k], o[k]) < DERIV_RTOL, k
    for k in ("cost", "Lu", "Luu"):
        assert rel(g[k], o[k]) < DERIV_RTOL, k
    cost, xn = p.calc(xs, us)
    assert rel(cost.cpu().numpy(), o["cost"]) < DERIV_RTOL and node_rel(xn.cpu().numpy(), o["xnext"]) < DERIV_RTOL


def test_panda9_solve_fixed_and_converged(solver_mod, orc, t9):
    B, T = 128, 50
    w = _goal9(t9, B, T, 2, orc)
    p = _problem(solver_mod, t9, w["dts"], w["refs"], B)
    for fixed, iters in ((True, 10), (False, 100)):
        opts = _abi.default_fddp_opts(fixed_iters=fixed)
        o = orc.solve(w["m"], w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], iters, opts)
        g = {k: v.cpu().numpy() for k, v in p.solve(w["x0"], w["xs_ws"], w["us_ws"], iters, opts).items()}
        np.testing.assert_array_equal(g["iters"], o["iters"])
        np.testing.assert_array_equal(g["status"], o["status"])
        for k in ("xs", "us", "cost"):
            assert rel(g[k], o[k]) < TRAJ_RTOL, k
        assert rel(g["K"], o["K"]) < 1e-5
    assert (g["status"] == _abi.AGX_STATUS_CONVERGED).mean() > 0.9


def test_cfg4_with_fingers(solver_mod, orc, t9):
    """BASELINE config 4 as stated: nv = 9 with fingers, T = 100, quintic pick-and-place move, two capsule pairs under
    QuadExp.  Per-node derivatives and solves against the CPU restatement on a slab, then the full 4096-problem batch:
    EVERY problem against the oracle after 3 iterations, plus size-independent properties."""
    m9 = t9.to_struct()
    rn = lambda q, v, a: orc.rnea(m9, q, v, a)  # noqa: E731
    B, T = 64, 100
    w = pick_and_place_collision_batch(B, T=T, rnea=rn, alpha=1e-3, w_col=(20.0, 20.0), lock_fingers=False)
    m = w["table"].to_struct()
    assert m.nv == 9 and m.n_pairs == 2
    p = _problem(solver_mod, w["table"], w["dts"], w["refs"], B)
    rng = np.random.default_rng(5)
    xs = w["xs_ws"] + rng.uniform(-0.05, 0.05, w["xs_ws"].shape)
    us = w["us_ws"] + rng.uniform(-2, 2, w["us_ws"].shape)
    o = orc.calc_diff(m, w["refs"], w["dts"], xs, us)
    g = {k: v.cpu().numpy() for k, v in p.calc_diff(xs, us).items()}
    for k in ("xnext", "Fx", "Fu", "Lx", "Lxx"):
        assert node_rel(g[k], o[k]) < DERIV_RTOL, k
    assert rel(g["cost"], o["cost"]) < DERIV_RTOL
    terms = p.cost_terms(xs, us)
    for b, t in ((0, 0), (5, 40), (63, 100)):
        for k in range(2):
            d, _, act = orc.collision(m, xs[b, t, :9], k)
            assert abs(float(terms["collision_distance"][b, t, k]) - d) < 1e-12
    for fixed, iters in ((True, 3), (False, 60)):
        opts = _abi.default_fddp_opts(fixed_iters=fixed)
        so = orc.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], iters, opts)
        sg = {k: v.cpu().numpy() for k, v in p.solve(w["x0"], w["xs_ws"], w["us_ws"], iters, opts).items()}
        np.testing.assert_array_equal(sg["iters"], so["iters"])
        np.testing.assert_array_equal(sg["status"], so["status"])
        for k in ("xs", "us", "cost"):
            assert rel(sg[k], so[k]) < TRAJ_RTOL, k
    # full size, every problem against the oracle
    Bf = 4096
    wf = pick_and_place_collision_batch(Bf, T=T, rnea=rn, alpha=1e-3, w_col=(20.0, 20.0), lock_fingers=False)
    pf = _problem(solver_mod, wf["table"], wf["dts"], wf["refs"], Bf)
    opts = _abi.default_fddp_opts(fixed_iters=True)
    gf = pf.solve(wf["x0"], wf["xs_ws"], wf["us_ws"], 3, opts)
    of = orc.solve(m, wf["refs"], wf["dts"], wf["x0"], wf["xs_ws"], wf["us_ws"], 3, opts)
    np.testing.assert_array_equal(gf["iters"].cpu().numpy(), of["iters"])
    np.testing.assert_array_equal(gf["status"].cpu().numpy(), of["status"])
    for k in ("xs", "us", "cost"):
        assert rel(gf[k].cpu().numpy(), of[k]) < TRAJ_RTOL, k
    assert bool(torch.isfinite(gf["xs"]).all())
    assert float((pf.rollout(wf["x0"], gf["us"]) - gf["xs"]).abs().max()) < 1e-7
    cost_nodes, _ = pf.calc(gf["xs"], gf["us"])
    assert rel(cost_nodes.sum(1).cpu().numpy(), gf["cost"].cpu().numpy()) < 1e-10
    # the fingers follow their opening reference
    assert float(gf["xs"][:, -1, 7:9].mean()) > 0.005


def test_chain7_through_the_tree_kernels(solver_mod, orc):
    """AGX_TREE=1: the 7-joint chain on the general-tree kernels agrees with the tuned chain kernels and the oracle."""
    m7 = panda_table().to_struct()
    B, T = 128, 50
    w = goal_reaching_batch(B, T=T, rnea=lambda q, v, a: orc.rnea(m7, q, v, a))
    opts = _abi.default_fddp_opts(fixed_iters=True)
    pc = _problem(solver_mod, w["table"], w["dts"], w["refs"], B)
    chain = {k: v.cpu().numpy() for k, v in pc.solve(w["x0"], w["xs_ws"], w["us_ws"], 10, opts).items()}
    os.environ["AGX_TREE"] = "1"
    try:
        pt = _problem(solver_mod, w["table"], w["dts"], w["refs"], B)
    finally:
        del os.environ["AGX_TREE"]
    tree = {k: v.cpu().numpy() for k, v in pt.solve(w["x0"], w["xs_ws"], w["us_ws"], 10, opts).items()}
    assert pt.launch_count != pc.launch_count
    o = orc.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 10, opts)
    for k in ("xs", "us", "cost"):
        assert rel(tree[k], chain[k]) < TRAJ_RTOL, k
        assert rel(tree[k], o[k]) < TRAJ_RTOL, k


def test_panda9_shift_and_reference_window(solver_mod, orc, t9):
    B, T = 16, 6
    w = _goal9(t9, B, T, 3, orc)
    dts = np.array([0.01, 0.01, 0.01, 0.02, 0.02, 0.04])
    p = _problem(solver_mod, t9, dts, w["refs"], B)
    xs = p.rollout(w["x0"], w["us_ws"])
    assert rel(xs.cpu().numpy(), orc.rollout(w["m"], w["refs"], dts, w["x0"], w["us_ws"])) < 1e-12
    oxs, ous = p.shift_warmstart(xs, w["us_ws"])
    xs_h = xs.cpu().numpy()
    assert np.array_equal(oxs[:, 0].cpu().numpy(), xs_h[:, 1])
    for i in (3, 4, 5):
        assert rel(oxs[:, i].cpu().numpy(), orc.integrate(w["m"], xs_h[:, i], w["us_ws"][:, i], 0.01)) < 1e-12
    # reference stream window (agx_set_refs_window) with the 9-DoF record size
    n_pts = 40
    stream = np.repeat(w["refs"][0, :1], n_pts, 0)
    stream[:, 0] += 0.01 * np.arange(n_pts)
    p.set_refs_window(torch.as_tensor(stream, device="cuda"), 3)
    hidx = np.concatenate([[0], np.cumsum(np.rint(dts / dts[0]).astype(int))])
    refs_host = np.broadcast_to(stream[3 + hidx][None], (B, T + 1, stream.shape[1])).copy()
    c_dev, _ = p.calc(xs, w["us_ws"])
    c_ref, _ = orc.calc(w["m"], refs_host, dts, xs_h, w["us_ws"])
    # the stream rows are running-node records: the terminal node keeps its control weights but has no control
    assert rel(c_dev.cpu().numpy(), c_ref) < 1e-12


def test_panda9_sqp_mode(solver_mod, orc, t9):
    """agx_solve_sqp (the solver the reference instantiates, unconstrained form) on the 9-DoF tree."""
    B, T = 64, 50
    w = _goal9(t9, B, T, 6, orc)
    p = _problem(solver_mod, t9, w["dts"], w["refs"], B)
    for max_iter in (3, 60):
        o = orc.solve_sqp(w["m"], w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], max_iter)
        g = {k: v.cpu().numpy() for k, v in p.solve_sqp(w["x0"], w["xs_ws"], w["us_ws"], max_iter).items()}
        np.testing.assert_array_equal(g["iters"], o["iters"])
        np.testing.assert_array_equal(g["status"], o["status"])
        for k in ("xs", "us", "cost", "stop"):
            assert rel(g[k], o[k]) < TRAJ_RTOL, k
        assert rel(g["K"], o["K"]) < 1e-5
    assert (g["status"] == _abi.AGX_STATUS_CONVERGED).mean() > 0.9
