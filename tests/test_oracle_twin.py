"""Oracle (C++ restatement) vs the independent numpy twin: two implementations that must agree.

Derivatives are checked with complex-step differentiation of the twin (exact to rounding), the
log maps with scipy's matrix logarithm.
"""
import numpy as np
import pytest

import np_twin as tw
from agimus_controller_b200 import PANDA_Q_NOMINAL, panda_table
from agimus_controller_b200 import _abi
from agimus_controller_b200.workloads import goal_reaching_batch


@pytest.fixture(scope="module", params=[True, False], ids=["panda7", "panda9"])
def table(request):
    return panda_table(lock_fingers=request.param, armature=0.1)


def _rand_state(t, rng):
    q = rng.uniform(-1.0, 1.0, t.nv)
    q[:7] += PANDA_Q_NOMINAL
    if t.nv > 7:
        q[7:] = rng.uniform(0.0, 0.04, t.nv - 7)
    return q, rng.uniform(-1, 1, t.nv), rng.uniform(-3, 3, t.nv)


def test_rnea_crba_match_twin(orc, table):
    rng = np.random.default_rng(1)
    m = table.to_struct()
    for _ in range(5):
        q, v, a = _rand_state(table, rng)
        np.testing.assert_allclose(orc.rnea(m, q, v, a), tw.rnea(table, q, v, a), rtol=0, atol=1e-11)
        M = orc.crba(m, q)
        np.testing.assert_allclose(M, tw.mass_matrix(table, q), rtol=0, atol=1e-12)
        np.testing.assert_allclose(M, M.T, atol=0)


def test_rnea_derivatives_match_complex_step(orc, table):
    rng = np.random.default_rng(2)
    m = table.to_struct()
    for _ in range(5):
        q, v, a = _rand_state(table, rng)
        tau, dq, dv, M = orc.rnea_derivatives(m, q, v, a)
        dq_cs = tw.complex_step_jac(lambda z: tw.rnea(table, z, v, a), q)
        dv_cs = tw.complex_step_jac(lambda z: tw.rnea(table, q, z, a), v)
        scale = max(np.abs(dq_cs).max(), 1.0)
        assert np.abs(dq - dq_cs).max() / scale < 1e-12
        assert np.abs(dv - dv_cs).max() / max(np.abs(dv_cs).max(), 1.0) < 1e-12
        np.testing.assert_allclose(M, tw.mass_matrix(table, q), atol=1e-12)
        np.testing.assert_allclose(tau, tw.rnea(table, q, v, a), atol=1e-11)


def test_forward_dynamics_matches_twin(orc, table):
    rng = np.random.default_rng(3)
    m = table.to_struct()
    q, v, u = _rand_state(table, rng)
    a, Minv = orc.forward_dynamics(m, q, v, u)
    np.testing.assert_allclose(a, tw.forward_dynamics(table, q, v, u), rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(Minv, np.linalg.inv(tw.mass_matrix(table, q) + np.diag(table.armature)), rtol=1e-10,
                               atol=1e-10)


def test_frame_and_log6_match_logm(orc, table):
    rng = np.random.default_rng(4)
    m = table.to_struct()
    for _ in range(5):
        q, _, _ = _rand_state(table, rng)
        R, p = orc.frame_placement(m, q)
        Mt = tw.frame_placement(table, q)
        np.testing.assert_allclose(R, Mt[:3, :3], atol=1e-13)
        np.testing.assert_allclose(p, Mt[:3, 3], atol=1e-13)
        np.testing.assert_allclose(orc.log6(R, p), tw.log6_logm(Mt), atol=1e-9)


def test_jlog6_is_right_jacobian_of_log6(orc):
    rng = np.random.default_rng(5)
    for theta_scale in (1e-6, 0.5, 2.0, 3.1):
        xi = rng.normal(size=6)
        xi[3:] *= theta_scale / np.linalg.norm(xi[3:])
        M = tw.exp6(xi)
        J = orc.Jlog6(M[:3, :3], M[:3, 3])
        h = 1e-6
        Jfd = np.zeros((6, 6))
        for k in range(6):
            e = np.zeros(6)
            e[k] = h
            Mp, Mm = M @ tw.exp6(e), M @ tw.exp6(-e)
            Jfd[:, k] = (orc.log6(Mp[:3, :3], Mp[:3, 3]) - orc.log6(Mm[:3, :3], Mm[:3, 3])) / (2 * h)
        tol = 2e-7 if theta_scale > 3.0 else 2e-8  # FD noise grows as 1/sin(theta)
        assert np.abs(J - Jfd).max() / max(1.0, np.abs(Jfd).max()) < tol, theta_scale


def test_log3_near_pi_branch(orc):
    """Panda's tcp at q = 0 against Rref = I sits on the theta = pi cut (SURVEY §7 hard parts)."""
    m = panda_table().to_struct()
    R, p = orc.frame_placement(m, np.zeros(7))
    r = orc.log6(R, p)
    assert np.isfinite(r).all()
    assert abs(np.linalg.norm(r[3:]) - np.pi) < 1e-5
    J = orc.Jlog6(R, p)
    assert np.isfinite(J).all()


def test_frame_jacobian_matches_complex_step(orc, table):
    rng = np.random.default_rng(6)
    m = table.to_struct()
    q, _, _ = _rand_state(table, rng)
    Jl, Jw = orc.frame_jacobian(m, q)
    Jp = tw.complex_step_jac(lambda z: tw.frame_placement(table, z)[:3, 3], q)
    np.testing.assert_allclose(Jw[:3], Jp, atol=1e-12)
    R, _ = orc.frame_placement(m, q)
    np.testing.assert_allclose(R @ Jl[:3], Jw[:3], atol=1e-13)
    np.testing.assert_allclose(R @ Jl[3:], Jw[3:], atol=1e-13)


def _node_cost_twin(table, ref, x, u, dt, terminal):
    nv = table.nv
    nx = 2 * nv
    xref, wx, uref, wu = ref[:nx], ref[nx:2 * nx], ref[2 * nx:2 * nx + nv], ref[2 * nx + nv:2 * nx + 2 * nv]
    o = 2 * nx + 2 * nv
    Rref, pref, wpose = ref[o:o + 9].reshape(3, 3), ref[o + 9:o + 12], ref[o + 12:o + 18]
    c = 0.5 * np.sum(wx * (x - xref) ** 2)
    if not terminal:
        c += 0.5 * np.sum(wu * (u - uref) ** 2)
    M = tw.frame_placement(table, x[:nv])
    Mr = np.eye(4)
    Mr[:3, :3], Mr[:3, 3] = Rref, pref
    r = tw.log6_logm(np.linalg.inv(Mr) @ M)
    c += 0.5 * np.sum(wpose * r ** 2)
    return c if terminal else dt * c


def test_node_calc_diff_matches_twin(orc):
    """IAM-Euler calc/calcDiff: xnext, cost by direct evaluation; Fx, Fu by complex step of the twin;
    Lx, Lu by central differences of the twin's cost (logm is not complex-safe)."""
    w = goal_reaching_batch(2, T=3, seed=7, rnea=None)
    table = w["table"]
    m = table.to_struct()
    nv, nx = 7, 14
    rng = np.random.default_rng(8)
    xs = w["xs_ws"] + rng.normal(scale=0.05, size=w["xs_ws"].shape)
    us = rng.normal(scale=5.0, size=w["us_ws"].shape)
    out = orc.calc_diff(m, w["refs"], w["dts"], xs, us)
    cost, xnext = orc.calc(m, w["refs"], w["dts"], xs, us)
    np.testing.assert_array_equal(cost, out["cost"])
    np.testing.assert_array_equal(xnext, out["xnext"])
    dt = w["dts"][0]
    for b in range(2):
        for t in range(4):
            term = t == 3
            x = xs[b, t]
            u = us[b, t] if not term else np.zeros(nv)
            ref = w["refs"][b, t]
            assert abs(out["cost"][b, t] - _node_cost_twin(table, ref, x, u, dt, term)) < 1e-9 * max(1, abs(out["cost"][b, t]))
            h = 1e-6
            gx = np.array([(_node_cost_twin(table, ref, x + h * e, u, dt, term) -
                            _node_cost_twin(table, ref, x - h * e, u, dt, term)) / (2 * h) for e in np.eye(nx)])
            assert np.abs(out["Lx"][b, t] - gx).max() < 2e-7 * max(1.0, np.abs(gx).max())
            if term:
                np.testing.assert_array_equal(out["Fx"][b, t], np.eye(nx))
                np.testing.assert_array_equal(out["xnext"][b, t], x)
                continue

            def step(z):
                q, v, uu = z[:nv], z[nv:nx], z[nx:]
                a = tw.forward_dynamics(table, q, v, uu)
                vn = v + dt * a
                return np.concatenate([q + dt * vn, vn])

            z = np.concatenate([x, u])
            np.testing.assert_allclose(out["xnext"][b, t], step(z), rtol=1e-12, atol=1e-12)
            Jz = tw.complex_step_jac(step, z)
            sc = np.abs(Jz).max()
            assert np.abs(out["Fx"][b, t] - Jz[:, :nx]).max() / sc < 1e-11
            assert np.abs(out["Fu"][b, t] - Jz[:, nx:]).max() / sc < 1e-11
            gu = np.array([(_node_cost_twin(table, ref, x, u + h * e, dt, term) -
                            _node_cost_twin(table, ref, x, u - h * e, dt, term)) / (2 * h) for e in np.eye(nv)])
            assert np.abs(out["Lu"][b, t] - gu).max() < 1e-7 * max(1.0, np.abs(gu).max())
            # KAT-5 identities (tests/test_ocp_croco_generic.py:48-52, :68-72): quadratic blocks
            np.testing.assert_allclose(np.diag(out["Luu"][b, t]), dt * ref[2 * nx + nv:2 * nx + 2 * nv], rtol=1e-15)
            np.testing.assert_allclose(np.diag(out["Lxx"][b, t])[nv:], dt * ref[nx + nv:2 * nx], rtol=1e-15)
            assert np.abs(out["Lxx"][b, t] - out["Lxx"][b, t].T).max() < 1e-12 * np.abs(out["Lxx"][b, t]).max()


def test_fddp_converges_and_is_stationary(orc):
    """FDDP on a config-2 style problem: cost decreases monotonically in accepted steps, converged point is feasible
    (rollout reproduces xs) and the end effector reaches the target (KAT-7 style, test_ocp_croco_generic.py:147-221)."""
    w = goal_reaching_batch(2, T=30, seed=3, rnea=None)
    table = w["table"]
    m = table.to_struct()
    z = np.zeros((2, 7))
    u0 = orc.rnea(m, w["x0"][:, :7], z, z)
    us_ws = np.repeat(u0[:, None, :], 30, axis=1)
    opts = _abi.default_fddp_opts()
    res = orc.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], us_ws, 100, opts)
    assert (res["status"] == _abi.AGX_STATUS_CONVERGED).all(), (res["status"], res["iters"], res["stop"])
    xs_roll = orc.rollout(m, w["refs"], w["dts"], w["x0"], res["us"])
    np.testing.assert_allclose(xs_roll, res["xs"], atol=1e-9)
    res10 = orc.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], us_ws, 10, _abi.default_fddp_opts(fixed_iters=True))
    assert (res10["iters"] == 10).all()
    assert (res10["cost"] >= res["cost"] - 1e-9).all()
    for b in range(2):
        _, p = orc.frame_placement(m, res["xs"][b, -1, :7])
        assert np.linalg.norm(p - np.array([0.5, 0.2, 0.5])) < 0.05


def test_frame_translation_and_rotation_residuals(orc):
    """ResidualModelFrameTranslation / FrameRotation (ocp_croco_generic.py:252-357) through the pose slot:
    pose_mode = 1 makes the linear part p_f - pref in the world (Rq = oRf fJf[:3]); with zero linear weights the
    placement record is the rotation residual log3(Rref^T oRf).  Checked against the twin's complex-step kinematics and
    scipy's matrix logarithm."""
    import scipy.linalg

    from agimus_controller_b200.problem import pack_refs

    t = panda_table().with_pose_mode(_abi.AGX_POSE_TRANSLATION_WORLD)
    m = t.to_struct()
    nv = t.nv
    rng = np.random.default_rng(7)
    q = PANDA_Q_NOMINAL + rng.uniform(-0.5, 0.5, nv)
    x = np.concatenate([q, np.zeros(nv)])[None, None]
    pref = np.array([0.4, 0.1, 0.6])
    th = 0.3
    Rref = np.diag([1.0, -1.0, -1.0]) @ np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    w_lin = np.array([3.0, 5.0, 7.0])
    # translation only
    refs = pack_refs(nv, 0, 1, np.zeros(2 * nv), np.zeros(2 * nv), np.zeros(nv), np.zeros(nv), Rref, pref,
                     np.concatenate([w_lin, np.zeros(3)]))
    o = orc.calc_diff(m, refs, np.zeros(0), x, np.zeros((1, 0, nv)))
    pos = lambda z: tw.frame_placement(t, z)[:3, 3]  # noqa: E731
    J = tw.complex_step_jac(pos, q)
    r = pos(q) - pref
    assert abs(o["cost"][0, 0] - 0.5 * np.sum(w_lin * r * r)) < 1e-13
    np.testing.assert_allclose(o["Lx"][0, 0, :nv], J.T @ (w_lin * r), atol=1e-12)
    np.testing.assert_allclose(o["Lxx"][0, 0, :nv, :nv], J.T @ np.diag(w_lin) @ J, atol=1e-12)
    # rotation only (placement record, zero linear weights): r = log3(Rref^T R_f)
    w_ang = np.array([2.0, 4.0, 6.0])
    refs = pack_refs(nv, 0, 1, np.zeros(2 * nv), np.zeros(2 * nv), np.zeros(nv), np.zeros(nv), Rref, pref,
                     np.concatenate([np.zeros(3), w_ang]))
    o = orc.calc_diff(m, refs, np.zeros(0), x, np.zeros((1, 0, nv)))

    def rot_res(z):
        R = tw.frame_placement(t, z)[:3, :3]
        L = np.real(scipy.linalg.logm(Rref.T @ R))
        return np.array([L[2, 1], L[0, 2], L[1, 0]])

    ra = rot_res(q)
    assert abs(o["cost"][0, 0] - 0.5 * np.sum(w_ang * ra * ra)) < 1e-12
    h = 1e-6
    Ja = np.stack([(rot_res(q + h * e) - rot_res(q - h * e)) / (2 * h) for e in np.eye(nv)], axis=1)
    np.testing.assert_allclose(o["Lx"][0, 0, :nv], Ja.T @ (w_ang * ra), atol=1e-7)
    np.testing.assert_allclose(o["Lxx"][0, 0, :nv, :nv], Ja.T @ np.diag(w_ang) @ Ja, atol=1e-6)
