"""Pins the CPU restatement (oracle/) on the reference's own golden data and known answers.

KAT numbering follows SURVEY.md §8(c).  Golden: agimus_controller/tests/resources/
simple_ocp_croco_results.pkl, compared by agimus_controller/tests/test_ocp_croco_base.py:175-204;
extracted (without unpickling) by tests/golden/extract_golden.py.
"""
import numpy as np

from agimus_controller_b200 import _abi

from agimus_controller_b200 import PANDA_Q_NOMINAL, panda_table
from agimus_controller_b200.workloads import golden_problem


def test_golden_spot_values(golden):
    assert golden["states"].shape == (10, 14)
    assert golden["ricatti_gains"].shape == (9, 7, 14)
    assert golden["feed_forward_terms"].shape == (9, 7)
    assert abs(golden["states"][9][0] - (-0.0467182276)) < 1e-9
    assert abs(golden["feed_forward_terms"][0][0] - (-1965.9911627546)) < 1e-6


def test_kat1_forward_dynamics_on_golden(orc, golden):
    """(M(q)+0.1 I) a + b(q,v) = u along the golden trajectory, a = (v_{k+1}-v_k)/1e-3; symplectic Euler."""
    m = panda_table(armature=0.1).to_struct()
    xs, us = golden["states"], golden["feed_forward_terms"]
    for k in range(9):
        a, _ = orc.forward_dynamics(m, xs[k, :7], xs[k, 7:], us[k])
        a_gold = (xs[k + 1, 7:] - xs[k, 7:]) / 1e-3
        assert np.abs(a - a_gold).max() / np.abs(a_gold).max() < 1e-9
        np.testing.assert_allclose(xs[k + 1, :7], xs[k, :7] + 1e-3 * xs[k + 1, 7:], atol=1e-15)
        xn = orc.integrate(m, xs[k], us[k], 1e-3)
        assert np.abs(xn - xs[k + 1]).max() < 2e-9  # cancellation-limited (|us| up to 1.4e4)


def test_kat2_golden_is_stationary(orc, golden):
    """The golden point is a KKT point of the restated OCP: |Lu + Fu^T lambda|_inf <= 1e-7, gaps <= 1e-9."""
    p = golden_problem()
    m = p["table"].to_struct()
    _, _, kkt = orc.riccati_sigma(m, p["refs"][0], p["dts"], p["x0"][0], golden["states"],
                                  golden["feed_forward_terms"], 1e-6)
    assert kkt[0] < 1e-7
    assert kkt[1] < 1e-9


CSQP_DIAG = 1e-6 + 1e-9  # SolverCSQP's proximal sigma plus the solver's regularisation at its floor reg_min


def test_kat3_golden_gains_need_csqp_sigma(orc, golden):
    """Golden K equals the Riccati gains at the golden point once CSQP's diagonal terms are in: the proximal
    sigma = 1e-6 AND the regularisation at its floor reg_min = 1e-9, both on Quu, Qxx and Vxx_T.  With 1.001e-6 every
    gain matrix agrees to < 1e-9 relative (1e-11 measured); sigma alone leaves 3..8e-4, no sigma is 60..470 % off."""
    p = golden_problem()
    m = p["table"].to_struct()
    Kg = golden["ricatti_gains"]
    K, _, _ = orc.riccati_sigma(m, p["refs"][0], p["dts"], p["x0"][0], golden["states"],
                                golden["feed_forward_terms"], CSQP_DIAG)
    for t in range(9):
        assert np.abs(K[t] - Kg[t]).max() / np.abs(Kg[t]).max() < 1e-9, t
    K, _, _ = orc.riccati_sigma(m, p["refs"][0], p["dts"], p["x0"][0], golden["states"],
                                golden["feed_forward_terms"], 1e-6)
    for t in range(9):
        assert 1e-4 < np.abs(K[t] - Kg[t]).max() / np.abs(Kg[t]).max() < 1e-3
    K0, _, _ = orc.riccati_sigma(m, p["refs"][0], p["dts"], p["x0"][0], golden["states"],
                                 golden["feed_forward_terms"], 0.0)
    assert np.abs(K0[0] - Kg[0]).max() / np.abs(Kg[0]).max() > 0.5


def test_kat9_csqp_replay_reproduces_the_golden_file(orc, golden):
    """The reference's own check, verbatim (tests/test_ocp_croco_base.py:140-204): solve the test's OCP from the zero
    warm start with the solver the reference instantiates (SolverCSQP, termination_tolerance 1e-3, 100 iterations
    allowed) and compare states / gains / feed-forward terms with the pickle at 6 decimals
    (assert_array_almost_equal: |a - b| < 1.5e-6 on values up to 1.4e4).  The restated iteration stops by the KKT
    criterion after 33 steps exactly where the reference did."""
    p = golden_problem()
    m = p["table"].to_struct()
    o = orc.solve_sqp(m, p["refs"], p["dts"], p["x0"], p["xs_ws"], p["us_ws"], 100)
    assert int(o["status"][0]) == _abi.AGX_STATUS_CONVERGED and int(o["iters"][0]) == 33
    assert float(o["stop"][0]) <= 1e-3
    np.testing.assert_array_almost_equal(o["xs"][0], golden["states"], decimal=6)
    np.testing.assert_array_almost_equal(o["K"][0], golden["ricatti_gains"], decimal=6)
    np.testing.assert_array_almost_equal(o["us"][0], golden["feed_forward_terms"], decimal=6)
    # measured: 9e-10 (states), 5e-8 (controls), 4e-8 (gains)
    assert np.abs(o["xs"][0] - golden["states"]).max() < 1e-8
    assert np.abs(o["us"][0] - golden["feed_forward_terms"]).max() < 5e-7
    # each ingredient matters: no regularisation floor in the QP, or another merit weight, miss the file
    for bad in (dict(reg=0.0), dict(mu=1.0), dict(termination_tolerance=1e-4)):
        opts = _abi.default_sqp_opts()
        for k, v in bad.items():
            setattr(opts, k, v)
        ob = orc.solve_sqp(m, p["refs"], p["dts"], p["x0"], p["xs_ws"], p["us_ws"], 100, opts)
        assert np.abs(ob["xs"][0] - golden["states"]).max() > 1e-5, bad


def test_kat4_ik_6d_known_answer(orc):
    """test_sin_wave_cartesian_space.py:190-218 — IK restated from
    trajectories/sine_wave_cartesian_space.py:62-111: start at q0+0.1, iterate
    dq = -J^T (J J^T)^-1 log6(Mdes^-1 M(q)) (LOCAL Jacobian) to 1e-4, then
    dq = J_LWA^T (J_LWA J_LWA^T)^-1 [0.1,0.2,0.3,0,0,0]; the test stores -dq to 1e-6."""
    m = panda_table(armature=0.0).to_struct()
    q0 = np.array([-0.3619834760502907, -1.3575006398318104, 0.969610481368033, -2.6028532848927295,
                   0.2040785081450368, 1.9436352693107668, 0.6423896937386857])
    R_des, p_des = orc.frame_placement(m, q0)
    q = q0 + 0.1
    for _ in range(10000):
        R, p = orc.frame_placement(m, q)
        err = orc.log6(R_des.T @ R, R_des.T @ (p - p_des))
        if np.linalg.norm(err) < 1e-4:
            break
        Jl, _ = orc.frame_jacobian(m, q)
        q = q - Jl.T @ np.linalg.solve(Jl @ Jl.T, err)
    else:
        raise AssertionError("IK did not converge")
    _, Jw = orc.frame_jacobian(m, q)
    dq = Jw.T @ np.linalg.solve(Jw @ Jw.T, np.array([0.1, 0.2, 0.3, 0.0, 0.0, 0.0]))
    expect = -np.array([0.640289, -0.419278, 0.146452, -1.156815, 0.21497, 0.43003, 0.108381])
    np.testing.assert_allclose(dq, expect, atol=1e-6)


def test_appendix_c_check_values(orc):
    """SURVEY.md Appendix C intermediate values (surveyor's scratch implementation, not reference data)."""
    t = panda_table(armature=0.0)
    m = t.to_struct()
    z = np.zeros(7)
    np.testing.assert_allclose(orc.rnea(m, z, z, z), [0, -4.0398866698, 0, -3.2668560499, 0, 2.2996715606, 0],
                               atol=1e-9)
    np.testing.assert_allclose(np.diag(orc.crba(m, z)), [0.1210851151, 2.8570265213, 0.0837476666, 0.6330609647,
                                                         0.0401557058, 0.0530412366, 0.0066841520], atol=1e-9)
    assert abs(t.mass[6] - 1.495522) < 1e-12
    qn = PANDA_Q_NOMINAL
    np.testing.assert_allclose(orc.rnea(m, qn, z, z), [0, -4.1369383180, -0.6405290606, 22.0167540519, 0.6338666395,
                                                       2.2783370255, 0], atol=1e-9)
    R, p = orc.frame_placement(m, qn)
    np.testing.assert_allclose(p, [0.3083482517, 0, 0.4880749934], atol=1e-9)
    v = np.array([0.1, -0.2, 0.3, -0.4, 0.5, -0.6, 0.7])
    m1 = panda_table(armature=0.1).to_struct()
    a, _ = orc.forward_dynamics(m1, qn, v, np.arange(1.0, 8.0))
    np.testing.assert_allclose(a, [-3.2985189959, -6.4305935203, 3.2042919333, -27.5839018674, 34.4103255304,
                                   46.1667931284, 65.8594064759], atol=1e-8)


def test_kat8_converged_fddp_lands_on_the_golden_solution(orc, golden):
    """KAT-8: the golden file is an (unconverged) iterate of the reference's solver on the OCP of
    tests/test_ocp_croco_base.py; any exact DDP-type solver run to convergence on that OCP must land within 3e-3
    (states) / 0.15 (controls, values up to 1.4e4, i.e. 1e-5 relative) of it, at cost 202.6215."""
    from agimus_controller_b200 import _abi

    w = golden_problem()
    m = w["table"].to_struct()
    o = orc.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 1000, _abi.default_fddp_opts())
    assert o["status"][0] == _abi.AGX_STATUS_CONVERGED and o["iters"][0] < 100
    assert np.abs(o["xs"][0] - golden["states"]).max() < 3e-3
    assert np.abs(o["us"][0] - golden["feed_forward_terms"]).max() < 0.15
    assert abs(o["cost"][0] - 202.6215) < 1e-3
