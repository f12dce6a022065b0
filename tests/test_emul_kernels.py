"""The shipped CUDA kernels, compiled for the CPU SIMT emulator (tests/emul), against the oracle.

These run on a machine without a GPU: same translation unit as libagx.so (agx_api.cu + the kernels), with
CUDA threads emulated as fibers.  The GPU twin of this file is tests/test_gpu_parity.py.
"""
import numpy as np
import pytest

from agimus_controller_b200 import PANDA_Q_NOMINAL, _abi, panda_table
from agimus_controller_b200.workloads import goal_reaching_batch, golden_problem
from emul import emu

import pathlib

ROOT = pathlib.Path(__file__).resolve().parent.parent


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300))


@pytest.fixture(scope="module")
def m7():
    return panda_table().to_struct()


def _workload(orc, m, B, T, **kw):
    return goal_reaching_batch(B, T=T, rnea=lambda q, v, a: orc.rnea(m, q, v, a), **kw)


def test_rnea_and_integrate(orc, m7):
    rng = np.random.default_rng(0)
    q = PANDA_Q_NOMINAL + rng.uniform(-1, 1, (9, 7))
    v, a = rng.uniform(-1, 1, (9, 7)), rng.uniform(-3, 3, (9, 7))
    assert rel(emu.rnea(m7, q, v, a), orc.rnea(m7, q, v, a)) < 1e-13
    x = np.concatenate([q, v], 1)
    assert rel(emu.integrate(m7, x, a, 0.01), orc.integrate(m7, x, a, 0.01)) < 1e-13


@pytest.mark.parametrize("target_R", ["tool_down", "identity"])
def test_calc_diff_per_node(orc, m7, target_R):
    B, T = 3, 6
    kw = {} if target_R == "tool_down" else dict(target_R=np.eye(3))
    w = _workload(orc, m7, B, T, **kw)
    rng = np.random.default_rng(1)
    xs = w["xs_ws"] + rng.uniform(-0.2, 0.2, w["xs_ws"].shape)
    us = w["us_ws"] + rng.uniform(-2, 2, w["us_ws"].shape)
    c0, xn0 = orc.calc(m7, w["refs"], w["dts"], xs, us)
    c1, xn1 = emu.calc(m7, w["refs"], w["dts"], xs, us)
    assert rel(c1, c0) < 1e-12 and rel(xn1, xn0) < 1e-12
    o = orc.calc_diff(m7, w["refs"], w["dts"], xs, us)
    e = emu.calc_diff(m7, w["refs"], w["dts"], xs, us)
    for k in ("cost", "xnext", "Fx", "Fu", "Lx", "Lu", "Lxx", "Luu"):
        assert rel(e[k], o[k]) < 1e-9, k
    assert np.abs(e["Lxu"]).max() == 0.0


def test_ragged_dts_and_rollout(orc, m7):
    """Variable step sizes (DTFactorsNSeq, ocp_param_base.py:67-78): 2 x dt, 2 x 2dt, 1 x 4dt."""
    B, T = 2, 5
    w = _workload(orc, m7, B, T)
    dts = np.array([0.01, 0.01, 0.02, 0.02, 0.04])
    xs = emu.rollout(m7, w["refs"], dts, w["x0"], w["us_ws"])
    assert rel(xs, orc.rollout(m7, w["refs"], dts, w["x0"], w["us_ws"])) < 1e-12
    o = orc.calc_diff(m7, w["refs"], dts, xs, w["us_ws"])
    e = emu.calc_diff(m7, w["refs"], dts, xs, w["us_ws"])
    for k in ("Fx", "Fu", "Lx", "Lxx"):
        assert rel(e[k], o[k]) < 1e-9, k


@pytest.mark.parametrize("fixed,iters", [(True, 3), (False, 40)])
def test_solve_matches_oracle(orc, m7, fixed, iters):
    B, T = 3, 12
    w = _workload(orc, m7, B, T)
    opts = _abi.default_fddp_opts(fixed_iters=fixed)
    o = orc.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], iters, opts)
    e = emu.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], iters, opts)
    np.testing.assert_array_equal(e["iters"], o["iters"])
    np.testing.assert_array_equal(e["status"], o["status"])
    for k in ("xs", "us", "cost", "K", "k"):
        assert rel(e[k], o[k]) < 1e-6, k
    if fixed:
        # init + first cost records + 5 launches per round + finalize; two rounds past the budget serve the problems
        # whose line search was deferred (accept_linesearch_kernel)
        assert e["launches"] == 5 * (iters + 2) + 3
    else:
        assert e["launches"] <= 5 * (iters + 2) + 3  # budgets above 32 iterations stop once every problem is done


def test_solve_golden_problem_shapes(orc):
    """The reference's golden OCP (T = 9, dt = 1e-3, ill-conditioned Quu): 3 iterations, same iterates."""
    w = golden_problem()
    m = w["table"].to_struct()
    opts = _abi.default_fddp_opts(fixed_iters=True)
    o = orc.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 3, opts)
    e = emu.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 3, opts)
    assert rel(e["cost"], o["cost"]) < 1e-6
    assert rel(e["xs"], o["xs"]) < 1e-5


def test_regularisation_failure_path(orc, m7):
    """Zero control weight and zero terminal Hessian make Quu singular: the sweep must raise the
    regularisation exactly as the oracle does (same final status / reg decisions)."""
    B, T = 2, 4
    w = _workload(orc, m7, B, T, w_u=0.0, w_pose=0.0, w_q=0.0, w_v=0.0)
    opts = _abi.default_fddp_opts(fixed_iters=True)
    o = orc.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 2, opts)
    e = emu.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 2, opts)
    np.testing.assert_array_equal(e["status"], o["status"])
    np.testing.assert_array_equal(e["iters"], o["iters"])


def test_golden_gains_through_the_shipped_kernels(orc, golden):
    """KAT-1 / KAT-3 of SURVEY.md 8(c) on the product kernels (emulated): the reference's golden states follow
    from its golden controls through `integrate`, and its golden Riccati gains are reproduced by the backward
    sweep with CSQP's diagonal terms, proximal sigma = 1e-6 + regularisation floor 1e-9 (and are far off without)."""
    p = golden_problem()
    m = p["table"].to_struct()
    xs, us, Kg = golden["states"], golden["feed_forward_terms"], golden["ricatti_gains"]
    xn = emu.integrate(m, xs[:9], us, 1e-3)
    assert np.abs(xn - xs[1:]).max() < 2e-9
    K, k, status = emu.riccati(m, p["refs"][0], p["dts"], p["x0"][0], xs, us, 1e-6 + 1e-9)
    assert status != _abi.AGX_STATUS_REGMAX
    for t in range(9):
        assert np.abs(K[t] - Kg[t]).max() / np.abs(Kg[t]).max() < 1e-9
    Ko, ko, _ = orc.riccati_sigma(m, p["refs"][0], p["dts"], p["x0"][0], xs, us, 1e-6 + 1e-9)
    assert rel(K, Ko) < 1e-9 and rel(k, ko) < 1e-6
    K0, _, _ = emu.riccati(m, p["refs"][0], p["dts"], p["x0"][0], xs, us, 0.0)
    assert np.abs(K0[0] - Kg[0]).max() / np.abs(Kg[0]).max() > 0.5


def test_shift_warmstart_matches_the_reference_semantics(orc, m7):
    """warm_start_shift_previous_solution.py:85-104 and tests/test_warm_start_shift_previous_reference.py:108-117:
    timesteps (dt, dt, 2dt, 2dt): fine nodes shift, coarse nodes are re-integrated over dt with their own control."""
    B, T = 2, 4
    w = _workload(orc, m7, B, T)
    dts = np.array([0.01, 0.01, 0.02, 0.02])
    rng = np.random.default_rng(5)
    us = w["us_ws"] + rng.uniform(-1, 1, w["us_ws"].shape)
    xs = orc.rollout(m7, w["refs"], dts, w["x0"], us)
    oxs, ous = emu.shift_warmstart(m7, w["refs"], dts, xs, us)
    for b in range(B):
        exp_x, exp_u = xs[b].copy(), us[b].copy()
        for i, dt in enumerate(dts):
            if dt == dts[0]:
                exp_x[i] = xs[b, i + 1]
                if i < T - 1:
                    exp_u[i] = us[b, i + 1]
            else:
                exp_x[i] = orc.integrate(m7, xs[b, i], us[b, i], dts[0])
        np.testing.assert_allclose(oxs[b], exp_x, rtol=0, atol=1e-12)
        np.testing.assert_allclose(ous[b], exp_u, rtol=0, atol=0)


def test_cost_terms_add_up(orc, m7):
    """Per-cost values: sum of the named costs == node cost / dt (terminal: unscaled); residual identities of
    agimus_controller/tests/test_ocp_croco_generic.py:48-52 (cost == sum 1/2 w r^2)."""
    B, T = 2, 5
    w = _workload(orc, m7, B, T)
    rng = np.random.default_rng(2)
    xs = w["xs_ws"] + rng.uniform(-0.2, 0.2, w["xs_ws"].shape)
    us = w["us_ws"] + rng.uniform(-2, 2, w["us_ws"].shape)
    terms = emu.cost_terms(m7, w["refs"], w["dts"], xs, us)
    cost, _ = orc.calc(m7, w["refs"], w["dts"], xs, us)
    scale = np.concatenate([w["dts"], [1.0]])
    np.testing.assert_allclose(terms[..., :3].sum(-1) * scale, cost, rtol=1e-12)
    wx = w["refs"][..., 14:28]
    np.testing.assert_allclose(terms[..., 0], 0.5 * (wx * (xs - w["refs"][..., :14]) ** 2).sum(-1), rtol=1e-12)
    wp = w["refs"][..., 54:60]
    np.testing.assert_allclose(terms[..., 2], 0.5 * (wp * terms[..., 3:9] ** 2).sum(-1), rtol=1e-12)
    assert np.all(terms[:, -1, 1] == 0.0)


def test_edge_shapes_and_error_codes(orc, m7):
    """Smallest shapes (B = 1, T = 1), a zero iteration budget, and the error paths of the ABI (no exception crosses it:
    bad arguments come back as AGX_EINVAL with a message)."""
    import ctypes as C

    w = _workload(orc, m7, 1, 1)
    opts = _abi.default_fddp_opts(fixed_iters=True)
    o = orc.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 2, opts)
    e = emu.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 2, opts)
    assert rel(e["xs"], o["xs"]) < 1e-9 and rel(e["K"], o["K"]) < 1e-9
    e0 = emu.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 0, opts)
    np.testing.assert_array_equal(e0["xs"], w["xs_ws"])   # no iteration: the warm start comes back
    np.testing.assert_array_equal(e0["iters"], [0])
    lib = emu.lib()
    h = C.c_void_p()
    assert lib.agx_create(None, 1, None, 1, 1, 0, C.byref(h)) == _abi.AGX_EINVAL
    arr = (_abi.AgxModel * 1)(m7)
    dts = np.full(3, 0.01)
    assert lib.agx_create(arr, 2, dts.ctypes.data, 4, 3, 0, C.byref(h)) == _abi.AGX_EINVAL   # n_models must be 1 or B
    lib.agx_destroy(h)
    hd = emu.Handle(m7, dts, 2, 3)
    assert lib.agx_solve(hd.h, None, None, None, 1, None, None, None, None, None, None, None, None, None, None) == _abi.AGX_EINVAL
    x = np.zeros((2, 4, 14)); u = np.zeros((2, 3, 7))
    assert lib.agx_shift_warmstart(hd.h, x.ctypes.data, u.ctypes.data, x.ctypes.data, u.ctypes.data, None) == _abi.AGX_EINVAL
    bad = _abi.default_fddp_opts(); bad.n_alphas = 11
    out = [np.zeros(s) for s in ((2, 4, 14), (2, 3, 7), (2, 3, 7, 14), (2,))]
    it = np.zeros(2, np.int32); stt = np.zeros(2, np.int32)
    rc = lib.agx_solve(hd.h, x[:, 0].copy().ctypes.data, x.ctypes.data, u.ctypes.data, 1, C.byref(bad), out[0].ctypes.data,
                       out[1].ctypes.data, out[2].ctypes.data, None, out[3].ctypes.data, it.ctypes.data, stt.ctypes.data,
                       None, None)
    assert rc == _abi.AGX_EINVAL and b"n_alphas" in lib.agx_last_error(hd.h)
    assert lib.agx_destroy(None) == _abi.AGX_OK


def test_nan_inputs_do_not_poison_neighbours(orc, m7):
    """A problem with a NaN initial state ends with a failure status; its neighbours are solved as usual."""
    B, T = 3, 6
    w = _workload(orc, m7, B, T)
    opts = _abi.default_fddp_opts(fixed_iters=True)
    good = emu.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 3, opts)
    x0 = w["x0"].copy(); xs = w["xs_ws"].copy()
    x0[1, 3] = np.nan; xs[1, :, 3] = np.nan
    e = emu.solve(m7, w["refs"], w["dts"], x0, xs, w["us_ws"], 3, opts)
    assert e["status"][1] == _abi.AGX_STATUS_REGMAX
    for b in (0, 2):
        np.testing.assert_array_equal(e["xs"][b], good["xs"][b])
        assert e["status"][b] == good["status"][b]
    o = orc.solve(m7, w["refs"], w["dts"], x0, xs, w["us_ws"], 3, opts)
    np.testing.assert_array_equal(e["status"], o["status"])
    np.testing.assert_array_equal(e["iters"], o["iters"])


def test_reference_window_follows_the_buffer_horizon_indexes(orc, m7):
    """agx_set_refs_window == TrajectoryBuffer.horizon (trajectory.py:199-222): with factors [1, 2, 4] x n_steps
    [2, 2, 1] the horizon reads points start + [0, 1, 2, 4, 6, 10]; past the end the last point repeats
    (tests/test_buffer.py:82-93 pins the index rule)."""
    from agimus_controller_b200.workloads import sine_configuration_reference

    table, rows, q, v, u = sine_configuration_reference(40, rnea=lambda q_, v_, a_: orc.rnea(m7, q_, v_, a_))
    dts = np.array([0.01, 0.01, 0.02, 0.02, 0.04])
    hidx = np.array([0, 1, 2, 4, 6, 10])
    B, T = 3, 5
    rng = np.random.default_rng(0)
    xs = rng.uniform(-0.3, 0.3, (B, T + 1, 14)) + np.concatenate([q[0], v[0]])
    us = rng.uniform(-2, 2, (B, T, 7))
    start = np.array([0, 7, 33])
    cost = emu.refs_window_cost(m7, dts, rows, start, xs, us)
    refs = np.stack([rows[np.minimum(s0 + hidx, 39)] for s0 in start])
    expect, _ = orc.calc(m7, refs, dts, xs, us)
    np.testing.assert_allclose(cost, expect, rtol=1e-12)
    cost0 = emu.refs_window_cost(m7, dts, rows, 7, xs, us)
    np.testing.assert_allclose(cost0[1], cost[1], rtol=0, atol=0)


def test_octet_sweep_still_matches(orc, m7, tmp_path):
    """The DFMA (octet) Riccati sweep is kept as a cross-check of the tensor-core sweep: AGX_BW=octet selects it at
    library load, so it runs in a subprocess; both must reproduce the oracle's iterates."""
    import os
    import subprocess
    import sys

    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from agimus_controller_b200 import _abi, panda_table
from agimus_controller_b200.workloads import goal_reaching_batch
from emul import emu
from oracle import orc
m = panda_table().to_struct()
w = goal_reaching_batch(3, T=10, rnea=lambda q, v, a: orc.rnea(m, q, v, a))
opts = _abi.default_fddp_opts(fixed_iters=True)
o = orc.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 3, opts)
e = emu.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 3, opts)
for k in ("xs", "us", "K", "cost"):
    assert np.abs(e[k] - o[k]).max() / np.abs(o[k]).max() < 1e-8, k
print("octet ok")
''' % (str(ROOT), str(ROOT / "tests"))
    env = dict(os.environ, AGX_BW="octet")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "octet ok" in r.stdout, r.stderr[-2000:]


def test_two_warp_rollout_matches(orc, m7):
    """The latency mode (eager_exit, at most 64 problems) runs the two-warp forward pass (rollout_try2_kernel);
    AGX_ROLLOUT=2w forces it for every solve at library load, so it runs in a subprocess: it must reproduce the oracle's
    iterates, line searches included."""
    import os
    import subprocess
    import sys

    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from agimus_controller_b200 import _abi, panda_table
from agimus_controller_b200.workloads import goal_reaching_batch
from emul import emu
from oracle import orc
m = panda_table().to_struct()
w = goal_reaching_batch(6, T=10, rnea=lambda q, v, a: orc.rnea(m, q, v, a), q_spread=0.8, target_p=(0.3, -0.4, 0.7))
for fixed, iters in ((True, 3), (False, 16)):
    opts = _abi.default_fddp_opts(fixed_iters=fixed)
    o = orc.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], iters, opts)
    e = emu.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], iters, opts)
    assert (e["iters"] == o["iters"]).all() and (e["status"] == o["status"]).all()
    for k in ("xs", "us", "cost"):
        assert np.abs(e[k] - o[k]).max() / np.abs(o[k]).max() < 1e-8, k
print("two-warp ok")
''' % (str(ROOT), str(ROOT / "tests"))
    env = dict(os.environ, AGX_ROLLOUT="2w")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "two-warp ok" in r.stdout, r.stderr[-2000:]


def test_converged_fddp_on_the_shipped_kernels_lands_on_the_golden_solution(orc, golden):
    """KAT-8 on the product kernels (emulated): zero warm start, run to convergence -> the reference's golden
    states / controls within 3e-3 / 0.15, same iteration count and iterates as the oracle."""
    w = golden_problem()
    m = w["table"].to_struct()
    opts = _abi.default_fddp_opts()
    e = emu.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 60, opts)
    o = orc.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 60, opts)
    assert e["status"][0] == _abi.AGX_STATUS_CONVERGED
    assert np.abs(e["xs"][0] - golden["states"]).max() < 3e-3
    assert np.abs(e["us"][0] - golden["feed_forward_terms"]).max() < 0.15
    assert abs(e["cost"][0] - 202.6215) < 1e-3
    assert abs(int(e["iters"][0]) - int(o["iters"][0])) <= 2
    assert rel(e["xs"], o["xs"]) < 1e-6 and rel(e["cost"], o["cost"]) < 1e-9


def test_long_budget_stops_early(orc, m7):
    """max_iter = 1000 (the controller's first solve): the launch loop ends once every problem has converged."""
    w = _workload(orc, m7, 2, 8)
    opts = _abi.default_fddp_opts()
    o = orc.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 1000, opts)
    e = emu.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 1000, opts)
    np.testing.assert_array_equal(e["iters"], o["iters"])
    np.testing.assert_array_equal(e["status"], o["status"])
    assert rel(e["xs"], o["xs"]) < 1e-6
    assert e["launches"] < 5 * (int(o["iters"].max()) + 16) + 20


# ----------------------------------------------------------------------------------------------------------------
# Collision-distance residuals: capsule pairs + ActivationModelQuadExp.  Parity is against the CPU
# restatement only (colmpc / coal are absent from the reference tree: "parity unpinned" for this row).
@pytest.fixture(scope="module")
def col_case(orc, m7):
    from agimus_controller_b200.robot_model import PANDA_CAPSULES, PANDA_COLLISION_PAIRS
    table = panda_table().with_capsules(PANDA_CAPSULES, PANDA_COLLISION_PAIRS, alpha=0.02)
    B, T = 3, 8
    w = _workload(orc, m7, B, T)
    refs = w["refs"].copy()
    refs[..., 60] = 30.0
    refs[..., 61] = 50.0
    refs[:, T, 60:62] = [0.0, 20.0]   # per-node weights: the terminal node keeps one pair only
    rng = np.random.default_rng(11)
    xs = w["xs_ws"] + rng.uniform(-0.2, 0.2, w["xs_ws"].shape)
    us = w["us_ws"] + rng.uniform(-2, 2, w["us_ws"].shape)
    return dict(w, m=table.to_struct(), refs=refs, xs=xs, us=us, refs_plain=w["refs"])


def test_collision_calc_diff_per_node(orc, col_case):
    c = col_case
    m = c["m"]
    c0, xn0 = orc.calc(m, c["refs"], c["dts"], c["xs"], c["us"])
    c1, xn1 = emu.calc(m, c["refs"], c["dts"], c["xs"], c["us"])
    assert rel(c1, c0) < 1e-12 and rel(xn1, xn0) < 1e-12
    o = orc.calc_diff(m, c["refs"], c["dts"], c["xs"], c["us"])
    e = emu.calc_diff(m, c["refs"], c["dts"], c["xs"], c["us"])
    for k in ("cost", "xnext", "Fx", "Fu", "Lx", "Lu", "Lxx", "Luu"):
        assert rel(e[k], o[k]) < 1e-9, k
    # the collision terms are really there: they change cost, gradient and Hessian, not the dynamics
    z = orc.calc_diff(m, c["refs_plain"], c["dts"], c["xs"], c["us"])
    assert rel(z["Lx"], o["Lx"]) > 1e-3 and rel(z["Lxx"], o["Lxx"]) > 1e-3 and rel(z["Fx"], o["Fx"]) == 0.0


def test_collision_gradient_is_the_distance_derivative(orc, col_case):
    """Rq of the restatement against central differences of its own distance (the oracle's only pin for A10)."""
    m = col_case["m"]
    rng = np.random.default_rng(3)
    for _ in range(6):
        q = PANDA_Q_NOMINAL + rng.uniform(-1.0, 1.0, 7)
        for k in range(2):
            d, Rq, act = orc.collision(m, q, k)
            fd = np.array([(orc.collision(m, q + h, k)[0] - orc.collision(m, q - h, k)[0]) / 2e-6
                           for h in 1e-6 * np.eye(7)])
            assert np.abs(fd - Rq).max() < 1e-8
            a = np.exp(-d * d / m.col_alpha)
            np.testing.assert_allclose(act, [a, -2 * d / m.col_alpha * a, (4 * d * d / m.col_alpha**2 - 2 / m.col_alpha) * a],
                                       rtol=1e-13)


def test_collision_cost_terms(orc, col_case):
    c = col_case
    m = c["m"]
    terms = emu.cost_terms(m, c["refs"], c["dts"], c["xs"], c["us"])
    for b in range(2):
        for t in (0, 3, 8):
            for k in range(2):
                d, _, act = orc.collision(m, c["xs"][b, t, :7], k)
                assert abs(terms[b, t, 11 + k] - d) < 1e-13
                assert abs(terms[b, t, 9 + k] - c["refs"][b, t, 60 + k] * act[0]) < 1e-12 * max(1.0, act[0])
    cost, _ = orc.calc(m, c["refs"], c["dts"], c["xs"], c["us"])
    scale = np.concatenate([c["dts"], [1.0]])
    total = (terms[..., 0] + terms[..., 1] + terms[..., 2] + terms[..., 9] + terms[..., 10]) * scale
    assert rel(total, cost) < 1e-12


@pytest.mark.parametrize("fixed,iters", [(True, 3), (False, 40)])
def test_collision_solve_matches_oracle(orc, col_case, fixed, iters):
    c = col_case
    m = c["m"]
    opts = _abi.default_fddp_opts(fixed_iters=fixed)
    o = orc.solve(m, c["refs"], c["dts"], c["x0"], c["xs_ws"], c["us_ws"], iters, opts)
    e = emu.solve(m, c["refs"], c["dts"], c["x0"], c["xs_ws"], c["us_ws"], iters, opts)
    np.testing.assert_array_equal(e["iters"], o["iters"])
    np.testing.assert_array_equal(e["status"], o["status"])
    for k in ("xs", "us", "cost", "K", "k"):
        assert rel(e[k], o[k]) < 1e-6, k
    z = orc.solve(m, c["refs_plain"], c["dts"], c["x0"], c["xs_ws"], c["us_ws"], iters, opts)
    assert rel(z["xs"], o["xs"]) > 1e-3  # the obstacle changes the plan


def test_collision_mixed_models_in_one_batch(orc, col_case):
    """One model per problem, only some with collision pairs: the COL kernels run for the whole batch and the
    pair-free problems must come out exactly as without them."""
    c = col_case
    plain = panda_table().to_struct()
    models = [c["m"], plain, c["m"]]
    o = orc.calc_diff(models, c["refs"], c["dts"], c["xs"], c["us"])
    e = emu.calc_diff(models, c["refs"], c["dts"], c["xs"], c["us"])
    for k in ("cost", "Lx", "Lxx", "Fx"):
        assert rel(e[k], o[k]) < 1e-9, k
    z = emu.calc_diff(plain, c["refs"], c["dts"], c["xs"], c["us"])
    np.testing.assert_allclose(e["Lxx"][1], z["Lxx"][1], rtol=0, atol=1e-12 * np.abs(z["Lxx"]).max())


def test_collision_model_errors(m7):
    import copy
    from agimus_controller_b200.robot_model import PANDA_CAPSULES, PANDA_COLLISION_PAIRS
    table = panda_table().with_capsules(PANDA_CAPSULES, PANDA_COLLISION_PAIRS)
    bad = table.to_struct()
    bad.pair_b[1] = 3  # capsule 3 does not exist
    w = goal_reaching_batch(1, T=2)
    with pytest.raises(RuntimeError, match="capsule that does not exist"):
        emu.calc(bad, w["refs"], w["dts"], w["xs_ws"], w["us_ws"])
    bad = table.to_struct()
    bad.col_alpha = 0.0
    with pytest.raises(RuntimeError, match="col_alpha"):
        emu.calc(bad, w["refs"], w["dts"], w["xs_ws"], w["us_ws"])


# ----------------------------------------------------------------------------------------------------------------
# SQP mode = mim_solvers.SolverCSQP without active constraints
def test_sqp_reproduces_the_reference_golden_file(golden):
    """KAT-9 on the shipped kernels: the reference's golden test (tests/test_ocp_croco_base.py:140-204) replayed
    through agx_solve_sqp meets the reference's own 6-decimal comparison of states, gains and feed-forward terms."""
    p = golden_problem()
    m = p["table"].to_struct()
    e = emu.solve_sqp(m, p["refs"], p["dts"], p["x0"], p["xs_ws"], p["us_ws"], 100)
    assert int(e["status"][0]) == _abi.AGX_STATUS_CONVERGED and int(e["iters"][0]) == 33
    np.testing.assert_array_almost_equal(e["xs"][0], golden["states"], decimal=6)
    np.testing.assert_array_almost_equal(e["K"][0], golden["ricatti_gains"], decimal=6)
    np.testing.assert_array_almost_equal(e["us"][0], golden["feed_forward_terms"], decimal=6)


@pytest.mark.parametrize("max_iter", [0, 4, 40])
def test_sqp_matches_oracle(orc, m7, max_iter):
    B, T = 3, 10
    w = _workload(orc, m7, B, T)
    o = orc.solve_sqp(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], max_iter)
    e = emu.solve_sqp(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], max_iter)
    np.testing.assert_array_equal(e["iters"], o["iters"])
    np.testing.assert_array_equal(e["status"], o["status"])
    for k in ("xs", "us", "cost", "K", "stop"):
        assert rel(e[k], o[k]) < 1e-6, k
    if max_iter == 40:
        assert (o["status"] == _abi.AGX_STATUS_CONVERGED).all() and (o["stop"] <= 1e-3).all()
        # SQP and FDDP agree on the optimum they approach
        f = orc.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 100)
        assert rel(o["cost"], f["cost"]) < 1e-4


def test_sqp_line_search_failure_and_nan_are_per_problem(orc, m7):
    """A problem whose merit cannot decrease (one step length only, hostile warm start) or that meets a NaN ends with
    its own status; its neighbours are solved as if alone."""
    B, T = 3, 6
    w = _workload(orc, m7, B, T)
    us = w["us_ws"].copy()
    us[1] = np.nan
    opts = _abi.default_sqp_opts()
    o = orc.solve_sqp(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], us, 12, opts)
    e = emu.solve_sqp(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], us, 12, opts)
    np.testing.assert_array_equal(e["status"], o["status"])
    assert e["status"][1] != _abi.AGX_STATUS_CONVERGED and e["status"][0] == _abi.AGX_STATUS_CONVERGED
    alone = emu.solve_sqp(m7, w["refs"][:1], w["dts"], w["x0"][:1], w["xs_ws"][:1], w["us_ws"][:1], 12, opts)
    np.testing.assert_array_equal(e["xs"][0], alone["xs"][0])


def test_eager_exit_gives_the_same_results_with_fewer_launches(orc, m7):
    """eager_exit (the latency mode of a single MPC tick): same iterates and statuses, the launches of the unused part
    of the budget are not queued."""
    B, T = 2, 8
    w = _workload(orc, m7, B, T)
    ref = emu.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 24, _abi.default_fddp_opts())
    opts = _abi.default_fddp_opts()
    opts.eager_exit = 1
    e = emu.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 24, opts)
    # the latency mode evaluates the cost records on the octet path of calc_diff_kernel (the throughput mode on the
    # thread-per-node kernel): the same numbers to rounding, the same decisions
    for k in ("iters", "status"):
        np.testing.assert_array_equal(e[k], ref[k])
    for k in ("xs", "us", "K", "cost"):
        assert rel(e[k], ref[k]) < 1e-9, k
    assert int(ref["iters"].max()) < 24 and e["launches"] < ref["launches"]
    assert e["launches"] <= 5 * (int(ref["iters"].max()) + 1) + 3
    so = _abi.default_sqp_opts()
    sref = emu.solve_sqp(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 12, so)
    so.eager_exit = 1
    se = emu.solve_sqp(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 12, so)
    for k in ("xs", "us", "K", "iters", "status"):
        np.testing.assert_array_equal(se[k], sref[k])
    assert se["launches"] < sref["launches"]


@pytest.mark.parametrize("iters", [1, 4])
def test_deferred_line_search_follows_the_sequential_search(orc, m7, iters):
    """Half of this batch rejects its alpha = 1 trial at some iteration (a few problems need alpha < 1/2 too).  The
    kernels defer the alpha = 1/2 trial to the next round's forward pass and run one round past the budget; per problem
    the result must be SolverFDDP's sequential search: same iterates after every budget, same iteration counts."""
    B, T = 10, 10
    w = goal_reaching_batch(B, T=T, rnea=lambda q, v, a: orc.rnea(m7, q, v, a), q_spread=0.8, target_p=(0.3, -0.4, 0.7))
    opts = _abi.default_fddp_opts(fixed_iters=True)
    o = orc.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], iters, opts)
    e = emu.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], iters, opts)
    np.testing.assert_array_equal(e["iters"], o["iters"])
    np.testing.assert_array_equal(e["status"], o["status"])
    assert (e["iters"] == iters).all()
    for k in ("xs", "us", "cost"):
        assert rel(e[k], o[k]) < 1e-8, k
    assert rel(e["K"], o["K"]) < 1e-5  # the gains amplify rounding through the ill-conditioned Quu
    # the search really runs: without it (one step length only) the oracle ends elsewhere for several problems
    one = _abi.default_fddp_opts(fixed_iters=True)
    one.n_alphas = 1
    o1 = orc.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], iters, one)
    assert (np.abs(o1["xs"] - o["xs"]).max(axis=(1, 2)) > 1e-9).sum() >= 3
    e1 = emu.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], iters, one)
    assert rel(e1["xs"], o1["xs"]) < 1e-8
    two = _abi.default_fddp_opts(fixed_iters=True)
    two.n_alphas = 2
    o2 = orc.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], iters, two)
    e2 = emu.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], iters, two)
    np.testing.assert_array_equal(e2["iters"], o2["iters"])
    assert rel(e2["xs"], o2["xs"]) < 1e-8


def test_moving_an_obstacle_capsule(orc, col_case):
    """update_geometry_placement (ocp_base_croco.py:110-131): new end points of a capsule in the device tables give the
    costs of a model built with the capsule there; an unknown capsule is refused."""
    import copy
    from agimus_controller_b200.robot_model import PANDA_CAPSULES, PANDA_COLLISION_PAIRS

    c = col_case
    a0, a1, radius = (0.30, -0.25, 0.35), (0.42, 0.15, 0.28), 0.04
    caps = dict(PANDA_CAPSULES)
    caps["obstacle_capsule"] = (None, a0, a1, radius)
    moved = panda_table().with_capsules(caps, PANDA_COLLISION_PAIRS, alpha=0.02).to_struct()
    expect, _ = orc.calc(moved, c["refs"], c["dts"], c["xs"], c["us"])
    before, _ = orc.calc(c["m"], c["refs"], c["dts"], c["xs"], c["us"])
    got, _ = emu.calc_with_moved_capsule(c["m"], c["refs"], c["dts"], c["xs"], c["us"], 2, a0, a1, radius)
    assert rel(got, expect) < 1e-12 and rel(before, expect) > 1e-4
    with pytest.raises(RuntimeError, match="no such capsule"):
        emu.calc_with_moved_capsule(c["m"], c["refs"], c["dts"], c["xs"], c["us"], 3, a0, a1, radius)


def test_frame_translation_mode_matches_oracle(orc):
    """pose_mode = 1 (ResidualModelFrameTranslation: world-frame p_f - pref for the linear part, log3 for the angular
    part) through both cost paths of the chain kernels: the octet one (agx_calc_diff) and the thread-per-node one (the
    solve's node_cost_kernel)."""
    t = panda_table().with_pose_mode(_abi.AGX_POSE_TRANSLATION_WORLD)
    m = t.to_struct()
    B, T = 3, 6
    w = _workload(orc, m, B, T, w_pose=50.0)
    rng = np.random.default_rng(3)
    xs = w["xs_ws"] + rng.uniform(-0.2, 0.2, w["xs_ws"].shape)
    us = w["us_ws"] + rng.uniform(-2, 2, w["us_ws"].shape)
    o = orc.calc_diff(m, w["refs"], w["dts"], xs, us)
    e = emu.calc_diff(m, w["refs"], w["dts"], xs, us)
    for k in ("cost", "Lx", "Lxx"):
        assert rel(e[k], o[k]) < 1e-9, k
    # not the placement residual
    o0 = orc.calc_diff(panda_table().to_struct(), w["refs"], w["dts"], xs, us)
    assert rel(o["cost"], o0["cost"]) > 1e-3
    opts = _abi.default_fddp_opts(fixed_iters=True)
    so = orc.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 3, opts)
    se = emu.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 3, opts)
    for k in ("xs", "us", "cost"):
        assert rel(se[k], so[k]) < 1e-6, k
    te = emu.cost_terms(m, w["refs"], w["dts"], xs, us)
    R, p = t.frame_placement(xs[1, 2, :7])
    assert np.abs(te[1, 2, 3:6] - (p - w["refs"][1, 2, 51:54])).max() < 1e-12


def test_per_cost_derivatives(orc):
    """agx_cost_derivatives: the per-cost gradients the debugger reads (mpc_debugger_node.py:303-323) — each named
    cost's unscaled Lx / Lu; they sum to the node's total gradient, and each equals the oracle's calcDiff on a
    reference table that keeps only that cost's weights."""
    from agimus_controller_b200.workloads import pick_and_place_collision_batch

    m0 = panda_table().to_struct()
    w = pick_and_place_collision_batch(2, T=4, rnea=lambda q, v, a: orc.rnea(m0, q, v, a), alpha=1e-3, w_col=(20.0, 20.0))
    m = w["table"].to_struct()
    refs = w["refs"].copy()
    refs[..., 42:51] = np.diag([1.0, -1.0, -1.0]).reshape(9)
    refs[..., 51:54] = [0.5, 0.2, 0.5]
    refs[..., 54:60] = 10.0
    rng = np.random.default_rng(9)
    xs = w["xs_ws"] + rng.uniform(-0.05, 0.05, w["xs_ws"].shape)
    us = w["us_ws"] + rng.uniform(-1, 1, w["us_ws"].shape)
    Lx, Lu = emu.cost_derivatives(m, refs, w["dts"], xs, us)
    s = np.concatenate([w["dts"], [1.0]])[None, :, None]
    o = orc.calc_diff(m, refs, w["dts"], xs, us)
    assert rel(Lx.sum(2) * s, o["Lx"]) < 1e-9
    assert rel(Lu[:, :-1].sum(2) * s[:, :-1], o["Lu"][:, :-1]) < 1e-9
    keep = {0: slice(14, 28), 1: slice(35, 42), 2: slice(54, 60), 3: slice(60, 61), 4: slice(61, 62)}
    for slot, sl in keep.items():
        r1 = refs.copy()
        for other, so in keep.items():
            if other != slot:
                r1[..., so] = 0.0
        o1 = orc.calc_diff(m, r1, w["dts"], xs, us)
        assert rel(Lx[:, :, slot] * s, o1["Lx"]) < 1e-9, slot
        assert np.abs(Lx[:, :, slot]).max() > 0.0 or slot == 1, slot


def test_deferred_line_search_two_levels_matches_oracle(orc, m7):
    """Problems whose alpha = 1 AND alpha = 1/2 trials are rejected: the deferred search (two levels, then in line)
    must reproduce the sequential search of the oracle decision for decision."""
    B, T = 4, 10
    w = _workload(orc, m7, B, T, w_pose=1e5, target_p=(0.9, 0.6, 0.9), q_spread=0.8, v_spread=1.0)
    opts = _abi.default_fddp_opts(fixed_iters=True)
    o = orc.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 8, opts)
    # the scenario really needs short steps: with alpha = 1 only, the outcome differs
    o1 = _abi.default_fddp_opts(fixed_iters=True)
    o1.n_alphas = 1
    o2 = _abi.default_fddp_opts(fixed_iters=True)
    o2.n_alphas = 2
    assert rel(orc.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 8, o1)["cost"], o["cost"]) > 1e-6
    assert rel(orc.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 8, o2)["cost"], o["cost"]) > 1e-6
    e = emu.solve(m7, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 8, opts)
    np.testing.assert_array_equal(e["iters"], o["iters"])
    np.testing.assert_array_equal(e["status"], o["status"])
    for k in ("xs", "us", "cost"):
        assert rel(e[k], o[k]) < 1e-6, k
