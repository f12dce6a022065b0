"""Synthetic workloads of BASELINE.json / SURVEY.md §8(d), as host-side tables.

Each builder returns a dict with ``table`` (RobotTable), ``refs`` [B,T+1,rs], ``dts`` [T], ``x0`` [B,nx],
``xs_ws`` [B,T+1,nx], ``us_ws`` [B,T,nu].  ``rnea`` is the inverse-dynamics callable used for the
gravity-compensation warm start (``warm_start_reference.py:77-87``): the device one in the product
path, the oracle's in CPU tests.
"""
from __future__ import annotations

import numpy as np

from .problem import pack_refs
from .robot_model import PANDA_CAPSULES, PANDA_COLLISION_PAIRS, PANDA_Q_NOMINAL, panda_table

TOOL_DOWN = np.diag([1.0, -1.0, -1.0])  # quaternion x = 1 (dummy_mpc_test.py:104-107)


def goal_reaching_batch(B, T=50, dt=0.01, seed=0, rnea=None, target_R=TOOL_DOWN, target_p=(0.5, 0.2, 0.5),
                        w_q=0.01, w_v=0.01, w_u=1e-4, w_pose=1e3, armature=0.1, q_spread=0.3, v_spread=0.1):
    """Config 2: B Panda goal-reaching OCPs (``ocp_goal_reaching.yaml`` cost stack, weights of
    ``tests/test_ocp_croco_generic.py:182-188``), randomised initial states."""
    table = panda_table(lock_fingers=True, armature=armature)
    nv = table.nv
    rng = np.random.default_rng(seed)
    q0 = PANDA_Q_NOMINAL + rng.uniform(-q_spread, q_spread, size=(B, nv))
    v0 = rng.uniform(-v_spread, v_spread, size=(B, nv))
    x0 = np.concatenate([q0, v0], axis=1)
    xref = np.concatenate([PANDA_Q_NOMINAL, np.zeros(nv)])
    wx = np.concatenate([np.full(nv, w_q), np.full(nv, w_v)])
    refs = pack_refs(nv, T, B, xref, wx, np.zeros(nv), np.full(nv, w_u), np.asarray(target_R), np.asarray(target_p),
                     np.full(6, w_pose))
    dts = np.full(T, dt)
    xs_ws = np.repeat(x0[:, None, :], T + 1, axis=1)
    z = np.zeros_like(q0)
    u0 = rnea(q0, z, z) if rnea is not None else np.zeros_like(q0)
    us_ws = np.repeat(np.asarray(u0).reshape(B, 1, nv), T, axis=1)
    return dict(table=table, refs=refs, dts=dts, x0=x0, xs_ws=np.ascontiguousarray(xs_ws),
                us_ws=np.ascontiguousarray(us_ws))


def golden_problem():
    """The reference's own golden OCP (``tests/test_ocp_croco_base.py:21-98, :140-158``): Panda, T = 9,
    IAM-Euler default step 1e-3, x0 = 0, zero warm start, target SE3(I, [1,1,1])."""
    table = panda_table(lock_fingers=True, armature=0.1)
    nv, T = table.nv, 9
    refs = pack_refs(nv, T, 1, np.zeros(2 * nv), np.full(2 * nv, 0.1), np.zeros(nv), np.full(nv, 1e-4), np.eye(3),
                     np.ones(3), np.ones(6), wpose_terminal=np.full(6, 50.0))
    return dict(table=table, refs=refs, dts=np.full(T, 1e-3), x0=np.zeros((1, 2 * nv)),
                xs_ws=np.zeros((1, T + 1, 2 * nv)), us_ws=np.zeros((1, T, nv)))


def quintic(t, scale_duration):
    """Quintic ramp 0 -> 1 over ``scale_duration`` (trajectories/quintic_trajectory.py:16-42)."""
    s = np.clip(np.asarray(t, dtype=np.float64) / scale_duration, 0.0, 1.0)
    return 10 * s**3 - 15 * s**4 + 6 * s**5


def cartesian_sine_batch(B, T=50, dt=0.01, rnea=None, amplitude=(0.2, 0.2, 0.2), period=4.0, scale_duration=1.0,
                         w_q=1e-2, w_v=1e-2, w_u=1e-4, w_pose=1.0, armature=0.1, t0=1.0):
    """Config 3: end-effector tracking of a Cartesian sine wave with frame-placement residuals
    (trajectories/sine_wave_cartesian_space.py:113-142: pose = initial pose + amplitude * quintic(t) * sin(w t)),
    one phase offset per problem, phi_b = 2 pi b / B.  The horizon starts at time ``t0``.  Joint references stay at the
    nominal posture with small weights (the reference's per-point IK is an input generator, not part of the solve)."""
    table = panda_table(lock_fingers=True, armature=armature)
    nv = table.nv
    R0, p0 = table.frame_placement(PANDA_Q_NOMINAL)
    tt = t0 + dt * np.arange(T + 1)
    w = 2 * np.pi / period
    phi = 2 * np.pi * np.arange(B) / B
    amp = np.asarray(amplitude, dtype=np.float64)
    pref = p0[None, None, :] + amp[None, None, :] * (quintic(tt, scale_duration)[None, :, None]
                                                      * np.sin(w * tt[None, :, None] + phi[:, None, None]))
    x_nom = np.concatenate([PANDA_Q_NOMINAL, np.zeros(nv)])
    z = np.zeros((1, nv))
    u_grav = np.asarray(rnea(PANDA_Q_NOMINAL[None], z, z)).reshape(nv) if rnea is not None else np.zeros(nv)
    refs = pack_refs(nv, T, B, x_nom, np.concatenate([np.full(nv, w_q), np.full(nv, w_v)]), u_grav, np.full(nv, w_u),
                     R0, pref, np.full(6, w_pose))
    x0 = np.repeat(x_nom[None], B, axis=0)
    xs_ws = np.repeat(x0[:, None, :], T + 1, axis=1)
    us_ws = np.repeat(np.broadcast_to(u_grav, (B, 1, nv)), T, axis=1)
    return dict(table=table, refs=refs, dts=np.full(T, dt), x0=x0, xs_ws=np.ascontiguousarray(xs_ws),
                us_ws=np.ascontiguousarray(us_ws))


# measured (x0, u0) operating points of the model-sensibility study are not redistributed; five synthetic points
# around the nominal posture stand in for state_and_control_expe_data.yaml
def model_sensibility_batch(B, T=50, dt=0.01, rnea=None, delta=0.01, seed=5):
    """Config 5: the cfg-2 problem with one inertial parameter perturbed per problem — 7 links x {6 inertia entries,
    3 COM coordinates, mass}, ``delta * s`` with ``s ~ U(-1, 1)`` (evaluate_model_sensibility.py:9-49, :71-73).
    Returns one RobotTable per problem in ``tables``."""
    w = goal_reaching_batch(B, T=T, dt=dt, seed=seed, rnea=rnea)
    base = w["table"]
    rng = np.random.default_rng(seed)
    s = rng.uniform(-1.0, 1.0, size=B)
    tables = [base.perturbed((b % 70) // 10, (b % 70) % 10, delta * s[b]) for b in range(B)]
    w["tables"] = tables
    return w


# start posture of the reference's pick-and-place example (panda_pick_and_place/main.py:81-91), arm joints only
PICK_AND_PLACE_Q_INIT = np.array([-0.3619834760502907, -1.3575006398318104, 0.969610481368033, -2.6028532848927295,
                                  0.2040785081450368, 1.9436352693107668, 0.6423896937386857])


def pick_and_place_collision_batch(B, T=100, dt=0.01, rnea=None, seed=4, alpha=1e-4, w_col=(10.0, 10.0), w_q=3.0,
                                   w_v=0.12, w_u=8e-4, armature=0.1, move_duration=1.0, q_spread=0.05,
                                   capsules=None, pairs=None, lock_fingers=True, finger_opening=0.02):
    """Config 4: joint-space quintic move ``q_init -> q_nominal`` over ``move_duration`` (generic_trajectory.py with
    the weights of panda_pick_and_place/config/trajectory_weigths_params.yaml:4-9) with ``ResidualDistanceCollision`` +
    ``ActivationModelQuadExp(alpha)`` costs on the link7/link3 and link7/obstacle capsule pairs
    (config/agimus_controller_params.yaml:16-20, tests/resources/environment.xacro:23-24).  ``lock_fingers=False`` is
    BASELINE config 4 as stated — nv = 9, the two prismatic finger joints branching off the hand
    (factory/robot_model.py:231-259 locks them only when asked to): the fingers start closed (``q_init`` of
    panda_pick_and_place/main.py:81-91 ends in 0, 0) and are asked to open to ``finger_opening`` along the move.  Capsule
    geometry is the synthetic table of robot_model.PANDA_CAPSULES.  Initial states are ``q_init`` plus a per-problem
    offset on the arm joints."""
    table = panda_table(lock_fingers=lock_fingers, armature=armature).with_capsules(
        PANDA_CAPSULES if capsules is None else capsules, PANDA_COLLISION_PAIRS if pairs is None else pairs, alpha)
    nv = table.nv
    rng = np.random.default_rng(seed)
    off = np.zeros((B, nv))
    off[:, :7] = rng.uniform(-q_spread, q_spread, size=(B, 7))
    q_init = np.concatenate([PICK_AND_PLACE_Q_INIT, np.zeros(nv - 7)])
    q_goal = np.concatenate([PANDA_Q_NOMINAL, np.full(nv - 7, finger_opening)])
    tt = dt * np.arange(T + 1)
    s = np.clip(tt / move_duration, 0.0, 1.0)
    p = 10 * s**3 - 15 * s**4 + 6 * s**5
    dp = np.where(s < 1.0, (30 * s**2 - 60 * s**3 + 30 * s**4) / move_duration, 0.0)
    ddp = np.where(s < 1.0, (60 * s - 180 * s**2 + 120 * s**3) / move_duration**2, 0.0)
    dq = (q_goal - q_init)[None, None, :]
    q = q_init[None, None, :] + off[:, None, :] * (1.0 - p)[None, :, None] + dq * p[None, :, None]
    v = (dq - off[:, None, :]) * dp[None, :, None]
    a = (dq - off[:, None, :]) * ddp[None, :, None]
    u = (np.asarray(rnea(q.reshape(-1, nv), v.reshape(-1, nv), a.reshape(-1, nv))).reshape(B, T + 1, nv)
         if rnea is not None else np.zeros((B, T + 1, nv)))
    xref = np.concatenate([q, v], axis=2)
    wx = np.concatenate([np.full(nv, w_q), np.full(nv, w_v)])
    refs = pack_refs(nv, T, B, xref, wx, u, np.full(nv, w_u), np.eye(3), np.zeros(3), np.zeros(6), wcol=np.asarray(w_col))
    x0 = np.ascontiguousarray(xref[:, 0, :])
    return dict(table=table, refs=refs, dts=np.full(T, dt), x0=x0, xs_ws=np.ascontiguousarray(xref),
                us_ws=np.ascontiguousarray(u[:, :T, :]))


def sine_configuration_reference(n_points, dt=0.01, amplitude=0.2, period=4.0, scale_duration=1.0, rnea=None,
                                 w_q=1.0, w_v=0.1, w_u=1e-3, w_pose=0.1, armature=0.1):
    """Config 1 reference stream: sine wave in configuration space around the nominal posture
    (trajectories/sine_wave_configuration_space.py:41-72, parameters of trajectory_weights_parameters.yaml:27-76):
    q(t) = q_nom + A quintic(t) sin(w t), its derivatives, u = rnea(q, v, a), end-effector pose = FK(q).
    Returns the table and per-point reference records ``[n_points, ref_size]`` (running-node form)."""
    table = panda_table(lock_fingers=True, armature=armature)
    nv = table.nv
    t = dt * np.arange(n_points)
    w = 2 * np.pi / period
    s = np.clip(t / scale_duration, 0.0, 1.0)
    p = 10 * s**3 - 15 * s**4 + 6 * s**5
    dp = np.where(s < 1.0, (30 * s**2 - 60 * s**3 + 30 * s**4) / scale_duration, 0.0)
    ddp = np.where(s < 1.0, (60 * s - 180 * s**2 + 120 * s**3) / scale_duration**2, 0.0)
    sn, cs = np.sin(w * t), np.cos(w * t)
    off = amplitude * p * sn
    doff = amplitude * (dp * sn + p * w * cs)
    ddoff = amplitude * (ddp * sn + 2 * dp * w * cs - p * w * w * sn)
    q = PANDA_Q_NOMINAL[None, :] + off[:, None]
    v = np.repeat(doff[:, None], nv, axis=1)
    a = np.repeat(ddoff[:, None], nv, axis=1)
    u = np.asarray(rnea(q, v, a)).reshape(n_points, nv) if rnea is not None else np.zeros((n_points, nv))
    rows = np.zeros((n_points, 6 * nv + 20))
    for i in range(n_points):
        R, pos = table.frame_placement(q[i])
        rows[i] = pack_refs(nv, 0, 1, np.concatenate([q[i], v[i]]), np.concatenate([np.full(nv, w_q), np.full(nv, w_v)]),
                            u[i], np.full(nv, w_u), R, pos, np.full(6, w_pose))[0, 0]
        rows[i, 5 * nv: 6 * nv] = w_u  # pack_refs zeroes the control weights of its last (terminal) node
    return table, rows, q, v, u
