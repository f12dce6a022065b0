"""ctypes loader for the CPU restatement ``oracle/liborc.so``.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs import this module; nothing under ``agimus_controller_b200/`` does.
"""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess

import numpy as np

from agimus_controller_b200 import _abi

_HERE = pathlib.Path(__file__).resolve().parent
_LIB = None
_P = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int32)


def build() -> None:
    subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = _HERE / "liborc.so"
        if not so.exists():
            build()
        _LIB = C.CDLL(str(so))
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(_P)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _models(models):
    if isinstance(models, _abi.AgxModel):
        return C.byref(models), 1, models.nv
    arr = (_abi.AgxModel * len(models))(*models)
    return arr, len(models), models[0].nv


def rnea(m, q, v, a):
    q, v, a = _c(q), _c(v), _c(a)
    n = q.size // m.nv
    tau = np.zeros_like(q)
    lib().orc_rnea(C.byref(m), _p(q), _p(v), _p(a), n, _p(tau))
    return tau


def crba(m, q):
    M = np.zeros((m.nv, m.nv))
    lib().orc_crba(C.byref(m), _p(_c(q)), _p(M))
    return M


def rnea_derivatives(m, q, v, a):
    nv = m.nv
    tau, dq, dv, M = np.zeros(nv), np.zeros((nv, nv)), np.zeros((nv, nv)), np.zeros((nv, nv))
    lib().orc_rnea_derivatives(C.byref(m), _p(_c(q)), _p(_c(v)), _p(_c(a)), _p(tau), _p(dq), _p(dv), _p(M))
    return tau, dq, dv, M


def forward_dynamics(m, q, v, u):
    nv = m.nv
    a, Minv = np.zeros(nv), np.zeros((nv, nv))
    rc = lib().orc_forward_dynamics(C.byref(m), _p(_c(q)), _p(_c(v)), _p(_c(u)), _p(a), _p(Minv))
    assert rc == 0
    return a, Minv


def frame_placement(m, q):
    R, p = np.zeros((3, 3)), np.zeros(3)
    lib().orc_frame_placement(C.byref(m), _p(_c(q)), _p(R), _p(p))
    return R, p


def frame_jacobian(m, q):
    Jl, Jw = np.zeros((6, m.nv)), np.zeros((6, m.nv))
    lib().orc_frame_jacobian(C.byref(m), _p(_c(q)), _p(Jl), _p(Jw))
    return Jl, Jw


def collision(m, q, k=0):
    """Signed capsule distance of collision pair k, its gradient and the QuadExp activation (a, a', a'')."""
    d, Rq, act = np.zeros(1), np.zeros(m.nv), np.zeros(3)
    lib().orc_collision(C.byref(m), _p(_c(q)), int(k), _p(d), _p(Rq), _p(act))
    return float(d[0]), Rq, act


def log6(R, p):
    out = np.zeros(6)
    lib().orc_log6(_p(_c(R)), _p(_c(p)), _p(out))
    return out


def Jlog6(R, p):
    J = np.zeros((6, 6))
    lib().orc_Jlog6(_p(_c(R)), _p(_c(p)), _p(J))
    return J


def calc(models, refs, dts, xs, us):
    mp, nm, nv = _models(models)
    refs, dts, xs, us = _c(refs), _c(dts), _c(xs), _c(us)
    B, T1, nx = xs.shape
    T = T1 - 1
    cost, xnext = np.zeros((B, T1)), np.zeros((B, T1, nx))
    rc = lib().orc_calc(mp, nm, _p(refs), _p(dts), B, T, _p(xs), _p(us), _p(cost), _p(xnext))
    assert rc == 0
    return cost, xnext


def calc_diff(models, refs, dts, xs, us):
    mp, nm, nv = _models(models)
    refs, dts, xs, us = _c(refs), _c(dts), _c(xs), _c(us)
    B, T1, nx = xs.shape
    T = T1 - 1
    out = dict(
        cost=np.zeros((B, T1)), xnext=np.zeros((B, T1, nx)), Fx=np.zeros((B, T1, nx, nx)),
        Fu=np.zeros((B, T1, nx, nv)), Lx=np.zeros((B, T1, nx)), Lu=np.zeros((B, T1, nv)),
        Lxx=np.zeros((B, T1, nx, nx)), Lxu=np.zeros((B, T1, nx, nv)), Luu=np.zeros((B, T1, nv, nv)),
    )
    rc = lib().orc_calc_diff(
        mp, nm, _p(refs), _p(dts), B, T, _p(xs), _p(us), *[_p(out[k]) for k in
        ("cost", "xnext", "Fx", "Fu", "Lx", "Lu", "Lxx", "Lxu", "Luu")])
    assert rc == 0
    return out


def rollout(models, refs, dts, x0, us):
    mp, nm, nv = _models(models)
    refs, dts, x0, us = _c(refs), _c(dts), _c(x0), _c(us)
    B, T, _ = us.shape
    xs = np.zeros((B, T + 1, 2 * nv))
    rc = lib().orc_rollout(mp, nm, _p(refs), _p(dts), B, T, _p(x0), _p(us), _p(xs))
    assert rc == 0
    return xs


def integrate(m, x, u, dt):
    x, u = _c(x), _c(u)
    n = x.size // (2 * m.nv)
    out = np.zeros_like(x)
    f = lib().orc_integrate
    f.argtypes = [C.c_void_p, _P, _P, C.c_double, C.c_int, _P]
    rc = f(C.addressof(m), _p(x), _p(u), float(dt), n, _p(out))
    assert rc == 0
    return out


def solve(models, refs, dts, x0, xs_ws, us_ws, max_iter, opts=None, nthreads=0):
    mp, nm, nv = _models(models)
    refs, dts, x0, xs_ws, us_ws = _c(refs), _c(dts), _c(x0), _c(xs_ws), _c(us_ws)
    B, T1, nx = xs_ws.shape
    T = T1 - 1
    if opts is None:
        opts = _abi.default_fddp_opts()
    out = dict(
        xs=np.zeros((B, T1, nx)), us=np.zeros((B, T, nv)), K=np.zeros((B, T, nv, nx)),
        k=np.zeros((B, T, nv)), cost=np.zeros(B), iters=np.zeros(B, dtype=np.int32),
        status=np.zeros(B, dtype=np.int32), stop=np.zeros(B),
    )
    rc = lib().orc_solve(
        mp, nm, _p(refs), _p(dts), B, T, _p(x0), _p(xs_ws), _p(us_ws), int(max_iter), C.byref(opts),
        _p(out["xs"]), _p(out["us"]), _p(out["K"]), _p(out["k"]), _p(out["cost"]),
        out["iters"].ctypes.data_as(_PI), out["status"].ctypes.data_as(_PI), _p(out["stop"]), int(nthreads))
    assert rc == 0
    return out


def mpc_latency(m, stream_refs, dts, x0, xs0, us0, ticks, max_iter, opts=None, node_threads=1):
    """Closed-loop single-problem MPC on the CPU (``orc_mpc_latency``): per-tick solve times in ns, iterations, final x."""
    stream_refs, dts, x0, xs0, us0 = _c(stream_refs), _c(dts), _c(x0), _c(xs0), _c(us0)
    T = len(dts)
    if opts is None:
        opts = _abi.default_fddp_opts()
    ns = np.zeros(ticks, dtype=np.int64)
    iters = np.zeros(ticks, dtype=np.int32)
    xf = np.zeros(2 * m.nv)
    f = lib().orc_mpc_latency
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, _P, C.c_int, _P, C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                  C.c_void_p, _P]
    rc = f(C.addressof(m), _p(stream_refs), stream_refs.shape[0], _p(dts), T, _p(x0), _p(xs0), _p(us0), int(ticks),
           int(max_iter), C.addressof(opts), int(node_threads), ns.ctypes.data, iters.ctypes.data, _p(xf))
    assert rc == 0
    return ns, iters, xf


def solve_sqp(models, refs, dts, x0, xs_ws, us_ws, max_iter, opts=None, nthreads=0):
    """mim_solvers.SolverCSQP (unconstrained) restated: see ``Sqp`` in agx_oracle.cpp."""
    mp, nm, nv = _models(models)
    refs, dts, x0, xs_ws, us_ws = _c(refs), _c(dts), _c(x0), _c(xs_ws), _c(us_ws)
    B, T1, nx = xs_ws.shape
    T = T1 - 1
    if opts is None:
        opts = _abi.default_sqp_opts()
    out = dict(
        xs=np.zeros((B, T1, nx)), us=np.zeros((B, T, nv)), K=np.zeros((B, T, nv, nx)),
        k=np.zeros((B, T, nv)), cost=np.zeros(B), iters=np.zeros(B, dtype=np.int32),
        status=np.zeros(B, dtype=np.int32), stop=np.zeros(B),
    )
    f = lib().orc_solve_sqp
    f.restype = C.c_int
    rc = f(mp, nm, _p(refs), _p(dts), B, T, _p(x0), _p(xs_ws), _p(us_ws), int(max_iter), C.byref(opts),
           _p(out["xs"]), _p(out["us"]), _p(out["K"]), _p(out["k"]), _p(out["cost"]),
           out["iters"].ctypes.data_as(_PI), out["status"].ctypes.data_as(_PI), _p(out["stop"]), int(nthreads))
    assert rc == 0
    return out


def riccati_sigma(m, refs, dts, x0, xs, us, sigma):
    refs, dts, x0, xs, us = _c(refs), _c(dts), _c(x0), _c(xs), _c(us)
    T = us.shape[0]
    nv = m.nv
    K, k, kkt = np.zeros((T, nv, 2 * nv)), np.zeros((T, nv)), np.zeros(2)
    f = lib().orc_riccati_sigma
    f.argtypes = [C.c_void_p, _P, _P, C.c_int, _P, _P, _P, C.c_double, _P, _P, _P]
    rc = f(C.addressof(m), _p(refs), _p(dts), T, _p(x0), _p(xs), _p(us), float(sigma), _p(K), _p(k), _p(kkt))
    assert rc == 0, rc
    return K, k, kkt


def num_threads() -> int:
    return int(lib().orc_num_threads())
