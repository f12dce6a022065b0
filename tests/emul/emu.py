"""numpy front end of the C ABI built for the CPU SIMT emulator (``tests/emul/libagx_emul.so``).

TEST INFRASTRUCTURE ONLY.  The emulator compiles the product's CUDA translation unit
(``agimus_controller_b200/csrc/agx_api.cu``) with g++: CUDA threads become fibers, "device" memory is
host memory.  It lets the CPU test-suite run the very kernels that ship against the oracle; the function
signatures mirror ``oracle/orc.py`` so the two are interchangeable in the tests.
"""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess

import numpy as np

from agimus_controller_b200 import _abi

_HERE = pathlib.Path(__file__).resolve().parent
_LIB = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)
        _LIB = _abi.bind(C.CDLL(str(_HERE / "libagx_emul.so")))
    return _LIB


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data


class Handle:
    def __init__(self, models, dts, B, T):
        if isinstance(models, _abi.AgxModel):
            arr, n = (_abi.AgxModel * 1)(models), 1
        else:
            arr, n = (_abi.AgxModel * len(models))(*models), len(models)
        self.dts = _c(dts)
        self.h = C.c_void_p()
        self.B, self.T = B, T
        rc = lib().agx_create(arr, n, _p(self.dts), B, T, 0, C.byref(self.h))
        if rc != 0:
            msg = lib().agx_last_error(self.h).decode()
            lib().agx_destroy(self.h)
            self.h = None
            raise RuntimeError(f"agx_create failed ({rc}): {msg}")

    def check(self, rc):
        if rc != 0:
            raise RuntimeError(f"agx error {rc}: {lib().agx_last_error(self.h).decode()}")

    def set_capsule(self, capsule, a0, a1, radius):
        a0 = (C.c_double * 3)(*a0)
        a1 = (C.c_double * 3)(*a1)
        self.check(lib().agx_set_capsule(self.h, int(capsule), a0, a1, float(radius), None))

    def set_refs(self, refs):
        self.refs = _c(refs)
        self.check(lib().agx_set_refs(self.h, _p(self.refs), None))

    def __del__(self):
        if getattr(self, "h", None):
            lib().agx_destroy(self.h)
            self.h = None


def _handle(models, refs, dts, B, T):
    h = Handle(models, dts, B, T)
    h.set_refs(refs)
    return h


def calc(models, refs, dts, xs, us):
    xs, us = _c(xs), _c(us)
    B, T1, nx = xs.shape
    h = _handle(models, refs, dts, B, T1 - 1)
    cost, xnext = np.zeros((B, T1)), np.zeros((B, T1, nx))
    h.check(lib().agx_calc(h.h, _p(xs), _p(us), _p(cost), _p(xnext), None))
    return cost, xnext


def calc_with_moved_capsule(models, refs, dts, xs, us, capsule, a0, a1, radius):
    """`calc` after `agx_set_capsule` on a fresh handle."""
    xs, us = _c(xs), _c(us)
    B, T1, nx = xs.shape
    h = _handle(models, refs, dts, B, T1 - 1)
    h.set_capsule(capsule, a0, a1, radius)
    cost, xnext = np.zeros((B, T1)), np.zeros((B, T1, nx))
    h.check(lib().agx_calc(h.h, _p(xs), _p(us), _p(cost), _p(xnext), None))
    return cost, xnext


def calc_diff(models, refs, dts, xs, us):
    xs, us = _c(xs), _c(us)
    B, T1, nx = xs.shape
    nv = nx // 2
    h = _handle(models, refs, dts, B, T1 - 1)
    out = dict(
        cost=np.zeros((B, T1)), xnext=np.zeros((B, T1, nx)), Fx=np.zeros((B, T1, nx, nx)),
        Fu=np.zeros((B, T1, nx, nv)), Lx=np.zeros((B, T1, nx)), Lu=np.zeros((B, T1, nv)),
        Lxx=np.zeros((B, T1, nx, nx)), Lxu=np.zeros((B, T1, nx, nv)), Luu=np.zeros((B, T1, nv, nv)),
    )
    h.check(lib().agx_calc_diff(h.h, _p(xs), _p(us), *[_p(out[k]) for k in
            ("cost", "xnext", "Fx", "Fu", "Lx", "Lu", "Lxx", "Lxu", "Luu")], None))
    return out


def rollout(models, refs, dts, x0, us):
    x0, us = _c(x0), _c(us)
    B, T, nv = us.shape
    h = _handle(models, refs, dts, B, T)
    xs = np.zeros((B, T + 1, 2 * nv))
    h.check(lib().agx_rollout(h.h, _p(x0), _p(us), _p(xs), None))
    return xs


def integrate(m, x, u, dt):
    x, u = _c(x), _c(u)
    n = x.size // (2 * m.nv)
    h = Handle(m, np.array([dt]), 1, 1)
    out = np.zeros_like(x)
    h.check(lib().agx_integrate(h.h, _p(x), _p(u), float(dt), n, _p(out), None))
    return out


def rnea(m, q, v, a):
    q, v, a = _c(q), _c(v), _c(a)
    n = q.size // m.nv
    h = Handle(m, np.array([1.0]), 1, 1)
    tau = np.zeros_like(q)
    h.check(lib().agx_rnea(h.h, _p(q), _p(v), _p(a), n, _p(tau), None))
    return tau


def solve(models, refs, dts, x0, xs_ws, us_ws, max_iter, opts=None):
    x0, xs_ws, us_ws = _c(x0), _c(xs_ws), _c(us_ws)
    B, T1, nx = xs_ws.shape
    T, nv = T1 - 1, nx // 2
    h = _handle(models, refs, dts, B, T)
    if opts is None:
        opts = _abi.default_fddp_opts()
    out = dict(
        xs=np.zeros((B, T1, nx)), us=np.zeros((B, T, nv)), K=np.zeros((B, T, nv, nx)),
        k=np.zeros((B, T, nv)), cost=np.zeros(B), iters=np.zeros(B, dtype=np.int32),
        status=np.zeros(B, dtype=np.int32), stop=np.zeros(B),
    )
    h.check(lib().agx_solve(
        h.h, _p(x0), _p(xs_ws), _p(us_ws), int(max_iter), C.byref(opts), _p(out["xs"]), _p(out["us"]),
        _p(out["K"]), _p(out["k"]), _p(out["cost"]), _p(out["iters"]), _p(out["status"]), _p(out["stop"]), None))
    out["launches"] = lib().agx_launch_count(h.h)
    return out


def solve_sqp(models, refs, dts, x0, xs_ws, us_ws, max_iter, opts=None):
    x0, xs_ws, us_ws = _c(x0), _c(xs_ws), _c(us_ws)
    B, T1, nx = xs_ws.shape
    T, nv = T1 - 1, nx // 2
    h = _handle(models, refs, dts, B, T)
    if opts is None:
        opts = _abi.default_sqp_opts()
    out = dict(
        xs=np.zeros((B, T1, nx)), us=np.zeros((B, T, nv)), K=np.zeros((B, T, nv, nx)),
        k=np.zeros((B, T, nv)), cost=np.zeros(B), iters=np.zeros(B, dtype=np.int32),
        status=np.zeros(B, dtype=np.int32), stop=np.zeros(B),
    )
    h.check(lib().agx_solve_sqp(
        h.h, _p(x0), _p(xs_ws), _p(us_ws), int(max_iter), C.byref(opts), _p(out["xs"]), _p(out["us"]),
        _p(out["K"]), _p(out["k"]), _p(out["cost"]), _p(out["iters"]), _p(out["status"]), _p(out["stop"]), None))
    out["launches"] = lib().agx_launch_count(h.h)
    return out


def riccati(m, refs, dts, x0, xs, us, reg):
    """Single problem, same signature as ``orc.riccati_sigma`` (returns K, k, status)."""
    xs, us, x0 = _c(xs)[None], _c(us)[None], _c(x0)[None]
    T, nv = us.shape[1], us.shape[2]
    h = _handle(m, _c(refs)[None], dts, 1, T)
    K, k = np.zeros((1, T, nv, 2 * nv)), np.zeros((1, T, nv))
    status = np.zeros(1, dtype=np.int32)
    h.check(lib().agx_riccati(h.h, _p(x0), _p(xs), _p(us), float(reg), _p(K), _p(k), _p(status), None))
    return K[0], k[0], int(status[0])


def cost_terms(models, refs, dts, xs, us):
    xs, us = _c(xs), _c(us)
    B, T1, nx = xs.shape
    h = _handle(models, refs, dts, B, T1 - 1)
    out = np.zeros((B, T1, 13))
    h.check(lib().agx_cost_terms(h.h, _p(xs), _p(us), _p(out), None))
    return out


def cost_derivatives(models, refs, dts, xs, us):
    xs, us = _c(xs), _c(us)
    B, T1, nx = xs.shape
    h = _handle(models, refs, dts, B, T1 - 1)
    Lx, Lu = np.zeros((B, T1, _abi.AGX_N_COSTS, nx)), np.zeros((B, T1, _abi.AGX_N_COSTS, nx // 2))
    h.check(lib().agx_cost_derivatives(h.h, _p(xs), _p(us), _p(Lx), _p(Lu), None))
    return Lx, Lu


def shift_warmstart(models, refs, dts, xs, us):
    xs, us = _c(xs), _c(us)
    B, T1, nx = xs.shape
    h = _handle(models, refs, dts, B, T1 - 1)
    oxs, ous = np.zeros_like(xs), np.zeros_like(us)
    h.check(lib().agx_shift_warmstart(h.h, _p(xs), _p(us), _p(oxs), _p(ous), None))
    return oxs, ous


def refs_window_cost(models, dts, stream_refs, start, xs, us):
    """Set the references from a stream window, then problem.calc (exercises agx_set_refs_window)."""
    xs, us = _c(xs), _c(us)
    B, T1, nx = xs.shape
    h = Handle(models, dts, B, T1 - 1)
    sr = _c(stream_refs)
    if sr.ndim == 2:
        sr = sr[None]
    if np.ndim(start) == 0:
        h.check(lib().agx_set_refs_window(h.h, _p(sr), sr.shape[0], sr.shape[1], None, int(start), None))
    else:
        st = np.ascontiguousarray(start, dtype=np.int32)
        h.check(lib().agx_set_refs_window(h.h, _p(sr), sr.shape[0], sr.shape[1], st.ctypes.data, 0, None))
    cost, xnext = np.zeros((B, T1)), np.zeros((B, T1, nx))
    h.check(lib().agx_calc(h.h, _p(xs), _p(us), _p(cost), _p(xnext), None))
    return cost
