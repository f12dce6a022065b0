"""``BatchedShootingProblem`` look-alike on the CPU SIMT emulator (TEST INFRASTRUCTURE ONLY).

Same method names and result shapes as ``agimus_controller_b200.solver.BatchedShootingProblem`` (the subset
``OCPBatchedFDDP`` uses), but every call goes to ``tests/emul/libagx_emul.so`` — the product's CUDA translation unit
compiled with g++ — on host arrays wrapped as torch CPU tensors.  It lets the CPU test-suite drive the OCP class (and
the reference's unmodified ``MPC.run``) without a GPU.  Tests monkeypatch it in; nothing under ``agimus_controller_b200/``
ever imports it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from agimus_controller_b200 import _abi
from emul import emu


def _np(x):
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=np.float64)


class EmuShootingProblem:
    def __init__(self, tables, dts, B, device=None):
        self.table = tables if not isinstance(tables, (list, tuple)) else tables[0]
        structs = self.table.to_struct() if not isinstance(tables, (list, tuple)) else [t.to_struct() for t in tables]
        self.dts = np.ascontiguousarray(dts, dtype=np.float64)
        self.B, self.T = int(B), len(self.dts)
        self.nv, self.nx = self.table.nv, 2 * self.table.nv
        self.ref_size = _abi.ref_size(self.nv)
        self.device = torch.device("cpu")
        self._h = emu.Handle(structs, self.dts, self.B, self.T)
        self._refs_set = False

    def _chk(self, rc):
        self._h.check(rc)

    def set_refs(self, refs):
        self._h.set_refs(_np(refs).reshape(self.B, self.T + 1, self.ref_size))
        self._refs_set = True

    def set_capsule(self, capsule, a0, a1, radius):
        self._h.set_capsule(capsule, a0, a1, radius)

    def alloc_outputs(self, with_k=True):
        B, T, nx, nv = self.B, self.T, self.nx, self.nv
        out = dict(xs=torch.zeros(B, T + 1, nx, dtype=torch.float64), us=torch.zeros(B, T, nv, dtype=torch.float64),
                   K=torch.zeros(B, T, nv, nx, dtype=torch.float64), cost=torch.zeros(B, dtype=torch.float64),
                   iters=torch.zeros(B, dtype=torch.int32), status=torch.zeros(B, dtype=torch.int32),
                   stop=torch.zeros(B, dtype=torch.float64))
        if with_k:
            out["k"] = torch.zeros(B, T, nv, dtype=torch.float64)
        return out

    def _solve(self, fn, x0, xs, us, max_iter, opts, out):
        assert self._refs_set
        x0 = _np(x0).reshape(self.B, self.nx)
        xs = _np(xs).reshape(self.B, self.T + 1, self.nx)
        us = _np(us).reshape(self.B, self.T, self.nv)
        if out is None:
            out = self.alloc_outputs()
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
        self._chk(fn(self._h.h, emu._p(x0), emu._p(xs), emu._p(us), int(max_iter), C.byref(opts), p(out["xs"]), p(out["us"]),
                     p(out["K"]), p(out.get("k")), p(out["cost"]), p(out["iters"]), p(out["status"]), p(out.get("stop")),
                     None))
        return out

    def solve(self, x0, xs, us, max_iter, opts=None, out=None):
        return self._solve(emu.lib().agx_solve, x0, xs, us, max_iter, opts or _abi.default_fddp_opts(), out)

    def solve_sqp(self, x0, xs, us, max_iter, opts=None, out=None):
        return self._solve(emu.lib().agx_solve_sqp, x0, xs, us, max_iter, opts or _abi.default_sqp_opts(), out)

    def integrate(self, x, u, dt):
        x, u = _np(x).reshape(-1, self.nx), _np(u).reshape(-1, self.nv)
        out = np.zeros_like(x)
        self._chk(emu.lib().agx_integrate(self._h.h, emu._p(x), emu._p(u), float(dt), x.shape[0], emu._p(out), None))
        return torch.from_numpy(out)

    def rnea(self, q, v, a):
        q, v, a = (_np(t).reshape(-1, self.nv) for t in (q, v, a))
        out = np.zeros_like(q)
        self._chk(emu.lib().agx_rnea(self._h.h, emu._p(q), emu._p(v), emu._p(a), q.shape[0], emu._p(out), None))
        return torch.from_numpy(out)

    def shift_warmstart(self, xs, us):
        xs, us = _np(xs).reshape(self.B, self.T + 1, self.nx), _np(us).reshape(self.B, self.T, self.nv)
        oxs, ous = np.zeros_like(xs), np.zeros_like(us)
        self._chk(emu.lib().agx_shift_warmstart(self._h.h, emu._p(xs), emu._p(us), emu._p(oxs), emu._p(ous), None))
        return torch.from_numpy(oxs), torch.from_numpy(ous)

    def cost_terms(self, xs, us):
        xs, us = _np(xs).reshape(self.B, self.T + 1, self.nx), _np(us).reshape(self.B, self.T, self.nv)
        o = np.zeros((self.B, self.T + 1, _abi.AGX_N_COST_TERMS))
        self._chk(emu.lib().agx_cost_terms(self._h.h, emu._p(xs), emu._p(us), emu._p(o), None))
        o = torch.from_numpy(o)
        return dict(state_reg=o[..., 0], control_reg=o[..., 1], goal_tracking=o[..., 2], r_pose=o[..., 3:9],
                    collision=o[..., 9:11], collision_distance=o[..., 11:13])
