"""Batched shooting problem + FDDP solver on one B200: thin torch-tensor front end of the C ABI.

Replaces, for a batch of ``B`` independent OCPs, the objects the reference builds at
``agimus_controller/agimus_controller/ocp_base_croco.py:36-80`` (``crocoddyl.ShootingProblem`` and the
solver) and the calls it makes on them: ``problem.calc/calcDiff`` (``mpc_debugger_node.py:300-301``),
``problem.rollout`` (``tests/test_warm_start_shift_previous_reference.py:76``), ``solver.solve``
(``ocp_base_croco.py:172``), ``runningModels[0].calc`` (``ocp_base_croco.py:184-189``) and ``pin.rnea``
(``warm_start_reference.py:77-87``).  Every tensor is fp64, C-contiguous, on the handle's CUDA device;
calls are enqueued on torch's current stream and never synchronise.
"""
from __future__ import annotations

import ctypes as C
import typing as T

import numpy as np
import torch

from . import _abi
from ._lib import lib
from .robot_model import RobotTable


def _ptr(t: T.Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def probe_fp64_tflops(device: int = 0, seconds: float = 0.5) -> float:
    """Sustained FP64 FMA throughput of the device (roofline denominator of the fp64 kernels)."""
    out = C.c_double()
    rc = lib().agx_probe_fp64(int(device), float(seconds), C.byref(out), None)
    if rc != 0:
        raise RuntimeError(f"agx_probe_fp64 failed ({rc})")
    return float(out.value)


class BatchedShootingProblem:
    """``B`` shooting problems with ``T`` running nodes sharing one horizon layout."""

    def __init__(self, tables: T.Union[RobotTable, T.Sequence[RobotTable]], dts: T.Sequence[float], B: int,
                 device: T.Union[int, torch.device, str, None] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("agimus_controller_b200 needs a CUDA device (no CPU fallback)")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("agimus_controller_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        if isinstance(tables, RobotTable):
            structs = [tables.to_struct()]
            self.table = tables
        else:
            structs = [t.to_struct() for t in tables]
            self.table = tables[0]
        arr = (_abi.AgxModel * len(structs))(*structs)
        self.dts = np.ascontiguousarray(dts, dtype=np.float64)
        self.B, self.T = int(B), int(len(self.dts))
        self.nv = self.table.nv
        self.nx = 2 * self.nv
        self.ref_size = _abi.ref_size(self.nv)
        self._h = C.c_void_p()
        rc = lib().agx_create(arr, len(structs), self.dts.ctypes.data, self.B, self.T, self.device.index,
                              C.byref(self._h))
        if rc != 0:
            msg = lib().agx_last_error(self._h).decode() if self._h else ""
            if self._h:
                lib().agx_destroy(self._h)
            self._h = None
            raise RuntimeError(f"agx_create failed ({rc}): {msg}")
        self._refs_set = False

    # ------------------------------------------------------------------ helpers
    def close(self) -> None:
        if getattr(self, "_h", None):
            lib().agx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise RuntimeError(f"agx error {rc}: {lib().agx_last_error(self._h).decode()}")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _t(self, x, shape) -> torch.Tensor:
        if isinstance(x, np.ndarray) and not x.flags.writeable:
            x = x.copy()  # broadcast views are read-only; torch wants a writable buffer
        t = torch.as_tensor(x, dtype=torch.float64, device=self.device).contiguous()
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    def _empty(self, *shape, dtype=torch.float64) -> torch.Tensor:
        return torch.empty(shape, dtype=dtype, device=self.device)

    @property
    def launch_count(self) -> int:
        return int(lib().agx_launch_count(self._h))

    def set_timing(self, enable: bool) -> None:
        self._check(lib().agx_set_timing(self._h, 1 if enable else 0))

    def get_timing(self) -> dict:
        """Per-phase device time (ms) and launch counts since the last read (synchronises)."""
        ms = (C.c_double * 5)()
        n = (C.c_longlong * 5)()
        self._check(lib().agx_get_timing(self._h, ms, n))
        names = ("calc_diff", "backward", "rollout_try", "node_cost", "accept_linesearch")
        return {k: dict(ms=float(ms[i]), launches=int(n[i])) for i, k in enumerate(names)}

    # ------------------------------------------------------------------ problem data
    def set_refs(self, refs) -> None:
        """Per-node references and weights, ``[B, T+1, ref_size]`` (``problem.pack_refs``)."""
        r = self._t(refs, (self.B, self.T + 1, self.ref_size))
        self._check(lib().agx_set_refs(self._h, _ptr(r), self._stream()))
        self._refs_set = True

    def set_capsule(self, capsule: int, a0, a1, radius: float) -> None:
        """New end points (parent-joint frame, or world frame for an obstacle) and radius of one collision capsule, for
        every model of the batch (``OCPBaseCroco.update_geometry_placement``, ``ocp_base_croco.py:110-131``)."""
        a0 = (C.c_double * 3)(*[float(v) for v in a0])
        a1 = (C.c_double * 3)(*[float(v) for v in a1])
        self._check(lib().agx_set_capsule(self._h, int(capsule), a0, a1, float(radius), self._stream()))

    def set_refs_window(self, stream_refs: torch.Tensor, start) -> None:
        """Select the horizon window out of a device-resident reference stream ``[n_points, ref_size]`` (shared) or
        ``[B, n_points, ref_size]``; ``start`` is an int (all problems) or an int32 device tensor ``[B]``."""
        r = torch.as_tensor(stream_refs, dtype=torch.float64, device=self.device).contiguous()
        if r.dim() == 2:
            r = r[None]
        if r.shape[0] not in (1, self.B) or r.shape[2] != self.ref_size:
            raise ValueError(f"bad reference stream shape {tuple(r.shape)}")
        if isinstance(start, torch.Tensor):
            st = start.to(device=self.device, dtype=torch.int32).contiguous()
            self._check(lib().agx_set_refs_window(self._h, _ptr(r), r.shape[0], r.shape[1], _ptr(st), 0, self._stream()))
        else:
            self._check(lib().agx_set_refs_window(self._h, _ptr(r), r.shape[0], r.shape[1], None, int(start),
                                                  self._stream()))
        self._refs_set = True

    # ------------------------------------------------------------------ problem.calc / calcDiff / rollout
    def calc(self, xs, us):
        xs = self._t(xs, (self.B, self.T + 1, self.nx))
        us = self._t(us, (self.B, self.T, self.nv))
        cost, xnext = self._empty(self.B, self.T + 1), self._empty(self.B, self.T + 1, self.nx)
        self._check(lib().agx_calc(self._h, _ptr(xs), _ptr(us), _ptr(cost), _ptr(xnext), self._stream()))
        return cost, xnext

    def calc_diff(self, xs, us) -> dict:
        xs = self._t(xs, (self.B, self.T + 1, self.nx))
        us = self._t(us, (self.B, self.T, self.nv))
        B, T1, nx, nv = self.B, self.T + 1, self.nx, self.nv
        out = dict(cost=self._empty(B, T1), xnext=self._empty(B, T1, nx), Fx=self._empty(B, T1, nx, nx),
                   Fu=self._empty(B, T1, nx, nv), Lx=self._empty(B, T1, nx), Lu=self._empty(B, T1, nv),
                   Lxx=self._empty(B, T1, nx, nx), Lxu=self._empty(B, T1, nx, nv), Luu=self._empty(B, T1, nv, nv))
        out["Lu"].zero_()
        out["Luu"].zero_()
        self._check(lib().agx_calc_diff(
            self._h, _ptr(xs), _ptr(us), *[_ptr(out[k]) for k in
                                           ("cost", "xnext", "Fx", "Fu", "Lx", "Lu", "Lxx", "Lxu", "Luu")],
            self._stream()))
        return out

    def rollout(self, x0, us):
        x0 = self._t(x0, (self.B, self.nx))
        us = self._t(us, (self.B, self.T, self.nv))
        xs = self._empty(self.B, self.T + 1, self.nx)
        self._check(lib().agx_rollout(self._h, _ptr(x0), _ptr(us), _ptr(xs), self._stream()))
        return xs

    def integrate(self, x, u, dt: float):
        x = torch.as_tensor(x, dtype=torch.float64, device=self.device).contiguous().reshape(-1, self.nx)
        u = torch.as_tensor(u, dtype=torch.float64, device=self.device).contiguous().reshape(-1, self.nv)
        out = torch.empty_like(x)
        self._check(lib().agx_integrate(self._h, _ptr(x), _ptr(u), float(dt), x.shape[0], _ptr(out), self._stream()))
        return out

    def rnea(self, q, v, a):
        q = torch.as_tensor(q, dtype=torch.float64, device=self.device).contiguous().reshape(-1, self.nv)
        v = torch.as_tensor(v, dtype=torch.float64, device=self.device).contiguous().reshape(-1, self.nv)
        a = torch.as_tensor(a, dtype=torch.float64, device=self.device).contiguous().reshape(-1, self.nv)
        tau = torch.empty_like(q)
        self._check(lib().agx_rnea(self._h, _ptr(q), _ptr(v), _ptr(a), q.shape[0], _ptr(tau), self._stream()))
        return tau

    def cost_terms(self, xs, us) -> dict:
        """Per-cost values and the frame-placement residual of every node (debugger view)."""
        xs = self._t(xs, (self.B, self.T + 1, self.nx))
        us = self._t(us, (self.B, self.T, self.nv))
        out = self._empty(self.B, self.T + 1, _abi.AGX_N_COST_TERMS)
        self._check(lib().agx_cost_terms(self._h, _ptr(xs), _ptr(us), _ptr(out), self._stream()))
        return dict(state_reg=out[..., 0], control_reg=out[..., 1], goal_tracking=out[..., 2], r_pose=out[..., 3:9],
                    collision=out[..., 9:11], collision_distance=out[..., 11:13])

    COST_NAMES = ("state_reg", "control_reg", "goal_tracking", "collision_0", "collision_1")

    def cost_derivatives(self, xs, us) -> dict:
        """Per-cost gradients ``w * Lx``, ``w * Lu`` of every node, unscaled by the time step — what the debugger reads
        off ``runningDatas[i].differential.costs.costs[name]`` (``mpc_debugger_node.py:303-323``):
        ``{"Lx": [B, T+1, 5, nx], "Lu": [B, T+1, 5, nu]}``, costs in the order of ``COST_NAMES``."""
        xs = self._t(xs, (self.B, self.T + 1, self.nx))
        us = self._t(us, (self.B, self.T, self.nv))
        Lx = self._empty(self.B, self.T + 1, _abi.AGX_N_COSTS, self.nx)
        Lu = self._empty(self.B, self.T + 1, _abi.AGX_N_COSTS, self.nv)
        self._check(lib().agx_cost_derivatives(self._h, _ptr(xs), _ptr(us), _ptr(Lx), _ptr(Lu), self._stream()))
        return dict(Lx=Lx, Lu=Lu)

    def shift_warmstart(self, xs, us):
        """Previous solution shifted by the first time step (``WarmStartShiftPreviousSolution.shift``)."""
        xs = self._t(xs, (self.B, self.T + 1, self.nx))
        us = self._t(us, (self.B, self.T, self.nv))
        oxs, ous = torch.empty_like(xs), torch.empty_like(us)
        self._check(lib().agx_shift_warmstart(self._h, _ptr(xs), _ptr(us), _ptr(oxs), _ptr(ous), self._stream()))
        return oxs, ous

    def riccati(self, x0, xs, us, reg: float = 0.0):
        """calc + calcDiff at ``(xs, us)`` and one backward sweep with fixed regularisation ``reg`` -> ``K, k, status``
        (``reg = 1e-6`` = the proximal sigma of the reference's CSQP backward pass)."""
        x0 = self._t(x0, (self.B, self.nx))
        xs = self._t(xs, (self.B, self.T + 1, self.nx))
        us = self._t(us, (self.B, self.T, self.nv))
        K = self._empty(self.B, self.T, self.nv, self.nx)
        k = self._empty(self.B, self.T, self.nv)
        status = self._empty(self.B, dtype=torch.int32)
        self._check(lib().agx_riccati(self._h, _ptr(x0), _ptr(xs), _ptr(us), float(reg), _ptr(K), _ptr(k),
                                      _ptr(status), self._stream()))
        return K, k, status

    # ------------------------------------------------------------------ solver.solve
    def alloc_outputs(self, with_k: bool = True) -> dict:
        B, T, nx, nv = self.B, self.T, self.nx, self.nv
        out = dict(xs=self._empty(B, T + 1, nx), us=self._empty(B, T, nv), K=self._empty(B, T, nv, nx),
                   cost=self._empty(B), iters=self._empty(B, dtype=torch.int32),
                   status=self._empty(B, dtype=torch.int32), stop=self._empty(B))
        if with_k:
            out["k"] = self._empty(B, T, nv)
        return out

    def solve(self, x0, xs_ws, us_ws, max_iter: int, opts: T.Optional[_abi.AgxFddpOpts] = None,
              out: T.Optional[dict] = None) -> dict:
        """FDDP from the warm start ``(xs_ws, us_ws)`` with ``problem.x0 = x0``; stream-ordered, no sync."""
        if not self._refs_set:
            raise RuntimeError("set_refs() must be called before solve()")
        x0 = self._t(x0, (self.B, self.nx))
        xs_ws = self._t(xs_ws, (self.B, self.T + 1, self.nx))
        us_ws = self._t(us_ws, (self.B, self.T, self.nv))
        if out is None:
            out = self.alloc_outputs()
        if opts is None:
            opts = _abi.default_fddp_opts()
        self._check(lib().agx_solve(
            self._h, _ptr(x0), _ptr(xs_ws), _ptr(us_ws), int(max_iter), C.byref(opts), _ptr(out["xs"]),
            _ptr(out["us"]), _ptr(out["K"]), _ptr(out.get("k")), _ptr(out["cost"]), _ptr(out["iters"]),
            _ptr(out["status"]), _ptr(out.get("stop")), self._stream()))
        return out

    def solve_sqp(self, x0, xs_ws, us_ws, max_iter: int, opts: T.Optional[_abi.AgxSqpOpts] = None,
                  out: T.Optional[dict] = None) -> dict:
        """The solver the reference instantiates — ``mim_solvers.SolverCSQP(problem).solve(xs, us, max_iter)``
        (``ocp_base_croco.py:64-75, :172``) with no constraint active: Gauss-Newton SQP, KKT stop at
        ``termination_tolerance``, merit line search; ``out["stop"]`` is the KKT norm, ``out["K"]`` the gains of the
        solver's last backward pass (with its proximal sigma).  Stream-ordered, no sync for budgets up to 32."""
        if not self._refs_set:
            raise RuntimeError("set_refs() must be called before solve_sqp()")
        x0 = self._t(x0, (self.B, self.nx))
        xs_ws = self._t(xs_ws, (self.B, self.T + 1, self.nx))
        us_ws = self._t(us_ws, (self.B, self.T, self.nv))
        if out is None:
            out = self.alloc_outputs()
        if opts is None:
            opts = _abi.default_sqp_opts()
        self._check(lib().agx_solve_sqp(
            self._h, _ptr(x0), _ptr(xs_ws), _ptr(us_ws), int(max_iter), C.byref(opts), _ptr(out["xs"]),
            _ptr(out["us"]), _ptr(out["K"]), _ptr(out.get("k")), _ptr(out["cost"]), _ptr(out["iters"]),
            _ptr(out["status"]), _ptr(out.get("stop")), self._stream()))
        return out


class SolvePipeline:
    """Several independent batches in flight on one GPU: ``n_in_flight`` handles of the same problem layout, each with
    its own stream, served round robin.

    Why: at the benchmark's batch (4096 problems) the per-problem kernels of one solve (Riccati sweep: one warp per
    problem; forward pass: four problems per warp) leave SMs idle in their last wave, and the forward pass fills less
    than one wave.  Kernels of another, independent batch run in those gaps: 379 k -> 434 k solves/s with three batches
    in flight on one B200 (``bench.py``).  Splitting ONE batch into slabs does not help (each slab's kernels are
    latency-bound and take almost as long as the whole batch's).

    Every submitted batch is an ordinary ``BatchedShootingProblem.solve``: same results, bit for bit, as the same
    batch solved alone (``tests/test_gpu_tick_graph.py``).
    """

    class Ticket:
        def __init__(self, out, event, index):
            self.out, self.event, self.index = out, event, index

        def wait(self) -> dict:
            """Blocks the host until this batch's results are complete; returns them."""
            self.event.synchronize()
            return self.out

    def __init__(self, tables, dts, B: int, n_in_flight: int = 3, device=None):
        if n_in_flight < 1:
            raise ValueError("n_in_flight must be at least 1")
        self.problems = [BatchedShootingProblem(tables, dts, B, device=device) for _ in range(n_in_flight)]
        self.device = self.problems[0].device
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(n_in_flight)]
        self._outs = [p.alloc_outputs() for p in self.problems]
        self._next = 0

    @property
    def launch_count(self) -> int:
        return sum(p.launch_count for p in self.problems)

    def set_refs(self, refs) -> None:
        """The same reference rows for every handle (a per-batch set goes through ``submit(refs=...)``)."""
        for p in self.problems:
            p.set_refs(refs)

    def submit(self, x0, xs_ws, us_ws, max_iter: int, opts=None, refs=None, out: T.Optional[dict] = None,
               after_current_stream: bool = True, sqp: bool = False) -> "SolvePipeline.Ticket":
        """Queues one batch on the next handle's stream and never blocks the host.  By default the batch is ordered
        after the work already queued on the caller's current stream (which produced the inputs); a caller whose inputs
        are ready and whose current stream carries work that waits on EARLIER batches (a collective over their results,
        say) passes ``after_current_stream=False`` so that the batches do not serialise through it.  Without ``out`` the
        results land in the handle's own buffers, valid until that handle's next turn (``n_in_flight`` submits later).
        ``sqp=True`` runs ``solve_sqp`` (``opts``: ``AgxSqpOpts``) instead of FDDP."""
        j = self._next
        self._next = (j + 1) % len(self.problems)
        p, s = self.problems[j], self.streams[j]
        if after_current_stream:
            s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            if refs is not None:
                p.set_refs(refs)
            res = (p.solve_sqp if sqp else p.solve)(x0, xs_ws, us_ws, max_iter, opts,
                                                    out=out if out is not None else self._outs[j])
            ev = torch.cuda.Event()
            ev.record(s)
        return SolvePipeline.Ticket(res, ev, j)

    def join(self) -> None:
        """Orders the caller's current stream after everything submitted so far."""
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            cur.wait_stream(s)

    def close(self) -> None:
        for p in self.problems:
            p.close()
