"""Device-side warm starts for the batched OCP.

Same behaviour as the reference's warm-start classes, for ``B`` problems at once and without host round trips:

* ``WarmStartReference``             ``agimus_controller/agimus_controller/warm_start_reference.py:33-96`` —
  ``x_init`` = [x0, reference states 1..T], ``u_init`` = ``pin.rnea(q, v, a)`` at [x0-point, reference 1..T-1];
* ``WarmStartShiftPreviousSolution`` ``agimus_controller/agimus_controller/warm_start_shift_previous_solution.py:30-104`` —
  previous solution shifted by the first time step, coarse nodes re-integrated.

Both accept either tensors (batched form) or the reference's ``TrajectoryPoint`` lists (single problem).
"""
from __future__ import annotations

import numpy as np
import torch

from .solver import BatchedShootingProblem


def _points_to_arrays(points):
    q = np.stack([np.asarray(p.robot_configuration, dtype=np.float64) for p in points])
    v = np.stack([np.asarray(p.robot_velocity, dtype=np.float64) for p in points])
    a = np.stack([np.asarray(p.robot_acceleration, dtype=np.float64) for p in points])
    return q, v, a


class WarmStartReference:
    def __init__(self, problem: BatchedShootingProblem) -> None:
        self._p = problem

    def generate_batched(self, x0, ref_q, ref_v, ref_a):
        """``x0 [B, nx]``, references ``[B, T+1, nv]`` -> ``x0, xs_init [B, T+1, nx], us_init [B, T, nv]`` (device)."""
        p = self._p
        dev = p.device
        x0 = torch.as_tensor(x0, dtype=torch.float64, device=dev)
        q, v, a = (torch.as_tensor(t, dtype=torch.float64, device=dev) for t in (ref_q, ref_v, ref_a))
        xs = torch.cat([q, v], dim=-1).contiguous()
        xs[:, 0] = x0
        nv = p.nv
        uq, uv, ua = xs[:, :-1, :nv], xs[:, :-1, nv:], a[:, :-1].clone()
        us = p.rnea(uq.reshape(-1, nv), uv.reshape(-1, nv), ua.reshape(-1, nv)).reshape(p.B, p.T, nv)
        return x0, xs, us

    def generate(self, initial_state, reference_trajectory):
        """Single-problem form with TrajectoryPoint inputs and list outputs, as the reference."""
        assert self._p.B == 1
        q, v, a = _points_to_arrays([initial_state] + list(reference_trajectory[1:]))
        x0 = np.concatenate([q[0], v[0]])
        _, xs, us = self.generate_batched(x0[None], q[None], v[None], a[None])
        return x0, list(xs[0].cpu().numpy()), list(us[0].cpu().numpy())


class WarmStartShiftPreviousSolution:
    def __init__(self, problem: BatchedShootingProblem) -> None:
        self._p = problem
        self._prev = None
        self._fallback = WarmStartReference(problem)

    def update_previous_solution(self, results) -> None:
        """``results``: the batched result dict of ``solve`` (device tensors) or an ``OCPResults``."""
        if isinstance(results, dict):
            self._prev = (results["xs"], results["us"])
        else:
            dev = self._p.device
            xs = torch.as_tensor(np.stack(results.states)[None], dtype=torch.float64, device=dev)
            us = torch.as_tensor(np.stack(results.feed_forward_terms)[None], dtype=torch.float64, device=dev)
            self._prev = (xs, us)

    def generate_batched(self, x0):
        assert self._prev is not None, "update_previous_solution must be called before generate"
        xs, us = self._p.shift_warmstart(*self._prev)
        return torch.as_tensor(x0, dtype=torch.float64, device=self._p.device), xs, us

    def generate(self, initial_state, reference_trajectory):
        if self._prev is None:
            return self._fallback.generate(initial_state, reference_trajectory)
        x0 = np.concatenate([initial_state.robot_configuration, initial_state.robot_velocity])
        _, xs, us = self.generate_batched(x0[None])
        return x0, list(xs[0].cpu().numpy()), list(us[0].cpu().numpy())
