"""``OCPBatchedFDDP`` — the reference's generic OCP (``OCPCrocoGeneric``) re-hosted on the CUDA solve path.

Same constructor inputs and the same ``OCPBase`` contract as
``agimus_controller/agimus_controller/ocp/ocp_croco_generic.py:764-897`` on top of
``agimus_controller/agimus_controller/ocp_base_croco.py:15-215``:

* the YAML cost stack (``ocp/ocp_goal_reaching.yaml``) is *flattened* into cost-slot weights instead of being
  turned into Crocoddyl objects (``DifferentialActionModelFreeFwdDynamics.build``, ``:687-711``);
* ``set_reference_weighted_trajectory`` fills the ``[B, T+1, ref_size]`` reference table that replaces the
  per-node ``residual.reference`` / ``activation.weights`` setters (``:855-892``, ``:158-210``);
* ``solve`` hands the warm start to ``agx_solve`` (FDDP on the device) where the reference calls
  ``solver.solve(xs, us, max_iters)`` (``ocp_base_croco.py:172``) and packs ``OCPResults`` the same way
  (``:173-177``).

Batched use: pass ``batch_size = B`` and give ``solve`` torch tensors ``x0 [B, nx]``, ``xs [B, T+1, nx]``,
``us [B, T, nu]``; results stay on the device in ``ocp_results_batched``.  With ``batch_size = 1`` and numpy
inputs the class is a drop-in behind an unmodified ``MPC.run``.
"""
from __future__ import annotations

import typing as T

import numpy as np
import torch
import yaml

from . import _abi
from .ocp_interface import OCPBase, OCPDebugData, OCPParamsBaseCroco, OCPResults
from .problem import pack_refs
from .robot_model import RobotTable
from .solver import BatchedShootingProblem

_SUPPORTED_RESIDUALS = ("ResidualModelState", "ResidualModelControl", "ResidualModelFramePlacement",
                        "ResidualDistanceCollision")


def flatten_cost_stack(model_def: dict, terminal: bool) -> dict:
    """``{slot: CostModelSum weight}`` for the three residual slots the kernels implement.

    ``model_def`` is the ``running_model`` / ``terminal_model`` subtree of the OCP definition YAML
    (``ocp_goal_reaching.yaml:1-63``).  Anything the device path does not cover raises instead of being dropped.
    """
    if model_def.get("class") != "IntegratedActionModelEuler":
        raise NotImplementedError(f"integrator {model_def.get('class')} is not supported on the device path")
    diff = model_def["differential"]
    if diff.get("class") != "DifferentialActionModelFreeFwdDynamics":
        raise NotImplementedError(f"differential model {diff.get('class')} is not supported on the device path")
    if diff.get("constraints"):
        raise NotImplementedError("constraints need the CSQP solver mode (SURVEY.md 8f N1): not on the FDDP device path")
    slots = {"state": 0.0, "control": 0.0, "pose": 0.0}
    names = {}
    collisions = []
    for item in diff.get("costs", []):
        cost = item["cost"]
        if cost.get("class") != "CostModelResidual":
            raise NotImplementedError(f"cost class {cost.get('class')}")
        act = cost.get("activation")
        rcls = cost["residual"].get("class")
        if rcls not in _SUPPORTED_RESIDUALS:
            raise NotImplementedError(f"residual {rcls} is not supported on the device path")
        if rcls == "ResidualDistanceCollision":
            # colmpc distance residual under the squared-exponential activation
            # (ocp_croco_generic.py:119-147, :499-535; ocp_traj_tracking_collision_avoidance.yaml:36-46)
            acls = (act or {}).get("class")
            quad_exp = acls == "ActivationModelQuadExp" or (acls == "ActivationModelExp" and int(act.get("exponent", 1)) == 2)
            if not quad_exp:
                raise NotImplementedError(f"activation {acls} on a collision residual is not supported on the device path")
            res = cost["residual"]
            pair = tuple(res["collision_pair"]) if "collision_pair" in res else int(res.get("collision_pair_id", 0))
            collisions.append(dict(name=item["name"], pair=pair, alpha=float(act.get("alpha", 1.0)),
                                   weight=float(item.get("weight", 1.0)) if item.get("active", True) else 0.0,
                                   update=bool(item.get("update", False))))
            continue
        if act is not None and act.get("class") != "ActivationModelWeightedQuad":
            raise NotImplementedError(f"activation {act.get('class')} is not supported on the device path")
        slot = {"ResidualModelState": "state", "ResidualModelControl": "control",
                "ResidualModelFramePlacement": "pose"}[rcls]
        if slot == "control" and terminal:
            continue  # the terminal node has no control
        if names.get(slot):
            raise NotImplementedError(f"two costs on the {slot} residual")
        names[slot] = item["name"]
        slots[slot] = float(item.get("weight", 1.0)) if item.get("active", True) else 0.0
    return {"weights": slots, "names": names, "collisions": collisions}


def resolve_collision_pairs(table: RobotTable, running: dict, terminal: dict) -> RobotTable:
    """Table whose collision pairs / alpha are the ones the cost stacks name (``_collision_pair_id``,
    ``ocp_croco_generic.py:504-521``: a pair is added to the geometry model when a residual asks for it); each
    stack's collision entries get the ``slot`` of their pair in the reference record."""
    pairs: list = []
    alphas = set()
    for stack in (running, terminal):
        for col in stack.get("collisions", []):
            pair = col["pair"]
            if isinstance(pair, int):
                if not 0 <= pair < len(table.collision_pairs):
                    raise ValueError(f"collision_pair_id {pair}: the model has {len(table.collision_pairs)} pairs")
                pair = tuple(table.collision_pairs[pair])
            for name in pair:
                if name not in table.capsules:
                    raise ValueError(f"Geometry object '{name}' not found.")
            if pair not in pairs:
                pairs.append(pair)
            col["slot"] = pairs.index(pair)
            alphas.add(col["alpha"])
    if not pairs:
        return table
    if len(pairs) > _abi.AGX_MAX_COLLISION_PAIRS:
        raise NotImplementedError(f"{len(pairs)} collision pairs: the device records hold {_abi.AGX_MAX_COLLISION_PAIRS}")
    if len(alphas) != 1:
        raise NotImplementedError("collision costs with different activation alphas are not supported on the device path")
    return table.with_capsules({n: c for n, c in table.capsules.items()}, pairs, alphas.pop())


def build_reference_rows(table: RobotTable, running: dict, terminal: dict, horizon: list) -> np.ndarray:
    """``[T+1, ref_size]`` reference records of one horizon: what ``DifferentialActionModelFreeFwdDynamics.update``
    (``ocp_croco_generic.py:712-724``) writes into the Crocoddyl residuals / activations of every node, with the
    CostModelSum weight folded into the activation weights.  The last point feeds the terminal model."""
    T1 = len(horizon)
    nv = table.nv
    rows = np.zeros((T1, _abi.ref_size(nv)))
    for t, wp in enumerate(horizon):
        stack = terminal if t == T1 - 1 else running
        w = stack["weights"]
        pt, wt = wp.point, wp.weights
        xref, wx = np.zeros(2 * nv), np.zeros(2 * nv)
        if w["state"] != 0.0:
            xref = np.asarray(pt.robot_state, dtype=np.float64)
            wx = w["state"] * np.asarray(wt.w_robot_state, dtype=np.float64)
        uref, wu = np.zeros(nv), np.zeros(nv)
        if w["control"] != 0.0:
            uref = np.asarray(pt.robot_effort, dtype=np.float64)
            wu = w["control"] * np.asarray(wt.w_robot_effort, dtype=np.float64)
        Rref, pref, wpose = np.eye(3), np.zeros(3), np.zeros(6)
        if w["pose"] != 0.0:
            assert len(pt.end_effector_poses) == 1, (
                "ResidualModelFramePlacement requires exactly one end-effector pose, current is "
                f"{pt.end_effector_poses}.")
            ee_name, ee_pose = next(iter(pt.end_effector_poses.items()))
            if ee_name != table.frame_name:
                raise NotImplementedError(
                    f"the device tables were built for frame '{table.frame_name}', got '{ee_name}'")
            Rref = np.asarray(ee_pose.rotation, dtype=np.float64)
            pref = np.asarray(ee_pose.translation, dtype=np.float64)
            wpose = w["pose"] * np.asarray(wt.w_end_effector_poses[ee_name], dtype=np.float64)
        wcol = np.zeros(_abi.AGX_MAX_COLLISION_PAIRS)
        for col in stack.get("collisions", []):
            # the collision activation has no weight vector: update() sets the scalar CostModelSum weight
            # (ocp_croco_generic.py:714-719)
            wcol[col["slot"]] += float(wt.w_collision_avoidance) if col["update"] else col["weight"]
        rows[t] = pack_refs(nv, 0, 1, xref, wx, uref, wu, Rref, pref, wpose, wcol=wcol)[0, 0]
        if t == T1 - 1:
            rows[t, 5 * nv: 6 * nv] = 0.0  # the terminal node has no control cost
        else:
            rows[t, 5 * nv: 6 * nv] = wu   # pack_refs treats its last node as terminal
    return rows


class OCPBatchedFDDP(OCPBase):
    def __init__(self, robot_table: RobotTable, params: OCPParamsBaseCroco,
                 yaml_file: T.Union[str, dict, T.IO], batch_size: int = 1, device=None,
                 fddp_opts: T.Optional[_abi.AgxFddpOpts] = None, solver: str = "fddp") -> None:
        """``solver = "fddp"`` (the solver BASELINE.json's north_star names) or ``"csqp"``: the solver the reference
        instantiates (``mim_solvers.SolverCSQP``, ``ocp_base_croco.py:64-75``) in its unconstrained form, configured
        from ``params.termination_tolerance`` as the reference does."""
        if solver not in ("fddp", "csqp"):
            raise ValueError(f"solver must be 'fddp' or 'csqp', got {solver!r}")
        self._solver = solver
        if isinstance(yaml_file, dict):
            data = yaml_file
        elif hasattr(yaml_file, "read"):
            data = yaml.safe_load(yaml_file)
        else:
            with open(yaml_file, "r") as f:
                data = yaml.safe_load(f)
        self._running = flatten_cost_stack(data["running_model"], terminal=False)
        self._terminal = flatten_cost_stack(data["terminal_model"], terminal=True)
        robot_table = resolve_collision_pairs(robot_table, self._running, self._terminal)
        self._table = robot_table
        self._ocp_params = params
        self._B = int(batch_size)
        self._problem = BatchedShootingProblem(robot_table, params.timesteps, self._B, device=device)
        self._opts = fddp_opts if fddp_opts is not None else _abi.default_fddp_opts()
        self._sqp_opts = _abi.default_sqp_opts(getattr(params, "termination_tolerance", 1e-3))
        if self._B == 1:
            # a single MPC tick waits for its result: stop queueing iterations as soon as the problem has finished
            if fddp_opts is None:
                self._opts.eager_exit = 1
            self._sqp_opts.eager_exit = 1
        self._ocp_results: T.Optional[OCPResults] = None
        self._results_batched: T.Optional[dict] = None
        self._debug_data = OCPDebugData()
        self._out = self._problem.alloc_outputs()

    # ------------------------------------------------------------------ OCPBase properties
    @property
    def n_controls(self) -> int:
        return self._ocp_params.n_controls

    @property
    def dt(self) -> float:
        return self._ocp_params.dt

    @property
    def batch_size(self) -> int:
        return self._B

    @property
    def problem(self) -> BatchedShootingProblem:
        """The device-side shooting problem (``calc`` / ``calc_diff`` / ``rollout``), as ``OCPBaseCroco.problem``."""
        return self._problem

    # ------------------------------------------------------------------ references
    def reference_table(self, reference_weighted_trajectory: list) -> np.ndarray:
        """``[T+1, ref_size]`` rows from a list of WeightedTrajectoryPoint (one MPC horizon)."""
        assert len(reference_weighted_trajectory) == self.n_controls + 1
        return build_reference_rows(self._table, self._running, self._terminal, reference_weighted_trajectory)

    def set_reference_weighted_trajectory(self, reference_weighted_trajectory: list) -> None:
        """One horizon for every problem of the batch (list of points) or one horizon per problem (list of lists)."""
        if reference_weighted_trajectory and isinstance(reference_weighted_trajectory[0], (list, tuple)):
            assert len(reference_weighted_trajectory) == self._B
            refs = np.stack([self.reference_table(h) for h in reference_weighted_trajectory])
        else:
            rows = self.reference_table(reference_weighted_trajectory)
            refs = np.broadcast_to(rows, (self._B,) + rows.shape)
        self._problem.set_refs(np.ascontiguousarray(refs))

    def set_reference_table(self, refs) -> None:
        """Device-resident form: a ``[B, T+1, ref_size]`` tensor built by the caller (no per-point host loop)."""
        self._problem.set_refs(refs)

    # ------------------------------------------------------------------ solve
    def solve(self, x0, x_warmstart, u_warmstart, use_iteration_limits_and_timeout: bool = True) -> None:
        max_iters = self._ocp_params.solver_iters if use_iteration_limits_and_timeout else 1000
        batched = isinstance(x0, torch.Tensor) and x0.dim() == 2
        run = ((lambda *a: self._problem.solve_sqp(*a, self._sqp_opts, out=self._out)) if self._solver == "csqp"
               else (lambda *a: self._problem.solve(*a, self._opts, out=self._out)))
        if batched:
            out = run(x0, x_warmstart, u_warmstart, max_iters)
            self._results_batched = out
            self._ocp_results = None
            return
        assert self._B == 1, "numpy / list inputs are the single-problem form; pass torch tensors for a batch"
        nx, nv, T_ = self._problem.nx, self._problem.nv, self.n_controls
        xs = np.asarray(x_warmstart, dtype=np.float64).reshape(1, T_ + 1, nx)
        us = np.asarray(u_warmstart, dtype=np.float64).reshape(1, T_, nv)
        out = run(np.asarray(x0, dtype=np.float64).reshape(1, nx), xs, us, max_iters)
        self._results_batched = out
        xs_h, us_h, K_h = out["xs"][0].cpu().numpy(), out["us"][0].cpu().numpy(), out["K"][0].cpu().numpy()
        ocp_results = OCPResults(states=list(xs_h), ricatti_gains=list(K_h), feed_forward_terms=list(us_h))
        if self._ocp_params.use_debug_data:
            self._debug_data.problem_solved = bool(int(out["status"][0]) == _abi.AGX_STATUS_CONVERGED)
            self._debug_data.result = ocp_results
            self._debug_data.kkt_norm = float(out["stop"][0])
            self._debug_data.nb_iter = int(out["iters"][0])
            self._debug_data.nb_qp_iter = 0
        self._ocp_results = ocp_results

    def update_geometry_placement(self, geometry_name: str, placement) -> None:
        """Updates placement of the obstacles (``OCPBaseCroco.update_geometry_placement``, ``ocp_base_croco.py:110-131``;
        called by the controller for every obstacle pose it receives, ``agimus_controller.py:406``).  ``placement`` is
        the pose of the capsule in the frame of its parent (the world for an obstacle): the capsule keeps its length
        and radius, its axis is the placement's z axis, its centre the placement's translation."""
        names = list(self._table.capsules)
        if geometry_name not in names:
            raise RuntimeError(f"Unknown geometry name '{geometry_name}' in collision model!")
        _, a0, a1, radius = self._table.capsules[geometry_name]
        half = 0.5 * float(np.linalg.norm(np.asarray(a1) - np.asarray(a0)))
        R = np.asarray(placement.rotation, dtype=np.float64)
        p = np.asarray(placement.translation, dtype=np.float64)
        n0, n1 = p - half * R[:, 2], p + half * R[:, 2]
        self._problem.set_capsule(names.index(geometry_name), n0, n1, radius)
        par = self._table.capsules[geometry_name][0]
        self._table.capsules[geometry_name] = (par, n0, n1, radius)

    def integrate(self, state, control):
        if isinstance(state, torch.Tensor):
            return self._problem.integrate(state, control, self.dt)
        return self._problem.integrate(np.asarray(state, dtype=np.float64), np.asarray(control, dtype=np.float64),
                                       self.dt)[0].cpu().numpy()

    # ------------------------------------------------------------------ results
    @property
    def ocp_results(self) -> OCPResults:
        return self._ocp_results

    @ocp_results.setter
    def ocp_results(self, value: OCPResults) -> None:
        self._ocp_results = value

    @property
    def ocp_results_batched(self) -> T.Optional[dict]:
        """Device tensors ``xs, us, K, k, cost, iters, status, stop`` of the last solve (stream-ordered)."""
        return self._results_batched

    @property
    def debug_data(self) -> OCPDebugData:
        return self._debug_data

    @debug_data.setter
    def debug_data(self, value: OCPDebugData) -> None:
        self._debug_data = value
