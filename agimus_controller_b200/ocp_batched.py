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

# YAML residual classes (ocp_croco_generic.py:153-550) -> (slot of the reference record, kind)
_RESIDUALS = {
    "ResidualModelState": ("state", "state"),
    "ResidualModelControl": ("control", "control"),
    "ResidualModelFramePlacement": ("pose", "placement"),
    "ResidualModelFramePlacementStatic": ("pose", "placement"),
    "ResidualModelFrameTranslation": ("pose", "translation"),
    "ResidualModelFrameTranslationStatic": ("pose", "translation"),
    "ResidualModelFrameRotation": ("pose", "rotation"),
    "ResidualModelFrameRotationStatic": ("pose", "rotation"),
    "ResidualModelVisualServoing": ("pose", "visual_servoing"),
    "ResidualDistanceCollision": ("collision", "collision"),
    "ResidualDistanceCollision2": ("collision", "collision"),  # same residual, evaluated on the state's shared geometry
}
# not on the device path: their Hessians need the Lxu / Lqv blocks the compact cost record does not carry
_REFUSED = {
    "ResidualModelControlGrav": "its Gauss-Newton Hessian couples q and u (Lxu != 0)",
    "ResidualModelFrameVelocity": "its residual depends on q and v (Lqv != 0)",
    "ResidualModelFrameVelocityStatic": "its residual depends on q and v (Lqv != 0)",
}
_NR = {"state": None, "control": None, "placement": 6, "translation": 3, "rotation": 3, "visual_servoing": 6}


def _static_weights(act: T.Optional[dict], nr: int) -> np.ndarray:
    """Activation weights as ``ActivationModelWeightedQuad.build`` resolves them (ocp_croco_generic.py:106-114):
    no activation / ``weights: null`` -> ones, a scalar -> that scalar, a list -> itself."""
    w = None if act is None else act.get("weights")
    if w is None:
        return np.ones(nr)
    try:
        return float(w) * np.ones(nr)
    except (ValueError, TypeError):
        w = np.asarray(w, dtype=np.float64)
        assert w.size == nr, f"activation weights have {w.size} entries, the residual {nr}"
        return w


def flatten_cost_stack(model_def: dict, terminal: bool, nv: int = 7) -> dict:
    """One cost stack (``running_model`` / ``terminal_model`` subtree of the OCP definition YAML,
    ``ocp_goal_reaching.yaml:1-63``) flattened into descriptors of the reference record's slots:

    ``{"state": d | None, "control": d | None, "pose": [d, ...], "collisions": [d, ...], "weights": {slot: w}, "names":
    {slot: name}}`` with ``d = dict(name, kind, weight (CostModelSum weight, 0 when inactive), update, ref (the YAML's
    static reference or None), w (the YAML's static activation weights), publish, frame, ...)``.

    ``update: false`` keeps the YAML's reference and weights for every tick, ``update: true`` takes them from the
    trajectory point (``DifferentialActionModelFreeFwdDynamics.update``, ``ocp_croco_generic.py:712-724``).  Anything
    the device path does not cover raises instead of being dropped."""
    if model_def.get("class") != "IntegratedActionModelEuler":
        raise NotImplementedError(f"integrator {model_def.get('class')} is not supported on the device path")
    diff = model_def["differential"]
    if diff.get("class") != "DifferentialActionModelFreeFwdDynamics":
        raise NotImplementedError(f"differential model {diff.get('class')} is not supported on the device path")
    if diff.get("constraints"):
        raise NotImplementedError("constraints need CSQP's ADMM loop for inequality constraints: not on the device path")
    out = {"state": None, "control": None, "pose": [], "collisions": []}
    for item in diff.get("costs", []):
        cost = item["cost"]
        if cost.get("class") != "CostModelResidual":
            raise NotImplementedError(f"cost class {cost.get('class')}")
        act = cost.get("activation")
        res = cost["residual"]
        rcls = res.get("class")
        if rcls in _REFUSED:
            raise NotImplementedError(f"residual {rcls} is not supported on the device path: {_REFUSED[rcls]}")
        if rcls not in _RESIDUALS:
            raise NotImplementedError(f"residual {rcls} is not supported on the device path")
        slot, kind = _RESIDUALS[rcls]
        weight = float(item.get("weight", 1.0)) if item.get("active", True) else 0.0
        d = dict(name=item["name"], kind=kind, weight=weight, update=bool(item.get("update", False)),
                 publish=bool(item.get("publish_residual", False)))
        if slot == "collision":
            # colmpc distance residual under the squared-exponential activation
            # (ocp_croco_generic.py:119-147, :499-550; ocp_traj_tracking_collision_avoidance.yaml:36-46)
            acls = (act or {}).get("class")
            quad_exp = acls == "ActivationModelQuadExp" or (acls == "ActivationModelExp" and int(act.get("exponent", 1)) == 2)
            if not quad_exp:
                raise NotImplementedError(f"activation {acls} on a collision residual is not supported on the device path")
            pair = tuple(res["collision_pair"]) if "collision_pair" in res else int(res.get("collision_pair_id", 0))
            d.update(pair=pair, alpha=float(act.get("alpha", 1.0)))
            out["collisions"].append(d)
            continue
        if act is not None and act.get("class") != "ActivationModelWeightedQuad":
            raise NotImplementedError(f"activation {act.get('class')} is not supported on the device path")
        if slot == "control" and terminal:
            continue  # the terminal node has no control
        if slot in ("state", "control"):
            if out[slot] is not None:
                raise NotImplementedError(f"two costs on the {slot} residual")
            nr = 2 * nv if slot == "state" else nv
            ref = res.get("xref" if slot == "state" else "uref")
            d.update(ref=None if ref is None else np.asarray(ref, dtype=np.float64), w=_static_weights(act, nr))
            out[slot] = d
            continue
        # task-frame costs share the pose slot
        static_frame = rcls.endswith("Static")
        d.update(static_frame=static_frame, frame=res.get("frame_id") if static_frame else res.get("id"),
                 ref=None if res.get("pref") is None else np.asarray(res["pref"], dtype=np.float64),
                 w=_static_weights(act, _NR[kind]))
        if kind == "visual_servoing":
            d.update(frame=res["robot_frame"], static_frame=True, input_key=res["robot_frame"] + "_vs",
                     transforms_key=(res["world_frame"], res["object_frame"]))
        if any(o["kind"] == kind for o in out["pose"]):
            raise NotImplementedError(f"two {kind} costs on the task frame")
        out["pose"].append(d)
    kinds = {d["kind"] for d in out["pose"]}
    if len(kinds) > 1 and kinds != {"translation", "rotation"}:
        raise NotImplementedError(f"task-frame costs {sorted(kinds)} cannot share the pose record")
    # summary kept for callers that only need the CostModelSum weights / names per slot
    out["weights"] = {"state": out["state"]["weight"] if out["state"] else 0.0,
                      "control": out["control"]["weight"] if out["control"] else 0.0,
                      "pose": max([d["weight"] for d in out["pose"]], default=0.0)}
    out["names"] = {k: out[k]["name"] for k in ("state", "control") if out[k]}
    if out["pose"]:
        out["names"]["pose"] = out["pose"][0]["name"]
    return out


def pose_mode_of(running: dict, terminal: dict) -> int:
    """``agx_model.pose_mode`` the two stacks ask for: world-frame translation when a FrameTranslation cost is there."""
    kinds = [{d["kind"] for d in st["pose"]} for st in (running, terminal)]
    tr = ["translation" in k for k in kinds]
    if any(tr) and any(("placement" in k or "visual_servoing" in k) for k in kinds):
        raise NotImplementedError("FrameTranslation and FramePlacement costs in one OCP: the pose record has one form")
    return _abi.AGX_POSE_TRANSLATION_WORLD if any(tr) else _abi.AGX_POSE_PLACEMENT


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


def _pose_of(pt, key):
    poses = pt.end_effector_poses
    return poses[key] if key is not None else next(iter(poses.values()))


def node_references(table: RobotTable, stack: dict, wp, transforms: T.Optional[dict] = None) -> dict:
    """What ``DifferentialActionModelFreeFwdDynamics.update`` (``ocp_croco_generic.py:712-724``) leaves in the
    residuals / activations of ONE node: ``{cost name: (reference, activation weights)}`` plus the packed slots
    ``xref, wx, uref, wu, Rref, pref, wpose, wcol`` (CostModelSum weights folded in)."""
    nv = table.nv
    pt, wt = wp.point, wp.weights
    out = {"by_name": {}}
    xref, wx = np.zeros(2 * nv), np.zeros(2 * nv)
    d = stack["state"]
    if d is not None:
        if d["update"]:
            ref, w = np.asarray(pt.robot_state, dtype=np.float64), np.asarray(wt.w_robot_state, dtype=np.float64)
        else:
            ref = d["ref"] if d["ref"] is not None else np.zeros(2 * nv)  # ResidualModelState(state): xref = state.zero()
            w = d["w"]
        out["by_name"][d["name"]] = (ref, w)
        xref, wx = ref, d["weight"] * w
    uref, wu = np.zeros(nv), np.zeros(nv)
    d = stack["control"]
    if d is not None:
        if d["update"]:
            ref, w = np.asarray(pt.robot_effort, dtype=np.float64), np.asarray(wt.w_robot_effort, dtype=np.float64)
        else:
            ref, w = (d["ref"] if d["ref"] is not None else np.zeros(nv)), d["w"]
        out["by_name"][d["name"]] = (ref, w)
        uref, wu = ref, d["weight"] * w
    Rref, pref, wpose = np.eye(3), np.zeros(3), np.zeros(6)
    for d in stack["pose"]:
        kind = d["kind"]
        sl = {"placement": slice(0, 6), "visual_servoing": slice(0, 6), "translation": slice(0, 3),
              "rotation": slice(3, 6)}[kind]
        if d["update"]:
            poses = pt.end_effector_poses
            assert len(poses) == 1, (
                f"{kind} residual requires exactly one end-effector pose, current is {poses}.")
            key = d.get("input_key") if kind == "visual_servoing" else (d["frame"] if d["static_frame"] else None)
            if key is not None:
                assert key in poses, f"end_effector_poses should contain the key {key}"
            ee_name = key if key is not None else next(iter(poses))
            frame_name = d["frame"] if d["static_frame"] else ee_name
            if frame_name != table.frame_name:
                raise NotImplementedError(f"the device tables were built for frame '{table.frame_name}', got '{frame_name}'")
            pose = poses[ee_name]
            R, p = np.asarray(pose.rotation, dtype=np.float64), np.asarray(pose.translation, dtype=np.float64)
            w6 = np.asarray(wt.w_end_effector_poses[ee_name], dtype=np.float64)
            if kind == "visual_servoing":
                # reference = wMo_vision * oMf_target when the transform is known (ocp_croco_generic.py:455-475)
                wMo = (transforms or {}).get(d["transforms_key"])
                assert not np.any(w6 != 0) or wMo is not None, (
                    f"Weights are not all zeros and no transform for {d['transforms_key']}")
                if wMo is not None:
                    Rw, pw = np.asarray(wMo.rotation, dtype=np.float64), np.asarray(wMo.translation, dtype=np.float64)
                    R, p = Rw @ R, pw + Rw @ p
            w = w6[sl] if kind in ("translation", "rotation") else w6
        else:
            if d["frame"] is not None and isinstance(d["frame"], str) and d["frame"] != table.frame_name:
                raise NotImplementedError(f"the device tables were built for frame '{table.frame_name}', got '{d['frame']}'")
            if d["ref"] is None:
                R, p = np.eye(3), np.zeros(3)
            else:
                from .robot_model import xyzquat_to_se3

                R, p = xyzquat_to_se3(d["ref"]) if len(d["ref"]) >= 7 else (np.eye(3), np.asarray(d["ref"][:3]))
            w = d["w"]
        if kind in ("placement", "visual_servoing"):
            Rref, pref = R, p
            out["by_name"][d["name"]] = ((R, p), w)
        elif kind == "translation":
            pref = p
            out["by_name"][d["name"]] = (p, w)
        else:
            Rref = R
            out["by_name"][d["name"]] = (R, w)
        wpose[sl] = d["weight"] * w
    wcol = np.zeros(_abi.AGX_MAX_COLLISION_PAIRS)
    for col in stack.get("collisions", []):
        # the collision activation has no weight vector: update() sets the scalar CostModelSum weight
        # (ocp_croco_generic.py:714-719)
        wcol[col["slot"]] += float(wt.w_collision_avoidance) if col["update"] else col["weight"]
    out.update(xref=xref, wx=wx, uref=uref, wu=wu, Rref=Rref, pref=pref, wpose=wpose, wcol=wcol)
    return out


def _pack_node_row(row: np.ndarray, nv: int, r: dict, terminal: bool) -> None:
    """One node's slots into its reference record (layout of ``include/agx.h``)."""
    nx = 2 * nv
    row[0:nx] = r["xref"]
    row[nx:2 * nx] = r["wx"]
    o = 2 * nx
    row[o:o + nv] = r["uref"]
    row[o + nv:o + 2 * nv] = 0.0 if terminal else r["wu"]   # the terminal node has no control cost
    o += 2 * nv
    row[o:o + 9] = np.asarray(r["Rref"], dtype=np.float64).reshape(9)
    row[o + 9:o + 12] = r["pref"]
    row[o + 12:o + 18] = r["wpose"]
    row[o + 18:o + 18 + len(r["wcol"])] = r["wcol"]


_POSE_SLICES = {"placement": slice(0, 6), "translation": slice(0, 3), "rotation": slice(3, 6)}


def _running_rows_vectorised(table: RobotTable, stack: dict, horizon: list, rows: np.ndarray) -> bool:
    """The running nodes' records, one numpy operation per field over the whole horizon instead of a Python pass per
    node (the per-node form costs 1.5 ms for 20 nodes: five times the solve).  Covers the stacks whose costs all read
    the trajectory point (``update: true`` state / control / Frame* residuals, collisions); returns False without
    touching ``rows`` for anything else (static references, visual servoing), which then takes the per-node path.
    Same values, bit for bit (``tests/test_host_logic.py``)."""
    nv = table.nv
    nx = 2 * nv
    for d in (stack["state"], stack["control"]):
        if d is not None and not d["update"]:
            return False
    for d in stack["pose"]:
        if not d["update"] or d["kind"] not in _POSE_SLICES:
            return False
    n = len(horizon)
    pts = [wp.point for wp in horizon]
    wts = [wp.weights for wp in horizon]
    out = np.zeros((n, rows.shape[1]))
    d = stack["state"]
    if d is not None:
        out[:, 0:nv] = [p.robot_configuration for p in pts]
        out[:, nv:nx] = [p.robot_velocity for p in pts]
        w = np.empty((n, nx))
        w[:, 0:nv] = [w_.w_robot_configuration for w_ in wts]
        w[:, nv:nx] = [w_.w_robot_velocity for w_ in wts]
        out[:, nx:2 * nx] = d["weight"] * w
    o = 2 * nx
    d = stack["control"]
    if d is not None:
        out[:, o:o + nv] = [p.robot_effort for p in pts]
        out[:, o + nv:o + 2 * nv] = d["weight"] * np.asarray([w_.w_robot_effort for w_ in wts], dtype=np.float64)
    o += 2 * nv
    out[:, o:o + 9] = np.eye(3).reshape(9)
    for d in stack["pose"]:
        kind = d["kind"]
        sl = _POSE_SLICES[kind]
        key = d["frame"] if d["static_frame"] else None
        names = []
        for p in pts:
            poses = p.end_effector_poses
            assert len(poses) == 1, (
                f"{kind} residual requires exactly one end-effector pose, current is {poses}.")
            if key is not None:
                assert key in poses, f"end_effector_poses should contain the key {key}"
            ee_name = key if key is not None else next(iter(poses))
            frame_name = d["frame"] if d["static_frame"] else ee_name
            if frame_name != table.frame_name:
                raise NotImplementedError(f"the device tables were built for frame '{table.frame_name}', got '{frame_name}'")
            names.append(ee_name)
        if kind in ("placement", "rotation"):
            out[:, o:o + 9] = np.asarray([p.end_effector_poses[k].rotation for p, k in zip(pts, names)],
                                         dtype=np.float64).reshape(n, 9)
        if kind in ("placement", "translation"):
            out[:, o + 9:o + 12] = [p.end_effector_poses[k].translation for p, k in zip(pts, names)]
        w6 = np.asarray([w_.w_end_effector_poses[k] for w_, k in zip(wts, names)], dtype=np.float64)
        out[:, o + 12 + sl.start:o + 12 + sl.stop] = d["weight"] * w6[:, sl]
    for col in stack.get("collisions", []):
        c = o + 18 + col["slot"]
        if col["update"]:
            out[:, c] += [float(w_.w_collision_avoidance) for w_ in wts]
        else:
            out[:, c] += col["weight"]
    rows[:n] = out
    return True


def build_reference_rows(table: RobotTable, running: dict, terminal: dict, horizon: list,
                         transforms: T.Optional[dict] = None, vectorised: bool = True) -> np.ndarray:
    """``[T+1, ref_size]`` reference records of one horizon: what ``set_reference_weighted_trajectory``
    (``ocp_croco_generic.py:855-892``) writes into the Crocoddyl residuals / activations of every node, with the
    CostModelSum weight folded into the activation weights.  The last point feeds the terminal model."""
    T1 = len(horizon)
    nv = table.nv
    rows = np.zeros((T1, _abi.ref_size(nv)))
    first = 0
    if vectorised and T1 > 1 and _running_rows_vectorised(table, running, horizon[:-1], rows):
        first = T1 - 1
    for t in range(first, T1):
        last = t == T1 - 1
        _pack_node_row(rows[t], nv, node_references(table, terminal if last else running, horizon[t], transforms), last)
    return rows


class OCPBatchedFDDP(OCPBase):
    def __init__(self, robot_table, params: OCPParamsBaseCroco,
                 yaml_file: T.Union[str, dict, T.IO], batch_size: int = 1, device=None,
                 fddp_opts: T.Optional[_abi.AgxFddpOpts] = None, solver: str = "fddp", frame: T.Optional[str] = None) -> None:
        """``robot_table``: a ``RobotTable``, or the reference's ``RobotModels`` (anything with ``robot_model`` /
        ``collision_model`` / ``armature``: flattened with ``RobotTable.from_robot_models``; ``frame`` then names the task
        frame).  ``solver = "fddp"`` (the solver BASELINE.json's north_star names) or ``"csqp"``: the solver the
        reference instantiates (``mim_solvers.SolverCSQP``, ``ocp_base_croco.py:64-75``) in its unconstrained form,
        configured from ``params.termination_tolerance`` as the reference does."""
        if solver not in ("fddp", "csqp"):
            raise ValueError(f"solver must be 'fddp' or 'csqp', got {solver!r}")
        self._solver = solver
        if getattr(params, "use_filter_line_search", False):
            raise NotImplementedError("use_filter_line_search: the device solvers use the merit / expected-improvement "
                                      "line searches only")
        if isinstance(yaml_file, dict):
            data = yaml_file
        elif hasattr(yaml_file, "read"):
            data = yaml.safe_load(yaml_file)
        else:
            with open(yaml_file, "r") as f:
                data = yaml.safe_load(f)
        if not isinstance(robot_table, RobotTable):
            robot_table = RobotTable.from_robot_models(robot_table, frame)
        elif frame is not None:
            robot_table = robot_table.with_frame(frame)
        nv = robot_table.nv
        self._running = flatten_cost_stack(data["running_model"], terminal=False, nv=nv)
        self._terminal = flatten_cost_stack(data["terminal_model"], terminal=True, nv=nv)
        robot_table = resolve_collision_pairs(robot_table, self._running, self._terminal)
        robot_table = robot_table.with_pose_mode(pose_mode_of(self._running, self._terminal))
        if not robot_table.frame_name:
            # the frame a static task-frame cost names, else any frame: a stack without pose costs never reads it
            named = [d["frame"] for st in (self._running, self._terminal) for d in st["pose"]
                     if isinstance(d.get("frame"), str)]
            robot_table = robot_table.with_frame(named[0] if named else next(iter(robot_table.frames)))
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
        # single-problem form (numpy / list inputs): pinned host staging for the three inputs and for the results, so that
        # a tick costs three asynchronous H2D copies, the solve, seven asynchronous D2H copies and ONE synchronisation
        self._pin = None
        dev_ = getattr(self._problem, "device", None)
        if self._B == 1 and isinstance(dev_, torch.device) and dev_.type == "cuda":
            nx_, T_ = 2 * nv, self._problem.T
            dev = self._problem.device
            host = lambda *shape, dtype=torch.float64: torch.empty(shape, dtype=dtype).pin_memory()  # noqa: E731

            def carve(buf, shapes):
                """Views of consecutive pieces of one flat buffer (every piece starts on an even index: 16 bytes)."""
                out_, o = {}, 0
                for k, shp in shapes.items():
                    n = int(np.prod(shp))
                    out_[k] = buf[o:o + n].view(*shp)
                    o += n + (n & 1)
                return out_

            size = lambda shapes: sum(int(np.prod(v)) + (int(np.prod(v)) & 1) for v in shapes.values())  # noqa: E731
            # inputs, results and integer results each live in ONE device buffer with ONE pinned mirror: a tick is one
            # H2D copy, the solve, two D2H copies and one synchronisation
            in_shapes = dict(x0=(1, nx_), xs=(1, T_ + 1, nx_), us=(1, T_, nv))
            f_shapes = dict(xs=(1, T_ + 1, nx_), us=(1, T_, nv), K=(1, T_, nv, nx_), k=(1, T_, nv), cost=(1,), stop=(1,))
            i_shapes = dict(iters=(1,), status=(1,))
            h_in, d_in = host(size(in_shapes)), torch.empty(size(in_shapes), dtype=torch.float64, device=dev)
            h_f, d_f = host(size(f_shapes)), torch.empty(size(f_shapes), dtype=torch.float64, device=dev)
            h_i = host(size(i_shapes), dtype=torch.int32)
            d_i = torch.empty(size(i_shapes), dtype=torch.int32, device=dev)
            self._out = {**carve(d_f, f_shapes), **carve(d_i, i_shapes)}
            self._pin = dict(
                h_in=h_in, d_in=d_in, np_in={k: v.numpy() for k, v in carve(h_in, in_shapes).items()},
                d_in_views=carve(d_in, in_shapes), h_f=h_f, d_f=d_f, h_i=h_i, d_i=d_i,
                h_out={**carve(h_f, f_shapes), **carve(h_i, i_shapes)},
                h_refs=host(1, T_ + 1, self._problem.ref_size))
            self._pin["d_refs"] = torch.empty_like(self._pin["h_refs"], device=dev)
        # transforms requested by the OCP and provided externally (BuildData.transforms, ocp_croco_generic.py:84-88)
        self._transforms: dict = {d["transforms_key"]: None for st in (self._running, self._terminal)
                                  for d in st["pose"] if d["kind"] == "visual_servoing"}
        self._node0_refs: dict = {}
        self._last_refs: T.Optional[np.ndarray] = None
        self.init_debug_data_attributes()

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

    @property
    def input_transforms(self) -> dict:
        """``OCPCrocoGeneric.input_transforms`` (``ocp_croco_generic.py:894-897``): ``{(parent, child): SE3 | None}``."""
        return self._transforms

    # ------------------------------------------------------------------ debug data (ocp_croco_generic.py:814-853)
    def init_debug_data_attributes(self) -> None:
        for d in self._running_costs():
            if d["update"] and d["kind"] != "collision":
                self._debug_data.references.append((d["name"], None))
            if d["publish"]:
                self._debug_data.residuals.append((d["name"], None))

    def _running_costs(self) -> list:
        r = self._running
        return [d for d in (r["control"], r["state"]) if d is not None] + list(r["pose"]) + list(r["collisions"])

    def _fill_references_and_residuals(self, out: dict) -> None:
        """References of the first running node and residual predictions of the running nodes, as
        ``OCPCrocoGeneric.fill_debug_data`` reads them off the Crocoddyl data (``ocp_croco_generic.py:827-853``)."""
        from .robot_model import se3_to_xyzquat

        dd = self._debug_data
        for i, (name, _) in enumerate(dd.references):
            ref = self._node0_refs.get(name, (None, None))[0]
            if isinstance(ref, tuple):  # an SE3 reference is published as XYZQUAT
                ref = se3_to_xyzquat(*ref)
            dd.references[i] = (name, None if ref is None else np.array(ref, copy=True))
        if not dd.residuals or self._last_refs is None:
            return
        T_ = self.n_controls
        xs, us = out["xs"][:1], out["us"][:1]
        terms = self._problem.cost_terms(out["xs"], out["us"])
        xs_h, us_h = xs[0].cpu().numpy(), us[0].cpu().numpy()
        r6 = terms["r_pose"][0].cpu().numpy()
        dist = terms["collision_distance"][0].cpu().numpy()
        nx, nv = self._problem.nx, self._problem.nv
        refs = self._last_refs
        by_name = {d["name"]: d for d in self._running_costs()}
        for i, (name, _) in enumerate(dd.residuals):
            d = by_name[name]
            if d["kind"] == "state":
                r = xs_h[:T_] - refs[:T_, :nx]
            elif d["kind"] == "control":
                r = us_h[:T_] - refs[:T_, 2 * nx: 2 * nx + nv]
            elif d["kind"] in ("placement", "visual_servoing"):
                r = r6[:T_]
            elif d["kind"] == "translation":
                r = r6[:T_, :3]
            elif d["kind"] == "rotation":
                r = r6[:T_, 3:]
            else:
                r = dist[:T_, d["slot"]: d["slot"] + 1]
            dd.residuals[i] = (name, np.array(r, copy=True))

    # ------------------------------------------------------------------ references
    def reference_table(self, reference_weighted_trajectory: list) -> np.ndarray:
        """``[T+1, ref_size]`` rows from a list of WeightedTrajectoryPoint (one MPC horizon)."""
        assert len(reference_weighted_trajectory) == self.n_controls + 1
        return build_reference_rows(self._table, self._running, self._terminal, reference_weighted_trajectory,
                                    self._transforms)

    def set_reference_weighted_trajectory(self, reference_weighted_trajectory: list) -> None:
        """One horizon for every problem of the batch (list of points) or one horizon per problem (list of lists)."""
        if reference_weighted_trajectory and isinstance(reference_weighted_trajectory[0], (list, tuple)):
            assert len(reference_weighted_trajectory) == self._B
            refs = np.stack([self.reference_table(h) for h in reference_weighted_trajectory])
            first = reference_weighted_trajectory[0]
        else:
            rows = self.reference_table(reference_weighted_trajectory)
            refs = np.broadcast_to(rows, (self._B,) + rows.shape)
            first = reference_weighted_trajectory
        if self._ocp_params.use_debug_data and (self._debug_data.references or self._debug_data.residuals):
            self._node0_refs = node_references(self._table, self._running, first[0], self._transforms)["by_name"]
            self._last_refs = np.array(refs[0], copy=True)
        if self._pin is not None:
            # single problem: through the pinned staging (the previous tick's copy has completed: solve synchronises)
            self._pin["h_refs"].numpy()[...] = refs
            self._pin["d_refs"].copy_(self._pin["h_refs"], non_blocking=True)
            self._problem.set_refs(self._pin["d_refs"])
        else:
            self._problem.set_refs(np.ascontiguousarray(refs))

    def set_reference_table(self, refs) -> None:
        """Device-resident form: a ``[B, T+1, ref_size]`` tensor built by the caller (no per-point host loop)."""
        self._problem.set_refs(refs)

    # ------------------------------------------------------------------ solve
    def solve(self, x0, x_warmstart, u_warmstart, use_iteration_limits_and_timeout: bool = True) -> None:
        max_iters = self._ocp_params.solver_iters if use_iteration_limits_and_timeout else 1000
        # max_solve_time: passed on only when the user set it, and lifted for the unlimited first solve
        # (ocp_base_croco.py:160-171); the deadline runs on the device clock (include/agx.h)
        mst = getattr(self._ocp_params, "max_solve_time", None)
        timeout = float(mst) if (mst is not None and use_iteration_limits_and_timeout and np.isfinite(mst)) else 0.0
        self._opts.max_solve_time = timeout
        self._sqp_opts.max_solve_time = timeout
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
        pin = self._pin
        if pin is None:
            # a problem object without a CUDA device (the CPU SIMT emulator of the tests): plain copies
            out = run(np.asarray(x0, dtype=np.float64).reshape(1, nx),
                      np.asarray(x_warmstart, dtype=np.float64).reshape(1, T_ + 1, nx),
                      np.asarray(u_warmstart, dtype=np.float64).reshape(1, T_, nv), max_iters)
            pin = dict(h_out={k: out[k].cpu() for k in ("xs", "us", "K", "cost", "iters", "status", "stop")})
            return self._pack_single_result(out, pin["h_out"])
        pin["np_in"]["x0"][...] = np.asarray(x0, dtype=np.float64).reshape(1, nx)
        pin["np_in"]["xs"][...] = np.asarray(x_warmstart, dtype=np.float64).reshape(1, T_ + 1, nx)
        pin["np_in"]["us"][...] = np.asarray(u_warmstart, dtype=np.float64).reshape(1, T_, nv)
        pin["d_in"].copy_(pin["h_in"], non_blocking=True)
        dv = pin["d_in_views"]
        out = run(dv["x0"], dv["xs"], dv["us"], max_iters)   # writes into self._out: views of d_f / d_i
        self._results_batched = out
        pin["h_f"].copy_(pin["d_f"], non_blocking=True)
        pin["h_i"].copy_(pin["d_i"], non_blocking=True)
        torch.cuda.current_stream(self._problem.device).synchronize()   # the one synchronisation of the tick
        self._pack_single_result(out, pin["h_out"])

    def _pack_single_result(self, out: dict, ho: dict) -> None:
        """Host copies of one problem's results -> ``OCPResults`` / ``OCPDebugData`` (``ocp_base_croco.py:134-140,
        :173-177``)."""
        self._results_batched = out
        xs_h, us_h, K_h = ho["xs"][0].numpy().copy(), ho["us"][0].numpy().copy(), ho["K"][0].numpy().copy()
        ocp_results = OCPResults(states=list(xs_h), ricatti_gains=list(K_h), feed_forward_terms=list(us_h))
        if self._ocp_params.use_debug_data:
            self._debug_data.problem_solved = bool(int(ho["status"][0]) == _abi.AGX_STATUS_CONVERGED)
            self._debug_data.result = ocp_results
            self._debug_data.kkt_norm = float(ho["stop"][0])
            self._debug_data.nb_iter = int(ho["iters"][0])
            self._debug_data.nb_qp_iter = 0  # no constraint is active: the QP is solved by one Riccati sweep
            self._fill_references_and_residuals(out)
        self._ocp_results = ocp_results

    def update_geometry_placement(self, geometry_name: str, placement) -> None:
        """Updates placement of the obstacles (``OCPBaseCroco.update_geometry_placement``, ``ocp_base_croco.py:110-131``;
        called by the controller for every obstacle pose it receives, ``agimus_controller.py:406``).  ``placement`` is
        the pose of the capsule in the frame of its parent (the world for an obstacle): the capsule keeps its length
        and radius, its axis is the placement's z axis, its centre the placement's translation."""
        if geometry_name not in self._table.capsules:
            raise RuntimeError(f"Unknown geometry name '{geometry_name}' in collision model!")
        names = self._table.device_capsules()
        _, a0, a1, radius = self._table.capsules[geometry_name]
        half = 0.5 * float(np.linalg.norm(np.asarray(a1) - np.asarray(a0)))
        R = np.asarray(placement.rotation, dtype=np.float64)
        p = np.asarray(placement.translation, dtype=np.float64)
        n0, n1 = p - half * R[:, 2], p + half * R[:, 2]
        if geometry_name in names:  # geometries no collision pair uses are not on the device
            self._problem.set_capsule(names.index(geometry_name), n0, n1, radius)
        par = self._table.capsules[geometry_name][0]
        self._table.capsules[geometry_name] = (par, n0, n1, radius)

    def integrate(self, state, control):
        if isinstance(state, torch.Tensor):
            return self._problem.integrate(state, control, self._ocp_params.timesteps[0])
        # runningModels[0].calc: the first running node's step (ocp_base_croco.py:184-189)
        return self._problem.integrate(np.asarray(state, dtype=np.float64), np.asarray(control, dtype=np.float64),
                                       self._ocp_params.timesteps[0])[0].cpu().numpy()

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
