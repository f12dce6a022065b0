"""Kinematic-tree tables — the flattened counterpart of ``RobotModels.robot_model``.

Reference: ``agimus_controller/agimus_controller/factory/robot_model.py:88-351`` builds a Pinocchio
model from a URDF, locks joints with ``buildReducedModel`` (``:231-259``) and exposes ``robot_model``
and ``armature``.  Neither Pinocchio nor a URDF parser dependency is needed on the solve path: the
kernels only consume the numeric table built here (``agx_model`` in ``include/agx.h``).

``RobotTable.from_links`` does the work of ``buildReducedModel`` for fixed / locked joints: bodies
behind a locked joint are merged into their moving ancestor (mass, centre of mass, inertia with the
parallel-axis term).  ``panda_table`` instantiates it with the Franka Panda description used by every
reference test (``agimus_controller/tests/test_ocp_croco_base.py:109-135``; numbers in SURVEY.md
Appendix A, validated against the reference's golden file by ``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import dataclasses
import typing as T

import numpy as np

from . import _abi

# Angles exactly as written in the public URDF (truncated pi/2, pi/4).
_HALF_PI = 1.57079632679
_QUARTER_PI = 0.785398163397


def rpy_to_matrix(r: float, p: float, y: float) -> np.ndarray:
    """URDF fixed-axis roll/pitch/yaw -> rotation matrix ``Rz(y) Ry(p) Rx(r)``."""
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return rz @ ry @ rx


def _sym(i6: T.Sequence[float]) -> np.ndarray:
    xx, xy, xz, yy, yz, zz = i6
    return np.array([[xx, xy, xz], [xy, yy, yz], [xz, yz, zz]], dtype=np.float64)


def quat_to_matrix(q: T.Sequence[float]) -> np.ndarray:
    """Unit quaternion ``(x, y, z, w)`` (Pinocchio / ROS order) -> rotation matrix."""
    x, y, z, w = (float(v) for v in q)
    n = np.sqrt(x * x + y * y + z * z + w * w)
    x, y, z, w = x / n, y / n, z / n, w / n
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def matrix_to_quat(R: np.ndarray) -> np.ndarray:
    """Rotation matrix -> unit quaternion ``(x, y, z, w)`` with ``w >= 0``."""
    R = np.asarray(R, dtype=np.float64)
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        s_ = 2.0 * np.sqrt(1.0 + tr)
        q = np.array([(R[2, 1] - R[1, 2]) / s_, (R[0, 2] - R[2, 0]) / s_, (R[1, 0] - R[0, 1]) / s_, 0.25 * s_])
    else:
        i = int(np.argmax([R[0, 0], R[1, 1], R[2, 2]]))
        j, k = (i + 1) % 3, (i + 2) % 3
        s_ = 2.0 * np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k])
        q = np.zeros(4)
        q[i] = 0.25 * s_
        q[j] = (R[j, i] + R[i, j]) / s_
        q[k] = (R[k, i] + R[i, k]) / s_
        q[3] = (R[k, j] - R[j, k]) / s_
    return q if q[3] >= 0 else -q


def xyzquat_to_se3(v: T.Sequence[float]) -> tuple[np.ndarray, np.ndarray]:
    """``pinocchio.XYZQUATToSE3``: ``[x y z qx qy qz qw]`` -> ``(R, p)``."""
    v = np.asarray(v, dtype=np.float64)
    return quat_to_matrix(v[3:7]), v[:3].copy()


def se3_to_xyzquat(R: np.ndarray, p: np.ndarray) -> np.ndarray:
    """``pinocchio.SE3ToXYZQUAT``."""
    return np.concatenate([np.asarray(p, dtype=np.float64), matrix_to_quat(R)])


# Pinocchio joint short names -> (joint type, axis in the joint frame); *Unaligned joints carry their own axis
_PIN_JOINTS = {
    "JointModelRX": (0, (1.0, 0.0, 0.0)), "JointModelRY": (0, (0.0, 1.0, 0.0)), "JointModelRZ": (0, (0.0, 0.0, 1.0)),
    "JointModelPX": (1, (1.0, 0.0, 0.0)), "JointModelPY": (1, (0.0, 1.0, 0.0)), "JointModelPZ": (1, (0.0, 0.0, 1.0)),
    "JointModelRevoluteUnaligned": (0, None), "JointModelPrismaticUnaligned": (1, None),
}


def _rp(se3) -> tuple[np.ndarray, np.ndarray]:
    """(rotation, translation) of a pinocchio.SE3-like object."""
    return np.asarray(se3.rotation, dtype=np.float64), np.asarray(se3.translation, dtype=np.float64).reshape(3)


@dataclasses.dataclass
class Link:
    """One URDF link + the joint that attaches it to ``parent`` (``None`` = world)."""

    name: str
    parent: T.Optional[str]
    joint_name: str
    joint_type: str  # "revolute" | "prismatic" | "fixed"
    xyz: T.Sequence[float]
    rpy: T.Sequence[float]
    axis: T.Sequence[float] = (0.0, 0.0, 1.0)
    mass: float = 0.0
    com: T.Sequence[float] = (0.0, 0.0, 0.0)
    inertia: T.Sequence[float] = (0.0, 0.0, 0.0, 0.0, 0.0, 0.0)  # xx xy xz yy yz zz about the COM


@dataclasses.dataclass
class RobotTable:
    """Numeric kinematic tree (what the device kernels consume)."""

    joint_names: list[str]
    parent: np.ndarray  # [nv] int, -1 = world
    jtype: np.ndarray  # [nv] int
    axis: np.ndarray  # [nv,3]
    placement_R: np.ndarray  # [nv,3,3]
    placement_p: np.ndarray  # [nv,3]
    mass: np.ndarray  # [nv]
    com: np.ndarray  # [nv,3]
    inertia: np.ndarray  # [nv,3,3] about the COM
    armature: np.ndarray  # [nv]
    gravity: np.ndarray  # [3]
    frames: dict[str, tuple[int, np.ndarray, np.ndarray]]  # name -> (parent joint, R, p)
    frame_name: str = ""
    # collision geometry: name -> (parent joint or -1 = world, a0, a1, radius), segment endpoints in the parent frame
    capsules: dict = dataclasses.field(default_factory=dict)
    collision_pairs: list = dataclasses.field(default_factory=list)  # [(capsule name, capsule name)], at most two
    collision_alpha: float = 1e-4  # ActivationModelQuadExp alpha (ocp_traj_tracking_collision_avoidance.yaml:44)
    # form of the task-frame residual: 0 = log6 placement (ResidualModelFramePlacement / FrameRotation),
    # 1 = world-frame translation for the linear part (ResidualModelFrameTranslation); include/agx.h
    pose_mode: int = 0

    @property
    def nv(self) -> int:
        return len(self.joint_names)

    @property
    def nq(self) -> int:
        return self.nv

    @property
    def nx(self) -> int:
        return 2 * self.nv

    def with_frame(self, frame_name: str) -> "RobotTable":
        assert frame_name in self.frames, f"Frame '{frame_name}' does not exist!"
        return dataclasses.replace(self, frame_name=frame_name)

    def with_pose_mode(self, pose_mode: int) -> "RobotTable":
        assert pose_mode in (_abi.AGX_POSE_PLACEMENT, _abi.AGX_POSE_TRANSLATION_WORLD)
        return dataclasses.replace(self, pose_mode=int(pose_mode))

    def with_armature(self, armature) -> "RobotTable":
        a = np.broadcast_to(np.asarray(armature, dtype=np.float64), (self.nv,)).copy()
        return dataclasses.replace(self, armature=a)

    def with_capsules(self, capsules: dict, pairs: list, alpha: float = 1e-4) -> "RobotTable":
        """Attach collision capsules ``{name: (parent joint name or index or None, a0, a1, radius)}`` and the pairs whose
        distance feeds ``ResidualDistanceCollision`` (``factory/robot_model.py:261-330`` builds the reference's capsules
        from the URDF cylinders; the solve path only needs this flat table)."""
        caps = {}
        for name, (parent, a0, a1, radius) in capsules.items():
            if parent is None:
                pi = -1
            elif isinstance(parent, str):
                pi = self.joint_names.index(parent)
            else:
                pi = int(parent)
            caps[name] = (pi, np.asarray(a0, dtype=np.float64), np.asarray(a1, dtype=np.float64), float(radius))
        assert len(pairs) <= _abi.AGX_MAX_COLLISION_PAIRS
        for a, b in pairs:
            assert a in caps and b in caps, f"Geometry object '{a if a not in caps else b}' not found."
        return dataclasses.replace(self, capsules=caps, collision_pairs=list(pairs), collision_alpha=float(alpha))

    def perturbed(self, link: int, param: int, delta: float) -> "RobotTable":
        """One inertial parameter of body ``link`` shifted by ``delta``: ``param`` 0-5 = inertia
        xx xy xz yy yz zz, 6-8 = COM x y z, 9 = mass
        (agimus_controller_examples/.../model_sensibility/evaluate_model_sensibility.py:9-49)."""
        t = dataclasses.replace(
            self, mass=self.mass.copy(), com=self.com.copy(), inertia=self.inertia.copy()
        )
        if param < 6:
            r, c = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)][param]
            t.inertia[link, r, c] += delta
            if r != c:
                t.inertia[link, c, r] += delta
        elif param < 9:
            t.com[link, param - 6] += delta
        else:
            t.mass[link] += delta
        return t

    def frame_placement(self, q) -> tuple[np.ndarray, np.ndarray]:
        """World placement ``(R, p)`` of the task frame at configuration ``q`` (host-side forward kinematics,
        used to build references; the solve path itself never calls it)."""
        q = np.asarray(q, dtype=np.float64)
        Rw = [None] * self.nv
        pw = [None] * self.nv
        for i in range(self.nv):
            a = self.axis[i]
            if self.jtype[i] == _abi.AGX_JOINT_REVOLUTE:
                K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
                Rj = np.eye(3) + np.sin(q[i]) * K + (1 - np.cos(q[i])) * (K @ K)
                pj = np.zeros(3)
            else:
                Rj, pj = np.eye(3), a * q[i]
            Rl = self.placement_R[i] @ Rj
            pl = self.placement_p[i] + self.placement_R[i] @ pj
            par = int(self.parent[i])
            if par < 0:
                Rw[i], pw[i] = Rl, pl
            else:
                Rw[i], pw[i] = Rw[par] @ Rl, pw[par] + Rw[par] @ pl
        fpar, fR, fp = self.frames[self.frame_name]
        return Rw[fpar] @ np.asarray(fR), pw[fpar] + Rw[fpar] @ np.asarray(fp)

    def to_struct(self) -> _abi.AgxModel:
        assert self.nv <= _abi.AGX_MAX_NV
        assert self.frame_name, "select the task frame with with_frame() first"
        m = _abi.AgxModel()
        m.nv = self.nv
        fpar, fR, fp = self.frames[self.frame_name]
        m.frame_parent = int(fpar)
        for i in range(self.nv):
            m.parent[i] = int(self.parent[i])
            m.jtype[i] = int(self.jtype[i])
            I = self.inertia[i]
            i6 = [I[0, 0], I[0, 1], I[0, 2], I[1, 1], I[1, 2], I[2, 2]]
            for k in range(3):
                m.axis[i][k] = float(self.axis[i, k])
                m.placement_p[i][k] = float(self.placement_p[i, k])
                m.com[i][k] = float(self.com[i, k])
            for k in range(9):
                m.placement_R[i][k] = float(self.placement_R[i].reshape(9)[k])
            for k in range(6):
                m.inertia[i][k] = float(i6[k])
            m.mass[i] = float(self.mass[i])
            m.armature[i] = float(self.armature[i])
        for k in range(3):
            m.gravity[k] = float(self.gravity[k])
            m.frame_p[k] = float(fp[k])
        for k in range(9):
            m.frame_R[k] = float(np.asarray(fR).reshape(9)[k])
        names = self.device_capsules()
        m.n_capsules = len(names)
        for c, name in enumerate(names):
            pi, a0, a1, radius = self.capsules[name]
            m.cap_parent[c] = int(pi)
            m.cap_radius[c] = float(radius)
            for k in range(3):
                m.cap_a0[c][k] = float(a0[k])
                m.cap_a1[c][k] = float(a1[k])
        m.n_pairs = len(self.collision_pairs)
        for k, (a, b) in enumerate(self.collision_pairs):
            m.pair_a[k] = names.index(a)
            m.pair_b[k] = names.index(b)
        m.col_alpha = float(self.collision_alpha)
        m.pose_mode = int(self.pose_mode)
        return m

    # ------------------------------------------------------------------ from a Pinocchio model
    @staticmethod
    def from_pinocchio_like(model, collision_model=None, armature=None, frame: T.Optional[str] = None,
                            alpha: float = 1e-4) -> "RobotTable":
        """Flatten a (reduced) Pinocchio model into the device table — what the solve path needs of
        ``RobotModels.robot_model`` / ``.collision_model`` / ``.armature``
        (``agimus_controller/agimus_controller/factory/robot_model.py:88-351``).

        Duck-typed: only attribute access, so the object may be a real ``pinocchio.Model`` or anything with the same
        fields: ``njoints``, ``names``, ``parents``, ``jointPlacements[i].rotation/.translation``,
        ``inertias[i].mass/.lever/.inertia``, ``joints[i].shortname()`` (+ ``.axis`` for the *Unaligned joints),
        ``gravity.linear``, ``frames[k].name/.parentJoint/.placement``; for the collision model
        ``geometryObjects[k].name/.parentJoint/.placement/.geometry`` (capsule: ``radius`` + ``halfLength`` along the
        local z axis, as ``coal.Capsule``; sphere: ``radius`` only) and ``collisionPairs[k].first/.second``.
        Joints with nq != nv (free-flyer, spherical, continuous) are refused: the state is a vector space on the device.
        """
        nj = int(model.njoints)
        nv = nj - 1
        if nv < 1 or nv > _abi.AGX_MAX_NV:
            raise NotImplementedError(f"{nv} joints: the device tables hold 1..{_abi.AGX_MAX_NV}")
        names, parent = [], np.full(nv, -1, dtype=np.int32)
        jtype, axis = np.zeros(nv, dtype=np.int32), np.zeros((nv, 3))
        pl_R, pl_p = np.zeros((nv, 3, 3)), np.zeros((nv, 3))
        mass, com, inertia = np.zeros(nv), np.zeros((nv, 3)), np.zeros((nv, 3, 3))
        for i in range(1, nj):
            j = i - 1
            names.append(str(model.names[i]))
            parent[j] = int(model.parents[i]) - 1
            jm = model.joints[i]
            short = jm.shortname() if callable(getattr(jm, "shortname", None)) else str(jm.shortname)
            if short not in _PIN_JOINTS:
                raise NotImplementedError(f"joint '{names[-1]}' is a {short}: only 1-DoF revolute / prismatic joints "
                                          "(nq = nv) are supported on the device path")
            jt, ax = _PIN_JOINTS[short]
            jtype[j] = jt
            axis[j] = np.asarray(jm.axis if ax is None else ax, dtype=np.float64).reshape(3)
            axis[j] /= np.linalg.norm(axis[j])
            pl_R[j], pl_p[j] = _rp(model.jointPlacements[i])
            Y = model.inertias[i]
            mass[j] = float(Y.mass)
            com[j] = np.asarray(Y.lever, dtype=np.float64).reshape(3)
            inertia[j] = np.asarray(Y.inertia, dtype=np.float64).reshape(3, 3)
        frames = {}
        for f in model.frames:
            pj = int(getattr(f, "parentJoint", getattr(f, "parent", 0)))
            if pj >= 1:
                R, p_ = _rp(f.placement)
                frames[str(f.name)] = (pj - 1, R, p_)
        g = getattr(getattr(model, "gravity", None), "linear", (0.0, 0.0, -9.81))
        arm = np.zeros(nv) if armature is None else np.broadcast_to(np.asarray(armature, dtype=np.float64), (nv,)).copy()
        t = RobotTable(joint_names=names, parent=parent, jtype=jtype, axis=axis, placement_R=pl_R, placement_p=pl_p,
                       mass=mass, com=com, inertia=inertia, armature=arm,
                       gravity=np.asarray(g, dtype=np.float64).reshape(3), frames=frames)
        if frame is not None:
            t = t.with_frame(frame)
        if collision_model is not None:
            caps, gnames = {}, []
            for go in collision_model.geometryObjects:
                geo = go.geometry
                gnames.append(str(go.name))
                if not hasattr(geo, "radius"):
                    continue  # boxes / meshes: no distance residual on the device path
                half = float(getattr(geo, "halfLength", 0.0))  # a sphere is a capsule of zero length
                R, p_ = _rp(go.placement)
                pj = int(go.parentJoint)
                caps[str(go.name)] = (pj - 1 if pj >= 1 else -1, p_ - half * R[:, 2], p_ + half * R[:, 2], float(geo.radius))
            pairs = []
            for cp in getattr(collision_model, "collisionPairs", []):
                a, b = gnames[int(cp.first)], gnames[int(cp.second)]
                if a in caps and b in caps:
                    pairs.append((a, b))
            t = dataclasses.replace(t, capsules=caps, collision_pairs=pairs[: _abi.AGX_MAX_COLLISION_PAIRS] if
                                    len(caps) <= _abi.AGX_MAX_CAPSULES else [], collision_alpha=float(alpha))
        return t

    @staticmethod
    def from_robot_models(robot_models, frame: T.Optional[str] = None) -> "RobotTable":
        """``RobotModels`` (factory/robot_model.py:88) -> device table: its reduced ``robot_model``, its
        ``collision_model`` (capsules / spheres) and its ``armature``."""
        return RobotTable.from_pinocchio_like(robot_models.robot_model, getattr(robot_models, "collision_model", None),
                                              getattr(robot_models, "armature", None), frame)

    def device_capsules(self) -> list:
        """Names of the capsules that go to the device, in table order: all of them while they fit, otherwise the ones
        the collision pairs use (a full robot carries dozens: the fer model of the reference's tests has 142 pairs)."""
        if len(self.capsules) <= _abi.AGX_MAX_CAPSULES:
            return list(self.capsules)
        used = []
        for a, b in self.collision_pairs:
            for n in (a, b):
                if n not in used:
                    used.append(n)
        if len(used) > _abi.AGX_MAX_CAPSULES:
            raise NotImplementedError(f"{len(used)} capsules in collision pairs: the device table holds "
                                      f"{_abi.AGX_MAX_CAPSULES}")
        return used

    @staticmethod
    def from_links(
        links: list[Link],
        locked_joints: T.Iterable[str] = (),
        frames: T.Optional[dict[str, tuple[str, T.Sequence[float], T.Sequence[float]]]] = None,
        armature: T.Union[float, T.Sequence[float]] = 0.0,
        gravity: T.Sequence[float] = (0.0, 0.0, -9.81),
        geometries: T.Sequence[tuple] = (),
    ) -> "RobotTable":
        """Reduce a link list: fixed and locked joints (at q = 0) are folded into the moving ancestor.
        ``geometries``: collision primitives ``(name, link, xyz, rpy, radius, length)`` — a capsule along the local z axis
        (length 0: a sphere) attached to ``link``; they end up in ``capsules`` expressed in the moving ancestor's frame."""
        locked = set(locked_joints)
        by_name = {l.name: l for l in links}
        moving: list[Link] = [
            l for l in links if l.joint_type != "fixed" and l.joint_name not in locked
        ]
        index = {l.name: i for i, l in enumerate(moving)}
        nv = len(moving)
        # placement of every link frame in its moving ancestor's body frame
        anchor: dict[str, tuple[int, np.ndarray, np.ndarray]] = {}

        def resolve(name: str) -> tuple[int, np.ndarray, np.ndarray]:
            if name in anchor:
                return anchor[name]
            l = by_name[name]
            if name in index:
                res = (index[name], np.eye(3), np.zeros(3))
            else:
                R = rpy_to_matrix(*l.rpy)
                p = np.asarray(l.xyz, dtype=np.float64)
                if l.parent is None:
                    res = (-1, R, p)
                else:
                    pi, Rp, pp = resolve(l.parent)
                    res = (pi, Rp @ R, pp + Rp @ p)
            anchor[name] = res
            return res

        parent = np.full(nv, -1, dtype=np.int32)
        jtype = np.zeros(nv, dtype=np.int32)
        axis = np.zeros((nv, 3))
        pl_R = np.zeros((nv, 3, 3))
        pl_p = np.zeros((nv, 3))
        for i, l in enumerate(moving):
            R = rpy_to_matrix(*l.rpy)
            p = np.asarray(l.xyz, dtype=np.float64)
            if l.parent is None:
                parent[i], pl_R[i], pl_p[i] = -1, R, p
            else:
                pi, Rp, pp = resolve(l.parent)
                parent[i], pl_R[i], pl_p[i] = pi, Rp @ R, pp + Rp @ p
            jtype[i] = (
                _abi.AGX_JOINT_REVOLUTE if l.joint_type == "revolute" else _abi.AGX_JOINT_PRISMATIC
            )
            axis[i] = np.asarray(l.axis, dtype=np.float64)
        # merge inertias
        parts: list[list[tuple[float, np.ndarray, np.ndarray]]] = [[] for _ in range(nv)]
        for l in links:
            if l.mass <= 0.0:
                continue
            bi, R, p = resolve(l.name)
            if bi < 0:
                continue  # welded to the world: no dynamics
            parts[bi].append((l.mass, p + R @ np.asarray(l.com), R @ _sym(l.inertia) @ R.T))
        mass = np.zeros(nv)
        com = np.zeros((nv, 3))
        inertia = np.zeros((nv, 3, 3))
        for i in range(nv):
            m = sum(pt[0] for pt in parts[i])
            mass[i] = m
            if m > 0:
                com[i] = sum(pt[0] * pt[1] for pt in parts[i]) / m
            for mk, ck, Ik in parts[i]:
                d = ck - com[i]
                inertia[i] += Ik + mk * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
        fr: dict[str, tuple[int, np.ndarray, np.ndarray]] = {}
        for l in links:
            bi, R, p = resolve(l.name)
            if bi >= 0:
                fr[l.name] = (bi, R, p)
        for name, (plink, xyz, rpy) in (frames or {}).items():
            bi, R, p = resolve(plink)
            Rf = rpy_to_matrix(*rpy)
            fr[name] = (bi, R @ Rf, p + R @ np.asarray(xyz, dtype=np.float64))
        arm = np.broadcast_to(np.asarray(armature, dtype=np.float64), (nv,)).copy()
        caps = {}
        for gname, glink, gxyz, grpy, gradius, glength in geometries:
            bi, R, p = resolve(glink)
            Rg = R @ rpy_to_matrix(*grpy)
            pg = p + R @ np.asarray(gxyz, dtype=np.float64)
            half = 0.5 * float(glength)
            caps[gname] = (bi, pg - half * Rg[:, 2], pg + half * Rg[:, 2], float(gradius))
        return RobotTable(
            capsules=caps,
            joint_names=[l.joint_name for l in moving],
            parent=parent,
            jtype=jtype,
            axis=axis,
            placement_R=pl_R,
            placement_p=pl_p,
            mass=mass,
            com=com,
            inertia=inertia,
            armature=arm,
            gravity=np.asarray(gravity, dtype=np.float64),
            frames=fr,
        )


def panda_links() -> list[Link]:
    """Franka Panda (example-robot-data description)."""
    h = _HALF_PI
    L = Link
    return [
        L("panda_link1", None, "panda_joint1", "revolute", (0, 0, 0.333), (0, 0, 0), (0, 0, 1),
          4.970684, (0.003875, 0.002081, -0.04762),
          (0.70337, -0.000139, 0.006772, 0.70661, 0.019169, 0.009117)),
        L("panda_link2", "panda_link1", "panda_joint2", "revolute", (0, 0, 0), (-h, 0, 0), (0, 0, 1),
          0.646926, (-0.003141, -0.02872, 0.003495),
          (0.007962, -0.003925, 0.010254, 0.02811, 0.000704, 0.025995)),
        L("panda_link3", "panda_link2", "panda_joint3", "revolute", (0, -0.316, 0), (h, 0, 0), (0, 0, 1),
          3.228604, (0.027518, 0.039252, -0.066502),
          (0.037242, -0.004761, -0.011396, 0.036155, -0.012805, 0.01083)),
        L("panda_link4", "panda_link3", "panda_joint4", "revolute", (0.0825, 0, 0), (h, 0, 0), (0, 0, 1),
          3.587895, (-0.05317, 0.104419, 0.027454),
          (0.025853, 0.007796, -0.001332, 0.019552, 0.008641, 0.028323)),
        L("panda_link5", "panda_link4", "panda_joint5", "revolute", (-0.0825, 0.384, 0), (-h, 0, 0), (0, 0, 1),
          1.225946, (-0.011953, 0.041065, -0.038437),
          (0.035549, -0.002117, -0.004037, 0.029474, 0.000229, 0.008627)),
        L("panda_link6", "panda_link5", "panda_joint6", "revolute", (0, 0, 0), (h, 0, 0), (0, 0, 1),
          1.666555, (0.060149, -0.014117, -0.010517),
          (0.001964, 0.000109, -0.001158, 0.004354, 0.000341, 0.005433)),
        L("panda_link7", "panda_link6", "panda_joint7", "revolute", (0.088, 0, 0), (h, 0, 0), (0, 0, 1),
          0.735522, (0.010517, -0.004252, 0.061597),
          (0.012516, -0.000428, -0.001196, 0.010027, -0.000741, 0.004815)),
        L("panda_link8", "panda_link7", "panda_joint8", "fixed", (0, 0, 0.107), (0, 0, 0)),
        L("panda_hand", "panda_link8", "panda_hand_joint", "fixed", (0, 0, 0), (0, 0, -_QUARTER_PI), (0, 0, 1),
          0.73, (-0.01, 0, 0.03), (0.001, 0, 0, 0.0025, 0, 0.0017)),
        L("panda_leftfinger", "panda_hand", "panda_finger_joint1", "prismatic", (0, 0, 0.0584), (0, 0, 0),
          (0, 1, 0), 0.015, (0, 0, 0), (2.375e-6, 0, 0, 2.375e-6, 0, 7.5e-7)),
        L("panda_rightfinger", "panda_hand", "panda_finger_joint2", "prismatic", (0, 0, 0.0584), (0, 0, 0),
          (0, -1, 0), 0.015, (0, 0, 0), (2.375e-6, 0, 0, 2.375e-6, 0, 7.5e-7)),
    ]


PANDA_FRAMES = {"panda_hand_tcp": ("panda_hand", (0, 0, 0.1034), (0, 0, 0))}
PANDA_Q_NOMINAL = np.array([0.0, -0.78, 0.0, -2.35, 0.0, 1.57, 0.78])  # dummy_mpc_test.py:89


# Synthetic capsule table (the reference's `fer_link*_sc_capsule_*` geometry comes from franka_description, which is
# not in the tree): one capsule on link 3, one on link 7 / hand, one obstacle in the world
# (agimus_controller/tests/resources/environment.xacro:23-24: direction x, radius 0.1, length 0.4, moved into reach).
PANDA_CAPSULES = {
    "link3_capsule": ("panda_joint3", (0.0, 0.0, -0.12), (0.0, 0.0, -0.02), 0.07),
    "link7_capsule": ("panda_joint7", (0.0, 0.0, 0.06), (0.0, 0.0, 0.20), 0.06),
    "obstacle_capsule": (None, (0.35, -0.2, 0.30), (0.35, 0.2, 0.30), 0.05),
}
PANDA_COLLISION_PAIRS = [("link7_capsule", "link3_capsule"), ("link7_capsule", "obstacle_capsule")]


def panda_table(
    lock_fingers: bool = True,
    armature: T.Union[float, T.Sequence[float]] = 0.1,
    frame: str = "panda_hand_tcp",
) -> RobotTable:
    """7-DoF (fingers locked at 0, as every reference test does) or 9-DoF Panda table."""
    locked = ("panda_finger_joint1", "panda_finger_joint2") if lock_fingers else ()
    t = RobotTable.from_links(panda_links(), locked, PANDA_FRAMES, armature=armature)
    return t.with_frame(frame)
