"""A small URDF reader: URDF text -> ``RobotTable`` (the device table of the solve path).

The reference builds its model with ``pinocchio.buildModelFromXML`` + ``buildReducedModel``
(``agimus_controller/agimus_controller/factory/robot_model.py:160-259``); Pinocchio is not needed on the solve path, so
this reads the same description directly: links (inertial with its origin), joints (revolute / continuous /
prismatic / fixed, origin, axis), collision cylinders / spheres / capsules.  Joints that are not listed in
``moving_joint_names`` are locked at zero and their bodies merged into the moving ancestor
(``RobotTable.from_links``), cylinders become capsules named ``<link>_capsule_<i>`` when ``collision_as_capsule``
(``factory/robot_model.py:261-302``), spheres keep ``<link>_<k>`` (Pinocchio's geometry naming).
Meshes and boxes are skipped: they carry no distance residual on the device path.
"""
from __future__ import annotations

import pathlib
import typing as T
import xml.etree.ElementTree as ET

import numpy as np

from .robot_model import Link, RobotTable, _sym, rpy_to_matrix


def _floats(text: T.Optional[str], n: int, default: float = 0.0) -> tuple:
    if not text:
        return tuple([default] * n)
    v = [float(x) for x in text.split()]
    assert len(v) == n, f"expected {n} numbers, got '{text}'"
    return tuple(v)


def _origin(elem) -> tuple:
    o = elem.find("origin") if elem is not None else None
    if o is None:
        return (0.0, 0.0, 0.0), (0.0, 0.0, 0.0)
    return _floats(o.get("xyz"), 3), _floats(o.get("rpy"), 3)


def parse_urdf(xml: T.Union[str, pathlib.Path]) -> tuple[list[Link], list[tuple]]:
    """``(links, geometries)`` of a URDF given as a string or a path; ``geometries`` as ``RobotTable.from_links`` takes."""
    text = pathlib.Path(xml).read_text() if isinstance(xml, pathlib.Path) or "<" not in str(xml) else str(xml)
    root = ET.fromstring(text)
    inertials, geoms = {}, []
    for le in root.findall("link"):
        name = le.get("name")
        ie = le.find("inertial")
        if ie is not None:
            xyz, rpy = _origin(ie)
            me, te = ie.find("mass"), ie.find("inertia")
            mass = float(me.get("value")) if me is not None else 0.0
            i6 = [float(te.get(k, "0")) for k in ("ixx", "ixy", "ixz", "iyy", "iyz", "izz")] if te is not None else [0.0] * 6
            R = rpy_to_matrix(*rpy)
            I = R @ _sym(i6) @ R.T  # inertia about the COM, in link axes
            inertials[name] = (mass, xyz, (I[0, 0], I[0, 1], I[0, 2], I[1, 1], I[1, 2], I[2, 2]))
        for k, ce in enumerate(le.findall("collision")):
            ge = ce.find("geometry")
            if ge is None:
                continue
            xyz, rpy = _origin(ce)
            cyl, sph, cap = ge.find("cylinder"), ge.find("sphere"), ge.find("capsule")
            if cyl is not None or cap is not None:
                e = cyl if cyl is not None else cap
                geoms.append((name, k, "cylinder" if cyl is not None else "capsule", xyz, rpy, float(e.get("radius")),
                              float(e.get("length"))))
            elif sph is not None:
                geoms.append((name, k, "sphere", xyz, rpy, float(sph.get("radius")), 0.0))
    joints = {}
    children = set()
    for je in root.findall("joint"):
        child = je.find("child").get("link")
        jt = je.get("type")
        if jt == "continuous":
            jt = "revolute"
        if jt not in ("revolute", "prismatic", "fixed"):
            raise NotImplementedError(f"joint '{je.get('name')}' of type {jt}")
        xyz, rpy = _origin(je)
        ae = je.find("axis")
        axis = _floats(ae.get("xyz"), 3) if ae is not None else (1.0, 0.0, 0.0)
        joints[child] = (je.get("name"), jt, je.find("parent").get("link"), xyz, rpy, axis)
        children.add(child)
    links = []
    order = [le.get("name") for le in root.findall("link")]
    roots = [n for n in order if n not in children]

    def visit(name, parent):
        mass, com, i6 = inertials.get(name, (0.0, (0.0, 0.0, 0.0), (0.0,) * 6))
        if parent is None:
            links.append(Link(name, None, name + "_root_joint", "fixed", (0, 0, 0), (0, 0, 0), (0, 0, 1), mass, com, i6))
        else:
            jn, jt, _, xyz, rpy, axis = joints[name]
            n = float(np.linalg.norm(axis)) or 1.0
            links.append(Link(name, parent, jn, jt, xyz, rpy, tuple(a / n for a in axis), mass, com, i6))
        for c in order:
            if c in joints and joints[c][2] == name:
                visit(c, name)

    for r in roots:
        visit(r, None)
    return links, geoms


def load_urdf(xml: T.Union[str, pathlib.Path], moving_joint_names: T.Optional[T.Sequence[str]] = None,
              frame: T.Optional[str] = None, armature: T.Union[float, T.Sequence[float]] = 0.0,
              collision_as_capsule: bool = True, collision_pairs: T.Sequence[tuple] = (), alpha: float = 1e-4,
              gravity: T.Sequence[float] = (0.0, 0.0, -9.81)) -> RobotTable:
    """URDF -> reduced ``RobotTable``: the counterpart of ``RobotModels(RobotModelParameters(...))``
    (``factory/robot_model.py:13-110``): ``moving_joint_names`` (``None`` = every non-fixed joint), ``armature``,
    ``collision_as_capsule``, ``collision_pairs`` (names of geometry objects)."""
    links, geoms = parse_urdf(xml)
    names = {l.joint_name for l in links if l.joint_type != "fixed"}
    if moving_joint_names is not None:
        for jn in moving_joint_names:
            if jn not in names:
                raise ValueError(jn + " not in the model.")
        locked = names - set(moving_joint_names)
    else:
        locked = set()
    geometries, n_caps = [], {}
    for link, k, kind, xyz, rpy, radius, length in geoms:
        if kind == "sphere":
            gname = f"{link}_{k}"
        elif kind == "capsule" or collision_as_capsule:
            i = n_caps.get(link, 0)
            n_caps[link] = i + 1
            gname = f"{link}_capsule_{i}"
        else:
            continue  # a bare cylinder has no closed-form distance on the device path
        geometries.append((gname, link, xyz, rpy, radius, length))
    t = RobotTable.from_links(links, locked, None, armature=armature, gravity=gravity, geometries=geometries)
    # the kinematic order of from_links follows the link list; moving_joint_names only selects
    for a, b in collision_pairs:
        for g in (a, b):
            if g not in t.capsules:
                raise ValueError(f"Invalid collision pair with name {g}")
    if collision_pairs:
        import dataclasses

        t = dataclasses.replace(t, collision_pairs=list(collision_pairs), collision_alpha=float(alpha))
    return t.with_frame(frame) if frame is not None else t
