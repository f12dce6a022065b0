"""Known-answer cases for the collision-distance residual (A10): closed-form distances and gradients of capsule /
sphere pairs, so that the row has a pin that is not the restatement itself.

A capsule is a segment with a radius (``coal.Capsule(radius, halfLength)`` along the local z axis,
factory/robot_model.py:269-300), a sphere a capsule of zero length.  colmpc's residual is the signed distance between
the two surfaces; its gradient is n^T (J_a(c_a) - J_b(c_b)) at the closest points.  The oracle, the emulated chain
kernels and the emulated tree kernels are all checked against the geometry worked out by hand below."""
import numpy as np
import pytest

from agimus_controller_b200 import _abi
from agimus_controller_b200.problem import pack_refs
from agimus_controller_b200.robot_model import Link, RobotTable
from emul import emu


def _gantry(nv=6):
    """A Cartesian gantry: three prismatic joints along x, y, z carrying body A, then three more (a second, independent
    branch off the world) carrying body B — positions of A and B are read off q directly."""
    L = Link
    box = (0.01, 0, 0, 0.01, 0, 0.01)
    links = [
        L("ax", None, "ax", "prismatic", (0, 0, 0), (0, 0, 0), (1, 0, 0), 1.0, (0, 0, 0), box),
        L("ay", "ax", "ay", "prismatic", (0, 0, 0), (0, 0, 0), (0, 1, 0), 1.0, (0, 0, 0), box),
        L("az", "ay", "az", "prismatic", (0, 0, 0), (0, 0, 0), (0, 0, 1), 1.0, (0, 0, 0), box),
        L("bx", None, "bx", "prismatic", (0, 0, 0), (0, 0, 0), (1, 0, 0), 1.0, (0, 0, 0), box),
        L("by", "bx", "by", "prismatic", (0, 0, 0), (0, 0, 0), (0, 1, 0), 1.0, (0, 0, 0), box),
        L("bz", "by", "bz", "prismatic", (0, 0, 0), (0, 0, 0), (0, 0, 1), 1.0, (0, 0, 0), box),
    ]
    return RobotTable.from_links(links, (), {"tool": ("az", (0, 0, 0), (0, 0, 0))}, armature=0.1).with_frame("tool")


CASES = {
    # name: (capsule A in body A, capsule B in body B, position of A, position of B, distance, unit normal B -> A)
    "sphere_sphere": (((0, 0, 0), (0, 0, 0), 0.10), ((0, 0, 0), (0, 0, 0), 0.20), (0.0, 0.0, 0.0), (0.6, 0.8, 0.0), 1.0 - 0.3,
                      (-0.6, -0.8, 0.0)),
    "sphere_capsule_side": (((0, 0, 0), (0, 0, 0), 0.05), ((0, 0, -0.5), (0, 0, 0.5), 0.10), (0.4, 0.0, 0.2), (0.0, 0.0, 0.0),
                            0.4 - 0.15, (1.0, 0.0, 0.0)),
    "sphere_capsule_endcap": (((0, 0, 0), (0, 0, 0), 0.05), ((0, 0, -0.5), (0, 0, 0.5), 0.10), (0.0, 0.3, 0.9), (0.0, 0.0, 0.0),
                              0.5 - 0.15, (0.0, 0.6, 0.8)),
    "capsules_crossing": (((-0.5, 0, 0), (0.5, 0, 0), 0.05), ((0, -0.5, 0), (0, 0.5, 0), 0.07), (0.1, 0.0, 0.4), (0.0, 0.2, 0.0),
                          0.4 - 0.12, (0.0, 0.0, 1.0)),
    "capsules_collinear_endcaps": (((0, 0, -0.2), (0, 0, 0.2), 0.05), ((0, 0, -0.3), (0, 0, 0.3), 0.05), (0.0, 0.0, 1.0),
                                   (0.0, 0.0, 0.0), 1.0 - 0.2 - 0.3 - 0.1, (0.0, 0.0, 1.0)),
    "capsules_penetrating": (((0, 0, 0), (0, 0, 0), 0.3), ((0, 0, 0), (0, 0, 0), 0.3), (0.0, 0.0, 0.0), (0.5, 0.0, 0.0),
                             0.5 - 0.6, (-1.0, 0.0, 0.0)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_distance_and_gradient_known_answers(orc, name):
    capA, capB, pA, pB, dist, normal = CASES[name]
    alpha, w = 0.05, 3.0
    t = _gantry().with_capsules({"A": ("az", *capA), "B": ("bz", *capB)}, [("A", "B")], alpha)
    m = t.to_struct()
    q = np.array(list(pA) + list(pB), dtype=float)
    n = np.asarray(normal, dtype=float)
    n /= np.linalg.norm(n)
    # oracle: distance, gradient d r / d q = [n, -n] on a gantry, activation exp(-r^2 / alpha)
    d, Rq, act = orc.collision(m, q, 0)
    assert abs(d - dist) < 1e-12, (d, dist)
    np.testing.assert_allclose(Rq, np.concatenate([n, -n]), atol=1e-12)
    assert abs(act[0] - np.exp(-dist * dist / alpha)) < 1e-14
    # the same through the product kernels (general-tree path: the gantry is not a 7-joint chain)
    nv = 6
    refs = pack_refs(nv, 1, 1, np.zeros(2 * nv), np.zeros(2 * nv), np.zeros(nv), np.zeros(nv), np.eye(3), np.zeros(3),
                     np.zeros(6), wcol=[w, 0.0])
    xs = np.concatenate([q, np.zeros(nv)])[None, None].repeat(2, 1)
    us = np.zeros((1, 1, nv))
    e = emu.calc_diff(m, refs, np.array([0.01]), xs, us)
    a, a1, a2 = np.exp(-dist**2 / alpha), -2 * dist / alpha * np.exp(-dist**2 / alpha), (4 * dist**2 / alpha**2 - 2 / alpha) * np.exp(-dist**2 / alpha)
    g = np.concatenate([n, -n])
    assert abs(e["cost"][0, 1] - w * a) < 1e-13                       # terminal node: unscaled
    np.testing.assert_allclose(e["Lx"][0, 1, :nv], w * a1 * g, atol=1e-12)
    np.testing.assert_allclose(e["Lxx"][0, 1, :nv, :nv], w * a2 * np.outer(g, g), atol=1e-11)
    te = emu.cost_terms(m, refs, np.array([0.01]), xs, us)
    assert abs(te[0, 1, 11] - dist) < 1e-12


def test_capsule_on_a_rotating_link_gradient():
    """A sphere on the tip of a 1-DoF pendulum (revolute about y through the origin, arm along x) against a world
    sphere above it: r(q) = | p(q) - c | - r_a - r_b with p(q) = l (cos q, 0, -sin q): the gradient is the analytic
    derivative (this exercises the rotational Jacobian z x (c_a - p) that the gantry cannot)."""
    L = Link
    links = [L("arm", None, "hinge", "revolute", (0, 0, 0), (0, 0, 0), (0, 1, 0), 1.0, (0.5, 0, 0), (0.01, 0, 0, 0.01, 0, 0.01)),
             L("d1", "arm", "d1", "prismatic", (0, 0, 0), (0, 0, 0), (1, 0, 0), 0.1, (0, 0, 0), (1e-3, 0, 0, 1e-3, 0, 1e-3)),
             L("d2", "d1", "d2", "prismatic", (0, 0, 0), (0, 0, 0), (0, 1, 0), 0.1, (0, 0, 0), (1e-3, 0, 0, 1e-3, 0, 1e-3)),
             L("d3", "d2", "d3", "prismatic", (0, 0, 0), (0, 0, 0), (0, 0, 1), 0.1, (0, 0, 0), (1e-3, 0, 0, 1e-3, 0, 1e-3)),
             L("d4", "d3", "d4", "revolute", (0, 0, 0), (0, 0, 0), (0, 0, 1), 0.1, (0, 0, 0), (1e-3, 0, 0, 1e-3, 0, 1e-3)),
             L("d5", "d4", "d5", "revolute", (0, 0, 0), (0, 0, 0), (1, 0, 0), 0.1, (0, 0, 0), (1e-3, 0, 0, 1e-3, 0, 1e-3))]
    t = RobotTable.from_links(links, (), {"tool": ("d5", (0, 0, 0), (0, 0, 0))}, armature=0.1).with_frame("tool")
    l, c, ra, rb = 0.8, np.array([0.3, 0.0, 0.9]), 0.05, 0.1
    t = t.with_capsules({"tip": ("hinge", (l, 0, 0), (l, 0, 0), ra), "ball": (None, tuple(c), tuple(c), rb)}, [("tip", "ball")], 0.05)
    m = t.to_struct()
    from oracle import orc

    for q0 in (-0.4, 0.1, 0.7):
        q = np.array([q0, 0, 0, 0, 0, 0], dtype=float)
        p = l * np.array([np.cos(q0), 0.0, -np.sin(q0)])
        dp = l * np.array([-np.sin(q0), 0.0, -np.cos(q0)])
        dist = np.linalg.norm(p - c) - ra - rb
        d, Rq, _ = orc.collision(m, q, 0)
        assert abs(d - dist) < 1e-12
        assert abs(Rq[0] - (p - c) @ dp / np.linalg.norm(p - c)) < 1e-12
        assert np.abs(Rq[1:]).max() == 0.0
        nv = 6
        refs = pack_refs(nv, 1, 1, np.zeros(2 * nv), np.zeros(2 * nv), np.zeros(nv), np.zeros(nv), np.eye(3), np.zeros(3),
                         np.zeros(6), wcol=[2.0, 0.0])
        xs = np.concatenate([q, np.zeros(nv)])[None, None].repeat(2, 1)
        e = emu.calc_diff(m, refs, np.array([0.01]), xs, np.zeros((1, 1, nv)))
        a1 = -2 * dist / 0.05 * np.exp(-dist**2 / 0.05)
        assert abs(e["Lx"][0, 1, 0] - 2.0 * a1 * Rq[0]) < 1e-12
