"""Reference-record packing: per-node references and weights -> the ``[B][T+1][ref_size]`` table.

This is the flattened form of what ``OCPCrocoGeneric.set_reference_weighted_trajectory``
(``agimus_controller/agimus_controller/ocp/ocp_croco_generic.py:855-892``) writes into Crocoddyl
objects one Boost.Python attribute at a time.  Record layout (``include/agx.h``):
``[xref nx][wx nx][uref nu][wu nu][Rref 9][pref 3][wpose 6][wcol 2]`` with the CostModelSum weight folded
into the activation weights; ``wcol`` are the scalar weights of the (up to two) collision pairs.
"""
from __future__ import annotations

import numpy as np

from ._abi import ref_size


def pack_refs(nv, T, B, xref, wx, uref, wu, Rref, pref, wpose, wpose_terminal=None, wx_terminal=None, wcol=None):
    """Broadcast the given references/weights to ``[B, T+1, ref_size]`` (float64, C order).

    Every argument broadcasts against ``[B, T+1, n]``; ``*_terminal`` overrides the last node.
    """
    nx = 2 * nv
    rs = ref_size(nv)
    r = np.zeros((B, T + 1, rs))
    r[..., 0:nx] = np.broadcast_to(np.asarray(xref, dtype=np.float64), (B, T + 1, nx))
    r[..., nx : 2 * nx] = np.broadcast_to(np.asarray(wx, dtype=np.float64), (B, T + 1, nx))
    o = 2 * nx
    r[..., o : o + nv] = np.broadcast_to(np.asarray(uref, dtype=np.float64), (B, T + 1, nv))
    r[..., o + nv : o + 2 * nv] = np.broadcast_to(np.asarray(wu, dtype=np.float64), (B, T + 1, nv))
    o += 2 * nv
    r[..., o : o + 9] = np.broadcast_to(np.asarray(Rref, dtype=np.float64).reshape(-1, 9) if np.ndim(Rref) == 2
                                         else np.asarray(Rref, dtype=np.float64).reshape(np.shape(Rref)[:-2] + (9,)),
                                         (B, T + 1, 9))
    r[..., o + 9 : o + 12] = np.broadcast_to(np.asarray(pref, dtype=np.float64), (B, T + 1, 3))
    r[..., o + 12 : o + 18] = np.broadcast_to(np.asarray(wpose, dtype=np.float64), (B, T + 1, 6))
    if wcol is not None:
        r[..., o + 18 : o + 20] = np.broadcast_to(np.asarray(wcol, dtype=np.float64), (B, T + 1, 2))
    if wpose_terminal is not None:
        r[:, T, o + 12 : o + 18] = np.broadcast_to(np.asarray(wpose_terminal, dtype=np.float64), (B, 6))
    if wx_terminal is not None:
        r[:, T, nx : 2 * nx] = np.broadcast_to(np.asarray(wx_terminal, dtype=np.float64), (B, nx))
    r[:, T, 2 * nx + nv : 2 * nx + 2 * nv] = 0.0  # terminal node has no control cost
    return r
