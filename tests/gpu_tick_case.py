"""Closed-loop MPC ticks in latency mode (eager_exit, at most 64 problems) through the C ABI; the results of every tick go
to an .npz file.  test_gpu_tick_graph.py runs this once with the tick graph and once with AGX_TICK_GRAPH=0 (the stream
path) and compares the two files bit for bit."""
import sys

import numpy as np
import torch

from agimus_controller_b200 import _abi, panda_table
from agimus_controller_b200.solver import BatchedShootingProblem
from agimus_controller_b200.workloads import goal_reaching_batch


def run(B, ticks, queue_without_sync, mode="fddp"):
    table = panda_table()
    helper = BatchedShootingProblem(table, np.full(2, 0.01), 1)
    rn = lambda q, v, a: helper.rnea(q, v, a).cpu().numpy()  # noqa: E731
    if mode.endswith("_col"):
        # capsule-pair collision costs: the COL instantiations of the kernels inside the graph
        from agimus_controller_b200.workloads import pick_and_place_collision_batch

        w = pick_and_place_collision_batch(B, T=20, rnea=rn)
        table = w["table"]
        mode = mode[:-4]
    elif mode.endswith("_nv9"):
        # the nine-joint Panda (fingers unlocked): general-tree kernels inside the graph
        from agimus_controller_b200.workloads import pick_and_place_collision_batch

        t9 = panda_table(lock_fingers=False)
        h9 = BatchedShootingProblem(t9, np.full(2, 0.01), 1)
        w = pick_and_place_collision_batch(B, T=20, rnea=lambda q, v, a: h9.rnea(q, v, a).cpu().numpy(),
                                           lock_fingers=False)
        table = w["table"]
        mode = mode[:-4]
    else:
        w = goal_reaching_batch(B, T=20, rnea=rn, seed=5)
    p = BatchedShootingProblem(table, w["dts"], B)
    p.set_refs(w["refs"])
    max_iter = 10
    if mode == "fddp_long":
        # a budget above 32 iterations solved to convergence: the graph serves it at any batch size
        mode, max_iter = "fddp", 60
        opts = _abi.default_fddp_opts()
    else:
        opts = _abi.default_fddp_opts() if mode == "fddp" else _abi.default_sqp_opts()
        opts.eager_exit = 1
    solve = p.solve if mode == "fddp" else p.solve_sqp
    x = torch.as_tensor(w["x0"], device="cuda")
    xs = torch.as_tensor(w["xs_ws"], device="cuda")
    us = torch.as_tensor(w["us_ws"], device="cuda")
    res = {}
    outs = [p.alloc_outputs() for _ in range(ticks)] if queue_without_sync else None
    for k in range(ticks):
        out = solve(x, xs, us, max_iter, opts, out=outs[k] if outs else None)
        if not queue_without_sync:
            for name in ("xs", "us", "K", "cost", "iters", "status"):
                res[f"{name}_{k}"] = out[name].cpu().numpy()
        xs, us = p.shift_warmstart(out["xs"], out["us"])
        x = xs[:, 0].contiguous()
    if queue_without_sync:
        torch.cuda.synchronize()
        for k in range(ticks):
            for name in ("xs", "us", "K", "cost", "iters", "status"):
                res[f"{name}_{k}"] = outs[k][name].cpu().numpy()
    res["launches"] = np.array([p.launch_count])
    return res


if __name__ == "__main__":
    out_path, B, ticks, queue = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    mode = sys.argv[5] if len(sys.argv) > 5 else "fddp"
    np.savez(out_path, **run(B, ticks, bool(queue), mode))
