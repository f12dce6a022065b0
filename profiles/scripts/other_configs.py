"""BASELINE.json configs 3, 4 and 5 at full size (secondary figures; bench.py measures config 2, the metric's).

Run under torchrun like bench.py (`--nproc-per-node N`), or plainly for one GPU.  Per config: the batch is sharded in
contiguous slabs over the ranks, inputs are resident, W warm-up solves, K timed solves bracketed by a barrier and a
device synchronisation, max over ranks; rank 0 prints one JSON line per config.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    import torch
    import torch.distributed as dist

    from agimus_controller_b200 import _abi, panda_table
    from agimus_controller_b200.sharding import shard_range
    from agimus_controller_b200.solver import BatchedShootingProblem
    from agimus_controller_b200.workloads import (cartesian_sine_batch, model_sensibility_batch,
                                                  pick_and_place_collision_batch)

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    helper = BatchedShootingProblem(panda_table(), np.full(2, 0.01), 1, device=dev)
    rn = lambda q, v, a: helper.rnea(q, v, a).cpu().numpy()  # noqa: E731
    helper9 = BatchedShootingProblem(panda_table(lock_fingers=False), np.full(2, 0.01), 1, device=dev)
    rn9 = lambda q, v, a: helper9.rnea(q, v, a).cpu().numpy()  # noqa: E731
    steps, warm = int(os.environ.get("AGX_STEPS", "10")), 3
    which = os.environ.get("AGX_CONFIGS", "3,4,5").split(",")
    per_gpu5 = int(os.environ.get("AGX_CFG5_PER_GPU", "8192"))
    cases = {
        "3": ("cfg3: 16384 Cartesian sine end-effector tracking OCPs, T=50, 10 FDDP iterations", 16384, 10,
              lambda B: cartesian_sine_batch(B, T=50, rnea=rn)),
        "4": ("cfg4: 4096 pick-and-place OCPs with two capsule-pair collision costs (fingers locked), T=100, 3 FDDP "
              "iterations", 4096, 3, lambda B: pick_and_place_collision_batch(B, T=100, rnea=rn)),
        "4f": ("cfg4 as BASELINE.json states it: 4096 pick-and-place OCPs, nv=9 WITH the finger joints (general-tree "
               "kernels), two capsule-pair collision costs, T=100, 3 FDDP iterations", 4096, 3,
               lambda B: pick_and_place_collision_batch(B, T=100, rnea=rn9, lock_fingers=False)),
        "5": (f"cfg5: {per_gpu5 * world} goal-reaching OCPs with per-problem inertial tables, T=50, 10 FDDP iterations",
              per_gpu5 * world, 10, lambda B: model_sensibility_batch(B, T=50, rnea=rn)),
    }
    for key in which:
        name, B, iters, build = cases[key]
        slab = shard_range(B, world, rank)
        lo, hi = slab.start, slab.stop
        w = build(B)                                   # the whole batch is generated, the rank keeps its slab
        tables = w["tables"][lo:hi] if "tables" in w else w["table"]
        prob = BatchedShootingProblem(tables, w["dts"], hi - lo, device=dev)
        prob.set_refs(torch.as_tensor(w["refs"][lo:hi], device=dev))
        x0, xs, us = (torch.as_tensor(w[k][lo:hi], device=dev) for k in ("x0", "xs_ws", "us_ws"))
        out = prob.alloc_outputs()
        opts = _abi.default_fddp_opts(fixed_iters=True)
        for _ in range(warm):
            prob.solve(x0, xs, us, iters, opts, out=out)
        torch.cuda.synchronize()
        phases = None
        if os.environ.get("AGX_PHASES"):
            prob.set_timing(True)
            prob.solve(x0, xs, us, iters, opts, out=out)
            phases = {k: v for k, v in prob.get_timing().items() if v["launches"]}
            prob.set_timing(False)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            prob.solve(x0, xs, us, iters, opts, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        finite = bool(torch.isfinite(out["xs"]).all())
        if rank == 0:
            print(json.dumps({"config": name, "n_gpus": world, "B_total": B, "iters": iters, "steps": steps,
                              "ms_per_step": float(ms) / steps, "solves_per_s": B * steps / (float(ms) * 1e-3),
                              "finite": finite, "mean_cost_rank0": float(out["cost"].mean()), "phases": phases}),
                  flush=True)
        del prob
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
