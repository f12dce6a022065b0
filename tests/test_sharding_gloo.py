"""N > 1 path on the CPU: world_size-2 gloo run of the sharding + statistics gather.

Each rank solves its slab through the emulated kernels (tests/emul) and the gathered cost/iters/status must
be bitwise identical to the single-process solve of the whole batch (no cross-problem arithmetic).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, T, iters, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from agimus_controller_b200 import _abi, panda_table
    from agimus_controller_b200.sharding import gather_stats, shard_range
    from agimus_controller_b200.workloads import goal_reaching_batch
    from emul import emu
    from oracle import orc

    m = panda_table().to_struct()
    w = goal_reaching_batch(B, T=T, rnea=lambda q_, v, a: orc.rnea(m, q_, v, a))
    r = shard_range(B, world, rank)
    sl = slice(r.start, r.stop)
    opts = _abi.default_fddp_opts(fixed_iters=True)
    e = emu.solve(m, w["refs"][sl], w["dts"], w["x0"][sl], w["xs_ws"][sl], w["us_ws"][sl], iters, opts)
    g = gather_stats(torch.as_tensor(e["cost"]), torch.as_tensor(e["iters"]), torch.as_tensor(e["status"]), B)
    if rank == 0:
        q.put({k: v.numpy() for k, v in g.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    from agimus_controller_b200.sharding import shard_range

    for B in (1, 7, 4096, 16384, 65537):
        for world in (1, 2, 4, 8):
            idx = [i for r in range(world) for i in shard_range(B, world, r)]
            assert idx == list(range(B))
            sizes = [len(shard_range(B, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gloo_matches_single_process(orc):
    from agimus_controller_b200 import _abi, panda_table
    from agimus_controller_b200.workloads import goal_reaching_batch
    from emul import emu

    B, T, iters, world = 5, 8, 3, 2
    emu.lib()  # build once before forking workers
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, T, iters, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m = panda_table().to_struct()
    w = goal_reaching_batch(B, T=T, rnea=lambda q_, v, a: orc.rnea(m, q_, v, a))
    e = emu.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], iters, _abi.default_fddp_opts(fixed_iters=True))
    np.testing.assert_array_equal(got["cost"], e["cost"])
    np.testing.assert_array_equal(got["iters"], e["iters"])
    np.testing.assert_array_equal(got["status"], e["status"])
