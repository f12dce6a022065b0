import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from agimus_controller_b200 import _abi
from agimus_controller_b200.solver import BatchedShootingProblem
from agimus_controller_b200.workloads import sine_configuration_reference
dev = torch.device("cuda", 0)
T, dt = 20, 0.01
from agimus_controller_b200 import panda_table
p0 = BatchedShootingProblem(panda_table(), np.full(T, dt), 1, device=dev)
table, rows, q, v, u = sine_configuration_reference(400, dt=dt, rnea=lambda q_, v_, a_: p0.rnea(q_, v_, a_).cpu().numpy())
p1 = BatchedShootingProblem(table, np.full(T, dt), 1, device=dev)
rows_d = torch.as_tensor(rows, device=dev)
out = p1.alloc_outputs()
x = torch.as_tensor(np.concatenate([q[0], v[0]])[None], device=dev)
xs = torch.cat([torch.as_tensor(q[: T + 1]), torch.as_tensor(v[: T + 1])], dim=1)[None].to(dev).contiguous()
us = torch.as_tensor(u[:T][None], device=dev).contiguous()
eo = _abi.default_fddp_opts(); eo.eager_exit = 1
for mode, opts, iters in (("fddp", _abi.default_fddp_opts(), 10), ("fddp_eager", eo, 10)):
    acc = {k: [] for k in ("refs", "solve_call", "sync", "d2h")}
    for k in range(200):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); p1.set_refs_window(rows_d, k)
        t1 = time.perf_counter(); p1.solve(x, xs, us, iters, opts, out=out)
        t2 = time.perf_counter(); torch.cuda.synchronize()
        t3 = time.perf_counter(); u0 = out["us"][0, 0].cpu(); K0 = out["K"][0, 0].cpu()
        t4 = time.perf_counter()
        for n, d in zip(acc, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)): acc[n].append(d * 1e6)
    print(mode, {n: round(float(np.median(v[20:])), 1) for n, v in acc.items()}, "launches/solve", 0)
