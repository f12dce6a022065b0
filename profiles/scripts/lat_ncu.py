import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from agimus_controller_b200 import _abi, panda_table
from agimus_controller_b200.solver import BatchedShootingProblem
from agimus_controller_b200.workloads import sine_configuration_reference
dev = torch.device("cuda", 0)
T, dt = 20, 0.01
p0 = BatchedShootingProblem(panda_table(), np.full(T, dt), 1, device=dev)
table, rows, q, v, u = sine_configuration_reference(60, dt=dt, rnea=lambda q_, v_, a_: p0.rnea(q_, v_, a_).cpu().numpy())
p1 = BatchedShootingProblem(table, np.full(T, dt), 1, device=dev)
rows_d = torch.as_tensor(rows, device=dev)
out = p1.alloc_outputs()
x = torch.as_tensor(np.concatenate([q[0], v[0]])[None], device=dev)
xs = torch.cat([torch.as_tensor(q[: T + 1]), torch.as_tensor(v[: T + 1])], dim=1)[None].to(dev).contiguous()
us = torch.as_tensor(u[:T][None], device=dev).contiguous()
opts = _abi.default_fddp_opts()
opts.eager_exit = 1  # latency mode
for k in range(6):
    p1.set_refs_window(rows_d, k)
    p1.solve(x, xs, us, 10, opts, out=out)
    torch.cuda.synchronize()
    x = p1.integrate(x, out["us"][:, 0], dt)
    xs, us = p1.shift_warmstart(out["xs"], out["us"]); xs[:, 0] = x
print("iters", int(out["iters"][0]))
