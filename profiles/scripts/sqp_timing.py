import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from agimus_controller_b200 import _abi, panda_table
from agimus_controller_b200.solver import BatchedShootingProblem
from agimus_controller_b200.workloads import goal_reaching_batch
B, T = 4096, 50
dev = torch.device("cuda", 0)
prob = BatchedShootingProblem(panda_table(), np.full(T, 0.01), B, device=dev)
w = goal_reaching_batch(B, T=T, rnea=lambda q, v, a: prob.rnea(q, v, a).cpu().numpy())
prob.set_refs(w["refs"])
x0, xs, us = (torch.as_tensor(w[k], device=dev) for k in ("x0", "xs_ws", "us_ws"))
out = prob.alloc_outputs()
for _ in range(2):
    prob.solve_sqp(x0, xs, us, 10, None, out=out)
torch.cuda.synchronize()
prob.set_timing(True)
prob.solve_sqp(x0, xs, us, 10, None, out=out)
t = prob.get_timing()
print({k: (round(v["ms"], 3), v["launches"]) for k, v in t.items()})
print("iters hist", torch.bincount(out["iters"]).tolist(), "status", torch.bincount(out["status"]).tolist())
