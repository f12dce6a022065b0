"""Does a kernel's code stay in the SM's instruction cache between launches?  B = 1, T = 20: per-launch GPU time of
calc_diff alone (back to back) against calc_diff interleaved with other large kernels."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
import numpy as np, torch
from agimus_controller_b200 import panda_table
from agimus_controller_b200._lib import lib
from agimus_controller_b200.solver import BatchedShootingProblem, _ptr
from agimus_controller_b200.workloads import goal_reaching_batch

T = 20
h0 = BatchedShootingProblem(panda_table(), np.full(2, 0.01), 1)
w = goal_reaching_batch(1, T=T, rnea=lambda q, v, a: h0.rnea(q, v, a).cpu().numpy())
p = BatchedShootingProblem(panda_table(), w["dts"], 1)
p.set_refs(w["refs"])
xs = torch.as_tensor(w["xs_ws"], device="cuda"); us = torch.as_tensor(w["us_ws"], device="cuda"); x0 = torch.as_tensor(w["x0"], device="cuda")
o = p.calc_diff(xs, us)
oxs = torch.empty_like(xs); from agimus_controller_b200 import _abi; terms = torch.empty(1, T + 1, _abi.AGX_N_COST_TERMS, dtype=torch.float64, device="cuda"); tau = torch.empty(1, 7, dtype=torch.float64, device="cuda")
q = x0[:, :7].contiguous(); v = x0[:, 7:].contiguous(); a = torch.zeros_like(q)
L = lib(); st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def cd():
    L.agx_calc_diff(p._h, _ptr(xs), _ptr(us), *[_ptr(o[k]) for k in ("cost", "xnext", "Fx", "Fu", "Lx", "Lu", "Lxx", "Lxu", "Luu")], st)
def others():
    L.agx_rollout(p._h, _ptr(x0), _ptr(us), _ptr(oxs), st)
    L.agx_cost_terms(p._h, _ptr(xs), _ptr(us), _ptr(terms), st)
    L.agx_rnea(p._h, _ptr(q), _ptr(v), _ptr(a), 1, _ptr(tau), st)
def timed(f, n=300):
    for _ in range(20): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
ta = timed(cd); tc = timed(others); tb = timed(lambda: (cd(), others()))
print(f"calc_diff+expand back to back: {ta:.1f} us; three other kernels: {tc:.1f} us; interleaved: {tb:.1f} us; excess {tb - ta - tc:.1f} us")
