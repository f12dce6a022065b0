"""experiment: n independent cfg-2 batches in flight on n streams (one handle each) against one batch at a time"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from agimus_controller_b200 import _abi, panda_table
from agimus_controller_b200.solver import BatchedShootingProblem
from agimus_controller_b200.workloads import goal_reaching_batch

dev = torch.device("cuda", 0)
B, T, IT = int(os.environ.get("PB", "4096")), 50, 10
helper = BatchedShootingProblem(panda_table(), np.full(2, 0.01), 1, device=dev)
rn = lambda q, v, a: helper.rnea(q, v, a).cpu().numpy()
opts = _abi.default_fddp_opts(fixed_iters=True)
for n in (1, 2, 3, 4):
    probs, ins, outs, streams = [], [], [], []
    for i in range(n):
        w = goal_reaching_batch(B, T=T, rnea=rn, seed=i)
        p = BatchedShootingProblem(panda_table(), w["dts"], B, device=dev)
        p.set_refs(torch.as_tensor(w["refs"], device=dev))
        probs.append(p)
        ins.append(tuple(torch.as_tensor(w[k], device=dev) for k in ("x0", "xs_ws", "us_ws")))
        outs.append(p.alloc_outputs())
        streams.append(torch.cuda.Stream(device=dev))
    def run(steps):
        for s in range(steps):
            i = s % n
            with torch.cuda.stream(streams[i]):
                probs[i].solve(*ins[i], IT, opts, out=outs[i])
    run(3 * n)
    torch.cuda.synchronize()
    steps = 12 * n
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for st in streams:
        st.wait_event(e0)
    run(steps)
    for st in streams:
        torch.cuda.current_stream().wait_stream(st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{n} batches in flight: {ms:.3f} ms per {B}-problem solve, {B / ms * 1e3:.0f} solves/s", flush=True)
    del probs, ins, outs
