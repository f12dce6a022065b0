#!/usr/bin/env python
"""bench.py — OCP solves/s of the batched FDDP solve path (BASELINE.json metric, config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path over one batch: B = 4096 Panda goal-reaching OCPs (T = 50,
dt = 0.01, randomised initial states), exactly 10 FDDP iterations each (cfg 2).
N > 1 (torchrun, one rank per GPU): every rank solves its own slab of 4096 problems (weak scaling, no
data-path collective; one NCCL all_gather of cost/iters/status per step).

Printed JSON (rank 0, one line):
  value        solves/s with inputs resident in HBM (device-timed with CUDA events, max over ranks); three batches
               are in flight (step i on handle/stream i % 3: independent batches run in the idle last waves of each
               other's sequential kernels); `serial` = one batch at a time
  e2e          same metric through the public API with HOST buffers: pinned H2D of x0 / warm start /
               references and D2H of xs, us, K[:, 0], cost, iters, status inside the timed region
  roofline     the dominant kernel against the FP64 (non-tensor) peak measured in-run by a DFMA probe,
               plus its HBM side; durations from CUDA-event pairs around every launch of the serial pass
  cpu_baseline the CPU restatement (oracle, OpenMP one problem per thread) on a bounded sample
--impl reference times that CPU restatement as the reference arm (Crocoddyl itself cannot be installed:
its sources are not in the reference tree and there is no network).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU = int(os.environ.get("AGX_BENCH_B", "4096"))  # 4096 is the metric's batch; the override is for scaling studies
T_NODES = 50
DT = 0.01
N_ITERS = 10
METRIC = "OCP solves/sec, Panda T=50 batch 4096, fixed 10 FDDP iterations"
WORKLOAD = "cfg2: 4096 Panda goal-reaching OCPs per GPU, T=50, dt=0.01, randomised x0, fixed 10 FDDP iterations"

# agreed algorithmic work per node and iteration, FMA = 2 flops.  calc+calcDiff (16 000) is split
# between the dynamics kernel (ABA 2500 + RNEA derivatives 5500 + Minv 1500 + two Minv products 1400 + Euler
# scaling 1000 + regs 1000 = 12 900) and the cost kernel (FK / frame Jacobian / log6 / Jlog6 1800 + Gauss-Newton
# assembly 1300 = 3 100, which also yields the forward pass's cost evaluation, 600); forward (3 500) = rollout
# (ABA 2500 + K dx / integrate 400 = 2 900) + that cost evaluation.
FLOP = {"calc_diff": 12900.0, "backward": 23500.0, "rollout_try": 2900.0, "node_cost": 3700.0,
        "accept_linesearch": 0.0}
REC_BYTES = 184 * 8    # dynamics record
CREC_BYTES = 64 * 8    # cost record


def build_workload(B, seed, rnea):
    """cfg-2 tables; `rnea` supplies the gravity-compensation warm start (device kernel in our arm)."""
    from agimus_controller_b200 import panda_table
    from agimus_controller_b200.workloads import goal_reaching_batch

    m = panda_table().to_struct()
    w = goal_reaching_batch(B, T=T_NODES, dt=DT, seed=seed, rnea=rnea)
    return w, m


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs (one streaming `-lms` process)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        time.sleep(0.3)  # let the first samples arrive before the timed region starts

    def stop(self):
        rows = []
        if self.proc is not None:
            try:
                self.proc.terminate()
                out, _ = self.proc.communicate(timeout=5)
                rows = [[c.strip() for c in ln.split(",")] for ln in out.strip().splitlines()]
            except Exception:
                pass
        sm, smax, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def time_cpu(orc, m, w, n, max_iter, threads=0):
    from agimus_controller_b200 import _abi

    opts = _abi.default_fddp_opts(fixed_iters=True)
    sl = slice(0, n)
    t0 = time.perf_counter()
    orc.solve(m, w["refs"][sl], w["dts"], w["x0"][sl], w["xs_ws"][sl], w["us_ws"][sl], max_iter, opts, nthreads=threads)
    return time.perf_counter() - t0


def run_reference(args):
    """Reference arm: the CPU restatement of the path on the host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from oracle import orc

    # torchrun exports OMP_NUM_THREADS=1: the thread count is passed explicitly so that the reference arm
    # uses every host core it may run on
    cores = len(os.sched_getaffinity(0))
    # the whole batch of the metric per step (same config as our arm); --sample bounds it on slow hosts
    n = B_PER_GPU if args.sample is None else min(args.sample, B_PER_GPU)
    from agimus_controller_b200 import panda_table

    m0 = panda_table().to_struct()
    w, m = build_workload(B_PER_GPU, 0, lambda q, v, a: orc.rnea(m0, q, v, a))
    for _ in range(args.warmup):
        time_cpu(orc, m, w, n, N_ITERS, threads=cores)
    t = [time_cpu(orc, m, w, n, N_ITERS, threads=cores) for _ in range(args.steps)]
    total = sum(t)
    value = n * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "B_per_gpu": B_PER_GPU, "T": T_NODES, "dt": DT, "fddp_iters": N_ITERS,
                   "sample": f"{n} of the {B_PER_GPU} problems per step"},
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": "port",
                         "sample": f"{n} problems x {args.steps} steps, OpenMP one problem per thread, "
                                   "CPU restatement of Crocoddyl FDDP (oracle/agx_oracle.cpp)"},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def mpc_latency(prob, dev, ticks, solver="fddp", with_cpu=False):
    """p50 / p99 of one MPC tick's solve (reference: MPCDebugData.duration_ocp_solve_ns, mpc.py:52-64): reference
    update, warm start, solve, read-back of the control the node publishes (us[0], K[0])."""
    import torch

    from agimus_controller_b200 import _abi
    from agimus_controller_b200.solver import BatchedShootingProblem
    from agimus_controller_b200.workloads import sine_configuration_reference

    T, dt = 20, DT
    table, rows, q, v, u = sine_configuration_reference(ticks + T + 2, dt=dt,
                                                        rnea=lambda q_, v_, a_: prob.rnea(q_, v_, a_).cpu().numpy())
    p1 = BatchedShootingProblem(table, np.full(T, dt), 1, device=dev)
    nv = table.nv
    rows_d = torch.as_tensor(rows, device=dev)
    opts = _abi.default_fddp_opts()
    opts.eager_exit = 1   # latency mode: one graph launch per FDDP tick, loop condition on the device (include/agx.h)
    sqp_opts = _abi.default_sqp_opts()
    sqp_opts.eager_exit = 1
    out = p1.alloc_outputs()
    x = torch.as_tensor(np.concatenate([q[0], v[0]])[None], device=dev)
    xs = torch.cat([torch.as_tensor(q[: T + 1]), torch.as_tensor(v[: T + 1])], dim=1)[None].to(dev).contiguous()
    us = torch.as_tensor(u[:T][None], device=dev).contiguous()
    ts, iters = [], []
    # what Control(feedback_gain, feedforward) carries lands in pinned host buffers: two asynchronous copies and ONE
    # stream synchronisation per tick (a `.cpu()` per tensor costs a pageable allocation and a blocking copy each)
    u0_d, K0_d = out["us"][0, 0], out["K"][0, 0]
    u0, K0 = torch.empty_like(u0_d, device="cpu").pin_memory(), torch.empty_like(K0_d, device="cpu").pin_memory()
    stream = torch.cuda.current_stream(dev)
    for k in range(ticks):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        p1.set_refs_window(rows_d, k)                                 # horizon window of the device-resident stream
        if solver == "fddp":
            p1.solve(x, xs, us, N_ITERS, opts, out=out)
        else:
            p1.solve_sqp(x, xs, us, N_ITERS, sqp_opts, out=out)
        u0.copy_(u0_d, non_blocking=True)
        K0.copy_(K0_d, non_blocking=True)
        stream.synchronize()
        ts.append(time.perf_counter() - t0)
        iters.append(int(out["iters"][0]))
        x = p1.integrate(x, out["us"][:, 0], dt)                      # plant = the OCP's integrator
        xs, us = p1.shift_warmstart(out["xs"], out["us"])
        xs[:, 0] = x
    del u0, K0
    ts = np.array(ts[20:]) * 1e3
    track = float(np.abs(x[0, :nv].cpu().numpy() - q[ticks]).max())
    cpu = None
    if solver == "fddp" and with_cpu:
        # the same closed loop on the host: CPU restatement, one thread, then node-parallel calc/calcDiff
        # (ShootingProblem.nthreads, ocp_base_croco.py:62); timed span = horizon window + solve, as on the device
        from oracle import orc

        m = table.to_struct()
        x0h, xs0h, us0h = np.concatenate([q[0], v[0]]), np.concatenate([q[: T + 1], v[: T + 1]], 1), u[:T]
        cpu = {"kind": "port", "what": "CPU restatement of Crocoddyl FDDP (oracle/agx_oracle.cpp), same closed loop, same "
                                       "ticks; wall time of horizon window + solve per tick"}
        cores = len(os.sched_getaffinity(0))
        for name, nt in (("threads_1", 1), ("node_parallel", min(8, cores))):
            ns, it_c, xf = orc.mpc_latency(m, rows, np.full(T, dt), x0h, xs0h, us0h, ticks, N_ITERS, None, nt)
            cpu[name] = {"threads": nt, "p50_ms": float(np.percentile(ns[20:], 50)) * 1e-6,
                         "p99_ms": float(np.percentile(ns[20:], 99)) * 1e-6, "mean_iters": float(it_c[20:].mean()),
                         "final_tracking_error_rad": float(np.abs(xf[:nv] - q[ticks]).max())}
    return {"p50_ms": float(np.percentile(ts, 50)), "p99_ms": float(np.percentile(ts, 99)), "ticks": int(len(ts)),
            "mean_iters": float(np.mean(iters[20:])), "final_tracking_error_rad": track, "cpu_baseline": cpu,
            "workload": "cfg1: B=1, T=20, dt=0.01, sine in configuration space (0.2 rad, 4 s), closed loop with shift "
                        "warm start, <=10 FDDP iterations per tick (eager_exit: a tick is one graph launch whose WHILE node stops at convergence, in FDDP and in CSQP mode); host wall clock of set_refs_window + solve + D2H of us[0], K[0] into pinned host buffers"}


def ocp_class_latency(dev, ticks):
    """The same closed loop through the drop-in OCP class, the way the reference's MPC.run drives an OCP (mpc.py:38-66):
    per tick set_reference_weighted_trajectory(list of T+1 WeightedTrajectoryPoint), solve(x0, list of states, list of
    controls), ocp_results (lists of numpy arrays).  Timed span = those three calls: the host-side packing of the
    reference list, H2D, solve, D2H of the whole result."""
    from agimus_controller_b200 import PANDA_Q_NOMINAL, panda_table
    from agimus_controller_b200.ocp_batched import OCPBatchedFDDP
    from agimus_controller_b200.ocp_interface import (DTFactorsNSeq, OCPParamsBaseCroco, SE3, TrajectoryPoint,
                                                      TrajectoryPointWeights, WeightedTrajectoryPoint)

    T, nv = 20, 7
    yaml_path = os.path.join(ROOT, "tests", "golden", "ocp_goal_reaching.yaml")
    params = OCPParamsBaseCroco(dt=DT, solver_iters=N_ITERS, dt_factor_n_seq=DTFactorsNSeq([1], [T]), horizon_size=T)
    ocp = OCPBatchedFDDP(panda_table(), params, yaml_path, batch_size=1, solver="fddp")
    pose = SE3(np.diag([1.0, -1.0, -1.0]), np.array([0.5, 0.2, 0.5]))

    def wpoint(i):
        q = PANDA_Q_NOMINAL + 0.2 * np.sin(2 * np.pi * i * DT / 4.0) * np.ones(nv)
        return WeightedTrajectoryPoint(
            point=TrajectoryPoint(id=i, time_ns=i * 10_000_000, robot_configuration=q, robot_velocity=np.zeros(nv),
                                  robot_acceleration=np.zeros(nv), robot_effort=np.zeros(nv),
                                  end_effector_poses={"panda_hand_tcp": pose}),
            weights=TrajectoryPointWeights(w_robot_configuration=np.full(nv, 1.0), w_robot_velocity=np.full(nv, 0.1),
                                           w_robot_acceleration=np.zeros(nv), w_robot_effort=np.full(nv, 1e-3),
                                           w_end_effector_poses={"panda_hand_tcp": np.full(6, 0.1)}))

    buffer = [wpoint(i) for i in range(ticks + T + 2)]
    x = np.concatenate([PANDA_Q_NOMINAL, np.zeros(nv)])
    u_grav = ocp.problem.rnea(PANDA_Q_NOMINAL, np.zeros(nv), np.zeros(nv))[0].cpu().numpy()
    xs, us = [x] * (T + 1), [u_grav] * T
    ts = []
    for k in range(ticks):
        t0 = time.perf_counter()
        ocp.set_reference_weighted_trajectory(buffer[k: k + T + 1])
        ocp.solve(x, xs, us)
        res = ocp.ocp_results
        ts.append(time.perf_counter() - t0)
        x = ocp.integrate(x, res.feed_forward_terms[0])
        xs = [x] + list(res.states[2:]) + [res.states[-1]]
        us = list(res.feed_forward_terms[1:]) + [res.feed_forward_terms[-1]]
    ts = np.array(ts[20:]) * 1e3
    return {"p50_ms": float(np.percentile(ts, 50)), "p99_ms": float(np.percentile(ts, 99)), "ticks": int(len(ts)),
            "what": "OCPBatchedFDDP (the OCPBase drop-in) driven as MPC.run drives an OCP: reference list of T+1 "
                    "points packed on the host, solve from Python lists, results back as lists of numpy arrays"}


def pin_to_gpu_numa_node(local):
    """CPU affinity of this process := the cores of the NUMA node GPU `local` hangs off (sysfs); None when unknown."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return {"node": None, "why": f"sysfs reports numa_node {node} for {bdf} (no NUMA topology exposed)"}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"node": node, "why": "none of the node's cores is in this process's affinity mask"}
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cores": len(cpus)}
    except Exception as e:  # no sysfs entry, no permission: report, do not pin
        return {"node": None, "why": f"{type(e).__name__}: {e}"}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from agimus_controller_b200 import _abi
    from agimus_controller_b200.solver import BatchedShootingProblem, probe_fp64_tflops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the solve path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = None
    if world > 1:
        # one rank per GPU: keep the rank (and the pinned host buffers it first-touches) on the NUMA node of its GPU, so
        # that eight ranks do not pull their end-to-end copies through one socket
        numa = pin_to_gpu_numa_node(local)
        dist.init_process_group("nccl", device_id=dev)

    B = B_PER_GPU
    from agimus_controller_b200 import panda_table

    opts = _abi.default_fddp_opts(fixed_iters=True)
    prob = BatchedShootingProblem(panda_table(), np.full(T_NODES, DT), B, device=dev)
    seed = rank + int(os.environ.get("AGX_BENCH_SEED", "0"))  # rank r solves its own slab (seed r)
    w, m = build_workload(B, seed, lambda q, v, a: prob.rnea(q, v, a).cpu().numpy())
    # resident inputs
    refs_d = torch.as_tensor(w["refs"], device=dev)
    x0_d = torch.as_tensor(w["x0"], device=dev)
    xs_d = torch.as_tensor(w["xs_ws"], device=dev)
    us_d = torch.as_tensor(w["us_ws"], device=dev)
    prob.set_refs(refs_d)
    out = prob.alloc_outputs()
    stats = torch.empty(B, 3, dtype=torch.float64, device=dev)
    gathered = torch.empty(world * B, 3, dtype=torch.float64, device=dev) if world > 1 else None
    # Several batches in flight: the sequential kernels of one batch (one warp per problem, or per four) leave SMs idle in
    # their last wave and the forward pass fills less than one wave, so a second, independent batch on its own handle
    # and stream runs in those gaps (measured: 392 k -> 427 k solves/s with two in flight, 434 k with three, 427 k with four; splitting ONE batch into
    # two slabs gains nothing).  Every step is still one complete solve of B problems; step i runs on handle i % IN_FLIGHT.
    IN_FLIGHT = max(1, int(os.environ.get("AGX_IN_FLIGHT", "3")))
    probs, outs = [prob], [out]
    for _ in range(IN_FLIGHT - 1):
        pb = BatchedShootingProblem(panda_table(), np.full(T_NODES, DT), B, device=dev)
        pb.set_refs(refs_d)
        probs.append(pb)
        outs.append(pb.alloc_outputs())
    s_solve = [torch.cuda.Stream(device=dev) for _ in range(IN_FLIGHT)]
    solved_ev = [torch.cuda.Event() for _ in range(IN_FLIGHT)]
    gathered_ev = [None] * IN_FLIGHT
    pipe_state = {"i": 0}

    def step_pipelined():
        i = pipe_state["i"]
        pipe_state["i"] = i + 1
        j = i % IN_FLIGHT
        main_stream = torch.cuda.current_stream()
        if gathered_ev[j] is not None:
            s_solve[j].wait_event(gathered_ev[j])   # this handle's previous results have been gathered
        with torch.cuda.stream(s_solve[j]):
            probs[j].solve(x0_d, xs_d, us_d, N_ITERS, opts, out=outs[j])
            solved_ev[j].record(s_solve[j])
        if world > 1:
            main_stream.wait_event(solved_ev[j])
            stats[:, 0] = outs[j]["cost"]
            stats[:, 1] = outs[j]["iters"]
            stats[:, 2] = outs[j]["status"]
            dist.all_gather_into_tensor(gathered, stats)
            gathered_ev[j] = torch.cuda.Event()
            gathered_ev[j].record(main_stream)

    def drain_pipelined():
        for st_ in s_solve:
            torch.cuda.current_stream().wait_stream(st_)

    solve_events = []

    def step_resident():
        if world > 1 and len(solve_events) < 64:
            # this rank's solve alone (the all_gather below makes every rank's step as long as the slowest slab's)
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            prob.solve(x0_d, xs_d, us_d, N_ITERS, opts, out=out)
            eb.record()
            solve_events.append((ea, eb))
        else:
            prob.solve(x0_d, xs_d, us_d, N_ITERS, opts, out=out)
        if world > 1:
            stats[:, 0] = out["cost"]
            stats[:, 1] = out["iters"]
            stats[:, 2] = out["status"]
            dist.all_gather_into_tensor(gathered, stats)

    # ---- end-to-end arm: HOST buffers in, HOST results out, every step.  The public API is stream-ordered, so the
    # copies of step i+1 (H2D) and of step i-1 (D2H) run on their own streams while step i solves, and IN_FLIGHT solves
    # overlap on their own handles: IN_FLIGHT + 1 slots of device inputs / outputs / pinned results.  Every step still
    # moves all its bytes inside the timed region and the host waits for step (i - IN_FLIGHT)'s results before it issues
    # step i + 1.
    pin = lambda a: torch.as_tensor(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
    h_in = [pin(w[k]) for k in ("refs", "x0", "xs_ws", "us_ws")]
    h2d_bytes = sum(t.numel() * t.element_size() for t in h_in)
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    class Slot:
        def __init__(self, full_K=False):
            self.d_in = [torch.empty_like(t, device=dev) for t in h_in]
            self.out = prob.alloc_outputs()
            self.K_full = torch.empty(self.out["K"].shape, dtype=torch.float64).pin_memory() if full_K else None
            self.res = dict(xs=torch.empty(self.out["xs"].shape, dtype=torch.float64).pin_memory(),
                            us=torch.empty(self.out["us"].shape, dtype=torch.float64).pin_memory(),
                            K0=torch.empty((B,) + tuple(self.out["K"].shape[2:]), dtype=torch.float64).pin_memory(),
                            cost=torch.empty(B, dtype=torch.float64).pin_memory(),
                            iters=torch.empty(B, dtype=torch.int32).pin_memory(),
                            status=torch.empty(B, dtype=torch.int32).pin_memory())
            self.ev_in, self.ev_solved, self.ev_out = (torch.cuda.Event() for _ in range(3))
            self.ev_gathered = None
            self.busy = False

    slots = [Slot() for _ in range(IN_FLIGHT + 1)]
    d2h_bytes = sum(t.numel() * t.element_size() for t in slots[0].res.values())
    e2e_state = {"i": 0}

    def step_e2e():
        i = e2e_state["i"]
        e2e_state["i"] = i + 1
        ns = len(slots)
        sl, prev = slots[i % ns], slots[(i - IN_FLIGHT) % ns]   # ns = IN_FLIGHT + 1 slots
        cur_stream = torch.cuda.current_stream()
        if sl.busy:
            sl.ev_out.synchronize()       # this slot's previous solve and result copy (two steps ago) are complete
        with torch.cuda.stream(s_in):     # H2D of this step's inputs
            for d, h_ in zip(sl.d_in, h_in):
                d.copy_(h_, non_blocking=True)
            sl.ev_in.record(s_in)
        j = i % IN_FLIGHT
        s_solve[j].wait_event(sl.ev_in)
        if sl.ev_gathered is not None:
            s_solve[j].wait_event(sl.ev_gathered)
        with torch.cuda.stream(s_solve[j]):   # step i solves on handle i % IN_FLIGHT while step i-1 still runs on the other
            probs[j].set_refs(sl.d_in[0])
            probs[j].solve(sl.d_in[1], sl.d_in[2], sl.d_in[3], N_ITERS, opts, out=sl.out)
            sl.ev_solved.record(s_solve[j])
        if world > 1:
            cur_stream.wait_event(sl.ev_solved)
            stats[:, 0] = sl.out["cost"]
            stats[:, 1] = sl.out["iters"]
            stats[:, 2] = sl.out["status"]
            dist.all_gather_into_tensor(gathered, stats)
            sl.ev_gathered = torch.cuda.Event()
            sl.ev_gathered.record(cur_stream)
        with torch.cuda.stream(s_out):    # D2H of this step's results
            s_out.wait_event(sl.ev_solved)
            sl.res["xs"].copy_(sl.out["xs"], non_blocking=True)
            sl.res["us"].copy_(sl.out["us"], non_blocking=True)
            if sl.K_full is not None:
                sl.K_full.copy_(sl.out["K"], non_blocking=True)   # every gain matrix, as OCPResults.ricatti_gains
            else:
                sl.res["K0"].copy_(sl.out["K"][:, 0], non_blocking=True)
            sl.res["cost"].copy_(sl.out["cost"], non_blocking=True)
            sl.res["iters"].copy_(sl.out["iters"], non_blocking=True)
            sl.res["status"].copy_(sl.out["status"], non_blocking=True)
            sl.ev_out.record(s_out)
        sl.busy = True
        if i >= IN_FLIGHT and prev.busy:
            prev.ev_out.synchronize()     # the caller reads step (i - IN_FLIGHT)'s results on the host

    def drain_e2e():
        for sl in slots:
            if sl.busy:
                sl.ev_out.synchronize()
        torch.cuda.current_stream().wait_stream(s_out)
        for st_ in s_solve:
            torch.cuda.current_stream().wait_stream(st_)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    per_rank = {}

    def timed(step, steps, finish=None, tag=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            step()
        if finish is not None:
            finish()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            allms = torch.empty(world, dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(allms, ms)
            if tag:
                per_rank[tag] = [float(v) / steps for v in allms.cpu()]
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        elif tag:
            per_rank[tag] = [float(ms.item()) / steps]
        return float(ms.item())

    fp64_peak = probe_fp64_tflops(local, 0.5) if (rank == 0 and not args.no_probe) else None

    for _ in range(max(args.warmup, 1)):
        step_resident()
    torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    # (a) the headline: K steps, IN_FLIGHT batches in flight
    for _ in range(IN_FLIGHT):
        step_pipelined()
    drain_pipelined()
    l0 = sum(p_.launch_count for p_ in probs)
    ms_total = timed(step_pipelined, args.steps, finish=drain_pipelined, tag="resident")
    launches = sum(p_.launch_count for p_ in probs) - l0
    # (b) the same K steps one at a time on one stream, with event pairs around every kernel: per-kernel durations for
    # the roofline (kernels of two batches overlapping would inflate each other's event intervals)
    prob.set_timing(True)
    solve_events.clear()
    ms_serial = timed(step_resident, args.steps, tag="resident_serial")
    if world > 1:
        mine = torch.tensor([float(np.mean([x.elapsed_time(y) for x, y in solve_events])) if solve_events else 0.0],
                            dtype=torch.float64, device=dev)
        allm = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allm, mine)
        per_rank["solve_only"] = [float(v) for v in allm.cpu()]
    phases = prob.get_timing()
    prob.set_timing(False)
    clocks = sampler.stop() if sampler else None

    for _ in range(max(min(args.warmup, 2), 1)):
        step_e2e()
    drain_e2e()
    ms_e2e = timed(step_e2e, args.steps, finish=drain_e2e, tag="e2e")
    # the results that came back are the solver's: same costs as the resident run
    assert torch.equal(slots[0].res["cost"], out["cost"].cpu()), "e2e results differ from the resident run"
    # the same end-to-end step returning EVERY gain matrix (OCPResults.ricatti_gains holds all T of them,
    # ocp_base_croco.py:173-177): + 157 MB of D2H per step
    e2e_full = None
    if not args.no_full_k:
        slots[:] = [Slot(full_K=True) for _ in range(IN_FLIGHT + 1)]
        e2e_state["i"] = 0
        for _ in range(2):
            step_e2e()
        drain_e2e()
        n_fk = max(3, min(args.steps, 10))
        ms_fk = timed(step_e2e, n_fk, finish=drain_e2e)
        d2h_fk = d2h_bytes - slots[0].res["K0"].numel() * 8 + slots[0].K_full.numel() * 8
        assert torch.equal(slots[0].K_full, slots[0].out["K"].cpu())
        e2e_full = {"value": world * B * n_fk / (ms_fk * 1e-3), "unit": "solves/s", "ms_per_step": ms_fk / n_fk,
                    "d2h_bytes_per_step": d2h_fk, "steps": n_fk,
                    "returned": "xs, us, K (all T gain matrices), cost, iters, status"}
        slots[:] = []
        torch.cuda.empty_cache()

    # strong scaling (driver-visible beside the weak-scaling headline): a FIXED total batch split over the ranks in
    # contiguous slabs, inputs resident, same 10 fixed iterations; max over ranks
    strong = None
    if not args.no_strong:
        from agimus_controller_b200.sharding import shard_range

        strong = {"what": "cfg-2 problems, total batch fixed, contiguous slabs per rank, inputs resident, 10 fixed "
                          "FDDP iterations; ms = max over ranks", "cases": []}
        for B_total in (4096, 16384):
            ws, _ = build_workload(B_total, 0, lambda q, v, a: prob.rnea(q, v, a).cpu().numpy()) if B_total != B or world > 1 or seed != 0 else (w, m)
            sl_ = shard_range(B_total, world, rank)
            nb = sl_.stop - sl_.start
            ps = BatchedShootingProblem(panda_table(), np.full(T_NODES, DT), nb, device=dev)
            ps.set_refs(torch.as_tensor(ws["refs"][sl_], device=dev))
            sx0, sxs, sus = (torch.as_tensor(ws[k][sl_], device=dev) for k in ("x0", "xs_ws", "us_ws"))
            so = ps.alloc_outputs()
            for _ in range(2):
                ps.solve(sx0, sxs, sus, N_ITERS, opts, out=so)
            n_s = max(3, min(args.steps, 10))
            ms_s = timed(lambda: ps.solve(sx0, sxs, sus, N_ITERS, opts, out=so), n_s)
            strong["cases"].append({"B_total": B_total, "B_per_gpu": nb, "ms_per_step": ms_s / n_s,
                                    "solves_per_s": B_total * n_s / (ms_s * 1e-3)})
            del ps, so, sx0, sxs, sus
            torch.cuda.empty_cache()

    # single-MPC latency (BASELINE config 1): B = 1, T = 20, dt = 0.01, sine wave in configuration space, closed loop
    # with the shift warm start, <= 10 FDDP iterations per tick (early exit allowed), rank 0 only
    lat = None
    if rank == 0 and not args.no_latency:
        lat = mpc_latency(prob, dev, args.latency_ticks, with_cpu=not args.no_cpu)
        lat_sqp = mpc_latency(prob, dev, args.latency_ticks, solver="csqp")
        lat["csqp_mode"] = {k: lat_sqp[k] for k in ("p50_ms", "p99_ms", "mean_iters", "final_tracking_error_rad")}
        lat["ocp_class"] = ocp_class_latency(dev, min(args.latency_ticks, 300))

    # the reference's own solver mode on the same workload (secondary figure, rank 0, N = 1): SQP = SolverCSQP without
    # active constraints, same budget of 10 iterations, per-problem KKT stop at the reference's tolerance 1e-3
    sqp = None
    if rank == 0 and world == 1 and not args.no_sqp:
        sq_out = prob.alloc_outputs()
        for _ in range(2):
            prob.solve_sqp(x0_d, xs_d, us_d, N_ITERS, None, out=sq_out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_sq = max(3, min(args.steps, 10))
        e0.record()
        for _ in range(n_sq):
            prob.solve_sqp(x0_d, xs_d, us_d, N_ITERS, None, out=sq_out)
        e1.record()
        torch.cuda.synchronize()
        ms_sq = e0.elapsed_time(e1) / n_sq
        # the same with the headline's batches in flight (the handles of the resident leg, one stream each)
        for j in range(IN_FLIGHT):
            with torch.cuda.stream(s_solve[j]):
                probs[j].solve_sqp(x0_d, xs_d, us_d, N_ITERS, None, out=outs[j])
        torch.cuda.synchronize()
        e0.record()
        for i in range(n_sq):
            with torch.cuda.stream(s_solve[i % IN_FLIGHT]):
                probs[i % IN_FLIGHT].solve_sqp(x0_d, xs_d, us_d, N_ITERS, None, out=outs[i % IN_FLIGHT])
        drain_pipelined()
        e1.record()
        torch.cuda.synchronize()
        ms_sqp = e0.elapsed_time(e1) / n_sq
        # the same solver run until its own stop criterion (KKT <= 1e-3) with a budget of 100 iterations, one batch at a time
        torch.cuda.synchronize()
        prob.solve_sqp(x0_d, xs_d, us_d, 100, None, out=sq_out)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            prob.solve_sqp(x0_d, xs_d, us_d, 100, None, out=sq_out)
        e1.record()
        torch.cuda.synchronize()
        ms_sqc = e0.elapsed_time(e1) / 3
        sqp_conv = {"value": B / (ms_sqc * 1e-3), "ms_per_step": ms_sqc, "max_iter": 100,
                    "mean_iters": float(sq_out["iters"].double().mean()), "max_iters": int(sq_out["iters"].max()),
                    "converged_frac": float((sq_out["status"] == 0).double().mean())}
        prob.solve_sqp(x0_d, xs_d, us_d, N_ITERS, None, out=sq_out)   # back to the 10-iteration results for the fields below
        torch.cuda.synchronize()
        sqp = {"value": B / (ms_sqp * 1e-3), "unit": "solves/s", "ms_per_step": ms_sqp, "steps": n_sq,
               "to_convergence": sqp_conv,
               "batches_in_flight": IN_FLIGHT, "serial": {"value": B / (ms_sq * 1e-3), "ms_per_step": ms_sq},
               "same_costs_as_serial": bool(torch.equal(outs[0]["cost"], sq_out["cost"])),
               "max_iter": N_ITERS, "mean_iters": float(sq_out["iters"].double().mean()),
               "converged_frac": float((sq_out["status"] == 0).double().mean()),
               "kkt_median": float(sq_out["stop"].median()),
               "what": "agx_solve_sqp (mim_solvers.SolverCSQP, unconstrained form, termination_tolerance 1e-3) on the "
                       "cfg-2 batch, inputs resident; includes the final sigma sweep for the reported gains"}

    # converged mode ("report also converged-mode"): the same batch solved until every problem stops on
    # SolverFDDP's own criterion (stop < th_stop = 1e-9) or runs out of a 100-iteration budget; one graph launch per
    # solve whose WHILE node ends with the last problem (no polling from the host)
    conv = None
    if rank == 0 and world == 1 and not args.no_sqp:
        copts = _abi.default_fddp_opts()
        c_out = prob.alloc_outputs()
        for _ in range(2):
            prob.solve(x0_d, xs_d, us_d, 100, copts, out=c_out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_c = max(3, min(args.steps, 5))
        e0.record()
        for _ in range(n_c):
            prob.solve(x0_d, xs_d, us_d, 100, copts, out=c_out)
        e1.record()
        torch.cuda.synchronize()
        ms_c = e0.elapsed_time(e1) / n_c
        it_c = c_out["iters"].double()
        # the same with the headline's batches in flight: the long tail of a batch (a few problems still iterating,
        # every round a sequential sweep's latency) runs beside the next batches' full rounds
        for j in range(IN_FLIGHT):
            with torch.cuda.stream(s_solve[j]):
                probs[j].solve(x0_d, xs_d, us_d, 100, copts, out=outs[j])
        torch.cuda.synchronize()
        n_cp = n_c * IN_FLIGHT
        e0.record()
        for i in range(n_cp):
            with torch.cuda.stream(s_solve[i % IN_FLIGHT]):
                probs[i % IN_FLIGHT].solve(x0_d, xs_d, us_d, 100, copts, out=outs[i % IN_FLIGHT])
        drain_pipelined()
        e1.record()
        torch.cuda.synchronize()
        ms_cp = e0.elapsed_time(e1) / n_cp
        conv = {"value": B / (ms_cp * 1e-3), "unit": "solves/s", "ms_per_step": ms_cp, "steps": n_cp, "max_iter": 100,
                "batches_in_flight": IN_FLIGHT, "serial": {"value": B / (ms_c * 1e-3), "ms_per_step": ms_c},
                "same_costs_as_serial": bool(torch.equal(outs[0]["cost"], c_out["cost"])),
                "mean_iters": float(it_c.mean()), "max_iters": int(it_c.max()),
                "converged_frac": float((c_out["status"] == 0).double().mean()),
                "what": "FDDP to convergence (th_stop 1e-9, budget 100 iterations) on the cfg-2 batch, inputs "
                        "resident; a solve lasts as long as its slowest problem"}

    # BASELINE config 4 as stated (secondary figure, rank 0, N = 1): 4096 pick-and-place OCPs, nv = 9 WITH the finger
    # joints (general-tree kernels), T = 100, two capsule-pair collision costs, 3 fixed FDDP iterations; the CPU
    # restatement beside it on a bounded sample
    cfg4 = None
    if rank == 0 and world == 1 and not args.no_cfg4:
        from agimus_controller_b200.workloads import pick_and_place_collision_batch

        t9 = panda_table(lock_fingers=False)
        h9 = BatchedShootingProblem(t9, np.full(2, DT), 1, device=dev)
        w4 = pick_and_place_collision_batch(B, T=100, rnea=lambda q, v, a: h9.rnea(q, v, a).cpu().numpy(),
                                            lock_fingers=False)
        p4 = BatchedShootingProblem(w4["table"], w4["dts"], B, device=dev)
        p4.set_refs(torch.as_tensor(w4["refs"], device=dev))
        a4 = [torch.as_tensor(w4[k], device=dev) for k in ("x0", "xs_ws", "us_ws")]
        o4 = p4.alloc_outputs()
        for _ in range(2):
            p4.solve(*a4, 3, opts, out=o4)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n4 = max(3, min(args.steps, 10))
        e0.record()
        for _ in range(n4):
            p4.solve(*a4, 3, opts, out=o4)
        e1.record()
        torch.cuda.synchronize()
        ms4 = e0.elapsed_time(e1) / n4
        # the same with two batches in flight (SolvePipeline: the second batch runs in the first one's idle waves)
        from agimus_controller_b200.solver import SolvePipeline

        pipe4 = SolvePipeline(w4["table"], w4["dts"], B, n_in_flight=2, device=dev)
        pipe4.set_refs(torch.as_tensor(w4["refs"], device=dev))
        for _ in range(2):
            pipe4.submit(*a4, 3, opts)
        pipe4.join()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n4):
            pipe4.submit(*a4, 3, opts, after_current_stream=False)
        pipe4.join()
        e1.record()
        torch.cuda.synchronize()
        ms4p = e0.elapsed_time(e1) / n4
        same4 = bool(torch.equal(pipe4.problems[0].solve(*a4, 3, opts)["cost"], o4["cost"]))
        pipe4.close()
        cfg4 = {"workload": "cfg4: 4096 pick-and-place OCPs, nv=9 with the finger joints (branching tree, prismatic "
                            "joints; general-tree kernels), T=100, two capsule-pair collision costs (QuadExp), 3 fixed "
                            "FDDP iterations, inputs resident",
                "value": B / (ms4p * 1e-3), "unit": "solves/s", "ms_per_step": ms4p, "steps": n4,
                "batches_in_flight": 2, "same_costs_as_serial": same4,
                "serial": {"value": B / (ms4 * 1e-3), "ms_per_step": ms4},
                "finite": bool(torch.isfinite(o4["xs"]).all())}
        if not args.no_cpu:
            from oracle import orc

            cores = len(os.sched_getaffinity(0))
            n_c = min(B, 32 * cores)
            m4 = w4["table"].to_struct()
            run4 = lambda: orc.solve(m4, w4["refs"][:n_c], w4["dts"], w4["x0"][:n_c], w4["xs_ws"][:n_c],  # noqa: E731
                                     w4["us_ws"][:n_c], 3, opts, nthreads=cores)
            run4()
            t0, reps4 = time.perf_counter(), 0
            while time.perf_counter() - t0 < 5.0 and reps4 < 64:
                run4()
                reps4 += 1
            dt4 = time.perf_counter() - t0
            cfg4["cpu_baseline"] = {"value": n_c * reps4 / dt4, "unit": "solves/s", "cores": cores, "kind": "port",
                                    "sample": f"first {n_c} of the 4096 problems x {reps4} passes ({dt4:.1f} s), OpenMP one "
                                              "problem per thread; CPU restatement of Crocoddyl FDDP, not Crocoddyl itself"}
        del p4, o4, a4
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    solves = world * B * args.steps
    value = solves / (ms_total * 1e-3)
    e2e_value = solves / (ms_e2e * 1e-3)

    # dominant kernel: algorithmic flops per launch / mean launch duration
    T1 = T_NODES + 1
    flops = {"calc_diff": FLOP["calc_diff"] * B * T_NODES, "backward": FLOP["backward"] * B * T_NODES,
             "rollout_try": FLOP["rollout_try"] * B * T_NODES, "node_cost": FLOP["node_cost"] * B * T1,
             "accept_linesearch": 0.0}
    # algorithmic bytes per launch: what the kernel must read and write once
    bytes_alg = {"calc_diff": B * T1 * (REC_BYTES + (14 + 7) * 8),
                 "backward": B * (T1 * (REC_BYTES + CREC_BYTES) + T_NODES * (98 + 7) * 8 + T1 * (14 * 3) * 8),
                 "rollout_try": B * T1 * ((14 * 3 + 7 * 3 + 98) * 8),
                 "node_cost": B * T1 * (CREC_BYTES + (14 + 7 + 62) * 8),
                 "accept_linesearch": B * T1 * 8}
    per = {}
    timing_steps = args.steps
    for k, v in phases.items():
        if v["launches"]:
            # a solve queues one round past its budget for problems whose line search was deferred; that round's
            # launches are all but empty, so the phase time is charged to the N_ITERS launches that do the work
            # (dividing by the launch count would flatter the per-launch figure)
            full = min(v["launches"], N_ITERS * timing_steps) if k != "node_cost" else min(v["launches"], (N_ITERS + 1) * timing_steps)
            mean_ms = v["ms"] / full
            per[k] = {"ms_per_launch": mean_ms, "launches": v["launches"], "share_of_step": v["ms"] / ms_serial,
                      "tflops": flops[k] / (mean_ms * 1e-3) / 1e12, "gbs": bytes_alg[k] / (mean_ms * 1e-3) / 1e9}
    top = max((k for k in per if flops[k] > 0), key=lambda k: per[k]["ms_per_launch"] * per[k]["launches"]) if per else None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (per launch, same command line)
    traffic = None
    try:
        import csv

        rows = list(csv.reader(open(os.path.join(ROOT, "profiles", "r2_ncu_full_raw.csv"))))
        hdr, units = rows[0], rows[1]
        kname = {"backward": "backward_mma_kernel", "node_cost": "node_cost_kernel"}.get(top, (top or "") + "_kernel")
        for row in rows[2:]:
            if kname in row[hdr.index("Kernel Name")]:
                def _bytes(col):
                    v, u = float(row[hdr.index(col)]), units[hdr.index(col)]
                    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
                traffic = _bytes("dram__bytes_read.sum") + _bytes("dram__bytes_write.sum")
                break
    except Exception:
        traffic = None
    roofline = None
    if top:
        roofline = {
            "kernel": {"backward": "backward_mma_kernel" if os.environ.get("AGX_BW", "mma") != "octet" else "backward_kernel",
                       "node_cost": "node_cost_kernel<true>"}.get(top, top + "_kernel"),
            "bound": "fp64", "achieved": per[top]["tflops"], "peak": fp64_peak,
            "unit": "TFLOP/s", "frac": per[top]["tflops"] / fp64_peak if fp64_peak else None, "traffic": traffic,
            "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one launch, from "
                            "profiles/r2_ncu_full_raw.csv; algorithmic bytes per launch: "
                            f"{bytes_alg[top]:.0f}",
            "peak_source": "in-run DFMA probe (agx_probe_fp64); MEASURED_PEAKS.json carries no FP64 figure",
            "hbm": {"achieved": per[top]["gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": per[top]["gbs"] / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"},
            "phases": per,
        }
        # the whole step against the same peak: algorithmic flops of the 10 iterations (+ the first cost pass) over the
        # step time, with the batches in flight (the headline) and one batch at a time
        step_flop = N_ITERS * (flops["calc_diff"] + flops["backward"] + flops["rollout_try"]) + (N_ITERS + 1) * flops["node_cost"]
        if fp64_peak:
            roofline["whole_step"] = {
                "flop_per_step": step_flop,
                "tflops": step_flop / (ms_total / args.steps * 1e-3) / 1e12,
                "frac": step_flop / (ms_total / args.steps * 1e-3) / 1e12 / fp64_peak,
                "serial_tflops": step_flop / (ms_serial / args.steps * 1e-3) / 1e12,
                "serial_frac": step_flop / (ms_serial / args.steps * 1e-3) / 1e12 / fp64_peak}

    # CPU baseline beside it (bounded sample, rank 0, N = 1 only)
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import orc

        cores = len(os.sched_getaffinity(0))
        n = min(B, 48 * cores)
        time_cpu(orc, m, w, min(n, 4 * cores), N_ITERS, threads=cores)
        t = time_cpu(orc, m, w, n, N_ITERS, threads=cores)
        reps = 1
        while t < 8.0 and reps < 16:
            t += time_cpu(orc, m, w, n, N_ITERS, threads=cores)
            reps += 1
        t1 = min(time_cpu(orc, m, w, 1, N_ITERS, threads=1) for _ in range(5))
        cpu = {"value": n * reps / t, "unit": "solves/s", "cores": cores, "kind": "port",
               "single_solve_ms_1thread": 1e3 * t1,
               "sample": f"{n} of the 4096 problems x {reps} passes ({t:.1f} s), OpenMP one problem per thread; "
                         "CPU restatement of Crocoddyl FDDP (oracle/agx_oracle.cpp), not Crocoddyl itself"}

    line = {
        "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "serial": {"value": solves / (ms_serial * 1e-3), "ms_per_step": ms_serial / args.steps,
                   "what": "the same K steps one batch at a time on one stream (the roofline's kernel durations come "
                           "from this pass)"},
        "config": {"workload": WORKLOAD, "B_per_gpu": B, "T": T_NODES, "dt": DT, "fddp_iters": N_ITERS,
                   "batches_in_flight": IN_FLIGHT,
                   "l2": "inputs_larger_than_L2 (node records 481 MB + gains 161 MB per step vs 126 MB L2)",
                   "parallelism": f"independent slabs x{world}, NCCL all_gather of cost/iters/status only"},
        "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e / args.steps,
                "returned": "xs, us, K[:,0] (the gain the controller applies), cost, iters, status",
                "pipelining": "copies of step i+1 (H2D) and i-1 (D2H) overlap the solve of step i on separate streams, "
                              f"step i solves on handle i % {IN_FLIGHT} while the previous ones finish on theirs; the host "
                              f"waits for step i-{IN_FLIGHT}'s results before issuing step i+1"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "latency_b1": lat, "sqp_mode": sqp, "converged_mode": conv, "e2e_full_K": e2e_full, "strong_scaling": strong, "cfg4_nv9": cfg4,
        "ms_per_step_per_rank": per_rank, "rank0_numa": numa,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sample", type=int, default=None, help="problems per step of the reference arm")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-latency", action="store_true", help="skip the B=1 latency leg")
    ap.add_argument("--latency-ticks", type=int, default=1000, help="MPC ticks of the B=1 latency leg")
    ap.add_argument("--no-probe", action="store_true", help="skip the FP64 peak probe (profiler runs)")
    ap.add_argument("--no-sqp", action="store_true", help="skip the SQP-mode leg")
    ap.add_argument("--no-full-k", action="store_true", help="skip the end-to-end leg that returns every gain matrix")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg")
    ap.add_argument("--no-cfg4", action="store_true", help="skip the config-4 (nv = 9, general-tree kernels) leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
