"""Latency mode as one graph launch (agx_api.cu: TickGraph, a conditional WHILE node around the FDDP round) against
the stream path that reads the completion flags back after every round: same bits, tick after tick, closed loop."""
import os
import pathlib
import subprocess
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = pathlib.Path(__file__).resolve().parent.parent
CASE = ROOT / "tests" / "gpu_tick_case.py"


def _run(tmp_path, tag, B, ticks, queue, graph, mode="fddp"):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    out = tmp_path / f"{tag}.npz"
    env = dict(os.environ, PYTHONPATH=str(ROOT), AGX_TICK_GRAPH="1" if graph else "0")
    subprocess.run([sys.executable, str(CASE), str(out), str(B), str(ticks), str(int(queue)), mode], check=True, env=env,
                   cwd=str(ROOT), timeout=600)
    return dict(np.load(out))


@pytest.mark.parametrize("B", [1, 8, 64])
def test_tick_graph_gives_the_bits_of_the_stream_path(tmp_path, B):
    ticks = 12
    g = _run(tmp_path, "graph", B, ticks, False, True)
    s = _run(tmp_path, "stream", B, ticks, False, False)
    for k in range(ticks):
        for name in ("iters", "status", "xs", "us", "K", "cost"):
            np.testing.assert_array_equal(g[f"{name}_{k}"], s[f"{name}_{k}"], err_msg=f"{name} tick {k}")
    # the ticks did real work, with different iteration counts over the loop (the WHILE node ran 1..n rounds)
    iters = np.stack([g[f"iters_{k}"] for k in range(ticks)])
    assert iters.max() > 1 and iters.min() >= 1
    assert np.isfinite(g[f"xs_{ticks - 1}"]).all()


@pytest.mark.parametrize("B", [1, 16])
def test_sqp_tick_graph_gives_the_bits_of_the_stream_path(tmp_path, B):
    """agx_solve_sqp in latency mode: nested WHILE nodes (iterations around the line search) against the stream path
    that reads a counter back per step length and the completion flags per iteration."""
    ticks = 10
    g = _run(tmp_path, "graph", B, ticks, False, True, "sqp")
    s = _run(tmp_path, "stream", B, ticks, False, False, "sqp")
    for k in range(ticks):
        for name in ("iters", "status", "xs", "us", "K", "cost"):
            np.testing.assert_array_equal(g[f"{name}_{k}"], s[f"{name}_{k}"], err_msg=f"{name} tick {k}")
    iters = np.stack([g[f"iters_{k}"] for k in range(ticks)])
    assert iters.max() > 1 and np.isfinite(g[f"xs_{ticks - 1}"]).all()


@pytest.mark.parametrize("mode", ["fddp_col", "sqp_col", "fddp_nv9"])
def test_tick_graph_with_collision_costs(tmp_path, mode):
    """The graphs record the collision-pair instantiations of the kernels when the model carries pairs, and the
    general-tree kernels for the nine-joint Panda."""
    ticks = 6
    g = _run(tmp_path, "graph", 4, ticks, False, True, mode)
    s = _run(tmp_path, "stream", 4, ticks, False, False, mode)
    for k in range(ticks):
        for name in ("iters", "status", "xs", "us", "K", "cost"):
            np.testing.assert_array_equal(g[f"{name}_{k}"], s[f"{name}_{k}"], err_msg=f"{name} tick {k}")
    assert np.isfinite(g[f"xs_{ticks - 1}"]).all()


def test_long_budget_to_convergence_runs_as_a_graph_at_any_batch_size(tmp_path):
    """max_iter > 32 without fixed_iters: the stream path polls the device every 16 rounds, the graph loops on the
    device until the last problem stops; 256 problems from a cold start (tens of iterations), then warm ticks."""
    ticks = 3
    g = _run(tmp_path, "graph", 256, ticks, False, True, "fddp_long")
    s = _run(tmp_path, "stream", 256, ticks, False, False, "fddp_long")
    for k in range(ticks):
        for name in ("iters", "status", "xs", "us", "K", "cost"):
            np.testing.assert_array_equal(g[f"{name}_{k}"], s[f"{name}_{k}"], err_msg=f"{name} tick {k}")
    assert g["iters_0"].max() > 10
    assert g["launches"][0] < s["launches"][0]   # the graph counts one round per solve, the stream path every launch


def test_ticks_queued_back_to_back_without_a_synchronisation(tmp_path):
    """The graph path never blocks the host: ticks queued on the stream one after the other (each into its own output
    buffers, each reading the previous one's shifted solution) give the results of the tick-by-tick loop."""
    ticks = 6
    q = _run(tmp_path, "queued", 4, ticks, True, True)
    s = _run(tmp_path, "stream", 4, ticks, False, False)
    for k in range(ticks):
        for name in ("iters", "status", "xs", "us", "K", "cost"):
            np.testing.assert_array_equal(q[f"{name}_{k}"], s[f"{name}_{k}"], err_msg=f"{name} tick {k}")


def test_solve_pipeline_gives_the_bits_of_solving_each_batch_alone():
    """SolvePipeline (independent batches in flight on their own handles and streams): every batch gets the results it
    gets when solved alone."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from agimus_controller_b200 import _abi, panda_table
    from agimus_controller_b200.solver import BatchedShootingProblem, SolvePipeline
    from agimus_controller_b200.workloads import goal_reaching_batch

    B, T = 512, 50
    table = panda_table()
    helper = BatchedShootingProblem(table, np.full(2, 0.01), 1)
    rn = lambda q, v, a: helper.rnea(q, v, a).cpu().numpy()  # noqa: E731
    batches = [goal_reaching_batch(B, T=T, rnea=rn, seed=s) for s in range(5)]
    opts = _abi.default_fddp_opts()
    alone = BatchedShootingProblem(table, batches[0]["dts"], B)
    want = []
    for w in batches:
        alone.set_refs(w["refs"])
        want.append({k: v.cpu().numpy() for k, v in alone.solve(w["x0"], w["xs_ws"], w["us_ws"], 10, opts).items()})
    pipe = SolvePipeline(table, batches[0]["dts"], B, n_in_flight=3)
    dev_in = [{k: torch.as_tensor(w[k], device="cuda") for k in ("refs", "x0", "xs_ws", "us_ws")} for w in batches]
    outs = [pipe.problems[0].alloc_outputs() for _ in batches]
    torch.cuda.synchronize()
    tickets = [pipe.submit(d["x0"], d["xs_ws"], d["us_ws"], 10, opts, refs=d["refs"], out=o)
               for d, o in zip(dev_in, outs)]
    for t, w_ in zip(tickets, want):
        got = t.wait()
        for k in ("iters", "status", "xs", "us", "K", "cost"):
            np.testing.assert_array_equal(got[k].cpu().numpy(), w_[k], err_msg=k)
    assert [t.index for t in tickets] == [0, 1, 2, 0, 1]
    pipe.join()
    pipe.close()


def test_solves_can_be_recorded_into_the_callers_own_graph():
    """A caller that records its stream into a CUDA graph gets the plain stream path of both solvers (kernels only: no
    pointer-table copy, no nested graph launch), and the replay gives the direct call's results."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from agimus_controller_b200 import _abi, panda_table
    from agimus_controller_b200.solver import BatchedShootingProblem
    from agimus_controller_b200.workloads import goal_reaching_batch

    B, T = 128, 30
    table = panda_table()
    helper = BatchedShootingProblem(table, np.full(2, 0.01), 1)
    w = goal_reaching_batch(B, T=T, rnea=lambda q, v, a: helper.rnea(q, v, a).cpu().numpy(), seed=3)
    p = BatchedShootingProblem(table, w["dts"], B)
    p.set_refs(w["refs"])
    x0, xs, us = (torch.as_tensor(w[k], device="cuda") for k in ("x0", "xs_ws", "us_ws"))
    fo = _abi.default_fddp_opts(fixed_iters=True)
    for name, call in (("fddp", lambda out: p.solve(x0, xs, us, 5, fo, out=out)),
                       ("sqp", lambda out: p.solve_sqp(x0, xs, us, 5, None, out=out))):
        want = {k: v.clone() for k, v in call(p.alloc_outputs()).items()}   # also allocates what the call allocates once
        out = p.alloc_outputs()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            call(out)
        for v in out.values():
            v.zero_()
        g.replay()
        torch.cuda.synchronize()
        for k in ("iters", "status", "xs", "us", "K", "cost"):
            assert torch.equal(out[k], want[k]), (name, k)


def test_graph_is_rebuilt_when_the_options_change():
    """One handle, options changing from call to call (the graph is keyed on the iteration budget and the options):
    every call gives what a fresh throughput-mode solve with the same options gives (same decisions, iterates to
    rounding), and returning to the first options reproduces the first result bit for bit."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from agimus_controller_b200 import _abi, panda_table
    from agimus_controller_b200.solver import BatchedShootingProblem
    from agimus_controller_b200.workloads import goal_reaching_batch

    B, T = 4, 20
    table = panda_table()
    helper = BatchedShootingProblem(table, np.full(2, 0.01), 1)
    w = goal_reaching_batch(B, T=T, rnea=lambda q, v, a: helper.rnea(q, v, a).cpu().numpy(), seed=9)
    p = BatchedShootingProblem(table, w["dts"], B)
    p.set_refs(w["refs"])
    ref = BatchedShootingProblem(table, w["dts"], B)
    ref.set_refs(w["refs"])

    def run(th_stop, max_iter, eager):
        o = _abi.default_fddp_opts()
        o.th_stop, o.eager_exit = th_stop, int(eager)
        h = p if eager else ref
        return {k: v.cpu().numpy().copy() for k, v in h.solve(w["x0"], w["xs_ws"], w["us_ws"], max_iter, o).items()}

    cases = [(1e-9, 30), (1e-2, 30), (1e-9, 12), (1e-9, 30)]
    got = [run(th, mi, True) for th, mi in cases]
    for (th, mi), g in zip(cases, got):
        want = run(th, mi, False)
        np.testing.assert_array_equal(g["iters"], want["iters"])
        np.testing.assert_array_equal(g["status"], want["status"])
        for k in ("xs", "us", "cost"):
            assert np.abs(g[k] - want[k]).max() <= 1e-9 * max(np.abs(want[k]).max(), 1.0), k
    assert got[1]["iters"].max() < got[0]["iters"].max()      # the looser stop criterion ends earlier
    assert got[2]["iters"].max() <= 12
    for k in ("iters", "status", "xs", "us", "K", "cost"):
        np.testing.assert_array_equal(got[3][k], got[0][k])
