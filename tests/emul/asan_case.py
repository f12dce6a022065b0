import sys, ctypes as C; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from agimus_controller_b200 import _abi, panda_table
from emul import emu
emu._LIB = _abi.bind(C.CDLL(sys.argv[1]))
from oracle import orc
from agimus_controller_b200.workloads import goal_reaching_batch
m = panda_table().to_struct()
B, T = 5, 7
w = goal_reaching_batch(B, T=T, rnea=lambda q, v, a: orc.rnea(m, q, v, a))
o = emu.solve(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 3, _abi.default_fddp_opts())
emu.calc_diff(m, w["refs"], w["dts"], o["xs"], o["us"]); emu.calc(m, w["refs"], w["dts"], o["xs"], o["us"])
emu.rollout(m, w["refs"], w["dts"], w["x0"], o["us"]); emu.cost_terms(m, w["refs"], w["dts"], o["xs"], o["us"])
emu.shift_warmstart(m, w["refs"], w["dts"], o["xs"], o["us"]); emu.riccati(m, w["refs"][0], w["dts"], w["x0"][0], o["xs"][0], o["us"][0], 1e-6)
# SQP mode (direction / try / accept kernels) and the collision variants of the cost kernels
emu.solve_sqp(m, w["refs"], w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 2)
from agimus_controller_b200.robot_model import PANDA_CAPSULES, PANDA_COLLISION_PAIRS
mc = panda_table().with_capsules(PANDA_CAPSULES, PANDA_COLLISION_PAIRS, alpha=0.02).to_struct()
rc = w["refs"].copy(); rc[..., 60:62] = 10.0
oc = emu.solve(mc, rc, w["dts"], w["x0"], w["xs_ws"], w["us_ws"], 2, _abi.default_fddp_opts())
emu.calc_diff(mc, rc, w["dts"], oc["xs"], oc["us"]); emu.cost_terms(mc, rc, w["dts"], oc["xs"], oc["us"])
# a hostile start makes the line search (deferred and in line) run
wh = goal_reaching_batch(4, T=T, rnea=lambda q, v, a: orc.rnea(m, q, v, a), q_spread=0.8, target_p=(0.3, -0.4, 0.7))
emu.solve(m, wh["refs"], wh["dts"], wh["x0"], wh["xs_ws"], wh["us_ws"], 3, _abi.default_fddp_opts(fixed_iters=True))
print("asan case ok", o["cost"].sum())
