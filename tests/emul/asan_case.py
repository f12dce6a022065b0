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
# per-cost gradients (masked reference passes) on the chain kernels
emu.cost_derivatives(mc, rc, w["dts"], oc["xs"], oc["us"])
# the general-tree kernels: 9-DoF Panda with fingers and two collision pairs, ragged sizes (B = 3, T = 5): FDDP, SQP,
# calc / calcDiff / rollout / shift / cost terms / per-cost gradients / Riccati
from agimus_controller_b200.workloads import pick_and_place_collision_batch
m9 = panda_table(lock_fingers=False).to_struct()
w9 = pick_and_place_collision_batch(3, T=5, rnea=lambda q, v, a: orc.rnea(m9, q, v, a), alpha=1e-3, w_col=(20.0, 20.0),
                                    lock_fingers=False)
mm9 = w9["table"].to_struct()
o9 = emu.solve(mm9, w9["refs"], w9["dts"], w9["x0"], w9["xs_ws"], w9["us_ws"], 3, _abi.default_fddp_opts())
emu.solve_sqp(mm9, w9["refs"], w9["dts"], w9["x0"], w9["xs_ws"], w9["us_ws"], 2)
emu.calc_diff(mm9, w9["refs"], w9["dts"], o9["xs"], o9["us"]); emu.calc(mm9, w9["refs"], w9["dts"], o9["xs"], o9["us"])
emu.rollout(mm9, w9["refs"], w9["dts"], w9["x0"], o9["us"]); emu.cost_terms(mm9, w9["refs"], w9["dts"], o9["xs"], o9["us"])
emu.shift_warmstart(mm9, w9["refs"], w9["dts"], o9["xs"], o9["us"]); emu.cost_derivatives(mm9, w9["refs"], w9["dts"], o9["xs"], o9["us"])
emu.riccati(mm9, w9["refs"][0], w9["dts"], w9["x0"][0], o9["xs"][0], o9["us"][0], 1e-6)
q9 = o9["xs"][:, 0, :9]; emu.rnea(mm9, q9, q9, q9); emu.integrate(mm9, o9["xs"][:, 0], o9["us"][:, 0], 0.01)
print("asan case ok", o["cost"].sum(), o9["cost"].sum())
