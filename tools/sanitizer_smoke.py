"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck / initcheck): every kernel
of the library on a few hundred problems, both limit sets (the toy limits exercise the root
solver and the work list). Usage: compute-sanitizer --tool memcheck python tools/sanitizer_smoke.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longtermplanner_b200 import LongTermPlanner, devtools, workloads as W  # noqa: E402

for lim, n in ((W.FRANKA7, 333), (W.REF_RANDOM6, 257), (W.FRANKA12, 100), (W.REF_GRID, 500)):
    ltp = LongTermPlanner(lim.dof, lim.t_sample, *lim.arrays(), device=0)
    ins = devtools.random_states_device(lim, n, 11)
    sol = ltp.solve(*ins, with_opt=True, with_cases=True)
    for layout in ("time_major", "rows"):
        traj = ltp.sample(ins[1], ins[2], ins[3], sol, layout=layout)
        fixed = ltp.sample(ins[1], ins[2], ins[3], sol, horizon=300, layout=layout)
    ltp.advance(fixed if fixed.layout == "time_major" else ltp.sample(ins[1], ins[2], ins[3], sol, horizon=300),
                9, *[t.clone() for t in ins[1:]])
    ltp.planStream(*ins, chunk=128, capacity=max(int(sol.traj_len.max()), 1))
    ltp.setSolveMode(True)
    ltp.solve(*ins)
    ltp.optBrakingBatch(ins[2], ins[3])
    o = ltp.optSwitchTimesBatch(*ins, torch.full_like(ins[0], float(lim.v_max[0])))
    ltp.timeScalingBatch(*ins, o["dir"], (o["t"][6] + 0.2).contiguous())
    torch.cuda.synchronize()
    # the latency path: single plans and single items through the mapped staging block
    from longtermplanner_b200 import Trajectory
    host = [t.cpu().numpy().T.copy() for t in ins]
    ltp.setSolveMode(False)
    for i in range(3):
        ltp.planTrajectory(host[0][i], host[1][i], host[2][i], host[3][i], Trajectory())
        ltp.optBraking(0, float(host[2][i][0]), float(host[3][i][0]))
        ok, t7, d, m = ltp.optSwitchTimes(0, float(host[0][i][0]), float(host[1][i][0]), float(host[2][i][0]),
                                          float(host[3][i][0]), float(lim.v_max[0]))
        ltp.timeScaling(0, float(host[0][i][0]), float(host[1][i][0]), float(host[2][i][0]), float(host[3][i][0]),
                        d, float(t7[6]) + 0.2)
    print(lim.name, "ok", int(sol.reached.sum()), "reached")

# item mode (batches of 8192 problems and more): tail / pending / search kernels, transpose bridge
for lim in (W.REF_RANDOM6, W.REF_GRID, W.FRANKA7):
    n = 8192 + 37
    ltp = LongTermPlanner(lim.dof, lim.t_sample, *lim.arrays(), device=0)
    ins = devtools.random_states_device(lim, n, 13)
    sol = ltp.solve(*ins, with_opt=True, with_cases=True)
    pm = [ltp.transpose(t) for t in ins]
    sol2, traj2, _ = ltp.planEnvs(*pm, horizon=64)
    torch.cuda.synchronize()
    assert torch.equal(sol.t_scaled, sol2.t_scaled)
    print(lim.name, "item mode ok", int((sol.ts_case >= 3).sum()), "joints past the second candidate")
