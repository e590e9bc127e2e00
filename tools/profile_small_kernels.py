"""Driver for ncu: the latency-path kernels (single plan) and the ordering kernels of the sorted-slot
streaming mode.  ncu --set full --clock-control none -k regex:"row_latency|order_|generic" -o ... python tools/profile_small_kernels.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longtermplanner_b200 import LongTermPlanner, Trajectory, devtools, workloads as W  # noqa: E402

lim = W.FRANKA7
ltp = LongTermPlanner(lim.dof, lim.t_sample, *lim.arrays(), device=0)
qg, q0, v0, a0 = W.random_states(lim, 4, W.SEEDS[1])
for i in range(3):
    ltp.planTrajectory(qg[i], q0[i], v0[i], a0[i], Trajectory())
lim12 = W.FRANKA12
ltp12 = LongTermPlanner(lim12.dof, lim12.t_sample, *lim12.arrays(), device=0)
ins = devtools.random_states_device(lim12, 8192, W.SEEDS[5])
ltp12.planStream(*ins, chunk=8192, capacity=4096, sorted_slots=True)
torch.cuda.synchronize()
print("done")
