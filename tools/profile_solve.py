"""Short ncu driver: one warm + one captured pass of the solve kernels on 2^20 FRANKA7 problems.
  ncu --set full --clock-control none --import-source on -k regex:ltp_solve -c 6 -o gpurun_out/x python tools/profile_solve.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longtermplanner_b200 import LongTermPlanner, workloads as W  # noqa: E402

lim = {"7": W.FRANKA7, "12": W.FRANKA12, "ref6": W.REF_RANDOM6}[sys.argv[1] if len(sys.argv) > 1 else "7"]
ltp = LongTermPlanner(lim.dof, lim.t_sample, *lim.arrays(), device=0)
n = 1 << 20
ins = [torch.from_numpy(W.to_joint_major(x)).cuda() for x in W.random_states(lim, n, W.SEEDS[2])]
sol = ltp.alloc_solution(n)
for _ in range(2):
    ltp.solve(*ins, out=sol)
torch.cuda.synchronize()
print("done")
