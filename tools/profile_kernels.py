"""Short driver for ncu: a few launches of the solve kernels (configs[1], 2^20 problems) and
of the sampler (configs[2], 4096 x 7 x 2001). Usage (on the GPU box, one GPU):
  ncu --set full --clock-control none --import-source on -k regex:ltp_ -o gpurun_out/prof python tools/profile_kernels.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longtermplanner_b200 import LongTermPlanner, workloads as W  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
lim = W.FRANKA7
ltp = LongTermPlanner(lim.dof, lim.t_sample, *lim.arrays(), device=0)
n = 1 << 20
ins = [torch.from_numpy(W.to_joint_major(x)).cuda() for x in W.random_states(lim, n, W.SEEDS[2])]
sol = ltp.alloc_solution(n)
for _ in range(reps):
    ltp.solve(*ins, out=sol)
n2, H = 4096, 2001
ins2 = [torch.from_numpy(W.to_joint_major(x)).cuda() for x in W.random_states(lim, n2, W.SEEDS[3])]
sol2 = ltp.alloc_solution(n2)
for layout in ("time_major", "rows"):
    traj = ltp.alloc_trajectories(n2, H, layout)
    for _ in range(reps):
        ltp.solve(*ins2, out=sol2)
        ltp.sample(ins2[1], ins2[2], ins2[3], sol2, horizon=H, out=traj)
torch.cuda.synchronize()
print("done")
