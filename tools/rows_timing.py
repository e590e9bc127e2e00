"""Kernel-level timing of the rows-layout sampler (configs[2]: 4096 x 7 x 2001, fixed horizon; and an
exact-length run) through the library's event hooks.  LTP_B200_LIB=variant.so python tools/rows_timing.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longtermplanner_b200 import LongTermPlanner, workloads as W  # noqa: E402

lim = W.FRANKA7
ltp = LongTermPlanner(lim.dof, lim.t_sample, *lim.arrays(), device=0)
for n2 in (4096, 16384):
    H = 2001
    ins2 = [torch.from_numpy(W.to_joint_major(x)).cuda() for x in W.random_states(lim, n2, W.SEEDS[3])]
    sol2 = ltp.solve(*ins2)
    for layout in ("rows", "time_major"):
        traj = ltp.alloc_trajectories(n2, H, layout)
        for _ in range(3):
            ltp.sample(ins2[1], ins2[2], ins2[3], sol2, horizon=H, out=traj)
        ltp.setProfiling(True)
        ltp.kernelTime("sample_rows"), ltp.kernelTime("sample_time_major")
        for _ in range(10):
            ltp.sample(ins2[1], ins2[2], ins2[3], sol2, horizon=H, out=traj)
        ms, cnt = ltp.kernelTime("sample_rows" if layout == "rows" else "sample_time_major")
        ltp.setProfiling(False)
        print(f"{os.path.basename(os.environ.get('LTP_B200_LIB', 'default'))}: n {n2} {layout:10s} {ms / cnt:.4f} ms -> "
              f"{n2 * 7 * H * 32 / (ms / cnt) / 1e6:.0f} GB/s", flush=True)
        del traj
