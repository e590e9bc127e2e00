"""Time-major sampler, fixed horizon 2001: GB/s against the number of environments (is the
configs[2] launch, 896 one-warp CTAs, large enough to saturate HBM?)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longtermplanner_b200 import LongTermPlanner, devtools, workloads as W  # noqa: E402

lim, H = W.FRANKA7, 2001
ltp = LongTermPlanner(lim.dof, lim.t_sample, *lim.arrays(), device=0)
for n in [int(x) for x in sys.argv[1:]] or (1024, 2048, 4096, 4736, 8192, 16384, 32768):
    ins = devtools.random_states_device(lim, n, W.SEEDS[3])
    sol = ltp.solve(*ins)
    traj = ltp.alloc_trajectories(n, H, "time_major")
    for _ in range(3):
        ltp.sample(ins[1], ins[2], ins[3], sol, horizon=H, out=traj)
    ltp.setProfiling(True)
    ltp.kernelTime("sample_time_major")
    for _ in range(10):
        ltp.sample(ins[1], ins[2], ins[3], sol, horizon=H, out=traj)
    ms, cnt = ltp.kernelTime("sample_time_major")
    ltp.setProfiling(False)
    print(f"n = {n:6d} ({n * 7 // 32:5d} warps, {n * 7 / 32 / 148:5.1f} per SM): {ms / cnt:.4f} ms -> "
          f"{n * 7 * H * 32 / (ms / cnt) / 1e6:6.0f} GB/s", flush=True)
    del traj
