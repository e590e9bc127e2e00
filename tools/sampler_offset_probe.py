"""Time-major sampler, fixed horizon 2001: does the bandwidth at a given number of environments depend
on where the four output arrays (q, v, a, j) sit relative to each other? One allocation, field k at
k * (field bytes + pad). Usage: python tools/sampler_offset_probe.py [n ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longtermplanner_b200 import LongTermPlanner, devtools, workloads as W  # noqa: E402
from longtermplanner_b200.planner import BatchTrajectories  # noqa: E402

lim, H = W.FRANKA7, 2001
ltp = LongTermPlanner(lim.dof, lim.t_sample, *lim.arrays(), device=0)
PADS = (0, 256, 1024, 4096, 16384, 65536, 65536 + 4096, 1 << 20, (1 << 20) + 65536, 3 << 20, (5 << 20) + 8192)
for n in [int(x) for x in sys.argv[1:]] or (4096, 8192, 16384):
    ins = devtools.random_states_device(lim, n, W.SEEDS[3])
    sol = ltp.solve(*ins)
    field = H * n * lim.dof  # doubles
    for pad in PADS:
        step = field + pad // 8
        big = torch.empty(4 * step + 64, dtype=torch.float64, device="cuda")
        base = (-big.data_ptr() // 8) % 32  # 256-byte aligned start
        f = [big[base + k * step: base + k * step + field].view(H, n, lim.dof) for k in range(4)]
        traj = BatchTrajectories("time_major", H, H, f[0], f[1], f[2], f[3],
                                 torch.empty(n, dtype=torch.uint8, device="cuda"), sol.traj_len)
        for _ in range(3):
            ltp.sample(ins[1], ins[2], ins[3], sol, horizon=H, out=traj)
        ltp.setProfiling(True)
        ltp.kernelTime("sample_time_major")
        for _ in range(10):
            ltp.sample(ins[1], ins[2], ins[3], sol, horizon=H, out=traj)
        ms, cnt = ltp.kernelTime("sample_time_major")
        ltp.setProfiling(False)
        print(f"n = {n:6d} pad {pad:8d} B (field {field * 8} B): {ms / cnt:.4f} ms -> "
              f"{n * 7 * H * 32 / (ms / cnt) / 1e6:6.0f} GB/s", flush=True)
        del traj, f, big
