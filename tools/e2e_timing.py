"""End-to-end solve through ltp_solve_host (pinned host buffers, 2^20 FRANKA7 problems): ms per call, for
A/B of the host pipeline.  LTP_B200_LIB=variant.so python tools/e2e_timing.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longtermplanner_b200 import LongTermPlanner, workloads as W  # noqa: E402

lim, n = W.FRANKA7, 1 << 20
ltp = LongTermPlanner(lim.dof, lim.t_sample, *lim.arrays(), device=0)
host_in = [torch.from_numpy(W.to_joint_major(x)).pin_memory().numpy() for x in W.random_states(lim, n, W.SEEDS[2])]
out = {k: torch.empty(s, dtype=d).pin_memory().numpy() for k, s, d in (
    ("records", (lim.dof, n, 8), torch.float64), ("dir", (lim.dof, n), torch.float64),
    ("mod", (lim.dof, n), torch.uint8),
    ("slowest", (n,), torch.int32), ("traj_len", (n,), torch.int32), ("reached", (n,), torch.uint8))}
for _ in range(3):
    ltp.solve_host(*host_in, out=out)
ts = []
for _ in range(10):
    t0 = time.perf_counter()
    ltp.solve_host(*host_in, out=out)
    ts.append((time.perf_counter() - t0) * 1e3)
ts.sort()
print(f"{os.path.basename(os.environ.get('LTP_B200_LIB', 'default'))}: median {ts[5]:.3f} ms, best {ts[0]:.3f} ms -> "
      f"{n / ts[5] / 1e3:.1f} M plans/s; checksum {int(out['traj_len'].sum())}", flush=True)
