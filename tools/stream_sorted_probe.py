"""Does the exact-length sampler lose its 5 % to ragged rows inside a warp? Streams the same
2^k random 12-DoF problems (configs[4] shape) three times: as generated, sorted by trajectory
length inside every chunk, sorted globally. Coalescing is identical in all three (the problem
order itself is permuted); only the spread of row lengths inside a warp changes."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longtermplanner_b200 import LongTermPlanner, devtools, workloads as W  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 21
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
lim = W.FRANKA7 if len(sys.argv) > 3 and sys.argv[3] == "7" else W.FRANKA12
n = 1 << log2n
ltp = LongTermPlanner(lim.dof, lim.t_sample, *lim.arrays(), device=0)
ins = devtools.random_states_device(lim, n, W.SEEDS[5])
tl = ltp.solve(*ins).traj_len.long()
order_chunk = torch.cat([c * chunk + torch.argsort(tl[c * chunk:(c + 1) * chunk]) for c in range(n // chunk)])
order_all = torch.argsort(tl)
ltp.planStream(*[t[:, :2 * chunk].contiguous() for t in ins], chunk=chunk, capacity=4096)
for name, order in (("as generated", None), ("sorted inside each chunk", order_chunk), ("sorted globally", order_all)):
    x = ins if order is None else [t[:, order].contiguous() for t in ins]
    torch.cuda.synchronize()
    ltp.setProfiling(True)
    ltp.kernelTime("sample_time_major")
    t0 = time.perf_counter()
    stats = ltp.planStream(*x, chunk=chunk, capacity=4096)
    dt = time.perf_counter() - t0
    k_ms, _ = ltp.kernelTime("sample_time_major")
    ltp.setProfiling(False)
    print(f"{name:26s}: {stats['bytes'] / dt / 1e9:7.0f} GB/s whole run, sampler kernels alone "
          f"{stats['bytes'] / (k_ms * 1e-3) / 1e9:7.0f} GB/s, {dt:.3f} s", flush=True)
