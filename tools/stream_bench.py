"""BASELINE.json configs[4] on one GPU: N random 12-DoF dual-arm problems with full dense
sampling, streamed through the two-slot ring of ltp_plan_stream (the ~50 TB of samples of the
full 2^26-problem run never exist at once). Usage: python tools/stream_bench.py [log2_n] [chunk]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longtermplanner_b200 import LongTermPlanner, devtools, workloads as W  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
lim = W.FRANKA12
n = 1 << log2n
ltp = LongTermPlanner(lim.dof, lim.t_sample, *lim.arrays(), device=0)
ins = devtools.random_states_device(lim, n, W.SEEDS[5])
torch.cuda.synchronize()
ltp.planStream(*[t[:, :2 * chunk].contiguous() for t in ins], chunk=chunk, capacity=4096)  # warm: allocates the ring
torch.cuda.synchronize()
ltp.setProfiling(True)
t0 = time.perf_counter()
stats = ltp.planStream(*ins, chunk=chunk, capacity=4096)
dt = time.perf_counter() - t0
k_ms, k_cnt = ltp.kernelTime("sample_time_major")
s_ms, s_cnt = ltp.kernelTime("solve_fast")
print(json.dumps({"workload": f"2^{log2n} random 12-DoF problems (FRANKA12), exact-length dense sampling, chunk {chunk}",
                  "seconds": dt, "plans_per_s": n / dt, "write_gbs": stats["bytes"] / dt / 1e9,
                  "sampler_kernel_gbs": stats["bytes"] / (k_ms * 1e-3) / 1e9, "sampler_kernel_ms_total": k_ms,
                  "solve_kernel_ms_total": s_ms, "stats": stats}), flush=True)
