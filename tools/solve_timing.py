"""Kernel-level timing of the solve (configs[1], 2^20 Franka problems) and of the time-major
sampler (configs[2]) through the library's own event hooks. Used to compare builds:
  LTP_B200_LIB=/path/to/variant.so python tools/solve_timing.py [6|7|12] [rest]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longtermplanner_b200 import LongTermPlanner, workloads as W  # noqa: E402

lim = W.FRANKA7 if len(sys.argv) < 2 or sys.argv[1] != "12" else W.FRANKA12
if len(sys.argv) > 1 and sys.argv[1] == "ref6":  # the reference's test limits: a fifth of the problems needs polynomial roots
    lim = W.REF_RANDOM6
if len(sys.argv) > 1 and sys.argv[1] == "6":  # a six-joint arm: the first six joints of the 7-DoF limits
    lim = W.Limits("franka6", 0.001, *[x[:6] for x in (W.FRANKA7.q_min, W.FRANKA7.q_max, W.FRANKA7.v_max,
                                                       W.FRANKA7.a_max, W.FRANKA7.j_max)])
n = 1 << 20
ltp = LongTermPlanner(lim.dof, lim.t_sample, *lim.arrays(), device=0)
ins = [torch.from_numpy(W.to_joint_major(x)).cuda() for x in W.random_states(lim, n, W.SEEDS[2])]
if len(sys.argv) > 2 and sys.argv[2] == "rest":  # every problem starts at rest: zero numerators
    ins[2].zero_()
    ins[3].zero_()
sol = ltp.alloc_solution(n)
for _ in range(3):
    ltp.solve(*ins, out=sol)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    ltp.solve(*ins, out=sol)
e1.record()
torch.cuda.synchronize()
step_ms = e0.elapsed_time(e1) / 20
ltp.setProfiling(True)
for _ in range(20):
    ltp.solve(*ins, out=sol)
ms, cnt = ltp.kernelTime("solve_fast")
ms2, cnt2 = ltp.kernelTime("solve_generic")
chk = int(sol.traj_len.sum().item())
ms3, cnt3 = ltp.kernelTime("solve_attempt2")
try:
    ms4, cnt4 = ltp.kernelTime("solve_queues")
except Exception:
    ms4, cnt4 = 0.0, 0
try:
    ms5, cnt5 = ltp.kernelTime("solve_items")
except Exception:
    ms5, cnt5 = 0.0, 0
print(f"{os.environ.get('LTP_B200_LIB', 'default')}: dof {lim.dof} step {step_ms:.4f} ms = {n / step_ms / 1e3:.1f} M plans/s; kernel slot0 {ms / 20:.4f} ms + slot4 {ms3 / 20:.4f} ms + slot5 {ms4 / 20:.4f} ms per solve; "
      f"generic {ms2 / 20:.4f} ms; items {ms5 / 20:.4f} ms; traj_len checksum {chk}", flush=True)
if lim.dof == 7:
    n2, H = 4096, 2001
    ins2 = [torch.from_numpy(W.to_joint_major(x)).cuda() for x in W.random_states(lim, n2, W.SEEDS[3])]
    sol2 = ltp.solve(*ins2)
    traj = ltp.alloc_trajectories(n2, H, "time_major")
    ltp.setProfiling(False)
    for _ in range(3):
        ltp.sample(ins2[1], ins2[2], ins2[3], sol2, horizon=H, out=traj)
    ltp.setProfiling(True)
    for _ in range(20):
        ltp.sample(ins2[1], ins2[2], ins2[3], sol2, horizon=H, out=traj)
    ms, cnt = ltp.kernelTime("sample_time_major")
    print(f"   sampler time-major {ms / cnt:.4f} ms -> {n2 * 7 * H * 32 / (ms / cnt) / 1e6:.0f} GB/s", flush=True)
