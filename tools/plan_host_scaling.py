"""Full planTrajectory (solve + dense sampling, trajectories delivered in HOST memory) for a
batch of n problems per call: ltp_plan_host against the reference's CPU code. Where is the
crossover for a caller who needs the samples on the host?"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longtermplanner_b200 import LongTermPlanner, _capi as capi, workloads as W  # noqa: E402
from oracle.bindings import OraclePort, Reference  # noqa: E402

lim = W.FRANKA7
ltp = LongTermPlanner(lim.dof, lim.t_sample, *lim.arrays(), device=0)
chk = (Reference if Reference.available() else OraclePort).from_limits(lim)
cores = os.cpu_count()
vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
cap = 3200
print(f"{'n':>6s} {'gpu us/plan':>12s} {'gpu pinned':>11s} {'cpu 1 thr':>10s} {'cpu %d thr' % cores:>11s}")
for n in (1, 4, 16, 64, 256, 1024, 4096):
    qg, q0, v0, a0 = W.random_states(lim, n, 77)
    hin = [np.ascontiguousarray(W.to_joint_major(x)) for x in (qg, q0, v0, a0)]
    res = []
    for pinned in (False, True):
        if pinned:
            rows = [torch.empty(n, lim.dof, cap, dtype=torch.float64).pin_memory().numpy() for _ in range(4)]
        else:
            rows = [np.empty((n, lim.dof, cap)) for _ in range(4)]
        ln, ok, needed = np.zeros(n, np.int32), np.zeros(n, np.uint8), capi.i64(0)
        reps = max(3, min(200, 2000 // n))
        for k in range(reps + 2):
            if k == 2:
                t0 = time.perf_counter()
            rc = capi.plan_host(ltp._h, n, *[vp(x) for x in hin], 0, cap, *[vp(r) for r in rows], vp(ln), vp(ok),
                                C.byref(needed))
            assert rc == 0, (rc, needed.value)
        res.append((time.perf_counter() - t0) / reps / n * 1e6)
        del rows
    cpu = []
    for thr in (1, cores):
        m = max(n, 64) if thr == 1 else max(n, 1024)
        g2, s0, sv, sa = W.random_states(lim, m, 77)
        chk.plan_batch(g2[:16], s0[:16], sv[:16], sa[:16], threads=thr)
        t0 = time.perf_counter()
        chk.plan_batch(g2, s0, sv, sa, threads=thr)
        cpu.append((time.perf_counter() - t0) / m * 1e6)
    print(f"{n:6d} {res[0]:12.1f} {res[1]:11.1f} {cpu[0]:10.1f} {cpu[1]:11.1f}", flush=True)
