"""A few single-plan host calls (ltp_plan_host, n = 1) for ncu / latency break-down:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/single_plan_probe.py 10"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longtermplanner_b200 import LongTermPlanner, _capi as capi, workloads as W  # noqa: E402

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 200
lim = W.FRANKA7
ltp = LongTermPlanner(lim.dof, lim.t_sample, *lim.arrays(), device=0)
qg, q0, v0, a0 = W.random_states(lim, calls + 20, W.SEEDS[1])
cap = 4096
rows = [np.empty((lim.dof, cap)) for _ in range(4)]
ln, ok, needed = np.zeros(1, np.int32), np.zeros(1, np.uint8), capi.i64(0)
vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
us = []
for k in range(calls + 20):
    ins = [np.ascontiguousarray(x[k]) for x in (qg, q0, v0, a0)]
    t0 = time.perf_counter()
    rc = capi.plan_host(ltp._h, 1, *[vp(x) for x in ins], 0, cap, *[vp(r) for r in rows], vp(ln), vp(ok), C.byref(needed))
    us.append((time.perf_counter() - t0) * 1e6)
    assert rc == 0
us = np.array(us[20:])
print(f"plan_host n=1: median {np.median(us):.1f} us, p10 {np.percentile(us, 10):.1f}, p90 {np.percentile(us, 90):.1f}")
# solve only (no rows): the same call without row buffers returns LTP_ERR_ARG after the solve
us2 = []
for k in range(calls):
    ins = [np.ascontiguousarray(x[k]) for x in (qg, q0, v0, a0)]
    t0 = time.perf_counter()
    capi.plan_host(ltp._h, 1, *[vp(x) for x in ins], 0, cap, None, None, None, None, vp(ln), vp(ok), C.byref(needed))
    us2.append((time.perf_counter() - t0) * 1e6)
print(f"solve part only: median {np.median(us2):.1f} us")
