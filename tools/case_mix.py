"""Measured case mix of the solve (SURVEY.md 8d asked for it; BASELINE.md records it): the share of
joints per phase-solve case, per accepted time-scaling attempt and per modified-profile flag, from
the CPU oracle on the seeded workloads of bench.py. TEST/ANALYSIS TOOLING (imports oracle/).
  python tools/case_mix.py [log2n]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longtermplanner_b200 import workloads as W  # noqa: E402
from oracle.bindings import OraclePort, build  # noqa: E402

build()
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 18)
out = {}
for lim, seed in ((W.FRANKA7, W.SEEDS[2]), (W.FRANKA12, W.SEEDS[5]), (W.REF_RANDOM6, W.SEEDS[2])):
    o = OraclePort.from_limits(lim)
    r = o.solve(*W.random_states(lim, n, seed), threads=os.cpu_count())
    oc, tc, fc = r["opt_case"], r["ts_case"], r["final_case"]
    d = {"n": n, "reached": float(r["reached"].mean())}
    d["opt_case_low_nibble"] = {int(k): float(v) / oc.size for k, v in zip(*np.unique(oc & 15, return_counts=True))}
    d["opt_case_flags"] = {f: float(((oc & b) != 0).mean()) for f, b in (("mod", 16), ("both", 32), ("nop2", 64), ("nop6", 128))}
    d["ts_case"] = {int(k): float(v) / tc.size for k, v in zip(*np.unique(tc, return_counts=True))}
    d["final_case_low_nibble"] = {int(k): float(v) / fc.size for k, v in zip(*np.unique(fc & 15, return_counts=True))}
    d["final_mod_share"] = float(((fc & 16) != 0).mean())
    nonslow = tc != 0
    d["final_mod_share_by_ts_case"] = {int(k): float((((fc & 16) != 0) & (tc == k)).sum() / max((tc == k).sum(), 1)) for k in np.unique(tc)}
    prob_root = ((tc >= 3) & (tc <= 9)).any(axis=1) | np.isin(oc & 15, (6, 7, 8)).any(axis=1)
    d["problems_with_root_solve_or_fail"] = float(prob_root.mean())
    d["problems_with_quartic_tail_in_stage1"] = float(np.isin(oc & 15, (6, 7, 8)).any(axis=1).mean())
    d["mean_traj_len"] = float(r["traj_len"].mean())
    out[lim.name] = d
print(json.dumps(out, indent=1))
