"""Generates tests/golden/ref_vectors.npz by running THE REFERENCE'S OWN CODE
(oracle/_ref/libltp_ref.so = /root/reference/src/long_term_planner.cc, unmodified, built
by oracle/Makefile against the Eigen shim) on seeded synthetic inputs.

Run in the build container (needs /root/reference to build oracle/_ref):
    python tests/golden/make_reference_vectors.py
The fixture lets the CPU suite and the GPU suite check the oracle port and the CUDA path
against reference outputs without /root/reference being present.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from longtermplanner_b200 import workloads as W  # noqa: E402
from oracle.bindings import Reference, build  # noqa: E402


def main():
    build()
    out = {}
    for tag, lim, n, seed, n_traj in (("franka7", W.FRANKA7, 192, 0x601D01, 2),
                                      ("franka12", W.FRANKA12, 64, 0x601D02, 1),
                                      ("random6", W.REF_RANDOM6, 256, 0x601D03, 3)):
        R = Reference.from_limits(lim)
        qg, q0, v0, a0 = W.random_states(lim, n, seed)
        s = R.solve(qg, q0, v0, a0)
        out[f"{tag}_seed"] = np.int64(seed)
        for k in ("t_opt", "t_scaled", "dir", "v_drive", "mod", "slowest", "reached"):
            out[f"{tag}_{k}"] = s[k]
        pb = R.plan_batch(qg, q0, v0, a0)
        out[f"{tag}_success"] = pb["success"]
        out[f"{tag}_length"] = pb["length"]
        for i in range(n_traj):
            full = R.plan(qg[i], q0[i], v0[i], a0[i])
            for k in "qvaj":
                out[f"{tag}_traj{i}_{k}"] = full[k]
    # the reference's time-scaling grid, every 97th point, at +0.2 s and +1.0 s
    lim = W.REF_GRID
    R = Reference.from_limits(lim)
    qg, v0, a0 = W.reference_grid_points(True)
    sel = np.arange(0, len(qg), 97)
    qg, v0, a0 = qg[sel], v0[sel], a0[sel]
    q0 = np.full_like(qg, 0.5)
    o = R.opt_switch_times(qg, q0, v0, a0, np.full_like(qg, 1.0))
    out["grid_sel_stride"] = np.int64(97)
    for k in ("t", "dir", "mod", "ok"):
        out[f"grid_ost_{k}"] = o[k]
    for inc in (0.2, 1.0):
        ts = R.time_scaling(qg, q0, v0, a0, o["dir"], o["t"][:, 6] + inc)
        for k in ("t", "v_drive", "mod", "ok"):
            out[f"grid_ts{inc}_{k}"] = ts[k]
    path = os.path.join(ROOT, "tests", "golden", "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", R.build_info())


if __name__ == "__main__":
    main()
