"""Root-finder cross-check (the Eigen boundary): every polynomial the solve hands to the
root finder on a workload is traced out of the C restatement (oracle/ltp_oracle.c, whose QR is the same
algorithm as the device's and the Eigen shim's, bit for bit) and solved again by LAPACK (numpy.linalg.eigvals
of the same companion matrix, roots.h:28-31). Reported: how often the two disagree on which roots are real
(roots.h:47 tests imag == 0 exactly), on whether a usable root exists, and on the chosen root
(smallest real root > 1e-7, roots.h:43-50). TEST/ANALYSIS TOOLING (imports oracle/).
  python tools/root_crosscheck.py [--json profiles/r02_root_crosscheck.json]"""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longtermplanner_b200 import workloads as W  # noqa: E402
from oracle.bindings import OraclePort, build  # noqa: E402


def traced_polynomials(lim, states, capacity=4_000_000, run=None):
    """-> (records [m, 8], total count) of the polynomials a single-threaded solve of `states`
    (or run(P), any single-threaded sequence of oracle calls) root-solves"""
    P = OraclePort.from_limits(lim)
    buf = np.zeros((capacity, 8))
    P.lib.ltpo_trace_roots.restype = None
    P.lib.ltpo_trace_count.restype = C.c_int64
    P.lib.ltpo_trace_roots(buf.ctypes.data_as(C.c_void_p), C.c_int64(capacity))
    try:
        if run is not None:
            run(P)
        else:
            P.solve(*states, threads=1)
        total = int(P.lib.ltpo_trace_count())
    finally:
        P.lib.ltpo_trace_roots(None, C.c_int64(0))
    return buf[:min(total, capacity)], total, P


def lapack_choice(rec):
    deg = int(rec[0])
    c = rec[1:2 + deg]
    comp = np.zeros((deg, deg))
    for i in range(deg - 1):
        comp[i + 1, i] = 1.0
    comp[:, deg - 1] = (-1.0 * c[::-1][:deg]) / c[0]
    if not np.all(np.isfinite(comp)):
        return None, None
    ev = np.linalg.eigvals(comp)
    real = ev[(ev.imag == 0) & (ev.real > 1e-7)].real
    return (real.min() if real.size else np.inf), ev


def crosscheck(lim, states, name, run=None):
    recs, total, P = traced_polynomials(lim, states, run=run)
    out = {"workload": name, "polynomials": total, "compared": len(recs), "by_degree": {},
           "non_finite_companion": 0, "existence_disagree": 0, "chosen_root_rel_gt_1e-9": 0,
           "chosen_root_rel_gt_1e-6": 0, "real_count_disagree": 0, "worst_rel": 0.0}
    for rec in recs:
        deg = int(rec[0])
        out["by_degree"][deg] = out["by_degree"].get(deg, 0) + 1
        want, ev = lapack_choice(rec)
        re, im = np.zeros(deg), np.zeros(deg)
        got = P.lib.ltpo_roots
        got.restype = C.c_double
        mine = got(rec[1:].ctypes.data_as(C.c_void_p), C.c_int(deg), re.ctypes.data_as(C.c_void_p),
                   im.ctypes.data_as(C.c_void_p))
        if want is None:
            out["non_finite_companion"] += 1
            continue
        if int((im == 0).sum()) != int((ev.imag == 0).sum()):
            out["real_count_disagree"] += 1
        if np.isinf(mine) != np.isinf(want):
            out["existence_disagree"] += 1
            continue
        if np.isinf(mine):
            continue
        rel = abs(mine - want) / max(abs(want), 1e-300)
        out["worst_rel"] = max(out["worst_rel"], float(rel))
        out["chosen_root_rel_gt_1e-9"] += int(rel > 1e-9)
        out["chosen_root_rel_gt_1e-6"] += int(rel > 1e-6)
    return out


def grid_states(m=40):
    """the REF_GRID sweep shape (tests.cc:345-363 refined): 1 joint, q_0 = 0.5"""
    lim = W.REF_GRID
    e = 1e-6
    qg = np.linspace(-6, 7, m)
    v0 = np.linspace(-(1 - e), 1 - e, m)
    u = np.linspace(0, 1, m)
    G, V, U = np.meshgrid(qg, v0, u, indexing="ij")
    a_max, j_max, v_max = lim.a_max[0], lim.j_max[0], lim.v_max[0]
    root = np.sqrt(2.0 * j_max * (v_max - np.abs(V)))
    pos = V >= 0
    a_lb = np.where(pos, -(a_max - e), np.maximum(-(a_max - e), -root))
    a_ub = np.where(pos, np.minimum(a_max - e, root), a_max)
    A = a_lb + U * (a_ub - a_lb)
    n = G.size
    return lim, (G.reshape(n, 1), np.full((n, 1), 0.5), V.reshape(n, 1), A.reshape(n, 1))


if __name__ == "__main__":
    build()
    res = []
    lim, st = grid_states(48)
    res.append(crosscheck(lim, st, "REF_GRID 48^3 single-joint grid (time-optimal solves: quartic tails)"))
    res.append(crosscheck(W.REF_RANDOM6, W.random_states(W.REF_RANDOM6, 60000, W.SEEDS[2]),
                          "REF_RANDOM6, 60000 random 6-DoF problems (stage-1 tails + TS3..TS8)"))

    # a one-joint limit set under which the quintic and the sextic candidate are ACCEPTED (the
    # reference's own limits only ever reject them): tests/test_gpu_parity.py uses the same set
    lim1 = W.random_limits(1, 1001)
    qg, q0, v0, a0 = (x[:, 0].copy() for x in W.random_states(lim1, 100000, 78))

    def search(P):
        o = P.opt_switch_times(qg, q0, v0, a0, np.full(qg.size, lim1.v_max[0]), threads=1)
        for inc in (0.02, 0.05, 0.2):
            P.time_scaling(qg, q0, v0, a0, o["dir"], o["t"][:, 6] + inc, threads=1)
    res.append(crosscheck(lim1, None, "random one-joint limit set 1001: searches at t_opt + {0.02, 0.05, 0.2} "
                                      "(every candidate 3..8 accepted somewhere)", run=search))
    print(json.dumps(res, indent=1))
    if len(sys.argv) > 2 and sys.argv[1] == "--json":
        json.dump(res, open(sys.argv[2], "w"), indent=1)
