"""BASELINE.json configs at their stated sizes, through the C ABI, against the CPU oracle:
configs[1] all 2^20 random 7-DoF problems (solve), configs[2] 4096 environments x 7 joints x 2001
samples (both trajectory layouts). Bar: exact fields equal, values within 1e-9 rel / 1e-12 abs."""
import os

import numpy as np
import pytest

from helpers import bitdiff, count_bad, jm, pm
from longtermplanner_b200 import workloads as W
from oracle.bindings import OraclePort

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
CORES = os.cpu_count() or 1


def _planner(lim):
    from longtermplanner_b200 import LongTermPlanner
    return LongTermPlanner(lim.dof, lim.t_sample, *lim.arrays(), device=0)


def test_config1_all_2_pow_20_problems_match_the_oracle():
    lim, n = W.FRANKA7, 1 << 20
    qg, q0, v0, a0 = W.random_states(lim, n, W.SEEDS[2])
    ltp = _planner(lim)
    ins = [torch.from_numpy(jm(x)).cuda() for x in (qg, q0, v0, a0)]
    sol = ltp.solve(*ins, with_opt=True, with_cases=True)
    torch.cuda.synchronize()
    ref = OraclePort.from_limits(lim).solve(qg, q0, v0, a0, threads=CORES)
    for k in ("reached", "slowest", "traj_len"):
        assert np.array_equal(getattr(sol, k).cpu().numpy(), ref[k]), k
    for k in ("mod", "opt_case", "ts_case", "final_case", "dir"):
        assert np.array_equal(pm(getattr(sol, k).cpu().numpy()), ref[k]), k
    for k in ("t_opt", "t_scaled", "v_drive"):
        got = pm(getattr(sol, k).cpu().numpy())
        assert count_bad(got, ref[k]) == 0, (k, bitdiff(got, ref[k]))
        assert bitdiff(got, ref[k]) < 1e-4 * got.size, k   # last-bit differences of pow(x, 3|4) only
    assert ref["reached"].all() and len(np.unique(ref["ts_case"])) >= 4


@pytest.mark.parametrize("layout", ["time_major", "rows"])
def test_config2_4096_envs_2001_samples_match_the_oracle(layout):
    lim, n, H = W.FRANKA7, 4096, 2001
    qg, q0, v0, a0 = W.random_states(lim, n, W.SEEDS[3])
    ltp = _planner(lim)
    ins = [torch.from_numpy(jm(x)).cuda() for x in (qg, q0, v0, a0)]
    sol = ltp.solve(*ins)
    traj = ltp.sample(ins[1], ins[2], ins[3], sol, horizon=H, layout=layout)
    torch.cuda.synchronize()
    P = OraclePort.from_limits(lim)
    tl = sol.traj_len.cpu().numpy()
    succ = traj.success.cpu().numpy()
    ref = P.plan_batch(qg, q0, v0, a0, threads=CORES)          # every environment: length and final check
    assert np.array_equal(ref["length"], tl)
    assert np.array_equal(ref["success"].astype(np.uint8), succ)
    # every sample of every 8th environment against the oracle's own sampler
    picks = np.arange(0, n, 8)
    idx = torch.from_numpy(picks).cuda()
    got = {}
    for k in "qvaj":
        x = getattr(traj, k)
        x = x[:, idx, :].permute(1, 2, 0) if layout == "time_major" else x[idx, :, :H]
        got[k] = x.cpu().numpy()                                 # [len(picks), dof, H]
    bits = 0
    for c, i in enumerate(picks):
        full = P.plan(qg[i], q0[i], v0[i], a0[i])
        m = min(full["length"], H)
        for k in "qvaj":
            assert count_bad(got[k][c, :, :m], full[k][:, :m]) == 0, (k, i)
            bits += bitdiff(got[k][c, :, :m], full[k][:, :m])
            if k != "q" and H > m + 1:                           # past the end: the recurrence's steady state
                assert not got[k][c, :, m + 1:].any(), (k, i)
    assert bits < 1e-3 * len(picks) * lim.dof * H * 4
