"""GPU tests of the device-side workload tooling and of the size-independent properties used
at BASELINE.json's full sizes: generator == numpy recipe bit for bit, per-row checksums of
sampled trajectories == the oracle's, limits respected and goal reached on every row."""
import numpy as np
import pytest

from helpers import count_bad, jm
from longtermplanner_b200 import workloads as W
from oracle.bindings import OraclePort

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _planner(lim):
    from longtermplanner_b200 import LongTermPlanner
    return LongTermPlanner(lim.dof, lim.t_sample, *lim.arrays(), device=0)


@pytest.mark.parametrize("lim,n,start,seed", [(W.FRANKA7, 10_000, 0, W.SEEDS[2]), (W.FRANKA12, 4097, 123_456_789, W.SEEDS[5]),
                                              (W.REF_GRID, 1000, 7, 3)])
def test_device_generator_equals_numpy_recipe(lim, n, start, seed):
    from longtermplanner_b200 import devtools
    dev = devtools.random_states_device(lim, n, seed, start=start)
    torch.cuda.synchronize()
    host = W.random_states(lim, n, seed, start=start)
    for d, h in zip(dev, host):
        assert np.array_equal(d.cpu().numpy(), jm(h))


@pytest.mark.parametrize("layout", ["time_major", "rows"])
def test_row_checksums_match_oracle(layout):
    """checksum of checksums: sequential per-row sums of q, v, a, j over the exact length"""
    from longtermplanner_b200 import devtools
    lim, n = W.FRANKA12, 48
    qg, q0, v0, a0 = W.random_states(lim, n, W.SEEDS[5])
    ltp = _planner(lim)
    ins = [torch.from_numpy(jm(x)).cuda() for x in (qg, q0, v0, a0)]
    sol, traj = ltp.planTrajectories(*ins, layout=layout)
    st = devtools.row_stats(traj, sol.traj_len).cpu().numpy()
    P = OraclePort.from_limits(lim)
    for i in range(n):
        full = P.plan(qg[i], q0[i], v0[i], a0[i])
        for c, k in enumerate("qvaj"):
            ref = np.zeros(lim.dof)
            for s in range(full["length"]):  # same order as the device reducer
                ref = ref + full[k][:, s]
            assert count_bad(st[i, :, c], ref) == 0, (i, k)
        assert count_bad(st[i, :, 6], full["q"][:, -1]) == 0


def test_exact_length_rows_are_clipped_to_the_capacity():
    """horizon = 0 with a capacity below traj_len: rows are clipped, nothing is written past the
    capacity, success still refers to the complete trajectory"""
    lim, n = W.FRANKA7, 64
    qg, q0, v0, a0 = W.random_states(lim, n, 5)
    ltp = _planner(lim)
    ins = [torch.from_numpy(jm(x)).cuda() for x in (qg, q0, v0, a0)]
    sol = ltp.solve(*ins)
    full = ltp.sample(ins[1], ins[2], ins[3], sol)
    tl = sol.traj_len.cpu().numpy()
    cap = int(np.median(tl))
    for layout in ("time_major", "rows"):
        ref = ltp.sample(ins[1], ins[2], ins[3], sol, layout=layout)
        if layout == "time_major":
            # addressing does not depend on the capacity: 8 guard samples behind it
            out = ltp.alloc_trajectories(n, cap + 8, layout)
            for t in (out.q, out.v, out.a, out.j):
                t.fill_(-7.0)
            out.stride = cap
        else:
            out = ltp.alloc_trajectories(n, cap, layout)  # capacity == row stride
        cap_eff = out.stride
        ltp.sample(ins[1], ins[2], ins[3], sol, out=out)
        torch.cuda.synchronize()
        assert torch.equal(out.success, full.success)
        for k in "qvaj":
            g, r = getattr(out, k).cpu().numpy(), getattr(ref, k).cpu().numpy()
            for i in range(n):
                m = min(tl[i], cap_eff)
                if layout == "time_major":
                    assert np.array_equal(g[:m, i, :], r[:m, i, :])
                else:
                    assert np.array_equal(g[i, :, :m], r[i, :, :m])
            if layout == "time_major":
                assert (g[cap:] == -7.0).all()  # guard samples untouched


@pytest.mark.parametrize("lim,n,seed", [(W.FRANKA12, 4096, W.SEEDS[5]), (W.FRANKA7, 4096, W.SEEDS[3])])
def test_domain_properties_on_a_generated_chunk(lim, n, seed):
    """what the streamed full-size runs check per chunk: every plan is reached, every row stays
    inside v_max / a_max (up to the discretisation overshoot of the forward-Euler recurrence:
    about two samples' worth on the oracle, four / two allowed here), ends at rest and
    within 0.02 rad of the goal (the reference's own grid-test bar, tests.cc:318)"""
    from longtermplanner_b200 import devtools
    ltp = _planner(lim)
    ins = devtools.random_states_device(lim, n, seed, start=10 * n)
    sol, traj = ltp.planTrajectories(*ins)
    st = devtools.row_stats(traj, sol.traj_len)
    torch.cuda.synchronize()
    assert bool(sol.reached.all()) and bool(traj.success.all())
    q_min, q_max, v_max, a_max, j_max = (torch.from_numpy(x).cuda() for x in lim.arrays())
    ts = lim.t_sample
    assert bool((st[:, :, 4] <= v_max + 4 * a_max * ts).all())
    assert bool((st[:, :, 5] <= a_max + 2 * j_max * ts).all())
    assert bool((st[:, :, 7] == 0).all())  # v pinned to 0 after the last switching time
    err = (st[:, :, 6] - ins[0].T).abs()
    assert float(err.max()) < 0.02, float(err.max())
