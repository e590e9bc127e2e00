"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on
the same seeded inputs. Exact: reached/success, case ids, slowest joint, trajectory
length, dir, mod. Numeric: 1e-9 rel / 1e-12 abs on times, v_drive and samples."""
import numpy as np
import pytest

from helpers import bitdiff, close, count_bad, jm, pm
from longtermplanner_b200 import workloads as W
from oracle.bindings import OraclePort, Reference

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _planner(lim):
    from longtermplanner_b200 import LongTermPlanner
    return LongTermPlanner(lim.dof, lim.t_sample, *lim.arrays(), device=0)


def _dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _solve_both(lim, n, seed):
    qg, q0, v0, a0 = W.random_states(lim, n, seed)
    ltp = _planner(lim)
    ins = [_dev(jm(x)) for x in (qg, q0, v0, a0)]
    sol = ltp.solve(*ins, with_opt=True, with_cases=True)
    torch.cuda.synchronize()
    ref = OraclePort.from_limits(lim).solve(qg, q0, v0, a0, threads=8)
    return ltp, ins, sol, ref, (qg, q0, v0, a0)


# a cruise speed so low that the braking half of the time-optimal solve needs the cc:153 fix-up in
# joint 0 (v_max / a_max < a_max / j_max): the half prepared on the host is then not usable and the
# closed-form kernel evaluates it per thread
_SLOW3 = W.Limits("slow_cruise", 0.004, (-3.0,) * 3, (3.0,) * 3, (0.05, 0.4, 1.0), (2.0, 2.0, 2.0), (4.0, 4.0, 4.0))


@pytest.mark.parametrize("lim,n,seed", [(W.FRANKA7, 100_000, W.SEEDS[2]), (W.FRANKA12, 30_000, W.SEEDS[5]),
                                        (W.REF_RANDOM6, 60_000, 7), (W.FRANKA7, 1, 3), (W.FRANKA7, 33, 4)])
def test_solve_matches_oracle(lim, n, seed):
    ltp, ins, sol, ref, _ = _solve_both(lim, n, seed)
    assert np.array_equal(sol.reached.cpu().numpy(), ref["reached"])
    assert np.array_equal(sol.slowest.cpu().numpy(), ref["slowest"])
    assert np.array_equal(sol.traj_len.cpu().numpy(), ref["traj_len"])
    for k in ("mod", "opt_case", "ts_case", "final_case"):
        assert np.array_equal(pm(getattr(sol, k).cpu().numpy()), ref[k]), k
    assert np.array_equal(pm(sol.dir.cpu().numpy()), ref["dir"])
    for k in ("t_opt", "t_scaled", "v_drive"):
        got = pm(getattr(sol, k).cpu().numpy())
        assert count_bad(got, ref[k]) == 0, (k, bitdiff(got, ref[k]))


@pytest.mark.parametrize("lim,n,seed", [(_SLOW3, 40_000, 17), (_SLOW3, 2_000, 18), (W.random_limits(8, 5), 40_000, 19),
                                        (W.random_limits(5, 11), 40_000, 20)], ids=lambda x: getattr(x, "name", str(x)))
def test_solve_matches_oracle_on_other_limit_sets(lim, n, seed):
    """limit sets under which the reference gives up on part of the problems (reached = 0: the
    other fields are then unspecified, include/ltp_b200.h), in item mode and as a small batch;
    everything the closed-form kernel reads from the per-joint factors prepared on the host"""
    ltp, ins, sol, ref, _ = _solve_both(lim, n, seed)
    assert np.array_equal(sol.reached.cpu().numpy(), ref["reached"])
    assert np.array_equal(sol.traj_len.cpu().numpy(), ref["traj_len"])
    r = ref["reached"].astype(bool)
    assert r.sum() > n // 20
    assert np.array_equal(sol.slowest.cpu().numpy()[r], ref["slowest"][r])
    for k in ("mod", "opt_case", "ts_case", "final_case", "dir"):
        assert np.array_equal(pm(getattr(sol, k).cpu().numpy())[r], ref[k][r]), k
    for k in ("t_opt", "t_scaled", "v_drive"):
        got = pm(getattr(sol, k).cpu().numpy())
        assert count_bad(got[r], ref[k][r]) == 0, (k, bitdiff(got[r], ref[k][r]))


@pytest.mark.parametrize("n", [40_000, 3_000])  # item mode and the small-batch sequence
@pytest.mark.parametrize("kind", ["tiny", "signed_zero_rest", "on_the_limits"])
def test_states_that_leave_the_division_window_match_oracle(kind, n):
    """The closed-form and attempt-2 kernels test the range of their prepared-reciprocal divisions
    once per stage and repeat a flagged thread out of line (DivDeferred, csrc/ltp_b200.cu). States
    that set the flag in a large share of the threads -- magnitudes down to 1e-320 -- and states
    full of zero numerators of either sign (which must NOT need the second pass to be right),
    all bit-compared with the oracle like any other batch."""
    lim = W.FRANKA7
    rng = np.random.default_rng(20261018)
    qg, q0, v0, a0 = W.random_states(lim, n, 977)
    if kind == "tiny":
        v0 = v0 * 10.0 ** rng.integers(-320, -80, v0.shape)
        a0 = a0 * 10.0 ** rng.integers(-320, -80, a0.shape)
    elif kind == "signed_zero_rest":
        v0 = np.where(rng.random(v0.shape) < 0.5, 0.0, -0.0)
        a0 = np.where(rng.random(a0.shape) < 0.5, 0.0, -0.0)
    else:
        a_max, v_max = np.asarray(lim.a_max), np.asarray(lim.v_max)
        a0 = np.where(rng.random(a0.shape) < 0.5, a_max, -a_max) * (rng.random(a0.shape) < 0.6)
        v0 = np.where(rng.random(v0.shape) < 0.2, v_max * rng.choice([-1.0, 1.0], v0.shape), v0 * 0.2)
    v0, a0 = np.ascontiguousarray(v0), np.ascontiguousarray(a0)
    ltp = _planner(lim)
    ins = [_dev(jm(x)) for x in (qg, q0, v0, a0)]
    sol = ltp.solve(*ins, with_opt=True, with_cases=True)
    torch.cuda.synchronize()
    ref = OraclePort.from_limits(lim).solve(qg, q0, v0, a0, threads=8)
    for k in ("reached", "traj_len"):
        assert np.array_equal(getattr(sol, k).cpu().numpy(), ref[k]), k
    # (a start state that checkInputs rejects -- cc:14-15, a fifth of the states on the limits --
    # has no per-joint results in the reference; the rest is compared field by field)
    ok = ref["reached"].astype(bool)
    assert ok.sum() > 0.5 * n
    assert np.array_equal(sol.slowest.cpu().numpy()[ok], ref["slowest"][ok])
    for k in ("mod", "opt_case", "ts_case", "final_case", "dir"):
        assert np.array_equal(pm(getattr(sol, k).cpu().numpy())[ok], ref[k][ok]), k
    for k in ("t_opt", "t_scaled", "v_drive"):
        got, want = pm(getattr(sol, k).cpu().numpy())[ok], np.asarray(ref[k])[ok]
        assert count_bad(got, want) == 0, (k, bitdiff(got, want))
        # zeros carry the oracle's sign (a zero numerator divided through the reciprocal keeps it)
        z = want == 0.0
        assert np.array_equal(np.signbit(got[z]), np.signbit(want[z])), k


@pytest.mark.parametrize("lim,n,seed", [(W.FRANKA7, 50_000, 301), (W.REF_RANDOM6, 30_000, 302), (W.FRANKA12, 10_000, 303)])
def test_controller_like_states_match_oracle(lim, n, seed):
    """what a replanning controller feeds the planner (workloads.edge_states): joints holding
    position, goals inside the brake-only window, tiny moves, states on a limit, whole arms at
    rest -- solve (all three kernels) and sampled trajectories against the oracle"""
    qg, q0, v0, a0 = W.edge_states(lim, n, seed)
    ltp = _planner(lim)
    ins = [_dev(jm(x)) for x in (qg, q0, v0, a0)]
    sol = ltp.solve(*ins, with_opt=True, with_cases=True)
    torch.cuda.synchronize()
    P = OraclePort.from_limits(lim)
    ref = P.solve(qg, q0, v0, a0, threads=8)
    for k in ("reached", "slowest", "traj_len"):
        assert np.array_equal(getattr(sol, k).cpu().numpy(), ref[k]), k
    for k in ("mod", "opt_case", "ts_case", "final_case", "dir"):
        assert np.array_equal(pm(getattr(sol, k).cpu().numpy()), ref[k]), k
    for k in ("t_opt", "t_scaled", "v_drive"):
        assert count_bad(pm(getattr(sol, k).cpu().numpy()), ref[k]) == 0, k
    m = 64
    sub = [t[:, :m].contiguous() for t in ins]
    sol_m = ltp.solve(*sub)
    traj = ltp.sample(sub[1], sub[2], sub[3], sol_m)
    torch.cuda.synchronize()
    rows = _rows(traj)
    succ = traj.success.cpu().numpy()
    for i in range(m):
        full = P.plan(qg[i], q0[i], v0[i], a0[i])
        assert full["length"] == ref["traj_len"][i] and bool(succ[i]) == full["success"], i
        for k in "qvaj":
            assert count_bad(rows[k][i][:, :full["length"]], full[k]) == 0, (i, k)


def test_solve_matches_reference_build():
    """same, against the reference's own .cc (oracle/_ref), observable fields only"""
    if not Reference.available():
        pytest.skip("oracle/_ref/libltp_ref.so not present")
    lim, n = W.REF_RANDOM6, 40_000
    qg, q0, v0, a0 = W.random_states(lim, n, 99)
    ltp = _planner(lim)
    sol = ltp.solve(*[_dev(jm(x)) for x in (qg, q0, v0, a0)])
    torch.cuda.synchronize()
    ref = Reference.from_limits(lim).solve(qg, q0, v0, a0, threads=8)
    assert np.array_equal(sol.reached.cpu().numpy(), ref["reached"])
    assert np.array_equal(sol.slowest.cpu().numpy(), ref["slowest"])
    assert np.array_equal(pm(sol.mod.cpu().numpy()), ref["mod"])
    assert count_bad(pm(sol.t_scaled.cpu().numpy()), ref["t_scaled"]) == 0
    assert count_bad(pm(sol.v_drive.cpu().numpy()), ref["v_drive"]) == 0


def _rows(traj):
    """-> {field: array [n, dof, samples]} for either layout"""
    out = {}
    for k in "qvaj":
        x = getattr(traj, k).cpu().numpy()
        out[k] = x if traj.layout == "rows" else np.ascontiguousarray(x.transpose(1, 2, 0))
    return out


@pytest.mark.parametrize("layout", ["time_major", "rows"])
@pytest.mark.parametrize("lim,n,seed", [(W.FRANKA7, 96, W.SEEDS[1]), (W.REF_RANDOM6, 300, 11), (W.FRANKA12, 40, 5)])
def test_sampled_trajectories_match_oracle(lim, n, seed, layout):
    ltp, ins, sol, ref, (qg, q0, v0, a0) = _solve_both(lim, n, seed)
    traj = ltp.sample(ins[1], ins[2], ins[3], sol, layout=layout)
    torch.cuda.synchronize()
    P = OraclePort.from_limits(lim)
    rows = _rows(traj)
    succ = traj.success.cpu().numpy()
    tl = sol.traj_len.cpu().numpy()
    bad = 0
    for i in range(n):
        full = P.plan(qg[i], q0[i], v0[i], a0[i])
        assert full["length"] == tl[i]
        assert bool(succ[i]) == full["success"]
        for k in "qvaj":
            bad += count_bad(rows[k][i, :, :tl[i]], full[k])
    assert bad == 0


@pytest.mark.parametrize("layout", ["time_major", "rows"])
def test_fixed_horizon_mode(layout):
    """horizon > traj_len continues with the recurrence's steady state (q_last, 0, 0, 0);
    horizon < traj_len clips; the first min(horizon, traj_len) samples equal the exact mode"""
    lim, n = W.FRANKA7, 64
    ltp, ins, sol, ref, _ = _solve_both(lim, n, 21)
    exact = ltp.sample(ins[1], ins[2], ins[3], sol, layout=layout)
    tl = sol.traj_len.cpu().numpy()
    erows = _rows(exact)
    for H in (int(tl.max()) + 37, int(tl.min()) // 2):
        fixed = ltp.sample(ins[1], ins[2], ins[3], sol, horizon=H, layout=layout)
        torch.cuda.synchronize()
        assert np.array_equal(fixed.success.cpu().numpy(), exact.success.cpu().numpy())
        frows = _rows(fixed)
        for k in "qvaj":
            e, f = erows[k], frows[k]
            for i in range(n):
                m = min(H, tl[i])
                # sample index tl-1+1 may carry the late jerk impulse the exact mode drops
                assert np.array_equal(e[i, :, :m], f[i, :, :m]), (k, i)
                if H > tl[i] + 1:
                    tail = f[i, :, tl[i] + 1:H]
                    if k == "q":
                        assert np.array_equal(tail, np.repeat(f[i, :, tl[i]:tl[i] + 1], tail.shape[1], axis=1))
                    else:
                        assert not tail.any()


def test_per_joint_primitives_on_reference_grids():
    """optSwitchTimes / timeScaling on the exact point sets of the reference's two grid
    tests (tests/src/long_term_planner_tests.cc:264-407)"""
    lim = W.REF_GRID
    ltp = _planner(lim)
    P = OraclePort.from_limits(lim)
    for ts in (False, True):
        qg, v0, a0 = W.reference_grid_points(ts)
        q0 = np.full_like(qg, 0.5)
        vd = np.full_like(qg, 1.0)
        d = [_dev(x[None, :]) for x in (qg, q0, v0, a0, vd)]
        got = ltp.optSwitchTimesBatch(*d)
        torch.cuda.synchronize()
        ref = P.opt_switch_times(qg, q0, v0, a0, vd)
        assert np.array_equal(got["ok"].cpu().numpy()[0], ref["ok"])
        assert np.array_equal(got["case"].cpu().numpy()[0], ref["case"])
        assert np.array_equal(got["mod"].cpu().numpy()[0], ref["mod"])
        assert np.array_equal(got["dir"].cpu().numpy()[0], ref["dir"])
        assert count_bad(pm(got["t"].cpu().numpy())[:, 0, :], ref["t"]) == 0
        if ts:
            for inc in (0.05, 0.1, 0.2, 0.5, 1.0, 2.0):
                tr = ref["t"][:, 6] + inc
                g2 = ltp.timeScalingBatch(d[0], d[1], d[2], d[3], _dev(ref["dir"][None, :]), _dev(tr[None, :]))
                torch.cuda.synchronize()
                r2 = P.time_scaling(qg, q0, v0, a0, ref["dir"], tr, threads=8)
                for k in ("ok", "mod", "ts_case", "final_case"):
                    assert np.array_equal(g2[k].cpu().numpy()[0], r2[k]), (inc, k)
                assert count_bad(pm(g2["t"].cpu().numpy())[:, 0, :], r2["t"]) == 0
                assert count_bad(g2["v_drive"].cpu().numpy()[0], r2["v_drive"]) == 0


def test_opt_braking_batch():
    lim = W.REF_RANDOM6
    rng = np.random.default_rng(5)
    n = 5000
    v0 = rng.uniform(-1, 1, (n, 6))
    a0 = rng.uniform(-2, 2, (n, 6))
    ltp = _planner(lim)
    got = ltp.optBrakingBatch(_dev(jm(v0)), _dev(jm(a0)))
    torch.cuda.synchronize()
    P = OraclePort.from_limits(lim)
    joint = np.tile(np.arange(6, dtype=np.int32), n)
    ref = P.opt_braking(v0.ravel(), a0.ravel(), joint)
    assert np.array_equal(pm(got["dir"].cpu().numpy()).ravel(), ref["dir"])
    assert count_bad(pm(got["q"].cpu().numpy()).ravel(), ref["q"]) == 0
    assert count_bad(pm(got["t_rel"].cpu().numpy()).reshape(-1, 3), ref["t_rel"]) == 0


def test_single_plan_api_matches_oracle():
    """LongTermPlanner.planTrajectory / protected methods, one problem at a time"""
    from longtermplanner_b200 import Trajectory
    lim = W.FRANKA7
    ltp = _planner(lim)
    P = OraclePort.from_limits(lim)
    qg, q0, v0, a0 = W.random_states(lim, 5, 1234)
    for i in range(5):
        tr = Trajectory()
        ok = ltp.planTrajectory(qg[i], q0[i], v0[i], a0[i], tr)
        full = P.plan(qg[i], q0[i], v0[i], a0[i])
        assert ok == full["success"] and tr.length == full["length"] and tr.dof == 7
        for k in "qvaj":
            assert count_bad(np.asarray(getattr(tr, k)), full[k]) == 0
    # out-of-limits start state: early false, trajectory untouched (cc:14-15)
    tr = Trajectory()
    bad_q0 = q0[0].copy()
    bad_q0[0] = 10.0
    assert ltp.planTrajectory(qg[0], bad_q0, v0[0], a0[0], tr) is False and tr.length == 0
    assert ltp.checkInputs(bad_q0, v0[0], a0[0]) is False


def test_host_entry_point_equals_device_entry_point():
    lim, n = W.FRANKA7, 5000
    qg, q0, v0, a0 = W.random_states(lim, n, 77)
    ltp = _planner(lim)
    host = ltp.solve_host(*[jm(x) for x in (qg, q0, v0, a0)], with_opt=True, with_cases=True)
    sol = ltp.solve(*[_dev(jm(x)) for x in (qg, q0, v0, a0)], with_opt=True, with_cases=True)
    torch.cuda.synchronize()
    for k in ("t_scaled", "dir", "v_drive", "mod", "slowest", "traj_len", "reached", "t_opt", "opt_case",
              "ts_case", "final_case"):
        a, b = host[k], getattr(sol, k).cpu().numpy()
        assert np.array_equal(a, b, equal_nan=True) if a.dtype.kind == "f" else np.array_equal(a, b), k


@pytest.mark.parametrize("lim,n", [(W.FRANKA7, 3 * 65536 + 777), (W.REF_RANDOM6, 65536 + 1), (W.FRANKA12, 65536)])
def test_host_entry_point_chunk_pipeline(lim, n):
    """ltp_solve_host above one chunk (2^16 problems) runs a two-slot copy/solve/copy pipeline
    with 2-D copies into the joint-major host arrays; ragged last chunk; pinned and pageable"""
    qg, q0, v0, a0 = W.random_states(lim, n, 78)
    ltp = _planner(lim)
    ins = [jm(x) for x in (qg, q0, v0, a0)]
    host = ltp.solve_host(*ins, with_cases=True)  # pageable numpy buffers
    pinned = [torch.from_numpy(x).pin_memory().numpy() for x in ins]
    host2 = ltp.solve_host(*pinned)
    sol = ltp.solve(*[_dev(x) for x in ins], with_cases=True)
    torch.cuda.synchronize()
    for k in ("t_scaled", "dir", "v_drive", "mod", "slowest", "traj_len", "reached", "opt_case", "ts_case",
              "final_case"):
        a, b = host[k], getattr(sol, k).cpu().numpy()
        assert np.array_equal(a, b, equal_nan=(a.dtype.kind == "f")), k
        if host2.get(k) is not None:
            assert np.array_equal(host2[k], b, equal_nan=(a.dtype.kind == "f")), k


def test_layouts_agree_and_limit_violations_are_flagged():
    """both layouts hold the same samples; a plan that ends outside [q_min, q_max] reports
    success = 0 with the trajectory still written (reference cc:59-61)"""
    lim = W.FRANKA7
    ltp = _planner(lim)
    qg, q0, v0, a0 = W.random_states(lim, 257, 41)
    qg[5, 3] = lim.q_max[3] + 0.5    # goal outside the joint range: reached, but final check fails
    qg[200, 0] = lim.q_min[0] - 0.5
    ins = [_dev(jm(x)) for x in (qg, q0, v0, a0)]
    sol = ltp.solve(*ins)
    a = ltp.sample(ins[1], ins[2], ins[3], sol, layout="time_major")
    b = ltp.sample(ins[1], ins[2], ins[3], sol, layout="rows")
    torch.cuda.synchronize()
    sa, sb = a.success.cpu().numpy(), b.success.cpu().numpy()
    assert np.array_equal(sa, sb) and sa[5] == 0 and sa[200] == 0 and sa.sum() == 255
    P = OraclePort.from_limits(lim)
    assert P.plan(qg[5], q0[5], v0[5], a0[5])["success"] is False
    tl = sol.traj_len.cpu().numpy()
    ra, rb = _rows(a), _rows(b)
    for k in "qvaj":
        for i in (0, 5, 100, 200, 256):
            assert np.array_equal(ra[k][i, :, :tl[i]], rb[k][i, :, :tl[i]])


def test_rejected_inputs_and_empty_batch():
    lim = W.FRANKA7
    ltp = _planner(lim)
    qg, q0, v0, a0 = W.random_states(lim, 64, 5)
    q0[3, 2] = 99.0      # outside [q_min, q_max]
    v0[7, 0] = 1e3       # outside v_max
    sol = ltp.solve(*[_dev(jm(x)) for x in (qg, q0, v0, a0)])
    torch.cuda.synchronize()
    reached = sol.reached.cpu().numpy()
    ref = OraclePort.from_limits(lim).solve(qg, q0, v0, a0)
    assert np.array_equal(reached, ref["reached"]) and reached[3] == 0 and reached[7] == 0
    assert sol.traj_len.cpu().numpy()[3] == 0
    empty = torch.empty(7, 0, dtype=torch.float64, device="cuda")
    s0 = ltp.solve(empty, empty, empty, empty)
    assert s0.n == 0


@pytest.mark.parametrize("lim,n,seed", [(W.FRANKA7, 50_000, 31), (W.REF_RANDOM6, 50_000, 32), (W.REF_GRID, 100_000, 33)])
def test_two_kernel_solve_equals_generic_kernel(lim, n, seed):
    """LTP_SOLVE_AUTO (closed-form kernel + work list drained by the generic kernel) must be
    bit-identical to running every problem through the generic kernel"""
    qg, q0, v0, a0 = W.random_states(lim, n, seed)
    ins = [_dev(jm(x)) for x in (qg, q0, v0, a0)]
    ltp = _planner(lim)
    auto = ltp.solve(*ins, with_opt=True, with_cases=True)
    ltp.setSolveMode(True)
    gen = ltp.solve(*ins, with_opt=True, with_cases=True)
    torch.cuda.synchronize()
    for k in ("t_scaled", "dir", "v_drive", "mod", "slowest", "traj_len", "reached", "t_opt", "opt_case",
              "ts_case", "final_case"):
        a, b = getattr(auto, k), getattr(gen, k)
        if a.dtype == torch.float64:
            assert torch.equal(a.view(torch.int64), b.view(torch.int64)) or torch.equal(a.nan_to_num(), b.nan_to_num()), k
        else:
            assert torch.equal(a, b), k
    # and the same call twice gives the same answer (work list is reset per call)
    again = ltp.solve(*ins)
    ltp.setSolveMode(False)
    auto2 = ltp.solve(*ins)
    torch.cuda.synchronize()
    assert torch.equal(again.traj_len, auto2.traj_len) and torch.equal(auto.traj_len, auto2.traj_len)


@pytest.mark.parametrize("lim,seed", [(W.FRANKA7, 31), (W.REF_RANDOM6, 32), (W.FRANKA12, 33)])
def test_small_batch_host_path_equals_batch_kernels(lim, seed):
    """ltp_plan_host with a handful of problems (the drop-in planTrajectory is n = 1) takes the
    latency path: mapped host staging, every-branch solve kernel, piece-wise row sampler. Its
    rows must be the batch kernels' rows bit for bit -- exact length, fixed horizon longer and
    shorter than the plan, capacity not a multiple of four, the too-small-capacity answer."""
    import ctypes as C
    from longtermplanner_b200 import _capi as capi
    n_all = 48
    qg, q0, v0, a0 = W.random_states(lim, n_all, seed)
    ltp = _planner(lim)
    ins = [_dev(jm(x)) for x in (qg, q0, v0, a0)]
    sol = ltp.solve(*ins)
    tl = sol.traj_len.cpu().numpy()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731

    def call(sel, horizon, cap):
        n = len(sel)
        rows = [np.full((n, lim.dof, cap), np.nan) for _ in range(4)]
        ln, ok, needed = np.zeros(n, np.int32), np.zeros(n, np.uint8), capi.i64(0)
        hin = [np.ascontiguousarray(jm(x[sel])) for x in (qg, q0, v0, a0)]
        rc = capi.plan_host(ltp._h, n, *[vp(x) for x in hin], horizon, cap, *[vp(r) for r in rows], vp(ln), vp(ok),
                            C.byref(needed))
        return rc, rows, ln, ok, int(needed.value)

    for horizon in (0, int(tl.max()) + 13, max(int(tl.min()) // 2, 4)):
        want = ltp.sample(ins[1], ins[2], ins[3], sol, horizon=horizon, layout="rows")
        torch.cuda.synchronize()
        wrows = [getattr(want, k).cpu().numpy() for k in "qvaj"]
        wok = want.success.cpu().numpy()
        for sel in ([0], [7], [1, 2, 3], list(range(10, 15)), list(range(16, 48))):
            sel = np.asarray(sel)
            cap = (horizon if horizon else int(tl[sel].max())) + 3  # not a multiple of four in general
            if 4 * len(sel) * lim.dof * (cap + 3) * 8 > (4 << 20):
                continue  # beyond the staging block: the general path, covered elsewhere
            rc, rows, ln, ok, needed = call(sel, horizon, cap)
            assert rc == 0
            assert np.array_equal(ln, tl[sel]) and np.array_equal(ok, wok[sel])
            for f in range(4):
                for i, pidx in enumerate(sel):
                    m = horizon if horizon else tl[pidx]
                    assert np.array_equal(rows[f][i, :, :m], wrows[f][pidx, :, :m]), (horizon, f, pidx)
    # capacity too small: the needed capacity comes back, nothing is reported as success
    rc, rows, ln, ok, needed = call(np.asarray([5]), 0, 8)
    assert rc == capi.LTP_ERR_CAPACITY and needed == tl[5] and not ok.any()


def test_single_item_calls_through_mapped_staging():
    """optBraking / optSwitchTimes / timeScaling / getTrajectory one item at a time: inputs
    and results travel through pinned, device-mapped host memory (no copies); same answers as
    the batched primitives and the oracle."""
    lim = W.REF_RANDOM6
    ltp = _planner(lim)
    P = OraclePort.from_limits(lim)
    qg, q0, v0, a0 = W.random_states(lim, 40, 99)
    for i in range(40):
        jt = i % lim.dof
        ok, qs, trel, d = ltp.optBraking(jt, v0[i, jt], a0[i, jt])
        rb = P.opt_braking(np.array([v0[i, jt]]), np.array([a0[i, jt]]), joint=np.array([jt]))
        assert count_bad([qs, *trel], [rb["q"][0], *rb["t_rel"][0]]) == 0 and d == rb["dir"][0]
        full = P.plan(qg[i], q0[i], v0[i], a0[i])
        from longtermplanner_b200 import Trajectory
        tr = Trajectory()
        assert ltp.planTrajectory(qg[i], q0[i], v0[i], a0[i], tr) == full["success"]
        assert tr.length == full["length"]
        for k in "qvaj":
            assert count_bad(np.asarray(getattr(tr, k)), full[k]) == 0


def _random_limits(dof, seed):
    return W.random_limits(dof, seed)


@pytest.mark.parametrize("dof", [1, 2, 3, 5, 8, 9, 16, 17, 31, 32])
def test_every_joint_count_bucket(dof):
    """The kernels are instantiated per CTA-size bucket (1, 2, 4, 8, 16, 32 warps; 6, 7, 12 exact)
    and the row sampler packs floor(32 / dof) problems per warp: every bucket, with random limit
    sets (mixed ratios a_max/j_max, three sample times), against the oracle -- solve in both
    modes, both sampler layouts, the small-batch host path."""
    import ctypes as C
    from longtermplanner_b200 import _capi as capi
    lim = _random_limits(dof, 1000 + dof)
    n = 700 if dof <= 9 else 200
    ltp, ins, sol, ref, (qg, q0, v0, a0) = _solve_both(lim, n, 4000 + dof)
    assert np.array_equal(sol.reached.cpu().numpy(), ref["reached"])
    assert np.array_equal(sol.traj_len.cpu().numpy(), ref["traj_len"])
    # a problem the reference gives up on (start state rejected, a joint without a solution) has
    # reached = 0 and traj_len = 0; its other fields are unspecified (include/ltp_b200.h)
    r = ref["reached"].astype(bool)
    assert r.sum() > n // 2
    assert np.array_equal(sol.slowest.cpu().numpy()[r], ref["slowest"][r])
    assert np.array_equal(pm(sol.final_case.cpu().numpy())[r], ref["final_case"][r])
    assert np.array_equal(pm(sol.ts_case.cpu().numpy())[r], ref["ts_case"][r])
    assert np.array_equal(pm(sol.mod.cpu().numpy())[r], ref["mod"][r])
    assert np.array_equal(pm(sol.dir.cpu().numpy())[r], ref["dir"][r])
    assert count_bad(pm(sol.t_scaled.cpu().numpy())[r], ref["t_scaled"][r]) == 0
    assert count_bad(pm(sol.v_drive.cpu().numpy())[r], ref["v_drive"][r]) == 0
    ltp.setSolveMode(True)
    gen = ltp.solve(*ins, with_opt=True, with_cases=True)
    ltp.setSolveMode(False)
    for k in ("t_scaled", "v_drive", "dir", "mod", "traj_len", "reached", "slowest", "final_case", "ts_case"):
        assert bitdiff(getattr(gen, k).cpu().numpy(), getattr(sol, k).cpu().numpy()) == 0, k
    P = OraclePort.from_limits(lim)
    tl = sol.traj_len.cpu().numpy()
    rows = {}
    for layout in ("time_major", "rows"):
        traj = ltp.sample(ins[1], ins[2], ins[3], sol, layout=layout)
        torch.cuda.synchronize()
        rows[layout] = _rows(traj)
        succ = traj.success.cpu().numpy()
        for i in range(0, n, max(n // 12, 1)):
            full = P.plan(qg[i], q0[i], v0[i], a0[i])
            assert bool(succ[i]) == full["success"]
            if not r[i]:  # early false of the reference: no trajectory at all
                assert full["length"] <= 0 and tl[i] == 0 and not succ[i]
                continue
            assert full["length"] == tl[i]
            for k in "qvaj":
                assert count_bad(rows[layout][k][i, :, :tl[i]], full[k]) == 0, (layout, k, i)
    for k in "qvaj":
        for i in range(n):
            assert np.array_equal(rows["rows"][k][i, :, :tl[i]], rows["time_major"][k][i, :, :tl[i]])
    # small-batch host path: three problems at once
    sel = np.array([0, n // 2, n - 1])
    cap = int(tl[sel].max()) + 1
    if 4 * 3 * dof * (cap + 3) * 8 <= (4 << 20):
        out = [np.full((3, dof, cap), np.nan) for _ in range(4)]
        ln, ok, needed = np.zeros(3, np.int32), np.zeros(3, np.uint8), capi.i64(0)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        hin = [np.ascontiguousarray(jm(x[sel])) for x in (qg, q0, v0, a0)]
        assert capi.plan_host(ltp._h, 3, *[vp(x) for x in hin], 0, cap, *[vp(r) for r in out], vp(ln), vp(ok),
                              C.byref(needed)) == 0
        assert np.array_equal(ln, tl[sel])
        for f, k in enumerate("qvaj"):
            for i, pidx in enumerate(sel):
                assert np.array_equal(out[f][i, :, :tl[pidx]], rows["rows"][k][pidx, :, :tl[pidx]]), (k, pidx)


@pytest.mark.parametrize("lim,n", [(W.FRANKA7, 1111), (W.FRANKA12, 640), (W.REF_RANDOM6, 500)])
def test_sorted_slot_sampling_equals_problem_order_sampling(lim, n):
    """ltp_sample_batch_sorted: slot k of the time-major tensors holds problem order[k] (longest
    trajectory first); gathered back, the samples are those of ltp_sample_batch bit for bit and
    the success flags (indexed by problem in both) are equal"""
    ltp, ins, sol, ref, _ = _solve_both(lim, n, 515)
    plain = ltp.sample(ins[1], ins[2], ins[3], sol)
    srt = ltp.sample(ins[1], ins[2], ins[3], sol, sorted_slots=True)
    torch.cuda.synchronize()
    order = srt.order.long()
    assert torch.equal(torch.sort(order).values, torch.arange(n, device="cuda"))
    tl = (sol.traj_len.long() * sol.reached.long())
    lens = tl[order]
    assert bool((lens[:-1] + 8 > lens[1:]).all())
    assert torch.equal(plain.success, srt.success)
    tlc = tl.cpu().numpy()
    oc = order.cpu().numpy()
    for k in "qvaj":
        a, b = getattr(plain, k).cpu().numpy(), getattr(srt, k).cpu().numpy()
        for slot in range(0, n, 7):
            p = oc[slot]
            assert np.array_equal(a[:tlc[p], p, :], b[:tlc[p], slot, :]), (k, slot, p)


@pytest.mark.parametrize("lim,n", [(W.FRANKA7, 1024), (W.FRANKA12, 257), (W.random_limits(1, 71), 3000),
                                   (W.random_limits(3, 72), 700), (W.random_limits(17, 73), 130)])
@pytest.mark.parametrize("mode", ["exact", "fixed_odd", "clipped"])
def test_rows_layout_equals_time_major_on_batches(lim, n, mode):
    """batches in the rows layout: same samples, bit for bit, as the time-major
    kernel and the oracle -- exact lengths (odd and even), an odd fixed horizon, rows clipped by a
    capacity below their length; problems that were not planned hold their start position in
    fixed-horizon mode and are left alone in exact-length mode"""
    assert n * lim.dof >= 1024
    qg, q0, v0, a0 = W.random_states(lim, n, 77)
    q0[3, 0] = lim.q_max[0] + 1.0     # rejected by checkInputs: reached = 0
    qg[5, 0] = lim.q_max[0] + 0.5     # reached, but ends outside the joint range: success = 0
    ltp = _planner(lim)
    ins = [_dev(jm(x)) for x in (qg, q0, v0, a0)]
    sol = ltp.solve(*ins)
    tl = sol.traj_len.cpu().numpy()
    longest = int(tl.max())
    if mode == "exact":
        kw, cap = dict(horizon=0), longest
    elif mode == "fixed_odd":
        kw, cap = dict(horizon=longest // 2 * 2 + 1), longest // 2 * 2 + 1
    else:
        kw, cap = dict(horizon=0), max(int(np.median(tl[tl > 0])) // 4 * 4, 4)
    out_r = ltp.alloc_trajectories(n, cap, "rows")
    out_t = ltp.alloc_trajectories(n, cap, "time_major")
    for o in (out_r, out_t):
        for k in "qvaj":
            getattr(o, k).fill_(-777.0)
    r = ltp.sample(ins[1], ins[2], ins[3], sol, out=out_r, **kw)
    t = ltp.sample(ins[1], ins[2], ins[3], sol, out=out_t, **kw)
    torch.cuda.synchronize()
    assert np.array_equal(r.success.cpu().numpy(), t.success.cpu().numpy())
    assert r.success[3].item() == 0 and r.success[5].item() == 0 and sol.reached[3].item() == 0
    rr, tt = _rows(r), _rows(t)
    H = kw["horizon"]
    for k in "qvaj":
        for i in range(n):
            m = H if H > 0 else min(int(tl[i]), cap)
            assert np.array_equal(rr[k][i, :, :m], tt[k][i, :, :m]), (k, i)
            assert np.all(rr[k][i, :, m:cap] == -777.0), (k, i)   # nothing written behind the row
    if H > 0:   # the unplanned problem holds (q_0, 0, 0, 0)
        assert np.array_equal(rr["q"][3, :, :H], np.broadcast_to(q0[3][:, None], (lim.dof, H)))
        assert np.all(rr["v"][3, :, :H] == 0) and np.all(rr["a"][3, :, :H] == 0) and np.all(rr["j"][3, :, :H] == 0)
    else:
        assert np.all(rr["q"][3] == -777.0)
    # against the oracle's own sampler on a few problems
    P = OraclePort.from_limits(lim)
    for i in (0, 1, 5, n - 1):
        full = P.plan(qg[i], q0[i], v0[i], a0[i], stride=max(longest, 8))
        m = min(full["length"], cap) if H == 0 else min(full["length"], H)
        for k in "qvaj":
            assert count_bad(rr[k][i, :, :m], full[k][:, :m]) == 0, (k, i)


def test_every_root_finder_candidate_is_accepted_somewhere_and_matches_the_oracle():
    """The reference's own limits never let the sextic candidate (cc:606-629) win the search; a
    one-joint random limit set does (found by scanning limit sets with the oracle): candidates
    3..8 -- quartic, quartic, quintic, quartic, quartic, sextic -- are all the ACCEPTED attempt of
    some joint here, so smallest_root<4|5|6> is checked on the accept path as well."""
    lim = W.random_limits(1, 1001)
    n = 400_000
    qg, q0, v0, a0 = (x[:, 0].copy() for x in W.random_states(lim, n, 78))
    ltp = _planner(lim)
    P = OraclePort.from_limits(lim)
    vd = np.full(n, lim.v_max[0])
    o = P.opt_switch_times(qg, q0, v0, a0, vd, threads=8)
    d = [_dev(x[None, :]) for x in (qg, q0, v0, a0)]
    seen = np.zeros(10, np.int64)
    for inc in (0.02, 0.05, 0.2):
        tr = o["t"][:, 6] + inc
        got = ltp.timeScalingBatch(d[0], d[1], d[2], d[3], _dev(o["dir"][None, :]), _dev(tr[None, :]))
        torch.cuda.synchronize()
        ref = P.time_scaling(qg, q0, v0, a0, o["dir"], tr, threads=8)
        for k in ("ok", "mod", "ts_case", "final_case"):
            assert np.array_equal(got[k].cpu().numpy()[0], ref[k]), (inc, k)
        assert count_bad(pm(got["t"].cpu().numpy())[:, 0, :], ref["t"]) == 0
        assert count_bad(got["v_drive"].cpu().numpy()[0], ref["v_drive"]) == 0
        seen += np.bincount(ref["ts_case"], minlength=10)[:10]
    assert all(seen[k] > 0 for k in range(1, 9)), seen
    assert seen[8] >= 50 and seen[5] >= 50, seen
