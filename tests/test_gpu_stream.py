"""GPU tests of the streaming driver (BASELINE.json configs[4] shape: more trajectories than fit
in memory, recycled ring) and of the receding-horizon step (configs[2]: replanning every 10 ms
from the state the previous plan has reached)."""
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


def _view_as_traj(view):
    from longtermplanner_b200.planner import BatchTrajectories
    return BatchTrajectories("time_major", view["horizon"], view["capacity"], view["q"], view["v"], view["a"],
                             view["j"], view["success"], view["traj_len"])


@pytest.mark.parametrize("lim,n,chunk", [(W.FRANKA12, 5000, 1536), (W.FRANKA7, 3000, 3000), (W.FRANKA7, 1000, 4096)])
def test_streamed_run_equals_one_shot(lim, n, chunk):
    """per-row checksums, lengths and flags gathered chunk by chunk from the ring equal those
    of one planTrajectories call over the whole batch, bit for bit"""
    from longtermplanner_b200 import devtools
    ltp = _planner(lim)
    ins = devtools.random_states_device(lim, n, W.SEEDS[5], start=77)
    sol, traj = ltp.planTrajectories(*ins)
    want = devtools.row_stats(traj, sol.traj_len)
    got = torch.zeros_like(want)
    succ = torch.zeros(n, dtype=torch.uint8, device="cuda")
    tl = torch.zeros(n, dtype=torch.int32, device="cuda")
    seen = []

    def consumer(view, stream):
        a, c = view["first"], view["count"]
        seen.append((a, c))
        got[a:a + c] = devtools.row_stats(_view_as_traj(view), view["traj_len"])
        succ[a:a + c] = view["success"]
        tl[a:a + c] = view["traj_len"]

    stats = ltp.planStream(*ins, chunk=chunk, capacity=4096, consumer=consumer)
    torch.cuda.synchronize()
    assert seen == [(a, min(chunk, n - a)) for a in range(0, n, chunk)]
    assert torch.equal(got.view(torch.int64), want.view(torch.int64))
    assert torch.equal(succ, traj.success) and torch.equal(tl, sol.traj_len)
    assert stats["problems"] == n and stats["chunks"] == len(seen)
    assert stats["reached"] == int(sol.reached.sum()) and stats["success"] == int(traj.success.sum())
    assert stats["samples"] == int(sol.traj_len.long().sum()) * lim.dof and stats["bytes"] == stats["samples"] * 32
    assert stats["clipped"] == 0 and stats["max_traj_len"] == int(sol.traj_len.max())


@pytest.mark.parametrize("lim,n,chunk", [(W.FRANKA12, 5000, 1536), (W.FRANKA7, 4097, 2048), (W.REF_RANDOM6, 1500, 700)])
def test_streamed_run_sorted_slots(lim, n, chunk):
    """sorted-slot mode: every chunk's problems ordered by trajectory length on the device, slot k
    of the chunk's trajectories holds problem order[k]. Same samples bit for bit once the slots
    are mapped back; order is a permutation, longest first; flags and totals unchanged."""
    from longtermplanner_b200 import devtools
    ltp = _planner(lim)
    ins = devtools.random_states_device(lim, n, W.SEEDS[5], start=123)
    sol, traj = ltp.planTrajectories(*ins)
    want = devtools.row_stats(traj, sol.traj_len)
    got = torch.zeros_like(want)
    succ = torch.zeros(n, dtype=torch.uint8, device="cuda")
    orders = []

    def consumer(view, stream):
        a, c = view["first"], view["count"]
        order = view["order"].long()
        orders.append(order.clone())
        # row statistics of slot k are those of problem order[k]; the slot's own length applies
        slot_len = view["traj_len"][order].contiguous()
        got[a + order] = devtools.row_stats(_view_as_traj(view), slot_len)
        succ[a:a + c] = view["success"]

    stats = ltp.planStream(*ins, chunk=chunk, capacity=4096, consumer=consumer, sorted_slots=True)
    torch.cuda.synchronize()
    assert torch.equal(got.view(torch.int64), want.view(torch.int64))
    assert torch.equal(succ, traj.success)
    tl = sol.traj_len.long() * sol.reached.long()
    for k, order in enumerate(orders):
        c = order.numel()
        assert torch.equal(torch.sort(order).values, torch.arange(c, device="cuda"))
        lens = tl[k * chunk + order]
        assert bool((lens[:-1] + 8 > lens[1:]).all())  # longest first, buckets of 8 samples at this capacity
    assert stats["samples"] == int(sol.traj_len.long().sum()) * lim.dof and stats["success"] == int(traj.success.sum())
    # the switch is per call: the default run afterwards is in problem order again
    seen = []
    ltp.planStream(*ins, chunk=chunk, capacity=4096, consumer=lambda v, s: seen.append(v["order"]))
    assert all(o is None for o in seen)


def test_streamed_run_fixed_horizon_and_clipping():
    from longtermplanner_b200 import devtools
    lim, n, H = W.FRANKA7, 2500, 600
    ltp = _planner(lim)
    ins = devtools.random_states_device(lim, n, 99)
    sol = ltp.solve(*ins)
    stats = ltp.planStream(*ins, chunk=1024, horizon=H, capacity=H)
    tl = sol.traj_len.cpu().numpy()
    assert stats["samples"] == n * lim.dof * H
    assert stats["clipped"] == int((tl > H).sum()) > 0
    # exact-length mode with a capacity below the longest trajectory
    cap = int(np.median(tl))
    stats = ltp.planStream(*ins, chunk=1024, horizon=0, capacity=cap)
    assert stats["clipped"] == int((tl > cap).sum())
    assert stats["samples"] == int(np.minimum(tl, cap).sum()) * lim.dof


def test_streamed_run_without_consumer_and_empty_batch():
    from longtermplanner_b200 import devtools
    lim = W.FRANKA7
    ltp = _planner(lim)
    ins = devtools.random_states_device(lim, 300, 5)
    stats = ltp.planStream(*ins, chunk=128)
    assert stats["problems"] == 300 and stats["chunks"] == 3
    empty = [t[:, :0].contiguous() for t in ins]
    assert ltp.planStream(*empty, chunk=128)["problems"] == 0


def test_consumer_exception_propagates():
    from longtermplanner_b200 import devtools
    lim = W.FRANKA7
    ltp = _planner(lim)
    ins = devtools.random_states_device(lim, 300, 5)

    def consumer(view, stream):
        raise KeyError("boom")

    with pytest.raises(KeyError):
        ltp.planStream(*ins, chunk=128, consumer=consumer)
    torch.cuda.synchronize()
    assert ltp.planStream(*ins, chunk=128)["problems"] == 300  # the planner is still usable


def test_advance_takes_the_sample_at_the_tick():
    lim, n, H, tick = W.FRANKA7, 777, 64, 9
    ltp = _planner(lim)
    qg, q0, v0, a0 = W.random_states(lim, n, 17)
    ins = [torch.from_numpy(jm(x)).cuda() for x in (qg, q0, v0, a0)]
    sol = ltp.solve(*ins)
    traj = ltp.sample(ins[1], ins[2], ins[3], sol, horizon=H)
    nq, nv, na = (t.clone() for t in ins[1:])
    ltp.advance(traj, tick, nq, nv, na, clamp=False)
    torch.cuda.synchronize()
    assert torch.equal(nq, traj.q[tick].T.contiguous()) and torch.equal(nv, traj.v[tick].T.contiguous())
    assert torch.equal(na, traj.a[tick].T.contiguous())
    # exact-length trajectories: a tick past the end gives the final state
    exact = ltp.sample(ins[1], ins[2], ins[3], sol)
    big = exact.stride - 1
    ltp.advance(exact, big, nq, nv, na, clamp=False)
    tl = sol.traj_len.long() - 1
    idx = torch.minimum(tl, torch.tensor(big, device="cuda"))
    want = exact.q[idx, torch.arange(n, device="cuda"), :].T.contiguous()
    assert torch.equal(nq, want)
    assert bool((nv == 0).all()) and bool((na == 0).all())


def test_receding_horizon_loop_matches_oracle_and_never_rejects_its_own_state():
    """configs[2] in small: replan every 10 samples towards a fresh goal from the state the
    previous plan reached; every replan is accepted (clamped Euler drift), and the solve from
    that state agrees with the oracle on the same numbers"""
    lim, n, H, tick = W.FRANKA7, 256, 200, 9
    ltp = _planner(lim)
    P = OraclePort.from_limits(lim)
    qg, q0, v0, a0 = W.random_states(lim, n, 23)
    state = [torch.from_numpy(jm(x)).cuda() for x in (q0, v0, a0)]
    for step in range(12):
        goal = torch.from_numpy(jm(W.random_states(lim, n, 100 + step)[0])).cuda()
        sol = ltp.solve(goal, *state, with_cases=True)
        traj = ltp.sample(*state, sol, horizon=H)
        torch.cuda.synchronize()
        assert bool(sol.reached.all()), step
        if step % 4 == 0:
            ref = P.solve(goal.cpu().numpy().T.copy(), *[s.cpu().numpy().T.copy() for s in state], threads=4)
            assert np.array_equal(sol.traj_len.cpu().numpy(), ref["traj_len"])
            assert np.array_equal(sol.final_case.cpu().numpy().T, ref["final_case"])
            assert count_bad(sol.t_scaled.cpu().numpy().transpose(2, 1, 0), ref["t_scaled"]) == 0
        ltp.advance(traj, tick, *state)
    # states stay physically plausible
    q_min, q_max, v_max, a_max, _ = (torch.from_numpy(x).cuda()[:, None] for x in lim.arrays())
    assert bool((state[0] >= q_min).all() and (state[0] <= q_max).all())
    assert bool((state[1].abs() <= v_max).all() and (state[2].abs() <= a_max).all())


def test_replanning_step_can_be_captured_in_a_cuda_graph():
    lim, n, H, tick = W.FRANKA7, 512, 128, 9
    ltp = _planner(lim)
    qg, q0, v0, a0 = W.random_states(lim, n, 31)
    goal = torch.from_numpy(jm(qg)).cuda()

    def fresh():
        return [torch.from_numpy(jm(x)).cuda() for x in (q0, v0, a0)]

    def run(state, sol, traj, steps):
        for _ in range(steps):
            ltp.solve(goal, *state, out=sol)
            ltp.sample(*state, sol, horizon=H, out=traj)
            ltp.advance(traj, tick, *state)

    eager_state = fresh()
    sol_e, traj_e = ltp.alloc_solution(n), ltp.alloc_trajectories(n, H)
    run(eager_state, sol_e, traj_e, 3)
    torch.cuda.synchronize()

    state = fresh()
    sol, traj = ltp.alloc_solution(n), ltp.alloc_trajectories(n, H)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        run(state, sol, traj, 1)  # warm-up on the capture stream: allocates the work list
        for a, b in zip(state, fresh()):
            a.copy_(b)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            run(state, sol, traj, 1)
        for a, b in zip(state, fresh()):  # capture does not execute
            a.copy_(b)
        for _ in range(3):
            g.replay()
    torch.cuda.synchronize()
    for a, b in zip(state, eager_state):
        assert torch.equal(a, b)
    assert torch.equal(traj.q, traj_e.q) and torch.equal(sol.traj_len, sol_e.traj_len)


def test_refined_single_joint_grid_matches_oracle_at_every_point():
    """configs[3] at reduced size (48^3 points): the sweep tool's three passes run and EVERY point
    is compared with the oracle (case bytes, accepted attempt, flags exact; times and cruise
    speeds within tolerance), so the case histograms are the oracle's. Which cases a grid
    populates depends on its resolution; the committed 256^3 run (profiles/) populates all
    eight phase patterns and attempts 1..7."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "grid_sweep.py"), "48", "8", "1"],
                         capture_output=True, text=True, check=True).stdout
    r = json.loads(out)
    par = r["parity_vs_cpu_oracle"]
    assert par["points"] == 48 ** 3
    assert all(v == 0 for v in par["exact_field_mismatches"].values()), par
    assert all(v == 0 for v in par["value_mismatches"].values()), par
    h = r["case_histogram_pass1_plus_nested"]
    assert sum(1 for k in range(1, 9) if h.get(str(k), 0) > 0) >= 5, h
    acc = r["pass3_goal_accuracy"]["results"]
    assert acc["optimal"]["max_abs_goal_error"] < 0.02  # the reference's own bar, tests.cc:318
    assert all(acc[k]["max_abs_goal_error"] < 0.02 for k in acc)


class _DlpackOnly:
    """a producer that is NOT a torch tensor: only the DLPack protocol (what CuPy / JAX arrays offer)"""

    def __init__(self, t):
        self._t = t

    def __dlpack__(self, stream=None):
        return self._t.__dlpack__(stream=stream) if stream is not None else self._t.__dlpack__()

    def __dlpack_device__(self):
        return self._t.__dlpack_device__()


def test_problem_major_and_dlpack_inputs_on_a_side_stream_match_the_oracle():
    """the binding a vectorised environment uses: [n, dof] state, any DLPack producer, the caller's
    current (non-default) stream -- inputs are still being written by a kernel on that stream when
    the planner is called; results equal the joint-major path bit for bit and the oracle"""
    lim, n, H = W.FRANKA7, 3000, 300
    ltp = _planner(lim)
    qg, q0, v0, a0 = W.random_states(lim, n, 91)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        # produced on the side stream right before the call (no synchronisation in between)
        pm = [torch.from_numpy(np.ascontiguousarray(x)).cuda(non_blocking=True) for x in (qg, q0, v0, a0)]
        pm = [(t * 2.0) * 0.5 for t in pm]
        sol, traj, jm_in = ltp.planEnvs(*[_DlpackOnly(t) for t in pm], horizon=H)
        stats_q = traj.q.sum(dim=0)          # consumer on the same stream
    side.synchronize()
    ref_jm = [torch.from_numpy(jm(x)).cuda() for x in (qg, q0, v0, a0)]
    for a, b in zip(jm_in, ref_jm):
        assert torch.equal(a, b)
    sol2 = ltp.solve(*ref_jm)
    traj2 = ltp.sample(ref_jm[1], ref_jm[2], ref_jm[3], sol2, horizon=H)
    torch.cuda.synchronize()
    assert torch.equal(sol.t_scaled, sol2.t_scaled) and torch.equal(sol.traj_len, sol2.traj_len)
    for k in "qvaj":
        assert torch.equal(getattr(traj, k), getattr(traj2, k)), k
    assert torch.equal(stats_q, traj2.q.sum(dim=0))
    ref = OraclePort.from_limits(lim).solve(qg, q0, v0, a0, threads=4)
    assert np.array_equal(sol.traj_len.cpu().numpy(), ref["traj_len"])
    assert count_bad(sol.t_scaled.cpu().numpy().transpose(2, 1, 0), ref["t_scaled"]) == 0
    # the outputs are DLPack producers themselves (zero copy out)
    back = torch.from_dlpack(traj.q)
    assert back.data_ptr() == traj.q.data_ptr()
    # transpose round trip, odd shapes
    x = torch.randn(1237, 7, dtype=torch.float64, device="cuda")
    assert torch.equal(ltp.transpose(ltp.transpose(x)), x) and torch.equal(ltp.transpose(x), x.t().contiguous())


def test_env_batch_replans_in_the_callers_layout():
    from longtermplanner_b200 import EnvBatch
    lim, n, H, tick = W.FRANKA7, 512, 200, 9
    ltp = _planner(lim)
    qg, q0, v0, a0 = W.random_states(lim, n, 93)
    envs = EnvBatch(ltp, *[torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (q0, v0, a0)])
    state_jm = [torch.from_numpy(jm(x)).cuda() for x in (q0, v0, a0)]
    for step in range(4):
        goal = W.random_states(lim, n, 200 + step)[0]
        traj = envs.replan(torch.from_numpy(np.ascontiguousarray(goal)).cuda(), H)
        sol = ltp.solve(torch.from_numpy(jm(goal)).cuda(), *state_jm)
        ref = ltp.sample(*state_jm, sol, horizon=H)
        torch.cuda.synchronize()
        assert traj.q.shape == (H, n, lim.dof) and torch.equal(traj.q, ref.q) and torch.equal(traj.j, ref.j)
        envs.advance(tick)
        ltp.advance(ref, tick, *state_jm, valid=sol.reached)
        q_pm, v_pm, a_pm = envs.state()
        assert torch.equal(q_pm, state_jm[0].t().contiguous()) and torch.equal(a_pm, state_jm[2].t().contiguous())


def test_advance_respects_the_sample_capacity_and_reserve_preallocates():
    """ltp_advance_batch reads at most sample capacity - 1 (a trajectory clipped by the capacity is
    shorter than its traj_len says) and refuses a tick past the capacity when no lengths are given;
    ltp_reserve sizes the solve scratch up front so that a captured solve never allocates"""
    from longtermplanner_b200 import _capi as capi
    lim, n = W.FRANKA7, 640
    ltp = _planner(lim)
    qg, q0, v0, a0 = W.random_states(lim, n, 97)
    ins = [torch.from_numpy(jm(x)).cuda() for x in (qg, q0, v0, a0)]
    ltp.reserve(n)
    g = torch.cuda.CUDAGraph()
    sol = ltp.alloc_solution(n)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ltp.solve(*ins, out=sol)            # warm (lazy module loading must not happen inside the capture)
        with torch.cuda.graph(g, stream=s):
            ltp.solve(*ins, out=sol)        # would fail with an allocation inside the capture
        sol.traj_len.zero_()
        g.replay()
    s.synchronize()
    ref = OraclePort.from_limits(lim).solve(qg, q0, v0, a0, threads=4)
    assert np.array_equal(sol.traj_len.cpu().numpy(), ref["traj_len"])
    # exact-length trajectories clipped by a small capacity: the tick is limited to the last stored sample
    cap = 64
    short = ltp.alloc_trajectories(n, cap, "time_major")
    ltp.sample(ins[1], ins[2], ins[3], sol, out=short)
    nq, nv, na = (t.clone() for t in ins[1:])
    cs = sol.c_struct()
    rc = capi.advance_batch(ltp._h, n, 10_000, 0, cap, sol.traj_len.data_ptr(), None, short.q.data_ptr(),
                            short.v.data_ptr(), short.a.data_ptr(), nq.data_ptr(), nv.data_ptr(), na.data_ptr(),
                            torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(nq, short.q[cap - 1].T.contiguous())      # every trajectory here is longer than 64 samples
    # fixed-horizon tensors (no lengths): a tick past the capacity is an argument error, nothing is read
    rc = capi.advance_batch(ltp._h, n, cap, 0, cap, None, None, short.q.data_ptr(), short.v.data_ptr(),
                            short.a.data_ptr(), nq.data_ptr(), nv.data_ptr(), na.data_ptr(),
                            torch.cuda.current_stream().cuda_stream)
    assert rc == capi.ERR_ARG if hasattr(capi, "ERR_ARG") else rc != 0
    del cs
