"""CPU suite, part 2: the product's device math (csrc/ltp_math.cuh) compiled for the host
by tests/host_shadow.cc, against the oracle. This catches formula/branch errors without a
GPU; the real parity tests are the `-m gpu` ones, which call the CUDA kernels through the
C ABI. The shadow library is a test artefact, never loaded by the product."""
import os
import subprocess

import numpy as np
import pytest

from helpers import bitdiff, count_bad
from longtermplanner_b200 import workloads as W
from oracle.bindings import OraclePort

HERE = os.path.dirname(os.path.abspath(__file__))
SHADOW = os.path.join(HERE, "_build", "libltp_shadow.so")


@pytest.fixture(scope="module", autouse=True)
def _shadow_lib():
    os.makedirs(os.path.dirname(SHADOW), exist_ok=True)
    src = os.path.join(HERE, "host_shadow.cc")
    hdr = os.path.join(HERE, "..", "longtermplanner_b200", "csrc", "ltp_math.cuh")
    if (not os.path.exists(SHADOW)) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(SHADOW):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", src, "-o", SHADOW],
                       check=True)


class Shadow(OraclePort):
    prefix = "shadow_"
    libname = os.path.join("..", "..", "tests", "_build", "libltp_shadow.so")


class ShadowAuto(Shadow):
    """closed-form pass + deferral to the generic sequence (what LTP_SOLVE_AUTO does)"""

    def _fn(self, name, restype=None):
        if name == "solve_batch":
            import ctypes
            f = getattr(self.lib, "shadow_solve_batch_auto")
            f.restype = ctypes.c_int64

            def call(*a):
                self.deferred = f(*a)
            return call
        return super()._fn(name, restype)


@pytest.mark.parametrize("lim,n,seed", [(W.FRANKA7, 30_000, 221), (W.REF_RANDOM6, 30_000, 222), (W.REF_GRID, 50_000, 223)])
def test_closed_form_pass_plus_deferral_equals_generic(lim, n, seed):
    qg, q0, v0, a0 = W.random_states(lim, n, seed)
    gen = Shadow.from_limits(lim).solve(qg, q0, v0, a0)
    S = ShadowAuto.from_limits(lim)
    auto = S.solve(qg, q0, v0, a0)
    for k in gen:
        a, b = gen[k], auto[k]
        assert np.array_equal(a, b, equal_nan=True) if a.dtype.kind == "f" else np.array_equal(a, b), k
    assert 0 <= S.deferred < n
    if lim is W.FRANKA7:  # realistic limits: the root solver is essentially never needed
        assert S.deferred < 0.01 * n


@pytest.mark.parametrize("lim,n,seed", [(W.FRANKA7, 40_000, 201), (W.FRANKA12, 10_000, 202), (W.REF_RANDOM6, 40_000, 203)])
def test_device_math_solve_matches_oracle(lim, n, seed):
    qg, q0, v0, a0 = W.random_states(lim, n, seed)
    ref = OraclePort.from_limits(lim).solve(qg, q0, v0, a0, threads=4)
    got = Shadow.from_limits(lim).solve(qg, q0, v0, a0)
    for k in ("dir", "mod", "opt_case", "ts_case", "final_case", "slowest", "traj_len", "reached"):
        assert np.array_equal(got[k], ref[k]), k
    for k in ("t_opt", "t_scaled", "v_drive"):
        assert count_bad(got[k], ref[k]) == 0, k
        # the only legitimate source of bit differences is glibc pow() vs the correctly
        # rounded power: a handful per million values
        assert bitdiff(got[k], ref[k]) < 1e-3 * ref[k].size, k


@pytest.mark.parametrize("lim,n,seed", [(W.FRANKA7, 30_000, 301), (W.REF_RANDOM6, 30_000, 302)])
def test_device_math_on_controller_like_states(lim, n, seed):
    """joints holding position, goals inside the brake-only window, tiny moves, states on a
    limit (workloads.edge_states): closed-form pass + deferral vs the oracle, all fields"""
    qg, q0, v0, a0 = W.edge_states(lim, n, seed)
    ref = OraclePort.from_limits(lim).solve(qg, q0, v0, a0, threads=4)
    got = ShadowAuto.from_limits(lim).solve(qg, q0, v0, a0)
    for k in ("dir", "mod", "opt_case", "ts_case", "final_case", "slowest", "traj_len", "reached"):
        assert np.array_equal(got[k], ref[k]), k
    for k in ("t_opt", "t_scaled", "v_drive"):
        assert count_bad(got[k], ref[k]) == 0, k
    assert (ref["ts_case"] == 9).mean() > 0.2  # the brake-only shortcut is what this exercises


def test_device_math_sampler_is_bit_exact():
    for lim, n, seed in ((W.FRANKA7, 40, 211), (W.REF_RANDOM6, 300, 212)):
        qg, q0, v0, a0 = W.random_states(lim, n, seed)
        P, S = OraclePort.from_limits(lim), Shadow.from_limits(lim)
        s = P.solve(qg, q0, v0, a0)
        for i in range(n):
            a = P.get_trajectory(s["t_scaled"][i], s["dir"][i], s["mod"][i], q0[i], v0[i], a0[i], s["v_drive"][i])
            b = S.get_trajectory(s["t_scaled"][i], s["dir"][i], s["mod"][i], q0[i], v0[i], a0[i], s["v_drive"][i])
            assert a["length"] == b["length"]
            for k in "qvaj":
                assert bitdiff(b[k], a[k]) == 0, (lim.name, i, k)


def test_closed_form_jump_to_the_end_of_a_row():
    """peek_position (used for the limit check of clipped rows, with a 1e-9 guard band) agrees
    with the stepped recurrence to 1e-11 from any start sample"""
    import ctypes
    f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
    u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
    for lim, n, seed in ((W.FRANKA7, 60, 231), (W.REF_RANDOM6, 200, 232)):
        qg, q0, v0, a0 = W.random_states(lim, n, seed)
        P, S = OraclePort.from_limits(lim), Shadow.from_limits(lim)
        fn = S.lib.shadow_peek_error
        fn.restype = ctypes.c_double
        fn.argtypes = [ctypes.c_void_p, f64p, f64p, u8p, f64p, f64p, f64p, f64p]
        s = P.solve(qg, q0, v0, a0)
        worst = 0.0
        for i in range(n):
            if not s["reached"][i]:
                continue
            worst = max(worst, fn(S.h, np.ascontiguousarray(s["t_scaled"][i]), np.ascontiguousarray(s["dir"][i]),
                                  np.ascontiguousarray(s["mod"][i]), np.ascontiguousarray(q0[i]),
                                  np.ascontiguousarray(v0[i]), np.ascontiguousarray(a0[i]),
                                  np.ascontiguousarray(s["v_drive"][i])))
        assert worst < 1e-11, (lim.name, worst)


def test_device_math_on_reference_time_scaling_grid():
    lim = W.REF_GRID
    P, S = OraclePort.from_limits(lim), Shadow.from_limits(lim)
    qg, v0, a0 = W.reference_grid_points(True)
    sel = np.arange(0, len(qg), 5)
    qg, v0, a0 = qg[sel], v0[sel], a0[sel]
    q0, vd = np.full_like(qg, 0.5), np.full_like(qg, 1.0)
    a, b = P.opt_switch_times(qg, q0, v0, a0, vd, threads=4), S.opt_switch_times(qg, q0, v0, a0, vd)
    for k in ("dir", "mod", "case", "ok"):
        assert np.array_equal(a[k], b[k]), k
    assert count_bad(b["t"], a["t"]) == 0
    for inc in (0.05, 0.5, 2.0):
        tr = a["t"][:, 6] + inc
        x = P.time_scaling(qg, q0, v0, a0, a["dir"], tr, threads=4)
        y = S.time_scaling(qg, q0, v0, a0, a["dir"], tr)
        for k in ("mod", "ts_case", "final_case", "ok"):
            assert np.array_equal(x[k], y[k]), (inc, k)
        assert count_bad(y["t"], x["t"]) == 0 and count_bad(y["v_drive"], x["v_drive"]) == 0


@pytest.mark.parametrize("dof", [1, 3, 8, 17, 32])
def test_device_math_on_random_limit_sets(dof):
    """the device arithmetic (compiled for the host) against the oracle on random limit sets:
    exact fields equal and values within tolerance wherever the reference produces a plan"""
    lim = W.random_limits(dof, 1000 + dof)
    n = 3000 if dof <= 8 else 800
    qg, q0, v0, a0 = W.random_states(lim, n, 4000 + dof)
    ref = OraclePort.from_limits(lim).solve(qg, q0, v0, a0, threads=4)
    got = Shadow.from_limits(lim).solve(qg, q0, v0, a0)
    r = ref["reached"].astype(bool)
    assert np.array_equal(got["reached"], ref["reached"]) and np.array_equal(got["traj_len"], ref["traj_len"])
    assert r.sum() > n // 2
    for k in ("dir", "mod", "ts_case", "final_case", "slowest"):
        assert np.array_equal(got[k][r], ref[k][r]), k
    for k in ("t_scaled", "v_drive"):
        assert count_bad(got[k][r], ref[k][r]) == 0, k



def test_device_math_with_every_root_finder_candidate_accepted():
    """a one-joint limit set under which candidates 3..8 of the search (quartics, quintic, sextic)
    are each the accepted attempt of some joint: device math (host build) == oracle, exact fields
    and values"""
    lim = W.random_limits(1, 1001)
    n = 200_000
    qg, q0, v0, a0 = (x[:, 0].copy() for x in W.random_states(lim, n, 78))
    P, S = OraclePort.from_limits(lim), Shadow.from_limits(lim)
    o = P.opt_switch_times(qg, q0, v0, a0, np.full(n, lim.v_max[0]), threads=4)
    seen = np.zeros(10, np.int64)
    for inc in (0.02, 0.05, 0.2):
        tr = o["t"][:, 6] + inc
        x = P.time_scaling(qg, q0, v0, a0, o["dir"], tr, threads=4)
        y = S.time_scaling(qg, q0, v0, a0, o["dir"], tr)
        for k in ("ok", "mod", "ts_case", "final_case"):
            assert np.array_equal(x[k], y[k]), (inc, k)
        assert count_bad(y["t"], x["t"]) == 0 and count_bad(y["v_drive"], x["v_drive"]) == 0
        seen += np.bincount(x["ts_case"], minlength=10)[:10]
    assert all(seen[k] > 0 for k in range(1, 9)), seen


class ShadowItems(Shadow):
    """item mode (tail items -> pending problems -> search items), replayed on the host"""

    def _fn(self, name, restype=None):
        if name == "solve_batch":
            import ctypes
            f = getattr(self.lib, "shadow_solve_batch_items")
            f.restype = ctypes.c_int64
            self.stats = np.zeros(4, np.int64)

            def call(*a):
                self.whole = f(*a[:-1], ctypes.c_void_p(self.stats.ctypes.data))
            return call
        return super()._fn(name, restype)


@pytest.mark.parametrize("lim,n,seed,kind", [
    (W.REF_RANDOM6, 40_000, 251, "random"), (W.REF_GRID, 60_000, 252, "random"), (W.FRANKA7, 30_000, 253, "random"),
    (W.REF_RANDOM6, 30_000, 254, "edge"), (W.FRANKA7, 30_000, 255, "edge"), (W.random_limits(5, 27), 30_000, 256, "random")])
def test_item_mode_hand_overs_equal_the_every_branch_sequence(lim, n, seed, kind):
    """what ltp_solve_batch does with large batches -- joints that need a polynomial root handed on
    one by one (tail items, pending problems, search items) -- gives, with the device's own per-joint
    functions, exactly what the every-branch sequence gives, and every list is exercised"""
    qg, q0, v0, a0 = W.random_states(lim, n, seed) if kind == "random" else W.edge_states(lim, n, seed)
    gen = Shadow.from_limits(lim).solve(qg, q0, v0, a0)
    S = ShadowItems.from_limits(lim)
    it = S.solve(qg, q0, v0, a0)
    for k in gen:
        a, b = gen[k], it[k]
        assert np.array_equal(a, b, equal_nan=True) if a.dtype.kind == "f" else np.array_equal(a, b), k
    tails, pending, searches, whole = (int(x) for x in S.stats)
    assert whole == S.whole and pending <= tails
    if lim in (W.REF_RANDOM6, W.REF_GRID) and kind == "random":
        assert tails > 0.01 * n and pending > 0
    if lim is W.REF_RANDOM6 and kind == "random":
        assert searches > 0
    if lim is W.FRANKA7 and kind == "random":
        assert tails == 0 and pending == 0


def test_division_by_prepared_reciprocal_has_the_bits_of_the_plain_division():
    """div_by (ltp_math.cuh: quotient estimate, exact remainder, one correction) against x / d for
    every divisor the limit sets of the workloads produce (a_max, j_max, v_max, 3, 12, and J^2, J^3,
    6 J^3, A J of the second cruise-speed candidate): random
    numerators over the whole exponent range plus the special values, bit for bit."""
    import ctypes
    lib = ctypes.CDLL(SHADOW)
    f = lib.shadow_div_by_mismatches
    f.restype = ctypes.c_int64
    f.argtypes = [ctypes.c_double, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    rng = np.random.default_rng(20261017)
    n = 400_000
    # mantissas uniform in [1, 2), exponents uniform over the doubles, both signs ...
    x = np.ldexp(1.0 + rng.random(n), rng.integers(-1070, 1023, n)) * rng.choice([-1.0, 1.0], n)
    # ... a block of everyday magnitudes, and the special values
    x[: n // 2] = rng.standard_normal(n // 2) * np.exp(rng.uniform(-12, 12, n // 2))
    x[-12:] = [0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, -5e-324, 2.2250738585072014e-308,
               1.7976931348623157e308, -1.7976931348623157e308, 1e-300, -1e300]
    x = np.ascontiguousarray(x)
    from fractions import Fraction
    divisors = {3.0, 12.0, 0.001, 0.004, 0.01}  # (the last three: sample times)
    for lim in (W.FRANKA7, W.FRANKA12, W.REF_RANDOM6, W.random_limits(6, 11), W.random_limits(32, 12)):
        divisors.update(lim.v_max, lim.a_max, lim.j_max)
        for a, j in zip(lim.a_max, lim.j_max):  # the divisors of the second cruise-speed candidate
            j3 = float(Fraction(j) ** 3)         # pow3() is the correctly rounded cube
            divisors.update([j * j, j3, 6 * j3, a * j])
    bad = ctypes.c_double(0.0)
    for d in sorted(divisors):
        miss = f(float(d), n, x.ctypes.data, ctypes.byref(bad))
        assert miss == 0, (d, bad.value)


@pytest.mark.parametrize("kind", ["random", "at_rest", "tiny", "on_the_limits"])
def test_deferred_range_test_of_the_divisions_changes_no_bit(kind):
    """The closed-form kernel runs stage 1 and attempt 1 with DivDeferred and repeats a flagged
    thread with DivChecked (csrc/ltp_b200.cu). Unflagged runs must carry the bits of the checked
    functions -- including starts at rest (zero numerators, either sign of zero), states on the
    limits (exact cancellation) and magnitudes near the ends of the exponent range."""
    import ctypes
    lib = ctypes.CDLL(SHADOW)
    f = lib.shadow_deferred_vs_checked
    f.restype = ctypes.c_int64
    lim = W.FRANKA7
    n = 60_000
    rng = np.random.default_rng(77)
    qg, q0, v0, a0 = [np.ascontiguousarray(x.reshape(-1)) for x in W.random_states(lim, n // lim.dof, 4242)]
    m = qg.size
    joint = np.ascontiguousarray(np.tile(np.arange(lim.dof, dtype=np.int32), m // lim.dof))
    if kind == "at_rest":
        v0 = np.where(rng.random(m) < 0.5, 0.0, -0.0)
        a0 = np.where(rng.random(m) < 0.5, 0.0, -0.0)
    elif kind == "tiny":
        scale = 10.0 ** rng.integers(-320, -100, m)
        v0 = v0 * scale
        a0 = a0 * 10.0 ** rng.integers(-320, -100, m)
    elif kind == "on_the_limits":
        a_max = np.asarray(lim.a_max)[joint]
        v_max = np.asarray(lim.v_max)[joint]
        a0 = np.where(rng.random(m) < 0.5, a_max, -a_max) * (rng.random(m) < 0.7)
        v0 = np.where(rng.random(m) < 0.3, v_max * rng.choice([-1.0, 1.0], m), v0 * 0.2)
    t_req = np.ascontiguousarray(rng.uniform(0.05, 3.0, m))
    v0, a0 = np.ascontiguousarray(v0, dtype=np.float64), np.ascontiguousarray(a0, dtype=np.float64)
    sh = Shadow(lim.dof, lim.t_sample, *lim.arrays())
    flagged = (ctypes.c_int64 * 2)(0, 0)
    f.argtypes = [ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 6 + [ctypes.c_void_p]
    miss = f(sh.h, m, joint.ctypes.data, qg.ctypes.data, q0.ctypes.data, v0.ctypes.data, a0.ctypes.data,
             t_req.ctypes.data, flagged)
    assert miss == 0
    if kind in ("random", "at_rest", "on_the_limits"):
        # the point of exempting zero numerators: these everyday cases stay on the fast path
        # (attempt 1 is flagged whenever its candidate is a NaN, which the made-up end times here
        # cause far more often than the real ones)
        assert flagged[0] < 0.001 * m, list(flagged)


_SLOW = W.Limits("slow_cruise", 0.004, (-3.0,) * 3, (3.0,) * 3, (0.05, 0.4, 1.0), (2.0, 2.0, 2.0), (4.0, 4.0, 4.0))


@pytest.mark.parametrize("lim", [W.FRANKA7, W.FRANKA12, W.REF_RANDOM6, W.REF_GRID, W.random_limits(8, 5),
                                 W.random_limits(5, 11), _SLOW], ids=lambda l: l.name)
def test_limit_only_factors_from_the_host_have_the_bits_of_the_per_item_expressions(lim):
    """The closed-form kernel for up to 8 joints reads what depends on the limits alone from
    JointLimits (derive_limits: powers of A/J, the braking half of the time-optimal solve, product
    prefixes of the first candidate, constant terms of the no-cruise radicand -- DivDeferredWide);
    the checked functions form all of it per item. Same bits for every limit set, including one
    whose cruise speed is so low that the braking half needs the cc:153 fix-up (part2v = NaN: that
    half is then evaluated per item) and joints with v_max / a_max == a_max / j_max."""
    import ctypes
    lib = ctypes.CDLL(SHADOW)
    f = lib.shadow_deferred_vs_checked
    f.restype = ctypes.c_int64
    f.argtypes = [ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 6 + [ctypes.c_void_p]
    n = 60_000 // lim.dof
    qg, q0, v0, a0 = W.random_states(lim, n, 515)
    # end times as stage 2 would hand them to the search: the slowest joint's, from the oracle
    ref = OraclePort.from_limits(lim).solve(qg, q0, v0, a0, threads=4)
    t_req = np.repeat(ref["t_opt"][:, :, 6].max(axis=1), lim.dof)
    qg, q0, v0, a0 = [np.ascontiguousarray(x.reshape(-1)) for x in (qg, q0, v0, a0)]
    m = qg.size
    joint = np.ascontiguousarray(np.tile(np.arange(lim.dof, dtype=np.int32), n))
    t_req = np.ascontiguousarray(t_req, dtype=np.float64)
    assert t_req.size == m
    sh = Shadow(lim.dof, lim.t_sample, *lim.arrays())
    flagged = (ctypes.c_int64 * 2)(0, 0)
    miss = f(sh.h, m, joint.ctypes.data, qg.ctypes.data, q0.ctypes.data, v0.ctypes.data, a0.ctypes.data,
             t_req.ctypes.data, flagged)
    assert miss == 0
    assert flagged[0] < 0.001 * m, list(flagged)


@pytest.mark.parametrize("lim", [W.FRANKA7, W.REF_RANDOM6, W.random_limits(8, 5)], ids=lambda l: l.name)
def test_second_candidate_through_reciprocals_has_the_bits_of_the_written_out_divisions(lim):
    import ctypes
    lib = ctypes.CDLL(SHADOW)
    f = lib.shadow_candidate2_mismatches
    f.restype = ctypes.c_int64
    f.argtypes = [ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 7
    rng = np.random.default_rng(5)
    qg, q0, v0, a0 = [np.ascontiguousarray(x.reshape(-1)) for x in W.random_states(lim, 40_000, 99)]
    m = qg.size
    joint = np.ascontiguousarray(np.tile(np.arange(lim.dof, dtype=np.int32), m // lim.dof))
    v0[: m // 8] = 0.0
    a0[m // 16: m // 4] = 0.0
    a0[m // 4: m // 3] = (np.asarray(lim.a_max)[joint] * rng.choice([-1.0, 1.0], m))[m // 4: m // 3]
    direction = np.ascontiguousarray(rng.choice([-1.0, 1.0], m))
    t_req = np.ascontiguousarray(rng.uniform(0.01, 5.0, m))
    sh = Shadow(lim.dof, lim.t_sample, *lim.arrays())
    assert f(sh.h, m, joint.ctypes.data, qg.ctypes.data, q0.ctypes.data, v0.ctypes.data, a0.ctypes.data,
             direction.ctypes.data, t_req.ctypes.data) == 0
