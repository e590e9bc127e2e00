"""Synthetic inputs for the planning hot path.

Limit sets and the random start/goal recipe. The recipe follows the reference's
tests/randomConfiguration.m:14-34 (velocity inside the limits, acceleration bounded so that
the joint can still be stopped below v_max), which guarantees that the reference's
checkInputs (src/long_term_planner.cc:68-77) accepts every draw. Everything is a pure
function of (seed, problem index, joint, field), so any shard of a workload can be
regenerated on any rank, on the host (numpy) or compared against a device-side generator.

Arrays are returned problem-major, shape [n, dof]; the device layout of the C ABI is
joint-major [dof, n] (see include/ltp_b200.h), use ``to_joint_major``.
"""
from __future__ import annotations

import dataclasses
import numpy as np


@dataclasses.dataclass(frozen=True)
class Limits:
    name: str
    t_sample: float
    q_min: tuple
    q_max: tuple
    v_max: tuple
    a_max: tuple
    j_max: tuple

    @property
    def dof(self) -> int:
        return len(self.q_min)

    def arrays(self):
        return tuple(np.asarray(x, dtype=np.float64) for x in
                     (self.q_min, self.q_max, self.v_max, self.a_max, self.j_max))


# Franka Emika FCI datasheet values (public; NOT from the reference, which ships toy limits)
_FR_QMIN = (-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973)
_FR_QMAX = (2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973)
_FR_V = (2.175, 2.175, 2.175, 2.175, 2.61, 2.61, 2.61)
_FR_A = (15.0, 7.5, 10.0, 12.5, 15.0, 20.0, 20.0)
_FR_J = (7500.0, 3750.0, 5000.0, 6250.0, 7500.0, 10000.0, 10000.0)

FRANKA7 = Limits("franka7", 0.001, _FR_QMIN, _FR_QMAX, _FR_V, _FR_A, _FR_J)
# dual arm: the first six joints of FRANKA7, twice
FRANKA12 = Limits("franka12", 0.001, _FR_QMIN[:6] * 2, _FR_QMAX[:6] * 2, _FR_V[:6] * 2,
                  _FR_A[:6] * 2, _FR_J[:6] * 2)
# reference tests/src/long_term_planner_tests.cc:331-336 (GridTimeScalingTest), README.md:129-131
REF_GRID = Limits("ref_grid", 0.004, (-6.0,), (7.0,), (1.0,), (2.0,), (15.0,))
# reference tests/include/long_term_planner_fixture.h:73-79
REF_UNIT = Limits("ref_unit", 0.001, (-3.1,), (3.1,), (10.0,), (2.0,), (4.0,))
# reference tests/randomConfiguration.m:4-8 (6 joints, v 1, a 2, j 15, Ts 4 ms)
REF_RANDOM6 = Limits("ref_random6", 0.004, (-3.14,) * 6, (3.14,) * 6, (1.0,) * 6, (2.0,) * 6,
                     (15.0,) * 6)

def random_limits(dof: int, seed: int) -> Limits:
    """A random limit set for tests: mixed ratios a_max / j_max (1.7 ... 250 ms), three sample
    times, asymmetric joint ranges."""
    r = np.random.default_rng(seed)
    half = r.uniform(1.0, 3.1, dof)
    centre = r.uniform(-0.5, 0.5, dof)
    v = r.uniform(0.5, 3.0, dof)
    a = r.uniform(1.0, 20.0, dof)
    j = a * np.exp(r.uniform(np.log(4.0), np.log(600.0), dof))
    ts = float(r.choice([0.001, 0.004, 0.01]))
    return Limits(f"rand{dof}", ts, tuple(centre - half), tuple(centre + half), tuple(v), tuple(a), tuple(j))


SEEDS = {1: 0xB2000001, 2: 0xB2000002, 3: 0xB2000003, 4: 0xB2000004, 5: 0xB2000005}

_GOLD = np.uint64(0x9E3779B97F4A7C15)


def splitmix64(counter: np.ndarray, seed: int) -> np.ndarray:
    """k-th output of the splitmix64 stream started at ``seed`` (k = counter, uint64)."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (counter.astype(np.uint64) + np.uint64(1)) * _GOLD
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def uniform01(counter: np.ndarray, seed: int) -> np.ndarray:
    """53-bit uniform in [0, 1)."""
    return (splitmix64(counter, seed) >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def random_states(lim: Limits, n: int, seed: int, start: int = 0, margin: float = 0.05):
    """(q_goal, q_0, v_0, a_0), each [n, dof] float64, for problems start .. start+n-1.

    counter k = ((problem * dof + joint) * 4 + field), field 0..3 = q_0, q_goal, v_0, a_0.
    """
    dof = lim.dof
    q_min, q_max, v_max, a_max, j_max = lim.arrays()
    e = 1e-6
    p = np.arange(start, start + n, dtype=np.uint64)[:, None]
    j = np.arange(dof, dtype=np.uint64)[None, :]
    base = (p * np.uint64(dof) + j) * np.uint64(4)
    u = [uniform01(base + np.uint64(f), seed) for f in range(4)]
    q_0 = q_min + u[0] * (q_max - q_min)
    q_goal = (q_min + margin) + u[1] * ((q_max - margin) - (q_min + margin))
    v_0 = -(v_max - e) + u[2] * (2.0 * (v_max - e))
    pos = v_0 >= 0
    root = np.sqrt(2.0 * j_max * (v_max - np.abs(v_0)))
    a_lb = np.where(pos, -(a_max - e), np.maximum(-(a_max - e), -root))
    a_ub = np.where(pos, np.minimum(a_max - e, root), a_max)
    a_0 = a_lb + u[3] * (a_ub - a_lb)
    # the very top of the a-range can violate |a_0| <= a_max by rounding; keep it inside
    a_0 = np.clip(a_0, -a_max, a_max)
    return q_goal, q_0, v_0, a_0


def edge_states(lim: Limits, n: int, seed: int):
    """Random states with the situations a replanning controller produces mixed in per joint
    (the reference's own tests only visit them on a single joint, tests.cc:111-196): joints that
    hold position (goal = start, at rest), goals inside the 4e-3 brake-only window (cc:102),
    tiny moves with tiny start velocities/accelerations (1e-9 .. 1e-3), joints at a velocity or
    position limit, and whole problems at rest at their goal. Same return convention as
    random_states."""
    q_goal, q_0, v_0, a_0 = (x.copy() for x in random_states(lim, n, seed))
    q_min, q_max, v_max, a_max, j_max = lim.arrays()
    rng = np.random.default_rng(seed)
    kind = rng.random((n, lim.dof))
    hold = kind < 0.30
    window = (kind >= 0.30) & (kind < 0.40)
    tiny = (kind >= 0.40) & (kind < 0.50)
    at_v = (kind >= 0.50) & (kind < 0.55)
    at_q = (kind >= 0.55) & (kind < 0.58)
    v_0[hold | window] = 0.0
    a_0[hold | window] = 0.0
    q_goal[hold] = q_0[hold]
    q_goal[window] = (q_0 + rng.uniform(-3.9e-3, 3.9e-3, q_0.shape))[window]
    mag = 10.0 ** rng.uniform(-9, -3, q_0.shape)
    v_0[tiny] = (mag * rng.choice([-1.0, 1.0], q_0.shape))[tiny]
    a_0[tiny] = (mag[:, ::-1] * rng.choice([-1.0, 1.0], q_0.shape))[tiny]
    q_goal[tiny] = (q_0 + rng.uniform(-0.05, 0.05, q_0.shape))[tiny]
    v_0[at_v] = (np.broadcast_to(v_max, q_0.shape) * rng.choice([-1.0, 1.0], q_0.shape))[at_v]
    a_0[at_v] = 0.0
    q_0[at_q] = np.broadcast_to(q_min, q_0.shape)[at_q]
    v_0[at_q] = np.abs(v_0[at_q]) * 0.1
    a_0[at_q] = 0.0
    rest = rng.random(n) < 0.02
    v_0[rest] = 0.0
    a_0[rest] = 0.0
    q_goal[rest] = q_0[rest]
    q_goal = np.clip(q_goal, q_min, q_max)
    return q_goal, q_0, v_0, a_0


def to_joint_major(x: np.ndarray) -> np.ndarray:
    """[n, dof] -> contiguous [dof, n]."""
    return np.ascontiguousarray(x.T)


def grid_one_joint(m: int = 256, lim: Limits = REF_GRID, q_0: float = 0.5):
    """m^3 single-joint grid over (q_goal, v_0, a_0): the refinement of the reference's
    GridTimeScalingTest loops (tests/src/long_term_planner_tests.cc:345-363) described in
    SURVEY.md 8d, config 4. Returns flat arrays of m^3 points (q_goal, q_0, v_0, a_0)."""
    q_min, q_max, v_max, a_max, j_max = (x[0] for x in lim.arrays())
    e = 1e-6
    qg = np.linspace(q_min, q_max, m)
    v0 = np.linspace(-(v_max - e), v_max - e, m)
    uu = np.linspace(0.0, 1.0, m)
    QG, V0, U = np.meshgrid(qg, v0, uu, indexing="ij")
    root = np.sqrt(2.0 * j_max * (v_max - np.abs(V0)))
    a_lb = np.where(V0 >= 0, -(a_max - e), np.maximum(-(a_max - e), -root))
    a_ub = np.where(V0 >= 0, np.minimum(a_max - e, root), a_max)
    A0 = np.clip(a_lb + U * (a_ub - a_lb), -a_max, a_max)
    Q0 = np.full_like(QG, q_0)
    return QG.ravel(), Q0.ravel(), V0.ravel(), A0.ravel()


def _excl(bound: float) -> int:
    """exclusive integer upper bound of a C loop `k < bound` with a double bound"""
    return int(np.ceil(bound))


def reference_grid_points(time_scaling: bool):
    """The exact point sets of the reference's two C++ grid tests, including their
    integer-cast loop bounds (tests/src/long_term_planner_tests.cc:264-298 and 325-363).
    Returns (q_goal, v_0, a_0) flat arrays; q_0 = 0.5, limits REF_GRID (q range differs
    between the two tests but is irrelevant to optSwitchTimes/timeScaling)."""
    eps, step = 1e-6, 0.1
    v_max, a_max, j_max = 1.0, 2.0, 15.0
    out = []
    if not time_scaling:
        q_lo, q_hi = -3.1, 3.1
    else:
        q_lo, q_hi = -6.0, 7.0
    i0, i1 = int(int(q_lo) / step), int(int(q_hi) / step)  # (int)q_min[0]/step -> int/double -> int
    for i in range(i0, i1 + 1):
        q_goal = i * step
        for j in range(int(int(-v_max) / step), _excl(int(v_max) / step)):
            v_0 = j * step
            if time_scaling:
                v_0 = v_0 - eps if j > 0 else v_0 + eps
            if v_0 >= 0:
                a_lb = -(a_max - eps)
                a_ub = min(a_max - eps, np.sqrt(2 * j_max * (v_max - v_0)))
            else:
                a_lb = max(-(a_max - eps), -np.sqrt(2 * j_max * (v_max - abs(v_0))))
                a_ub = a_max
            if not time_scaling:
                for k in range(int(int(a_lb) / step), _excl(int(a_ub) / step)):
                    out.append((q_goal, v_0, k * step - eps))
            else:
                n_steps = int(np.floor((a_ub - a_lb) / step))
                for k in range(n_steps):
                    out.append((q_goal, v_0, a_lb + k * step))
    arr = np.asarray(out, dtype=np.float64)
    return arr[:, 0].copy(), arr[:, 1].copy(), arr[:, 2].copy()
