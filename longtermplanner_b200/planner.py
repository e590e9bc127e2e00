"""Python mirror of the reference's LongTermPlanner interface over the C ABI.

Same names, argument order and error behaviour as the reference class
(include/long_term_planner/long_term_planner.h:61-308 of yannickBurkhardt/LongTermPlanner):
``planTrajectory`` returns a bool and fills a ``Trajectory``; the protected per-joint
methods are exposed the way the reference's own test fixture exposes them
(tests/include/long_term_planner_fixture.h:34-57). ``planTrajectories`` is the new batched
entry point over structure-of-arrays CUDA tensors.

torch is used for device memory and streams only. Every call lands in the hand-written
kernels of csrc/ltp_b200.cu through include/ltp_b200.h; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _capi as capi


@dataclasses.dataclass
class Trajectory:
    """reference long_term_planner.h:37-45"""
    dof: int = 0
    t_sample: float = 0.0
    length: int = 0
    q: List[List[float]] = dataclasses.field(default_factory=list)
    v: List[List[float]] = dataclasses.field(default_factory=list)
    a: List[List[float]] = dataclasses.field(default_factory=list)
    j: List[List[float]] = dataclasses.field(default_factory=list)


@dataclasses.dataclass
class BatchSolution:
    """Device-resident result of stages 1-3 for n problems. `records` is what the library writes:
    [dof, n, 8] float64, one 64-byte record per (joint, problem) = the seven final switching times
    and v_drive (include/ltp_b200.h). `t_scaled` ([7, dof, n]) and `v_drive` ([dof, n]) are strided
    views of it, so code written for the round-1 field layout keeps working."""
    n: int
    dof: int
    records: torch.Tensor    # [dof, n, 8] f64
    dir: torch.Tensor        # [dof, n] f64
    mod: torch.Tensor        # [dof, n] u8
    slowest: torch.Tensor    # [n] i32
    traj_len: torch.Tensor   # [n] i32
    reached: torch.Tensor    # [n] u8
    t_opt: Optional[torch.Tensor] = None       # [7, dof, n]
    opt_case: Optional[torch.Tensor] = None    # [dof, n] u8
    ts_case: Optional[torch.Tensor] = None
    final_case: Optional[torch.Tensor] = None

    @property
    def t_scaled(self) -> torch.Tensor:
        """[7, dof, n] view: t_scaled[k, joint, problem]"""
        return self.records[:, :, :7].permute(2, 0, 1)

    @property
    def v_drive(self) -> torch.Tensor:
        """[dof, n] view"""
        return self.records[:, :, 7]

    @classmethod
    def from_fields(cls, t_scaled, dir, v_drive, mod, slowest, traj_len, reached) -> "BatchSolution":
        """pack separate [7, dof, n] times and [dof, n] v_drive tensors into records"""
        _, dof, n = t_scaled.shape
        rec = torch.empty(dof, n, 8, dtype=torch.float64, device=t_scaled.device)
        rec[:, :, :7] = t_scaled.permute(1, 2, 0)
        rec[:, :, 7] = v_drive
        return cls(n, dof, rec, dir, mod, slowest, traj_len, reached)

    def c_struct(self) -> capi.Solution:
        def ptr(t):
            return None if t is None else t.data_ptr()
        return capi.Solution(ptr(self.records), ptr(self.dir), None, ptr(self.mod),
                             ptr(self.slowest), ptr(self.traj_len), ptr(self.reached), ptr(self.t_opt),
                             ptr(self.opt_case), ptr(self.ts_case), ptr(self.final_case))


@dataclasses.dataclass
class BatchTrajectories:
    """Device-resident sampled trajectories. layout "time_major": tensors (stride, n, dof);
    layout "rows": tensors (n, dof, stride), sample-contiguous like Trajectory::q[joint]."""
    layout: str
    horizon: int
    stride: int
    q: torch.Tensor
    v: torch.Tensor
    a: torch.Tensor
    j: torch.Tensor
    success: torch.Tensor   # [n] u8
    traj_len: torch.Tensor  # [n] i32
    order: Optional[torch.Tensor] = None  # [n] i32: slot k holds problem order[k] (sorted-slot sampling)
    reached: Optional[torch.Tensor] = None  # [n] u8: the solution's flag (problems that were planned at all)


def as_tensor(x) -> torch.Tensor:
    """a torch tensor, or any DLPack producer (CuPy, JAX, Numba, Warp ... arrays): imported zero-copy.
    Everything the batched entry points return is a torch CUDA tensor, i.e. itself a DLPack producer
    (x.__dlpack__()), so results travel the other way without a copy as well."""
    if isinstance(x, torch.Tensor):
        return x
    if hasattr(x, "__dlpack__"):
        return torch.from_dlpack(x)
    raise TypeError(f"expected a torch tensor or a DLPack producer, got {type(x).__name__}")


class _DevMem:
    """a device allocation owned by someone else, exposed through __cuda_array_interface__ so that
    torch.as_tensor can alias it without copying"""

    def __init__(self, ptr: int, nbytes: int, shape, dtype):
        typestr = {torch.float64: "<f8", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2, "strides": None}


def _vec(x: Sequence[float], dof: int) -> np.ndarray:
    a = np.ascontiguousarray(x, dtype=np.float64)
    if a.shape != (dof,):
        raise ValueError(f"expected {dof} values, got shape {a.shape}")
    return a


def _np_ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class LongTermPlanner:
    def __init__(self, dof: int = 0, t_sample: float = 0.001, q_min=(), q_max=(), v_max=(), a_max=(),
                 j_max=(), device: Optional[int] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("longtermplanner_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self._h = capi.vp()
        lims = [_vec(x, dof) for x in (q_min, q_max, v_max, a_max, j_max)]
        capi.check(capi.create(C.byref(self._h), self.device, int(dof), float(t_sample),
                               *[_np_ptr(x) for x in lims]), "ltp_create")
        self.dof_, self.t_sample_ = int(dof), float(t_sample)
        self.limits_ = lims

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and capi is not None and getattr(capi, "destroy", None) is not None:  # interpreter shutdown
            capi.destroy(h)
            self._h = None

    # ---- setters (reference long_term_planner.h:176-205) --------------------------------
    def setLimits(self, q_min, q_max, v_max, a_max, j_max) -> None:
        lims = [_vec(x, self.dof_) for x in (q_min, q_max, v_max, a_max, j_max)]
        capi.check(capi.set_limits(self._h, *[_np_ptr(x) for x in lims]), "ltp_set_limits")
        self.limits_ = lims

    def setSampleTime(self, t_sample: float) -> None:
        capi.check(capi.set_sample_time(self._h, float(t_sample)), "ltp_set_sample_time")
        self.t_sample_ = float(t_sample)

    def setDoF(self, dof) -> None:
        capi.check(capi.set_dof(self._h, int(dof)), "ltp_set_dof")
        self.dof_ = int(dof)

    def setSolveMode(self, generic_only: bool) -> None:
        """validation switch: run every problem through the generic kernel (same results)"""
        capi.check(capi.set_solve_mode(self._h, 1 if generic_only else 0), "ltp_set_solve_mode")

    KERNELS = {"solve_fast": 0, "solve_generic": 1, "sample_time_major": 2, "sample_rows": 3, "solve_attempt2": 4,
               "solve_items": 5}

    def setProfiling(self, on: bool) -> None:
        """bracket every launch of the hot kernels with CUDA events on the launching stream"""
        capi.check(capi.set_profiling(self._h, 1 if on else 0), "ltp_set_profiling")

    def kernelTime(self, kernel: str, reset: bool = True):
        """-> (sum of launch durations in ms, number of launches) since the last reset"""
        ms, cnt = C.c_double(0), capi.i64(0)
        capi.check(capi.profile_read(self._h, self.KERNELS[kernel], C.byref(ms), C.byref(cnt), 1 if reset else 0),
                   "ltp_profile_read")
        return ms.value, int(cnt.value)

    @property
    def launches(self) -> int:
        return int(capi.launch_count(self._h))

    # ---- reference long_term_planner.cc:68-77 ---------------------------------------------
    def checkInputs(self, q_0, v_0, a_0) -> bool:
        q_min, q_max, v_max, a_max, j_max = self.limits_
        q_0, v_0, a_0 = (_vec(x, self.dof_) for x in (q_0, v_0, a_0))
        for i in range(self.dof_):
            if q_0[i] < q_min[i] or q_0[i] > q_max[i] or abs(v_0[i]) > v_max[i] or abs(a_0[i]) > a_max[i]:
                return False
            if abs(v_0[i] + 0.5 * a_0[i] * abs(a_0[i]) / j_max[i]) > v_max[i]:
                return False
        return True

    # ---- reference long_term_planner.cc:7-63 ----------------------------------------------
    def planTrajectory(self, q_goal, q_0, v_0, a_0, traj: Trajectory) -> bool:
        dof = self.dof_
        ins = [_vec(x, dof) for x in (q_goal, q_0, v_0, a_0)]
        # latency path: read the rows straight out of the planner's pinned staging block
        view, pitch, vlen, vok = (capi.vp * 4)(), capi.i64(0), capi.i32(0), C.c_uint8(0)
        rc = capi.plan_one_view(self._h, *[_np_ptr(x) for x in ins], C.byref(view), C.byref(pitch), C.byref(vlen),
                                C.byref(vok))
        if rc == capi.LTP_OK:
            n = int(vlen.value)
            if n <= 0:
                return False  # early `return false` of the reference: traj untouched
            traj.dof, traj.t_sample, traj.length = dof, self.t_sample_, n
            fields = []
            for f in range(4):
                buf = (C.c_double * (dof * pitch.value)).from_address(view[f])
                fields.append(np.frombuffer(buf, dtype=np.float64).reshape(dof, pitch.value)[:, :n].tolist())
            traj.q, traj.v, traj.a, traj.j = fields
            return bool(vok.value)
        if rc != capi.LTP_ERR_CAPACITY:
            capi.check(rc, "ltp_plan_one_view")
        cap = 4096
        while True:
            rows = [np.empty((dof, cap)) for _ in range(4)]
            ln = np.zeros(1, np.int32)
            ok = np.zeros(1, np.uint8)
            needed = capi.i64(0)
            rc = capi.plan_host(self._h, 1, *[_np_ptr(x) for x in ins], 0, cap, *[_np_ptr(r) for r in rows],
                                _np_ptr(ln), _np_ptr(ok), C.byref(needed))
            if rc == capi.LTP_ERR_CAPACITY:
                cap = int(needed.value)
                continue
            capi.check(rc, "ltp_plan_host")
            break
        n = int(ln[0])
        if n <= 0:
            return False  # early `return false` of the reference: traj untouched
        traj.dof, traj.t_sample, traj.length = dof, self.t_sample_, n
        traj.q, traj.v, traj.a, traj.j = (r[:, :n].tolist() for r in rows)
        return bool(ok[0])

    # ---- protected methods of the reference, single item ----------------------------------
    def optBraking(self, joint: int, v_0: float, a_0: float):
        """-> (True, q, t_rel[0..2], dir)   reference long_term_planner.cc:650-701"""
        q, d = C.c_double(), C.c_double()
        t3 = np.zeros(3)
        capi.check(capi.opt_braking_host(self._h, joint, v_0, a_0, C.byref(q), _np_ptr(t3), C.byref(d)))
        return True, q.value, t3, d.value

    def optSwitchTimes(self, joint, q_goal, q_0, v_0, a_0, v_drive):
        """-> (success, t[7], dir, mod_jerk_profile)   reference long_term_planner.cc:82-353"""
        t7 = np.zeros(7)
        d = C.c_double()
        mod, kase, ok = C.c_uint8(), C.c_uint8(), C.c_uint8()
        capi.check(capi.opt_switch_times_host(self._h, joint, q_goal, q_0, v_0, a_0, v_drive, _np_ptr(t7),
                                              C.byref(d), C.byref(mod), C.byref(kase), C.byref(ok)))
        return bool(ok.value), t7, d.value, int(mod.value)

    def timeScaling(self, joint, q_goal, q_0, v_0, a_0, dir, t_required):
        """-> (success, scaled_t[7], v_drive, mod_jerk_profile)   reference cc:358-645"""
        t7 = np.zeros(7)
        vd = C.c_double()
        mod, tsc, ok = C.c_uint8(), C.c_uint8(), C.c_uint8()
        capi.check(capi.time_scaling_host(self._h, joint, q_goal, q_0, v_0, a_0, dir, t_required, _np_ptr(t7),
                                          C.byref(vd), C.byref(mod), C.byref(tsc), C.byref(ok)))
        return bool(ok.value), t7, vd.value, int(mod.value)

    def getTrajectory(self, t, dir, mod_jerk_profile, q_0, v_0, a_0, v_drive) -> Trajectory:
        """reference long_term_planner.cc:706-841"""
        dof = self.dof_
        t7 = np.ascontiguousarray(t, dtype=np.float64).reshape(dof, 7)
        d, q0, v0, a0, vd = (_vec(x, dof) for x in (dir, q_0, v_0, a_0, v_drive))
        mod = np.ascontiguousarray(mod_jerk_profile, dtype=np.uint8).reshape(dof)
        cap = 4096
        while True:
            rows = [np.empty((dof, cap)) for _ in range(4)]
            ln = C.c_int32(0)
            needed = capi.i64(0)
            rc = capi.get_trajectory_host(self._h, _np_ptr(t7), _np_ptr(d), _np_ptr(mod), _np_ptr(q0),
                                          _np_ptr(v0), _np_ptr(a0), _np_ptr(vd), cap,
                                          *[_np_ptr(r) for r in rows], C.byref(ln), C.byref(needed))
            if rc == capi.LTP_ERR_CAPACITY:
                cap = int(needed.value)
                continue
            capi.check(rc, "ltp_get_trajectory_host")
            break
        n = int(ln.value)
        tr = Trajectory(dof=dof, t_sample=self.t_sample_, length=n)
        tr.q, tr.v, tr.a, tr.j = (r[:, :n].tolist() for r in rows)
        return tr

    # ---- batched entry points over SoA device buffers --------------------------------------
    def _chk(self, t, n: int, name: str) -> torch.Tensor:
        t = as_tensor(t)
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and tuple(t.shape) == (self.dof_, n)):
            raise ValueError(f"{name}: expected contiguous float64 CUDA tensor of shape ({self.dof_}, {n})")
        if t.device.index != self.device:
            raise ValueError(f"{name}: tensor lives on cuda:{t.device.index}, planner on cuda:{self.device}")
        return t

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def transpose(self, x, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[r, c] float64 CUDA matrix -> its transpose [c, r], contiguous (ltp_transpose): the bridge
        between a vectorised environment's problem-major state [n, dof] and the joint-major
        [dof, n] layout of the entry points below"""
        x = as_tensor(x)
        if not (x.is_cuda and x.dtype == torch.float64 and x.dim() == 2 and x.is_contiguous()):
            raise ValueError("transpose: expected a contiguous 2-D float64 CUDA tensor")
        r, c = x.shape
        if out is None:
            out = torch.empty(c, r, dtype=torch.float64, device=x.device)
        elif tuple(out.shape) != (c, r) or not out.is_contiguous() or out.dtype != torch.float64:
            raise ValueError("transpose: out must be a contiguous float64 tensor of the transposed shape")
        capi.check(capi.transpose(self._h, r, c, x.data_ptr(), out.data_ptr(), self._stream()), "ltp_transpose")
        return out

    def planEnvs(self, q_goal, q_0, v_0, a_0, horizon: int = 0, layout: str = "time_major"):
        """planTrajectories for callers whose state is problem-major: four [n, dof] float64 CUDA tensors
        (or DLPack producers), e.g. the observation tensors of a vectorised environment. Returns
        (BatchSolution, BatchTrajectories, joint-major copies of (q_goal, q_0, v_0, a_0)); the
        time-major trajectories are (samples, n, dof), i.e. already in the caller's layout."""
        jm = [self.transpose(x) for x in (q_goal, q_0, v_0, a_0)]
        sol = self.solve(*jm)
        return sol, self.sample(jm[1], jm[2], jm[3], sol, horizon, layout=layout), jm

    def reserve(self, n: int) -> None:
        """allocate the solve scratch for up to n problems now (ltp_reserve), so that no later
        solve allocates -- needed before capturing solves of a growing batch into a CUDA graph"""
        capi.check(capi.reserve(self._h, int(n)), "ltp_reserve")

    def alloc_solution(self, n: int, with_opt: bool = False, with_cases: bool = False) -> BatchSolution:
        dev = torch.device("cuda", self.device)
        dof = self.dof_
        f = lambda *s: torch.empty(*s, dtype=torch.float64, device=dev)
        b = lambda *s: torch.empty(*s, dtype=torch.uint8, device=dev)
        i = lambda *s: torch.empty(*s, dtype=torch.int32, device=dev)
        return BatchSolution(n, dof, f(dof, n, 8), f(dof, n), b(dof, n), i(n), i(n), b(n),
                             f(7, dof, n) if with_opt else None,
                             b(dof, n) if with_cases else None, b(dof, n) if with_cases else None,
                             b(dof, n) if with_cases else None)

    def solve(self, q_goal, q_0, v_0, a_0, out: Optional[BatchSolution] = None, with_opt=False,
              with_cases=False) -> BatchSolution:
        """stages 1-3 for n problems; inputs [dof, n] float64 CUDA tensors."""
        n = q_goal.shape[1]
        ins = [self._chk(t, n, nm) for t, nm in zip((q_goal, q_0, v_0, a_0), ("q_goal", "q_0", "v_0", "a_0"))]
        sol = out if out is not None else self.alloc_solution(n, with_opt, with_cases)
        cs = sol.c_struct()
        capi.check(capi.solve_batch(self._h, n, *[t.data_ptr() for t in ins], C.byref(cs), self._stream()),
                   "ltp_solve_batch")
        return sol

    def alloc_trajectories(self, n: int, samples: int, layout: str = "time_major") -> BatchTrajectories:
        """time_major: (samples, n, dof) tensors. Full store bandwidth needs n * dof to be a
        multiple of 32 (include/ltp_b200.h): pad the batch to a multiple of 32 problems."""
        dev = torch.device("cuda", self.device)
        if layout == "time_major":
            stride = samples
            shape = (stride, n, self.dof_)
        elif layout == "rows":
            stride = (samples + 3) // 4 * 4
            shape = (n, self.dof_, stride)
        else:
            raise ValueError(layout)
        rows = [torch.empty(shape, dtype=torch.float64, device=dev) for _ in range(4)]
        # 4-byte aligned, padded flag array (the kernel clears flags with word atomics)
        succ = torch.empty((n + 3) // 4 * 4, dtype=torch.uint8, device=dev)[:n]
        return BatchTrajectories(layout, 0, stride, *rows, succ, None)

    def sample(self, q_0, v_0, a_0, sol: BatchSolution, horizon: int = 0,
               out: Optional[BatchTrajectories] = None, layout: str = "time_major",
               sorted_slots: bool = False) -> BatchTrajectories:
        """stage 4. horizon = 0: exact length per problem (synchronises once to size the output
        unless `out` is given); horizon > 0: fixed number of samples per problem.
        sorted_slots (time-major, exact length): slot k of the tensors holds problem out.order[k],
        problems ordered by trajectory length on the device (ltp_sample_batch_sorted)."""
        n = sol.n
        ins = [self._chk(t, n, nm) for t, nm in zip((q_0, v_0, a_0), ("q_0", "v_0", "a_0"))]
        if out is None:
            samples = horizon if horizon > 0 else max(int(sol.traj_len.max().item()), 1)
            out = self.alloc_trajectories(n, samples, layout)
        out.horizon = horizon
        out.traj_len = sol.traj_len
        out.reached = sol.reached
        cs = sol.c_struct()
        if sorted_slots:
            if out.layout != "time_major" or horizon != 0:
                raise ValueError("sorted_slots needs the time-major layout and exact-length mode")
            out.order = torch.empty(n, dtype=torch.int32, device=out.q.device)
            capi.check(capi.sample_batch_sorted(self._h, n, *[t.data_ptr() for t in ins], C.byref(cs), out.stride,
                                                out.q.data_ptr(), out.v.data_ptr(), out.a.data_ptr(),
                                                out.j.data_ptr(), out.success.data_ptr(), out.order.data_ptr(),
                                                self._stream()), "ltp_sample_batch_sorted")
            return out
        out.order = None
        lay = capi.LAYOUT_TIME_MAJOR if out.layout == "time_major" else capi.LAYOUT_ROWS
        capi.check(capi.sample_batch(self._h, n, *[t.data_ptr() for t in ins], C.byref(cs), horizon, lay,
                                     out.stride, out.q.data_ptr(), out.v.data_ptr(), out.a.data_ptr(),
                                     out.j.data_ptr(), out.success.data_ptr(), self._stream()),
                   "ltp_sample_batch")
        return out

    def planTrajectories(self, q_goal, q_0, v_0, a_0, horizon: int = 0, layout: str = "time_major"):
        """Batched planTrajectory: -> (BatchSolution, BatchTrajectories)."""
        sol = self.solve(q_goal, q_0, v_0, a_0)
        return sol, self.sample(q_0, v_0, a_0, sol, horizon, layout=layout)

    def planStream(self, q_goal, q_0, v_0, a_0, chunk: int, horizon: int = 0, capacity: int = 4096,
                   consumer=None, sorted_slots: bool = False) -> dict:
        """planTrajectories for more problems than fit in memory at once (ltp_plan_stream): chunks
        of `chunk` problems are solved and sampled (time-major) into a two-slot ring; `consumer`,
        if given, is called per chunk as consumer(view, stream) with view a dict of CUDA tensors
        that alias the ring slot (valid for work enqueued on `stream` = a torch ExternalStream).
        sorted_slots (exact-length mode only): trajectory slot k of a chunk holds problem
        view["order"][k] (problems ordered by length on the device, ltp_set_stream_sorted); the
        solution, inputs and success flags stay indexed by problem. Returns the totals accumulated
        on the device."""
        capi.check(capi.set_stream_sorted(self._h, 1 if sorted_slots else 0), "ltp_set_stream_sorted")
        n = q_goal.shape[1]
        ins = [self._chk(t, n, nm) for t, nm in zip((q_goal, q_0, v_0, a_0), ("q_goal", "q_0", "v_0", "a_0"))]
        dof, dev = self.dof_, torch.device("cuda", self.device)
        err = []

        def _alias(ptr, shape, dtype):
            n_el = int(np.prod(shape))
            if n_el == 0 or not ptr:
                return torch.empty(shape, dtype=dtype, device=dev)
            itemsize = torch.empty((), dtype=dtype).element_size()
            arr = _DevMem(ptr, n_el * itemsize, shape, dtype)
            return torch.as_tensor(arr, device=dev)

        def _cb(user, chunk_p, stream):
            try:
                c = chunk_p.contents
                cnt, cap = int(c.count), int(c.capacity)
                sol = c.solution
                view = dict(
                    first=int(c.first), count=cnt, capacity=cap, horizon=int(c.horizon),
                    q_goal=_alias(c.q_goal, (dof, cnt), torch.float64), q_0=_alias(c.q_0, (dof, cnt), torch.float64),
                    v_0=_alias(c.v_0, (dof, cnt), torch.float64), a_0=_alias(c.a_0, (dof, cnt), torch.float64),
                    records=_alias(sol.t_scaled, (dof, cnt, 8), torch.float64),
                    traj_len=_alias(sol.traj_len, (cnt,), torch.int32),
                    reached=_alias(sol.reached, (cnt,), torch.uint8),
                    success=_alias(c.success, (cnt,), torch.uint8),
                    order=_alias(c.order, (cnt,), torch.int32) if c.order else None,
                    q=_alias(c.q, (cap, cnt, dof), torch.float64), v=_alias(c.v, (cap, cnt, dof), torch.float64),
                    a=_alias(c.a, (cap, cnt, dof), torch.float64), j=_alias(c.j, (cap, cnt, dof), torch.float64))
                view["t_scaled"] = view["records"][:, :, :7].permute(2, 0, 1)   # [7, dof, cnt] view
                view["v_drive"] = view["records"][:, :, 7]
                ext = torch.cuda.ExternalStream(int(stream), device=dev)
                with torch.cuda.stream(ext):
                    consumer(view, ext)
                return 0
            except Exception as e:  # never unwind through the C frame
                err.append(e)
                return 1

        cb = capi.CHUNK_CONSUMER(_cb) if consumer is not None else capi.CHUNK_CONSUMER()
        stats = capi.StreamStats()
        rc = capi.plan_stream(self._h, n, *[t.data_ptr() for t in ins], int(chunk), int(horizon), int(capacity),
                              cb, None, C.byref(stats), self._stream())
        if err:
            raise err[0]
        capi.check(rc, "ltp_plan_stream")
        return {k: int(getattr(stats, k)) for k, _ in capi.StreamStats._fields_}

    def advance(self, traj: BatchTrajectories, tick: int, q_0, v_0, a_0, valid: Optional[torch.Tensor] = None,
                clamp: bool = True) -> None:
        """receding horizon: the state at sample index `tick` of time-major trajectories becomes
        the next start state (written into q_0, v_0, a_0 in place), ltp_advance_batch. Problems
        that were not planned (valid, default: the `reached` flag of the solution the trajectories
        were sampled from) keep their state."""
        if valid is None:
            valid = traj.reached
        if traj.layout != "time_major":
            raise ValueError("advance needs time-major trajectories")
        n = q_0.shape[1]
        for t in (q_0, v_0, a_0):
            self._chk(t, n, "state")
        if not (0 <= tick < traj.stride):
            raise ValueError("tick outside the sampled range")
        tl = None if traj.horizon > 0 else traj.traj_len.data_ptr()
        capi.check(capi.advance_batch(self._h, n, int(tick), 1 if clamp else 0, int(traj.stride), tl,
                                      None if valid is None else valid.data_ptr(), traj.q.data_ptr(),
                                      traj.v.data_ptr(), traj.a.data_ptr(), q_0.data_ptr(), v_0.data_ptr(),
                                      a_0.data_ptr(), self._stream()), "ltp_advance_batch")

    # per-joint primitives, batched
    def optBrakingBatch(self, v_0, a_0):
        n = v_0.shape[1]
        self._chk(v_0, n, "v_0"), self._chk(a_0, n, "a_0")
        q, d = torch.empty_like(v_0), torch.empty_like(v_0)
        t_rel = torch.empty(3, self.dof_, n, dtype=torch.float64, device=v_0.device)
        capi.check(capi.opt_braking_batch(self._h, n, v_0.data_ptr(), a_0.data_ptr(), q.data_ptr(),
                                          t_rel.data_ptr(), d.data_ptr(), self._stream()))
        return dict(q=q, t_rel=t_rel, dir=d)

    def optSwitchTimesBatch(self, q_goal, q_0, v_0, a_0, v_drive):
        n = q_goal.shape[1]
        ins = [self._chk(t, n, "input") for t in (q_goal, q_0, v_0, a_0, v_drive)]
        dev = q_goal.device
        t = torch.zeros(7, self.dof_, n, dtype=torch.float64, device=dev)
        d = torch.empty_like(q_goal)
        mod, kase, ok = (torch.empty(self.dof_, n, dtype=torch.uint8, device=dev) for _ in range(3))
        capi.check(capi.opt_switch_times_batch(self._h, n, *[x.data_ptr() for x in ins], t.data_ptr(),
                                               d.data_ptr(), mod.data_ptr(), kase.data_ptr(), ok.data_ptr(),
                                               self._stream()))
        return dict(t=t, dir=d, mod=mod, case=kase, ok=ok)

    def timeScalingBatch(self, q_goal, q_0, v_0, a_0, dir, t_required):
        n = q_goal.shape[1]
        ins = [self._chk(t, n, "input") for t in (q_goal, q_0, v_0, a_0, dir, t_required)]
        dev = q_goal.device
        t = torch.zeros(7, self.dof_, n, dtype=torch.float64, device=dev)
        vd = torch.empty_like(q_goal)
        mod, tsc, fc, ok = (torch.empty(self.dof_, n, dtype=torch.uint8, device=dev) for _ in range(4))
        capi.check(capi.time_scaling_batch(self._h, n, *[x.data_ptr() for x in ins], t.data_ptr(), vd.data_ptr(),
                                           mod.data_ptr(), tsc.data_ptr(), fc.data_ptr(), ok.data_ptr(),
                                           self._stream()))
        return dict(t=t, v_drive=vd, mod=mod, ts_case=tsc, final_case=fc, ok=ok)

    # host-buffer forms (numpy, joint-major [dof, n]); copies are inside the call
    def solve_host(self, q_goal, q_0, v_0, a_0, out: Optional[dict] = None, with_opt=False,
                   with_cases=False) -> dict:
        """ltp_solve_host. `out` (optional) names the host arrays to fill -- records [dof, n, 8]
        (switching times + v_drive), dir, mod, slowest, traj_len, reached, t_opt, opt_case, ts_case,
        final_case, v_drive (a separate contiguous copy); an absent key is the output mask of the C
        call (not copied back). The returned dict also carries `t_scaled` ([7, dof, n]) and
        `v_drive` as views of the records."""
        dof = self.dof_
        ins = [np.ascontiguousarray(x, dtype=np.float64) for x in (q_goal, q_0, v_0, a_0)]
        n = ins[0].shape[1]
        if out is None:
            out = dict(records=np.empty((dof, n, 8)), dir=np.empty((dof, n)),
                       mod=np.empty((dof, n), np.uint8), slowest=np.empty(n, np.int32),
                       traj_len=np.empty(n, np.int32), reached=np.empty(n, np.uint8),
                       t_opt=np.empty((7, dof, n)) if with_opt else None,
                       opt_case=np.empty((dof, n), np.uint8) if with_cases else None,
                       ts_case=np.empty((dof, n), np.uint8) if with_cases else None,
                       final_case=np.empty((dof, n), np.uint8) if with_cases else None)
        rec = out.get("records")
        own_vd = out.get("v_drive") if (rec is None or out.get("v_drive") is None or
                                        not np.shares_memory(out["v_drive"], rec)) else None
        cs = capi.Solution(_np_ptr(rec), _np_ptr(out.get("dir")), _np_ptr(own_vd),
                           *[_np_ptr(out.get(k)) for k in ("mod", "slowest", "traj_len", "reached", "t_opt",
                                                           "opt_case", "ts_case", "final_case")])
        capi.check(capi.solve_host(self._h, n, *[_np_ptr(x) for x in ins], C.byref(cs)), "ltp_solve_host")
        if rec is not None:
            out["t_scaled"] = rec[:, :, :7].transpose(2, 0, 1)
            if own_vd is None:
                out["v_drive"] = rec[:, :, 7]
        return out
