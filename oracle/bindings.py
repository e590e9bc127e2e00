"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the two CPU checkers.

  OraclePort  -> oracle/_ref/libltp_oracle.so  (plain-C restatement, oracle/ltp_oracle.c)
  Reference   -> oracle/_ref/libltp_ref.so     (the reference's unmodified .cc + Eigen shim)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module. Array convention: problem-major, x[p, joint]; times [p, joint, 7].
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_OUT = os.path.join(_HERE, "_ref")


def build(quiet: bool = True) -> None:
    """Compile the checkers (the port always; the reference build only when /root/reference
    is present -- on the GPU box the prebuilt .so shipped with the snapshot is used)."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(x, shape=None):
    a = np.ascontiguousarray(x, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


class _Base:
    prefix = ""
    libname = ""

    def __init__(self, dof, t_sample, q_min, q_max, v_max, a_max, j_max):
        path = os.path.join(_OUT, self.libname)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` (oracle.bindings.build())")
        self.lib = C.CDLL(path)
        self.dof, self.t_sample = int(dof), float(t_sample)
        self.limits = [_f64(x, (dof,)) for x in (q_min, q_max, v_max, a_max, j_max)]
        create = getattr(self.lib, self.prefix + "create")
        create.restype = C.c_void_p
        self.h = C.c_void_p(create(C.c_int(self.dof), C.c_double(self.t_sample), *[_vp(x) for x in self.limits]))

    @classmethod
    def from_limits(cls, lim):
        return cls(lim.dof, lim.t_sample, *lim.arrays())

    @classmethod
    def available(cls) -> bool:
        return os.path.exists(os.path.join(_OUT, cls.libname))

    def __del__(self):
        try:
            d = getattr(self.lib, self.prefix + "destroy")
            d.restype = None
            d(self.h)
        except Exception:
            pass

    def _fn(self, name, restype=None):
        f = getattr(self.lib, self.prefix + name)
        f.restype = restype
        return f


class OraclePort(_Base):
    """oracle/ltp_oracle.c"""
    prefix = "ltpo_"
    libname = "libltp_oracle.so"
    kind = "port"

    def roots(self, coeffs):
        c = _f64(coeffs)
        deg = len(c) - 1
        re, im = np.zeros(deg), np.zeros(deg)
        r = self._fn("roots", C.c_double)(_vp(c), C.c_int(deg), _vp(re), _vp(im))
        return r, re, im

    def opt_braking(self, v_0, a_0, joint=None):
        v_0, a_0 = _f64(v_0), _f64(a_0)
        n = v_0.size
        jt = None if joint is None else np.ascontiguousarray(joint, dtype=np.int32)
        q, t3, d = np.zeros(n), np.zeros((n, 3)), np.zeros(n)
        self._fn("opt_braking_items")(self.h, C.c_int64(n), _vp(jt), _vp(v_0), _vp(a_0), _vp(q), _vp(t3), _vp(d))
        return dict(q=q, t_rel=t3, dir=d)

    def opt_switch_times(self, q_goal, q_0, v_0, a_0, v_drive, joint=None, threads=1):
        q_goal, q_0, v_0, a_0, v_drive = map(_f64, (q_goal, q_0, v_0, a_0, v_drive))
        n = q_goal.size
        jt = None if joint is None else np.ascontiguousarray(joint, dtype=np.int32)
        t, d = np.zeros((n, 7)), np.zeros(n)
        mod, kase, ok = (np.zeros(n, np.uint8) for _ in range(3))
        self._fn("opt_switch_times_items")(self.h, C.c_int64(n), _vp(jt), _vp(q_goal), _vp(q_0), _vp(v_0),
                                           _vp(a_0), _vp(v_drive), _vp(t), _vp(d), _vp(mod), _vp(kase),
                                           _vp(ok), C.c_int(threads))
        return dict(t=t, dir=d, mod=mod, case=kase, ok=ok)

    def time_scaling(self, q_goal, q_0, v_0, a_0, dir, t_required, joint=None, threads=1):
        q_goal, q_0, v_0, a_0, dir, t_required = map(_f64, (q_goal, q_0, v_0, a_0, dir, t_required))
        n = q_goal.size
        jt = None if joint is None else np.ascontiguousarray(joint, dtype=np.int32)
        t, vd = np.zeros((n, 7)), np.zeros(n)
        mod, tsc, fc, ok = (np.zeros(n, np.uint8) for _ in range(4))
        self._fn("time_scaling_items")(self.h, C.c_int64(n), _vp(jt), _vp(q_goal), _vp(q_0), _vp(v_0), _vp(a_0),
                                       _vp(dir), _vp(t_required), _vp(t), _vp(vd), _vp(mod), _vp(tsc),
                                       _vp(fc), _vp(ok), C.c_int(threads))
        return dict(t=t, v_drive=vd, mod=mod, ts_case=tsc, final_case=fc, ok=ok)

    def solve(self, q_goal, q_0, v_0, a_0, threads=1):
        dof = self.dof
        q_goal, q_0, v_0, a_0 = (_f64(x).reshape(-1, dof) for x in (q_goal, q_0, v_0, a_0))
        n = q_goal.shape[0]
        t_opt, t_sc = np.zeros((n, dof, 7)), np.zeros((n, dof, 7))
        d, vd = np.zeros((n, dof)), np.zeros((n, dof))
        mod, oc, tc, fc = (np.zeros((n, dof), np.uint8) for _ in range(4))
        slowest, tl = np.zeros(n, np.int32), np.zeros(n, np.int32)
        reached = np.zeros(n, np.uint8)
        self._fn("solve_batch")(self.h, C.c_int64(n), _vp(q_goal), _vp(q_0), _vp(v_0), _vp(a_0), _vp(t_opt),
                                _vp(t_sc), _vp(d), _vp(vd), _vp(mod), _vp(oc), _vp(tc), _vp(fc),
                                _vp(slowest), _vp(tl), _vp(reached), C.c_int(threads))
        return dict(t_opt=t_opt, t_scaled=t_sc, dir=d, v_drive=vd, mod=mod, opt_case=oc, ts_case=tc,
                    final_case=fc, slowest=slowest, traj_len=tl, reached=reached)

    def get_trajectory(self, t7, dir, mod, q_0, v_0, a_0, v_drive, stride=None):
        dof = self.dof
        t7 = _f64(t7, (dof, 7))
        dir, q_0, v_0, a_0, v_drive = (_f64(x, (dof,)) for x in (dir, q_0, v_0, a_0, v_drive))
        mod = np.ascontiguousarray(mod, dtype=np.uint8).reshape(dof)
        if stride is None:
            stride = int(np.ceil(np.nanmax(t7[:, 6]) / self.t_sample)) + 2
        out = [np.zeros((dof, stride)) for _ in range(4)]
        f = self._fn("get_trajectory", C.c_int)
        ln = f(self.h, _vp(t7), _vp(dir), _vp(mod), _vp(q_0), _vp(v_0), _vp(a_0), _vp(v_drive),
               C.c_int64(stride), *[_vp(x) for x in out])
        if ln < 0:
            return self.get_trajectory(t7, dir, mod, q_0, v_0, a_0, v_drive, stride=-ln)
        return dict(length=ln, q=out[0][:, :ln], v=out[1][:, :ln], a=out[2][:, :ln], j=out[3][:, :ln])

    def plan(self, q_goal, q_0, v_0, a_0, stride=8192):
        dof = self.dof
        q_goal, q_0, v_0, a_0 = (_f64(x, (dof,)) for x in (q_goal, q_0, v_0, a_0))
        out = [np.zeros((dof, stride)) for _ in range(4)]
        ln = C.c_int(-1)
        ok = self._fn("plan", C.c_int)(self.h, _vp(q_goal), _vp(q_0), _vp(v_0), _vp(a_0), C.c_int64(stride),
                                       *[_vp(x) for x in out], C.byref(ln))
        ln = ln.value
        if ln > stride:
            return self.plan(q_goal, q_0, v_0, a_0, stride=ln)
        k = max(ln, 0)
        return dict(success=bool(ok), length=ln, q=out[0][:, :k], v=out[1][:, :k], a=out[2][:, :k], j=out[3][:, :k])

    def plan_batch(self, q_goal, q_0, v_0, a_0, threads=1):
        dof = self.dof
        q_goal, q_0, v_0, a_0 = (_f64(x).reshape(-1, dof) for x in (q_goal, q_0, v_0, a_0))
        n = q_goal.shape[0]
        ok, ln = np.zeros(n, np.uint8), np.zeros(n, np.int32)
        s = self._fn("plan_batch", C.c_double)(self.h, C.c_int64(n), _vp(q_goal), _vp(q_0), _vp(v_0), _vp(a_0),
                                               _vp(ok), _vp(ln), C.c_int(threads))
        return dict(success=ok, length=ln, checksum=s)


class Reference(_Base):
    """the reference's own long_term_planner.cc behind oracle/ref_capi.cc"""
    prefix = "ref_"
    libname = "libltp_ref.so"
    kind = "reference"
    flags = "-O2 -ffp-contract=off"

    def build_info(self):
        return self._fn("build_info", C.c_char_p)().decode()

    def roots(self, coeffs, dtype=np.float64):
        c = np.ascontiguousarray(coeffs, dtype=dtype)
        deg = len(c) - 1
        re, im = np.zeros(deg, dtype), np.zeros(deg, dtype)
        if dtype == np.float32:
            r = self._fn("roots_f32", C.c_float)(_vp(c), C.c_int(deg), _vp(re), _vp(im))
        else:
            r = self._fn("roots_f64", C.c_double)(_vp(c), C.c_int(deg), _vp(re), _vp(im))
        return r, re, im

    def opt_braking(self, v_0, a_0, joint=None):
        v_0, a_0 = _f64(v_0), _f64(a_0)
        n = v_0.size
        jt = None if joint is None else np.ascontiguousarray(joint, dtype=np.int32)
        q, t3, d = np.zeros(n), np.zeros((n, 3)), np.zeros(n)
        self._fn("opt_braking")(self.h, C.c_int64(n), _vp(jt), _vp(v_0), _vp(a_0), _vp(q), _vp(t3), _vp(d))
        return dict(q=q, t_rel=t3, dir=d)

    def opt_switch_times(self, q_goal, q_0, v_0, a_0, v_drive, joint=None, threads=1):
        q_goal, q_0, v_0, a_0, v_drive = map(_f64, (q_goal, q_0, v_0, a_0, v_drive))
        n = q_goal.size
        jt = None if joint is None else np.ascontiguousarray(joint, dtype=np.int32)
        t, d = np.zeros((n, 7)), np.zeros(n)
        mod, ok = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        self._fn("opt_switch_times")(self.h, C.c_int64(n), _vp(jt), _vp(q_goal), _vp(q_0), _vp(v_0), _vp(a_0),
                                     _vp(v_drive), _vp(t), _vp(d), _vp(mod), _vp(ok), C.c_int(threads))
        return dict(t=t, dir=d, mod=mod, ok=ok)

    def time_scaling(self, q_goal, q_0, v_0, a_0, dir, t_required, joint=None, threads=1):
        q_goal, q_0, v_0, a_0, dir, t_required = map(_f64, (q_goal, q_0, v_0, a_0, dir, t_required))
        n = q_goal.size
        jt = None if joint is None else np.ascontiguousarray(joint, dtype=np.int32)
        t, vd = np.zeros((n, 7)), np.zeros(n)
        mod, ok = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        self._fn("time_scaling")(self.h, C.c_int64(n), _vp(jt), _vp(q_goal), _vp(q_0), _vp(v_0), _vp(a_0),
                                 _vp(dir), _vp(t_required), _vp(t), _vp(vd), _vp(mod), _vp(ok), C.c_int(threads))
        return dict(t=t, v_drive=vd, mod=mod, ok=ok)

    def solve(self, q_goal, q_0, v_0, a_0, threads=1):
        dof = self.dof
        q_goal, q_0, v_0, a_0 = (_f64(x).reshape(-1, dof) for x in (q_goal, q_0, v_0, a_0))
        n = q_goal.shape[0]
        t_opt, t_sc = np.zeros((n, dof, 7)), np.zeros((n, dof, 7))
        d, vd = np.zeros((n, dof)), np.zeros((n, dof))
        mod = np.zeros((n, dof), np.uint8)
        slowest, ts_ok = np.zeros(n, np.int32), np.zeros((n, dof), np.int32)
        reached = np.zeros(n, np.uint8)
        self._fn("solve_batch")(self.h, C.c_int64(n), _vp(q_goal), _vp(q_0), _vp(v_0), _vp(a_0), _vp(t_opt),
                                _vp(t_sc), _vp(d), _vp(vd), _vp(mod), _vp(slowest), _vp(ts_ok), _vp(reached),
                                C.c_int(threads))
        return dict(t_opt=t_opt, t_scaled=t_sc, dir=d, v_drive=vd, mod=mod, slowest=slowest, ts_ok=ts_ok,
                    reached=reached)

    def get_trajectory(self, t7, dir, mod, q_0, v_0, a_0, v_drive, stride=None):
        dof = self.dof
        t7 = _f64(t7, (dof, 7))
        dir, q_0, v_0, a_0, v_drive = (_f64(x, (dof,)) for x in (dir, q_0, v_0, a_0, v_drive))
        mod = np.ascontiguousarray(mod, dtype=np.uint8).reshape(dof)
        if stride is None:
            stride = int(np.ceil(np.nanmax(t7[:, 6]) / self.t_sample)) + 2
        out = [np.zeros((dof, stride)) for _ in range(4)]
        ln = self._fn("get_trajectory", C.c_int)(self.h, _vp(t7), _vp(dir), _vp(mod), _vp(q_0), _vp(v_0), _vp(a_0),
                                                 _vp(v_drive), C.c_int64(stride), *[_vp(x) for x in out])
        if ln < 0:
            return self.get_trajectory(t7, dir, mod, q_0, v_0, a_0, v_drive, stride=-ln)
        return dict(length=ln, q=out[0][:, :ln], v=out[1][:, :ln], a=out[2][:, :ln], j=out[3][:, :ln])

    def plan(self, q_goal, q_0, v_0, a_0, stride=8192):
        dof = self.dof
        q_goal, q_0, v_0, a_0 = (_f64(x, (dof,)) for x in (q_goal, q_0, v_0, a_0))
        out = [np.zeros((dof, stride)) for _ in range(4)]
        ln = C.c_int(-1)
        ok = self._fn("plan", C.c_int)(self.h, _vp(q_goal), _vp(q_0), _vp(v_0), _vp(a_0), C.c_int64(stride),
                                       *[_vp(x) for x in out], C.byref(ln))
        ln = ln.value
        if ln > stride:
            return self.plan(q_goal, q_0, v_0, a_0, stride=ln)
        k = max(ln, 0)
        return dict(success=bool(ok), length=ln, q=out[0][:, :k], v=out[1][:, :k], a=out[2][:, :k], j=out[3][:, :k])

    def plan_batch(self, q_goal, q_0, v_0, a_0, threads=1):
        dof = self.dof
        q_goal, q_0, v_0, a_0 = (_f64(x).reshape(-1, dof) for x in (q_goal, q_0, v_0, a_0))
        n = q_goal.shape[0]
        ok, ln = np.zeros(n, np.uint8), np.zeros(n, np.int32)
        s = self._fn("plan_batch", C.c_double)(self.h, C.c_int64(n), _vp(q_goal), _vp(q_0), _vp(v_0), _vp(a_0),
                                               _vp(ok), _vp(ln), C.c_int(threads))
        return dict(success=ok, length=ln, checksum=s)


class ReferenceO0(Reference):
    """the same sources at the reference's shipped flags (-std=c++17 only, i.e. -O0)"""
    libname = "libltp_ref_O0.so"
    kind = "reference"
    flags = "-O0 (shipped flags, CMakeLists.txt:5)"
