"""ctypes view of include/ltp_b200.h. There is no fallback: if the CUDA library has not
been built this module raises, and every entry point needs a CUDA device."""
from __future__ import annotations

import ctypes as C
import os

from ._build import LIB

# a differently built copy of the same library (kernel experiments); never a fallback
LIB = os.environ.get("LTP_B200_LIB", LIB)

if not os.path.exists(LIB):
    raise ImportError(
        f"{LIB} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(needs nvcc). longtermplanner_b200 has no CPU fallback.")

lib = C.CDLL(LIB)

LTP_OK, LTP_ERR_ARG, LTP_ERR_CUDA, LTP_ERR_CAPACITY = 0, -1, -2, -3
LTP_MAX_DOF = 32

vp, i64, i32, f64 = C.c_void_p, C.c_int64, C.c_int32, C.c_double


class Solution(C.Structure):
    """ltp_solution"""
    _fields_ = [(k, vp) for k in ("t_scaled", "dir", "v_drive", "mod", "slowest", "traj_len", "reached",
                                  "t_opt", "opt_case", "ts_case", "final_case")]


def _sig(name, restype, *argtypes):
    f = getattr(lib, name)
    f.restype = restype
    f.argtypes = list(argtypes)
    return f


EXPORTS = [
    "ltp_create", "ltp_set_limits", "ltp_set_sample_time", "ltp_set_dof", "ltp_set_solve_mode", "ltp_set_stream_sorted", "ltp_set_profiling", "ltp_profile_read", "ltp_get_dof", "ltp_get_device",
    "ltp_destroy", "ltp_status_string", "ltp_last_cuda_error", "ltp_launch_count",
    "ltp_opt_braking_batch", "ltp_opt_switch_times_batch", "ltp_time_scaling_batch", "ltp_reserve", "ltp_solve_batch",
    "ltp_sample_batch", "ltp_sample_batch_sorted", "ltp_plan_stream", "ltp_advance_batch", "ltp_transpose", "ltp_solve_host", "ltp_plan_host", "ltp_plan_one_view", "ltp_opt_braking_host",
    "ltp_opt_switch_times_host", "ltp_time_scaling_host", "ltp_get_trajectory_host",
]

class Chunk(C.Structure):
    """ltp_chunk"""
    _fields_ = [("first", i64), ("count", i64), ("capacity", i64), ("horizon", i32), ("solution", Solution),
                ("q_goal", vp), ("q_0", vp), ("v_0", vp), ("a_0", vp), ("q", vp), ("v", vp), ("a", vp), ("j", vp),
                ("success", vp), ("order", vp)]


class StreamStats(C.Structure):
    """ltp_stream_stats"""
    _fields_ = [(k, i64) for k in ("problems", "chunks", "reached", "success", "clipped", "samples", "bytes",
                                   "max_traj_len")]


CHUNK_CONSUMER = C.CFUNCTYPE(C.c_int, vp, C.POINTER(Chunk), vp)

create = _sig("ltp_create", C.c_int, C.POINTER(vp), C.c_int, C.c_int, f64, vp, vp, vp, vp, vp)
set_limits = _sig("ltp_set_limits", C.c_int, vp, vp, vp, vp, vp, vp)
set_sample_time = _sig("ltp_set_sample_time", C.c_int, vp, f64)
set_dof = _sig("ltp_set_dof", C.c_int, vp, C.c_int)
set_solve_mode = _sig("ltp_set_solve_mode", C.c_int, vp, C.c_int)
set_profiling = _sig("ltp_set_profiling", C.c_int, vp, C.c_int)
set_stream_sorted = _sig("ltp_set_stream_sorted", C.c_int, vp, C.c_int)
profile_read = _sig("ltp_profile_read", C.c_int, vp, C.c_int, C.POINTER(C.c_double), C.POINTER(i64), C.c_int)
get_dof = _sig("ltp_get_dof", C.c_int, vp)
get_device = _sig("ltp_get_device", C.c_int, vp)
destroy = _sig("ltp_destroy", None, vp)
status_string = _sig("ltp_status_string", C.c_char_p, C.c_int)
last_cuda_error = _sig("ltp_last_cuda_error", C.c_char_p)
launch_count = _sig("ltp_launch_count", i64, vp)
opt_braking_batch = _sig("ltp_opt_braking_batch", C.c_int, vp, i64, vp, vp, vp, vp, vp, vp)
opt_switch_times_batch = _sig("ltp_opt_switch_times_batch", C.c_int, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp,
                              vp, vp, vp)
time_scaling_batch = _sig("ltp_time_scaling_batch", C.c_int, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                          vp, vp, vp)
solve_batch = _sig("ltp_solve_batch", C.c_int, vp, i64, vp, vp, vp, vp, C.POINTER(Solution), vp)
sample_batch = _sig("ltp_sample_batch", C.c_int, vp, i64, vp, vp, vp, C.POINTER(Solution), i32, i32, i64, vp,
                    vp, vp, vp, vp, vp)
LAYOUT_ROWS, LAYOUT_TIME_MAJOR = 0, 1
sample_batch_sorted = _sig("ltp_sample_batch_sorted", C.c_int, vp, i64, vp, vp, vp, C.POINTER(Solution), i64, vp, vp,
                           vp, vp, vp, vp, vp)
plan_stream = _sig("ltp_plan_stream", C.c_int, vp, i64, vp, vp, vp, vp, i64, i32, i64, CHUNK_CONSUMER, vp,
                   C.POINTER(StreamStats), vp)
advance_batch = _sig("ltp_advance_batch", C.c_int, vp, i64, i32, i32, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp)
reserve = _sig("ltp_reserve", C.c_int, vp, i64)
transpose = _sig("ltp_transpose", C.c_int, vp, i64, i64, vp, vp, vp)
solve_host = _sig("ltp_solve_host", C.c_int, vp, i64, vp, vp, vp, vp, C.POINTER(Solution))
plan_host = _sig("ltp_plan_host", C.c_int, vp, i64, vp, vp, vp, vp, i32, i64, vp, vp, vp, vp, vp, vp,
                 C.POINTER(i64))
plan_one_view = _sig("ltp_plan_one_view", C.c_int, vp, vp, vp, vp, vp, C.POINTER(vp * 4), C.POINTER(i64),
                     C.POINTER(i32), C.POINTER(C.c_uint8))
opt_braking_host = _sig("ltp_opt_braking_host", C.c_int, vp, C.c_int, f64, f64, vp, vp, vp)
opt_switch_times_host = _sig("ltp_opt_switch_times_host", C.c_int, vp, C.c_int, f64, f64, f64, f64, f64, vp,
                             vp, vp, vp, vp)
time_scaling_host = _sig("ltp_time_scaling_host", C.c_int, vp, C.c_int, f64, f64, f64, f64, f64, f64, vp, vp,
                         vp, vp, vp)
get_trajectory_host = _sig("ltp_get_trajectory_host", C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp, vp,
                           vp, vp, vp, C.POINTER(i64))


class LtpError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc == LTP_OK:
        return
    msg = status_string(rc).decode()
    if rc == LTP_ERR_CUDA:
        msg += ": " + last_cuda_error().decode()
    raise LtpError(f"{what or 'ltp call'} failed: {msg}")
