"""Bench / test tooling on the device (csrc/ltp_workload.cu): a generator that is
bit-identical to ``workloads.random_states`` and a per-row reducer over sampled trajectories.
Not part of the planning path."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _build
from .workloads import Limits

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_build.WORKLOADLIB):
            _build.build_workload_library()
        _lib = C.CDLL(_build.WORKLOADLIB)
        _lib.ltp_wl_random_states.restype = C.c_int
        _lib.ltp_wl_random_states.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int64, C.c_int64, C.c_uint64,
                                                                              C.c_double] + [C.c_void_p] * 5
        _lib.ltp_wl_row_stats.restype = C.c_int
        _lib.ltp_wl_row_stats.argtypes = [C.c_int, C.c_int64, C.c_int, C.c_int64, C.c_int] + [C.c_void_p] * 7
    return _lib


def random_states_device(lim: Limits, n: int, seed: int, start: int = 0, margin: float = 0.05, device=None,
                         out=None):
    """-> [q_goal, q_0, v_0, a_0], each a [dof, n] float64 CUDA tensor (joint-major), equal bit
    for bit to ``to_joint_major(workloads.random_states(...))``"""
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    arrs = [np.ascontiguousarray(x) for x in lim.arrays()]
    if out is None:
        out = [torch.empty(lim.dof, n, dtype=torch.float64, device=dev) for _ in range(4)]
    st = torch.cuda.current_stream(dev).cuda_stream
    rc = lib().ltp_wl_random_states(lim.dof, *[a.ctypes.data for a in arrs], n, start, seed, margin,
                                    *[t.data_ptr() for t in out], st)
    if rc != 0:
        raise RuntimeError(f"ltp_wl_random_states failed ({rc})")
    return out


ROW_STATS = ("sum_q", "sum_v", "sum_a", "sum_j", "max_abs_v", "max_abs_a", "q_last", "v_last")


def row_stats(traj, traj_len, horizon: int = 0) -> torch.Tensor:
    """-> [n, dof, 8] float64 CUDA tensor (ROW_STATS) of a BatchTrajectories"""
    layout = 1 if traj.layout == "time_major" else 0
    if layout == 1:
        _, n, dof = traj.q.shape
    else:
        n, dof, _ = traj.q.shape
    out = torch.empty(n, dof, 8, dtype=torch.float64, device=traj.q.device)
    st = torch.cuda.current_stream(traj.q.device).cuda_stream
    rc = lib().ltp_wl_row_stats(layout, n, dof, traj.stride, horizon, traj_len.data_ptr(), traj.q.data_ptr(),
                                traj.v.data_ptr(), traj.a.data_ptr(), traj.j.data_ptr(), out.data_ptr(), st)
    if rc != 0:
        raise RuntimeError(f"ltp_wl_row_stats failed ({rc})")
    return out
