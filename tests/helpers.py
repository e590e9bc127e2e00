"""Shared comparison helpers for the parity tests.

Bar (BASELINE.json north_star): success flag, case ids and trajectory length exact;
phase times and sampled q/v/a/j within 1e-9 relative / 1e-12 absolute in FP64."""
import numpy as np

RTOL, ATOL = 1e-9, 1e-12


def close(got, ref):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        ok = np.abs(got - ref) <= ATOL + RTOL * np.abs(ref)
    return ok | (got == ref) | (np.isnan(got) & np.isnan(ref))


def count_bad(got, ref):
    return int((~close(got, ref)).sum())


def bitdiff(got, ref):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return int((~((got == ref) | (np.isnan(got) & np.isnan(ref)))).sum())


def jm(x):
    """problem-major [n, dof, ...] -> joint-major [..., dof, n] contiguous"""
    x = np.asarray(x)
    return np.ascontiguousarray(np.moveaxis(x, 0, -1) if x.ndim == 2 else x.transpose(2, 1, 0))


def pm(x):
    """joint-major [dof, n] or [7, dof, n] -> problem-major [n, dof] / [n, dof, 7]"""
    x = np.asarray(x)
    return np.ascontiguousarray(x.T if x.ndim == 2 else x.transpose(2, 1, 0))
