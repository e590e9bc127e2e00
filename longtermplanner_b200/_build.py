"""Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIBDIR = os.path.join(_HERE, "lib")
LIB = os.path.join(LIBDIR, "libltp_b200.so")
HOSTLIB = os.path.join(LIBDIR, "liblong_term_planner.so")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # numerical contract: no contraction of a*b+c (csrc/ltp_math.cuh)
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built")


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """nvcc -> longtermplanner_b200/lib/libltp_b200.so. Returns the path."""
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [os.path.join(CSRC, "ltp_b200.cu")]
    deps = srcs + [os.path.join(CSRC, "ltp_math.cuh"), os.path.join(INCLUDE, "ltp_b200.h")]
    if force or _stale(LIB, deps):
        cmd = [_nvcc(), *NVCC_FLAGS, "-I", INCLUDE, *srcs, "-o", LIB]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.run(cmd, check=True)
    return LIB


PROBELIB = os.path.join(LIBDIR, "libltp_probe.so")


def build_probe_library(force: bool = False) -> str:
    """bench-only roofline probes (FP64 FMA rate, HBM streaming-store bandwidth)"""
    os.makedirs(LIBDIR, exist_ok=True)
    src = os.path.join(CSRC, "ltp_probe.cu")
    if force or _stale(PROBELIB, [src]):
        subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                        "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared", src, "-o", PROBELIB], check=True)
    return PROBELIB


WORKLOADLIB = os.path.join(LIBDIR, "libltp_workload.so")


def build_workload_library(force: bool = False) -> str:
    """bench/test tooling: device-side workload generator + trajectory row statistics"""
    os.makedirs(LIBDIR, exist_ok=True)
    src = os.path.join(CSRC, "ltp_workload.cu")
    if force or _stale(WORKLOADLIB, [src]):
        subprocess.run([_nvcc(), *NVCC_FLAGS, src, "-o", WORKLOADLIB], check=True)
    return WORKLOADLIB


def build_host_library(force: bool = False) -> str:
    """g++ -> lib/liblong_term_planner.so: the C++ drop-in class over the C ABI."""
    os.makedirs(LIBDIR, exist_ok=True)
    src = os.path.join(CSRC, "long_term_planner.cc")
    hdr = os.path.join(INCLUDE, "long_term_planner", "long_term_planner.h")
    if not os.path.exists(src):
        return ""
    if force or _stale(HOSTLIB, [src, hdr, LIB]):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I", INCLUDE, src, "-o", HOSTLIB,
                        "-L", LIBDIR, "-lltp_b200", "-Wl,-rpath,$ORIGIN"], check=True)
    return HOSTLIB
