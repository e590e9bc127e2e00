"""Build named variants of libltp_b200.so into tools/_bin/ (git-ignored, shipped to the GPU
box) for A/B timing with LTP_B200_LIB=... python tools/solve_timing.py.
  python tools/build_variants.py name=-DFLAG=1,-DOTHER=2 name2=..."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longtermplanner_b200 import _build  # noqa: E402

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_bin")
os.makedirs(out, exist_ok=True)
procs = []
for spec in sys.argv[1:]:
    name, _, flags = spec.partition("=")
    cmd = [_build._nvcc(), *_build.NVCC_FLAGS, *[f for f in flags.split(",") if f], "-I", _build.INCLUDE,
           os.path.join(_build.CSRC, "ltp_b200.cu"), "-o", os.path.join(out, f"libltp_{name}.so")]
    procs.append((name, subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)))
for name, pr in procs:
    print(name, "ok" if pr.wait() == 0 else "FAILED")
