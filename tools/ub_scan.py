"""Which inputs drive the UNMODIFIED reference into undefined behaviour (impulse
writes past the end of the jerk array cc:771,776,793,807; float -> int conversions of non-finite
values)? The reference sources are built with -fsanitize=address,undefined in recover mode
(oracle/Makefile `san`) and run over the seeded workloads; every report is attributed to its input.
TEST/ANALYSIS TOOLING.   python tools/ub_scan.py [--json profiles/r02_ub_scan.json]"""
import collections
import json
import os
import re
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from longtermplanner_b200 import workloads as W  # noqa: E402


def scan(lim, states, name):
    qg, q0, v0, a0 = states
    n = qg.shape[0]
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
        f.write(struct.pack("<qqd", lim.dof, n, lim.t_sample))
        for x in lim.arrays():
            f.write(np.ascontiguousarray(x, dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(np.stack([qg, q0, v0, a0], axis=1), dtype=np.float64).tobytes())
        path = f.name
    env = dict(os.environ, ASAN_OPTIONS="halt_on_error=0:detect_leaks=0", UBSAN_OPTIONS="print_stacktrace=0")
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ub_scan"), path], capture_output=True, text=True, env=env)
    os.unlink(path)
    cur, per_line, hit = -1, collections.Counter(), collections.defaultdict(set)
    for line in r.stderr.splitlines():
        if line.startswith("@@ "):
            cur = int(line[3:])
            continue
        m = re.search(r"long_term_planner\.cc:(\d+):\d+: runtime error: (.*)", line)
        if m:
            key = f"cc:{m.group(1)} UBSan: {m.group(2)[:80]}"
        else:
            m = re.search(r"ERROR: AddressSanitizer: ([\w-]+)", line)
            if not m:
                m2 = re.search(r"#\d+ .* in .*long_term_planner\.cc:(\d+)", line)
                if m2 and cur in hit.get("_pending_asan", ()):  # first frame inside the reference
                    hit["_pending_asan"].discard(cur)
                    per_line[f"cc:{m2.group(1)} ASan"] += 1
                    hit[f"cc:{m2.group(1)} ASan"].add(cur)
                continue
            hit["_pending_asan"].add(cur)
            key = f"ASan: {m.group(1)}"
        per_line[key] += 1
        hit[key].add(cur)
    hit.pop("_pending_asan", None)
    bad = sorted(set().union(*hit.values())) if hit else []
    out = {"workload": name, "plans": n, "exit_code": r.returncode, "stdout": r.stdout.strip(),
           "plans_with_a_report": len(bad), "reports": dict(per_line),
           "plans_by_report": {k: len(v) for k, v in hit.items()}, "examples": []}
    for i in bad[:3]:
        out["examples"].append({"index": int(i), "q_goal": qg[i].tolist(), "q_0": q0[i].tolist(), "v_0": v0[i].tolist(),
                                "a_0": a0[i].tolist(), "reports": [k for k, v in hit.items() if i in v]})
    return out


if __name__ == "__main__":
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "san"], check=True, stdout=subprocess.DEVNULL)
    res = [scan(W.FRANKA7, W.random_states(W.FRANKA7, 20000, W.SEEDS[1]), "configs[0]/[1]: FRANKA7 random states"),
           scan(W.REF_RANDOM6, W.random_states(W.REF_RANDOM6, 20000, W.SEEDS[2]), "REF_RANDOM6 random states"),
           scan(W.FRANKA7, W.edge_states(W.FRANKA7, 20000, 5), "FRANKA7 controller-like edge states"),
           scan(W.FRANKA12, W.random_states(W.FRANKA12, 5000, W.SEEDS[5]), "configs[4]: FRANKA12 random states")]
    print(json.dumps(res, indent=1))
    if len(sys.argv) > 2 and sys.argv[1] == "--json":
        json.dump(res, open(sys.argv[2], "w"), indent=1)
