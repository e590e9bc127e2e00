"""Per-source-line instruction counts of one kernel: joins the SASS page of an .ncu-rep
(captured with --import-source on) with nvdisasm's line table of the same cubin.
Runs in the build container, no GPU needed.

  python tools/ncu_lines.py gpurun_out/x.ncu-rep ltp_solve_fast_kernelILi7 [--top 40] [--ranges]

Columns: warp instructions executed, share of the kernel, average active threads, the
FP64 share (D* opcodes), and stall samples. --ranges adds a per-function roll-up using the
line ranges of the functions in ltp_math.cuh / ltp_b200.cu (innermost inlined line)."""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "longtermplanner_b200", "lib", "libltp_b200.so")


def line_table(kernel_substr):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True,
                         text=True).stdout.splitlines()
    table, cur, inside = {}, None, False
    for ln in txt:
        if ln.startswith("//---") and ".text." in ln:
            inside = kernel_substr in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def function_ranges():
    """(file, first line, last line, name) of top-level functions, by a crude brace scan"""
    out = []
    for f in ("ltp_math.cuh", "ltp_b200.cu"):
        path = os.path.join(ROOT, "longtermplanner_b200", "csrc", f)
        lines = open(path).read().splitlines()
        name, start, pending_global = None, None, False
        for i, ln in enumerate(lines, 1):
            if name is None:
                m = re.match(r"^(?:LTP_HD(?:_NOINLINE)?|__device__ __forceinline__|inline|static)\s.*?(\w+)\(", ln)
                if ln.startswith("__global__"):
                    pending_global = True
                    m = None
                elif pending_global:
                    m = re.match(r"^(\w+)\(", ln)
                    pending_global = m is None
                if m:
                    name, start = m.group(1), i
                    if ln.rstrip().endswith("}"):  # one-liner
                        out.append((f, start, i, name))
                        name = None
            elif ln.startswith("}"):
                out.append((f, start, i, name))
                name = None
    return out


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    table = line_table(kern)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    # one section per captured launch: "Kernel Name",<demangled> / header / instructions
    m = re.match(r"(\w+?)ILi(\d+)$", kern)
    want = f"{m.group(1)}<(int){m.group(2)}" if m else kern  # (further template arguments may follow)
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    sect = next(i for i in starts if want in rows[i][1])
    end = next((i for i in starts if i > sect), len(rows))
    rows = rows[sect:end]
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hi]
    ia, ii, it, isamp = h.index("Address"), h.index("Instructions Executed"), h.index(
        "Thread Instructions Executed"), h.index("# Samples")
    body = rows[hi + 1:]
    base = int(body[0][ia], 16)
    per = defaultdict(lambda: [0, 0, 0, 0])  # inst, thread inst, fp64 inst, samples
    total = 0
    for r in body:
        off = int(r[ia], 16) - base
        key, op = table.get(off, (None, r[1]))
        n, t, s = int(r[ii]), int(r[it]), int(r[isamp])
        e = per[key]
        e[0] += n
        e[1] += t
        e[3] += s
        opc = op.split()[0] if not op.startswith("@") else op.split()[1]
        if opc.startswith(("DADD", "DMUL", "DFMA", "DSETP", "MUFU.RCP64H", "MUFU.RSQ64H")):
            e[2] += n
        total += n
    tot_s = sum(e[3] for e in per.values())
    print(f"kernel {kern}: {total} warp instructions, {tot_s} samples")
    print(f"{'line':28s} {'inst':>11s} {'share':>6s} {'thr':>5s} {'fp64':>5s} {'samples':>7s}")
    for key, e in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        nm = f"{key[0]}:{key[1]}" if key else "?"
        print(f"{nm:28s} {e[0]:11d} {100 * e[0] / total:5.1f}% {e[1] / max(e[0], 1):5.1f} "
              f"{100 * e[2] / max(e[0], 1):4.0f}% {100 * e[3] / max(tot_s, 1):6.1f}%")
    if "--ranges" in sys.argv:
        fr = function_ranges()
        agg = defaultdict(lambda: [0, 0, 0, 0])
        for key, e in per.items():
            name = "?"
            if key:
                for f, a, b, nme in fr:
                    if f == key[0] and a <= key[1] <= b:
                        name = nme
                        break
                else:
                    name = key[0] + ":other"
            for k in range(4):
                agg[name][k] += e[k]
        print()
        for name, e in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            print(f"{name:28s} {e[0]:11d} {100 * e[0] / total:5.1f}% {e[1] / max(e[0], 1):5.1f} "
                  f"{100 * e[2] / max(e[0], 1):4.0f}% {100 * e[3] / max(tot_s, 1):6.1f}%")


if __name__ == "__main__":
    main()
