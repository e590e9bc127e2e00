"""Condense an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file x.csv) into one line
per kernel: launches, mean / min / max duration, and the kernel's share of one solve step (the
first kernels of bench.py are its solve steps: closed-form, attempt 2, the three item kernels, the
work-list kernel, in that order, once per step).
  python tools/launch_summary.py gpurun_out/launches.csv [> profiles/xxx_launch_summary.txt]"""
import collections
import csv
import sys


def short(name):
    name = name.split("(")[0]
    for junk in ("void ", "<unnamed>::", "ltp::"):
        name = name.replace(junk, "")
    return name.strip()


def main(path):
    rows = [r for r in csv.reader(open(path, newline="")) if r]
    head = next(i for i, r in enumerate(rows) if r[0] == "ID")
    hdr = rows[head]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    pi = hdr.index("Process ID")
    seq = []
    for r in rows[head + 1:]:
        if len(r) <= vi or not r[0].isdigit():
            continue
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[ui], 1e-3)
        seq.append((short(r[ki]), v * scale, r[pi]))
    print(f"{path}: {len(seq)} launches, processes {sorted(set(p for _, _, p in seq))}")
    agg = collections.OrderedDict()
    for k, us, _ in seq:
        agg.setdefault(k, []).append(us)
    print(f"{'kernel':48s} {'launches':>8s} {'mean us':>10s} {'min us':>10s} {'max us':>10s} {'total ms':>10s}")
    for k, v in agg.items():
        print(f"{k[:48]:48s} {len(v):8d} {sum(v) / len(v):10.1f} {min(v):10.1f} {max(v):10.1f} {sum(v) / 1e3:10.3f}")
    # the first solve step: from the first closed-form launch up to (not including) the second
    first = [i for i, (k, _, _) in enumerate(seq) if k.startswith("ltp_solve_fast_kernel")]
    if len(first) >= 2:
        step = seq[first[0]:first[1]]
        total = sum(us for _, us, _ in step)
        print(f"\nfirst solve step under ncu (cold caches, serialised): {total:.1f} us in {len(step)} launches")
        for k, us, _ in step:
            print(f"  {k[:48]:48s} {us:10.1f} us  {100 * us / total:5.1f} %")


if __name__ == "__main__":
    main(sys.argv[1])
