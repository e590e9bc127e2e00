"""Dynamic opcode mix of one kernel from an .ncu-rep (source page): warp instructions, active threads
and stall samples per SASS mnemonic.  python tools/ncu_opcodes.py rep kernel_regex [top]"""
import collections
import csv
import re
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kern}", "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# one section per captured launch ("Kernel Name" row first): keep the longest launch only
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
if len(starts) > 2:
    a, b = max(zip(starts, starts[1:]), key=lambda ab: ab[1] - ab[0])
    sections = [(a, b)]
    def weight(ab):
        tot = 0
        for r in rows[ab[0]:ab[1]]:
            for c in r[2:6]:
                if c.isdigit():
                    tot += int(c)
                    break
        return tot
    a, b = max(zip(starts, starts[1:]), key=weight)
    rows = rows[a:b]
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
iS, iI, iT, iSm = (hdr.index(k) for k in ("Source", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
ops, thr, smp = collections.Counter(), collections.Counter(), collections.Counter()
for r in rows:
    if len(r) <= iSm or not r[iI].isdigit():
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS])
    if not m:
        continue
    op = m.group(2).split(".")[0]
    ops[op] += int(r[iI]); thr[op] += int(r[iT]); smp[op] += int(r[iSm])
tot, S = sum(ops.values()), max(sum(smp.values()), 1)
print(f"{kern}: {tot} warp instructions")
for op, n in ops.most_common(top):
    print(f"{op:10s} {n:12d} {100 * n / tot:5.1f}%  thr {thr[op] / max(n, 1):5.1f}  samples {100 * smp[op] / S:5.1f}%")
