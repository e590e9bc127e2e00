"""Condense an .ncu-rep (read here, no GPU needed) into the handful of numbers the roofline
discussion uses. Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--json profiles/ncu_facts.json] [> profiles/xxx.txt]"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "occ limit regs (blocks)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe util % (inst)"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe cycles active %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instr"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "L1 global store sectors"),
    ("smsp__sass_inst_executed_op_local_ld.sum", "local loads"),
    ("smsp__sass_inst_executed_op_local_st.sum", "local stores"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_drain_per_issue_active.ratio", "stall drain"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall dispatch"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
]


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value.replace(",", "")) * scale.get(unit, 1)


def main(path, json_out=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    seen = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].replace("void ", "").replace("<unnamed>::", "").split("(")[0]
        seen.setdefault(name, []).append(r)
    facts = {}
    for name, rs in seen.items():
        di = hdr.index("gpu__time_duration.sum")
        r = max(rs, key=lambda x: float(x[di]))  # the largest launch of each kernel
        print(f"=== {name}  ({len(rs)} launches captured; showing the longest)")
        for k, label in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {label:34s} {r[i]:>16s} {units[i]}")
        print()
        if json_out:
            g = lambda k: (r[hdr.index(k)], units[hdr.index(k)])
            facts[name.split("<")[0]] = {
                "dram_bytes_per_launch": to_bytes(*g("dram__bytes_read.sum")) + to_bytes(*g("dram__bytes_write.sum")),
                "dram_read_bytes": to_bytes(*g("dram__bytes_read.sum")),
                "dram_write_bytes": to_bytes(*g("dram__bytes_write.sum")),
                "duration": " ".join(g("gpu__time_duration.sum")),
                "fp64_pipe_pct": float(g("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")[0]),
                "grid": r[hdr.index("launch__grid_size")], "block": r[hdr.index("launch__block_size")],
                "registers": r[hdr.index("launch__registers_per_thread")],
                "source": "ncu --set full --clock-control none, " + path.split("/")[-1] + " (longest launch of the kernel)"}
    if json_out:
        # stamp: hash of the kernel sources the capture belongs to (bench.py refuses to quote a
        # capture of other code) -- so run this right after the capture, before editing csrc/
        import datetime
        import hashlib
        import json
        import os
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        h = hashlib.sha256()
        for f in ("ltp_b200.cu", "ltp_math.cuh"):
            h.update(open(os.path.join(root, "longtermplanner_b200", "csrc", f), "rb").read())
        facts["source_sha"] = h.hexdigest()[:16]
        facts["captured"] = datetime.date.today().isoformat() + ", " + path.split("/")[-1]
        json.dump(facts, open(json_out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[3] if len(sys.argv) > 3 and sys.argv[2] == "--json" else None)
