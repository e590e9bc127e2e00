#!/usr/bin/env python
"""bench.py -- the hot-path benchmark (contract in the task statement, metric from BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU planTrajectory path

Workload at every N (weak scaling, one process per GPU, no collective on the data path):
BASELINE.json configs[1] -- 2^20 random 7-DoF problems per GPU (FRANKA7 limits, recipe of the
reference's tests/randomConfiguration.m), phase times + synchronisation + time scaling.
One "step" = one pass of the solve over the rank's 2^20 problems.

  value   plans/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e     plans/s through the host-buffer C-ABI call (ltp_solve_host): pinned host inputs
          copied in and results copied out inside the timed region
  roofline        solver kernel vs the FP64 pipe (algorithmic 2470 flop per 7-DoF plan)
  sampler         BASELINE.json configs[2] (4096 envs x 7 DoF x 2001 samples): the dense
                  q/v/a/j sampler vs HBM bandwidth (32 B per sample), plus replan latency
  cpu_baseline    the reference's CPU code timed on this box's host cores (rank 0, N = 1)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from longtermplanner_b200 import workloads as W  # noqa: E402

N_PER_GPU = 1 << 20
FLOP_PER_PLAN_7DOF = 388 * 7 - 246  # W_solve(dof) = 388*dof - 246 (fast-path count)
METRIC = "7-DoF plans/sec (phase times + time scaling, 2^20 random problems per GPU)"


def env_int(k, d):
    return int(os.environ.get(k, d))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_checker(lim):
    """the reference's CPU code: oracle/_ref (reference .cc + Eigen shim) if built, else the port"""
    from oracle.bindings import OraclePort, Reference, build
    try:
        build()
    except Exception:
        pass
    cls = Reference if Reference.available() else OraclePort
    return cls.from_limits(lim), cls.kind


def cpu_solve_rate(n_sample, threads, seed):
    lim = W.FRANKA7
    chk, kind = cpu_checker(lim)
    qg, q0, v0, a0 = W.random_states(lim, n_sample, seed)
    chk.solve(qg[:2048], q0[:2048], v0[:2048], a0[:2048], threads=threads)  # warm
    t0 = time.perf_counter()
    chk.solve(qg, q0, v0, a0, threads=threads)
    dt = time.perf_counter() - t0
    return n_sample / dt, kind


RTOL, ATOL = 1e-9, 1e-12  # north_star: 1e-9 relative / 1e-12 absolute on FP64 values


def _close(got, ref):
    with np.errstate(invalid="ignore"):
        ok = np.abs(got - ref) <= ATOL + RTOL * np.abs(ref)
    return ok | (got == ref) | (np.isnan(got) & np.isnan(ref))


def parity_counts(sol, ref, t_sample, cases=None):
    """Counts of disagreement between a BatchSolution (joint-major CUDA tensors, with case bytes and
    t_opt) and the CPU checker's solve of the same problems (problem-major numpy). exact_mismatch:
    reached, slowest, traj_len, dir, mod (and, against `cases` = the C restatement's result, the
    three case bytes -- the reference itself emits no case ids); numeric_mismatch: switching times
    and v_drive outside 1e-9 rel / 1e-12 abs; bitdiff: values that are not bit-identical.
    The unmodified reference returns no trajectory length from the solve: it is derived from its
    switching times with the reference's own formula (cc:716-719), for reached problems."""
    ref = dict(ref)
    if "traj_len" not in ref:
        t6 = ref["t_scaled"][:, :, 6]
        with np.errstate(invalid="ignore"):
            ln = (np.ceil(t6 / t_sample) + 1).max(axis=1)
        ref["traj_len"] = np.where(ref["reached"] != 0, ln, 0).astype(np.int32)
    exact = {}
    for k in ("reached", "slowest", "traj_len"):
        exact[k] = int((getattr(sol, k).cpu().numpy() != ref[k]).sum())
    exact["dir"] = int((sol.dir.cpu().numpy().T != ref["dir"]).sum())
    exact["mod"] = int((sol.mod.cpu().numpy().T != ref["mod"]).sum())
    if cases is not None:
        for k in ("opt_case", "ts_case", "final_case"):
            exact[k] = int((getattr(sol, k).cpu().numpy().T != cases[k]).sum())
    numeric, bits, values = {}, 0, 0
    for k in ("t_scaled", "t_opt", "v_drive"):
        got = getattr(sol, k).cpu().numpy()
        got = got.transpose(2, 1, 0) if got.ndim == 3 else got.T
        numeric[k] = int((~_close(got, ref[k])).sum())
        bits += int((~((got == ref[k]) | (np.isnan(got) & np.isnan(ref[k])))).sum())
        values += got.size
    return {"checked": int(ref["reached"].shape[0]), "exact_mismatch": int(sum(exact.values())),
            "numeric_mismatch": int(sum(numeric.values())), "bitdiff": bits, "values_compared": values,
            "exact_by_field": exact, "numeric_by_field": numeric, "tolerance": {"rel": RTOL, "abs": ATOL}}


def full_parity(ltp, lim, dev_in, states, threads):
    """all problems of a workload: the CUDA solve against the reference build (values, flags) and
    the C restatement (case ids)"""
    import torch
    from oracle.bindings import OraclePort
    chk, kind = cpu_checker(lim)
    t0 = time.perf_counter()
    ref = chk.solve(*states, threads=threads)
    cpu_s = time.perf_counter() - t0
    cases = ref if kind == "port" else OraclePort.from_limits(lim).solve(*states, threads=threads)
    full = ltp.solve(*dev_in, with_opt=True, with_cases=True)
    torch.cuda.synchronize()
    par = parity_counts(full, ref, lim.t_sample, cases)
    par.update({"against": kind + (" (values, flags) + C restatement (case ids)" if kind != "port" else ""),
                "cpu_seconds": cpu_s, "cpu_threads": threads})
    return par, full, n_per_s(states, cpu_s)


def n_per_s(states, seconds):
    return states[0].shape[0] / seconds


def run_reference_arm(args):
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    lim = W.FRANKA7
    chk, kind = cpu_checker(lim)
    n_sample = 1 << 18
    qg, q0, v0, a0 = W.random_states(lim, n_sample, W.SEEDS[2])
    for _ in range(max(args.warmup, 1)):
        chk.solve(qg[:1 << 14], q0[:1 << 14], v0[:1 << 14], a0[:1 << 14], threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        chk.solve(qg, q0, v0, a0, threads=cores)
    dt = time.perf_counter() - t0
    value = n_sample * args.steps / dt
    sample = (f"{n_sample} of the 2^20 problems per step, solve only (reference cc:14-55 through the exposed "
              f"protected methods), {cores} host threads; "
              + ("reference long_term_planner.cc unmodified + Eigen shim, g++ -O2 -ffp-contract=off"
                 if kind == "reference" else "plain-C restatement oracle/ltp_oracle.c"))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "plans/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "configs[1]: 2^20 random 7-DoF problems (FRANKA7), solve only; CPU arm "
                                   "times a bounded sample per step", "dof": 7, "t_sample": 0.001},
            "cpu_baseline": {"value": value, "unit": "plans/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "plans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(index):
    """Run this rank's host threads on the CPUs of the GPU's NUMA node, so that the pinned
    buffers of the end-to-end path are first touched (= placed) next to the GPU's PCIe root.
    Returns the node, or None where the topology is not exposed."""
    try:
        import torch
        try:
            pr = torch.cuda.get_device_properties(index)
            bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        except AttributeError:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = vis.split(",")[index] if vis else str(index)
            out = subprocess.run(["nvidia-smi", f"--id={phys}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                 capture_output=True, text=True, timeout=20).stdout.strip()
            bdf = out.lower()[-12:]  # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def kernel_source_hash():
    """sha256 over the CUDA sources the kernels are built from (stamped into profiles/ncu_facts.json
    by tools/ncu_summary.py --json, compared here: numbers read off a capture of OTHER code are not
    printed)"""
    import hashlib
    h = hashlib.sha256()
    for f in ("ltp_b200.cu", "ltp_math.cuh"):
        h.update(open(os.path.join(ROOT, "longtermplanner_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def ncu_facts():
    """per-launch facts read off the committed ncu --set full capture (profiles/ncu_facts.json,
    written by tools/ncu_summary.py --json); {} if the capture has not been made or was taken from
    other kernel sources than the ones this run was built from"""
    try:
        facts = json.load(open(os.path.join(ROOT, "profiles", "ncu_facts.json")))
    except Exception:
        return {}, "no capture"
    if facts.get("source_sha") != kernel_source_hash():
        return {}, f"stale (capture of source {facts.get('source_sha')}, built from {kernel_source_hash()})"
    return facts, f"profiles/ncu_facts.json ({facts.get('captured', '?')})"


def pcie_probe(host_in, host_out_t, dev, iters, barrier, max_over_ranks):
    """The ceiling of the end-to-end path: the step's input bytes go host->device on one stream while
    its output bytes go device->host on another, as flat copies between pinned buffers and nothing
    else. Returns seconds per step (max over ranks)."""
    import torch
    d_in = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in host_in]
    d_out = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in host_out_t]
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def step():
        with torch.cuda.stream(s_in):
            for d, h in zip(d_in, host_in):
                d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s_out):
            for d, h in zip(d_out, host_out_t):
                h.copy_(d, non_blocking=True)

    step()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    torch.cuda.synchronize()
    return max_over_ranks(time.perf_counter() - t0) / iters


ROW_STATS = ("sum_q", "sum_v", "sum_a", "sum_j", "max_abs_v", "max_abs_a", "q_last", "v_last")


class StreamParity:
    """Consumer of the streamed configs[4] run: of every `stride`-th problem of the GLOBAL index
    space it reduces the sampled rows on the device (sums, maxima and the last sample of q, v, a,
    j per joint) and keeps traj_len and the success flag; after the run the same problems are
    planned by the CPU checker and compared. No synchronisation inside the callback."""

    def __init__(self, lim, seed, global_first, stride, dev):
        self.lim, self.seed, self.g0, self.stride, self.dev = lim, seed, global_first, stride, dev
        self.kept = []

    def __call__(self, view, stream):
        import torch
        first, cnt, cap = view["first"], view["count"], view["capacity"]
        g = self.g0 + first
        k0 = (-g) % self.stride
        if k0 >= cnt:
            return
        ks = list(range(k0, cnt, self.stride))
        idx = torch.tensor(ks, device=self.dev)
        tl = view["traj_len"][idx].long()
        live = (torch.arange(cap, device=self.dev)[:, None] < tl[None, :])[:, :, None]
        f = {k: view[k][:, idx, :] for k in "qvaj"}          # (cap, m, dof) copies of the selected slots
        z = {k: torch.where(live, f[k], torch.zeros((), dtype=torch.float64, device=self.dev)) for k in "qvaj"}
        last = (tl - 1).clamp(min=0)
        cols = torch.arange(len(ks), device=self.dev)
        stats = torch.stack([z["q"].sum(0), z["v"].sum(0), z["a"].sum(0), z["j"].sum(0), z["v"].abs().amax(0),
                             z["a"].abs().amax(0), f["q"][last, cols, :], f["v"][last, cols, :]])
        self.kept.append(([g + k for k in ks], stats, tl, view["success"][idx].clone()))

    def compare(self, checker):
        """-> counts against the CPU checker's plan of the same problems"""
        n_prob = exact_bad = numeric_bad = values = 0
        for gids, stats, tl, succ in self.kept:
            stats, tl, succ = stats.cpu().numpy(), tl.cpu().numpy(), succ.cpu().numpy()
            for c, gid in enumerate(gids):
                qg, q0, v0, a0 = W.random_states(self.lim, 1, self.seed, start=gid)
                ref = checker.plan(qg[0], q0[0], v0[0], a0[0], stride=16384)
                n_prob += 1
                ln = ref["length"]
                exact_bad += int(ln != tl[c]) + int(bool(ref["success"]) != bool(succ[c]))
                if ln != tl[c] or ln <= 0:
                    continue
                r = {k: ref[k][:, :ln] for k in "qvaj"}
                want = np.stack([r["q"].sum(1), r["v"].sum(1), r["a"].sum(1), r["j"].sum(1), np.abs(r["v"]).max(1),
                                 np.abs(r["a"]).max(1), r["q"][:, -1], r["v"][:, -1]])
                scale = np.stack([np.abs(r["q"]).sum(1), np.abs(r["v"]).sum(1), np.abs(r["a"]).sum(1),
                                  np.abs(r["j"]).sum(1), np.abs(r["v"]).max(1), np.abs(r["a"]).max(1),
                                  np.abs(r["q"][:, -1]), np.abs(r["v"][:, -1])])
                bad = np.abs(stats[:, c, :] - want) > ATOL + RTOL * scale
                numeric_bad += int(bad.sum())
                values += bad.size
        return n_prob, exact_bad, numeric_bad, values


def generic_section(LongTermPlanner, local, dev, steps, cores):
    """The every-branch (root-solving) kernel under load: 2^20 random 6-DoF problems under the
    reference's own toy limits (REF_RANDOM6: a fifth of the problems needs a polynomial root), the
    AUTO pipeline and the every-branch kernel alone, all problems checked against the CPU checker,
    whose own time on the same problems is the CPU baseline of this workload."""
    import torch
    lim = W.REF_RANDOM6
    n = 1 << 20
    ltp6 = LongTermPlanner(lim.dof, lim.t_sample, *lim.arrays(), device=local)
    qg, q0, v0, a0 = W.random_states(lim, n, W.SEEDS[2])
    ins = [torch.from_numpy(W.to_joint_major(x)).to(dev) for x in (qg, q0, v0, a0)]
    sol = ltp6.alloc_solution(n, with_opt=True, with_cases=True)
    out = {"workload": "2^20 random 6-DoF problems, reference test limits (REF_RANDOM6: v 1, a 2, j 4, t_sample "
                       "1 ms), solve only"}
    for mode, generic_only in (("auto", False), ("every_branch_kernel_only", True)):
        ltp6.setSolveMode(generic_only)
        for _ in range(2):
            ltp6.solve(*ins, out=sol)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            ltp6.solve(*ins, out=sol)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        ltp6.setProfiling(True)
        for k in ("solve_fast", "solve_attempt2", "solve_items", "solve_generic"):
            ltp6.kernelTime(k)
        for _ in range(steps):
            ltp6.solve(*ins, out=sol)
        kern = {}
        for k in ("solve_fast", "solve_attempt2", "solve_items", "solve_generic"):
            k_ms, k_cnt = ltp6.kernelTime(k)
            kern[k] = k_ms / steps
        ltp6.setProfiling(False)
        out[mode] = {"ms_per_step": ms, "plans_per_s": n / (ms * 1e-3), "kernels_ms": kern}
    ltp6.setSolveMode(False)
    ltp6.solve(*ins, out=sol)
    torch.cuda.synchronize()
    tc, oc = sol.ts_case.cpu().numpy(), sol.opt_case.cpu().numpy()
    rooty = ((tc >= 3) & (tc <= 9)).any(axis=0) | np.isin(oc & 15, (6, 7, 8)).any(axis=0)
    out["problems_with_a_root_solve_or_failed_search"] = float(rooty.mean())
    try:
        par, _, cpu_rate = full_parity(ltp6, lim, ins, (qg, q0, v0, a0), cores)
        out["cpu_baseline"] = {"value": cpu_rate, "unit": "plans/s", "cores": cores,
                               "kind": cpu_checker(lim)[1], "sample": "all 2^20 problems, solve only"}
        out["parity"] = par
    except Exception as e:
        out["parity"] = {"checked": 0, "unavailable": str(e)}
    return out



def replan_loop(ltp, lim, dev, hbm_peak, n_env=4096, horizon=2001, tick=9, ticks=50):
    """4096 environments, every control period (10 samples = 10 ms): new goals arrive from the
    host, every environment is re-solved from the state its previous plan has reached by then and
    re-sampled to the 2 s horizon; the next 10 samples go back to the host. Solve + sample +
    state gather are one CUDA graph."""
    import torch
    g0, q0, v0, a0 = W.random_states(lim, n_env, W.SEEDS[3])
    state = [torch.from_numpy(W.to_joint_major(x)).to(dev) for x in (q0, v0, a0)]
    goal = torch.from_numpy(W.to_joint_major(g0)).to(dev)
    pool = [torch.from_numpy(W.to_joint_major(W.random_states(lim, n_env, 1000 + k)[0])).pin_memory()
            for k in range(8)]
    sol = ltp.alloc_solution(n_env)
    traj = ltp.alloc_trajectories(n_env, horizon, "time_major")
    head = torch.empty(3, tick + 1, n_env, lim.dof, dtype=torch.float64).pin_memory()

    def step():
        ltp.solve(goal, *state, out=sol)
        ltp.sample(*state, sol, horizon=horizon, out=traj)
        ltp.advance(traj, tick, *state)

    s = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(s):
        step()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            step()
        for _ in range(3):
            graph.replay()
        s.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record()
        for k in range(ticks):
            graph.replay()
        e[1].record()
        s.synchronize()
        graph_ms = e[0].elapsed_time(e[1]) / ticks
        # end to end: goals from pinned host memory in, the next control period's samples out
        t0 = time.perf_counter()
        for k in range(ticks):
            goal.copy_(pool[k % len(pool)], non_blocking=True)
            graph.replay()
            head[0].copy_(traj.q[:tick + 1], non_blocking=True)
            head[1].copy_(traj.v[:tick + 1], non_blocking=True)
            head[2].copy_(traj.a[:tick + 1], non_blocking=True)
            s.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / ticks
    reached = float(sol.reached.double().mean().item())
    useful = n_env * lim.dof * horizon * 32
    return {"workload": f"configs[2] loop: {n_env} envs x {lim.dof} DoF, replanned every {tick + 1} samples towards new "
                        f"goals, dense sampling to {horizon} samples, {ticks} consecutive replans",
            "graph_replan_ms": graph_ms, "budget_ms": (tick + 1) * lim.t_sample * 1e3,
            "e2e_replan_ms": e2e_ms, "e2e_h2d_bytes": int(pool[0].numel() * 8), "e2e_d2h_bytes": int(head.numel() * 8),
            "write_gbs": useful / (graph_ms * 1e-3) / 1e9, "frac_of_hbm_peak": useful / (graph_ms * 1e-3) / 1e9 / hbm_peak,
            "reached_frac": reached, "kernels_per_replan": 4,
            "timing": "CUDA events around 50 graph replays (solve fast + work-list kernel + sampler + state gather)"}


def single_plan_latency(ltp, lim, calls=300, cpu_plans=2000):
    """configs[0]: one 7-DoF planTrajectory at a time through the host-buffer C ABI (ltp_plan_host,
    n = 1: inputs in, solve, dense sampling, the q/v/a/j rows back out), wall clock per call, next
    to the reference's own planTrajectory on one host core."""
    from longtermplanner_b200 import _capi as capi
    qg, q0, v0, a0 = W.random_states(lim, max(calls, cpu_plans), W.SEEDS[1])
    cap = 4096
    rows = [np.empty((lim.dof, cap)) for _ in range(4)]
    ln, ok, needed = np.zeros(1, np.int32), np.zeros(1, np.uint8), capi.i64(0)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    us = []
    for k in range(calls + 20):
        ins = [np.ascontiguousarray(x[k]) for x in (qg, q0, v0, a0)]
        t0 = time.perf_counter()
        rc = capi.plan_host(ltp._h, 1, *[vp(x) for x in ins], 0, cap, *[vp(r) for r in rows], vp(ln), vp(ok),
                            C.byref(needed))
        dt = time.perf_counter() - t0
        assert rc == 0, rc
        if k >= 20:
            us.append(dt * 1e6)
    out = {"workload": "configs[0]: single 7-DoF planTrajectory (FRANKA7, t_sample 1 ms, random states), "
                       "solve + dense sampling, pageable host buffers in and out, one call at a time",
           "gpu_us_per_plan_median": float(np.median(us)), "gpu_us_per_plan_p90": float(np.percentile(us, 90)),
           "calls": calls, "mean_samples_per_plan": None}
    if CPP_CLASS_LATENCY is not None:
        out["cpp_class"] = CPP_CLASS_LATENCY
    try:
        chk, kind = cpu_checker(lim)
        chk.plan_batch(qg[:64], q0[:64], v0[:64], a0[:64], threads=1)
        t0 = time.perf_counter()
        r = chk.plan_batch(qg[:cpu_plans], q0[:cpu_plans], v0[:cpu_plans], a0[:cpu_plans], threads=1)
        out["cpu_us_per_plan"] = (time.perf_counter() - t0) * 1e6 / cpu_plans
        out["cpu_kind"] = kind
        out["mean_samples_per_plan"] = float(np.mean(r["length"]))
    except Exception as e:
        out["cpu_us_per_plan"] = None
        out["cpu_kind"] = f"unavailable: {e}"
    return out


CPP_CLASS_LATENCY = None


def cpp_class_latency(calls=300):
    """One planTrajectory at a time through the drop-in C++ class (Trajectory's vectors filled),
    tests/cpp/single_plan_bench.cc. Run BEFORE this process creates its own CUDA context: two
    contexts on one GPU time-slice, which adds tens of microseconds to every launch."""
    exe = os.path.join(ROOT, "tests", "_build", "single_plan_bench")
    if not os.path.exists(exe):
        return None
    try:
        r = subprocess.run([exe, str(calls)], capture_output=True, text=True, timeout=120)
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:
        return {"unavailable": str(e)}


def load_probe():
    from longtermplanner_b200 import _build
    path = _build.PROBELIB
    if not os.path.exists(path):
        _build.build_probe_library()
    return C.CDLL(path)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=N_PER_GPU, help="problems per GPU")
    ap.add_argument("--no-sampler", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-stream", action="store_true")
    ap.add_argument("--no-replan", action="store_true")
    ap.add_argument("--no-single", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-generic", action="store_true")
    ap.add_argument("--stream-log2n", type=int, default=26, help="configs[4]: total problems = 2^k over all GPUs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference_arm(args)

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if world == 1 and not args.no_single:
        global CPP_CLASS_LATENCY
        CPP_CLASS_LATENCY = cpp_class_latency()

    import torch
    import torch.distributed as dist
    from longtermplanner_b200 import LongTermPlanner

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL announces its version on stdout when the first communicator comes up; stdout
        # carries exactly one JSON line, so that goes to stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    lim = W.FRANKA7
    n = args.n
    ltp = LongTermPlanner(lim.dof, lim.t_sample, *lim.arrays(), device=local)
    # this rank's shard of the global problem index space (contiguous, no overlap)
    qg, q0, v0, a0 = W.random_states(lim, n, W.SEEDS[2], start=rank * n)
    host_in = [torch.from_numpy(W.to_joint_major(x)).pin_memory() for x in (qg, q0, v0, a0)]
    dev_in = [t.to(dev) for t in host_in]
    sol = ltp.alloc_solution(n)

    # ---- device-resident solve: warm-up, then exactly K timed steps -------------------------
    for _ in range(args.warmup):
        ltp.solve(*dev_in, out=sol)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = ltp.launches
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e_start.record()
    for k in range(args.steps):
        ltp.solve(*dev_in, out=sol)
    e_stop.record()
    barrier()
    total_ms = max_over_ranks(e_start.elapsed_time(e_stop))
    gpu_launches = ltp.launches - launches0
    value = world * n * args.steps / (total_ms * 1e-3)
    # the dominant kernel's own launch duration: CUDA events recorded by the library on the
    # launching stream directly around each launch (ltp_set_profiling), K more steps
    ltp.setProfiling(True)
    for k in range(args.steps):
        ltp.solve(*dev_in, out=sol)
    per_kernel = {}  # per step: a large batch runs the first two kernels once per slice
    for kname in ("solve_fast", "solve_attempt2", "solve_items", "solve_generic"):
        k_ms, k_cnt = ltp.kernelTime(kname)
        per_kernel[kname] = k_ms / args.steps
    ltp.setProfiling(False)
    kernel_ms = per_kernel["solve_fast"]
    generic_kernel_ms = per_kernel["solve_generic"]

    # ---- end to end through the host-buffer C-ABI entry point ------------------------------
    host_np = [t.numpy() for t in host_in]
    host_out_t = {k: torch.empty(s, dtype=d).pin_memory() for k, s, d in (
        ("records", (lim.dof, n, 8), torch.float64), ("dir", (lim.dof, n), torch.float64),
        ("mod", (lim.dof, n), torch.uint8),
        ("slowest", (n,), torch.int32), ("traj_len", (n,), torch.int32), ("reached", (n,), torch.uint8))}
    host_out = {k: t.numpy() for k, t in host_out_t.items()}
    e2e_steps = max(3, min(args.steps, 10))

    def time_host_calls(out):
        for _ in range(2):
            ltp.solve_host(*host_np, out=out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ltp.solve_host(*host_np, out=out)  # synchronises internally; result lands in host memory
        torch.cuda.synchronize()
        return max_over_ranks(time.perf_counter() - t0)

    e2e_s = time_host_calls(host_out)
    e2e_value = world * n * e2e_steps / e2e_s
    assert np.array_equal(host_out["traj_len"], sol.traj_len.cpu().numpy())  # what was timed is what was solved
    h2d = sum(x.nbytes for x in host_np)
    d2h = sum(host_out[k].nbytes for k in host_out_t)  # (the call adds views of the records to the dict)
    # the same call with the output mask a caller sets who goes on to sample on the device or only
    # needs the durations: switching times, lengths and flags (ltp_solve_host: NULL = not copied)
    lean_out = {k: host_out[k] for k in ("records", "traj_len", "reached")}
    lean_s = time_host_calls(lean_out)
    lean_d2h = sum(host_out[k].nbytes for k in ("records", "traj_len", "reached"))
    # the ceiling: the same bytes as bare pinned copies in both directions at once
    probe_s = pcie_probe(host_in, list(host_out_t.values()), dev, 5, barrier, max_over_ranks)
    probe_lean_s = pcie_probe(host_in, [host_out_t[k] for k in ("records", "traj_len", "reached")], dev, 5, barrier,
                              max_over_ranks)
    clock_info = clocks.stop() if rank == 0 else None
    os.sched_setaffinity(0, all_cpus)  # the CPU baseline below uses every host core

    # a cheap integrity check on what was timed
    reached_frac = float(sol.reached.double().mean().item())
    assert reached_frac > 0.99, reached_frac

    # ---- parity of the timed workload at its full size: every one of this rank's 2^20 problems
    # against the reference's own CPU code (oracle/_ref; the checker, never the thing measured)
    parity = None
    if rank == 0 and not args.no_parity:
        try:
            parity, full, _ = full_parity(ltp, lim, dev_in, (qg, q0, v0, a0), len(all_cpus))
            parity.update({"workload": "configs[1], all problems of rank 0",
                           "matches_timed_output": bool(torch.equal(full.t_scaled, sol.t_scaled)
                                                        and torch.equal(full.traj_len, sol.traj_len))})
            del full
        except Exception as e:  # the checker is test infrastructure; never fail the bench on it
            parity = {"checked": 0, "unavailable": str(e)}

    extra = {}
    # free the configs[1] buffers before the memory-hungry sections
    del dev_in, host_in, host_np, host_out, host_out_t, lean_out
    torch.cuda.empty_cache()

    # ---- configs[4]: 2^26 random 12-DoF problems, full dense sampling, sharded over the ranks --
    if not args.no_stream:
        from longtermplanner_b200 import devtools
        lim12 = W.FRANKA12
        n_total = 1 << args.stream_log2n
        n_rank = n_total // world
        chunk, cap = 8192, 4096
        ltp12 = LongTermPlanner(lim12.dof, lim12.t_sample, *lim12.arrays(), device=local)
        ins12 = devtools.random_states_device(lim12, n_rank, W.SEEDS[5], start=rank * n_rank, device=local)
        ltp12.planStream(*[t[:, :2 * chunk + 7].contiguous() for t in ins12], chunk=chunk, capacity=cap)  # warm-up
        runs = {}
        # the problem-order run carries the parity consumer: every (2^k / 1024)-th problem of the run
        watch = StreamParity(lim12, W.SEEDS[5], rank * n_rank, max(n_total // 1024, 1), dev)
        for mode in ("sorted_slots", "problem_order"):
            launches12 = ltp12.launches
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            s0.record()
            st12 = ltp12.planStream(*ins12, chunk=chunk, capacity=cap, sorted_slots=(mode == "sorted_slots"),
                                    consumer=watch if (mode == "problem_order" and not args.no_parity) else None)
            s1.record()
            barrier()
            stream_ms = max_over_ranks(s0.elapsed_time(s1))
            tot = torch.tensor([st12["problems"], st12["bytes"], st12["success"], st12["clipped"], st12["reached"]],
                               dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tot)
            runs[mode] = (stream_ms, tot.tolist(), int(ltp12.launches - launches12))
        stream_ms, tot, launches_stream = runs["sorted_slots"]
        assert tot == runs["problem_order"][1], "both slot orders must produce the same totals"
        po_ms = runs["problem_order"][0]
        stream_parity = None
        if not args.no_parity:
            try:
                chk12, kind12 = cpu_checker(lim12)
                cnt = torch.tensor(watch.compare(chk12), dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(cnt)
                cnt = [int(x) for x in cnt.tolist()]
                stream_parity = {"checked": cnt[0], "exact_mismatch": cnt[1], "numeric_mismatch": cnt[2],
                                 "values_compared": cnt[3], "against": kind12,
                                 "what": "every (n/1024)-th problem of the run: traj_len and success exact; per joint "
                                         "the sums of q, v, a, j over the row, max |v|, max |a| and the last q, v "
                                         "(reduced on the device inside the stream consumer) within 1e-9 rel / "
                                         "1e-12 abs of the same reductions of the CPU checker's plan"}
            except Exception as e:
                stream_parity = {"checked": 0, "unavailable": str(e)}
        extra_stream = {
            "workload": f"configs[4]: 2^{args.stream_log2n} random 12-DoF dual-arm problems (FRANKA12), solve + "
                        "exact-length dense q/v/a/j sampling, contiguous problem-index shards, each rank streams its "
                        f"shard through a two-slot ring of {chunk}-problem chunks (time-major, capacity {cap} samples); "
                        "headline = sorted-slot mode (each chunk's problems ordered by trajectory length on the "
                        "device, slot k holds problem order[k]), problem-order mode beside it",
            "scaling": "strong", "seconds": stream_ms * 1e-3, "plans_per_s": tot[0] / (stream_ms * 1e-3),
            "write_gbs": tot[1] / (stream_ms * 1e-3) / 1e9, "write_gbs_per_gpu": tot[1] / (stream_ms * 1e-3) / 1e9 / world,
            "problem_order": {"seconds": po_ms * 1e-3, "plans_per_s": tot[0] / (po_ms * 1e-3),
                              "write_gbs": tot[1] / (po_ms * 1e-3) / 1e9,
                              "write_gbs_per_gpu": tot[1] / (po_ms * 1e-3) / 1e9 / world},
            "bytes_written": tot[1], "problems": tot[0], "reached": tot[4], "success": tot[2], "clipped": tot[3],
            "gpu_launches": launches_stream, "parity": stream_parity,
            "timing": "CUDA events around the call (it synchronises its two streams), max over ranks; the "
                      "problem-order run carries the parity consumer (a Python callback per chunk)"}
        del ins12, ltp12
        torch.cuda.empty_cache()
    if rank == 0:
        if not args.no_stream:
            extra["stream"] = extra_stream
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else \
            (6650.0, "fallback (B200_PROFILING.md)")
        # FP64 peak: no driver-written figure exists, so it is probed live
        probe = load_probe()
        tf, wgbs = C.c_double(0), C.c_double(0)
        probe.ltp_probe_fp64_tflops(local, 3, C.byref(tf))
        probe.ltp_probe_hbm_write_gbs(local, 3, C.byref(wgbs))
        fp64_peak = tf.value if tf.value > 0 else 37.0
        achieved_tf = (n / (kernel_ms * 1e-3)) * FLOP_PER_PLAN_7DOF / 1e12
        ncu, ncu_src = ncu_facts()
        alg_gbs = (n * (224 + 534) / (kernel_ms * 1e-3)) / 1e9
        extra["roofline"] = {
            "kernel": "ltp_solve_fast_kernel", "bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak,
            "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak,
            "traffic": ncu.get("ltp_solve_fast_kernel", {}).get("dram_bytes_per_launch"),
            "peak_source": "live DFMA probe (csrc/ltp_probe.cu); no FP64 figure in MEASURED_PEAKS.json",
            "algorithmic_flop_per_plan": FLOP_PER_PLAN_7DOF, "kernel_ms": kernel_ms,
            "kernel_share_of_step": kernel_ms / (total_ms / args.steps),
            "work_list_kernel_ms": generic_kernel_ms, "kernels_ms": per_kernel,
            "timing": "CUDA events on the launching stream around each launch, mean of K launches",
            "ncu_fp64_pipe_pct": ncu.get("ltp_solve_fast_kernel", {}).get("fp64_pipe_pct"),
            "ncu_capture": ncu_src,
            "hbm": {"achieved": alg_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": alg_gbs / hbm_peak,
                    "algorithmic_bytes_per_plan": 224 + 534, "peak_source": hbm_src}}

        # ---- sampler: configs[2], 4096 envs x 7 DoF, dense sampling to 2 s at 1 ms ----------
        if not args.no_sampler:
            n_env, horizon = 4096, 2001
            g2, s0, sv, sa = W.random_states(lim, n_env, W.SEEDS[3])
            d2 = [torch.from_numpy(W.to_joint_major(x)).to(dev) for x in (g2, s0, sv, sa)]
            sol2 = ltp.alloc_solution(n_env)
            reps = 20
            useful = n_env * lim.dof * horizon * 32
            res = {}
            for layout in ("rows", "time_major"):
                traj = ltp.alloc_trajectories(n_env, horizon, layout)
                for _ in range(3):
                    ltp.solve(*d2, out=sol2)
                    ltp.sample(d2[1], d2[2], d2[3], sol2, horizon=horizon, out=traj)
                torch.cuda.synchronize()
                ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(reps)]
                ltp.setProfiling(True)
                ltp.kernelTime("sample_time_major"), ltp.kernelTime("sample_rows")
                for r in range(reps):
                    ev[r][0].record()
                    ltp.solve(*d2, out=sol2)
                    ev[r][1].record()
                    ltp.sample(d2[1], d2[2], d2[3], sol2, horizon=horizon, out=traj)
                    ev[r][2].record()
                torch.cuda.synchronize()
                k_ms, k_cnt = ltp.kernelTime("sample_time_major" if layout == "time_major" else "sample_rows")
                ltp.setProfiling(False)
                res[layout] = (float(np.mean([e[0].elapsed_time(e[1]) for e in ev])),
                               float(np.mean([e[1].elapsed_time(e[2]) for e in ev])), k_ms / max(k_cnt, 1))
                del traj
            solve_ms, samp_ms, samp_kernel_ms = res["time_major"]
            gbs = useful / (samp_kernel_ms * 1e-3) / 1e9
            extra["sampler"] = {
                "workload": "configs[2]: 4096 envs x 7 DoF, fixed horizon 2001 samples (2 s at 1 ms), "
                            "output 1.84 GB per replan (larger than L2)",
                "layout": "time-major (samples, n, dof) f64 tensors q, v, a, j",
                "replan_ms": solve_ms + samp_ms, "solve_ms": solve_ms, "sample_ms": samp_ms,
                "sample_call_gbs": useful / (samp_ms * 1e-3) / 1e9,
                "rows_layout_sample_ms": res["rows"][1], "rows_layout_kernel_ms": res["rows"][2],
                "rows_layout_gbs": useful / (res["rows"][2] * 1e-3) / 1e9,
                "roofline": {"kernel": "ltp_sample_tm_kernel", "bound": "hbm", "achieved": gbs, "peak": hbm_peak,
                             "unit": "GB/s", "frac": gbs / hbm_peak,
                             "traffic": ncu.get("ltp_sample_tm_kernel", {}).get("dram_bytes_per_launch"),
                             "peak_source": hbm_src, "kernel_ms": samp_kernel_ms,
                             "timing": "CUDA events on the launching stream around each launch, mean of 20",
                             "algorithmic_bytes_per_sample": 32, "bytes_per_launch": useful,
                             "write_only_probe_gbs": wgbs.value}}

        # ---- configs[2] as a loop: replan every 10 ms from the state the previous plan reached ----
        if not args.no_replan and not args.no_sampler:
            extra["replan"] = replan_loop(ltp, lim, dev, hbm_peak)

        # ---- configs[0]: one planTrajectory at a time -----------------------------------------
        if not args.no_single and world == 1:
            extra["single_plan"] = single_plan_latency(ltp, lim)

        # ---- the root-solving path under load (reference test limits) ------------------------
        if not args.no_generic and world == 1:
            extra["generic"] = generic_section(LongTermPlanner, local, dev, max(3, min(args.steps, 10)),
                                               os.cpu_count() or 1)

        # ---- CPU baseline: the reference's code on this box's host cores ---------------------
        if not args.no_cpu and world == 1:
            cores = os.cpu_count() or 1
            try:
                r_all, kind = cpu_solve_rate(1 << 19, cores, W.SEEDS[2])
                r_one, _ = cpu_solve_rate(1 << 16, 1, W.SEEDS[2])
                o0 = None
                try:  # the reference's shipped flags (-std=c++17 only, CMakeLists.txt:5), labelled as such
                    from oracle.bindings import ReferenceO0
                    if ReferenceO0.available():
                        c0 = ReferenceO0.from_limits(W.FRANKA7)
                        s_qg, s_q0, s_v0, s_a0 = W.random_states(W.FRANKA7, 1 << 17, W.SEEDS[2])
                        c0.solve(s_qg[:2048], s_q0[:2048], s_v0[:2048], s_a0[:2048], threads=cores)
                        t0 = time.perf_counter()
                        c0.solve(s_qg, s_q0, s_v0, s_a0, threads=cores)
                        o0 = {"value": (1 << 17) / (time.perf_counter() - t0), "unit": "plans/s", "cores": cores,
                              "flags": ReferenceO0.flags, "sample": "first 2^17 problems"}
                except Exception as e:
                    o0 = {"unavailable": str(e)}
                extra["cpu_baseline"] = {
                    "value": r_all, "unit": "plans/s", "cores": cores, "kind": kind,
                    "single_thread_value": r_one, "shipped_flags": o0,
                    "sample": f"first 2^19 of the 2^20 problems with {cores} threads (and 2^16 with 1 thread), "
                              "solve only = reference cc:14-55; "
                              + ("reference .cc unmodified + Eigen shim" if kind == "reference" else "C restatement")
                              + ", g++ -O2 -ffp-contract=off"}
            except Exception as e:  # the checker is test infrastructure; never fail the bench on it
                extra["cpu_baseline"] = {"value": None, "unit": "plans/s", "cores": 0, "kind": "port",
                                         "sample": f"unavailable: {e}"}

        line = {
            "metric": METRIC, "value": value, "unit": "plans/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: 2^20 random 7-DoF problems per GPU (FRANKA7 limits, t_sample 1 ms), "
                                   "solve only (phase times, synchronisation, time scaling)",
                       "problems_per_gpu": n, "dof": lim.dof, "layout": "SoA joint-major [dof][n] f64",
                       "l2": "inputs 235 MB + outputs 560 MB per step, larger than the 126 MB L2",
                       "sharding": "contiguous problem-index shards, no collective on the data path"},
            "e2e": {"value": e2e_value, "unit": "plans/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "ltp_solve_host (pinned host buffers in and out), every solver output",
                    "host_numa_node": numa,
                    "achieved_gbs": world * (h2d + d2h) * e2e_steps / e2e_s / 1e9,
                    "pcie_peak_gbs": world * (h2d + d2h) / probe_s / 1e9,
                    "frac": probe_s / (e2e_s / e2e_steps),
                    "bound": "host<->device copies: the same bytes as bare concurrent pinned copies (both directions "
                             "at once, all ranks at once) take pcie_probe_ms per step; frac = that / the call",
                    "pcie_probe_ms": probe_s * 1e3,
                    "lean": {"value": world * n * e2e_steps / lean_s, "unit": "plans/s",
                             "outputs": "records (switching times + v_drive), traj_len, reached (output mask: the "
                                        "other fields NULL)",
                             "d2h_bytes_per_step": lean_d2h, "pcie_probe_ms": probe_lean_s * 1e3,
                             "frac": probe_lean_s / (lean_s / e2e_steps)}},
            "gpu_launches": int(gpu_launches), "clocks": clock_info, "reached_frac": reached_frac,
            "parity": parity,
        }
        line.update(extra)
        if "stream" in extra:  # configs[4] at the top level, so that a per-N table shows the strong scaling
            line["stream_seconds"] = extra["stream"]["seconds"]
            line["stream_plans_per_s"] = extra["stream"]["plans_per_s"]
            line["stream_write_gbs"] = extra["stream"]["write_gbs"]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
