"""CPU suite, part 4: the multi-GPU path shards by contiguous problem index with no collective
on the data path. Two gloo ranks each regenerate and solve their own shard (CPU checker stands
in for the device here; the sharding and reduction logic is what is under test), and the
gathered result must equal the single-process solve of the whole batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from longtermplanner_b200 import workloads as W


def test_shards_tile_the_global_index_space():
    lim, n = W.FRANKA7, 1000
    whole = W.random_states(lim, 4 * n, W.SEEDS[2])
    for r in range(4):
        part = W.random_states(lim, n, W.SEEDS[2], start=r * n)
        for a, b in zip(whole, part):
            assert np.array_equal(a[r * n:(r + 1) * n], b)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle.bindings import OraclePort
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lim = W.FRANKA7
    qg, q0, v0, a0 = W.random_states(lim, n, W.SEEDS[2], start=rank * n)
    s = OraclePort.from_limits(lim).solve(qg, q0, v0, a0)
    # what bench.py reduces: units processed and the max of the per-rank times
    stats = torch.tensor([float(n), float(rank + 1)], dtype=torch.float64)
    total = stats.clone()
    dist.all_reduce(total[:1], op=dist.ReduceOp.SUM)
    tmax = stats[1:].clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    lens = torch.from_numpy(s["traj_len"].astype(np.int64))
    gathered = [torch.empty_like(lens) for _ in range(world)] if rank == 0 else None
    dist.gather(lens, gathered, dst=0)
    if rank == 0:
        np.savez(out_path, traj_len=torch.cat(gathered).numpy(), total=total[0].item(), tmax=tmax[0].item())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_run_equals_single_process(tmp_path):
    from oracle.bindings import OraclePort
    world, n = 2, 3000
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(world, _free_port(), n, out), nprocs=world, join=True)
    z = np.load(out)
    lim = W.FRANKA7
    whole = OraclePort.from_limits(lim).solve(*W.random_states(lim, world * n, W.SEEDS[2]))
    assert np.array_equal(z["traj_len"], whole["traj_len"])
    assert z["total"] == world * n and z["tmax"] == world
