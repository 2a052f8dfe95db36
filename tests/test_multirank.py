"""World-size-2 tests on CPU (gloo): the N>1 path shards coarse cells into contiguous Morton chunks with no
data-path collective; what has to hold is that the chunks tile the cell range, that chunk contents do not
depend on the world size, and that the max-over-ranks reduction bench.py uses works."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_total = 512
    lo, hi = (rank * n_total) // world, ((rank + 1) * n_total) // world
    cells = bench.morton_cells(3, lo, hi)
    # every rank contributes its chunk; gather and compare with the single-rank enumeration
    gathered = [torch.zeros((n_total // world, 8, 3), dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.tensor(cells))
    full = torch.cat(gathered).numpy()
    ref = bench.morton_cells(3, 0, n_total)
    t = torch.tensor([10.0 * (rank + 1), 1.0], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[rank] = (bool(np.array_equal(full, ref)), float(t[0]), lo, hi)
    dist.barrier()
    dist.destroy_process_group()


def test_partition_and_reduction_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29591, out), nprocs=world, join=True)
    assert out[0][0] and out[1][0]
    assert out[0][1] == out[1][1] == 20.0          # max over ranks
    assert (out[0][2], out[0][3], out[1][2], out[1][3]) == (0, 256, 256, 512)


def test_morton_matches_oracle_enumeration():
    sys.path.insert(0, ROOT)
    import bench
    from oracle import msfec_oracle as mo
    assert np.array_equal(bench.morton_cells(3, 0, 512), mo.morton_cells(3))
    assert np.array_equal(bench.morton_cells(5, 1000, 1010), mo.morton_cells(5)[1000:1010])


def test_reference_arm_skips_nonzero_ranks():
    import subprocess
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
