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


def _pipeline_worker(rank, world, port, out):
    """One rank of the partitioned multiscale pipeline: basis build of the rank's Morton chunk on the GPU, exchange of
    the coarse element matrices (all_gather; stands in for the reference's distributed coarse assembly), replicated
    coarse solve, weights of the OWNED cells back to the device, per-cell norms, all_reduce(sum) of the squared norms."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import coarse_solve as cs
    from common import lib_problem
    from conftest import msfec_mod
    from oracle import msfec_oracle as mo
    m = msfec_mod()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g_ref, L, pairing = 2, 2, "NED_RT"
    cells = mo.morton_cells(g_ref)
    n_total = len(cells)
    lo, hi = (rank * n_total) // world, ((rank + 1) * n_total) // world
    bb = m.BasisBuilder(lib_problem(m, pairing, L), device=0).run(cells[lo:hi], np.arange(lo, hi))
    M_loc = torch.tensor(bb.get_global_element_matrix()); r_loc = torch.tensor(bb.get_global_element_rhs())
    Ms = [torch.zeros_like(M_loc) for _ in range(world)]; rs = [torch.zeros_like(r_loc) for _ in range(world)]
    dist.all_gather(Ms, M_loc); dist.all_gather(rs, r_loc)
    w = cs.solve_coarse(pairing, g_ref, cells, torch.cat(Ms).numpy(), torch.cat(rs).numpy())
    bb.set_global_weights(w[lo:hi])
    t = torch.tensor(bb.solution_norms(hi - lo).sum(0))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    out[rank] = (t.numpy().copy(), w.copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_partitioned_pipeline_matches_single_rank():
    """Two ranks (sharing cuda:0; gloo carries the host-side exchange) reproduce the single-rank final norms: the basis
    build has no cross-rank coupling, the only collectives are the coarse-data exchange and the norm reduction."""
    mgr = mp.Manager()
    res = {}
    for world, port in ((1, 29593), (2, 29594)):
        out = mgr.dict()
        mp.spawn(_pipeline_worker, args=(world, port, out), nprocs=world, join=True)
        res[world] = dict(out)
    n1, w1 = res[1][0]
    for rank in (0, 1):
        n2, w2 = res[2][rank]
        assert np.abs(w2 - w1).max() <= 1e-10 * np.abs(w1).max()
        assert np.abs(n2 - n1).max() <= 1e-10 * np.abs(n1).max(), (n1, n2)
    assert (n1 > 0).all()
