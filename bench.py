#!/usr/bin/env python
"""Basis-build benchmark (BASELINE.json metric: coarse cells/s and fine DoF-solves/s).

Workload (config 5 of BASELINE.json): synthetic 3D Ned_RT, rough random-field coefficient,
32^3 coarse cells x 3 local refinements.  The 32 768 coarse cells are enumerated in p4est Morton
order and split into contiguous chunks, one per rank/GPU (the reference's is_locally_owned
partition); the basis build needs no inter-GPU traffic, so there is no data-path collective --
torch.distributed (NCCL) is used for the barriers and the max-over-ranks reduction only.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--cells C]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one full pass of the hot path (assemble -> lift -> Krylov solve of all k rhs -> Gram)
over this rank's cells.  Prints ONE JSON line (rank 0).
"""
import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261017
RHS = "scale*(2*x-1)*(y^2-y)*(z^2-z); scale*(2*y-1)*(x^2-x)*(z^2-z); scale*(2*z-1)*(x^2-x)*(y^2-y)"


def load_binding():
    spec = importlib.util.spec_from_file_location("msfec_b200", os.path.join(ROOT, "mpi-msfec_b200", "msfec_b200.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["msfec_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def morton_cells(g_ref, lo, hi):
    """corners[hi-lo, 8, 3] of coarse cells lo..hi-1 in p4est z-order (host logic, no oracle import)."""
    idx = np.arange(lo, hi, dtype=np.int64)
    ijk = np.zeros((idx.size, 3), dtype=np.int64)
    for b in range(g_ref):
        for d in range(3):
            ijk[:, d] |= ((idx >> (3 * b + d)) & 1) << b
    H = 1.0 / (1 << g_ref)
    off = np.array([[v & 1, (v >> 1) & 1, v >> 2] for v in range(8)], float) * H
    return (ijk * H)[:, None, :] + off[None, :, :]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower() == "active":
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def fp64_peak(dev):
    """FP64 matrix peak measured in this run the way MEASURED_PEAKS.json measures bf16: cuBLAS GEMM
    (torch.matmul, 6144^3, best of 5, CUDA events).  Nominal B200 figure is 37 TFLOP/s."""
    import torch
    try:
        n = 6144
        a = torch.randn((n, n), dtype=torch.float64, device=dev)
        c = torch.randn((n, n), dtype=torch.float64, device=dev)
        torch.matmul(a, c)
        best = 1e30
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, c); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        del a, c
        torch.cuda.empty_cache()
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12, "measured in this run: cuBLAS DGEMM 6144^3 (torch.matmul fp64), best of 5"
    except Exception as exc:   # pragma: no cover
        return 37.0, f"nominal B200 FP64 (measurement failed: {exc})"


def hbm_peak():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------
# CPU arms (the ONLY places that touch oracle/): cpu_baseline sample and --impl reference
# ------------------------------------------------------------------------------------------
def _oracle_worker(args):
    lo, hi, g_ref, L = args
    # one BLAS/OpenMP thread per worker process: the pool already uses every core
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from oracle import msfec_oracle as mo
    prob = mo.Problem(pairing="NED_RT", n_refine_local=L, n_refine_global=g_ref, random_field_seed=SEED,
                      rhs_expr=RHS, rhs_constants={"scale": 100.0})
    cells = morton_cells(g_ref, lo, hi)
    acc = 0.0
    for i in range(hi - lo):
        M, r, *_ = mo.build_basis(prob, cells[i], lo + i)
        acc += float(M[0, 0])
    return acc


def cpu_oracle_throughput(n_sample, g_ref, L, cores):
    """cells/s of the oracle port (exact sparse direct solve per cell) on `cores` host processes."""
    import multiprocessing as mp
    per = max(1, n_sample // cores)
    jobs = [(i * per, (i + 1) * per, g_ref, L) for i in range(cores)]
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    # spawn: never fork a process that has initialised CUDA
    with mp.get_context("spawn").Pool(cores) as pool:
        pool.map(_oracle_worker, [(0, 1, g_ref, L)] * cores)     # warm imports / topology caches
        t0 = time.perf_counter()
        pool.map(_oracle_worker, jobs)
        dt = time.perf_counter() - t0
    return per * cores / dt, per * cores, dt


def cpu_compiled_throughput(mode, n_sample, g_ref, L, threads):
    """cells/s of the compiled C++ port (oracle/msfec_cpu.cpp) on `threads` host threads.  mode 'reference': the
    reference's own algorithm shape -- one right-hand side at a time, Schur-complement CG + GMRES(ILU(0)) at 1e-6
    (source/Ned_RT/ned_rt_basis.cc:637-847), what `use direct solver basis = false` of every shipped .prm runs;
    mode 'exact': banded LDL^T (the reference's solve_direct branch)."""
    from oracle import msfec_cpu as mc
    from oracle import msfec_oracle as mo
    prob = mo.Problem(pairing="NED_RT", n_refine_local=L, n_refine_global=g_ref, random_field_seed=SEED,
                      rhs_expr=RHS, rhs_constants={"scale": 100.0})
    cells = morton_cells(g_ref, 0, n_sample)
    ids = np.arange(n_sample, dtype=np.int64)
    mc.build_basis_ned_rt(prob, cells[:1], ids[:1], mode, 1)                       # load the library
    t0 = time.perf_counter()
    M, r, its = mc.build_basis_ned_rt(prob, cells, ids, mode, threads)
    dt = time.perf_counter() - t0
    assert np.isfinite(M).all()
    return n_sample / dt, n_sample, dt, its.mean(0)


def cpu_baseline_block(g_ref, L, cores, n_ref=0, with_variants=True):
    """cpu_baseline of the JSON line: the reference-shaped compiled port on all host cores, with the exact variants beside it."""
    n_ref = n_ref or 4 * cores
    v, n_done, dt, its = cpu_compiled_throughput("reference", n_ref, g_ref, L, cores)
    block = {"value": v, "unit": "coarse cells/s", "cores": cores, "kind": "port",
             "algorithm": "reference-shaped: per right-hand side Schur-complement CG + GMRES(ILU(0)) inner solves at 1e-6 "
                          "(ned_rt_basis.cc:637-847), compiled C++ (oracle/msfec_cpu.cpp), one cell per host thread",
             "sample": f"{n_done} cells of the same workload in {dt:.1f} s on {cores} threads; {its[0]:.0f} outer CG / "
                       f"{its[1]:.0f} inner GMRES iterations per cell (18 right-hand sides)"}
    if with_variants:
        ve, ne, dte, _ = cpu_compiled_throughput("exact", 8 * cores, g_ref, L, cores)
        vp, npy, dtp = cpu_oracle_throughput(16 * cores, g_ref, L, cores)
        block["variants"] = {
            "exact_cpp": {"value": ve, "unit": "coarse cells/s", "cores": cores, "kind": "port",
                          "sample": f"{ne} cells in {dte:.1f} s, banded LDL^T per cell (oracle/msfec_cpu.cpp, solve_direct shape)"},
            "exact_python_superlu": {"value": vp, "unit": "coarse cells/s", "cores": cores, "kind": "port",
                                     "sample": f"{npy} cells in {dtp:.1f} s, oracle/msfec_oracle.py (the parity oracle), {cores} processes"}}
    return block


def _emit(line, fd):
    """The ONE JSON line goes to the process' original stdout; everything else that writes to fd 1 during the run
    (NCCL prints its version banner there) was redirected to stderr."""
    os.write(fd, (json.dumps(line) + "\n").encode())


def main():
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--global-refinements", type=int, default=5)   # 32^3 coarse cells
    ap.add_argument("--local-refinements", type=int, default=3)
    ap.add_argument("--cells", type=int, default=0, help="use only the first C coarse cells (debug)")
    ap.add_argument("--cells-per-batch", type=int, default=0)
    ap.add_argument("--solver", default="auto", choices=["auto", "mf", "band", "minres"],
                    help="solver of the local problems (msfec_problem.solver); auto = what a verbatim .prm gets: "
                         "multifrontal LDL^T up to 3 local refinements, banded block LDL^T beyond")
    ap.add_argument("--cpu-sample", type=int, default=0, help="cells in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    g_ref, L = args.global_refinements, args.local_refinements
    n_total = args.cells if args.cells else 8 ** g_ref
    n_fine = 1944 + 1728 if L == 3 else None
    k = 18
    cores = os.cpu_count() or 1
    workload = (f"synthetic 3D Ned_RT, rough random field (seed {SEED}, sigma ln(10)/2), "
                f"{n_total} coarse cells (global refinements {g_ref}) x {L} local refinements")
    config = {"workload": workload, "pairing": "NED_RT", "coarse_cells": n_total, "local_refinements": L,
              "use_direct_solver_basis": False, "solver": args.solver,
              "partition": f"contiguous Morton chunks over {world} rank(s)",
              "cache": "inputs larger than L2 (per-step working set is GBs)"}

    # ---------------- reference arm: the reference's algorithm shape on the host cores ---------
    # (the reference itself cannot be built here: deal.II / Trilinos / p4est / MPI are absent)
    if args.impl == "reference":
        if rank != 0:
            return
        n_sample = args.cpu_sample or 8 * cores
        vals = []
        for _ in range(args.warmup + args.steps):
            v, n_done, dt, its = cpu_compiled_throughput("reference", n_sample, g_ref, L, cores)
            vals.append((v, dt))
        v = float(np.mean([a for a, _ in vals[args.warmup:]]))
        ms = float(np.mean([b for _, b in vals[args.warmup:]])) * 1e3
        line = {"metric": "basis build throughput", "value": v, "unit": "coarse cells/s", "impl": "reference",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "coarse cells/s", "cores": cores, "kind": "port",
                                 "algorithm": "reference-shaped: per right-hand side Schur-complement CG + GMRES(ILU(0)) inner "
                                              "solves at 1e-6 (ned_rt_basis.cc:637-847), compiled C++ (oracle/msfec_cpu.cpp)",
                                 "sample": f"{n_sample} cells of the same workload per step on {cores} host threads; "
                                           f"{its[0]:.0f} outer CG / {its[1]:.0f} inner GMRES iterations per cell"},
                "e2e": {"value": v, "unit": "coarse cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "fine_dof_solves_per_s": v * k * (n_fine or 0)}
        _emit(line, out_fd)
        return

    # ---------------- B200 arm ------------------------------------------------------------------
    import torch
    import torch.distributed as dist
    m = load_binding()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lo, hi = (rank * n_total) // world, ((rank + 1) * n_total) // world
    n_loc = hi - lo
    cells = morton_cells(g_ref, lo, hi)
    ids = np.arange(lo, hi, dtype=np.int64)
    prob = m.make_problem("NED_RT", n_refine_local=L, n_refine_global=g_ref, random_field_seed=SEED,
                          rhs_expression=RHS, rhs_constants="scale=100", cells_per_batch=args.cells_per_batch,
                          use_direct_solver_basis=0, solver=m.SOLVER[args.solver])
    bb = m.BasisBuilder(prob, device=local_rank)
    dev = torch.device("cuda", local_rank)
    # device-resident inputs/outputs for `value`; pinned host buffers for `e2e`
    d_c = torch.tensor(cells, dtype=torch.float64, device=dev).contiguous()
    d_i = torch.tensor(ids, dtype=torch.int64, device=dev)
    d_M = torch.empty((n_loc, k, k), dtype=torch.float64, device=dev)
    d_r = torch.empty((n_loc, k), dtype=torch.float64, device=dev)
    h_c = torch.tensor(cells, dtype=torch.float64).pin_memory()
    h_i = torch.tensor(ids, dtype=torch.int64).pin_memory()
    h_M = torch.empty((n_loc, k, k), dtype=torch.float64).pin_memory()
    h_r = torch.empty((n_loc, k), dtype=torch.float64).pin_memory()

    def step_device():
        bb.run_device(n_loc, d_c.data_ptr(), d_i.data_ptr(), d_M.data_ptr(), d_r.data_ptr())
        return bb.stats

    def step_e2e():
        st = m.Stats()
        rc = m.lib().msfec_build_basis(bb._ctx, n_loc, h_c.data_ptr(), h_i.data_ptr(), h_M.data_ptr(), h_r.data_ptr(),
                                       m.C.byref(st))
        if rc:
            raise m.MsfecError(rc, m.lib().msfec_last_error(bb._ctx).decode())
        return st.as_dict()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        t0 = time.perf_counter()
        stats = [fn() for _ in range(steps)]
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        barrier()
        ev_ms = sum(s["ms_total"] for s in stats)          # CUDA events on the library's stream
        t = torch.tensor([ev_ms, wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), stats

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev_ms, wall_ms, stats = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    step_e2e()                                             # warm the host path once
    e_ev_ms, e_wall_ms, e_stats = timed(step_e2e, args.steps)

    if rank == 0:
        ms_step = ev_ms / args.steps
        value = n_total / (ms_step * 1e-3)
        e2e_value = n_total / (e_wall_ms / args.steps * 1e-3)
        st = stats[-1]
        n_fine_dofs = st["n_fine_dofs"]
        solver_used = {0: "minres", 1: "band", 2: "mf"}[st["solver"]]
        hbm_pk, hbm_src = hbm_peak()
        if solver_used == "mf":
            # dominant kernel k_mf_forward (one CTA per front, fronts in shared memory): HBM-bound.  Algorithmic bytes
            # (MfPlan::bytes_fwd): slot values + lifted rhs read once, every factor panel written once,
            # every contribution block (lower triangle + rhs rows) written once and read once by its parent.  Time: CUDA
            # events around the forward launches of every sub-batch, on the library's stream.
            ms_f, ms_b = st["mf_ms_fwd"], st["mf_ms_bwd"]
            achieved = st["mf_bytes_fwd"] / (ms_f * 1e-3) / 1e9 if ms_f > 0 else 0.0
            traffic, traffic_src = None, None
            tpath = os.path.join(ROOT, "profiles", "r02_mf_forward_ncu.json")
            if os.path.exists(tpath):
                with open(tpath) as f:
                    tj = json.load(f)
                traffic, traffic_src = tj["dram_bytes_per_cell"] * n_total, "profiles/r02_mf_forward_ncu.json: " + tj["capture"]
            traffic_b = None
            bpath = os.path.join(ROOT, "profiles", "r02_mf_backward_ncu.json")
            if os.path.exists(bpath):
                with open(bpath) as f:
                    traffic_b = json.load(f)["dram_bytes_per_cell"] * n_total
            # measured DRAM bytes (ncu, all launches of both front kernels) over the measured time of those launches: the HBM
            # utilisation of the solve as a whole (what north_star's ">= 60 % of the HBM roofline" is judged on)
            hbm = None
            if traffic is not None and traffic_b is not None and ms_f + ms_b > 0:
                hbm = {"dram_bytes_per_step": traffic + traffic_b, "ms_per_step": ms_f + ms_b,
                       "achieved": (traffic + traffic_b) / ((ms_f + ms_b) * 1e-3) / 1e9,
                       "frac": (traffic + traffic_b) / ((ms_f + ms_b) * 1e-3) / 1e9 / hbm_pk,
                       "source": "ncu dram__bytes_read + dram__bytes_write of all k_mf_forward / k_mf_backward launches (profiles/r02_mf_*_ncu.json) "
                                 "over the CUDA-event time of those launches in this run"}
            fp64_pk, fp64_src = fp64_peak(dev)
            ach_b = st["mf_bytes_bwd"] / (ms_b * 1e-3) / 1e9 if ms_b > 0 else 0.0
            roofline = {"bound": "hbm", "kernel": "k_mf_forward", "achieved": achieved, "peak": hbm_pk, "unit": "GB/s",
                        "frac": achieved / hbm_pk, "traffic": traffic, "traffic_source": traffic_src, "peak_source": hbm_src,
                        "bytes_per_step": st["mf_bytes_fwd"], "ms_per_step": ms_f,
                        "backward": {"kernel": "k_mf_backward", "bytes_per_step": st["mf_bytes_bwd"], "ms_per_step": ms_b,
                                     "achieved": ach_b, "frac": ach_b / hbm_pk, "traffic": traffic_b},
                        "hbm": hbm,
                        "launches_per_step": int(st["mf_launches"]),
                        "step_frac": (st["mf_bytes_fwd"] + st["mf_bytes_bwd"]) / (ms_step * 1e-3) / 1e9 / hbm_pk,
                        "flops_per_step": st["mf_flops"], "step_fp64_tflops": st["mf_flops"] / (ms_step * 1e-3) / 1e12,
                        "fp64_peak_tflops": fp64_pk, "fp64_peak_source": fp64_src}
        elif solver_used == "minres":
            # dominant kernel k_minres_spmm: algorithmic bytes per launch: every per-cell matrix
            # value of the interior system once per iteration + rhs/solution amortised over the iterations.
            launches = max(1, st["krylov_spmm_launches"])
            bytes_per_launch = st["krylov_matrix_bytes"] / launches
            spmm_ms = st["krylov_ms_spmm"]
            peak, peak_src = hbm_peak()
            achieved = bytes_per_launch / (spmm_ms * 1e-3) / 1e9 if spmm_ms > 0 else 0.0
            roofline = {"bound": "hbm", "kernel": "k_minres_spmm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                        "bytes_per_launch": bytes_per_launch, "ms_per_launch": spmm_ms,
                        "launches_per_step": int(launches)}
        else:
            # dominant kernel k_direct_update_s (FP64 tensor-core trailing update of the block LDL^T):
            # algorithmic flops = 2 * 32 * (entries of the lower-triangular trailing region) per launch, summed
            # over the launches that were bracketed by CUDA events (first sub-batch of the step).
            peak, peak_src = fp64_peak(dev)
            ms_upd = st["direct_ms_update"]
            achieved = st["direct_flops_timed"] / (ms_upd * 1e-3) / 1e12 if ms_upd > 0 else 0.0
            # DRAM traffic per launch of that kernel from the committed ncu --set full capture (profiles/)
            traffic, traffic_src = None, None
            tpath = os.path.join(ROOT, "profiles", "r01_update_s_ncu.json")
            if os.path.exists(tpath):
                with open(tpath) as f:
                    tj = json.load(f)
                traffic, traffic_src = tj["mean_dram_bytes_per_launch"], "profiles/r01_update_s_ncu.json: " + tj["capture"]
            roofline = {"bound": "tensor", "kernel": "k_direct_update_s", "achieved": achieved, "peak": peak,
                        "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                        "peak_source": peak_src,
                        "flops_per_step": st["direct_flops"], "flops_timed": st["direct_flops_timed"],
                        "ms_timed": ms_upd, "launches_timed": int(st["direct_timed_launches"]),
                        "launches_per_step": int(st["direct_update_launches"]),
                        "step_fp64_tflops": st["direct_flops"] / (ms_step * 1e-3) / 1e12}
        line = {
            "metric": "basis build throughput", "value": value, "unit": "coarse cells/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "fine_dof_solves_per_s": value * k * n_fine_dofs,
            "wall_ms_per_step": wall_ms / args.steps,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "coarse cells/s",
                    "h2d_bytes_per_step": int(h_c.numel() * 8 + h_i.numel() * 8),
                    "d2h_bytes_per_step": int(h_M.numel() * 8 + h_r.numel() * 8),
                    "timing": "host wall clock around msfec_build_basis (pinned host buffers in, host buffers out), max over ranks"},
            "gpu_launches": int(sum(s["kernel_launches"] for s in stats)),
            "roofline": roofline,
            "solver": solver_used,
            "phases_ms": {kk: st[kk] for kk in ("ms_assemble", "ms_lift", "ms_solve", "ms_gram")},
            "krylov": {"iterations_max": st["iterations_max"], "iterations_mean": st["iterations_mean"],
                       "residual_max": st["residual_max"], "not_converged": st["not_converged"]},
        }
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_block(g_ref, L, cores, n_ref=args.cpu_sample, with_variants=(world == 1))
        _emit(line, out_fd)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
