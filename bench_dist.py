"""bench_dist.py — the N > 1 arm of bench.py (strong scaling of the S1 cube over the GPUs of one box, one process per
GPU, launched by torchrun).  Bench tooling, not product: it drives the public row-partitioned API
(iife_b200.dist.DistExtraction: numeric / rhs / solve) and reports, besides the step time, the per-phase device times
that explain the scaling curve and a PARITY CHECK of the partitioned result against a single-GPU run of the same
library on rank 0 (same N_b): sum(b_b), ||A_b 1||_2, ||u_b||_2 and the CG iteration count."""
from __future__ import annotations

import ctypes
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist


def _max_over_ranks(x, dev):
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _sum_over_ranks(vals, dev):
    t = torch.tensor([float(v) for v in vals], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t.tolist()]


def run(args, I, stream, peak, peak_src, metric, unit, ClockSampler):
    from iife_b200 import dist as idist
    from iife_b200 import synthetic
    from iife_b200._lib import check, lib
    from iife_b200.core import synth_cube

    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device())
    idist.init_comm()
    N = args.cells
    sz = synthetic.cube_sizes(N)
    n_f, n_b = sz["n_f"], sz["n_b"]
    nnzA, nnzM, nnzC = synthetic.cube_nnz(N)
    fpart = idist.row_partition(n_f, world)
    f0, f1 = int(fpart[rank]), int(fpart[rank + 1])
    b_f = torch.empty(f1 - f0, dtype=torch.float64, device=dev)
    A, M = synth_cube(N, 1.0, f0, f1, b_f=b_f)
    I.sync()

    def tensors_of(mat):
        n_rows, _, nnz = mat.info()
        rp = torch.empty(n_rows + 1, dtype=torch.int32, device=dev)
        ci = torch.empty(nnz, dtype=torch.int32, device=dev)
        v = torch.empty(nnz, dtype=torch.float64, device=dev)
        check(lib.iife_mat_get_csr(mat.handle, ctypes.c_void_p(rp.data_ptr()), ctypes.c_void_p(ci.data_ptr()),
                                   ctypes.c_void_p(v.data_ptr()), 4, I.MEM_DEVICE))
        I.sync()
        return rp, ci, v

    A_t, M_t = tensors_of(A), tensors_of(M)
    del A, M
    t0 = time.perf_counter()
    ex = idist.DistExtraction(n_f, n_b, M_t, A_t)
    ex.numeric(A_t[2])
    I.sync()
    t_setup = time.perf_counter() - t0
    x = torch.zeros(ex.n_owned, dtype=torch.float64, device=dev)
    state = {}

    marks = []  # per timed step: events before numeric / rhs / solve / after (the phases of THIS step, not separate loops)

    def step(record=False):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if record else None
        if record:
            ev[0].record(stream)
        ex.numeric(A_t[2])
        if record:
            ev[1].record(stream)
        bb = ex.rhs(b_f)
        x.zero_()
        if record:
            ev[2].record(stream)
        state["info"] = ex.solve(bb, x)
        if record:
            ev[3].record(stream)
            marks.append(ev)
        state["bb"] = bb

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    I.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(torch.cuda.current_device()) as clocks:
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step(record=True)
        e1.record(stream)
        barrier()
    launches = I.launch_count()
    ms_step = _max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    info = state["info"]
    t_numeric = _max_over_ranks(sum(ev[0].elapsed_time(ev[1]) for ev in marks) / len(marks), dev)
    t_rhs = _max_over_ranks(sum(ev[1].elapsed_time(ev[2]) for ev in marks) / len(marks), dev)
    t_cg = _max_over_ranks(sum(ev[2].elapsed_time(ev[3]) for ev in marks) / len(marks), dev)
    its_cg = max(info.iterations, 1)

    # ---- per-phase device times (max over ranks) were taken inside the timed steps above; the local SpMV separately
    def timed(fn, reps, warm=1):
        for _ in range(warm):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        barrier()
        return _max_over_ranks(a.elapsed_time(b) / reps, dev)

    bb = state["bb"]
    # SpMV of the local operator block (roofline of the dominant kernel, per GPU)
    n_loc = ex.n_owned
    n_ext = n_loc + int(ex.ghost_ids.numel())
    xs = torch.ones(n_ext, dtype=torch.float64, device=dev)
    ys = torch.empty(n_loc, dtype=torch.float64, device=dev)
    nnz_loc = ex.C_op.nnz
    t_spmv = timed(lambda: ex.C_op.spmv(xs, ys), 20)
    B_spmv = 12 * nnz_loc + 4 * (n_loc + 1) + 8 * n_ext + 8 * n_loc

    # ---- parity of the partitioned result: three rank-gathered scalars against a single-GPU run on rank 0
    ex.numeric(A_t[2])
    bb = ex.rhs(b_f)
    x.zero_()
    info_p = ex.solve(bb, x)
    ones = torch.ones(n_ext, dtype=torch.float64, device=dev)
    a1 = torch.empty(n_loc, dtype=torch.float64, device=dev)
    ex.C_op.spmv(ones, a1)
    torch.cuda.synchronize()
    sums = _sum_over_ranks([bb.sum().item(), (a1 * a1).sum().item(), (x * x).sum().item()], dev)
    got = {"sum_b_b": sums[0], "norm_A_b_ones": float(np.sqrt(sums[1])), "norm_u_b": float(np.sqrt(sums[2])),
           "cg_iterations": int(info_p.iterations)}
    parity = None
    if rank == 0:
        try:
            bf1 = torch.empty(n_f, dtype=torch.float64, device=dev)
            A1, M1 = synth_cube(N, 1.0, b_f=bf1)
            C1, _ = I.ptap(M1, A1)
            bb1 = torch.empty(n_b, dtype=torch.float64, device=dev)
            M1.spmv(bf1, bb1, trans=True)
            x1 = torch.zeros(n_b, dtype=torch.float64, device=dev)
            i1 = I.ksp_solve(C1, bb1, x1, I.KSP_CG, I.PC_JACOBI, rtol=1e-8, atol=1e-9)
            o1 = torch.ones(n_b, dtype=torch.float64, device=dev)
            y1 = torch.empty(n_b, dtype=torch.float64, device=dev)
            C1.spmv(o1, y1)
            torch.cuda.synchronize()
            ref = {"sum_b_b": float(bb1.sum().item()), "norm_A_b_ones": float(torch.linalg.vector_norm(y1).item()),
                   "norm_u_b": float(torch.linalg.vector_norm(x1).item()), "cg_iterations": int(i1.iterations)}
            rel = {k: abs(got[k] - ref[k]) / max(abs(ref[k]), 1e-300) for k in ("sum_b_b", "norm_A_b_ones", "norm_u_b")}
            # PtAP and M^T b are sums in a fixed order per row: 1e-10; the solution is compared at the north star's
            # bar for solutions at matched KSP tolerances (1e-8), the iteration count within one
            ok = (rel["sum_b_b"] <= 1e-10 and rel["norm_A_b_ones"] <= 1e-10 and rel["norm_u_b"] <= 1e-8
                  and abs(got["cg_iterations"] - ref["cg_iterations"]) <= 1)
            parity = {"parity_check": "ok" if ok else "FAILED", "partitioned": got, "single_gpu": ref, "rel_diff": rel}
            del A1, M1, C1, bf1, bb1, x1, o1, y1
            I.plan_cache_clear()
        except Exception as exc:
            parity = {"parity_check": f"not run: {str(exc)[:160]}", "partitioned": got}
    barrier()

    # ---- end to end at N GPUs through DistExtraction: every step uploads the CSR arrays of this rank's rows of A_f (a
    # freshly assembled matrix, as in the single-GPU figure) and its block of b_f from pinned host memory, and brings this
    # rank's block of u_b back.  The pattern was routed once at setup; numeric_csr checks the new one against it.
    e2e = None
    if not getattr(args, "no_e2e", False):
        try:
            hrp = torch.empty(A_t[0].numel(), dtype=torch.int32).pin_memory()
            hci = torch.empty(A_t[1].numel(), dtype=torch.int32).pin_memory()
            hv = torch.empty(A_t[2].numel(), dtype=torch.float64).pin_memory()
            hb = torch.empty(b_f.numel(), dtype=torch.float64).pin_memory()
            hx = torch.empty(ex.n_owned, dtype=torch.float64).pin_memory()
            hrp.copy_(A_t[0])
            hci.copy_(A_t[1])
            hv.copy_(A_t[2])
            hb.copy_(b_f)
            drp, dci = torch.empty_like(A_t[0]), torch.empty_like(A_t[1])
            torch.cuda.synchronize()

            def e2e_step():
                # a freshly assembled block of A_f: all three CSR arrays cross PCIe, as in the single-GPU figure
                drp.copy_(hrp, non_blocking=True)
                dci.copy_(hci, non_blocking=True)
                A_t[2].copy_(hv, non_blocking=True)
                b_f.copy_(hb, non_blocking=True)
                ex.numeric_csr(drp, dci, A_t[2])
                bbe = ex.rhs(b_f)
                x.zero_()
                ex.solve(bbe, x)
                hx.copy_(x, non_blocking=True)
                torch.cuda.synchronize()

            n_e2e = max(1, min(args.steps, getattr(args, "e2e_steps", 3)))
            e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                e2e_step()
            barrier()
            dt = _max_over_ranks((time.perf_counter() - t0) / n_e2e, dev)
            h2d = _sum_over_ranks([hrp.numel() * 4 + hci.numel() * 4 + hv.numel() * 8 + hb.numel() * 8, hx.numel() * 8], dev)
            e2e = {"value": n_f / dt / 1e6, "unit": unit, "h2d_bytes_per_step": int(h2d[0]),
                   "d2h_bytes_per_step": int(h2d[1]), "ms_per_step": dt * 1e3, "steps": n_e2e,
                   "api": "iife_b200.dist.DistExtraction.numeric_csr/rhs/solve: every rank uploads the full CSR arrays of its "
                          "rows of A_f and its block of b_f from pinned host memory (same accounting as the single-GPU "
                          "figure; the pattern is checked on the device against the one routed at setup), u_b back"}
            del hrp, hci, hv, hb, hx, drp, dci
        except Exception as exc:  # the device-resident line must survive
            e2e = {"error": str(exc)[:200]}
    if rank == 0:
        value = n_f / (ms_step * 1e-3) / 1e6
        achieved = B_spmv / (t_spmv * 1e-3) / 1e9
        config = {"workload": f"BASELINE config 5: synthetic S1 fitted cube N_b={N}, row-partitioned", "n_f": n_f,
                  "n_b": n_b, "nnz_A_f": nnzA, "nnz_M": nnzM, "nnz_A_b": nnzC,
                  "ksp": "cg+jacobi rtol=1e-8 atol=1e-9 zero guess", "cg_iterations": info.iterations,
                  "cg_reason": info.reason_name, "setup_plus_first_numeric_ms": t_setup * 1e3,
                  "parallelism": f"row blocks over {world} GPUs: ghost rows (PtAP) via NCCL, halo + reductions (CG) "
                                 "over NVLink peer memory",
                  "l2": "inputs larger than L2 (no flush)",
                  "phases_ms": {"ptap_numeric_ms": t_numeric, "rhs_ms": t_rhs, "cg_ms": t_cg,
                                "cg_us_per_iteration": t_cg * 1e3 / its_cg, "spmv_local_ms": t_spmv}}
        if parity:
            config.update(parity)
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config, "clocks": clocks.summary(),
            "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_spmv_sell (local block of A_b, per GPU)", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": B_spmv, "launch_ms": t_spmv,
                         "ptap_numeric": {"ms": t_numeric}, "cg_iteration": {"ms": t_cg / its_cg}},
            "cpu_baseline": None,
        }
        print(json.dumps(line))
    dist.barrier()
