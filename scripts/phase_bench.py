"""Per-phase device timings of the extraction path on the synthetic cube (development aid).
usage: python scripts/phase_bench.py N_b [N_b ...]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "interpolation-based-immersed-fea_b200"))
import numpy as np
import torch

import iife_b200 as I
from iife_b200 import synthetic

I.init(0)
torch.cuda.set_device(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
I.set_stream(stream.cuda_stream)
PEAK = 6544.3
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


for N in [int(a) for a in sys.argv[1:]] or [23]:
    sz = synthetic.cube_sizes(N)
    n_f, n_b = sz["n_f"], sz["n_b"]
    nnzA, nnzM, nnzC = synthetic.cube_nnz(N)
    t0 = time.time()
    bf = torch.empty(n_f, dtype=torch.float64, device="cuda")
    A, M = I.synth_cube(N, 1.0, b_f=bf)
    I.sync()
    print(f"N_b={N} n_f={n_f} n_b={n_b} nnzA={nnzA} nnzM={nnzM} nnzC={nnzC} gen {time.time()-t0:.2f}s dev_bytes={I.device_bytes()/1e9:.2f} GB", flush=True)
    t0 = time.time()
    plan = I.PtapPlan(M, A)
    I.sync()
    t_sym = time.time() - t0
    print(f"  symbolic (host wall, incl. transpose) {t_sym*1e3:.1f} ms  info={plan.info()}", flush=True)
    C = plan.numeric(M, A)
    plan.check()
    B_ptap = 12 * (nnzA + 2 * nnzM + nnzC) + 4 * (2 * (n_f + 1) + 2 * (n_b + 1))
    tmin, tmed = timed(lambda: plan.numeric(M, A, C=C))
    print(f"  numeric {tmin:.3f} ms (med {tmed:.3f})  alg {B_ptap/1e9:.2f} GB -> {(B_ptap-4*nnzC)/tmin/1e6:.0f} GB/s = {(B_ptap-4*nnzC)/tmin/1e6/PEAK:.3f} of measured", flush=True)
    # SpMV on A_b, M, Mt
    x = torch.ones(n_b, dtype=torch.float64, device="cuda")
    y = torch.empty(n_b, dtype=torch.float64, device="cuda")
    B_spmv = 12 * nnzC + 4 * (n_b + 1) + 16 * n_b
    tmin, tmed = timed(lambda: C.spmv(x, y), reps=10, warm=2)
    print(f"  spmv(A_b) {tmin:.3f} ms  {B_spmv/tmin/1e6:.0f} GB/s = {B_spmv/tmin/1e6/PEAK:.3f}", flush=True)
    uf = torch.empty(n_f, dtype=torch.float64, device="cuda")
    B = 12 * nnzM + 4 * (n_f + 1) + 8 * n_b + 8 * n_f
    tmin, tmed = timed(lambda: M.spmv(x, uf), reps=5, warm=2)
    print(f"  spmv(M) {tmin:.3f} ms  {B/tmin/1e6:.0f} GB/s = {B/tmin/1e6/PEAK:.3f}", flush=True)
    bb = torch.empty(n_b, dtype=torch.float64, device="cuda")
    M.spmv(bf, bb, trans=True)
    B = 12 * nnzM + 4 * (n_b + 1) + 8 * n_b + 8 * n_f
    tmin, tmed = timed(lambda: M.spmv(bf, bb, trans=True), reps=5, warm=1)
    print(f"  spmvT(M) {tmin:.3f} ms  {B/tmin/1e6:.0f} GB/s = {B/tmin/1e6/PEAK:.3f}", flush=True)
    # CG
    for kt, name in ((I.KSP_CG, "cg"), (I.KSP_FGMRES, "fgmres")):
        if name == "fgmres" and os.environ.get("SKIP_FGMRES"):
            continue
        xs = torch.zeros(n_b, dtype=torch.float64, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        info = I.ksp_solve(C, bb, xs, kt, I.PC_JACOBI)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        its = max(info.iterations, 1)
        B_it = B_spmv + 88 * n_b
        extra = f" cg-iter alg {B_it/1e9:.3f} GB -> {B_it*its/ms/1e6:.0f} GB/s = {B_it*its/ms/1e6/PEAK:.3f}" if name == "cg" else ""
        print(f"  {name}: {info.iterations} its, reason {info.reason_name}, {ms:.2f} ms, {ms/its:.4f} ms/it{extra}", flush=True)
    tmin, tmed = timed(lambda: C.spmv(x, y), reps=10, warm=2)
    print(f"  spmv(A_b) after KSP (SELL-32 if accepted) {tmin:.3f} ms  {B_spmv/tmin/1e6:.0f} GB/s = {B_spmv/tmin/1e6/PEAK:.3f}", flush=True)
    del plan, C, A, M
