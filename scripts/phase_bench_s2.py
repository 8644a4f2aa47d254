"""Per-phase device timings on the S2 "unfitted" stress case: wide rows, intermediate rows beyond
the slot plan's 256-entry limit (hashing kernels), thousands of empty rows of A_b.  Development aid.
usage: python scripts/phase_bench_s2.py N_f[:degree] [...]      e.g.  96:1 64:2"""
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


for spec in sys.argv[1:] or ["48:1"]:
    N, deg = (spec.split(":") + ["1"])[:2]
    N, deg = int(N), int(deg)
    t0 = time.time()
    g = synthetic.unfitted_operators(N, deg)
    n_f, n_b = g["n_f"], g["n_b"]
    A = I.DeviceMat.from_csr(n_f, n_f, *g["A"])
    M = I.DeviceMat.from_csr(n_f, n_b, *g["M"])
    nnzA, nnzM = A.nnz, M.nnz
    print(f"S2 N_f={N} p={deg}: n_f={n_f} n_b={n_b} nnzA={nnzA} nnzM={nnzM} host gen+upload {time.time()-t0:.1f}s", flush=True)
    t0 = time.time()
    plan = I.PtapPlan(M, A)
    I.sync()
    info = plan.info()
    print(f"  symbolic {1e3*(time.time()-t0):.1f} ms  {info}  bins {plan.bin_counts()}", flush=True)
    C = plan.numeric(M, A)
    plan.check()
    nnzC = info["nnz_c"]
    B = 12 * (nnzA + 2 * nnzM + nnzC) + 4 * (2 * (n_f + 1) + 2 * (n_b + 1)) - 4 * nnzC
    tmin, tmed = timed(lambda: plan.numeric(M, A, C=C))
    print(f"  numeric {tmin:.3f} ms (med {tmed:.3f})  alg {B/1e9:.3f} GB -> {B/tmin/1e6:.0f} GB/s = {B/tmin/1e6/PEAK:.3f} of measured;"
          f" nnz_inter {info['nnz_intermediate']}", flush=True)
    bf = torch.from_numpy(g["b_f"]).cuda()
    bb = torch.empty(n_b, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    M.spmv(bf, bb, trans=True)
    for kt, name in ((I.KSP_CG, "cg"), (I.KSP_FGMRES, "fgmres")):
        xs = torch.zeros(n_b, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        r = I.ksp_solve(C, bb, xs, kt, I.PC_JACOBI, max_it=20000)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"  {name}: {r.iterations} its, {r.reason_name}, {ms:.2f} ms, {ms/max(r.iterations,1)*1e3:.1f} us/it", flush=True)
    del plan, C, A, M
