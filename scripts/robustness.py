"""How much of the headline depends on the regularity of the synthetic cube?  Per-phase device times on operators that
defeat (a) the compact-column SELL layout of A_b, (b) the template numeric PtAP, (c) both:
  S1            lexicographic numbering (the headline workload)
  S1-bg         background numbering shuffled (columns of M, i.e. rows/columns of A_b): no shared column offsets, no
                shared row structure -> plain SELL columns, per-row PtAP kernels
  S1-bg-fg      foreground numbering shuffled as well (rows of A_f and M): random gathers in every product
  S2 p=1, p=2   the unfitted stress case (rotated background): wide rows, hashing kernels
Reports numeric PtAP, SpMV(A_b) and a CG iteration with their algorithmic-byte roofline fractions (MEASURED_PEAKS.json).
usage (GPU box): python scripts/robustness.py [N_b=184] [N_f(S2,p=1)=171] [N_f(S2,p=2)=128] > gpurun_out/robustness.md"""
import ctypes
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
from iife_b200._lib import check, lib

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
dev = torch.device("cuda", 0)


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
    return min(ts)


def tensors_of(mat):
    n_rows, _, nnz = mat.info()
    rp = torch.empty(n_rows + 1, dtype=torch.int32, device=dev)
    ci = torch.empty(nnz, dtype=torch.int32, device=dev)
    v = torch.empty(nnz, dtype=torch.float64, device=dev)
    check(lib.iife_mat_get_csr(mat.handle, ctypes.c_void_p(rp.data_ptr()), ctypes.c_void_p(ci.data_ptr()),
                               ctypes.c_void_p(v.data_ptr()), 4, I.MEM_DEVICE))
    I.sync()
    return rp, ci, v


def permute(mat, n_cols, row_perm=None, col_perm=None):
    """CSR of P_r A P_c^T: new row id = row_perm[old], new column id = col_perm[old]; columns re-sorted per row."""
    rp, ci, v = tensors_of(mat)
    n = rp.numel() - 1
    lens = (rp[1:] - rp[:-1]).to(torch.int64)
    rows = torch.repeat_interleave(torch.arange(n, device=dev), lens)
    if row_perm is not None:
        rows = row_perm[rows]
    cols = ci.to(torch.int64)
    if col_perm is not None:
        cols = col_perm[cols]
    key = rows * n_cols + cols
    del rows
    order = torch.argsort(key)
    key = key[order]
    v2 = v[order]
    del order, v, ci
    r2 = torch.div(key, n_cols, rounding_mode="floor")
    c2 = (key - r2 * n_cols).to(torch.int32)
    del key
    rp2 = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    rp2[1:] = torch.cumsum(torch.bincount(r2, minlength=n), 0)
    del r2
    torch.cuda.synchronize()
    out = I.DeviceMat.from_csr(n, n_cols, rp2.to(torch.int32), c2, v2)
    I.sync()
    return out


rows = []


def run_case(name, A, M, n_f, n_b, b_f):
    nnzA, nnzM = A.nnz, M.nnz
    I.plan_cache_clear()
    t0 = time.perf_counter()
    plan = I.PtapPlan(M, A)
    I.sync()
    t_sym = (time.perf_counter() - t0) * 1e3
    C = plan.numeric(M, A)
    plan.check()
    info = plan.info()
    nnzC = info["nnz_c"]
    t_num = timed(lambda: plan.numeric(M, A, C=C))
    ti = plan.tpl_info()
    bins = plan.bin_counts()
    B_num = 12 * (nnzA + 2 * nnzM + nnzC) + 4 * (2 * (n_f + 1) + 2 * (n_b + 1)) - 4 * nnzC
    bb = torch.empty(n_b, dtype=torch.float64, device=dev)
    M.spmv(b_f, bb, trans=True)
    x = torch.zeros(n_b, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    r = I.ksp_solve(C, bb, x, I.KSP_CG, I.PC_JACOBI, rtol=1e-8, atol=1e-9, max_it=400)  # builds the SELL copy
    xs = torch.ones(n_b, dtype=torch.float64, device=dev)
    ys = torch.empty(n_b, dtype=torch.float64, device=dev)
    t_spmv = timed(lambda: C.spmv(xs, ys), reps=10)
    B_spmv = 12 * nnzC + 4 * (n_b + 1) + 16 * n_b

    def cg():
        x.zero_()
        return I.ksp_solve(C, bb, x, I.KSP_CG, I.PC_JACOBI, rtol=1e-30, atol=1e-300, max_it=200)

    t_cg = timed(cg, reps=2) / 200.0
    B_cg = B_spmv + 88 * n_b
    rows.append((name, n_f, n_b, nnzA, nnzC, info["nnz_intermediate"], t_sym, t_num, B_num / t_num / 1e6 / PEAK,
                 f"{ti['rows']}/{sum(bins)} rows in {ti['templates']} templates", bins, t_spmv * 1e3, B_spmv / t_spmv / 1e6 / PEAK,
                 t_cg * 1e3, B_cg / t_cg / 1e6 / PEAK, r.iterations, r.reason_name))
    print(f"[{name}] done: numeric {t_num:.2f} ms, spmv {t_spmv*1e3:.0f} us, cg/it {t_cg*1e3:.0f} us", file=sys.stderr, flush=True)
    del plan, C


N = int(sys.argv[1]) if len(sys.argv) > 1 else 184
if not os.environ.get("ROBUST_ONLY_S2"):  # development: only the unfitted cases
    sz = synthetic.cube_sizes(N)
    n_f, n_b = sz["n_f"], sz["n_b"]
    b_f = torch.empty(n_f, dtype=torch.float64, device=dev)
    A, M = I.synth_cube(N, 1.0, b_f=b_f)
    I.sync()
    run_case(f"S1 N_b={N}", A, M, n_f, n_b, b_f)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    pb = torch.randperm(n_b, device=dev, generator=g)
    M2 = permute(M, n_b, col_perm=pb)
    run_case(f"S1 N_b={N}, background shuffled", A, M2, n_f, n_b, b_f)
    del M2
    pf = torch.randperm(n_f, device=dev, generator=g)
    M3 = permute(M, n_b, row_perm=pf, col_perm=pb)
    del M
    A3 = permute(A, n_f, row_perm=pf, col_perm=pf)
    del A
    bf3 = torch.empty_like(b_f)
    bf3[pf] = b_f
    run_case(f"S1 N_b={N}, background and foreground shuffled", A3, M3, n_f, n_b, bf3)
    del A3, M3, bf3, pf, pb, b_f
    torch.cuda.empty_cache()
for spec, deg in ((int(sys.argv[2]) if len(sys.argv) > 2 else 171, 1), (int(sys.argv[3]) if len(sys.argv) > 3 else 128, 2)):
    gg = synthetic.unfitted_operators(spec, deg)
    A = I.DeviceMat.from_csr(gg["n_f"], gg["n_f"], *gg["A"])
    M = I.DeviceMat.from_csr(gg["n_f"], gg["n_b"], *gg["M"])
    bf = torch.from_numpy(gg["b_f"]).to(dev)
    run_case(f"S2 N_f={spec} p={deg}", A, M, gg["n_f"], gg["n_b"], bf)
    del A, M, bf, gg

print("# Robustness of the headline: the same phases on operators without the cube's regularity\n")
print(f"`python scripts/robustness.py` on 1 x B200.  Fractions are algorithmic bytes (SURVEY.md §8d) over time over the measured "
      f"copy peak ({PEAK:.0f} GB/s).  CG: 200 iterations timed, tolerances disabled.\n")
print("| case | n_f | n_b | nnz(A_f) | nnz(A_b) | nnz(M^T A_f) | symbolic ms | numeric PtAP ms | frac | template rows | bins [5 hashing, 3 slot-plan: 128/32, 256/256, wide] | "
      "SpMV(A_b) us | frac | CG us/it | frac | CG its to 1e-8 |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---|---|---:|---:|---:|---:|---|")
for r in rows:
    print(f"| {r[0]} | {r[1]} | {r[2]} | {r[3]} | {r[4]} | {r[5]} | {r[6]:.0f} | {r[7]:.2f} | {r[8]:.3f} | {r[9]} | {r[10]} | {r[11]:.0f} | "
          f"{r[12]:.2f} | {r[13]:.0f} | {r[14]:.2f} | {r[15]} ({r[16]}) |")
