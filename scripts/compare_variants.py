"""A/B check of the experimental kernels against the default ones INSIDE ONE PROCESS (the switches are environment
variables read at call time), so that one GPU call tells which variant is wrong and where.  Development aid.

  IIFE_PTAP_V2=1      slot-plan numeric kernel v2           -> values of A_b against the default kernel, per bin
  IIFE_PTAP_V3=1      stage 2 as a gather program (small-row bin)
  IIFE_PTAP_CTAIL=0   default kernel without the compacted second pass
  IIFE_SPMV_SELL_T=1  SELL path for M^T x                  -> against the CSR path
  IIFE_SPMV_ILP=1     two rows in flight per lane group for short rows (M x)  -> against k_spmv
  IIFE_KSP_PERSIST=1  persistent cooperative CG            -> iterations / reason / history / solution

usage: python scripts/compare_variants.py [N_b ...]      (default 8 23 46; plus the S2 cases 20:1 and 14:2)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "interpolation-based-immersed-fea_b200"))
import numpy as np

import iife_b200 as I
from iife_b200 import synthetic

I.init(0)


class env:
    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update({k: str(v) for k, v in self.kw.items()})

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def report(tag, ref, got, scale=None):
    d = np.abs(got - ref)
    s = np.abs(ref).max() if scale is None else scale
    bad = np.flatnonzero(d > 1e-11 * max(s, 1e-300))
    print(f"  {tag:<28s} max |diff| {d.max() if d.size else 0.0:.3e} (scale {s:.3e})  entries off: {bad.size}"
          + (f"  first at {bad[:5].tolist()}" if bad.size else ""), flush=True)
    return bad


def cases(args):
    for a in args:
        n = int(a)
        g = synthetic.cube_operators(n)
        yield f"S1 cube N_b={n}", g
    for n, p in ((20, 1), (14, 2)):
        yield f"S2 N_f={n} p={p}", synthetic.unfitted_operators(n, p)


for name, g in cases(sys.argv[1:] or ["8", "23", "46"]):
    n_f, n_b = g["n_f"], g["n_b"]
    A = I.DeviceMat.from_csr(n_f, n_f, *g["A"])
    M = I.DeviceMat.from_csr(n_f, n_b, *g["M"])
    plan = I.PtapPlan(M, A)
    print(f"{name}: n_f={n_f} n_b={n_b} bins {plan.bin_counts()}", flush=True)
    C0 = plan.numeric(M, A)
    rp, ci, v0 = C0.to_csr(np.int64)
    row_of = np.repeat(np.arange(n_b), np.diff(rp))
    for tag, kw in (("PTAP_V2", dict(IIFE_PTAP_V2=1)), ("PTAP_V3", dict(IIFE_PTAP_V3=1)),
                    ("PTAP_V3 again (program cached)", dict(IIFE_PTAP_V3=1)), ("PTAP_CTAIL=0", dict(IIFE_PTAP_CTAIL=0))):
        with env(**kw):
            try:
                v = plan.numeric(M, A, check_errors=True).values()
                bad = report(tag, v0, v)
                if bad.size:
                    rows = np.unique(row_of[bad])
                    print(f"    rows off: {rows.size}, first {rows[:8].tolist()}, their lengths {np.diff(rp)[rows[:8]].tolist()}")
            except Exception as exc:
                print(f"  {tag}: FAILED {exc}", flush=True)
    b_f = np.ascontiguousarray(g["b_f"])
    bb0 = M.spmv(b_f, trans=True)
    with env(IIFE_SPMV_SELL_T=1):
        try:
            report("SPMV_SELL_T", bb0, M.spmv(b_f, trans=True))
        except Exception as exc:
            print(f"  SPMV_SELL_T: FAILED {exc}", flush=True)
    xb = np.linspace(-1.0, 1.0, n_b)
    uf0 = M.spmv(xb)
    with env(IIFE_SPMV_ILP=1):
        try:
            report("SPMV_ILP (M x)", uf0, M.spmv(xb))
        except Exception as exc:
            print(f"  SPMV_ILP: FAILED {exc}", flush=True)
    for kt, kname in ((I.KSP_CG, "cg"),):
        x0 = np.zeros(n_b)
        i0 = I.ksp_solve(C0, bb0, x0, kt, I.PC_JACOBI, hist_len=4000)
        with env(IIFE_KSP_PERSIST=1):
            try:
                x1 = np.zeros(n_b)
                i1 = I.ksp_solve(C0, bb0, x1, kt, I.PC_JACOBI, hist_len=4000)
                k = min(i0.iterations, i1.iterations) + 1
                hd = np.abs(i1.history[:k] - i0.history[:k]) / (np.abs(i0.history[:k]) + 1e-300)
                print(f"  KSP_PERSIST {kname}: its {i0.iterations} -> {i1.iterations}, reason {i0.reason} -> {i1.reason}, "
                      f"max rel history diff {hd.max():.2e} at it {int(hd.argmax())}, "
                      f"|dx|/|x| {np.linalg.norm(x1 - x0) / (np.linalg.norm(x0) + 1e-300):.2e}", flush=True)
            except Exception as exc:
                print(f"  KSP_PERSIST: FAILED {exc}", flush=True)
    del plan, C0, A, M
