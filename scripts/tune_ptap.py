"""Sweep the warp-kernel launch parameters of the numeric PtAP (development aid)."""
import os, sys, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "interpolation-based-immersed-fea_b200"))
import numpy as np, torch
import iife_b200 as I
from iife_b200 import synthetic
I.init(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); I.set_stream(stream.cuda_stream)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 92
nnzA, nnzM, nnzC = synthetic.cube_nnz(N)
sz = synthetic.cube_sizes(N)
B = 12 * (nnzA + 2 * nnzM + nnzC) + 4 * (2 * (sz["n_f"] + 1) + 2 * (sz["n_b"] + 1)) - 4 * nnzC
A, M = I.synth_cube(N)
plan = I.PtapPlan(M, A)
C = plan.numeric(M, A)
ref = None
def t(reps=3):
    plan.numeric(M, A, C=C); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); plan.numeric(M, A, C=C); e1.record(stream); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)
combos = [("new", lg1, lg2, wpc) for lg1 in (4,) for lg2 in (2, 3) for wpc in (2, 4, 8)]
for kind, lg1, lg2, wpc in combos:
    for k in ("IIFE_PTAP_LG1", "IIFE_PTAP_LG2", "IIFE_PTAP_WPC"): os.environ.pop(k, None)
    if kind == "old":
        continue
    os.environ["IIFE_PTAP_LG1"] = str(lg1); os.environ["IIFE_PTAP_LG2"] = str(lg2); os.environ["IIFE_PTAP_WPC"] = str(wpc)
    try:
        ms = t()
        plan.check()
        v = C.values()
        if ref is None: ref = v
        err = float(np.abs(v - ref).max() / np.abs(ref).max())
        print(f"N={N} lg1={lg1} lg2={lg2} wpc={wpc}: {ms:.3f} ms  {B/ms/1e6:.0f} GB/s  relerr_vs_first={err:.2e}", flush=True)
    except Exception as e:
        print(f"N={N} lg1={lg1} lg2={lg2} wpc={wpc}: FAILED {e}", flush=True)
