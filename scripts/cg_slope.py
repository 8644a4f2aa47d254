"""Per-iteration cost of the device CG without setup: slope between two max_it settings (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "interpolation-based-immersed-fea_b200"))
import numpy as np, torch
import iife_b200 as I
I.init(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); I.set_stream(stream.cuda_stream)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 184
from iife_b200 import synthetic
sz = synthetic.cube_sizes(N)
bf = torch.empty(sz["n_f"], dtype=torch.float64, device="cuda")
A, M = I.synth_cube(N, 1.0, b_f=bf)
C, _ = I.ptap(M, A)
bb = M.spmv(bf, trans=True)
def run(mi):
    x = torch.zeros(sz["n_b"], dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); info = I.ksp_solve(C, bb, x, I.KSP_CG, I.PC_JACOBI, rtol=1e-30, atol=1e-300, max_it=mi); e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), info.iterations
run(32)
for rep in range(2):
    t1, i1 = run(64); t2, i2 = run(192); t3, i3 = run(448)
    print(f"N={N}: {i1} its {t1:.2f} ms, {i2} its {t2:.2f} ms, {i3} its {t3:.2f} ms -> slope {1e3*(t2-t1)/(i2-i1):.1f} us/it, {1e3*(t3-t2)/(i3-i2):.1f} us/it; setup ~{t1 - i1*(t2-t1)/(i2-i1):.2f} ms", flush=True)
