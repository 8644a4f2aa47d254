"""Time y = M x (transferToForeground) and check it against the CSR kernel (development aid; env selects the variant)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "interpolation-based-immersed-fea_b200"))
import torch
import iife_b200 as I
I.init(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); I.set_stream(stream.cuda_stream)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 184
from iife_b200 import synthetic
sz = synthetic.cube_sizes(N)
A, M = I.synth_cube(N, 1.0)
del A
torch.manual_seed(0)
x = torch.rand(sz["n_b"], dtype=torch.float64, device="cuda")
y = torch.empty(sz["n_f"], dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    M.spmv(x, y)
ts = []
for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); M.spmv(x, y); e1.record(stream)
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
print(f"{os.environ.get('TAG', '')}: M x  min {ts[0]:.4f} ms  med {ts[5]:.4f} ms  checksum {float(y.sum()):.12e}", flush=True)
