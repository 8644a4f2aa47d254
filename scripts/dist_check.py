"""torchrun check of the row-partitioned path against the single-GPU path (same library) and the oracle.
usage: torchrun --nproc-per-node N scripts/dist_check.py [N_b]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "interpolation-based-immersed-fea_b200"))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import iife_b200 as I
from iife_b200 import dist as idist
from iife_b200 import synthetic

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
I.init(lr)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
I.set_stream(stream.cuda_stream)
idist.init_comm()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 12
dev = torch.device("cuda", lr)
g = synthetic.cube_operators(N)
n_f, n_b = g["n_f"], g["n_b"]
fpart, bpart = idist.row_partition(n_f, world), idist.row_partition(n_b, world)
f0, f1, b0, b1 = int(fpart[rank]), int(fpart[rank + 1]), int(bpart[rank]), int(bpart[rank + 1])


def block(t, r0, r1):
    rp, ci, v = t
    sl = slice(int(rp[r0]), int(rp[r1]))
    return (torch.from_numpy((rp[r0:r1 + 1] - rp[r0]).astype(np.int64)).to(dev), torch.from_numpy(ci[sl].astype(np.int64)).to(dev),
            torch.from_numpy(v[sl].copy()).to(dev))


ex = idist.DistExtraction(n_f, n_b, block(g["M"], f0, f1), block(g["A"], f0, f1))
A_loc_vals = torch.from_numpy(g["A"][2][int(g["A"][0][f0]):int(g["A"][0][f1])].copy()).to(dev)
A_blk = block(g["A"], f0, f1)
C = ex.numeric_csr(A_blk[0], A_blk[1], A_loc_vals)  # a freshly assembled block: pattern checked, then numeric
try:
    bad = A_blk[1].clone()
    if bad.numel():
        bad[0] = (bad[0] + 1) % n_f
    ex.numeric_csr(A_blk[0], bad, A_loc_vals)
    raise AssertionError("numeric_csr accepted a different pattern")
except ValueError:
    pass
rp, ci, v = C.to_csr(np.int64)
# single-GPU reference on every rank (small problem)
dM = I.DeviceMat.from_csr(n_f, n_b, *g["M"])
dA = I.DeviceMat.from_csr(n_f, n_f, *g["A"])
Cg, _ = I.ptap(dM, dA)
grp, gci, gv = Cg.to_csr(np.int64)
sl = slice(int(grp[b0]), int(grp[b1]))
assert np.array_equal(rp, grp[b0:b1 + 1] - grp[b0]), "row pointers differ"
assert np.array_equal(ci, gci[sl]), "columns differ"
assert np.allclose(v, gv[sl], rtol=1e-12, atol=1e-14 * np.abs(gv).max()), "values differ from the single-GPU product"
b_f = torch.from_numpy(g["b_f"][f0:f1].copy()).to(dev)
bb = ex.rhs(b_f)
bbg = dM.spmv(g["b_f"], trans=True)
assert np.allclose(bb.cpu().numpy(), bbg[b0:b1], rtol=1e-13, atol=0)
x = torch.zeros(b1 - b0, dtype=torch.float64, device=dev)
info = ex.solve(bb, x, rtol=1e-10, atol=1e-50)
xg = np.zeros(n_b)
ig = I.ksp_solve(Cg, bbg, xg, I.KSP_CG, I.PC_JACOBI, rtol=1e-10, atol=1e-50)
err = np.linalg.norm(x.cpu().numpy() - xg[b0:b1]) / np.linalg.norm(xg)
assert info.reason == ig.reason == 2, (info.reason, ig.reason)
assert abs(info.iterations - ig.iterations) <= 1, (info.iterations, ig.iterations)
assert err <= 1e-8, err
# FGMRES (restart shorter than the iteration count) over the NCCL layer
xf = torch.zeros(b1 - b0, dtype=torch.float64, device=dev)
infof = ex.solve(bb, xf, rtol=1e-10, atol=1e-50, method="gmres", restart=20)
xgf = np.zeros(n_b)
igf = I.ksp_solve(Cg, bbg, xgf, I.KSP_FGMRES, I.PC_JACOBI, rtol=1e-10, atol=1e-50, restart=20)
errf = np.linalg.norm(xf.cpu().numpy() - xgf[b0:b1]) / np.linalg.norm(xgf)
assert infof.reason == igf.reason == 2, (infof.reason, igf.reason)
assert abs(infof.iterations - igf.iterations) <= 2, (infof.iterations, igf.iterations)
assert errf <= 1e-8, errf
# transferToForeground: u_f = M u_b with ghost entries of u_b
u_f = ex.transfer_to_foreground(x, block(g["M"], f0, f1))
u_f_ref = dM.spmv(xg)[f0:f1]
assert np.allclose(u_f.cpu().numpy(), u_f_ref, rtol=1e-7, atol=1e-12 * np.abs(u_f_ref).max())
# second numeric call with new values reuses the plan
C2 = ex.numeric(A_loc_vals * 2.0)
v2 = C2.values()
assert np.allclose(v2, 2.0 * gv[sl], rtol=1e-14, atol=0)
dist.barrier()
if rank == 0:
    print(f"dist_check ok: world={world} N_b={N} cg its={info.iterations} (single {ig.iterations}) sol_err={err:.2e}; "
          f"fgmres its={infof.iterations} (single {igf.iterations}) sol_err={errf:.2e}")
dist.destroy_process_group()
