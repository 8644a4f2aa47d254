"""A/B of the row-partitioned CG iteration over the GPUs of one box (development aid, run under torchrun):
builds the S1 cube's partitioned operator once, then times the solve for every combination of the iteration's
switches (read per solve by the library) and checks the solution norm and the iteration count against the first
variant.  usage: torchrun --nproc-per-node N scripts/dist_cg_ab.py [N_b]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "interpolation-based-immersed-fea_b200"))
import torch
import torch.distributed as dist

import iife_b200 as I
from iife_b200 import dist as idist
from iife_b200 import synthetic
from iife_b200._lib import check, lib
from iife_b200.core import synth_cube

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
I.init(local)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
I.set_stream(stream.cuda_stream)
idist.init_comm()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 184
sz = synthetic.cube_sizes(N)
n_f, n_b = sz["n_f"], sz["n_b"]
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
ex = idist.DistExtraction(n_f, n_b, M_t, A_t)
ex.numeric(A_t[2])
bb = ex.rhs(b_f)
x = torch.zeros(ex.n_owned, dtype=torch.float64, device=dev)


def barrier():
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()


def solve_timed(reps=3):
    x.zero_()
    info = ex.solve(bb, x)  # warm-up of this variant (graph capture, one-off plans)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        x.zero_()
        info = ex.solve(bb, x)
    b.record(stream)
    barrier()
    t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    nrm = (x * x).sum().reshape(1)
    dist.all_reduce(nrm)
    return float(t.item()), info.iterations, float(nrm.sqrt().item())


variants = [
    ("round-2 baseline: three kernels, prologue wait, fence+flag reductions", {"IIFE_CG_INTERIOR_FIRST": "0", "IIFE_P2P_LL": "0", "IIFE_CG_MERGED": "0"}),
    ("three kernels, interior first + packed reductions", {"IIFE_CG_INTERIOR_FIRST": "1", "IIFE_P2P_LL": "1", "IIFE_CG_MERGED": "0"}),
    ("two kernels (update + p merged), prologue wait", {"IIFE_CG_INTERIOR_FIRST": "0", "IIFE_P2P_LL": "1", "IIFE_CG_MERGED": "1"}),
    ("two kernels, interior first + packed reductions (default)", {"IIFE_CG_INTERIOR_FIRST": "1", "IIFE_P2P_LL": "1", "IIFE_CG_MERGED": "1"}),
]
extra = os.environ.get("AB_EXTRA", "")  # e.g. "IIFE_CG_PDL=1;IIFE_KSP_CHUNK=64"
for item in [e for e in extra.split(";") if e]:
    k, v = item.split("=")
    variants.append((f"default + {item}", {"IIFE_CG_INTERIOR_FIRST": "1", "IIFE_P2P_LL": "1", "IIFE_CG_MERGED": "1", k: v}))
ref = None
for rep in range(int(os.environ.get("AB_REPS", "2"))):
    for name, env in variants:
        saved = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        t, its, nrm = solve_timed()
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        if ref is None:
            ref = (its, nrm)
        ok = abs(its - ref[0]) <= 1 and abs(nrm - ref[1]) <= 1e-8 * abs(ref[1])
        if rank == 0:
            print(f"[w{world} N_b={N}] {name}: {t:.3f} ms, {its} its, {t * 1e3 / max(its, 1):.1f} us/iteration, ||u|| {nrm:.12e} "
                  f"{'ok' if ok else 'MISMATCH'}", flush=True)
dist.barrier()
