"""Wall time per call of the reference-facing API on the small BASELINE configs (golden fixtures from the
reference's shipped operators) next to the CPU oracle: these sizes are latency-bound on a B200."""
import glob
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "interpolation-based-immersed-fea_b200"))
sys.path.insert(0, ROOT)
import numpy as np

from InterpolationBasedImmersedFEA import common as api
from oracle import oracle as O

O.build()
O.set_threads(int(os.environ.get("ORACLE_THREADS", "4")))


def best(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3


for f in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))):
    g = dict(np.load(f))
    n_f, n_b = int(g["n_f"]), int(g["n_b"])
    method = "cg" if ("cg_reason" in g and int(g["cg_reason"]) > 0) else "gmres"
    mi = int(g["max_it"])

    def gpu_extract():
        A = api.CSRMat((n_f, n_f), g["A_rowptr"], g["A_colind"], g["A_val"])
        M = api.CSRMat((n_f, n_b), g["M_rowptr"], g["M_colind"], g["M_val"])
        return api.assembleLinearSystemBackground(A, api.Vec(g["b_f"]), M)

    A_b, b_b = gpu_extract()

    def gpu_solve():
        u = api.Vec(np.zeros(n_b))
        api.solveKSP(A_b, b_b, u, method=method, PC="jacobi", max_it=mi, monitor=False)
        return u

    Ao = O.CSR(n_f, n_f, g["A_rowptr"], g["A_colind"], g["A_val"])
    Mo = O.CSR(n_f, n_b, g["M_rowptr"], g["M_colind"], g["M_val"])
    Co = O.AT_R_A(Mo, Ao)
    bo = O.AT_x(Mo, g["b_f"])
    t_ge, t_gs = best(gpu_extract), best(gpu_solve)
    t_ce = best(lambda: (O.AT_R_A(Mo, Ao), O.AT_x(Mo, g["b_f"])))
    t_cs = best(lambda: O.solve_ksp(Co, bo, method=method, max_it=mi))
    its = api.last_ksp_info.iterations
    print(f"{os.path.basename(f)[:-4]:40s} n_f={n_f:6d} n_b={n_b:5d} | extract GPU {t_ge:7.2f} ms CPU {t_ce:7.2f} ms | "
          f"{method:5s} {its:4d} its GPU {t_gs:7.2f} ms CPU {t_cs:7.2f} ms", flush=True)
