"""Generates the committed golden fixtures under tests/golden/ from the reference's shipped data
(run in the build container, where /root/reference exists):

    python tests/golden/make_golden.py

Inputs pinned per case: the real extraction operator M (``meshes/**/ExOp_Cons.csv`` read with the
mirror's ``readExOp``: 1-based ids, INSERT-overwrite of duplicates, field-major background blocks,
reference common.py:645-712) and a dolfin-free surrogate A_f / b_f on the real foreground mesh
(oracle/fixtures.py).  Outputs recorded: what the ORACLE (the CPU restatement of the reference's PETSc
call sequence — parity unpinned, no PETSc here) produces for them: pattern and values of A_b, b_b, and
the Jacobi-CG / FGMRES(300) iteration counts and solutions.  The GPU tests compare the CUDA path with
these files on the GPU box, where /root/reference does not exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "interpolation-based-immersed-fea_b200"))
REF = "/root/reference/meshes"

from InterpolationBasedImmersedFEA import common as mirror  # noqa: E402
from oracle import fixtures as fx  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle.mini_h5 import read_mesh  # noqa: E402

MAX_IT = 400

CASES = [
    # name, mesh dir, kind, NFields
    ("cfg1_square_linear_R4", "square/Linear/R4", "p1", 1),
    ("cfg2_hole_in_plate_linear_R3_2field", "hole_in_plate/Linear/R3", "elasticity", 2),
    ("cfg3_square_quadratic_R3", "square/Quadratic/R3", "p2", 1),
    ("cfg4_square_linear_R4_3field", "square/Linear/R4", "p1x3", 3),
    ("cube_linear_R1", "cube/Linear/R1", "p1", 1),
    ("cube_quadratic_R0", "cube/Quadratic/R0", "p2", 1),
]


def build_case(name, mdir, kind, nfields):
    d = os.path.join(REF, mdir)
    pts, cells, mat = read_mesh(os.path.join(d, "mesh.h5"))
    meta = {}
    if kind == "p1":
        A, b = fx.p1_operator(pts, cells, mat)
        n_f = len(pts)
    elif kind == "elasticity":
        A, b = fx.p1_elasticity(pts, cells, mat)
        n_f = 2 * len(pts)
    elif kind == "p1x3":
        # config 4: three fields on the scalar pattern (block-full coupling), nonsymmetric seeded values
        A1, _ = fx.p1_operator(pts, cells, mat)
        S = A1.to_scipy()
        S.data[:] = 1.0
        import scipy.sparse as sp

        blk = sp.kron(S, np.ones((3, 3)), format="csr")
        blk.sort_indices()
        rp, ci = blk.indptr.astype(np.int64), blk.indices.astype(np.int32)
        A = O.CSR(blk.shape[0], blk.shape[1], rp, ci, fx.seeded_spd_values(rp, ci, seed=0, skew=0.1))
        b = np.cos(np.arange(blk.shape[0]) * 0.37)
        n_f = blk.shape[0]
        meta["values"] = "seeded_spd_values(seed=0, skew=0.1)"
    elif kind == "p2":
        cn = fx.read_cell_nodes(os.path.join(d, "cell_nodes.csv"))
        n_f = int(cn.max()) + 1
        rp, ci = fx.p2_pattern(cn, n_f)
        A = O.CSR(n_f, n_f, rp, ci, fx.seeded_spd_values(rp, ci, seed=0))
        b = np.sin(np.arange(n_f) * 0.11) + 0.5
        meta["values"] = "seeded_spd_values(seed=0)"
    M = mirror.readExOp([os.path.join(d, "ExOp_Cons.csv")], NFields=nfields, n_f=n_f)
    Mo = O.CSR(M.getSize()[0], M.getSize()[1], M.rowptr, M.colind, M.val)
    return A, Mo, b, meta


def main():
    O.build()
    for name, mdir, kind, nfields in CASES:
        A, M, b, meta = build_case(name, mdir, kind, nfields)
        C, ATR = O.AT_R_A(M, A, return_intermediate=True)
        bb = O.AT_x(M, b)
        # max_it bounded: cut-cell conditioning makes Jacobi-Krylov slow on some cases (the reference uses
        # MUMPS there); DIVERGED_ITS (-3) is then the recorded, parity-checked outcome
        cg = O.solve_ksp(C, bb, method="cg", rtol=1e-8, atol=1e-9, max_it=MAX_IT, hist_len=MAX_IT + 1) if kind != "p1x3" else None
        gm = O.solve_ksp(C, bb, method="gmres", rtol=1e-8, atol=1e-9, restart=300, max_it=MAX_IT, hist_len=MAX_IT + 1)
        out = dict(
            n_f=A.n_rows, n_b=M.n_cols,
            A_rowptr=A.rowptr.astype(np.int32), A_colind=A.colind, A_val=A.val,
            M_rowptr=M.rowptr.astype(np.int32), M_colind=M.colind, M_val=M.val, b_f=b,
            C_rowptr=C.rowptr.astype(np.int32), C_colind=C.colind, C_val=C.val, nnz_inter=ATR.nnz, b_b=bb,
            gm_its=gm.iterations, gm_reason=gm.reason, gm_x=gm.x, gm_hist=gm.history[:gm.iterations + 1], max_it=MAX_IT,
        )
        if cg is not None:
            out.update(cg_its=cg.iterations, cg_reason=cg.reason, cg_x=cg.x, cg_hist=cg.history[:cg.iterations + 1])
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: n_f={A.n_rows} n_b={M.n_cols} nnz(A_f)={A.nnz} nnz(M)={M.nnz} nnz(A_b)={C.nnz} "
              f"max A_b row={int(np.diff(C.rowptr).max())} empty A_b rows={int((np.diff(C.rowptr) == 0).sum())} "
              f"cg={None if cg is None else (cg.iterations, cg.reason)} gmres={(gm.iterations, gm.reason)} "
              f"-> {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
