"""Generates tests/golden/full/cfg{1..4}_inputs.npz: the INPUTS of BASELINE configs 1-4 at their named sizes
(square/Linear/R6, hole_in_plate/Linear/R5 with 2 fields, square/Quadratic/R5, square/Linear/R6 with 3
fields) from the reference's shipped meshes and extraction operators (run in the build container, where /root/reference
exists):  python tests/golden/make_fullsize_inputs.py

Stored per case: the mesh arrays the dolfin-free foreground surrogate is assembled from (oracle/fixtures.py) and the
extraction operator M read by the mirror's readExOp.  Outputs are NOT stored: tests/test_fullsize_configs.py runs the
oracle on these inputs at test time (seconds) and compares the CUDA path with it (parity unpinned: PETSc is absent)."""
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
from oracle.mini_h5 import read_mesh  # noqa: E402

CASES = [
    ("cfg1", "square/Linear/R6", "p1", 1),
    ("cfg2", "hole_in_plate/Linear/R5", "elasticity", 2),
    ("cfg3", "square/Quadratic/R5", "p2", 1),
    ("cfg4", "square/Linear/R6", "p1x3", 3),
]


def main():
    os.makedirs(os.path.join(HERE, "full"), exist_ok=True)
    for name, mdir, kind, nfields in CASES:
        d = os.path.join(REF, mdir)
        pts, cells, mat = read_mesh(os.path.join(d, "mesh.h5"))
        out = dict(kind=kind, nfields=nfields, mesh_dir=mdir, points=pts, cells=cells.astype(np.int32), material=mat.astype(np.int8))
        if kind == "p2":
            cn = fx.read_cell_nodes(os.path.join(d, "cell_nodes.csv"))
            out["cell_nodes"] = cn.astype(np.int32)
            n_f = int(cn.max()) + 1
        else:
            n_f = len(pts) * (2 if kind == "elasticity" else 3 if kind == "p1x3" else 1)
        M = mirror.readExOp([os.path.join(d, "ExOp_Cons.csv")], NFields=nfields, n_f=n_f)
        out.update(n_f=n_f, n_b=M.getSize()[1], M_rowptr=M.rowptr.astype(np.int32), M_colind=M.colind.astype(np.int32), M_val=M.val)
        path = os.path.join(HERE, "full", name + "_inputs.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {mdir} n_f={n_f} n_b={out['n_b']} nnz(M)={M.val.size} -> {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
