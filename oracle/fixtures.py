"""Dolfin-free surrogate foreground operators on the reference's real meshes — TEST INFRASTRUCTURE
ONLY (used by tests/golden/make_golden.py and the tests).

The reference assembles A_f with FEniCS (reference common.py:158-159), which is not available; what
the hot path needs from A_f is a realistic AIJ matrix on the same foreground dofs as the shipped
extraction operators.  These builders produce one from the shipped mesh data:

  * ``p1_operator``          P1 stiffness + mass on the material-2 cells (the immersed block,
                             reference demos/poisson.py:135-136), zero rows elsewhere
  * ``p1_elasticity``        2-field plane-stress elasticity blocks, interleaved dofs (node*2 + field),
                             E = 200e9, nu = 0.3 (reference demos/linear_elasticity.py:86-89)
  * ``p2_pattern_operator``  the P2 connectivity pattern of ``cell_nodes.csv`` (all node pairs of a cell)
                             with seeded SPD-like values

A_b is invariant under a renumbering of the foreground dofs, so exodus node numbering is used directly
(what ``readExOp`` calls exoID); the extraction operators come from the mirror's ``readExOp``.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .oracle import CSR


def _simplex_grads(P):
    """P: (nc, d+1, d) vertex coordinates.  Returns (vol (nc), grads (nc, d+1, d))."""
    d = P.shape[2]
    B = np.transpose(P[:, 1:, :] - P[:, :1, :], (0, 2, 1))  # columns = edge vectors
    det = np.linalg.det(B)
    vol = np.abs(det) / (2.0 if d == 2 else 6.0)
    Binv = np.linalg.inv(B)  # rows = gradients of lambda_1..d
    G = np.zeros_like(P)
    G[:, 1:, :] = Binv
    G[:, 0, :] = -Binv.sum(axis=1)
    return vol, G


def p1_operator(points, cells, material, sigma=1.0, block=2.0):
    """A_f = K + sigma * Mass on cells with material == block; b_f = load of f = 1."""
    sel = cells[material == block]
    n = len(points)
    P = points[sel]
    vol, G = _simplex_grads(P)
    nv = sel.shape[1]
    d = points.shape[1]
    Ke = vol[:, None, None] * np.einsum("cid,cjd->cij", G, G)
    mass_scale = vol / ((d + 1) * (d + 2))
    Me = mass_scale[:, None, None] * (np.ones((nv, nv)) + np.eye(nv))[None]
    E = Ke + sigma * Me
    rows = np.repeat(sel, nv, axis=1).ravel()
    cols = np.tile(sel, (1, nv)).ravel()
    A = sp.coo_matrix((E.ravel(), (rows, cols)), shape=(n, n)).tocsr()
    A.sort_indices()
    b = np.zeros(n)
    np.add.at(b, sel.ravel(), np.repeat(vol / nv, nv))
    return CSR.from_scipy(A), b


def p1_elasticity(points, cells, material, E_mod=200e9, nu=0.3, block=2.0):
    """2D plane-stress elasticity, 2 fields, interleaved dofs (2*node + field)."""
    assert points.shape[1] == 2
    sel = cells[material == block]
    n = len(points)
    vol, G = _simplex_grads(points[sel])
    D = E_mod / (1 - nu ** 2) * np.array([[1, nu, 0], [nu, 1, 0], [0, 0, (1 - nu) / 2]])
    nc = len(sel)
    Bm = np.zeros((nc, 3, 6))
    for a in range(3):
        Bm[:, 0, 2 * a] = G[:, a, 0]
        Bm[:, 1, 2 * a + 1] = G[:, a, 1]
        Bm[:, 2, 2 * a] = G[:, a, 1]
        Bm[:, 2, 2 * a + 1] = G[:, a, 0]
    Ke = vol[:, None, None] * np.einsum("cki,kl,clj->cij", Bm, D, Bm)
    dofs = np.stack([2 * sel[:, a] + f for a in range(3) for f in range(2)], axis=1)
    rows = np.repeat(dofs, 6, axis=1).ravel()
    cols = np.tile(dofs, (1, 6)).ravel()
    A = sp.coo_matrix((Ke.ravel(), (rows, cols)), shape=(2 * n, 2 * n)).tocsr()
    A.sort_indices()
    b = np.zeros(2 * n)
    np.add.at(b, dofs[:, 1::2].ravel(), np.repeat(-1e10 * vol / 3.0, 3))  # body force in -y, scaled to the modulus
    return CSR.from_scipy(A), b


def p2_pattern(cell_nodes, n):
    """structural pattern of a P2 operator: all (a, b) node pairs of every cell"""
    k = cell_nodes.shape[1]
    rows = np.repeat(cell_nodes, k, axis=1).ravel()
    cols = np.tile(cell_nodes, (1, k)).ravel()
    A = sp.coo_matrix((np.ones(rows.size), (rows, cols)), shape=(n, n)).tocsr()
    A.sort_indices()
    return A.indptr.astype(np.int64), A.indices.astype(np.int32)


def seeded_spd_values(rowptr, colind, seed=0, skew=0.0):
    """Deterministic values on a symmetric pattern: symmetric, strictly diagonally dominant (SPD);
    ``skew`` adds an antisymmetric part (the nonsymmetric Navier-Stokes-like config 4)."""
    n = len(rowptr) - 1
    rng = np.random.default_rng(seed)
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    lo, hi = np.minimum(rows, colind), np.maximum(rows, colind)
    # value of an off-diagonal pair from a hash of the unordered pair -> symmetric without a transpose
    key = (lo.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) + hi.astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)
           + np.uint64(seed))
    key ^= key >> np.uint64(29)
    key *= np.uint64(0xBF58476D1CE4E5B9)
    key ^= key >> np.uint64(32)
    u = (key >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    val = -(0.1 + u)
    if skew:
        sign = np.where(rows < colind, 1.0, -1.0)
        val = val + skew * sign * (0.5 + u)
    diag = rows == colind
    val[diag] = 0.0
    rowsum = np.zeros(n)
    np.add.at(rowsum, rows, np.abs(val))
    val[diag] = rowsum[rows[diag]] * 1.05 + 1.0
    _ = rng
    return val


def read_cell_nodes(path):
    return np.loadtxt(path, delimiter=",", dtype=np.int64, ndmin=2)


def fullsize_case(g, values_seed=0):
    """(A_f, M, b_f) of a BASELINE config 1-4 at its named size from a tests/golden/full/*_inputs.npz dictionary
    (tests/golden/make_fullsize_inputs.py): the foreground surrogate SURVEY.md §8d specifies per config, assembled from
    the stored mesh arrays.  ``values_seed`` reseeds the value generator of the seeded cases (config 4's value updates)."""
    import scipy.sparse as sp

    from .oracle import CSR

    kind = str(g["kind"])
    pts, cells, mat = g["points"], g["cells"].astype(np.int64), g["material"].astype(np.float64)
    if kind == "p1":
        A, b = p1_operator(pts, cells, mat)
    elif kind == "elasticity":
        A, b = p1_elasticity(pts, cells, mat)
    elif kind == "p1x3":
        A1, _ = p1_operator(pts, cells, mat)
        S = A1.to_scipy()
        S.data[:] = 1.0
        blk = sp.kron(S, np.ones((3, 3)), format="csr")
        blk.sort_indices()
        rp, ci = blk.indptr.astype(np.int64), blk.indices.astype(np.int32)
        A = CSR(blk.shape[0], blk.shape[1], rp, ci, seeded_spd_values(rp, ci, seed=values_seed, skew=0.1))
        b = np.cos(np.arange(blk.shape[0]) * 0.37)
    elif kind == "p2":
        n_f = int(g["n_f"])
        rp, ci = p2_pattern(g["cell_nodes"].astype(np.int64), n_f)
        A = CSR(n_f, n_f, rp, ci, seeded_spd_values(rp, ci, seed=values_seed))
        b = np.sin(np.arange(n_f) * 0.11) + 0.5
    else:
        raise ValueError(kind)
    M = CSR(int(g["n_f"]), int(g["n_b"]), g["M_rowptr"], g["M_colind"], g["M_val"])
    return A, M, b
