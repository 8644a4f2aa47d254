"""Synthetic "S1 fitted cube" operands (BASELINE config 5) — bench / test tooling.

The reference's only mesh generator is ``generateUnfittedMesh(dim=3)`` (reference common.py:80-90:
dolfin ``BoxMesh`` = cubes split into 6 Kuhn tetrahedra); the shipped ``meshes/cube/Linear`` data are
fitted XTK decompositions where the foreground refines the background grid.  This module produces
the same kind of operands without dolfin:

  background : ``N^3`` cells, trilinear B-splines, ``n_b = (N+1)^3``, id = bx + (N+1)(by + (N+1) bz)
  foreground : each background cell split 2x2x2, each sub-cube into 6 Kuhn tetrahedra,
               ``n_f = (2N+1)^3`` P1 vertices, id = x + nv (y + nv z)
  A_f = K + sigma * Mass (P1), M = trilinear interpolation (rows sum to 1), b_f = load of f = 1.

``element_tables`` returns the per-cell contribution tables consumed by the device generator
(``iife_synth_cube_build`` in csrc/synth.cu); ``cube_operators`` is the host (numpy) generator that
adds the same table entries in the same order, so both agree bit for bit.
"""
from __future__ import annotations

import itertools

import numpy as np

KUHN_PERMS = list(itertools.permutations(range(3)))


def kuhn_tets():
    """The 6 tetrahedra of the Kuhn triangulation of the unit cube, as 4x3 integer vertex arrays."""
    tets = []
    for perm in KUHN_PERMS:
        v = np.zeros((4, 3), dtype=np.int64)
        for s, axis in enumerate(perm):
            v[s + 1] = v[s]
            v[s + 1, axis] += 1
        tets.append(v)
    return tets


def p1_element(verts: np.ndarray):
    """P1 stiffness, mass and load (f = 1) of one tetrahedron with vertex coordinates ``verts`` (4x3)."""
    B = (verts[1:] - verts[0]).T.astype(np.float64)  # columns = edge vectors
    vol = abs(np.linalg.det(B)) / 6.0
    Binv = np.linalg.inv(B)
    G = np.zeros((4, 3))
    G[1:] = Binv
    G[0] = -Binv.sum(axis=0)
    K = vol * (G @ G.T)
    Mm = vol / 20.0 * (np.ones((4, 4)) + np.eye(4))
    load = np.full(4, vol / 4.0)
    return K, Mm, load


def element_tables(n_bg_cells: int, sigma: float = 1.0, h: float | None = None):
    """coef[c, d] (8 x 27) and load8[c] (8): contribution of ONE foreground cell to the pair
    (vertex at local corner c, vertex at c + d) and to the load at corner c.  c = cx + 2 cy + 4 cz,
    d = (dx+1) + 3 (dy+1) + 9 (dz+1).  ``h`` = foreground cell size (default: the S1 cube's 1/(2N))."""
    if h is None:
        h = 1.0 / (2.0 * n_bg_cells)
    coef = np.zeros((8, 27))
    load8 = np.zeros(8)
    for tet in kuhn_tets():
        K, Mm, load = p1_element(tet.astype(np.float64) * h)
        E = K + sigma * Mm
        for a in range(4):
            ca = int(tet[a, 0] + 2 * tet[a, 1] + 4 * tet[a, 2])
            load8[ca] += load[a]
            for b in range(4):
                dd = tet[b] - tet[a]
                d = int((dd[0] + 1) + 3 * (dd[1] + 1) + 9 * (dd[2] + 1))
                coef[ca, d] += E[a, b]
    return coef, load8


def cube_sizes(n_bg_cells: int):
    nv = 2 * n_bg_cells + 1
    nb = n_bg_cells + 1
    return {"n_f": nv ** 3, "n_b": nb ** 3, "nv": nv, "nb": nb}


def _kuhn_grid_rows(nv: int, coef, load8, row_begin: int, row_end: int):
    """Rows [row_begin, row_end) of the P1 operator on an (nv-1)^3-cell Kuhn grid from the per-cell tables:
    (x, y, z, a_rowptr, a_col, a_val, b_f)."""
    j = np.arange(row_begin, row_end, dtype=np.int64)
    x, y, z = j % nv, (j // nv) % nv, j // (nv * nv)
    ncell = nv - 1
    cell_ok = []
    for c in range(8):
        cx, cy, cz = x - (c & 1), y - ((c >> 1) & 1), z - ((c >> 2) & 1)
        cell_ok.append((cx >= 0) & (cy >= 0) & (cz >= 0) & (cx < ncell) & (cy < ncell) & (cz < ncell))
    n = j.size
    cols = np.zeros((n, 27), dtype=np.int64)
    vals = np.zeros((n, 27))
    valid = np.zeros((n, 27), dtype=bool)
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                nonneg = dx >= 0 and dy >= 0 and dz >= 0
                nonpos = dx <= 0 and dy <= 0 and dz <= 0
                if not (nonneg or nonpos):
                    continue
                d = (dx + 1) + 3 * (dy + 1) + 9 * (dz + 1)
                xx, yy, zz = x + dx, y + dy, z + dz
                ok = (xx >= 0) & (yy >= 0) & (zz >= 0) & (xx < nv) & (yy < nv) & (zz < nv)
                v = np.zeros(n)
                for c in range(8):
                    v = np.where(cell_ok[c], v + coef[c, d], v)
                cols[:, d] = xx + nv * (yy + nv * zz)
                vals[:, d] = v
                valid[:, d] = ok
    a_len = valid.sum(axis=1)
    a_rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(a_len, out=a_rowptr[1:])
    a_col = cols[valid].astype(np.int32)
    a_val = vals[valid]
    b_f = np.zeros(n)
    for c in range(8):
        b_f = np.where(cell_ok[c], b_f + load8[c], b_f)
    return x, y, z, a_rowptr, a_col, a_val, b_f


def cube_operators(n_bg_cells: int, sigma: float = 1.0, row_begin: int = 0, row_end: int | None = None):
    """Host generator.  Returns dict with A=(rowptr,colind,val), M=(rowptr,colind,val), b_f, n_f, n_b for
    foreground rows [row_begin, row_end) (global column ids)."""
    sz = cube_sizes(n_bg_cells)
    nv, nb, n_f, n_b = sz["nv"], sz["nb"], sz["n_f"], sz["n_b"]
    if row_end is None:
        row_end = n_f
    coef, load8 = element_tables(n_bg_cells, sigma)
    x, y, z, a_rowptr, a_col, a_val, b_f = _kuhn_grid_rows(nv, coef, load8, row_begin, row_end)
    n = x.size
    # M: tensor product of 1D hats
    bx0, by0, bz0 = x >> 1, y >> 1, z >> 1
    nx, ny, nz = (x & 1) + 1, (y & 1) + 1, (z & 1) + 1
    wx = np.where(x & 1, 0.5, 1.0)
    wy = np.where(y & 1, 0.5, 1.0)
    wz = np.where(z & 1, 0.5, 1.0)
    mcols = np.zeros((n, 8), dtype=np.int64)
    mvalid = np.zeros((n, 8), dtype=bool)
    s = 0
    for kz in range(2):
        for ky in range(2):
            for kx in range(2):
                mcols[:, s] = (bx0 + kx) + nb * ((by0 + ky) + nb * (bz0 + kz))
                mvalid[:, s] = (kx < nx) & (ky < ny) & (kz < nz)
                s += 1
    m_len = mvalid.sum(axis=1)
    m_rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(m_len, out=m_rowptr[1:])
    m_col = mcols[mvalid].astype(np.int32)
    m_val = np.repeat(wx * wy * wz, m_len)
    return {
        "A": (a_rowptr, a_col, a_val),
        "M": (m_rowptr, m_col, m_val),
        "b_f": b_f,
        "n_f": n_f,
        "n_b": n_b,
        "n_rows": n,
    }


def cube_nnz(n_bg_cells: int):
    """Closed-form sizes: nnz(A_f), nnz(M), nnz(A_b)."""
    nv = 2 * n_bg_cells + 1
    # per direction: pairs (x, x+dx) inside the grid: dx=0 -> nv, dx=+-1 -> nv-1 each
    s0, s1 = nv, nv - 1
    # offsets with all components >= 0 (8 of them, incl. 0) plus all <= 0 (8) minus the double-counted 0
    pos = (s0 + s1) ** 3  # sum over dx,dy,dz in {0,1} of prod
    nnz_a = 2 * pos - s0 ** 3
    nnz_m = (3 * n_bg_cells + 1) ** 3
    nnz_ab = (3 * n_bg_cells + 1) ** 3
    return nnz_a, nnz_m, nnz_ab


# --------------------------------------------------------------------------------------------------
# S2 "unfitted" stress case: the foreground cube of generateUnfittedMesh(dim=3)
# (reference common.py:80-90, demos/poisson_unfitted.py:115-134) inside a rotated background grid
# --------------------------------------------------------------------------------------------------
def unfitted_sizes(n_fg_cells: int, degree: int = 1):
    """n_f, n_b of the S2 case: foreground L_f = 2 with N_f^3 cells, background L_b = 4 with h_b = 2 h_f
    (=> N_b = N_f cells per edge), uniform B-splines of the given degree (N_b + degree functions per edge)."""
    nv = n_fg_cells + 1
    nbx = n_fg_cells + degree
    return {"n_f": nv ** 3, "n_b": nbx ** 3, "nv": nv, "nb": nbx}


def unfitted_operators(n_fg_cells: int, degree: int = 1, sigma: float = 1.0, angle: float = np.pi / 6.0):
    """Host generator of the S2 operands.

      foreground : cube [-1, 1]^3, N_f^3 cells of 6 Kuhn tetrahedra, P1, A_f = K + sigma * Mass (15-pt stencil)
      background : cube [-2, 2]^3 rotated by ``angle`` about z and then about y (``mesh_b.rotate(angle, 2)``,
                   ``mesh_b.rotate(angle, 1)``: reference common.py:88-90), N_b = N_f cells per edge (h_b = 2 h_f),
                   uniform tensor-product B-splines of degree 1 or 2
      M[j, k]    : background function k evaluated at foreground vertex j  ((degree+1)^3 entries per row, rows
                   sum to 1; exact zeros are not stored).  Background functions whose support holds no
                   foreground vertex are empty columns => empty rows of A_b, as with the reference's XTK data.
    """
    if degree not in (1, 2):
        raise ValueError("degree must be 1 or 2")
    sz = unfitted_sizes(n_fg_cells, degree)
    nv, nbx, n_f, n_b = sz["nv"], sz["nb"], sz["n_f"], sz["n_b"]
    L_f, L_b = 2.0, 4.0
    h_f = L_f / n_fg_cells
    h_b = 2.0 * h_f
    coef, load8 = element_tables(n_fg_cells, sigma, h=h_f)
    x, y, z, a_rowptr, a_col, a_val, b_f = _kuhn_grid_rows(nv, coef, load8, 0, n_f)
    # foreground vertex coordinates in the background frame: xi = R^T X, R = R_y(angle) R_z(angle)
    X = np.stack([-L_f / 2 + h_f * x, -L_f / 2 + h_f * y, -L_f / 2 + h_f * z], axis=1)
    c, s = np.cos(angle), np.sin(angle)
    Rz = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    Ry = np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])
    R = Ry @ Rz
    t = ((X @ R) + L_b / 2) / h_b  # rows of X @ R = (R^T X_j)^T ; in units of background cells
    cell = np.floor(t).astype(np.int64)
    u = t - cell
    n_cells_b = n_fg_cells
    if cell.min() < 0 or cell.max() >= n_cells_b:
        raise ValueError("foreground leaves the background grid")
    if degree == 1:
        w1d = np.stack([1.0 - u, u], axis=2)  # (n, 3, 2): functions cell, cell+1
    else:
        w1d = np.stack([0.5 * (1.0 - u) ** 2, 0.5 * (-2.0 * u * u + 2.0 * u + 1.0), 0.5 * u * u], axis=2)
    q = degree + 1
    n = n_f
    mcols = np.zeros((n, q ** 3), dtype=np.int64)
    mvals = np.zeros((n, q ** 3))
    sidx = 0
    for kz in range(q):  # ascending background id within a row: x fastest
        for ky in range(q):
            for kx in range(q):
                mcols[:, sidx] = (cell[:, 0] + kx) + nbx * ((cell[:, 1] + ky) + nbx * (cell[:, 2] + kz))
                mvals[:, sidx] = w1d[:, 0, kx] * w1d[:, 1, ky] * w1d[:, 2, kz]
                sidx += 1
    keep = mvals != 0.0
    m_len = keep.sum(axis=1)
    m_rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(m_len, out=m_rowptr[1:])
    return {
        "A": (a_rowptr, a_col, a_val),
        "M": (m_rowptr, mcols[keep].astype(np.int32), mvals[keep]),
        "b_f": b_f,
        "n_f": n_f,
        "n_b": n_b,
        "n_rows": n,
    }
