"""Independent element-by-element assembly of the synthetic S1 cube — TEST
INFRASTRUCTURE ONLY.  It assembles the P1 stiffness + mass matrix tetrahedron by tetrahedron (the way
dolfin's ``assemble`` would, reference common.py:158-159) and the trilinear extraction operator by
evaluating hat functions, without sharing code with ``iife_b200.synthetic`` (which adds precomputed
per-cell stencil tables); tests compare the two (pattern identical, values to rounding).
"""
from __future__ import annotations

import itertools

import numpy as np
import scipy.sparse as sp

from .oracle import CSR


def assemble_cube(n_bg_cells: int, sigma: float = 1.0):
    N = n_bg_cells
    nv = 2 * N + 1
    nb = N + 1
    h = 1.0 / (2 * N)
    n_f = nv ** 3
    # all foreground cells
    cz, cy, cx = np.meshgrid(np.arange(nv - 1), np.arange(nv - 1), np.arange(nv - 1), indexing="ij")
    cx, cy, cz = cx.ravel(), cy.ravel(), cz.ravel()
    rows, cols, vals = [], [], []
    b = np.zeros(n_f)
    for perm in itertools.permutations(range(3)):
        # Kuhn tetrahedron: walk from (0,0,0) to (1,1,1) adding unit vectors in the order `perm`
        loc = np.zeros((4, 3), dtype=np.int64)
        for s, ax in enumerate(perm):
            loc[s + 1] = loc[s]
            loc[s + 1, ax] += 1
        P = loc * h
        B = (P[1:] - P[0]).T
        vol = abs(np.linalg.det(B)) / 6.0
        G = np.zeros((4, 3))
        G[1:] = np.linalg.inv(B)
        G[0] = -G[1:].sum(axis=0)
        Ke = vol * G @ G.T + sigma * vol / 20.0 * (np.ones((4, 4)) + np.eye(4))
        gid = [(cx + loc[a, 0]) + nv * ((cy + loc[a, 1]) + nv * (cz + loc[a, 2])) for a in range(4)]
        for a in range(4):
            np.add.at(b, gid[a], vol / 4.0)
            for c in range(4):
                rows.append(gid[a])
                cols.append(gid[c])
                vals.append(np.full(gid[a].shape, Ke[a, c]))
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n_f, n_f)).tocsr()
    A.sort_indices()  # duplicates summed, exact zeros (sigma = 0 face/body diagonals) stay stored
    # extraction operator: value of the trilinear hat of background node k at foreground vertex j
    j = np.arange(n_f)
    x, y, z = j % nv, (j // nv) % nv, j // (nv * nv)
    mr, mc, mv = [], [], []
    for kz in range(nb):
        wz = np.maximum(0.0, 1.0 - np.abs(z / 2.0 - kz))
        selz = wz > 0
        for ky in range(nb):
            wy = np.maximum(0.0, 1.0 - np.abs(y / 2.0 - ky))
            sely = selz & (wy > 0)
            if not sely.any():
                continue
            for kx in range(nb):
                wx = np.maximum(0.0, 1.0 - np.abs(x / 2.0 - kx))
                sel = sely & (wx > 0)
                idx = np.nonzero(sel)[0]
                if idx.size == 0:
                    continue
                mr.append(idx)
                mc.append(np.full(idx.shape, kx + nb * (ky + nb * kz)))
                mv.append(wx[idx] * wy[idx] * wz[idx])
    M = sp.coo_matrix((np.concatenate(mv), (np.concatenate(mr), np.concatenate(mc))), shape=(n_f, nb ** 3)).tocsr()
    M.sort_indices()
    return CSR.from_scipy(A), CSR.from_scipy(M), b


def cell_tables(n_bg_cells: int, sigma: float = 1.0):
    """coef[c, d] (8 x 27) and load8[c]: what ONE foreground cell (6 Kuhn tetrahedra, P1, K + sigma*Mass) adds to
    the pair (vertex at local corner c = cx + 2cy + 4cz, vertex at c + d), d = (dx+1) + 3(dy+1) + 9(dz+1), and to
    the load of f = 1 at corner c.  Tetrahedra and corners are visited in a fixed order, so the table — and the
    operators oracle_synth_cube_fill builds from it — agree bit for bit with the bench tooling's generator."""
    h = 1.0 / (2.0 * n_bg_cells)
    coef = np.zeros((8, 27))
    load8 = np.zeros(8)
    for perm in itertools.permutations(range(3)):
        loc = np.zeros((4, 3), dtype=np.int64)
        for s, ax in enumerate(perm):
            loc[s + 1] = loc[s]
            loc[s + 1, ax] += 1
        P = loc.astype(np.float64) * h
        B = (P[1:] - P[0]).T
        vol = abs(np.linalg.det(B)) / 6.0
        Binv = np.linalg.inv(B)
        G = np.zeros((4, 3))
        G[1:] = Binv
        G[0] = -Binv.sum(axis=0)
        E = vol * (G @ G.T) + sigma * (vol / 20.0 * (np.ones((4, 4)) + np.eye(4)))
        for a in range(4):
            ca = int(loc[a, 0] + 2 * loc[a, 1] + 4 * loc[a, 2])
            load8[ca] += vol / 4.0
            for b in range(4):
                dd = loc[b] - loc[a]
                coef[ca, int((dd[0] + 1) + 3 * (dd[1] + 1) + 9 * (dd[2] + 1))] += E[a, b]
    return coef, load8


def cube_operators_fast(n_bg_cells: int, sigma: float = 1.0):
    """(A_f, M, b_f) of the S1 cube through the threaded C generator of iife_oracle.c: the operands of bench.py's
    CPU arm at the headline size (N_b = 184: 50 M foreground rows in seconds, nothing of the product imported)."""
    import ctypes

    from . import oracle as O

    L = O.lib()
    N = int(n_bg_cells)
    nv, nb = 2 * N + 1, N + 1
    n_f = nv ** 3
    coef, load8 = cell_tables(N, sigma)
    a_rp = np.zeros(n_f + 1, dtype=np.int64)
    m_rp = np.zeros(n_f + 1, dtype=np.int64)
    i64p, i32p, f64p = ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double)
    L.oracle_synth_cube_lengths(ctypes.c_int64(N), a_rp[1:].ctypes.data_as(i64p), m_rp[1:].ctypes.data_as(i64p))
    np.cumsum(a_rp, out=a_rp)
    np.cumsum(m_rp, out=m_rp)
    a_ci = np.empty(int(a_rp[-1]), dtype=np.int32)
    a_v = np.empty(int(a_rp[-1]))
    m_ci = np.empty(int(m_rp[-1]), dtype=np.int32)
    m_v = np.empty(int(m_rp[-1]))
    b_f = np.empty(n_f)
    L.oracle_synth_cube_fill(ctypes.c_int64(N), coef.ctypes.data_as(f64p), load8.ctypes.data_as(f64p),
                             a_rp.ctypes.data_as(i64p), a_ci.ctypes.data_as(i32p), a_v.ctypes.data_as(f64p),
                             m_rp.ctypes.data_as(i64p), m_ci.ctypes.data_as(i32p), m_v.ctypes.data_as(f64p),
                             b_f.ctypes.data_as(f64p))
    return CSR(n_f, n_f, a_rp, a_ci, a_v), CSR(n_f, nb ** 3, m_rp, m_ci, m_v), b_f
