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
