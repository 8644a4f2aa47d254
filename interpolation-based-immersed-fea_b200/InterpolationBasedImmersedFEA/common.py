"""Drop-in mirror of the hot-path functions of the reference's ``InterpolationBasedImmersedFEA.common``
(reference InterpolationBasedImmersedFEA/common.py): ``assembleLinearSystemBackground`` (:142-163),
``transferToForeground`` (:123-140), ``zeroDofBackground`` (:120-121), ``solveKSP`` (:509-641) and the
extraction-operator import ``readExOp`` (:645-712, host side).  FEniCS assembly stays on the host
exactly as in the reference; everything PETSc did on this path runs in libiife.so on the GPU.
"""
from __future__ import annotations

import numpy as np

import iife_b200 as _iife
from .la_utils import *  # noqa: F401,F403  (the reference star-imports la_utils too, common.py:8)
from .la_utils import (HAVE_DOLFIN, HAVE_PETSC, CSRMat, Vec, _as_device, _vec_array, arg2m, arg2v, AT_R_A, AT_x,
                       updateU)

DEFAULT_LINEAR_SOLVER = 'gmres'  # reference common.py:36

_KRYLOV = {'gmres': _iife.KSP_FGMRES, 'cg': _iife.KSP_CG}
_DELEGATED_METHODS = ('mumps', 'gcr')          # direct / GCR: not on the north-star path
_DELEGATED_PCS = ('ASM', 'ICC', 'ILU', 'ILUT')  # heavy preconditioners: PETSc only

last_ksp_info = None  # KSPInfo of the most recent solveKSP call (iterations, reason, residual history)


def zeroDofBackground(M):
    """reference common.py:120-121"""
    return arg2m(M).createVecRight()


def transferToForeground(u_f, u_b, M):
    """u_f = M u_b, then ghost update (reference common.py:123-140)."""
    M_m = arg2m(M)
    y = _as_device(M_m).spmv(_vec_array(arg2v(u_b)))
    target = u_f.vector() if (HAVE_DOLFIN and hasattr(u_f, "vector")) else u_f
    _vec_array(arg2v(target))[:] = y
    updateU(u_f)


def assembleLinearSystemBackground(a_f, L_f, M):
    """Background system from foreground forms (reference common.py:142-163).

    With dolfin, ``a_f`` / ``L_f`` are UFL forms and are assembled on the host as in the reference
    (:158-159).  Without dolfin the caller passes the already assembled foreground matrix and vector
    (any type ``arg2m`` / ``arg2v`` accepts)."""
    if HAVE_DOLFIN and not isinstance(a_f, (CSRMat,)) and hasattr(a_f, "arguments"):
        import dolfin

        A_f = m2p(dolfin.assemble(a_f))  # noqa: F405
        b_f = dolfin.assemble(L_f)
    else:
        A_f, b_f = a_f, L_f
    A_b = AT_R_A(M, A_f)
    b_b = AT_x(M, b_f)
    return A_b, b_b


def solveKSP(A, b, u, method='gmres', PC='jacobi',
             remove_zero_diagonal=False, rtol=1E-8,
             atol=1E-9, max_it=1000000, bfr_tol=1E-9,
             monitor=True, gmr_res=3000, bfr_b=True):
    """solve linear system A*u=b (reference common.py:509-641).

    Krylov branch with Jacobi (the reference's default and the north-star path): 'gmres' is PETSc's
    FGMRES with restart 300 (:557, :574), 'cg' is KSPCG (:561), tolerances as given (:555, :631-633),
    the initial guess is whatever ``u`` holds (:634), non-convergence never raises (:635); ``u`` is
    updated in place and None is returned.  Direct solves and heavy preconditioners are PETSc-only in
    the reference and are not reimplemented: they raise NotImplementedError here."""
    global last_ksp_info
    if method is None:
        method = 'gmres'
    if PC is None:
        PC = 'jacobi'
    if method in _DELEGATED_METHODS or PC in _DELEGATED_PCS:
        raise NotImplementedError(
            f"solveKSP(method={method!r}, PC={PC!r}) is PETSc-only in the reference (MUMPS / ASM / ICC / HYPRE); "
            "the B200 path implements method in ('gmres', 'cg') with PC='jacobi'")
    if method not in _KRYLOV:
        raise NotImplementedError(f"unknown method {method!r}")
    if PC != 'jacobi':
        raise NotImplementedError(f"unknown PC {PC!r}")
    if remove_zero_diagonal and bfr_tol is not None:
        raise NotImplementedError("trimNodes (basis function removal) is outside the hot path (SURVEY.md §2 C10)")
    dA = _as_device(arg2m(A))
    bv, uv = arg2v(b), arg2v(u)
    x = np.ascontiguousarray(_vec_array(uv), dtype=np.float64).copy()
    info = _iife.ksp_solve(dA, _vec_array(bv), x, _KRYLOV[method], _iife.PC_JACOBI, rtol=rtol, atol=atol,
                           max_it=max_it, restart=300, hist_len=(4096 if monitor else 0))
    _vec_array(uv)[:] = x
    last_ksp_info = info
    if monitor:
        # the reference prints the iteration count and PETSc's (always empty) history (:638-641)
        print('Converged in', info.iterations, 'iterations.')
        print('Convergence history:', [])
    return None


def read_exop_triplets(fileNames):
    """Triplets of the ``ExOp_Cons*.csv`` files: 1-based foreground id, 1-based background id, weight
    (reference common.py:645-665; format: ``mesh_convert.py:135-157``, space separated)."""
    rows, cols, w = [], [], []
    for name in fileNames:
        data = np.loadtxt(name, dtype=np.float64, ndmin=2)
        rows.append(data[:, 0].astype(np.int64))
        cols.append(data[:, 1].astype(np.int64))
        w.append(data[:, 2])
    return np.concatenate(rows), np.concatenate(cols), np.concatenate(w)


def readExOp(fileNames, V=None, mesh=None, l_size=None, nodeFileNames=None, k=1, NFields=1,
             exo_to_dof=None, n_f=None):
    """Extraction operator M from the MORIS/XTK triplet files (reference common.py:645-712).

    Semantics kept from the reference: ids are 1-based (:700, :703); the background size per field is
    the largest background id (:668); field ``f`` occupies the background block ``id + f*m - 1``
    (field-major, :703); repeated triplets OVERWRITE (``setValue`` / INSERT, :707) instead of adding;
    foreground dofs without a map entry (``< 0``) are skipped (:706).  The reference fills a PETSc
    matrix entry by entry in Python; this builds the CSR arrays vectorised.

    ``exo_to_dof`` is the exodus-node -> FEniCS-dof map of each field (what ``convertDOFs*`` return,
    :681-692); with dolfin it is derived from ``V`` / ``mesh`` by the caller, without it the identity
    (exodus numbering) is used — A_b is invariant under a foreground renumbering."""
    r, c, w = read_exop_triplets(fileNames)
    m = int(c.max())
    if exo_to_dof is None:
        n_scalar = int(n_f // NFields) if n_f is not None else int(r.max())
        maps = [np.arange(n_scalar, dtype=np.int64) * NFields + f for f in range(NFields)] if NFields > 1 else \
            [np.arange(n_scalar, dtype=np.int64)]
    else:
        maps = exo_to_dof if isinstance(exo_to_dof, (list, tuple)) else [exo_to_dof]
    if n_f is None:
        n_f = int(max(mp.max() for mp in maps)) + 1
    fr, fc, fw = [], [], []
    for field in range(NFields):
        dof = np.asarray(maps[field])[r - 1]
        keep = dof >= 0
        fr.append(dof[keep])
        fc.append(c[keep] + field * m - 1)
        fw.append(w[keep])
    fr, fc, fw = np.concatenate(fr), np.concatenate(fc), np.concatenate(fw)
    # INSERT semantics: the LAST occurrence of a repeated (row, col) wins
    key = fr * (m * NFields) + fc
    order = np.argsort(key, kind="stable")
    key_s = key[order]
    last = np.ones(key_s.size, dtype=bool)
    last[:-1] = key_s[1:] != key_s[:-1]
    sel = order[last]
    fr, fc, fw = fr[sel], fc[sel], fw[sel]
    rowptr = np.zeros(n_f + 1, dtype=np.int64)
    np.add.at(rowptr, fr + 1, 1)
    np.cumsum(rowptr, out=rowptr)
    return CSRMat((n_f, m * NFields), rowptr.astype(np.int32), fc.astype(np.int32), fw)
