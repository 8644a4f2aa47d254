"""Drop-in mirror of the hot-path functions of the reference's ``InterpolationBasedImmersedFEA.common``
(reference InterpolationBasedImmersedFEA/common.py): ``assembleLinearSystemBackground`` (:142-163),
``transferToForeground`` (:123-140), ``zeroDofBackground`` (:120-121), ``solveKSP`` (:509-641), the
extraction-operator import ``readExOp`` (:645-712, host side) and — SURVEY.md §8f rows N1/N3 — the
basis-function-removal helpers ``createNonzeroDiagonal`` / ``removeZeroDiagonal`` / ``getIdentity`` /
``trimNodes`` (:207-332), the Newton driver for linear systems ``solveNewtonsLinear`` (:335-402) and
``estimateConditionNumber`` (:483-507).  FEniCS assembly stays on the host
exactly as in the reference; everything PETSc did on this path runs in libiife.so on the GPU.
"""
from __future__ import annotations

import numpy as np

import iife_b200 as _iife
from .la_utils import *  # noqa: F401,F403  (the reference star-imports la_utils too, common.py:8)
from .la_utils import (HAVE_DOLFIN, HAVE_PETSC, CSRMat, Vec, _as_device, _to_device, _vec_array, arg2m, arg2v, AT_R_A,
                       AT_x, updateU)

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
        A, b = trimNodes(A, b=b, bfr_tol=bfr_tol)  # reference common.py:565-566
    dA = _as_device(arg2m(A))
    bv, uv = arg2v(b), arg2v(u)
    kw = dict(rtol=rtol, atol=atol, max_it=max_it, restart=300, hist_len=(4096 if monitor else 0))
    uarr = _vec_array(uv)  # u is updated in place (:634-636)
    direct = isinstance(uarr, np.ndarray) and uarr.dtype == np.float64 and uarr.flags.c_contiguous and uarr.ndim == 1
    if isinstance(bv, Vec) and bv.device_tensor() is not None and direct and uarr.size > 0:
        # b was produced on the GPU (AT_x): solve there; only u crosses PCIe (guess in, solution out)
        import torch

        b_d = bv.device_tensor()
        x_d = torch.from_numpy(uarr).to(b_d.device)
        torch.cuda.current_stream(b_d.device).synchronize()
        info = _iife.ksp_solve(dA, b_d, x_d, _KRYLOV[method], _iife.PC_JACOBI, **kw)
        torch.from_numpy(uarr).copy_(x_d)
    elif direct:
        info = _iife.ksp_solve(dA, _vec_array(bv), uarr, _KRYLOV[method], _iife.PC_JACOBI, **kw)
    else:
        x = np.ascontiguousarray(uarr, dtype=np.float64).copy()
        info = _iife.ksp_solve(dA, _vec_array(bv), x, _KRYLOV[method], _iife.PC_JACOBI, **kw)
        uarr[:] = x
    last_ksp_info = info
    if monitor:
        # the reference prints the iteration count and PETSc's (always empty) history (:638-641)
        print('Converged in', info.iterations, 'iterations.')
        print('Convergence history:', [])
    return None


# --------------------------------------------------------------------------------------------------
# basis function removal (reference common.py:207-332) — SURVEY.md §8f row N3
# --------------------------------------------------------------------------------------------------
def createNonzeroDiagonal(A, bfr_tol=1E-9):
    """Vector with 1 where ``|A_ii| <= bfr_tol`` and 0 elsewhere (reference common.py:207-233).  The diagonal
    is extracted on the device; the comparison is one vectorised pass instead of the reference's
    ``getValue``/``setValue`` loop."""
    d = arg2m(A).getDiagonal().array
    return Vec(np.where(np.abs(d) <= bfr_tol, 1.0, 0.0))


def removeZeroDiagonal(A, bfr_tol=1E-9):
    """Ones onto the (near-)zero diagonal entries of A, in place (reference common.py:236-251): ``A += A0``
    with ``A0 = diag(createNonzeroDiagonal(A))`` — the pattern gains the full diagonal."""
    A = arg2m(A)
    vd = createNonzeroDiagonal(A, bfr_tol=bfr_tol)
    A.addDiagonal(vd)
    return A


def getIdentity(size):
    """Identity of the given ``(local, global)`` size (reference common.py:254-258)."""
    (size_l, size_g) = size
    A = zero_petsc_mat(size_g, size_g, row_loc=size_l, col_loc=size_l)  # noqa: F405
    return removeZeroDiagonal(A)


def trimNodes(A, b=None, bfr_tol=1E-9, target=None, zero_vec=None, monitor=False):
    """Rows whose diagonal is ``<= bfr_tol`` (signed, reference common.py:312) — or the rows listed in
    ``zero_vec`` — become unit rows, ``b`` there becomes ``target`` (or 0) (reference common.py:262-332).
    ``A`` and ``b`` are modified in place and returned.  Diagonal scan and row rewrite run on the device."""
    A = arg2m(A)
    bv = None if b is None else _vec_array(arg2v(b))
    tv = None if target is None else _vec_array(arg2v(target))
    if zero_vec is not None:
        ids = np.asarray(zero_vec, dtype=np.int64)
    else:
        ids = np.flatnonzero(A.getDiagonal().array <= bfr_tol)
    nz_val = 0
    if bv is not None:
        vals = np.zeros(ids.size) if tv is None else tv[ids]
        bv[ids] = vals
        nz_val = int(np.count_nonzero(vals > 1e-15))
    if zero_vec is None or monitor:  # the reference prints unconditionally on the scan branch (:322-323)
        print("number of nodes trimmed: ", int(ids.size))
        print("number of nonzero residuals set: ", nz_val)
    A.zeroRows(ids)
    return A, b


# --------------------------------------------------------------------------------------------------
# Newton iteration on an assembled linear system (reference common.py:335-402) — SURVEY.md §8f row N1
# --------------------------------------------------------------------------------------------------
def solveNewtonsLinear(A, L, u_f, M, u_p,
                       maxIters=20,
                       relativeTolerance=1e-7,
                       monitorNewtonConvergence=True,
                       moniterLinearConvergence=False,
                       linear_method=None,
                       linear_preconditioner=None,
                       relax_param=1,
                       zero_vec=None):
    """Newton / iterative-refinement loop on the background system (reference common.py:335-402; the keyword
    spelling ``moniterLinearConvergence`` is the reference's).  A_b, L_b, the iterate, the residual and the
    correction stay on the device for the whole loop; per iteration only two norms come back to the host, and
    ``u_f = M u_p`` is downloaded for the caller as the reference's ``transferToForeground`` does.  Returns
    ``u_p`` (a ``Vec``) on convergence; like the reference it stops the run if the loop does not converge."""
    import torch

    A_b, L_b = assembleLinearSystemBackground(A, L, M)
    u_p = zeroDofBackground(M)
    if zero_vec is not None:
        A_b, L_b = trimNodes(A_b, b=L_b, target=u_p, zero_vec=zero_vec, monitor=False)
    method = linear_method or 'gmres'
    pc = linear_preconditioner or 'jacobi'
    if method in _DELEGATED_METHODS or pc in _DELEGATED_PCS or method not in _KRYLOV or pc != 'jacobi':
        raise NotImplementedError(f"solveNewtonsLinear(linear_method={method!r}, linear_preconditioner={pc!r})")
    dA, dM = _as_device(arg2m(A_b)), _as_device(arg2m(M))
    n_b = dA.shape[0]
    Lb_d = _to_device(L_b)
    dev = Lb_d.device
    up_d = torch.zeros(n_b, dtype=torch.float64, device=dev)
    res_d = torch.empty_like(up_d)
    du_d = torch.empty_like(up_d)
    uf_d = torch.empty(dM.shape[0], dtype=torch.float64, device=dev)
    initialNorm = initialNormRes = None
    def tsync():  # torch's stream and the library's stream are different streams: order them by hand
        torch.cuda.current_stream().synchronize()

    for i in range(0, maxIters):
        res_d.copy_(Lb_d)
        tsync()
        dA.spmv(up_d, y=res_d, alpha=1.0, beta=1.0)  # A_b.multAdd(u_p, L_b, res_b)  (:364)
        _iife.sync()
        currentNormRes = float(torch.linalg.vector_norm(res_d))
        du_d.zero_()
        tsync()
        info = _iife.ksp_solve(dA, res_d, du_d, _KRYLOV[method], _iife.PC_JACOBI, rtol=1e-8, atol=1e-9,
                               max_it=1000000, restart=300)
        currentNorm = float(torch.linalg.vector_norm(du_d))
        if i == 0:
            initialNorm, initialNormRes = currentNorm, currentNormRes
        relativeNorm = currentNorm / initialNorm
        relativeNormRes = currentNormRes / initialNormRes
        if monitorNewtonConvergence:
            print("Newton solver iteration: " + str(i) + ", Relative norm of du: " + str(relativeNorm)
                  + ", Relative norm of res: " + str(relativeNormRes), flush=True)
        if moniterLinearConvergence:
            print('Converged in', info.iterations, 'iterations.')
        if (relativeNorm < relativeTolerance) or (relativeNormRes < relativeTolerance):
            print('converged')
            u_p.array[:] = up_d.cpu().numpy()
            return u_p
        up_d.add_(du_d, alpha=-float(relax_param))  # u_p += -du_p*relax_param  (:395)
        tsync()
        dM.spmv(up_d, y=uf_d)  # transferToForeground(u_f, u_p, M)  (:399)
        _iife.sync()
        target = u_f.vector() if (HAVE_DOLFIN and hasattr(u_f, "vector")) else u_f
        _vec_array(arg2v(target))[:] = uf_d.cpu().numpy()
        updateU(u_f)
    print("ERROR: Nonlinear solver failed to converge.")
    raise SystemExit(1)  # the reference calls exit() here (:401-402)


def estimateConditionNumber(A, b, u, bfr_tol=None, rtol=1E-8, atol=1E-9, max_it=100000, PC=None):
    """Extreme singular values from a GMRES(1000) solve (reference common.py:483-507: ``setComputeSingularValues``
    + ``computeExtremeSingularValues``): the device FGMRES keeps the triangular factor of its Hessenberg matrix,
    whose singular values are the Hessenberg's; ``u`` receives the solution as in the reference.  ``PC=None`` is
    PETSc's "none"; 'jacobi' preconditions from the right (A D^-1) where PETSc's GMRES would use D^-1 A.
    Returns ``(smax, smin)`` like the reference."""
    if bfr_tol is not None:
        A, b = trimNodes(A, b=b, bfr_tol=bfr_tol)
    if PC in (None, 'none'):
        pc = _iife.PC_NONE
    elif PC == 'jacobi':
        pc = _iife.PC_JACOBI
    else:
        raise NotImplementedError(f"estimateConditionNumber(PC={PC!r})")
    dA = _as_device(arg2m(A))
    uarr = _vec_array(arg2v(u))
    x = np.ascontiguousarray(uarr, dtype=np.float64).copy()
    info, R = _iife.ksp_hessenberg(dA, _vec_array(arg2v(b)), x, pc_type=pc, rtol=rtol, atol=atol, max_it=max_it, restart=1000)
    uarr[:] = x
    if R.size == 0:
        return 0.0, 0.0
    sv = np.linalg.svd(R, compute_uv=False)
    return float(sv.max()), float(sv.min())


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
