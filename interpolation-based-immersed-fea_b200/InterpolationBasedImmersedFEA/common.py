"""Drop-in mirror of the reference's ``InterpolationBasedImmersedFEA.common``
(reference InterpolationBasedImmersedFEA/common.py).

On the extraction hot path everything PETSc did runs in libiife.so on the GPU:
``assembleLinearSystemBackground`` (:142-163), ``transferToForeground`` (:123-140), ``zeroDofBackground`` (:120-121),
the Krylov + Jacobi branch of ``solveKSP`` (:554-574, :628-636: FGMRES, CG, GCR), ``readExOp`` (:645-712, host side), the
basis-function-removal helpers ``createNonzeroDiagonal`` / ``removeZeroDiagonal`` / ``getIdentity`` / ``trimNodes``
(:207-332), the Newton drivers ``solveNewtonsLinear`` (:335-402) and ``solveNonlinear`` (:404-480), ``L2Project``
(:172-195) and ``estimateConditionNumber`` (:483-507).  FEniCS assembly stays on the host exactly as in the reference.

Everything else keeps working as an OVERLAY of the reference package:
  * the MUMPS / ASM / ICC / HYPRE branches of ``solveKSP`` (:525-551, :576-616) are configured through
    petsc4py exactly as the reference does whenever petsc4py is importable (NotImplementedError otherwise);
  * with dolfin importable, dolfin's names are re-exported (the reference module star-imports dolfin, and the
    demos rely on it: demos/poisson.py:14-16) and ``worldcomm`` / ``mpirank`` / ``mpisize`` exist;
  * a name this module does not define (``generateUnfittedMesh``, ``mixedScalarSpace``, ``cellMetric``,
    ``convertDOFs*`` ... — dolfin-only helpers off the hot path) is looked up in the reference's own ``common.py``,
    loaded from ``IIFE_REFERENCE_PATH`` (the directory holding the reference's ``InterpolationBasedImmersedFEA/``)
    or from a second ``InterpolationBasedImmersedFEA`` package found on ``sys.path``; inside that module
    ``from InterpolationBasedImmersedFEA.la_utils import *`` resolves to the mirror, so those helpers run on top
    of the GPU products too.
"""
from __future__ import annotations

import math  # noqa: F401  (re-exported like the reference module's own imports, common.py:9-17)
import os
import sys

import numpy as np

import iife_b200 as _iife
from .la_utils import *  # noqa: F401,F403  (the reference star-imports la_utils too, common.py:8)
from .la_utils import (HAVE_DOLFIN, HAVE_PETSC, CSRMat, PETSc, Vec, _as_device, _to_device, _vec_array, _wrap_mat_like,
                       arg2m, arg2v, AT_R_A, AT_x, m2p, mpirank, updateU, zero_petsc_mat)

if HAVE_PETSC:
    import petsc4py  # noqa: F401  (demos call petsc4py.init(): demos/poisson.py:21)
try:
    from mpi4py import MPI as pyMPI  # type: ignore # noqa: F401
except Exception:  # pragma: no cover - depends on the environment
    pyMPI = None

DEFAULT_LINEAR_SOLVER = 'gmres'  # reference common.py:36

_KRYLOV = {'gmres': _iife.KSP_FGMRES, 'cg': _iife.KSP_CG, 'gcr': _iife.KSP_GCR}
_RESTART = {'gmres': 300, 'cg': 0, 'gcr': 30}  # FGMRES: common.py:574; GCR: PETSc's default (setGMRESRestart does not reach it)
_DELEGATED_METHODS = ('mumps',)                 # sparse direct solve: PETSc only
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


def _petsc_vec(x):
    """petsc4py Vec view of a vector argument (the PETSc branches of solveKSP)."""
    v = arg2v(x)
    if isinstance(v, Vec):
        return PETSc.Vec().createWithArray(v.array)
    return v


def _petsc_mat(A):
    A = arg2m(A)
    if isinstance(A, CSRMat):
        idt = np.dtype(PETSc.IntType)
        return PETSc.Mat().createAIJ(size=A.getSize(), csr=(A.rowptr.astype(idt), A.colind.astype(idt), A.val))
    return A


def _solve_with_petsc(A, b, u, method, PC, remove_zero_diagonal, rtol, atol, max_it, bfr_tol, monitor, gmr_res, bfr_b):
    """The branches of the reference's solveKSP that stay on PETSc (reference common.py:525-551 MUMPS, :553-561 GCR,
    :576-616 ASM / ICC / HYPRE euclid / HYPRE pilut): same KSP / PC configuration, same options, same order of calls."""
    ksp = PETSc.KSP().create()
    ksp.setTolerances(rtol=rtol, atol=atol, max_it=max_it)
    if method == 'mumps':
        if remove_zero_diagonal and bfr_tol is not None:
            if bfr_b:
                A, b = trimNodes(A, b=b, bfr_tol=bfr_tol)
            else:
                A, _ = trimNodes(A, bfr_tol=bfr_tol)
        opts = PETSc.Options("mat_mumps_")
        opts["icntl_24"] = 1      # detection of null pivot rows
        opts["cntl_3"] = 1e-12    # tolerance that defines a null pivot
        A = _petsc_mat(A)
        A.assemble()
        ksp.setOperators(A)
        ksp.setType('preonly')
        pc = ksp.getPC()
        pc.setType('lu')
        pc.setFactorSolverType('mumps')
        ksp.setUp()
        _petsc_ksp_solve(ksp, b, u, None)
        return None
    if method == 'gmres':
        ksp.setType(PETSc.KSP.Type.FGMRES)
    elif method == 'gcr':
        ksp.setType(PETSc.KSP.Type.GCR)
    elif method == 'cg':
        ksp.setType(PETSc.KSP.Type.CG)
    if remove_zero_diagonal and bfr_tol is not None:
        A, b = trimNodes(A, b=b, bfr_tol=bfr_tol)
    A = _petsc_mat(A)
    A.assemble()
    ksp.setOperators(A)
    pc = ksp.getPC()
    if PC == 'jacobi':
        pc.setType("jacobi")
        ksp.setUp()
        ksp.setGMRESRestart(300)
    else:
        ksp.setFromOptions()
        if PC == 'ASM':
            pc.setType("asm")
            pc.setASMOverlap(1)
            ksp.setUp()
            localKSP = pc.getASMSubKSP()[0]
            localKSP.setType(PETSc.KSP.Type.FGMRES)
            localKSP.getPC().setType("lu")
        elif PC == 'ICC':
            pc.setType("icc")
            ksp.setUp()
        elif PC == 'ILU':
            pc.setType("hypre")
            pc.setHYPREType("euclid")
            ksp.setUp()
        elif PC == 'ILUT':
            pc.setType("hypre")
            pc.setHYPREType("pilut")
            ksp.setUp()
        else:
            raise NotImplementedError(f"unknown PC {PC!r}")
        ksp.setGMRESRestart(gmr_res)
    _petsc_ksp_solve(ksp, b, u, dict(monitor=monitor, rtol=rtol, atol=atol, max_it=max_it))
    if monitor:
        print('Converged in', ksp.getIterationNumber(), 'iterations.')
        print('Convergence history:', ksp.getConvergenceHistory())
    return None


def _petsc_ksp_solve(ksp, b, u, params):
    """``PETScKrylovSolver(ksp).solve(PETScVector(u), PETScVector(b))`` with the reference's parameters
    (common.py:628-636) when dolfin is there; the plain petsc4py solve with the same settings otherwise."""
    bv, uv = _petsc_vec(b), _petsc_vec(u)
    if HAVE_DOLFIN:
        ksp_d = PETScKrylovSolver(ksp)  # noqa: F405
        if params is not None:
            if params["monitor"]:
                ksp_d.parameters['monitor_convergence'] = True
            ksp_d.parameters['absolute_tolerance'] = params["atol"]
            ksp_d.parameters['relative_tolerance'] = params["rtol"]
            ksp_d.parameters['maximum_iterations'] = params["max_it"]
            ksp_d.parameters['nonzero_initial_guess'] = True
            ksp_d.parameters['error_on_nonconvergence'] = False
        ksp_d.solve(PETScVector(uv), PETScVector(bv))  # noqa: F405
    else:
        if params is not None:
            ksp.setInitialGuessNonzero(True)
        ksp.solve(bv, uv)
    w = arg2v(u)
    if isinstance(w, Vec):  # the solve wrote through a PETSc view of a copy-free numpy array: nothing to do
        w.array[:] = uv.getArray()


def solveKSP(A, b, u, method='gmres', PC='jacobi',
             remove_zero_diagonal=False, rtol=1E-8,
             atol=1E-9, max_it=1000000, bfr_tol=1E-9,
             monitor=True, gmr_res=3000, bfr_b=True):
    """solve linear system A*u=b (reference common.py:509-641).

    Krylov branch with Jacobi (the reference's default and the north-star path): 'gmres' is PETSc's
    FGMRES with restart 300 (:557, :574), 'cg' is KSPCG (:561), tolerances as given (:555, :631-633),
    the initial guess is whatever ``u`` holds (:634), non-convergence never raises (:635); ``u`` is
    updated in place and None is returned.  'gcr' is KSPGCR (:559-560, restart 30).  Direct solves (MUMPS) and the heavy
    preconditioners (ASM / ICC / HYPRE) keep running on PETSc, configured as the reference configures them, whenever petsc4py is importable;
    without PETSc they raise NotImplementedError."""
    global last_ksp_info
    if method is None:
        method = 'gmres'
    if PC is None:
        PC = 'jacobi'
    if method in _DELEGATED_METHODS or PC in _DELEGATED_PCS:
        if HAVE_PETSC:
            return _solve_with_petsc(A, b, u, method, PC, remove_zero_diagonal, rtol, atol, max_it, bfr_tol, monitor,
                                     gmr_res, bfr_b)
        raise NotImplementedError(
            f"solveKSP(method={method!r}, PC={PC!r}) runs on PETSc (MUMPS / ASM / ICC / HYPRE) in the reference "
            "and petsc4py is not importable here; the B200 path implements method in ('gmres', 'cg', 'gcr') with PC='jacobi'")
    if method not in _KRYLOV:
        raise NotImplementedError(f"unknown method {method!r}")
    if PC != 'jacobi':
        raise NotImplementedError(f"unknown PC {PC!r}")
    if remove_zero_diagonal and bfr_tol is not None:
        A, b = trimNodes(A, b=b, bfr_tol=bfr_tol)  # reference common.py:565-566
    dA = _as_device(arg2m(A))
    bv, uv = arg2v(b), arg2v(u)
    kw = dict(rtol=rtol, atol=atol, max_it=max_it, restart=_RESTART[method], hist_len=(4096 if monitor else 0))
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
    d = _vec_array(arg2m(A).getDiagonal())
    return Vec(np.where(np.abs(d) <= bfr_tol, 1.0, 0.0))


def removeZeroDiagonal(A, bfr_tol=1E-9):
    """Ones onto the (near-)zero diagonal entries of A, in place (reference common.py:236-251): ``A += A0``
    with ``A0 = diag(createNonzeroDiagonal(A))`` — the pattern gains the full diagonal."""
    A = arg2m(A)
    vd = createNonzeroDiagonal(A, bfr_tol=bfr_tol)
    if isinstance(A, CSRMat):
        A.addDiagonal(vd)
        return A
    # a petsc4py Mat: A0 = diag(vd); A += A0 as the reference does (:243-249), the sum computed on the device
    return _wrap_mat_like(A, _as_device(A).add_diagonal(_vec_array(vd)))


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
    bvec = None if b is None else arg2v(b)
    tvec = None if target is None else arg2v(target)
    if zero_vec is not None:
        ids = np.asarray(zero_vec, dtype=np.int64)
    else:
        ids = np.flatnonzero(_vec_array(A.getDiagonal()) <= bfr_tol)
    nz_val = 0
    if bvec is not None:
        def on_device(v):
            return isinstance(v, Vec) and v.device_tensor() is not None

        if tvec is None:
            vals = np.zeros(ids.size)
        elif on_device(tvec):  # only the trimmed entries leave the GPU
            import torch

            t = tvec.device_tensor()
            vals = t[torch.from_numpy(ids).to(t.device)].cpu().numpy()
        else:
            vals = _vec_array(tvec)[ids]
        if on_device(bvec):
            import torch

            t = bvec.device_tensor()
            _iife.sync()
            t[torch.from_numpy(ids).to(t.device)] = torch.from_numpy(np.ascontiguousarray(vals)).to(t.device)
            torch.cuda.current_stream(t.device).synchronize()
        else:
            _vec_array(bvec)[ids] = vals
        nz_val = int(np.count_nonzero(vals > 1e-15))
    if zero_vec is None or monitor:  # the reference prints unconditionally on the scan branch (:322-323)
        print("number of nodes trimmed: ", int(ids.size))
        print("number of nonzero residuals set: ", nz_val)
    if isinstance(A, CSRMat):
        A.zeroRows(ids)
    else:  # petsc4py Mat: its own MatZeroRows, as in the reference (:284, :327)
        idt = np.dtype(PETSc.IntType)
        A.zeroRows(ids.astype(idt))
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
                               max_it=1000000, restart=_RESTART[method])
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
    # the reference never enables a nonzero initial guess on this KSP: PETSc zeroes u and builds the Krylov space from b
    x = np.zeros(uarr.shape[0], dtype=np.float64)
    info, R = _iife.ksp_hessenberg(dA, _vec_array(arg2v(b)), x, pc_type=pc, rtol=rtol, atol=atol, max_it=max_it, restart=1000)
    uarr[:] = x
    if R.size == 0:
        return 0.0, 0.0
    sv = np.linalg.svd(R, compute_uv=False)
    return float(sv.max()), float(sv.min())


# --------------------------------------------------------------------------------------------------
# Newton iteration on a nonlinear residual (reference common.py:404-480) — SURVEY.md §8f row N1
# --------------------------------------------------------------------------------------------------
def solveNonlinear(res_f, u_f, M, u_p,
                   maxIters=20,
                   relativeTolerance=1e-4,
                   monitorNewtonConvergence=True,
                   moniterLinearConvergence=False,
                   linear_method=None,
                   linear_preconditioner=None,
                   bfr_tol=None,
                   relax_param=1,
                   absoluteTolerance=1e-6,
                   absoluteToleranceRes=1e-9,
                   du_0_mag=None,
                   zero_IDs=None,
                   estimateCondNum=False,
                   assemble_cb=None):
    """Solve ``res_f = 0`` by Newton's iteration on the background space (reference common.py:404-480; same
    positional signature, same convergence tests, same messages; it ends the run if the loop does not converge).

    Per iteration the reference assembles ``J_f = derivative(res_f, u_f)`` and ``res_f`` on the host (dolfin), extracts
    them (``AT_R_A`` / ``AT_x``), trims, solves, updates ``u_p`` and transfers it to the foreground.  Here only the
    assembly stays on the host.  With the Krylov + Jacobi solvers the background iterate, the correction, the residual
    and both norms stay on the GPU for the whole loop: per iteration the new foreground values go up, ``u_f = M u_p``
    comes back for the next assembly, and two scalars are read.  The sparsity pattern never changes between
    iterations (:432-435), so every ``AT_R_A`` after the first finds its symbolic plan in the cache (pattern
    fingerprints) and runs the numeric phase only — more cheaply still when the callback hands back the SAME
    ``CSRMat`` after ``set_values`` (values-only upload).

    ``assemble_cb`` (extension for callers without dolfin; the reference has no such argument): a callable
    ``assemble_cb(u_f) -> (J_f, R_f)`` returning the assembled foreground Jacobian and residual at the current
    ``u_f``.  Without it ``res_f`` must be a UFL form and dolfin importable, as in the reference."""
    import torch

    method = linear_method or 'gmres'
    pc = linear_preconditioner or 'jacobi'
    on_device = method in _KRYLOV and pc == 'jacobi'
    if not on_device and not HAVE_PETSC:
        raise NotImplementedError(f"solveNonlinear(linear_method={method!r}, linear_preconditioner={pc!r}) needs PETSc")
    if assemble_cb is None:
        if not (HAVE_DOLFIN and hasattr(res_f, "arguments")):
            raise TypeError("solveNonlinear needs a UFL residual (dolfin) or an assemble_cb callable")

        def assemble_cb(u):  # noqa: F811 - the reference's two host assemblies (:432-433, :158-159)
            J_f = derivative(res_f, u)  # noqa: F405
            return m2p(assemble(J_f)), assemble(res_f)  # noqa: F405

    target_f = u_f.vector() if (HAVE_DOLFIN and hasattr(u_f, "vector")) else u_f
    M_m = arg2m(M)
    dM = _as_device(M_m)
    up_host = _vec_array(arg2v(u_p))
    up_d = _to_device(u_p).clone() if on_device else None
    uf_d = None
    converged = False
    initialNorm = initialNormRes = None

    def tsync():
        torch.cuda.current_stream().synchronize()

    for i in range(0, maxIters):
        J_f, R_f = assemble_cb(u_f)
        dR_b, R_b = assembleLinearSystemBackground(J_f, R_f, M)
        if on_device:
            up_now = Vec(device=up_d)  # trimNodes reads the target's entries on the trimmed rows only
        else:
            up_now = u_p
        if bfr_tol is not None:
            dR_b, R_b = trimNodes(dR_b, R_b, bfr_tol=bfr_tol, target=up_now)
        elif zero_IDs is not None:
            dR_b, R_b = trimNodes(dR_b, b=R_b, target=up_now, zero_vec=zero_IDs, monitor=True)
        if on_device:
            up_d = _to_device(up_now)  # trimNodes may have pulled the iterate to the host to read it
        du_p = zeroDofBackground(M)
        if estimateCondNum:
            estimateConditionNumber(dR_b, R_b, du_p)
        if on_device:
            dA = _as_device(arg2m(dR_b))
            Rb_d = _to_device(R_b)
            du_d = torch.zeros(dA.shape[0], dtype=torch.float64, device=Rb_d.device)
            tsync()
            info = _iife.ksp_solve(dA, Rb_d, du_d, _KRYLOV[method], _iife.PC_JACOBI, rtol=1e-8, atol=1e-9,
                                   max_it=1000000, restart=_RESTART[method])
            if moniterLinearConvergence:
                print('Converged in', info.iterations, 'iterations.')
                print('Convergence history:', [])
            currentNorm = float(torch.linalg.vector_norm(du_d))
            currentNormRes = float(torch.linalg.vector_norm(Rb_d))
        else:
            solveKSP(dR_b, R_b, du_p, method=linear_method, PC=linear_preconditioner, monitor=moniterLinearConvergence,
                     bfr_tol=None)
            currentNorm = arg2v(du_p).norm()
            currentNormRes = arg2v(R_b).norm()
        if i == 0:
            initialNorm = currentNorm
            initialNormRes = currentNormRes
        if du_0_mag is not None:
            initialNorm = du_0_mag
        relativeNorm = currentNorm / initialNorm
        relativeNormRes = currentNormRes / initialNormRes
        if monitorNewtonConvergence and mpirank == 0:
            print("Newton solver iteration: " + str(i) + ", Relative norm of du: " + str(relativeNorm)
                  + ", Relative norm of res: " + str(relativeNormRes), flush=True)
        if relativeNorm < relativeTolerance and relativeNormRes < relativeTolerance:
            converged = True
            break
        if i > 1:
            if currentNorm < absoluteTolerance or currentNormRes < absoluteToleranceRes:
                converged = True
                break
        if on_device:
            up_d.add_(du_d, alpha=-float(relax_param))  # u_p += -du_p*relax_param  (:474)
            if uf_d is None:
                uf_d = torch.empty(dM.shape[0], dtype=torch.float64, device=up_d.device)
            tsync()
            dM.spmv(up_d, y=uf_d)  # transferToForeground(u_f, u_p, M)  (:475)
            _iife.sync()
            _vec_array(arg2v(target_f))[:] = uf_d.cpu().numpy()
            updateU(u_f)
        else:
            up = arg2v(u_p)
            up += -du_p * relax_param
            transferToForeground(u_f, u_p, M)
    if on_device:
        up_host[:] = up_d.cpu().numpy()  # the caller's u_p is updated in place, as `u_p += ...` does in the reference
    if not converged:
        print("ERROR: Nonlinear solver failed to converge.")
        raise SystemExit(1)  # the reference calls exit() here (:478-479)
    return


def L2Project(u_p, u_f, expression_f, M, dx_=None, bfr_tol=None):
    """Project an initial condition onto the foreground and background spaces so that ``u_f = M u_p`` (reference
    common.py:172-195): mass matrix and load on the foreground (host, dolfin), extraction, default Krylov solve
    (FGMRES + Jacobi), transfer back.  Without dolfin ``expression_f`` may be the pair ``(A_f, b_f)`` of the already
    assembled foreground mass matrix and load vector."""
    if HAVE_DOLFIN and not isinstance(expression_f, tuple):
        if dx_ is None:
            dx_ = dx  # noqa: F405
        V_f = u_f.function_space()
        u_f_0 = TrialFunction(V_f)  # noqa: F405
        w_f = TestFunction(V_f)  # noqa: F405
        a_f = inner(u_f_0, w_f) * dx_  # noqa: F405
        L_f = inner(expression_f, w_f) * dx_  # noqa: F405
    elif isinstance(expression_f, tuple) and len(expression_f) == 2:
        a_f, L_f = expression_f
    else:
        raise TypeError("L2Project needs a UFL expression (dolfin) or an assembled (A_f, b_f) pair")
    A_b, b_b = assembleLinearSystemBackground(a_f, L_f, M)
    solveKSP(A_b, b_b, u_p, monitor=False, bfr_tol=bfr_tol)
    transferToForeground(u_f, u_p, M)


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


# --------------------------------------------------------------------------------------------------
# overlay: names this module does not define come from the reference's own common.py
# --------------------------------------------------------------------------------------------------
_REFERENCE_MODULE = None
_REFERENCE_TRIED = False


def _reference_common_path():
    """``common.py`` of the reference package: ``$IIFE_REFERENCE_PATH/InterpolationBasedImmersedFEA/common.py``, or
    the first ``InterpolationBasedImmersedFEA/common.py`` on ``sys.path`` that is not this file."""
    here = os.path.realpath(__file__)
    roots = [os.environ["IIFE_REFERENCE_PATH"]] if os.environ.get("IIFE_REFERENCE_PATH") else []
    roots += [p for p in sys.path if p]
    for root in roots:
        cand = os.path.join(root, "InterpolationBasedImmersedFEA", "common.py")
        if os.path.isfile(cand) and os.path.realpath(cand) != here:
            return cand
    return None


def _reference_public_names():
    """Top-level function / class / constant names of the reference's common.py (parsed, not imported)."""
    path = _reference_common_path()
    if path is None:
        return []
    import ast

    try:
        tree = ast.parse(open(path).read())
    except Exception:  # pragma: no cover
        return []
    names = []
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)):
            names.append(node.name)
        elif isinstance(node, ast.Assign):
            names += [t.id for t in node.targets if isinstance(t, ast.Name)]
    return [n for n in names if not n.startswith("_")]


def _load_reference_common():
    global _REFERENCE_MODULE, _REFERENCE_TRIED
    if _REFERENCE_TRIED:
        return _REFERENCE_MODULE
    _REFERENCE_TRIED = True
    path = _reference_common_path()
    if path is not None:
        import importlib.util

        spec = importlib.util.spec_from_file_location("_iife_reference_common", path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        try:
            spec.loader.exec_module(mod)  # its `from InterpolationBasedImmersedFEA.la_utils import *` resolves to the mirror
        except BaseException:
            del sys.modules[spec.name]
            raise
        _REFERENCE_MODULE = mod
    return _REFERENCE_MODULE


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    try:
        mod = _load_reference_common()
    except Exception as exc:
        raise AttributeError(f"{name!r} is not defined by the B200 mirror, and the reference's common.py could not be "
                             f"imported to provide it ({type(exc).__name__}: {exc})") from exc
    if mod is None or not hasattr(mod, name):
        raise AttributeError(f"module {__name__!r} has no attribute {name!r} (set IIFE_REFERENCE_PATH to the directory "
                             "holding the reference's InterpolationBasedImmersedFEA/ to fall through to it)")
    return getattr(mod, name)


# `from InterpolationBasedImmersedFEA.common import *` must deliver dolfin's names, the mirror's and the
# reference-only helpers alike
__all__ = sorted({n for n in globals() if not n.startswith("_")} | set(_reference_public_names()))

