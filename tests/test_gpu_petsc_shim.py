"""GPU test of the petsc4py adaptor of the mirror (reference la_utils.py:28-182 as the demos call it): with a petsc4py
importable (the duck-typed shim of tests/shims: PETSc itself is not installable in this image) the products read
their operands through ``Mat.getValuesCSR()`` / ``Vec.getArray()``, run in libiife.so on the GPU, and come back as
``PETSc.Mat`` (``createAIJ(csr=...)``) / ``PETSc.Vec`` objects that the callers keep using with PETSc's own methods.
Runs in a subprocess so that the shim does not leak into the other tests."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_products_and_solve_through_petsc_objects(iife):
    code = """
        import numpy as np
        from petsc4py import PETSc
        import InterpolationBasedImmersedFEA.common as common_mod
        from InterpolationBasedImmersedFEA.common import *
        from oracle import oracle as O
        from oracle.synthetic_cube import assemble_cube

        assert HAVE_PETSC and HAVE_DOLFIN
        A, M, b = assemble_cube(4)
        pA = PETSc.Mat().createAIJ(size=A.shape, csr=(A.rowptr.astype(np.int32), A.colind, A.val))
        pM = PETSc.Mat().createAIJ(size=M.shape, csr=(M.rowptr.astype(np.int32), M.colind, M.val))
        pb = PETSc.Vec().createWithArray(b.copy())
        # the funnel of every demo (common.py:142-163) with dolfin wrappers around the PETSc objects
        A_b, b_b = assembleLinearSystemBackground(PETScMatrix(pA), PETScVector(pb), pM)
        assert isinstance(A_b, PETSc.Mat) and isinstance(b_b, PETSc.Vec)
        C = O.AT_R_A(M, A)
        rp, ci, v = A_b.getValuesCSR()
        assert np.array_equal(rp, C.rowptr) and np.array_equal(ci, C.colind)
        bound = O.AT_R_A(O.CSR(M.n_rows, M.n_cols, M.rowptr, M.colind, np.abs(M.val)), O.CSR(A.n_rows, A.n_cols, A.rowptr, A.colind, np.abs(A.val)))
        assert np.all(np.abs(v - C.val) <= 1e-12 * bound.val)
        assert np.allclose(b_b.getArray(), O.AT_x(M, b), rtol=1e-13, atol=0)
        # M is unchanged (the reference transposes it in place twice), A_x_b writes into the caller's Vec
        assert np.array_equal(pM.getValuesCSR()[2], M.val) and pM.getSize() == M.shape
        y = pM.createVecLeft()
        x = np.random.default_rng(0).standard_normal(M.n_cols)
        A_x_b(pM, PETSc.Vec().createWithArray(x), y)
        assert np.allclose(y.getArray(), O.spmv(M, x), rtol=1e-13, atol=1e-15)
        # solveKSP: u is a PETSc.Vec created by the caller from the result matrix (demos/poisson.py:206), updated in place
        u = A_b.createVecLeft()
        solveKSP(A_b, b_b, u, method='cg', PC='jacobi', monitor=False)
        ro = O.solve_ksp(C, O.AT_x(M, b), method='cg', rtol=1e-8, atol=1e-9)
        assert np.linalg.norm(u.getArray() - ro.x) <= 1e-8 * np.linalg.norm(ro.x)
        info = common_mod.last_ksp_info  # (the star import copied the None the module started with)
        assert info.reason == ro.reason and abs(info.iterations - ro.iterations) <= 1
        u2 = A_b.createVecLeft()
        solveKSP(A_b, b_b, u2, method='gmres', monitor=False)
        assert np.linalg.norm(u2.getArray() - ro.x) <= 1e-6 * np.linalg.norm(ro.x)
        # the delegated branch on the SAME objects: preonly + LU (MUMPS in the reference) through PETSc
        u3 = A_b.createVecLeft()
        solveKSP(A_b, b_b, u3, method='mumps', monitor=False)
        assert np.linalg.norm(u3.getArray() - ro.x) <= 1e-6 * np.linalg.norm(ro.x)
        # transferToForeground into a dolfin Function backed by a PETSc.Vec (common.py:123-140)
        u_f = Function(None, PETScVector(pM.createVecLeft()))
        transferToForeground(u_f, u, pM)
        assert np.allclose(u_f.vector().vec().getArray(), O.spmv(M, u.getArray()), rtol=1e-12, atol=1e-15)
        # basis-function removal on PETSc matrices (common.py:207-332): zero a row/column pair, repair it
        S = C.to_scipy().tolil()
        S[5, :] = 0; S[:, 5] = 0
        S = S.tocsr(); S.eliminate_zeros(); S.sort_indices()
        pC = PETSc.Mat().createAIJ(size=S.shape, csr=(S.indptr, S.indices, S.data))
        fixed = removeZeroDiagonal(pC)
        assert isinstance(fixed, PETSc.Mat) and fixed.getDiagonal().getArray()[5] == 1.0
        pC2 = PETSc.Mat().createAIJ(size=S.shape, csr=(S.indptr, S.indices, S.data))
        bb = PETSc.Vec().createWithArray(np.ones(S.shape[0]))
        pC2, bb = trimNodes(pC2, b=bb)
        assert pC2.getDiagonal().getArray()[5] == 1.0 and bb.getArray()[5] == 0.0
        I5 = getIdentity((5, 5))
        assert isinstance(I5, PETSc.Mat) and np.array_equal(I5.getValuesCSR()[2], np.ones(5))
        print('petsc shim ok')
    """
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "tests", "shims"), os.path.join(ROOT, "interpolation-based-immersed-fea_b200"), ROOT])
    out = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and "petsc shim ok" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
