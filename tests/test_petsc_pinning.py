"""Pins the oracle against the reference's real arithmetic — PETSc through petsc4py — wherever that is importable
(it is not in this image: every test here is skipped, and DESIGN.md §5 says "parity unpinned" for that reason).
The calls are the reference's own (la_utils.py:165-182 AT_R_A, :143-163 AT_x, common.py:554-574 KSP setup)."""
import numpy as np
import pytest

from conftest import rand_csr

petsc4py = pytest.importorskip("petsc4py")
from petsc4py import PETSc  # noqa: E402


def _aij(C):
    idt = PETSc.IntType
    A = PETSc.Mat().createAIJ(size=(C.n_rows, C.n_cols),
                              csr=(C.rowptr.astype(idt), C.colind.astype(idt), C.val), comm=PETSc.COMM_SELF)
    A.assemble()
    return A


def _cases(oracle):
    from oracle.synthetic_cube import assemble_cube

    rng = np.random.default_rng(0)
    M = oracle.CSR(300, 90, *rand_csr(rng, 300, 90, 3, empty_frac=0.2))
    A = oracle.CSR(300, 300, *rand_csr(rng, 300, 300, 7, empty_frac=0.05))
    yield M, A, rng.standard_normal(300)
    A2, M2, b2 = assemble_cube(5, 0.0)  # sigma = 0: stored exact zeros must stay in the pattern
    yield M2, A2, b2


def test_at_r_a_and_at_x_against_petsc(oracle):
    for M, A, b in _cases(oracle):
        Mp, Ap = _aij(M), _aij(A)
        # reference la_utils.py:178-181, verbatim
        AT = Mp.transpose()
        ATR = AT.matMult(Ap)
        ATT = Mp.transpose()
        ATRA = ATR.matMult(ATT)
        rp, ci, v = ATRA.getValuesCSR()
        C = oracle.AT_R_A(M, A)
        assert np.array_equal(rp, C.rowptr) and np.array_equal(ci, C.colind), "structural pattern differs from PETSc"
        bound = oracle.AT_R_A(oracle.CSR(M.n_rows, M.n_cols, M.rowptr, M.colind, np.abs(M.val)),
                              oracle.CSR(A.n_rows, A.n_cols, A.rowptr, A.colind, np.abs(A.val))).val
        assert np.all(np.abs(v - C.val) <= 1e-12 * bound + 1e-300)
        x = PETSc.Vec().createWithArray(b.copy(), comm=PETSc.COMM_SELF)
        y = Mp.createVecRight()
        Mp.multTranspose(x, y)  # la_utils.py:162
        assert np.allclose(y.getArray(), oracle.AT_x(M, b), rtol=1e-13, atol=1e-300)


@pytest.mark.parametrize("method", ["cg", "gmres"])
def test_ksp_histories_against_petsc(oracle, method):
    from oracle.synthetic_cube import assemble_cube

    A, M, b = assemble_cube(6)
    C = oracle.AT_R_A(M, A)
    bb = oracle.AT_x(M, b)
    ro = oracle.solve_ksp(C, bb, method=method, hist_len=2000)
    ksp = PETSc.KSP().create(PETSc.COMM_SELF)  # common.py:554-574
    ksp.setTolerances(rtol=1e-8, atol=1e-9, max_it=1000000)
    ksp.setType(PETSc.KSP.Type.FGMRES if method == "gmres" else PETSc.KSP.Type.CG)
    ksp.setOperators(_aij(C))
    ksp.getPC().setType("jacobi")
    ksp.setUp()
    ksp.setGMRESRestart(300)
    ksp.setInitialGuessNonzero(True)
    ksp.setConvergenceHistory()
    x = PETSc.Vec().createWithArray(np.zeros(C.n_rows), comm=PETSc.COMM_SELF)
    ksp.solve(PETSc.Vec().createWithArray(bb.copy(), comm=PETSc.COMM_SELF), x)
    assert ksp.getIterationNumber() == ro.iterations and int(ksp.getConvergedReason()) == ro.reason
    h = np.asarray(ksp.getConvergenceHistory())
    k = min(len(h), ro.iterations + 1)
    assert np.allclose(h[:k], ro.history[:k], rtol=1e-6, atol=1e-12 * ro.history[0])
    assert np.linalg.norm(x.getArray() - ro.x) <= 1e-8 * np.linalg.norm(ro.x)
