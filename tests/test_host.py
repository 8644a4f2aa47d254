"""CPU tests of the host-side mirror of the reference interface (no GPU compute calls)."""
import os

import numpy as np
import pytest

REF = "/root/reference/meshes"


def test_mirror_exports_reference_names():
    from InterpolationBasedImmersedFEA import common, la_utils

    for name in ("v2p m2p arg2v arg2m zero_petsc_vec zero_petsc_mat updateU A_x_b AT_x AT_R_A").split():
        assert callable(getattr(la_utils, name)), name
    for name in ("assembleLinearSystemBackground transferToForeground zeroDofBackground solveKSP readExOp "
                 "createNonzeroDiagonal removeZeroDiagonal getIdentity trimNodes solveNewtonsLinear "
                 "estimateConditionNumber").split():
        assert callable(getattr(common, name)), name
    # star-import like the reference (common.py:8) re-exports la_utils
    assert common.AT_R_A is la_utils.AT_R_A
    import inspect

    sig = inspect.signature(common.solveKSP)
    assert list(sig.parameters) == ["A", "b", "u", "method", "PC", "remove_zero_diagonal", "rtol", "atol", "max_it",
                                    "bfr_tol", "monitor", "gmr_res", "bfr_b"]
    assert list(inspect.signature(common.trimNodes).parameters) == ["A", "b", "bfr_tol", "target", "zero_vec", "monitor"]
    assert list(inspect.signature(common.solveNewtonsLinear).parameters) == [
        "A", "L", "u_f", "M", "u_p", "maxIters", "relativeTolerance", "monitorNewtonConvergence",
        "moniterLinearConvergence", "linear_method", "linear_preconditioner", "relax_param", "zero_vec"]
    assert list(inspect.signature(common.estimateConditionNumber).parameters) == ["A", "b", "u", "bfr_tol", "rtol", "atol",
                                                                                   "max_it", "PC"]
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d["method"], d["PC"], d["rtol"], d["atol"], d["max_it"], d["gmr_res"]) == ("gmres", "jacobi", 1e-8, 1e-9, 1000000, 3000)


def test_type_errors_match_reference():
    from InterpolationBasedImmersedFEA import la_utils

    with pytest.raises(TypeError, match="is not supported yet"):
        la_utils.arg2v("nope")
    with pytest.raises(TypeError, match="is not supported yet"):
        la_utils.arg2m(3.0)


def test_delegated_branches_raise():
    from InterpolationBasedImmersedFEA import common

    A = common.CSRMat((1, 1), np.array([0, 1]), np.array([0]), np.array([1.0]))
    b, u = common.Vec(np.ones(1)), common.Vec(np.zeros(1))
    for kw in (dict(method="mumps"), dict(PC="ASM"), dict(PC="ILU"), dict(method="gcr", PC="ICC"), dict(method="bicg")):
        with pytest.raises(NotImplementedError):
            common.solveKSP(A, b, u, **kw)


def test_vec_and_csrmat_petsc_surface():
    from InterpolationBasedImmersedFEA import common

    v = common.Vec(np.arange(3.0))
    v += -common.Vec(np.ones(3)) * 0.5  # u_p += -du_p*relax (reference common.py:394,474)
    assert np.allclose(v.array, [-0.5, 0.5, 1.5]) and v.getSize() == 3 and np.isclose(v.norm(), np.linalg.norm(v.array))
    A = common.CSRMat((2, 3), np.array([0, 1, 2]), np.array([0, 2]), np.array([1.0, 2.0]))
    assert A.getSize() == (2, 3) and A.createVecLeft().getSize() == 2 and A.createVecRight().getSize() == 3
    assert common.zero_petsc_vec(4).getSize() == 4


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference meshes not present (GPU box)")
def test_readexop_semantics_on_shipped_operators():
    """1-based ids, n_b = max background id, INSERT-overwrite of duplicates, field-major blocks
    (reference common.py:645-712)."""
    from InterpolationBasedImmersedFEA import common

    M = common.readExOp([f"{REF}/square/Linear/R6/ExOp_Cons.csv"])
    assert M.getSize()[1] == 17368 and M.rowptr[-1] == 22885
    nonempty = np.diff(M.rowptr) > 0
    assert np.count_nonzero(~nonempty[:34271]) == 25262
    sums = np.add.reduceat(M.val, M.rowptr[:-1][nonempty])
    assert np.allclose(sums, 1.0, atol=1e-12)
    # duplicates carry identical weights and overwrite: rows still sum to 1 (a summing loader would give 2)
    f = f"{REF}/hole_in_plate/Quadratic/FG_R0/R1/ExOp_Cons.csv"
    M1 = common.readExOp([f])
    assert M1.rowptr[-1] == 1974 - 16
    M2 = common.readExOp([f], NFields=2)
    m = M1.getSize()[1]
    assert M2.getSize() == (2 * M1.getSize()[0], 2 * m) and M2.rowptr[-1] == 2 * M1.rowptr[-1]
    lens = np.diff(M2.rowptr)
    odd = next(i for i in range(1, len(lens), 2) if lens[i] > 0)   # dof 2k+1 = node k, field 1
    assert M2.colind[M2.rowptr[odd]:M2.rowptr[odd + 1]].min() >= m  # second background block (field-major)
    assert M2.colind[M2.rowptr[odd - 1]:M2.rowptr[odd]].max() < m
    s2 = np.add.reduceat(M2.val, M2.rowptr[:-1][np.diff(M2.rowptr) > 0])
    assert np.allclose(s2, 1.0, atol=1e-12)


def test_vec_is_lazy_about_device_data(monkeypatch):
    """A vector produced on the GPU stays there until ``.array`` is read; from then on the host copy is the only one
    (the mirror hands out a mutable numpy array).  The device tensor is faked: this is host logic."""
    from InterpolationBasedImmersedFEA import la_utils

    class FakeTensor:
        def __init__(self, a):
            self.a, self.downloads = np.asarray(a, dtype=np.float64), 0

        def numel(self):
            return self.a.size

        def cpu(self):
            self.downloads += 1
            return self

        def numpy(self):
            return self.a.copy()

    synced = []
    monkeypatch.setattr(la_utils._iife, "sync", lambda: synced.append(1))
    t = FakeTensor([1.0, 2.0, 3.0])
    v = la_utils.Vec(device=t)
    assert v.getSize() == 3 and len(v) == 3 and v.getOwnershipRange() == (0, 3) and t.downloads == 0
    assert v.device_tensor() is t
    a = v.array
    assert t.downloads == 1 and synced == [1] and np.array_equal(a, [1.0, 2.0, 3.0])
    assert v.device_tensor() is None  # the host array may be modified from here on
    a[0] = 5.0
    assert v.array[0] == 5.0 and t.downloads == 1
    v2 = la_utils.Vec(np.zeros(2))
    assert v2.device_tensor() is None and v2.getSize() == 2
    v2.array = [1, 2]
    assert v2.array.dtype == np.float64 and v2.norm() == pytest.approx(np.sqrt(5.0))
    with pytest.raises(ValueError):
        la_utils.Vec()
