"""CPU tests of the oracle (the restatement of the reference path) against dense numpy, its
known-answer tests and the committed golden fixtures.  PARITY UNPINNED: no PETSc."""
import numpy as np
import pytest

from conftest import rand_csr


def _csr(O, n_rows, n_cols, t):
    return O.CSR(n_rows, n_cols, t[0], t[1], t[2])


def test_transpose_sorted_and_dense(oracle):
    rng = np.random.default_rng(0)
    A = _csr(oracle, 37, 23, rand_csr(rng, 37, 23, 4, empty_frac=0.2))
    T = oracle.transpose(A)
    assert np.array_equal(T.todense(), A.todense().T)
    for i in range(T.n_rows):
        c = T.colind[T.rowptr[i]:T.rowptr[i + 1]]
        assert np.all(np.diff(c) > 0)


def test_ptap_dense_small(oracle):
    rng = np.random.default_rng(1)
    M = _csr(oracle, 60, 17, rand_csr(rng, 60, 17, 3, empty_frac=0.3))
    A = _csr(oracle, 60, 60, rand_csr(rng, 60, 60, 6, empty_frac=0.1))
    C = oracle.AT_R_A(M, A)
    Cd = M.todense().T @ A.todense() @ M.todense()
    assert np.allclose(C.todense(), Cd, rtol=0, atol=1e-13 * np.abs(Cd).max())
    pat = (M.pattern_dense().T.astype(int) @ A.pattern_dense().astype(int) @ M.pattern_dense().astype(int)) > 0
    assert np.array_equal(C.pattern_dense(), pat)


def test_identity_extraction_is_bit_exact(oracle):
    """M = I  =>  A_b == A_f bit for bit (mirrors getIdentity, reference common.py:254-258)."""
    rng = np.random.default_rng(2)
    A = _csr(oracle, 40, 40, rand_csr(rng, 40, 40, 5))
    n = 40
    I = oracle.CSR(n, n, np.arange(n + 1), np.arange(n), np.ones(n))
    C = oracle.AT_R_A(I, A)
    assert np.array_equal(C.rowptr, A.rowptr) and np.array_equal(C.colind, A.colind) and np.array_equal(C.val, A.val)


def test_structural_zeros_and_cancellation_stay(oracle):
    """[1,-1] . [1,1]^T cancels numerically but stays in the pattern; stored zeros count."""
    M = oracle.CSR(2, 1, [0, 1, 2], [0, 0], [1.0, 1.0])
    A = oracle.CSR(2, 2, [0, 2, 4], [0, 1, 0, 1], [1.0, -1.0, -1.0, 1.0])
    C = oracle.AT_R_A(M, A)
    assert C.nnz == 1 and C.val[0] == 0.0
    A0 = oracle.CSR(2, 2, [0, 2, 4], [0, 1, 0, 1], [0.0, 0.0, 0.0, 0.0])
    C0 = oracle.AT_R_A(M, A0)
    assert C0.nnz == 1


def test_empty_rows_and_cols_propagate(oracle):
    rng = np.random.default_rng(3)
    rp, ci, v = rand_csr(rng, 50, 20, 2, empty_frac=0.5)
    ci = ci.copy()
    ci[ci == 7] = 8  # leave column 7 unsupported (may create duplicates; rebuild cleanly)
    M = oracle.CSR.from_scipy(oracle.CSR(50, 20, rp, np.sort(ci), v).to_scipy().tocoo().tocsr())
    A = _csr(oracle, 50, 50, rand_csr(rng, 50, 50, 5))
    C = oracle.AT_R_A(M, A)
    unsupported = np.setdiff1d(np.arange(20), M.colind)
    for k in unsupported:
        assert C.rowptr[k + 1] == C.rowptr[k]
        assert not np.any(C.colind == k)


def test_symmetry_and_partition_of_unity(oracle):
    from oracle.synthetic_cube import assemble_cube

    A, M, b = assemble_cube(3)
    C = oracle.AT_R_A(M, A)
    D = C.todense()
    assert np.allclose(D, D.T, rtol=0, atol=1e-14 * np.abs(D).max())
    ones_f = np.ones(A.n_rows)
    assert np.allclose(oracle.spmv(M, np.ones(M.n_cols)), 1.0)
    assert np.isclose(np.ones(C.n_rows) @ D @ np.ones(C.n_rows), ones_f @ A.todense() @ ones_f, rtol=1e-12)
    assert np.allclose(oracle.AT_x(M, b), M.todense().T @ b)


def test_synthetic_generators_agree(oracle):
    from iife_b200 import synthetic
    from oracle.synthetic_cube import assemble_cube

    for sigma in (1.0, 0.0):
        A, M, b = assemble_cube(4, sigma)
        g = synthetic.cube_operators(4, sigma)
        assert np.array_equal(A.rowptr, g["A"][0]) and np.array_equal(A.colind, g["A"][1])
        assert np.allclose(A.val, g["A"][2], rtol=0, atol=1e-14 * np.abs(A.val).max())
        assert np.array_equal(M.rowptr, g["M"][0]) and np.array_equal(M.colind, g["M"][1])
        assert np.array_equal(M.val, g["M"][2])
        assert np.allclose(b, g["b_f"], rtol=1e-13, atol=0)
        assert (A.nnz, M.nnz) == synthetic.cube_nnz(4)[:2]
    # sigma = 0: the face/body diagonal couplings of the Kuhn stiffness are stored exact zeros
    A0, _, _ = assemble_cube(3, 0.0)
    assert np.count_nonzero(A0.val == 0.0) > 0


@pytest.mark.parametrize("method", ["cg", "gmres"])
def test_ksp_against_dense_solve(oracle, method):
    from oracle.synthetic_cube import assemble_cube

    A, M, b = assemble_cube(4)
    C = oracle.AT_R_A(M, A)
    bb = oracle.AT_x(M, b)
    r = oracle.solve_ksp(C, bb, method=method, rtol=1e-12, atol=1e-30, hist_len=400)
    xs = np.linalg.solve(C.todense(), bb)
    assert r.reason == 2
    assert np.linalg.norm(r.x - xs) <= 1e-9 * np.linalg.norm(xs)
    h = r.history[: r.iterations + 1]
    assert h[-1] <= 1e-12 * h[0] * 1.0000001 or r.reason == 2


@pytest.mark.parametrize("method", ["cg", "gmres"])
def test_zero_rhs_with_nonzero_guess_iterates(oracle, method):
    """KSPConvergedDefault: with b = 0 the reference norm falls back to the initial residual norm, so the solve
    reduces the guess by rtol instead of stopping at iteration 0 with DIVERGED_DTOL (any residual >= dtol * 0)."""
    from oracle.synthetic_cube import assemble_cube

    A, M, _ = assemble_cube(3)
    C = oracle.AT_R_A(M, A)
    x0 = np.random.default_rng(2).standard_normal(C.n_rows)
    r = oracle.solve_ksp(C, np.zeros(C.n_rows), x0=x0.copy(), method=method, rtol=1e-8, atol=1e-50)
    assert r.reason == 2 and r.iterations > 3
    assert np.linalg.norm(r.x) <= 1e-6 * np.linalg.norm(x0)


def test_gcr_against_dense_solve_and_restart(oracle):
    """KSPGCR restatement (reference common.py:559-560): converges to the dense solution, the residual history is
    the true residual norm and decreases monotonically (GCR minimises it over the current space), a restart shorter
    than the iteration count still converges."""
    from oracle.synthetic_cube import assemble_cube

    A, M, b = assemble_cube(3)
    C = oracle.AT_R_A(M, A)
    bb = oracle.AT_x(M, b)
    x_ref = np.linalg.solve(C.todense(), bb)
    for restart in (30, 5):
        r = oracle.solve_ksp(C, bb, method="gcr", rtol=1e-10, atol=1e-50, restart=restart, hist_len=400)
        assert r.reason == 2 and (restart == 30 or r.iterations > 5)
        assert np.linalg.norm(r.x - x_ref) <= 1e-8 * np.linalg.norm(x_ref)
        h = r.history[: r.iterations + 1]
        assert np.all(np.diff(h) <= 1e-12 * h[0])
        assert abs(h[-1] - np.linalg.norm(bb - oracle.spmv(C, r.x))) <= 1e-8 * h[0]
    # unsymmetric operator: GCR does not need symmetry
    rng = np.random.default_rng(0)
    S = C.to_scipy().tolil()
    S[0, 5] += 0.3
    S[7, 2] -= 0.2
    Cu = oracle.CSR.from_scipy(S.tocsr())
    xr = np.linalg.solve(Cu.todense(), bb)
    r = oracle.solve_ksp(Cu, bb, method="gcr", rtol=1e-10, atol=1e-50)
    assert r.reason == 2 and np.linalg.norm(r.x - xr) <= 1e-8 * np.linalg.norm(xr)


def test_cg_textbook_iteration_by_iteration(oracle):
    """The C CG equals a line-by-line numpy transcription of SURVEY A.6 (preconditioned norm)."""
    rng = np.random.default_rng(5)
    n = 30
    B = rng.standard_normal((n, n))
    S = B @ B.T + n * np.eye(n)
    S[np.abs(S) < 0.5] = 0.0
    import scipy.sparse as sp

    A = oracle.CSR.from_scipy(sp.csr_matrix(S))
    b = rng.standard_normal(n)
    r = oracle.solve_ksp(A, b, method="cg", rtol=1e-10, atol=1e-50, hist_len=100)
    d = 1.0 / np.diag(S)
    x = np.zeros(n)
    res = b.copy()
    z = d * res
    hist = [np.linalg.norm(z)]
    ttol = 1e-10 * np.linalg.norm(d * b)
    p = None
    beta_old = None
    while hist[-1] > ttol:
        beta = z @ res
        p = z.copy() if p is None else z + beta / beta_old * p
        w = S @ p
        a = beta / (p @ w)
        x += a * p
        res -= a * w
        z = d * res
        hist.append(np.linalg.norm(z))
        beta_old = beta
    assert r.iterations == len(hist) - 1
    assert np.allclose(r.history[: len(hist)], hist, rtol=1e-9)
    assert np.allclose(r.x, x, rtol=1e-9, atol=1e-14)


def test_jacobi_zero_diagonal_becomes_one(oracle):
    A = oracle.CSR(3, 3, [0, 1, 1, 3], [0, 1, 2], [2.0, 5.0, 4.0])  # row 1 empty, row 2 has a 0-diag? no: (2,1),(2,2)
    d = oracle.jacobi_inverse(A)
    assert d[0] == 0.5 and d[1] == 1.0 and d[2] == 0.25


# ---- independent cross-checks against scipy (Gustavson/SMMP structural SpGEMM, its own CG/GMRES).  scipy is
# not PETSc, so this does not pin the oracle to the reference's binary — it pins it to a second, unrelated
# implementation of the same published algorithms, including on the reference's own shipped operators.
def _scipy_triple(M, A):
    """(structural pattern, values) of M^T A M by scipy.  scipy's numeric pass drops sums that are exactly 0,
    so the STRUCTURAL pattern comes from the product of the all-ones patterns (positive sums never cancel)."""
    import scipy.sparse as sp

    Ms, As = M.to_scipy(), A.to_scipy()
    Mp = sp.csr_matrix((np.ones(Ms.nnz), M.colind, M.rowptr), shape=Ms.shape)
    Ap = sp.csr_matrix((np.ones(As.nnz), A.colind, A.rowptr), shape=As.shape)
    P = (Mp.T.tocsr() @ Ap) @ Mp
    P.sort_indices()
    V = (Ms.T.tocsr() @ As) @ Ms
    scale = ((abs(Ms).T.tocsr() @ abs(As)) @ abs(Ms)).max()  # forward-error scale: terms may cancel (1e13 -> 1e4)
    return P, V, float(scale)


def _assert_matches_scipy(C, P, V, tol):
    import scipy.sparse as sp

    assert np.array_equal(C.rowptr, P.indptr) and np.array_equal(C.colind, P.indices)
    D = sp.csr_matrix((C.val, C.colind, C.rowptr), shape=V.shape) - V
    assert D.nnz == 0 or np.abs(D.data).max() <= tol


def _golden_names():
    import pathlib

    return sorted(p.stem for p in (pathlib.Path(__file__).parent / "golden").glob("*.npz"))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_ptap_pattern_and_values_against_scipy_random(oracle, seed):
    rng = np.random.default_rng(100 + seed)
    n_f, n_b = 400 + 50 * seed, 90 + 10 * seed
    M = _csr(oracle, n_f, n_b, rand_csr(rng, n_f, n_b, 3, empty_frac=0.2))
    A = _csr(oracle, n_f, n_f, rand_csr(rng, n_f, n_f, 7, empty_frac=0.05))
    C = oracle.AT_R_A(M, A)
    P, V, scale = _scipy_triple(M, A)
    _assert_matches_scipy(C, P, V, 1e-13 * scale)


@pytest.mark.parametrize("name", _golden_names())
def test_ptap_against_scipy_on_shipped_operators(oracle, name):
    import pathlib

    z = np.load(pathlib.Path(__file__).parent / "golden" / f"{name}.npz")
    n_f, n_b = int(z["n_f"]), int(z["n_b"])
    M = oracle.CSR(n_f, n_b, z["M_rowptr"], z["M_colind"], z["M_val"])
    A = oracle.CSR(n_f, n_f, z["A_rowptr"], z["A_colind"], z["A_val"])
    C = oracle.AT_R_A(M, A)
    P, V, scale = _scipy_triple(M, A)
    _assert_matches_scipy(C, P, V, 1e-13 * scale)
    x = np.linspace(-1.0, 1.0, n_f)
    assert np.allclose(oracle.AT_x(M, x), M.to_scipy().T @ x, rtol=0, atol=1e-12 * np.abs(x).max() * 8)


def test_cg_solution_against_scipy(oracle):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    rng = np.random.default_rng(7)
    n = 300
    B = sp.random(n, n, density=0.02, random_state=7, format="csr")
    S = (B @ B.T + sp.diags(np.full(n, 2.0))).tocsr()
    S.sort_indices()
    A = oracle.CSR.from_scipy(S)
    b = rng.standard_normal(n)
    res = oracle.solve_ksp(A, b, method="cg", PC="jacobi", rtol=1e-12, atol=1e-30, max_it=2000)
    x = res.x
    assert res.reason > 0
    d = S.diagonal()
    xs, info = spla.cg(S, b, rtol=1e-13, atol=0.0, maxiter=5000, M=sp.diags(1.0 / d))
    assert info == 0
    assert np.linalg.norm(x - xs) <= 1e-9 * np.linalg.norm(xs)
    resg = oracle.solve_ksp(A, b, method="gmres", PC="jacobi", rtol=1e-12, atol=1e-30, max_it=2000, restart=40)
    assert resg.reason > 0 and np.linalg.norm(resg.x - xs) <= 1e-8 * np.linalg.norm(xs)


@pytest.mark.parametrize("degree,n_cells", [(1, 12), (2, 10)])
def test_unfitted_stress_case_generator(oracle, degree, n_cells):
    """S2 generator: partition of unity, (p+1)^3 entries per row of M, unsupported background
    functions give empty rows of A_b, total mass/stiffness conserved; oracle == scipy on it."""
    import sys, pathlib

    sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1] / "interpolation-based-immersed-fea_b200"))
    from iife_b200 import synthetic

    g = synthetic.unfitted_operators(n_cells, degree)
    M = oracle.CSR(g["n_f"], g["n_b"], *g["M"])
    A = oracle.CSR(g["n_f"], g["n_f"], *g["A"])
    ml = np.diff(M.rowptr)
    assert ml.max() == (degree + 1) ** 3 and ml.min() >= 1
    assert np.abs(np.add.reduceat(M.val, M.rowptr[:-1]) - 1.0).max() < 1e-14 and M.val.min() > 0.0
    C = oracle.AT_R_A(M, A)
    supported = np.zeros(g["n_b"], dtype=bool)
    supported[M.colind] = True
    assert np.array_equal(np.diff(C.rowptr) > 0, supported)
    assert abs(C.val.sum() - A.val.sum()) < 1e-11 * np.abs(A.val).sum()  # 1^T A_b 1 = 1^T A_f 1 (M 1 = 1)
    assert abs(A.val.sum() - 8.0) < 1e-11  # sigma * volume of [-1,1]^3 (stiffness rows sum to 0)
    P, V, scale = _scipy_triple(M, A)
    _assert_matches_scipy(C, P, V, 1e-13 * scale)


def test_row_edits_follow_petsc_semantics(oracle):
    """MatZeroRows without KEEP_NONZERO_PATTERN and `A += diag` over different patterns (trimNodes /
    removeZeroDiagonal, reference common.py:236-332): dense check + the pattern rules."""
    rng = np.random.default_rng(11)
    A = _csr(oracle, 40, 40, rand_csr(rng, 40, 40, 4, empty_frac=0.25))
    Ad = A.todense()
    rows = np.array([0, 3, 3, 17, 39])
    Z = oracle.zero_rows(A, rows, 1.0)
    Zd = Ad.copy()
    Zd[rows] = 0.0
    Zd[rows, rows] = 1.0
    assert np.array_equal(Z.todense(), Zd)
    lens = np.diff(Z.rowptr)
    assert np.all(lens[rows] == 1) and np.array_equal(np.delete(lens, rows), np.delete(np.diff(A.rowptr), rows))
    Z0 = oracle.zero_rows(A, rows, 0.0)
    assert np.all(np.diff(Z0.rowptr)[rows] == 0)
    d = rng.standard_normal(40)
    D = oracle.add_diagonal(A, d)
    assert np.array_equal(D.todense(), Ad + np.diag(d))
    for i in range(40):
        c = D.colind[D.rowptr[i]:D.rowptr[i + 1]]
        assert np.all(np.diff(c) > 0) and i in c
    assert D.nnz == A.nnz + int(np.sum(~np.diag(A.pattern_dense())))
    # trimNodes on a product with unsupported background functions: empty rows become unit rows
    M = _csr(oracle, 60, 40, rand_csr(rng, 60, 40, 2, empty_frac=0.5))
    S = _csr(oracle, 60, 60, rand_csr(rng, 60, 60, 5))
    C = oracle.AT_R_A(M, oracle.CSR.from_scipy(S.to_scipy() @ S.to_scipy().T))
    b = rng.standard_normal(40)
    C2, b2, ids = oracle.trim_nodes(C, b)
    dg = oracle.diagonal(C)
    assert np.array_equal(ids, np.flatnonzero(dg <= 1e-9)) and ids.size > 0
    assert np.all(oracle.diagonal(C2)[ids] == 1.0) and np.all(b2[ids] == 0.0)
    assert np.array_equal(np.delete(b2, ids), np.delete(b, ids))
    # getIdentity = removeZeroDiagonal(empty matrix) (reference common.py:254-258)
    E = oracle.CSR(5, 5, np.zeros(6, dtype=np.int64), np.zeros(0, dtype=np.int32), np.zeros(0))
    I5 = oracle.add_diagonal(E, oracle.create_nonzero_diagonal(E))
    assert np.array_equal(I5.todense(), np.eye(5))


def test_condition_estimate_from_the_hessenberg(oracle):
    """estimateConditionNumber (reference common.py:483-507): when GMRES runs through the whole Krylov space the
    Hessenberg matrix is orthogonally similar to A, so its extreme singular values are A's — up to the loss of
    orthogonality of classical Gram-Schmidt without refinement (PETSc's default, restated as such: 6e-5 here);
    after fewer steps they lie inside [smin(A), smax(A)]."""
    import scipy.sparse as sp

    rng = np.random.default_rng(5)
    n = 30
    S = sp.random(n, n, density=0.2, random_state=5, format="csr") + sp.diags(np.linspace(1.0, 9.0, n))
    S = S.tocsr()
    S.sort_indices()
    A = oracle.CSR.from_scipy(S)
    sv = np.linalg.svd(S.toarray(), compute_uv=False)
    b = rng.standard_normal(n)
    smax, smin, res = oracle.estimate_condition_number(A, b, rtol=0.0, atol=0.0, max_it=n)
    assert res.iterations == n
    assert abs(smax - sv.max()) <= 1e-3 * sv.max() and abs(smin - sv.min()) <= 1e-3 * sv.max()
    smax2, smin2, res2 = oracle.estimate_condition_number(A, b, rtol=1e-3, atol=0.0)
    assert 0 < res2.iterations < n and res2.reason == 2
    assert sv.min() * (1 - 1e-6) <= smin2 <= smax2 <= sv.max() * (1 + 1e-6)
    # the solve itself is the ordinary one
    x = oracle.solve_ksp(A, b, method="gmres", PC=None, rtol=1e-3, atol=0.0, restart=1000)
    assert x.iterations == res2.iterations and np.array_equal(x.x, res2.x)


def test_oracle_sweep_against_scipy(oracle):
    """hypothesis sweep: structural pattern and values of the oracle's triple product, its transpose and its
    row edits against scipy / dense numpy on irregular shapes (empty rows, heavy rows, stored zeros)."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    @settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(seed=st.integers(0, 2 ** 31 - 1), n_f=st.integers(1, 300), n_b=st.integers(1, 120),
           m_len=st.sampled_from([0.5, 3.0, 12.0]), a_len=st.sampled_from([1.0, 6.0, 30.0]),
           empty=st.sampled_from([0.0, 0.3, 0.8]))
    def sweep(seed, n_f, n_b, m_len, a_len, empty):
        rng = np.random.default_rng(seed)
        M = _csr(oracle, n_f, n_b, rand_csr(rng, n_f, n_b, m_len, empty_frac=empty))
        A = _csr(oracle, n_f, n_f, rand_csr(rng, n_f, n_f, a_len, empty_frac=empty / 2))
        A.val[rng.random(A.nnz) < 0.1] = 0.0  # stored zeros stay in the pattern
        C = oracle.AT_R_A(M, A)
        P, V, scale = _scipy_triple(M, A)
        _assert_matches_scipy(C, P, V, 1e-13 * max(scale, 1e-300))
        T = oracle.transpose(M)
        S = M.to_scipy().T.tocsr()
        S.sort_indices()
        assert np.array_equal(T.rowptr, S.indptr) and np.array_equal(T.colind, S.indices) and np.array_equal(T.val, S.data)
        rows = rng.choice(n_f, size=max(1, n_f // 4))
        Z = oracle.zero_rows(A, rows, 1.0)
        Zd = A.todense() if n_f <= 60 else None
        if Zd is not None:
            Zd[rows] = 0.0
            Zd[rows, rows] = 1.0
            assert np.array_equal(Z.todense(), Zd)
        assert np.all(np.diff(Z.rowptr)[rows] == 1)

    sweep()
