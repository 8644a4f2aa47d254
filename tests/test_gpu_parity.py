"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same inputs.

Bars (BASELINE.json north_star): sparsity pattern of A_b bit-exact; values within 1e-12 relative,
measured as |dC_ij| <= 1e-12 * (|M|^T |A_f| |M|)_ij (per-entry relative error is meaningless on
structurally present, numerically cancelled entries — SURVEY.md §8c); solutions within 1e-8 relative
at matched KSP tolerances; integer/index work bit-exact.
"""
import numpy as np
import pytest

from conftest import rand_csr

pytestmark = pytest.mark.gpu

VAL_TOL = 1e-12
SOL_TOL = 1e-8


def ocsr(O, n_rows, n_cols, t):
    return O.CSR(n_rows, n_cols, t[0], t[1], t[2])


def dmat(I, A):
    return I.DeviceMat.from_csr(A.n_rows, A.n_cols, A.rowptr, A.colind, A.val)


def abs_csr(O, A):
    return O.CSR(A.n_rows, A.n_cols, A.rowptr, A.colind, np.abs(A.val))


def check_ptap(I, O, M, A, plan=None):
    """Runs the CUDA PtAP and the oracle on (M, A); asserts pattern identity and the value bound."""
    dM, dA = dmat(I, M), dmat(I, A)
    own = plan is None
    if own:
        plan = I.PtapPlan(dM, dA)
    dC = plan.numeric(dM, dA, check_errors=True)
    rp, ci, v = dC.to_csr(np.int64)
    C, ATR = O.AT_R_A(M, A, return_intermediate=True)
    assert np.array_equal(rp, C.rowptr), "row pointers of A_b differ from the structural product"
    assert np.array_equal(ci, C.colind.astype(np.int64)), "column indices of A_b differ"
    info = plan.info()
    assert info["nnz_c"] == C.nnz and info["nnz_intermediate"] == ATR.nnz and info["n_b"] == M.n_cols
    bound = O.AT_R_A(abs_csr(O, M), abs_csr(O, A))
    assert np.array_equal(bound.colind, C.colind)
    err = np.abs(v - C.val)
    assert np.all(err <= VAL_TOL * bound.val + 1e-300), f"max scaled error {np.max(err / (bound.val + 1e-300))}"
    nf = np.linalg.norm(C.val)
    assert np.linalg.norm(v - C.val) <= VAL_TOL * nf + 1e-300
    return dM, dA, dC, C, plan


def test_library_is_native(iife):
    assert iife.device_count() >= 1
    assert iife.LIB_PATH.endswith("libiife.so")
    n0 = iife.launch_count(reset=True)
    A = iife.DeviceMat.from_csr(2, 2, np.array([0, 1, 2]), np.array([0, 1]), np.array([1.0, 2.0]))
    y = A.spmv(np.array([1.0, 1.0]))
    assert np.array_equal(y, [1.0, 2.0])
    assert iife.launch_count() > 0


def test_mat_roundtrip_and_validation(iife):
    rng = np.random.default_rng(0)
    for dt in (np.int32, np.int64):
        rp, ci, v = rand_csr(rng, 300, 211, 5, empty_frac=0.3, dtype=dt)
        A = iife.DeviceMat.from_csr(300, 211, rp, ci, v)
        assert A.shape == (300, 211) and A.nnz == len(v)
        for odt in (np.int32, np.int64):
            rp2, ci2, v2 = A.to_csr(odt)
            assert np.array_equal(rp2, rp) and np.array_equal(ci2, ci) and np.array_equal(v2, v)
    # empty matrix, empty rows
    E = iife.DeviceMat.from_csr(4, 3, np.zeros(5, dtype=np.int32), np.zeros(0, dtype=np.int32), np.zeros(0))
    assert E.nnz == 0 and np.array_equal(E.spmv(np.ones(3)), np.zeros(4))
    Z = iife.DeviceMat.from_csr(0, 0, np.zeros(1, dtype=np.int32), np.zeros(0, dtype=np.int32), np.zeros(0))
    assert Z.shape == (0, 0)
    # malformed inputs are refused with IIFE_ERR_ARG
    with pytest.raises(iife.IifeError):
        iife.DeviceMat.from_csr(2, 2, np.array([0, 3, 2]), np.array([0, 1]), np.array([1.0, 2.0]))
    with pytest.raises(iife.IifeError):
        iife.DeviceMat.from_csr(2, 2, np.array([0, 1, 2]), np.array([0, 5]), np.array([1.0, 2.0]))
    with pytest.raises(iife.IifeError):
        iife.DeviceMat.from_csr(1, 3, np.array([0, 2]), np.array([2, 1]), np.array([1.0, 2.0]))


def test_fingerprint_depends_on_pattern_only(iife):
    rng = np.random.default_rng(1)
    rp, ci, v = rand_csr(rng, 100, 100, 6)
    A = iife.DeviceMat.from_csr(100, 100, rp, ci, v)
    B = iife.DeviceMat.from_csr(100, 100, rp, ci, v * 2 + 1)
    assert A.fingerprint() == B.fingerprint()
    ci2 = ci.copy()
    # move one entry to a different (still sorted, unique) column if possible
    for i in range(100):
        row = ci2[rp[i]:rp[i + 1]]
        if len(row) and row[-1] < 99:
            row[-1] += 1
            break
    C = iife.DeviceMat.from_csr(100, 100, rp, ci2, v)
    assert C.fingerprint() != A.fingerprint()
    D = iife.DeviceMat.from_csr(100, 101, rp, ci, v)
    assert D.fingerprint() != A.fingerprint()


@pytest.mark.parametrize("shape", [(500, 37, 3.0), (64, 4000, 40.0), (3000, 5, 2.0), (200, 200, 90.0)])
def test_transpose_matches_oracle(iife, oracle, shape):
    n_rows, n_cols, mean = shape
    rng = np.random.default_rng(2)
    A = ocsr(oracle, n_rows, n_cols, rand_csr(rng, n_rows, n_cols, mean, empty_frac=0.1))
    T = oracle.transpose(A)
    dT = dmat(iife, A).transpose()
    rp, ci, v = dT.to_csr(np.int64)
    assert np.array_equal(rp, T.rowptr) and np.array_equal(ci, T.colind) and np.array_equal(v, T.val)


def test_transpose_very_long_rows(iife, oracle):
    """columns with > 4096 entries take the global-memory sort path (cf. cube/Quadratic/R0: 5672)."""
    rng = np.random.default_rng(3)
    n_rows, n_cols = 12000, 3
    lens = rng.integers(1, 4, n_rows)
    rp = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(lens, out=rp[1:])
    ci = np.concatenate([np.sort(rng.choice(3, l, replace=False)) for l in lens])
    A = oracle.CSR(n_rows, n_cols, rp, ci, rng.standard_normal(rp[-1]))
    T = oracle.transpose(A)
    rp2, ci2, v2 = dmat(iife, A).transpose().to_csr(np.int64)
    assert np.array_equal(rp2, T.rowptr) and np.array_equal(ci2, T.colind) and np.array_equal(v2, T.val)


@pytest.mark.parametrize("mean", [1.5, 4.0, 9.0, 18.0, 60.0])
def test_spmv_matches_oracle(iife, oracle, mean):
    rng = np.random.default_rng(4)
    A = ocsr(oracle, 2000, 1500, rand_csr(rng, 2000, 1500, mean, empty_frac=0.2))
    dA = dmat(iife, A)
    x = rng.standard_normal(1500)
    y = dA.spmv(x)
    yo = oracle.spmv(A, x)
    scale = oracle.spmv(abs_csr(oracle, A), np.abs(x)) + 1e-300
    assert np.all(np.abs(y - yo) <= 1e-14 * scale)
    # transpose product (AT_x) and the alpha/beta form (multAdd)
    xt = rng.standard_normal(2000)
    yt = dA.spmv(xt, trans=True)
    yto = oracle.AT_x(A, xt)
    scale_t = oracle.AT_x(abs_csr(oracle, A), np.abs(xt)) + 1e-300
    assert np.all(np.abs(yt - yto) <= 1e-14 * scale_t)
    y0 = rng.standard_normal(2000)
    y2 = dA.spmv(x, y=y0.copy(), alpha=-2.0, beta=0.5)
    assert np.all(np.abs(y2 - (-2.0 * yo + 0.5 * y0)) <= 1e-13 * (2 * scale + np.abs(y0)))


@pytest.mark.parametrize("n_rows,n_cols,mean,empty", [(1000, 300, 2.0, 0.3), (70001, 9000, 3.4, 0.1), (4097, 4097, 6.0, 0.0)])
def test_spmv_short_rows_stream_kernel(iife, oracle, monkeypatch, n_rows, n_cols, mean, empty):
    """Operators with short ragged rows (the extraction operator M: 1-8 entries) take the CSR-stream kernel, whose
    tiles are staged with bulk async copies (csrc/spmv.cu: k_spmv_stream): tiles that start off 16-byte boundaries,
    empty rows, a ragged last tile, against the oracle and against the plain CSR kernel."""
    rng = np.random.default_rng(int(n_rows))
    A = ocsr(oracle, n_rows, n_cols, rand_csr(rng, n_rows, n_cols, mean, max_len=8, empty_frac=empty))
    x = rng.standard_normal(n_cols)
    ref = oracle.spmv(A, x)
    scale = oracle.spmv(abs_csr(oracle, A), np.abs(x)) + 1e-300
    dA = dmat(iife, A)
    n0 = iife.launch_count(reset=True)
    y = dA.spmv(x)
    assert np.all(np.abs(y - ref) <= 1e-14 * scale)
    monkeypatch.setenv("IIFE_SPMV_STREAM", "0")
    y2 = dmat(iife, A).spmv(x)
    assert np.all(np.abs(y2 - ref) <= 1e-14 * scale)
    assert np.all(np.abs(y2 - y) <= 1e-14 * scale)


def test_diagonal(iife, oracle):
    rng = np.random.default_rng(5)
    A = ocsr(oracle, 400, 400, rand_csr(rng, 400, 400, 7, empty_frac=0.1))
    d = dmat(iife, A).diagonal()
    assert np.array_equal(d, np.diag(A.todense()))


@pytest.mark.parametrize("n_cells,sigma", [(2, 1.0), (5, 1.0), (4, 0.0)])
def test_ptap_cube(iife, oracle, n_cells, sigma):
    from oracle.synthetic_cube import assemble_cube

    A, M, _ = assemble_cube(n_cells, sigma)
    check_ptap(iife, oracle, M, A)


def test_ptap_known_answers(iife, oracle):
    rng = np.random.default_rng(6)
    n = 300
    A = ocsr(oracle, n, n, rand_csr(rng, n, n, 7, empty_frac=0.1))
    I = oracle.CSR(n, n, np.arange(n + 1), np.arange(n), np.ones(n))
    # (i) M = I  =>  A_b == A_f bit for bit
    dI, dA = dmat(iife, I), dmat(iife, A)
    dC, cached = iife.ptap(dI, dA)
    rp, ci, v = dC.to_csr(np.int64)
    assert np.array_equal(rp, A.rowptr) and np.array_equal(ci, A.colind) and np.array_equal(v, A.val)
    # (ii) A_f = I  =>  A_b = M^T M
    M = ocsr(oracle, n, 40, rand_csr(rng, n, 40, 3, empty_frac=0.4))
    check_ptap(iife, oracle, M, I)
    # (v) exact cancellation and stored zeros stay in the pattern
    M2 = oracle.CSR(2, 1, [0, 1, 2], [0, 0], [1.0, 1.0])
    A2 = oracle.CSR(2, 2, [0, 2, 4], [0, 1, 0, 1], [1.0, -1.0, -1.0, 1.0])
    _, _, dC2, _, _ = check_ptap(iife, oracle, M2, A2)
    assert dC2.nnz == 1 and dC2.values()[0] == 0.0


@pytest.mark.parametrize("case", [
    dict(n_f=2000, n_b=700, m_mean=3, a_mean=7, m_empty=0.4, a_empty=0.05),     # shipped-data-like
    dict(n_f=3000, n_b=40, m_mean=8, a_mean=20, m_empty=0.0, a_empty=0.0),      # fat rows: upper levels
    dict(n_f=1500, n_b=1500, m_mean=1, a_mean=2, m_empty=0.5, a_empty=0.5),     # hypersparse, many empties
    dict(n_f=6000, n_b=6, m_mean=2, a_mean=30, m_empty=0.0, a_empty=0.0),       # huge Mt rows: global level
    dict(n_f=900, n_b=300, m_mean=27, a_mean=60, m_empty=0.1, a_empty=0.0),     # quadratic-like widths
])
def test_ptap_random(iife, oracle, case):
    rng = np.random.default_rng(7)
    M = ocsr(oracle, case["n_f"], case["n_b"], rand_csr(rng, case["n_f"], case["n_b"], case["m_mean"], empty_frac=case["m_empty"]))
    A = ocsr(oracle, case["n_f"], case["n_f"], rand_csr(rng, case["n_f"], case["n_f"], case["a_mean"], empty_frac=case["a_empty"]))
    check_ptap(iife, oracle, M, A)


@pytest.mark.parametrize("wide_slots", [True, False])
def test_ptap_every_numeric_kernel_is_exercised(iife, oracle, monkeypatch, wide_slots):
    """Cases built to land rows in each bin of the numeric ladder (slot plans 128/32, 256/256 and the two-byte plan of
    wide rows, warp and CTA hashing levels, global-memory tables) — all against the oracle, and the plan reports the
    bins.  With IIFE_PTAP_SLOTS_WIDE=0 the wide rows go through the hashing kernels they used before."""
    if not wide_slots:
        monkeypatch.setenv("IIFE_PTAP_SLOTS_WIDE", "0")
    rng = np.random.default_rng(11)
    seen = np.zeros(8, dtype=np.int64)

    def one_entry_M(n_f, n_b):
        # every foreground row maps to exactly one background function: Mt rows have ~n_f/n_b entries
        cols = rng.integers(0, n_b, n_f)
        return oracle.CSR(n_f, n_b, np.arange(n_f + 1), cols, rng.random(n_f) + 0.5)

    cases = [
        (ocsr(oracle, 1200, 400, rand_csr(rng, 1200, 400, 3, empty_frac=0.3)), ocsr(oracle, 1200, 1200, rand_csr(rng, 1200, 1200, 6))),   # slots 128/32-ish
        (ocsr(oracle, 800, 200, rand_csr(rng, 800, 200, 4)), ocsr(oracle, 800, 800, rand_csr(rng, 800, 800, 12))),                          # slots 256/256
        (one_entry_M(3000, 100), ocsr(oracle, 3000, 3000, rand_csr(rng, 3000, 3000, 15))),                                                  # n1 ~ 450: wide slots / warp hashing
        (ocsr(oracle, 4000, 600, rand_csr(rng, 4000, 600, 2)), ocsr(oracle, 4000, 4000, rand_csr(rng, 4000, 4000, 50))),                     # CTA hashing
        (ocsr(oracle, 3000, 40, rand_csr(rng, 3000, 40, 8)), ocsr(oracle, 3000, 3000, rand_csr(rng, 3000, 3000, 20))),                       # fat rows
        (ocsr(oracle, 6000, 6, rand_csr(rng, 6000, 6, 2)), ocsr(oracle, 6000, 6000, rand_csr(rng, 6000, 6000, 30))),                         # global tables
        (one_entry_M(6000, 300), ocsr(oracle, 6000, 6000, rand_csr(rng, 6000, 6000, 60))),                                                   # n1 ~ 1100, n2 <= 300: wide slots
    ]
    for M, A in cases:
        _, _, _, _, plan = check_ptap(iife, oracle, M, A)
        seen += np.array(plan.bin_counts())
    assert seen[5] > 0 and seen[6] > 0, seen           # both one-byte slot-plan kernels
    if wide_slots:
        assert seen[7] > 0, seen                        # two-byte slot plan of wide rows
    else:
        assert seen[7] == 0, seen
        assert seen[1] + seen[2] + seen[3] > 0, seen    # shared-memory hashing levels
    assert seen[4] > 0, seen                            # global-memory tables
    print("rows per numeric kernel:", seen.tolist())


@pytest.mark.parametrize("min_rows", [None, "2"])
def test_ptap_template_kernel(iife, oracle, monkeypatch, min_rows):
    """Rows that share structure and M values run one precompiled gather program (csrc/ptap_tpl.cuh): the cube's
    interior rows with the default threshold, every face/edge variant too with a threshold of 2.  Checked against
    the oracle, against the per-row kernels (IIFE_PTAP_TPL=0), after a value update of A_f, and after a value
    update of M (which must rebuild the templates)."""
    from oracle.synthetic_cube import assemble_cube

    if min_rows:
        monkeypatch.setenv("IIFE_TPL_MIN_ROWS", min_rows)
    A, M, _ = assemble_cube(6)
    if not min_rows:  # the library's default: no templates for a problem this small (343 rows)
        monkeypatch.setenv("IIFE_TPL_MIN_PROBLEM", "32768")
        dM0, dA0 = dmat(iife, M), dmat(iife, A)
        p0 = iife.PtapPlan(dM0, dA0)
        p0.numeric(dM0, dA0, check_errors=True)
        assert p0.tpl_info()["rows"] == 0
        monkeypatch.setenv("IIFE_TPL_MIN_PROBLEM", "0")
    dM, dA, dC, C, plan = check_ptap(iife, oracle, M, A)
    ti = plan.tpl_info()
    n_b = M.n_cols
    assert ti["templates"] >= (1 if not min_rows else 8) and ti["rows"] >= 5 ** 3, ti
    if min_rows:
        assert ti["rows"] >= n_b - 8, ti  # everything but the 8 corner rows repeats
    assert ti["lane_use"][0] > 0.5 and ti["lane_use"][1] > 0.5, ti
    v_tpl = dC.values().copy()
    monkeypatch.setenv("IIFE_PTAP_TPL", "0")
    v_row = plan.numeric(dM, dA, check_errors=True).values()
    assert plan.tpl_info()["rows"] == 0
    monkeypatch.delenv("IIFE_PTAP_TPL")
    bound = oracle.AT_R_A(abs_csr(oracle, M), abs_csr(oracle, A))
    assert np.all(np.abs(v_tpl - v_row) <= VAL_TOL * bound.val)
    # new values of A_f on the same templates
    rng = np.random.default_rng(3)
    A2 = oracle.CSR(A.n_rows, A.n_cols, A.rowptr, A.colind, A.val * (1 + 0.3 * rng.standard_normal(A.nnz)))
    dA.update_values(A2.val)
    v2 = plan.numeric(dM, dA, check_errors=True).values()
    assert plan.tpl_info()["rows"] == ti["rows"]
    C2 = oracle.AT_R_A(M, A2)
    b2 = oracle.AT_R_A(abs_csr(oracle, M), abs_csr(oracle, A2))
    assert np.all(np.abs(v2 - C2.val) <= VAL_TOL * b2.val)
    # new values of M: the groups are value dependent, so the template plan is rebuilt (here: nothing repeats)
    M3 = oracle.CSR(M.n_rows, M.n_cols, M.rowptr, M.colind, M.val * (1 + 0.3 * rng.random(M.nnz)))
    dM.update_values(M3.val)
    v3 = plan.numeric(dM, dA, check_errors=True).values()
    assert plan.tpl_info()["rows"] < ti["rows"]
    C3 = oracle.AT_R_A(M3, A2)
    b3 = oracle.AT_R_A(abs_csr(oracle, M3), abs_csr(oracle, A2))
    assert np.all(np.abs(v3 - C3.val) <= VAL_TOL * b3.val)
    # and back
    dM.update_values(M.val)
    v4 = plan.numeric(dM, dA, check_errors=True).values()
    assert plan.tpl_info()["rows"] == ti["rows"] and np.array_equal(v4, v2)


def test_ptap_numeric_reuse_and_cache(iife, oracle):
    """config 4 pattern: same M, same A_f pattern, new values many times on one symbolic plan."""
    from oracle.synthetic_cube import assemble_cube

    rng = np.random.default_rng(8)
    A, M, _ = assemble_cube(4)
    dM, dA, dC, C, plan = check_ptap(iife, oracle, M, A)
    for step in range(3):
        A2 = oracle.CSR(A.n_rows, A.n_cols, A.rowptr, A.colind, A.val * (1 + 0.1 * rng.standard_normal(A.nnz)))
        dA.update_values(A2.val)
        plan.numeric(dM, dA, C=dC, check_errors=True)
        C2 = oracle.AT_R_A(M, A2)
        bound = oracle.AT_R_A(abs_csr(oracle, M), abs_csr(oracle, A2))
        assert np.all(np.abs(dC.values() - C2.val) <= VAL_TOL * bound.val)
    # handle-less AT_R_A entry: second call with a fresh upload of the same pattern hits the plan cache
    iife.plan_cache_clear()
    c1, cached1 = iife.ptap(dmat(iife, M), dmat(iife, A))
    c2, cached2 = iife.ptap(dmat(iife, M), dmat(iife, A2))
    assert not cached1 and cached2
    assert np.all(np.abs(c2.values() - C2.val) <= VAL_TOL * bound.val)
    # a plan refuses operands of another shape
    other = dmat(iife, oracle.CSR(3, 3, [0, 1, 2, 3], [0, 1, 2], [1.0, 1.0, 1.0]))
    with pytest.raises(iife.IifeError):
        plan.numeric(dM, other)


def test_ptap_is_deterministic(iife, oracle):
    from oracle.synthetic_cube import assemble_cube

    A, M, _ = assemble_cube(5)
    dM, dA = dmat(iife, M), dmat(iife, A)
    plan = iife.PtapPlan(dM, dA)
    v1 = plan.numeric(dM, dA).values()
    v2 = plan.numeric(dM, dA).values()
    assert np.array_equal(v1, v2)


@pytest.mark.parametrize("method", ["cg", "gmres"])
def test_ksp_matches_oracle(iife, oracle, method):
    from oracle.synthetic_cube import assemble_cube

    A, M, b = assemble_cube(6)
    C = oracle.AT_R_A(M, A)
    bb = oracle.AT_x(M, b)
    kt = iife.KSP_CG if method == "cg" else iife.KSP_FGMRES
    for rtol, atol in ((1e-8, 1e-9), (1e-12, 1e-50)):
        ro = oracle.solve_ksp(C, bb, method=method, rtol=rtol, atol=atol, hist_len=500)
        x = np.zeros(C.n_rows)
        info = iife.ksp_solve(dmat(iife, C), bb, x, kt, iife.PC_JACOBI, rtol=rtol, atol=atol, hist_len=500)
        assert info.reason == ro.reason, (info.reason_name, ro.reason)
        assert abs(info.iterations - ro.iterations) <= 1
        k = min(info.iterations, ro.iterations) + 1
        assert np.allclose(info.history[:k], ro.history[:k], rtol=1e-4, atol=1e-9 * ro.history[0])
        assert np.linalg.norm(x - ro.x) <= SOL_TOL * np.linalg.norm(ro.x)


@pytest.mark.parametrize("method", ["cg", "gmres"])
def test_ksp_zero_rhs_nonzero_guess(iife, oracle, method):
    """Homogeneous system with a nonzero guess: KSPConvergedDefault falls back to the initial residual norm as the
    reference norm (`if (!snorm) snorm = rnorm`), so the solve iterates to zero instead of stopping at iteration 0
    with DIVERGED_DTOL."""
    from oracle.synthetic_cube import assemble_cube

    A, M, _ = assemble_cube(4)
    C = oracle.AT_R_A(M, A)
    b = np.zeros(C.n_rows)
    x0 = np.random.default_rng(5).standard_normal(C.n_rows)
    ro = oracle.solve_ksp(C, b, x0=x0.copy(), method=method, rtol=1e-8, atol=1e-50, hist_len=400)
    assert ro.reason == 2 and ro.iterations > 3
    x = x0.copy()
    kt = iife.KSP_CG if method == "cg" else iife.KSP_FGMRES
    info = iife.ksp_solve(dmat(iife, C), b, x, kt, iife.PC_JACOBI, rtol=1e-8, atol=1e-50, hist_len=400)
    assert info.reason == ro.reason and abs(info.iterations - ro.iterations) <= 1
    assert np.linalg.norm(x) <= 1e-6 * np.linalg.norm(x0)


def test_cg_repeated_solves_on_device_vectors(iife, oracle):
    """The captured chunk of CG iterations is kept between solves while the kernels' arguments are unchanged (device
    vectors in place, same operator, work arrays from the arena): repeated solves, solves with other tolerances and
    iteration limits (device scalars, not arguments), a value update of the operator in place, another preconditioner
    and another system of the same size must all give what a fresh solve gives (the oracle)."""
    import torch
    from oracle.synthetic_cube import assemble_cube

    A, M, b = assemble_cube(6)
    C = oracle.AT_R_A(M, A)
    bb = oracle.AT_x(M, b)
    dC = dmat(iife, C)
    b_d = torch.from_numpy(bb).cuda()
    x_d = torch.zeros_like(b_d)

    def solve(pc, **kw):
        x_d.zero_()
        info = iife.ksp_solve(dC, b_d, x_d, iife.KSP_CG, pc, **kw)
        torch.cuda.synchronize()
        return info, x_d.cpu().numpy()

    ref = oracle.solve_ksp(C, bb, method="cg", rtol=1e-10, atol=1e-50)
    for _ in range(3):  # same arguments every time
        info, x = solve(iife.PC_JACOBI, rtol=1e-10, atol=1e-50)
        assert info.reason == ref.reason and info.iterations == ref.iterations
        assert np.linalg.norm(x - ref.x) <= 1e-9 * np.linalg.norm(ref.x)
    ref5 = oracle.solve_ksp(C, bb, method="cg", rtol=1e-10, atol=1e-50, max_it=5)
    info, x = solve(iife.PC_JACOBI, rtol=1e-10, atol=1e-50, max_it=5)  # limits live in device scalars
    assert info.reason == ref5.reason and info.iterations == 5
    assert np.linalg.norm(x - ref5.x) <= 1e-10 * np.linalg.norm(ref5.x)
    refn = oracle.solve_ksp(C, bb, method="cg", PC="none", rtol=1e-10, atol=1e-50)
    info, x = solve(iife.PC_NONE, rtol=1e-10, atol=1e-50)  # another preconditioner: another argument list
    assert info.reason == refn.reason and abs(info.iterations - refn.iterations) <= 1
    assert np.linalg.norm(x - refn.x) <= 1e-8 * np.linalg.norm(refn.x)
    C2 = oracle.CSR(C.n_rows, C.n_cols, C.rowptr, C.colind, C.val * 3.0)  # new values in place, same addresses
    dC.update_values(C2.val)
    ref2 = oracle.solve_ksp(C2, bb, method="cg", rtol=1e-10, atol=1e-50)
    info, x = solve(iife.PC_JACOBI, rtol=1e-10, atol=1e-50)
    assert info.reason == ref2.reason and abs(info.iterations - ref2.iterations) <= 1
    assert np.linalg.norm(x - ref2.x) <= 1e-9 * np.linalg.norm(ref2.x)


@pytest.mark.parametrize("restart", [30, 7])
def test_gcr_matches_oracle(iife, oracle, restart):
    """Device GCR (IIFE_KSP_GCR, reference common.py:559-560: method='gcr' -> KSPGCR) against the oracle's restatement:
    reason, iteration count, residual history and solution; also through solveKSP(method='gcr')."""
    from InterpolationBasedImmersedFEA import common as api
    from oracle.synthetic_cube import assemble_cube

    A, M, b = assemble_cube(6)
    C = oracle.AT_R_A(M, A)
    bb = oracle.AT_x(M, b)
    ro = oracle.solve_ksp(C, bb, method="gcr", rtol=1e-8, atol=1e-9, restart=restart, hist_len=600)
    x = np.zeros(C.n_rows)
    info = iife.ksp_solve(dmat(iife, C), bb, x, iife.KSP_GCR, iife.PC_JACOBI, rtol=1e-8, atol=1e-9, restart=restart, hist_len=600)
    assert info.reason == ro.reason and abs(info.iterations - ro.iterations) <= 1, (info.reason_name, info.iterations, ro.iterations)
    k = min(info.iterations, ro.iterations) + 1
    assert np.allclose(info.history[:k], ro.history[:k], rtol=1e-4, atol=1e-9 * ro.history[0])
    assert np.linalg.norm(x - ro.x) <= SOL_TOL * np.linalg.norm(ro.x)
    if restart == 30:
        u = api.Vec(np.zeros(C.n_rows))
        api.solveKSP(api.CSRMat((C.n_rows, C.n_cols), C.rowptr, C.colind, C.val), api.Vec(bb), u, method="gcr", monitor=False)
        assert api.last_ksp_info.reason == ro.reason and np.linalg.norm(u.array - ro.x) <= SOL_TOL * np.linalg.norm(ro.x)
        # max_it is honoured inside a cycle
        x2 = np.zeros(C.n_rows)
        i2 = iife.ksp_solve(dmat(iife, C), bb, x2, iife.KSP_GCR, iife.PC_JACOBI, rtol=1e-30, atol=1e-300, max_it=13)
        assert i2.reason == -3 and i2.iterations == 13


def test_ksp_singular_rows_and_nonzero_guess(iife, oracle):
    """A_b of real data has structurally empty rows (unsupported background functions): Jacobi maps the
    zero diagonal to 1 and those unknowns keep their initial value."""
    rng = np.random.default_rng(9)
    from oracle.synthetic_cube import assemble_cube

    A, M0, b = assemble_cube(4)
    # drop the support of a few background functions: empty columns of M -> empty rows/cols of A_b
    S = M0.to_scipy().tolil()
    dead = [3, 17, 60]
    for k in dead:
        S[:, k] = 0
    S = S.tocsr()
    S.eliminate_zeros()
    M = oracle.CSR.from_scipy(S)
    C = oracle.AT_R_A(M, A)
    bb = oracle.AT_x(M, b)
    x0 = rng.standard_normal(C.n_rows) * 1e-3
    for method, kt in (("cg", iife.KSP_CG), ("gmres", iife.KSP_FGMRES)):
        ro = oracle.solve_ksp(C, bb, x0=x0, method=method, rtol=1e-10, atol=1e-50, hist_len=300)
        x = x0.copy()
        info = iife.ksp_solve(dmat(iife, C), bb, x, kt, iife.PC_JACOBI, rtol=1e-10, atol=1e-50, hist_len=300)
        assert info.reason == ro.reason and abs(info.iterations - ro.iterations) <= 1
        assert np.array_equal(x[dead], x0[dead])
        live = np.setdiff1d(np.arange(C.n_rows), dead)
        assert np.linalg.norm(x[live] - ro.x[live]) <= SOL_TOL * np.linalg.norm(ro.x[live])


def test_ksp_reasons(iife, oracle):
    import scipy.sparse as sp

    n = 50
    # max_it exhaustion never raises (error_on_nonconvergence=False, reference common.py:635)
    rng = np.random.default_rng(10)
    B = rng.standard_normal((n, n))
    S = sp.csr_matrix(B @ B.T + 1e-3 * np.eye(n))
    A = oracle.CSR.from_scipy(S)
    b = rng.standard_normal(n)
    for method, kt in (("cg", iife.KSP_CG), ("gmres", iife.KSP_FGMRES)):
        ro = oracle.solve_ksp(A, b, method=method, rtol=1e-14, atol=1e-50, max_it=7)
        x = np.zeros(n)
        info = iife.ksp_solve(dmat(iife, A), b, x, kt, iife.PC_JACOBI, rtol=1e-14, atol=1e-50, max_it=7)
        assert info.reason == ro.reason == -3 and info.iterations == ro.iterations == 7
        assert np.allclose(x, ro.x, rtol=1e-8, atol=1e-12)
    # indefinite operator: CG reports DIVERGED_INDEFINITE_MAT (-10) or INDEFINITE_PC (-8) like the oracle
    Dm = sp.diags(np.concatenate([np.ones(n // 2), -np.ones(n - n // 2)])).tocsr()
    Ai = oracle.CSR.from_scipy(Dm)
    ro = oracle.solve_ksp(Ai, b, method="cg", PC="none")
    x = np.zeros(n)
    info = iife.ksp_solve(dmat(iife, Ai), b, x, iife.KSP_CG, iife.PC_NONE)
    assert info.reason == ro.reason and info.reason < 0
    # zero right-hand side with zero guess converges immediately
    x = np.zeros(n)
    info = iife.ksp_solve(dmat(iife, A), np.zeros(n), x, iife.KSP_CG, iife.PC_JACOBI)
    assert info.reason > 0 and info.iterations == 0 and np.all(x == 0)


def test_fgmres_restart(iife, oracle):
    """restart shorter than the iteration count exercises the cycle logic (true residual at restart)."""
    from oracle.synthetic_cube import assemble_cube

    A, M, b = assemble_cube(5)
    C = oracle.AT_R_A(M, A)
    bb = oracle.AT_x(M, b)
    ro = oracle.solve_ksp(C, bb, method="gmres", rtol=1e-10, atol=1e-50, restart=5, hist_len=400)
    x = np.zeros(C.n_rows)
    info = iife.ksp_solve(dmat(iife, C), bb, x, iife.KSP_FGMRES, iife.PC_JACOBI, rtol=1e-10, atol=1e-50, restart=5,
                          hist_len=400)
    assert info.reason == ro.reason == 2
    assert abs(info.iterations - ro.iterations) <= 2
    assert np.linalg.norm(x - ro.x) <= SOL_TOL * np.linalg.norm(ro.x)


def test_synthetic_device_generator_is_bit_exact(iife):
    from iife_b200 import synthetic

    for n_cells, sigma in ((3, 1.0), (6, 0.0)):
        g = synthetic.cube_operators(n_cells, sigma)
        import torch

        bf = torch.empty(g["n_f"], dtype=torch.float64, device="cuda:0")
        dA, dM = iife.synth_cube(n_cells, sigma, b_f=bf)
        iife.sync()
        for d, h in ((dA, g["A"]), (dM, g["M"])):
            rp, ci, v = d.to_csr(np.int64)
            assert np.array_equal(rp, h[0]) and np.array_equal(ci, h[1]) and np.array_equal(v, h[2])
        assert np.array_equal(bf.cpu().numpy(), g["b_f"])
    # a row slab carries global column ids
    g = synthetic.cube_operators(4, 1.0, 100, 300)
    dA, dM = iife.synth_cube(4, 1.0, 100, 300)
    rp, ci, v = dA.to_csr(np.int64)
    assert np.array_equal(rp, g["A"][0]) and np.array_equal(ci, g["A"][1]) and np.array_equal(v, g["A"][2])


def test_end_to_end_cube_pipeline(iife, oracle):
    """extraction + solve, device path end to end, against the oracle (N_b = 12: 15 625 fg dofs)."""
    from iife_b200 import synthetic

    g = synthetic.cube_operators(12)
    A = oracle.CSR(g["n_f"], g["n_f"], *g["A"])
    M = oracle.CSR(g["n_f"], g["n_b"], *g["M"])
    dM, dA, dC, C, plan = check_ptap(iife, oracle, M, A)
    bb = dM.spmv(g["b_f"], trans=True)
    bbo = oracle.AT_x(M, g["b_f"])
    assert np.allclose(bb, bbo, rtol=1e-13, atol=0)
    x = np.zeros(C.n_rows)
    info = iife.ksp_solve(dC, bb, x, iife.KSP_CG, iife.PC_JACOBI, rtol=1e-10, atol=1e-50)
    ro = oracle.solve_ksp(C, bbo, method="cg", rtol=1e-10, atol=1e-50)
    assert info.reason == ro.reason == 2 and abs(info.iterations - ro.iterations) <= 1
    assert np.linalg.norm(x - ro.x) <= SOL_TOL * np.linalg.norm(ro.x)
    uf = dM.spmv(x)  # transferToForeground (reference common.py:123-140)
    assert np.allclose(uf, oracle.spmv(M, ro.x), rtol=1e-7, atol=1e-12)


@pytest.mark.parametrize("degree,n_cells", [(1, 20), (2, 14)])
def test_unfitted_stress_case(iife, oracle, degree, n_cells):
    """S2: foreground cube inside a rotated background grid (reference common.py:80-90).
    Wide rows — A_b up to 119 (p=1) / 331 (p=2) entries, intermediate rows beyond the 256-entry slot plan —
    so the hashing kernels run on a realistic pattern, with thousands of unsupported (empty) rows."""
    from iife_b200 import synthetic

    g = synthetic.unfitted_operators(n_cells, degree)
    A = oracle.CSR(g["n_f"], g["n_f"], *g["A"])
    M = oracle.CSR(g["n_f"], g["n_b"], *g["M"])
    dM, dA, dC, C, plan = check_ptap(iife, oracle, M, A)
    lens = np.diff(C.rowptr)
    assert (lens == 0).sum() > 1000 and lens.max() > (100 if degree == 1 else 256)
    bb = dM.spmv(g["b_f"], trans=True)
    bbo = oracle.AT_x(M, g["b_f"])
    assert np.allclose(bb, bbo, rtol=0, atol=1e-13 * np.abs(bbo).max())
    bnorm = np.linalg.norm(bbo)
    for method, kt in (("cg", iife.KSP_CG), ("gmres", iife.KSP_FGMRES)):
        ro = oracle.solve_ksp(C, bbo, method=method, max_it=5000, hist_len=64)
        x = np.zeros(C.n_rows)
        info = iife.ksp_solve(dC, bb, x, kt, iife.PC_JACOBI, max_it=5000, hist_len=64)
        assert info.reason == ro.reason == 2, (method, info.reason_name)
        # cut-cell conditioning (p=2: 800 CG iterations at 2 700 dofs): the count moves with rounding
        assert abs(info.iterations - ro.iterations) <= max(2, ro.iterations // 10), (method, info.iterations)
        k = min(info.iterations, ro.iterations, 10) + 1
        assert np.allclose(info.history[:k], ro.history[:k], rtol=1e-3, atol=1e-9 * ro.history[0])
        r_gpu = np.linalg.norm(bbo - oracle.spmv(C, x))
        r_orc = np.linalg.norm(bbo - oracle.spmv(C, ro.x))
        assert r_gpu <= 10.0 * max(r_orc, 1e-8 * bnorm), (method, r_gpu, r_orc)
        # forward-error parity where it is meaningful: one more or less iteration moves the p=1 CG solution by
        # 4e-8 but the FGMRES one by 1e-6..1e-4 (true-residual test on a cut-cell operator)
        if degree == 1 and (method == "cg" or info.iterations == ro.iterations):
            assert np.linalg.norm(x - ro.x) <= 1e-6 * np.linalg.norm(ro.x)


def test_row_edits_match_oracle(iife, oracle):
    """iife_mat_zero_rows / iife_mat_add_diagonal against the oracle's restatement of MatZeroRows and
    `A += diag` (trimNodes / removeZeroDiagonal, reference common.py:236-332): bit-exact pattern and values."""
    rng = np.random.default_rng(21)
    for n, mean, empty in ((1, 1, 0.0), (37, 3, 0.3), (5000, 9, 0.1), (300, 60, 0.0)):
        A = ocsr(oracle, n, n, rand_csr(rng, n, n, mean, empty_frac=empty))
        dA = dmat(iife, A)
        rows = rng.choice(n, size=max(1, n // 7), replace=True)  # duplicates allowed
        for diag in (1.0, 0.0, -2.5):
            Z = oracle.zero_rows(A, rows, diag)
            rp, ci, v = dA.zero_rows(rows, diag).to_csr(np.int64)
            assert np.array_equal(rp, Z.rowptr) and np.array_equal(ci, Z.colind) and np.array_equal(v, Z.val)
        for dt in (np.int32, np.int64):
            rp, ci, v = dA.zero_rows(rows.astype(dt)).to_csr(np.int64)
            Z = oracle.zero_rows(A, rows, 1.0)
            assert np.array_equal(rp, Z.rowptr) and np.array_equal(ci, Z.colind) and np.array_equal(v, Z.val)
        d = rng.standard_normal(n)
        D = oracle.add_diagonal(A, d)
        rp, ci, v = dA.add_diagonal(d).to_csr(np.int64)
        assert np.array_equal(rp, D.rowptr) and np.array_equal(ci, D.colind) and np.array_equal(v, D.val)
        # the source matrix is untouched
        rp, ci, v = dA.to_csr(np.int64)
        assert np.array_equal(rp, A.rowptr) and np.array_equal(ci, A.colind) and np.array_equal(v, A.val)
    # rectangular: diagonal positions exist only for i < n_cols
    R = ocsr(oracle, 30, 12, rand_csr(rng, 30, 12, 3))
    dR = dmat(iife, R)
    rows = np.array([2, 11, 12, 29])
    Z = oracle.zero_rows(R, rows, 1.0)
    rp, ci, v = dR.zero_rows(rows).to_csr(np.int64)
    assert np.array_equal(rp, Z.rowptr) and np.array_equal(ci, Z.colind) and np.array_equal(v, Z.val)
    d = rng.standard_normal(30)
    D = oracle.add_diagonal(R, d)
    rp, ci, v = dR.add_diagonal(d).to_csr(np.int64)
    assert np.array_equal(rp, D.rowptr) and np.array_equal(ci, D.colind) and np.array_equal(v, D.val)
    # empty list, empty matrix, out-of-range index
    rp, ci, v = dA.zero_rows(np.zeros(0, dtype=np.int64)).to_csr(np.int64)
    assert np.array_equal(rp, A.rowptr) and np.array_equal(v, A.val)
    E = iife.DeviceMat.from_csr(4, 4, np.zeros(5, dtype=np.int32), np.zeros(0, dtype=np.int32), np.zeros(0))
    rp, ci, v = E.add_diagonal(np.ones(4)).to_csr(np.int64)
    assert np.array_equal(rp, np.arange(5)) and np.array_equal(ci, np.arange(4)) and np.array_equal(v, np.ones(4))
    with pytest.raises(iife.IifeError):
        dA.zero_rows(np.array([A.n_rows]))


def test_trim_nodes_and_newton_through_the_mirror(iife, oracle, capsys):
    """trimNodes / removeZeroDiagonal / getIdentity / solveKSP(remove_zero_diagonal=True) / solveNewtonsLinear
    through the reference's own names (reference common.py:207-402), against the oracle."""
    from InterpolationBasedImmersedFEA import common as api
    from iife_b200 import synthetic

    g = synthetic.unfitted_operators(12, 1)  # thousands of unsupported background functions => empty rows of A_b
    n_f, n_b = g["n_f"], g["n_b"]
    Ao, Mo = oracle.CSR(n_f, n_f, *g["A"]), oracle.CSR(n_f, n_b, *g["M"])
    A, M = api.CSRMat((n_f, n_f), *g["A"]), api.CSRMat((n_f, n_b), *g["M"])
    A_b, b_b = api.assembleLinearSystemBackground(A, api.Vec(g["b_f"]), M)
    Co, bbo = oracle.AT_R_A(Mo, Ao), oracle.AT_x(Mo, g["b_f"])
    # createNonzeroDiagonal / removeZeroDiagonal on a copy
    vd = api.createNonzeroDiagonal(A_b)
    assert np.array_equal(vd.array, oracle.create_nonzero_diagonal(Co)) and vd.array.sum() > 1000
    Cc = api.CSRMat((n_b, n_b), A_b.rowptr.copy(), A_b.colind.copy(), A_b.val.copy())
    R = api.removeZeroDiagonal(Cc)
    Ro = oracle.add_diagonal(Co, oracle.create_nonzero_diagonal(Co))
    assert R is Cc and np.array_equal(R.rowptr, Ro.rowptr) and np.array_equal(R.colind, Ro.colind)
    assert np.allclose(R.val, Ro.val, rtol=0, atol=1e-12 * np.abs(Ro.val).max())
    I7 = api.getIdentity((7, 7))
    assert np.array_equal(I7.to_scipy().toarray(), np.eye(7))
    # trimNodes, scan branch (prints like the reference) and list branch
    T_o, bt_o, ids = oracle.trim_nodes(Co, bbo)
    b_t = api.Vec(b_b.array.copy())
    T, b_ret = api.trimNodes(A_b, b=b_t)
    out = capsys.readouterr().out
    assert f"number of nodes trimmed:  {ids.size}" in out and "number of nonzero residuals set:  0" in out
    assert T is A_b and b_ret is b_t
    assert np.array_equal(T.rowptr, T_o.rowptr) and np.array_equal(T.colind, T_o.colind)
    assert np.allclose(T.val, T_o.val, rtol=0, atol=1e-12 * np.abs(T_o.val).max())
    assert np.all(b_t.array[ids] == 0.0)
    tgt = api.Vec(np.linspace(1.0, 2.0, n_b))
    b_l = api.Vec(b_b.array.copy())
    A_l, _ = api.assembleLinearSystemBackground(A, api.Vec(g["b_f"]), M)
    api.trimNodes(A_l, b=b_l, target=tgt, zero_vec=list(ids[:5]))
    assert np.array_equal(b_l.array[ids[:5]], tgt.array[ids[:5]])
    assert np.all(np.diff(A_l.rowptr)[ids[:5]] == 1)
    # the trimmed system is nonsingular: the Krylov solve agrees with the oracle's
    u = api.Vec(np.zeros(n_b))
    api.solveKSP(T, b_t, u, method="cg", PC="jacobi", monitor=False)
    ro = oracle.solve_ksp(T_o, bt_o, method="cg")
    assert api.last_ksp_info.reason == ro.reason == 2 and abs(api.last_ksp_info.iterations - ro.iterations) <= 2
    assert np.linalg.norm(u.array - ro.x) <= 1e-6 * np.linalg.norm(ro.x)
    # solveKSP(remove_zero_diagonal=True) trims first (reference common.py:565-566)
    A2, b2 = api.assembleLinearSystemBackground(A, api.Vec(g["b_f"]), M)
    u2 = api.Vec(np.zeros(n_b))
    api.solveKSP(A2, b2, u2, method="cg", PC="jacobi", remove_zero_diagonal=True, monitor=False)
    assert np.linalg.norm(u2.array - ro.x) <= 1e-6 * np.linalg.norm(ro.x)
    # solveNewtonsLinear: residual convention res = A_b u + L_b, u -= du  =>  converges to -A_b^{-1} L_b
    u_f = api.Vec(np.zeros(n_f))
    u_p = api.solveNewtonsLinear(A, api.Vec(g["b_f"]), u_f, M, None, linear_method="cg", linear_preconditioner="jacobi",
                                 monitorNewtonConvergence=False, zero_vec=list(ids))
    assert np.linalg.norm(u_p.array + ro.x) <= 1e-6 * np.linalg.norm(ro.x)
    assert np.allclose(u_f.array, oracle.spmv(Mo, u_p.array), rtol=0, atol=1e-10 * np.abs(u_f.array).max())


def test_solve_nonlinear_device_loop(iife, oracle, capsys):
    """solveNonlinear (reference common.py:404-480) with the assembly callback: a mildly nonlinear problem
    R(u_f) = A_f u_f + c u_f^3 - f on the cube, Newton on the background space.  The iterate stays on the GPU; every
    iteration hands back the SAME CSRMat after set_values, so the PtAP plan is reused (one symbolic phase in total).
    Checked against the same loop run with the oracle."""
    from InterpolationBasedImmersedFEA import common as api
    from oracle.synthetic_cube import assemble_cube

    A, M, f = assemble_cube(4)
    n_f, n_b = A.n_rows, M.n_cols
    c = 50.0
    diag_pos = np.array([A.rowptr[i] + np.searchsorted(A.colind[A.rowptr[i]:A.rowptr[i + 1]], i) for i in range(n_f)])

    def residual_and_jacobian_values(u):
        R = oracle.spmv(A, u) + c * u ** 3 * np.abs(f) - f
        Jv = A.val.copy()
        Jv[diag_pos] += 3.0 * c * u ** 2 * np.abs(f)
        return R, Jv

    # oracle loop (host)
    u_b = np.zeros(n_b)
    for it in range(12):
        R, Jv = residual_and_jacobian_values(oracle.spmv(M, u_b))
        J = oracle.CSR(n_f, n_f, A.rowptr, A.colind, Jv)
        dR = oracle.AT_R_A(M, J)
        Rb = oracle.AT_x(M, R)
        du = oracle.solve_ksp(dR, Rb, method="cg", rtol=1e-8, atol=1e-9).x
        if np.linalg.norm(du) < 1e-9 * max(np.linalg.norm(u_b), 1e-30):
            break
        u_b = u_b - du
    # mirror loop (device)
    Mh = api.CSRMat((n_f, n_b), M.rowptr, M.colind, M.val)
    Jh = api.CSRMat((n_f, n_f), A.rowptr, A.colind, A.val.copy())
    u_f = api.Vec(np.zeros(n_f))
    u_p = api.Vec(np.zeros(n_b))
    calls = []

    def assemble_cb(uf):
        R, Jv = residual_and_jacobian_values(uf.array)
        Jh.set_values(Jv)
        calls.append(1)
        return Jh, api.Vec(R)

    iife.plan_cache_clear()
    api.solveNonlinear(None, u_f, Mh, u_p, maxIters=20, relativeTolerance=1e-9, linear_method="cg",
                       linear_preconditioner="jacobi", assemble_cb=assemble_cb)
    out = capsys.readouterr().out
    assert "Newton solver iteration: 0" in out and len(calls) >= 3
    assert np.linalg.norm(u_p.array - u_b) <= 1e-7 * np.linalg.norm(u_b)
    assert np.allclose(u_f.array, oracle.spmv(M, u_p.array), rtol=0, atol=1e-7 * np.abs(u_f.array).max())
    # non-convergence ends the run like the reference's exit()
    with pytest.raises(SystemExit):
        api.solveNonlinear(None, api.Vec(np.zeros(n_f)), Mh, api.Vec(np.zeros(n_b)), maxIters=1, relativeTolerance=1e-30,
                           linear_method="cg", assemble_cb=assemble_cb)


def test_condition_estimate_matches_oracle(iife, oracle):
    """iife_ksp_solve_hessenberg / estimateConditionNumber (reference common.py:483-507) against the oracle."""
    from InterpolationBasedImmersedFEA import common as api
    from oracle.synthetic_cube import assemble_cube

    A, M, b = assemble_cube(5)
    C = oracle.AT_R_A(M, A)
    bb = oracle.AT_x(M, b)
    smax_o, smin_o, ro = oracle.estimate_condition_number(C, bb)
    info, R = iife.ksp_hessenberg(dmat(iife, C), bb, np.zeros(C.n_rows))
    assert info.reason == ro.reason and abs(info.iterations - ro.iterations) <= 1 and R.shape[0] == info.iterations
    assert np.allclose(np.tril(R, -1), 0.0)
    sv = np.linalg.svd(R, compute_uv=False)
    assert abs(sv.max() - smax_o) <= 1e-6 * smax_o and abs(sv.min() - smin_o) <= 1e-4 * smax_o
    u = api.Vec(np.zeros(C.n_rows))
    smax, smin = api.estimateConditionNumber(api.CSRMat((C.n_rows, C.n_cols), C.rowptr, C.colind, C.val), api.Vec(bb), u)
    assert abs(smax - smax_o) <= 1e-6 * smax_o and abs(smin - smin_o) <= 1e-4 * smax_o
    assert np.linalg.norm(u.array - ro.x) <= 1e-6 * np.linalg.norm(ro.x)
