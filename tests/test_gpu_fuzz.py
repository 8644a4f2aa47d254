"""Randomised parity sweep of the CUDA path against the oracle (hypothesis): shapes, row-length distributions,
empty rows/columns, single rows, rows far longer than a warp, stored zeros.  Same bars as test_gpu_parity.py.

First run on a B200 in round 2 (2 passed); part of the `-m gpu` suite since."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

pytestmark = [pytest.mark.gpu]


def _random_csr(rng, n_rows, n_cols, mean_len, heavy_rows, empty_frac, zero_frac):
    lens = np.minimum(rng.poisson(mean_len, n_rows), n_cols)
    if n_rows and heavy_rows:
        idx = rng.choice(n_rows, size=min(heavy_rows, n_rows), replace=False)
        lens[idx] = np.minimum(n_cols, rng.integers(n_cols // 2, n_cols + 1, size=idx.size))
    lens[rng.random(n_rows) < empty_frac] = 0
    rowptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(lens, out=rowptr[1:])
    colind = np.empty(int(rowptr[-1]), dtype=np.int32)
    for i in range(n_rows):
        colind[rowptr[i]:rowptr[i + 1]] = np.sort(rng.choice(n_cols, lens[i], replace=False))
    val = rng.standard_normal(int(rowptr[-1]))
    val[rng.random(val.size) < zero_frac] = 0.0  # stored zeros stay in the pattern
    return rowptr, colind, val


@settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 2 ** 31 - 1), n_f=st.integers(1, 900), n_b=st.integers(1, 400),
       m_len=st.sampled_from([0.5, 1.0, 3.0, 8.0, 27.0]), a_len=st.sampled_from([1.0, 5.0, 15.0, 60.0]),
       heavy=st.integers(0, 3), empty=st.sampled_from([0.0, 0.2, 0.7]), zeros=st.sampled_from([0.0, 0.1]))
def test_fuzz_ptap_spmv_transpose(iife, oracle, seed, n_f, n_b, m_len, a_len, heavy, empty, zeros):
    from test_gpu_parity import check_ptap, dmat

    rng = np.random.default_rng(seed)
    M = oracle.CSR(n_f, n_b, *_random_csr(rng, n_f, n_b, m_len, heavy, empty, zeros))
    A = oracle.CSR(n_f, n_f, *_random_csr(rng, n_f, n_f, a_len, heavy, empty / 2, zeros))
    dM, dA, dC, C, plan = check_ptap(iife, oracle, M, A)
    # second numeric call on new values through the same plan
    A2 = oracle.CSR(n_f, n_f, A.rowptr, A.colind, rng.standard_normal(A.nnz))
    dA.update_values(A2.val)
    C2 = oracle.AT_R_A(M, A2)
    v2 = plan.numeric(dM, dA, check_errors=True).values()
    scale = oracle.AT_R_A(oracle.CSR(n_f, n_b, M.rowptr, M.colind, np.abs(M.val)),
                          oracle.CSR(n_f, n_f, A2.rowptr, A2.colind, np.abs(A2.val))).val
    assert np.all(np.abs(v2 - C2.val) <= 1e-12 * scale + 1e-300)
    # transpose, SpMV in both directions
    T = oracle.transpose(M)
    rp, ci, v = dM.transpose().to_csr(np.int64)
    assert np.array_equal(rp, T.rowptr) and np.array_equal(ci, T.colind) and np.array_equal(v, T.val)
    x = rng.standard_normal(n_b)
    y = rng.standard_normal(n_f)
    absM = oracle.CSR(n_f, n_b, M.rowptr, M.colind, np.abs(M.val))
    assert np.all(np.abs(dM.spmv(x) - oracle.spmv(M, x)) <= 1e-13 * (oracle.spmv(absM, np.abs(x)) + 1e-300))
    assert np.all(np.abs(dM.spmv(y, trans=True) - oracle.AT_x(M, y)) <= 1e-13 * (oracle.AT_x(absM, np.abs(y)) + 1e-300))


@settings(max_examples=25, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(1, 700), mean=st.sampled_from([0.5, 4.0, 40.0]),
       empty=st.sampled_from([0.0, 0.5]), diag=st.sampled_from([1.0, 0.0, -3.0]))
def test_fuzz_row_edits_and_ksp(iife, oracle, seed, n, mean, empty, diag):
    from test_gpu_parity import dmat

    rng = np.random.default_rng(seed)
    A = oracle.CSR(n, n, *_random_csr(rng, n, n, mean, 1, empty, 0.05))
    dA = dmat(iife, A)
    rows = rng.choice(n, size=max(1, n // 5), replace=True)
    Z = oracle.zero_rows(A, rows, diag)
    rp, ci, v = dA.zero_rows(rows, diag).to_csr(np.int64)
    assert np.array_equal(rp, Z.rowptr) and np.array_equal(ci, Z.colind) and np.array_equal(v, Z.val)
    d = rng.standard_normal(n)
    D = oracle.add_diagonal(A, d)
    rp, ci, v = dA.add_diagonal(d).to_csr(np.int64)
    assert np.array_equal(rp, D.rowptr) and np.array_equal(ci, D.colind) and np.array_equal(v, D.val)
    # SPD system from the same pattern: S = B B^T + I (through the PtAP itself: M := B^T, A_f := I)
    I_n = oracle.CSR(n, n, np.arange(n + 1), np.arange(n, dtype=np.int32), np.ones(n))
    S = oracle.add_diagonal(oracle.AT_R_A(oracle.transpose(A), I_n), np.ones(n))
    b = rng.standard_normal(n)
    for method, kt in (("cg", iife.KSP_CG), ("gmres", iife.KSP_FGMRES)):
        ro = oracle.solve_ksp(S, b, method=method, max_it=4 * n + 50)
        x = np.zeros(n)
        info = iife.ksp_solve(dmat(iife, S), b, x, kt, iife.PC_JACOBI, max_it=4 * n + 50)
        assert info.reason == ro.reason, (method, info.reason_name, ro.reason)
        assert abs(info.iterations - ro.iterations) <= max(2, ro.iterations // 10)
        if ro.reason > 0:
            r_gpu = np.linalg.norm(b - oracle.spmv(S, x))
            r_orc = np.linalg.norm(b - oracle.spmv(S, ro.x))
            assert r_gpu <= 10.0 * max(r_orc, 1e-8 * np.linalg.norm(b))
