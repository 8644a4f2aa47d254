"""CPU tests of the template compiler of the numeric PtAP (csrc/ptap_tpl_host.h) through its host-only hook
iife_tpl_emulate_row: the gather program compiled for one output row of A_b = M^T A_f M (reference
la_utils.py:165-182), interpreted on the CPU, must reproduce that row.  The CUDA kernel k_ptap_numeric_tpl executes the
same program step for step; its own parity tests are in test_gpu_parity.py."""
import ctypes

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import rand_csr


def _row_description(M, A, i):
    """What k_tpl_extract hands the host for output row i: operand lengths, R values, slot bytes, M values."""
    MT = M.T.tocsr()
    MT.sort_indices()
    js = MT.indices[MT.indptr[i]:MT.indptr[i + 1]]
    w = MT.data[MT.indptr[i]:MT.indptr[i + 1]]
    len1 = np.array([A.indptr[j + 1] - A.indptr[j] for j in js], dtype=np.int32)
    cols1 = np.concatenate([A.indices[A.indptr[j]:A.indptr[j + 1]] for j in js]) if len(js) else np.zeros(0, np.int64)
    a_vals = np.concatenate([A.data[A.indptr[j]:A.indptr[j + 1]] for j in js]) if len(js) else np.zeros(0)
    inter = np.unique(cols1)
    slot1 = np.searchsorted(inter, cols1).astype(np.uint8)
    len2 = np.array([M.indptr[k + 1] - M.indptr[k] for k in inter], dtype=np.int32)
    cols2 = np.concatenate([M.indices[M.indptr[k]:M.indptr[k + 1]] for k in inter]) if len(inter) else np.zeros(0, np.int64)
    mval = np.concatenate([M.data[M.indptr[k]:M.indptr[k + 1]] for k in inter]) if len(inter) else np.zeros(0)
    out = np.unique(cols2)
    slot2 = np.searchsorted(out, cols2).astype(np.uint8)
    return dict(len1=len1, w=np.ascontiguousarray(w), slot1=slot1, n1=len(inter), len2=len2, mval=np.ascontiguousarray(mval),
                slot2=slot2, n2=len(out), a_vals=np.ascontiguousarray(a_vals), out_cols=out)


def _emulate(d):
    from iife_b200 import _lib

    vp = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    c = np.full(max(d["n2"], 1), np.nan)
    info = np.zeros(10, dtype=np.int32)
    rc = _lib.lib.iife_tpl_emulate_row(len(d["len1"]), vp(d["len1"]), vp(d["w"]), vp(d["slot1"]), d["n1"], vp(d["len2"]),
                                       vp(d["mval"]), vp(d["slot2"]), d["n2"], vp(d["a_vals"]), vp(c), vp(info))
    return rc, c[:d["n2"]], info


def _check_rows(M, A, rows):
    C = (M.T @ A @ M).tocsr()
    scale = (abs(M).T @ abs(A) @ abs(M)).tocsr()
    n_checked = 0
    for i in rows:
        d = _row_description(M, A, i)
        if len(d["len1"]) == 0 or d["n1"] == 0 or d["n2"] == 0:
            continue
        if len(d["len1"]) > 64 or d["len1"].sum() > 1023 or d["n1"] > 256 or d["n2"] > 256 or d["len2"].sum() > 4095:
            rc, _, _ = _emulate(d)
            assert rc == 4  # IIFE_ERR_UNSUPPORTED: stays on the per-row kernels
            continue
        rc, c, info = _emulate(d)
        assert rc == 0
        ref = np.asarray(C[i, d["out_cols"]].todense()).ravel()
        sc = np.asarray(scale[i, d["out_cols"]].todense()).ravel()
        assert np.all(np.abs(c - ref) <= 1e-13 * sc + 1e-300), (i, np.abs(c - ref).max())
        T1, T2 = int(d["len1"].sum()), int(d["len2"].sum())
        assert info[0] >= (T1 + 31) // 32 and info[1] >= (T2 + 31) // 32
        n_checked += 1
    return n_checked


def test_cube_rows_compile_and_match():
    from oracle.synthetic_cube import assemble_cube

    Ao, Mo, _ = assemble_cube(4)
    A = sp.csr_matrix((Ao.val, Ao.colind, Ao.rowptr), shape=(Ao.n_rows, Ao.n_cols))
    M = sp.csr_matrix((Mo.val, Mo.colind, Mo.rowptr), shape=(Mo.n_rows, Mo.n_cols))
    assert _check_rows(M, A, range(M.shape[1])) == M.shape[1]
    # the interior row: 27 operand rows of 15 entries; both gather stages keep most lanes busy
    i = 2 * 25 + 2 * 5 + 2
    d = _row_description(M, A, i)
    _, _, info = _emulate(d)
    assert len(d["len1"]) == 27 and d["n2"] == 27
    assert info[6] >= 800 and info[7] >= 800, info  # lane use x1000
    assert info[8] == 1000 and info[9] <= 1500, info  # conflict degree x1000: stage 1 is conflict free by construction


@pytest.mark.parametrize("seed", range(4))
def test_random_operators(seed):
    rng = np.random.default_rng(seed)
    n_f, n_b = 300, 90
    rp, ci, v = rand_csr(rng, n_f, n_f, 6 + 3 * seed)
    A = sp.csr_matrix((v, ci, rp), shape=(n_f, n_f))
    rp, ci, v = rand_csr(rng, n_f, n_b, 1 + seed, empty_frac=0.2)
    M = sp.csr_matrix((v, ci, rp), shape=(n_f, n_b))
    assert _check_rows(M, A, range(n_b)) > 0


def test_split_destinations():
    """one destination fed by far more terms than 32 lanes x the mean: the compiler must split it into pieces"""
    n_f, n_b = 40, 3
    A = sp.csr_matrix(np.ones((n_f, n_f)))  # every operand row hits every intermediate column
    M = sp.csr_matrix(np.ones((n_f, 1)) @ np.array([[1.0, 0.5, 0.25]]))
    M = sp.csr_matrix(M[:, :n_b])
    Ms = sp.csr_matrix(M[:20])
    As = sp.csr_matrix(A[:20, :20])
    d = _row_description(Ms, As, 0)
    rc, c, info = _emulate(d)
    assert rc == 0
    ref = np.asarray((Ms.T @ As @ Ms).todense())[0]
    assert np.allclose(c, ref, rtol=1e-14)
    assert info[3] > 0  # stage 2 needed extra slots: 3 destinations x 20 terms on 32 lanes


def test_stored_zeros_and_cancellation_stay():
    A = sp.csr_matrix((np.array([1.0, -1.0, 0.0, 2.0]), np.array([0, 1, 0, 1]), np.array([0, 2, 4])), shape=(2, 2))
    M = sp.csr_matrix((np.array([1.0, 1.0]), np.array([0, 0]), np.array([0, 1, 2])), shape=(2, 1))
    d = _row_description(M, A, 0)
    rc, c, _ = _emulate(d)
    assert rc == 0 and d["n2"] == 1 and c[0] == 2.0
