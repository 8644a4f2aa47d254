"""ctypes front end of oracle/iife_oracle.c — the CPU restatement of the reference's extraction path.

TEST INFRASTRUCTURE ONLY (see the header of iife_oracle.c): imported by tests/, by
__graft_entry__.smoke() as the checker and by bench.py's cpu_baseline / --impl reference legs.
The product (libiife.so, iife_b200, the la_utils/common mirror) never imports this module.

PARITY UNPINNED: PETSc is not available, the reference has no golden outputs.

The functions mirror the reference call sequence:
  * ``AT_R_A(M, A_f)``  reference la_utils.py:165-182:  AT = M^T; ATR = AT*A_f; ATT = M; ATRA = ATR*ATT
  * ``AT_x(M, b_f)``    reference la_utils.py:143-163:  MatMultTranspose into a fresh zero vector
  * ``solve_ksp``       reference common.py:509-641 (Krylov + Jacobi branch only)
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "iife_oracle.c")
_LIBDIR = os.path.join(_HERE, "_build")
_LIB = os.path.join(_LIBDIR, "liboracle.so")

_c_i64p = ctypes.POINTER(ctypes.c_int64)
_c_i32p = ctypes.POINTER(ctypes.c_int32)
_c_f64p = ctypes.POINTER(ctypes.c_double)


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc -O3 -fopenmp).  Returns the library path."""
    if not force and os.path.exists(_LIB) and os.path.getmtime(_LIB) >= os.path.getmtime(_SRC):
        return _LIB
    os.makedirs(_LIBDIR, exist_ok=True)
    cmd = ["gcc", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-o", _LIB, _SRC, "-lm"]
    try:
        subprocess.run(cmd, check=True, capture_output=True, text=True)
    except subprocess.CalledProcessError:
        # -march=native can be refused on exotic hosts: retry portable
        cmd.remove("-march=native")
        subprocess.run(cmd, check=True, capture_output=True, text=True)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _LIB
        if not os.path.exists(path) or (os.path.exists(_SRC) and os.path.getmtime(path) < os.path.getmtime(_SRC)):
            path = build()
        _lib = ctypes.CDLL(path)
        _lib.oracle_max_threads.restype = ctypes.c_int
        _lib.oracle_cg_jacobi.restype = ctypes.c_int
        _lib.oracle_fgmres_jacobi.restype = ctypes.c_int
        _lib.oracle_gcr_jacobi.restype = ctypes.c_int
        _lib.oracle_fgmres_hessenberg.restype = ctypes.c_int
    return _lib


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def set_threads(n: int) -> None:
    lib().oracle_set_threads(ctypes.c_int(int(n)))


@dataclass
class CSR:
    """AIJ matrix as PETSc stores it: int64 row pointers, int32 sorted column indices, fp64 values."""

    n_rows: int
    n_cols: int
    rowptr: np.ndarray
    colind: np.ndarray
    val: np.ndarray

    def __post_init__(self):
        self.rowptr = np.ascontiguousarray(self.rowptr, dtype=np.int64)
        self.colind = np.ascontiguousarray(self.colind, dtype=np.int32)
        self.val = np.ascontiguousarray(self.val, dtype=np.float64)
        assert self.rowptr.shape == (self.n_rows + 1,)
        assert self.colind.shape == self.val.shape == (int(self.rowptr[-1]),)

    @property
    def nnz(self) -> int:
        return int(self.rowptr[-1])

    @property
    def shape(self):
        return (self.n_rows, self.n_cols)

    def todense(self) -> np.ndarray:
        d = np.zeros((self.n_rows, self.n_cols))
        for i in range(self.n_rows):
            for p in range(self.rowptr[i], self.rowptr[i + 1]):
                d[i, self.colind[p]] += self.val[p]
        return d

    def pattern_dense(self) -> np.ndarray:
        d = np.zeros((self.n_rows, self.n_cols), dtype=bool)
        for i in range(self.n_rows):
            d[i, self.colind[self.rowptr[i]:self.rowptr[i + 1]]] = True
        return d

    @staticmethod
    def from_scipy(S) -> "CSR":
        """From a scipy CSR whose stored entries (explicit zeros included) are the AIJ entries."""
        S = S.tocsr()
        S.sort_indices()
        return CSR(S.shape[0], S.shape[1], S.indptr.astype(np.int64), S.indices.astype(np.int32), S.data.astype(np.float64))

    def to_scipy(self):
        import scipy.sparse as sp

        return sp.csr_matrix((self.val, self.colind, self.rowptr), shape=self.shape)


def _p(a, typ):
    return a.ctypes.data_as(typ)


def transpose(A: CSR) -> CSR:
    """MatTranspose (reference la_utils.py:178,180)."""
    t_rp = np.empty(A.n_cols + 1, dtype=np.int64)
    t_ci = np.empty(A.nnz, dtype=np.int32)
    t_v = np.empty(A.nnz, dtype=np.float64)
    lib().oracle_transpose(ctypes.c_int64(A.n_rows), ctypes.c_int64(A.n_cols), _p(A.rowptr, _c_i64p), _p(A.colind, _c_i32p),
                           _p(A.val, _c_f64p), _p(t_rp, _c_i64p), _p(t_ci, _c_i32p), _p(t_v, _c_f64p))
    return CSR(A.n_cols, A.n_rows, t_rp, t_ci, t_v)


def spgemm_symbolic(X: CSR, Y: CSR):
    """Structural product pattern of X*Y (MatMatMult symbolic): (rowptr, sorted colind)."""
    assert X.n_cols == Y.n_rows
    cnt = np.empty(X.n_rows, dtype=np.int64)
    lib().oracle_spgemm_count(ctypes.c_int64(X.n_rows), ctypes.c_int64(Y.n_cols), _p(X.rowptr, _c_i64p), _p(X.colind, _c_i32p),
                              _p(Y.rowptr, _c_i64p), _p(Y.colind, _c_i32p), _p(cnt, _c_i64p))
    rp = np.zeros(X.n_rows + 1, dtype=np.int64)
    np.cumsum(cnt, out=rp[1:])
    ci = np.empty(int(rp[-1]), dtype=np.int32)
    lib().oracle_spgemm_fill(ctypes.c_int64(X.n_rows), ctypes.c_int64(Y.n_cols), _p(X.rowptr, _c_i64p), _p(X.colind, _c_i32p),
                             _p(Y.rowptr, _c_i64p), _p(Y.colind, _c_i32p), _p(rp, _c_i64p), _p(ci, _c_i32p))
    return rp, ci


def spgemm_numeric(X: CSR, Y: CSR, rp: np.ndarray, ci: np.ndarray) -> CSR:
    v = np.empty(int(rp[-1]), dtype=np.float64)
    lib().oracle_spgemm_numeric(ctypes.c_int64(X.n_rows), ctypes.c_int64(Y.n_cols), _p(X.rowptr, _c_i64p), _p(X.colind, _c_i32p),
                                _p(X.val, _c_f64p), _p(Y.rowptr, _c_i64p), _p(Y.colind, _c_i32p), _p(Y.val, _c_f64p),
                                _p(rp, _c_i64p), _p(ci, _c_i32p), _p(v, _c_f64p))
    return CSR(X.n_rows, Y.n_cols, rp, ci, v)


def matmult(X: CSR, Y: CSR) -> CSR:
    """MatMatMult(MAT_INITIAL_MATRIX): symbolic + numeric (reference la_utils.py:179,181)."""
    rp, ci = spgemm_symbolic(X, Y)
    return spgemm_numeric(X, Y, rp, ci)


def AT_R_A(M: CSR, A_f: CSR, return_intermediate: bool = False):
    """Reference la_utils.py:165-182 with A := M, R := A_f:
    AT = M^T (:178); ATR = AT*R (:179); ATT = M (:180); ATRA = ATR*ATT (:181)."""
    AT = transpose(M)
    ATR = matmult(AT, A_f)
    ATRA = matmult(ATR, M)
    return (ATRA, ATR) if return_intermediate else ATRA


def spmv(A: CSR, x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float64)
    assert x.shape == (A.n_cols,)
    y = np.empty(A.n_rows, dtype=np.float64)
    lib().oracle_spmv(ctypes.c_int64(A.n_rows), _p(A.rowptr, _c_i64p), _p(A.colind, _c_i32p), _p(A.val, _c_f64p),
                      _p(x, _c_f64p), _p(y, _c_f64p))
    return y


def AT_x(M: CSR, x: np.ndarray) -> np.ndarray:
    """Reference la_utils.py:143-163: b = M^T x through MatMultTranspose into a new zero vector."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    assert x.shape == (M.n_rows,)
    y = np.empty(M.n_cols, dtype=np.float64)
    lib().oracle_spmv_transpose(ctypes.c_int64(M.n_rows), ctypes.c_int64(M.n_cols), _p(M.rowptr, _c_i64p), _p(M.colind, _c_i32p),
                                _p(M.val, _c_f64p), _p(x, _c_f64p), _p(y, _c_f64p))
    return y


def jacobi_inverse(A: CSR) -> np.ndarray:
    d = np.empty(A.n_rows, dtype=np.float64)
    lib().oracle_jacobi_inverse(ctypes.c_int64(A.n_rows), _p(A.rowptr, _c_i64p), _p(A.colind, _c_i32p), _p(A.val, _c_f64p),
                                _p(d, _c_f64p))
    return d


def diagonal(A: CSR) -> np.ndarray:
    """MatGetDiagonal (reference common.py:222,305): stored (i, i) or 0."""
    d = np.zeros(A.n_rows)
    for i in range(min(A.n_rows, A.n_cols)):
        seg = A.colind[A.rowptr[i]:A.rowptr[i + 1]]
        k = np.searchsorted(seg, i)
        if k < seg.size and seg[k] == i:
            d[i] = A.val[A.rowptr[i] + k]
    return d


def zero_rows(A: CSR, rows, diag: float = 1.0) -> CSR:
    """MatZeroRows as trimNodes calls it (reference common.py:284,327; AIJ, no KEEP_NONZERO_PATTERN): a listed
    row keeps the single entry (i, i) = diag when diag != 0 (and the position exists), nothing otherwise."""
    flag = np.zeros(A.n_rows, dtype=bool)
    flag[np.asarray(rows, dtype=np.int64)] = True
    rp = [0]
    ci, v = [], []
    for i in range(A.n_rows):
        if flag[i]:
            if diag != 0.0 and i < A.n_cols:
                ci.append(np.array([i], dtype=np.int32))
                v.append(np.array([diag]))
                rp.append(rp[-1] + 1)
            else:
                rp.append(rp[-1])
        else:
            b, e = A.rowptr[i], A.rowptr[i + 1]
            ci.append(A.colind[b:e])
            v.append(A.val[b:e])
            rp.append(rp[-1] + int(e - b))
    cat = (lambda parts, dt: np.concatenate(parts).astype(dt) if parts else np.zeros(0, dtype=dt))
    return CSR(A.n_rows, A.n_cols, np.array(rp, dtype=np.int64), cat(ci, np.int32), cat(v, np.float64))


def add_diagonal(A: CSR, d: np.ndarray) -> CSR:
    """``A += A0`` with ``A0.setDiagonal(vd)`` on an empty matrix (removeZeroDiagonal, reference common.py:243-249):
    MatAXPY over different patterns => pattern union(A, full diagonal), sorted rows, (i, i) = A_ii + d_i."""
    rp = [0]
    ci, v = [], []
    for i in range(A.n_rows):
        b, e = A.rowptr[i], A.rowptr[i + 1]
        cols, vals = A.colind[b:e], A.val[b:e].copy()
        if i < A.n_cols:
            k = int(np.searchsorted(cols, i))
            if k < cols.size and cols[k] == i:
                vals[k] = vals[k] + d[i]
            else:
                cols = np.insert(cols, k, i)
                vals = np.insert(vals, k, 0.0 + d[i])
        ci.append(cols)
        v.append(vals)
        rp.append(rp[-1] + cols.size)
    cat = (lambda parts, dt: np.concatenate(parts).astype(dt) if parts else np.zeros(0, dtype=dt))
    return CSR(A.n_rows, A.n_cols, np.array(rp, dtype=np.int64), cat(ci, np.int32), cat(v, np.float64))


def create_nonzero_diagonal(A: CSR, bfr_tol: float = 1e-9) -> np.ndarray:
    """createNonzeroDiagonal (reference common.py:207-233): 1 where |A_ii| <= bfr_tol, else 0."""
    return np.where(np.abs(diagonal(A)) <= bfr_tol, 1.0, 0.0)


def trim_nodes(A: CSR, b=None, bfr_tol: float = 1e-9, target=None, zero_vec=None):
    """trimNodes (reference common.py:262-332): rows with A_ii <= bfr_tol (SIGNED test, :312) — or the rows of
    ``zero_vec`` — become unit rows; b there becomes target or 0.  Returns (A', b', ids)."""
    ids = np.asarray(zero_vec, dtype=np.int64) if zero_vec is not None else np.flatnonzero(diagonal(A) <= bfr_tol)
    A2 = zero_rows(A, ids, 1.0)
    b2 = None
    if b is not None:
        b2 = np.array(b, dtype=np.float64, copy=True)
        b2[ids] = 0.0 if target is None else np.asarray(target)[ids]
    return A2, b2, ids


@dataclass
class KSPResult:
    x: np.ndarray
    iterations: int
    reason: int
    rnorm: float
    history: np.ndarray


def solve_ksp(A: CSR, b: np.ndarray, x0: np.ndarray | None = None, method: str = "gmres", PC: str = "jacobi",
              rtol: float = 1e-8, atol: float = 1e-9, max_it: int = 1000000, restart: int = 300, dtol: float = 1e4,
              hist_len: int = 0) -> KSPResult:
    """Krylov branch of reference common.py:509-641: 'gmres' -> FGMRES(300), 'cg' -> CG, 'gcr' -> GCR (PETSc's default
    restart of 30: pass restart=30; the reference's setGMRESRestart does not reach KSPGCR), PC jacobi."""
    assert A.n_rows == A.n_cols
    n = A.n_rows
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
    if PC == "jacobi":
        dinv = jacobi_inverse(A)
        dp = _p(dinv, _c_f64p)
    elif PC in (None, "none"):
        dinv = None
        dp = ctypes.cast(None, _c_f64p)
    else:
        raise NotImplementedError(PC)
    its = ctypes.c_int64(0)
    rn = ctypes.c_double(0.0)
    hist = np.zeros(max(hist_len, 1), dtype=np.float64)
    common = (ctypes.c_int64(n), _p(A.rowptr, _c_i64p), _p(A.colind, _c_i32p), _p(A.val, _c_f64p), dp, _p(b, _c_f64p),
              _p(x, _c_f64p), ctypes.c_double(rtol), ctypes.c_double(atol), ctypes.c_double(dtol), ctypes.c_int64(max_it))
    if method == "cg":
        reason = lib().oracle_cg_jacobi(*common, ctypes.byref(its), ctypes.byref(rn), _p(hist, _c_f64p), ctypes.c_int64(hist_len))
    elif method in ("gmres", None):
        reason = lib().oracle_fgmres_jacobi(*common, ctypes.c_int(restart), ctypes.byref(its), ctypes.byref(rn),
                                            _p(hist, _c_f64p), ctypes.c_int64(hist_len))
    elif method == "gcr":
        reason = lib().oracle_gcr_jacobi(*common, ctypes.c_int(30 if restart in (None, 300) else restart), ctypes.byref(its),
                                         ctypes.byref(rn), _p(hist, _c_f64p), ctypes.c_int64(hist_len))
    else:
        raise NotImplementedError(method)
    return KSPResult(x, int(its.value), int(reason), float(rn.value), hist[:hist_len])


def estimate_condition_number(A: CSR, b: np.ndarray, x0: np.ndarray | None = None, bfr_tol=None, rtol: float = 1e-8,
                              atol: float = 1e-9, max_it: int = 100000, PC=None, restart: int = 1000):
    """estimateConditionNumber (reference common.py:483-507): GMRES(1000) with ``ksp.setComputeSingularValues``,
    then ``computeExtremeSingularValues`` = extreme singular values of the Hessenberg matrix of the last cycle.
    PETSc's KSPGMRES and the FGMRES restated here build the same Hessenberg matrix when there is no
    preconditioner (the reference's default ``PC=None`` -> "none"); with Jacobi this restatement is the
    right-preconditioned operator A D^-1 (PETSc: left, D^-1 A).  Returns (smax, smin, KSPResult)."""
    if bfr_tol is not None:
        A, b, _ = trim_nodes(A, b, bfr_tol=bfr_tol)
    n = A.n_rows
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
    if PC == "jacobi":
        dinv = jacobi_inverse(A)
        dp = _p(dinv, _c_f64p)
    elif PC in (None, "none"):
        dp = ctypes.cast(None, _c_f64p)
    else:
        raise NotImplementedError(PC)
    m = max(1, int(restart))
    if max_it > 0:
        m = min(m, int(max_it))
    R = np.zeros((m, m), dtype=np.float64, order="F")
    its, k, rn = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_double(0.0)
    reason = lib().oracle_fgmres_hessenberg(ctypes.c_int64(n), _p(A.rowptr, _c_i64p), _p(A.colind, _c_i32p), _p(A.val, _c_f64p),
                                            dp, _p(b, _c_f64p), _p(x, _c_f64p), ctypes.c_double(rtol), ctypes.c_double(atol),
                                            ctypes.c_double(1e4), ctypes.c_int64(max_it), ctypes.c_int(m), ctypes.byref(its),
                                            ctypes.byref(rn), _p(R, _c_f64p), ctypes.byref(k))
    kk = int(k.value)
    res = KSPResult(x, int(its.value), int(reason), float(rn.value), np.zeros(0))
    if kk == 0:
        return 0.0, 0.0, res
    sv = np.linalg.svd(R[:kk, :kk], compute_uv=False)
    return float(sv.max()), float(sv.min()), res
