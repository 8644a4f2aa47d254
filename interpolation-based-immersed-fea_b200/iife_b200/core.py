"""Thin object layer over the C ABI: device matrices, PtAP plans, SpMV and KSP.

Host data are numpy arrays; device data are anything with ``data_ptr()`` (torch tensors) or a raw
integer address.  All arithmetic happens in libiife.so on the GPU bound by :func:`init`.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import check, lib

MEM_HOST, MEM_DEVICE = 0, 1
KSP_CG, KSP_FGMRES, KSP_GCR = 0, 1, 2
PC_NONE, PC_JACOBI = 0, 1

REASONS = {
    2: "CONVERGED_RTOL", 3: "CONVERGED_ATOL", 4: "CONVERGED_ITS", 0: "ITERATING", -3: "DIVERGED_ITS",
    -4: "DIVERGED_DTOL", -5: "DIVERGED_BREAKDOWN", -8: "DIVERGED_INDEFINITE_PC", -9: "DIVERGED_NANORINF",
    -10: "DIVERGED_INDEFINITE_MAT",
}

_initialised = False
_device_index = -1


def init(device: int = 0) -> None:
    """Bind this process to one GPU.  Raises if there is no usable sm_100 device (no CPU fallback)."""
    global _initialised, _device_index
    check(lib.iife_init(int(device)))
    _initialised = True
    _device_index = int(device)


def current_device() -> int:
    """Index of the GPU this process is bound to (-1 before init)."""
    return _device_index if _initialised else -1


def is_initialised() -> bool:
    return _initialised


def finalize() -> None:
    global _initialised
    check(lib.iife_finalize())
    _initialised = False


def device_count() -> int:
    n = ctypes.c_int(0)
    rc = lib.iife_device_count(ctypes.byref(n))
    return int(n.value) if rc == 0 else 0


def set_stream(stream_ptr) -> None:
    check(lib.iife_set_stream(ctypes.c_void_p(int(stream_ptr) if stream_ptr else 0)))


def sync() -> None:
    check(lib.iife_sync())


def device_bytes() -> int:
    b = ctypes.c_int64(0)
    check(lib.iife_device_bytes(ctypes.byref(b)))
    return int(b.value)


def launch_count(reset: bool = False) -> int:
    n = ctypes.c_int64(0)
    check(lib.iife_launch_count(ctypes.byref(n), 1 if reset else 0))
    return int(n.value)


def _ptr(a):
    """(address, keepalive) of a host numpy array, a device tensor or a raw address."""
    if a is None:
        return ctypes.c_void_p(0), None
    if isinstance(a, np.ndarray):
        return ctypes.c_void_p(a.ctypes.data), a
    if hasattr(a, "data_ptr"):
        return ctypes.c_void_p(int(a.data_ptr())), a
    return ctypes.c_void_p(int(a)), None


def _is_device(a) -> bool:
    return hasattr(a, "data_ptr") and getattr(a, "is_cuda", False)


def _host_f64(a, n=None):
    out = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and out.shape != (n,):
        raise ValueError(f"expected a vector of length {n}, got shape {out.shape}")
    return out


class DeviceMat:
    """Device-resident CSR (AIJ) matrix (handle of ``iife_mat``)."""

    def __init__(self, handle, owner=True):
        self._h = ctypes.c_void_p(handle)
        self._owner = owner

    # ---- construction
    @classmethod
    def from_csr(cls, n_rows, n_cols, rowptr, colind, val=None) -> "DeviceMat":
        dev = _is_device(rowptr)
        if dev:
            import torch

            idx_bytes = 8 if rowptr.dtype == torch.int64 else 4
            if colind.dtype != rowptr.dtype:
                raise TypeError("rowptr and colind must have the same integer dtype")
            if val is not None and val.dtype != torch.float64:
                raise TypeError("values must be float64")
            rp, ci, v = rowptr.contiguous(), colind.contiguous(), (val.contiguous() if val is not None else None)
        else:
            rp = np.ascontiguousarray(rowptr)
            if rp.dtype not in (np.int32, np.int64):
                rp = rp.astype(np.int64)
            ci = np.ascontiguousarray(colind, dtype=rp.dtype)
            v = None if val is None else np.ascontiguousarray(val, dtype=np.float64)
            idx_bytes = rp.dtype.itemsize
            if rp.shape != (int(n_rows) + 1,):
                raise ValueError(f"rowptr must have n_rows+1 = {int(n_rows) + 1} entries, got {rp.shape}")
            if ci.shape != (int(rp[-1]),) or (v is not None and v.shape != ci.shape):
                raise ValueError("colind/val length must equal rowptr[-1]")
        h = ctypes.c_void_p(0)
        check(lib.iife_mat_create_csr(int(n_rows), int(n_cols), _ptr(rp)[0], _ptr(ci)[0], _ptr(v)[0], idx_bytes,
                                      MEM_DEVICE if dev else MEM_HOST, ctypes.byref(h)))
        return cls(h.value)

    @classmethod
    def from_scipy(cls, S) -> "DeviceMat":
        S = S.tocsr()
        if not S.has_sorted_indices:
            S = S.sorted_indices()
        return cls.from_csr(S.shape[0], S.shape[1], S.indptr, S.indices, S.data)

    # ---- info
    @property
    def handle(self):
        return self._h

    def info(self):
        a, b, c = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
        check(lib.iife_mat_get_info(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return int(a.value), int(b.value), int(c.value)

    @property
    def shape(self):
        r, c, _ = self.info()
        return (r, c)

    @property
    def nnz(self):
        return self.info()[2]

    def fingerprint(self) -> int:
        fp = ctypes.c_uint64(0)
        check(lib.iife_mat_fingerprint(self._h, ctypes.byref(fp)))
        return int(fp.value)

    def device_ptrs(self):
        a, b, c = ctypes.c_void_p(0), ctypes.c_void_p(0), ctypes.c_void_p(0)
        check(lib.iife_mat_device_ptrs(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return a.value, b.value, c.value

    # ---- data movement
    def update_values(self, val) -> None:
        if _is_device(val):
            check(lib.iife_mat_update_values(self._h, _ptr(val.contiguous())[0], MEM_DEVICE))
        else:
            v = _host_f64(val, self.nnz)
            check(lib.iife_mat_update_values(self._h, _ptr(v)[0], MEM_HOST))

    def to_csr(self, index_dtype=np.int32, out=None):
        """(rowptr, colind, val) as host numpy arrays (optionally into preallocated ``out``)."""
        n_rows, _, nnz = self.info()
        index_dtype = np.dtype(index_dtype)
        if out is None:
            rp = np.empty(n_rows + 1, dtype=index_dtype)
            ci = np.empty(nnz, dtype=index_dtype)
            v = np.empty(nnz, dtype=np.float64)
        else:
            rp, ci, v = out
        check(lib.iife_mat_get_csr(self._h, _ptr(rp)[0], _ptr(ci)[0], _ptr(v)[0], index_dtype.itemsize, MEM_HOST))
        return rp, ci, v

    def values(self) -> np.ndarray:
        v = np.empty(self.nnz, dtype=np.float64)
        check(lib.iife_mat_get_csr(self._h, None, None, _ptr(v)[0], 4, MEM_HOST))
        return v

    def to_scipy(self):
        import scipy.sparse as sp

        rp, ci, v = self.to_csr()
        return sp.csr_matrix((v, ci, rp), shape=self.shape)

    # ---- operations
    def transpose(self) -> "DeviceMat":
        h = ctypes.c_void_p(0)
        check(lib.iife_mat_transpose(self._h, ctypes.byref(h)))
        return DeviceMat(h.value)

    def diagonal(self) -> np.ndarray:
        d = np.empty(self.shape[0], dtype=np.float64)
        check(lib.iife_mat_get_diagonal(self._h, _ptr(d)[0], MEM_HOST))
        return d

    def zero_rows(self, rows, diag=1.0) -> "DeviceMat":
        """New matrix with the listed rows reduced to the single entry (i, i) = diag (MatZeroRows as used by
        trimNodes, reference common.py:284,327); see include/iife.h."""
        r = np.ascontiguousarray(np.atleast_1d(rows))
        if r.dtype not in (np.int32, np.int64):
            r = r.astype(np.int64)
        h = ctypes.c_void_p(0)
        check(lib.iife_mat_zero_rows(self._h, _ptr(r)[0], int(r.size), r.dtype.itemsize, float(diag), MEM_HOST,
                                     ctypes.byref(h)))
        return DeviceMat(h.value)

    def add_diagonal(self, d) -> "DeviceMat":
        """New matrix A + diag(d) on the pattern union(A, full diagonal) (removeZeroDiagonal, common.py:243-249)."""
        dv = _host_f64(d, self.shape[0])
        h = ctypes.c_void_p(0)
        check(lib.iife_mat_add_diagonal(self._h, _ptr(dv)[0], MEM_HOST, ctypes.byref(h)))
        return DeviceMat(h.value)

    def spmv(self, x, y=None, trans=False, alpha=1.0, beta=0.0):
        """y = alpha*op(A) x + beta*y.  Host numpy in -> numpy out; device tensors in -> written in place."""
        n_rows, n_cols, _ = self.info()
        n_in, n_out = (n_rows, n_cols) if trans else (n_cols, n_rows)
        if _is_device(x):
            if y is None:
                import torch

                y = torch.empty(n_out, dtype=torch.float64, device=x.device)
            check(lib.iife_spmv(self._h, int(trans), float(alpha), _ptr(x)[0], float(beta), _ptr(y)[0], MEM_DEVICE))
            return y
        xh = _host_f64(x, n_in)
        if y is None:
            if beta != 0.0:
                raise ValueError("beta != 0 needs y")
            yh = np.empty(n_out, dtype=np.float64)
        else:
            yh = y if (isinstance(y, np.ndarray) and y.dtype == np.float64 and y.flags.c_contiguous) else _host_f64(y, n_out)
        check(lib.iife_spmv(self._h, int(trans), float(alpha), _ptr(xh)[0], float(beta), _ptr(yh)[0], MEM_HOST))
        return yh

    def destroy(self) -> None:
        if self._h and self._h.value and self._owner:
            lib.iife_mat_destroy(self._h)
        self._h = ctypes.c_void_p(0)

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class PtapPlan:
    """Symbolic PtAP plan: transpose of M, pattern of A_b, numeric row bins (``iife_plan``)."""

    def __init__(self, M: DeviceMat, A: DeviceMat):
        h = ctypes.c_void_p(0)
        check(lib.iife_ptap_symbolic(M.handle, A.handle, ctypes.byref(h)))
        self._h = h

    def info(self):
        a, b, c = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
        check(lib.iife_plan_get_info(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return {"n_b": int(a.value), "nnz_c": int(b.value), "nnz_intermediate": int(c.value)}

    def bin_counts(self):
        """rows handled by each numeric kernel (see include/iife.h: iife_plan_bin_counts)"""
        c = (ctypes.c_int64 * 8)()
        check(lib.iife_plan_bin_counts_n(self._h, c, 8))
        return list(c)

    def tpl_info(self):
        """template plan of the numeric phase (include/iife.h: iife_plan_tpl_info); zeros before the first numeric call"""
        a, b, c = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
        use = (ctypes.c_double * 2)()
        check(lib.iife_plan_tpl_info(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), use))
        return {"templates": int(a.value), "rows": int(b.value), "chunks": int(c.value), "lane_use": [use[0], use[1]]}

    def matches(self, M: DeviceMat, A: DeviceMat) -> bool:
        m = ctypes.c_int(0)
        check(lib.iife_plan_matches(self._h, M.handle, A.handle, ctypes.byref(m)))
        return bool(m.value)

    def numeric(self, M: DeviceMat, A: DeviceMat, C: DeviceMat | None = None, check_errors: bool = False) -> DeviceMat:
        h = ctypes.c_void_p(C.handle.value if C is not None else 0)
        check(lib.iife_ptap_numeric(self._h, M.handle, A.handle, ctypes.byref(h)))
        if check_errors:
            check(lib.iife_plan_check(self._h))
        return C if C is not None else DeviceMat(h.value)

    def check(self) -> None:
        check(lib.iife_plan_check(self._h))

    def destroy(self) -> None:
        if self._h and self._h.value:
            lib.iife_plan_destroy(self._h)
        self._h = ctypes.c_void_p(0)

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def ptap(M: DeviceMat, A: DeviceMat):
    """A_b = M^T A M with an internally cached symbolic plan (the AT_R_A path).  Returns (C, cached)."""
    h = ctypes.c_void_p(0)
    cached = ctypes.c_int(0)
    check(lib.iife_ptap(M.handle, A.handle, ctypes.byref(h), ctypes.byref(cached)))
    return DeviceMat(h.value), bool(cached.value)


def plan_cache_clear() -> None:
    check(lib.iife_plan_cache_clear())


@dataclass
class KSPInfo:
    iterations: int
    reason: int
    rnorm: float
    rnorm0: float
    history: np.ndarray

    @property
    def converged(self) -> bool:
        return self.reason > 0

    @property
    def reason_name(self) -> str:
        return REASONS.get(self.reason, str(self.reason))


def ksp_solve(A: DeviceMat, b, x, ksp_type=KSP_FGMRES, pc_type=PC_JACOBI, rtol=1e-8, atol=1e-9, dtol=1e4,
              max_it=1000000, restart=300, hist_len=0) -> KSPInfo:
    """Solve A x = b; ``x`` holds the initial guess and is overwritten (numpy array or device tensor)."""
    n = A.shape[0]
    res = _lib.KspResult()
    hist = np.zeros(max(int(hist_len), 1), dtype=np.float64)
    hp = _ptr(hist)[0] if hist_len > 0 else ctypes.c_void_p(0)
    if _is_device(x):
        if not _is_device(b):
            raise TypeError("b and x must live in the same memory space")
        check(lib.iife_ksp_solve(A.handle, ksp_type, pc_type, rtol, atol, dtol, int(max_it), int(restart),
                                 _ptr(b)[0], _ptr(x)[0], MEM_DEVICE, None, ctypes.byref(res), hp, int(hist_len)))
    else:
        if not (isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags.c_contiguous and x.shape == (n,)):
            raise TypeError("x must be a contiguous float64 numpy vector of length n (it is updated in place)")
        bh = _host_f64(b, n)
        check(lib.iife_ksp_solve(A.handle, ksp_type, pc_type, rtol, atol, dtol, int(max_it), int(restart),
                                 _ptr(bh)[0], _ptr(x)[0], MEM_HOST, None, ctypes.byref(res), hp, int(hist_len)))
    return KSPInfo(int(res.iterations), int(res.reason), float(res.rnorm), float(res.rnorm0), hist[:hist_len])


def ksp_hessenberg(A: DeviceMat, b, x, pc_type=PC_NONE, rtol=1e-8, atol=1e-9, dtol=1e4, max_it=100000, restart=1000):
    """FGMRES solve (host numpy ``b``/``x``; ``x`` is updated in place) that also returns the triangular factor R
    (k x k) of the last cycle's Hessenberg matrix: ``np.linalg.svd(R)`` gives the singular values behind the
    reference's ``estimateConditionNumber`` (common.py:483-507).  Returns (KSPInfo, R)."""
    n = A.shape[0]
    if not (isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags.c_contiguous and x.shape == (n,)):
        raise TypeError("x must be a contiguous float64 numpy vector of length n (it is updated in place)")
    bh = _host_f64(b, n)
    restart = int(restart) if int(restart) >= 1 else 30  # the library's default
    m = max(1, min(restart, int(max_it) if max_it > 0 else restart, 10000))
    R = np.zeros(m * m, dtype=np.float64)
    k = ctypes.c_int64(0)
    res = _lib.KspResult()
    check(lib.iife_ksp_solve_hessenberg(A.handle, int(pc_type), rtol, atol, dtol, int(max_it), int(restart), _ptr(bh)[0],
                                        _ptr(x)[0], MEM_HOST, ctypes.byref(res), _ptr(R)[0], int(R.size), ctypes.byref(k)))
    kk = int(k.value)
    info = KSPInfo(int(res.iterations), int(res.reason), float(res.rnorm), float(res.rnorm0), np.zeros(0))
    return info, R[:kk * kk].reshape((kk, kk), order="F").copy()


def synth_cube(n_bg_cells: int, sigma: float = 1.0, row_begin: int = 0, row_end: int | None = None, b_f=None):
    """Generate the S1 cube operands on the device.  Returns (A_f, M) as DeviceMat; fills ``b_f``
    (device tensor of length row_end-row_begin) if given."""
    from . import synthetic

    sz = synthetic.cube_sizes(n_bg_cells)
    if row_end is None:
        row_end = sz["n_f"]
    coef, load8 = synthetic.element_tables(n_bg_cells, sigma)
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    load8 = np.ascontiguousarray(load8, dtype=np.float64)
    ha, hm = ctypes.c_void_p(0), ctypes.c_void_p(0)
    check(lib.iife_synth_cube_build(int(n_bg_cells), int(row_begin), int(row_end), _ptr(coef)[0], _ptr(load8)[0],
                                    ctypes.byref(ha), ctypes.byref(hm), _ptr(b_f)[0]))
    return DeviceMat(ha.value), DeviceMat(hm.value)
