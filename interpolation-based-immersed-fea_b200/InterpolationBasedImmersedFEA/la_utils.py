"""Drop-in mirror of the reference's ``InterpolationBasedImmersedFEA.la_utils`` for the extraction
hot path: same names, same positional signatures, same error behaviour
(reference InterpolationBasedImmersedFEA/la_utils.py:28-182), with the arithmetic done by libiife.so
on a B200 instead of PETSc on the host.

What crosses the boundary
  * ``petsc4py.PETSc.Mat`` / ``Vec`` and dolfin wrappers when those packages are importable (the
    matrices are read through ``Mat.getValuesCSR()`` and results are wrapped back with
    ``Mat.createAIJ(csr=...)`` so that everything callers do next — createVecLeft, getDiagonal,
    zeroRows, ksp.setOperators, reference common.py:206-329 — keeps working on a real PETSc.Mat);
  * otherwise the light :class:`CSRMat` / :class:`Vec` objects below, which carry the PETSc method
    names the reference's callers use, plus scipy CSR matrices and numpy vectors.

There is no CPU fallback: every product runs in the CUDA library, and importing this module without
a built libiife.so fails.
"""
from __future__ import annotations

import numpy as np

import iife_b200 as _iife
from iife_b200 import DeviceMat

try:  # the reference's own imports (la_utils.py:6-8); absent in this image
    from petsc4py import PETSc  # type: ignore

    HAVE_PETSC = True
except Exception:  # pragma: no cover - depends on the environment
    PETSc = None
    HAVE_PETSC = False
try:
    import dolfin as _dolfin  # type: ignore

    HAVE_DOLFIN = True
except Exception:  # pragma: no cover
    _dolfin = None
    HAVE_DOLFIN = False

if HAVE_DOLFIN:
    # the reference module star-imports dolfin (la_utils.py:6-8) and every demo relies on getting dolfin's names
    # through `from InterpolationBasedImmersedFEA.la_utils import *` (demos/poisson.py:14-16)
    from dolfin import *  # type: ignore # noqa: F401,F403,E402
    try:
        from dolfin.cpp.log import *  # type: ignore # noqa: F401,F403,E402
    except Exception:  # pragma: no cover
        pass
    try:  # module-level settings of the reference (la_utils.py:12-26)
        INFO = LogLevel.INFO  # noqa: F405
        set_log_level(INFO)  # noqa: F405
        parameters['std_out_all_processes'] = False  # noqa: F405
        worldcomm = MPI.comm_world  # noqa: F405
        mpirank = MPI.rank(worldcomm)  # noqa: F405
        mpisize = MPI.size(worldcomm)  # noqa: F405
        DOLFIN_FUNCTION = function.function.Function  # noqa: F405
        DOLFIN_VECTOR = cpp.la.Vector  # noqa: F405
        DOLFIN_MATRIX = cpp.la.Matrix  # noqa: F405
        DOLFIN_PETSCVECTOR = cpp.la.PETScVector  # noqa: F405
        DOLFIN_PETSCMATRIX = cpp.la.PETScMatrix  # noqa: F405
    except Exception:  # pragma: no cover - a dolfin without these names
        worldcomm, mpirank, mpisize = None, 0, 1
else:
    worldcomm, mpirank, mpisize = None, 0, 1
if HAVE_PETSC:
    PETSC4PY_VECTOR = PETSc.Vec
    PETSC4PY_MATRIX = PETSc.Mat


def _ensure_init():
    if not _iife.is_initialised():
        import os

        _iife.init(int(os.environ.get("LOCAL_RANK", "0")))


# --------------------------------------------------------------------------------------------------
# light stand-ins with the PETSc method names used by the reference's callers
# --------------------------------------------------------------------------------------------------
class Vec:
    """Dense fp64 vector with the subset of the ``PETSc.Vec`` API the reference's callers use.

    Like :class:`CSRMat` the host array is lazy: a vector produced on the GPU (``AT_x``) stays there until
    somebody reads ``.array``; ``solveKSP`` consumes it without a round trip through host memory.  Reading
    ``.array`` hands out a mutable numpy array, so from then on the host copy is the only one."""

    def __init__(self, array=None, device=None):
        if array is None and device is None:
            raise ValueError("Vec needs a host array or a device tensor")
        self._array = None if array is None else np.ascontiguousarray(array, dtype=np.float64)
        self._dev = device if array is None else None  # float64 CUDA tensor holding the current value

    @property
    def array(self):
        if self._array is None:
            _iife.sync()  # the producer ran on the library's stream
            self._array = self._dev.cpu().numpy()
        self._dev = None
        return self._array

    @array.setter
    def array(self, a):
        self._array = np.ascontiguousarray(a, dtype=np.float64)
        self._dev = None

    def device_tensor(self):
        """The device-resident value, or None when the vector lives on the host."""
        return self._dev

    # PETSc.Vec API
    def getSize(self):
        return int(self._array.size if self._array is not None else self._dev.numel())

    def getSizes(self):
        n = self.getSize()
        return (n, n)

    def getArray(self):
        return self.array

    def setArray(self, a):
        self.array[:] = a

    def norm(self):
        return float(np.linalg.norm(self.array))

    def copy(self):
        return Vec(self.array.copy())

    def duplicate(self):
        return Vec(np.zeros_like(self.array))

    def set(self, alpha):
        self.array[:] = alpha

    def axpy(self, alpha, x):
        self.array += alpha * arg2v(x).array

    def assemble(self):
        return None

    def ghostUpdate(self, *args, **kwargs):
        return None

    def getOwnershipRange(self):
        return (0, self.getSize())

    def getValue(self, i):
        return float(self.array[i])

    def setValue(self, i, v):
        self.array[i] = v

    def getComm(self):
        return None

    # arithmetic used by the Newton drivers: u_p += -du_p*relax_param (reference common.py:394,474)
    def __neg__(self):
        return Vec(-self.array)

    def __mul__(self, alpha):
        return Vec(self.array * alpha)

    __rmul__ = __mul__

    def __iadd__(self, other):
        self.array += arg2v(other).array
        return self

    def __isub__(self, other):
        self.array -= arg2v(other).array
        return self

    def __len__(self):
        return self.getSize()


class CSRMat:
    """AIJ (CSR) matrix with the subset of the ``PETSc.Mat`` API the reference's callers use.  Host
    arrays and the device-resident copy are both lazy: a matrix produced by :func:`AT_R_A` lives on
    the GPU and is only downloaded if somebody looks at ``rowptr/colind/val``."""

    def __init__(self, shape, rowptr=None, colind=None, val=None, device: DeviceMat | None = None):
        self._shape = (int(shape[0]), int(shape[1]))
        self._rowptr = None if rowptr is None else np.ascontiguousarray(rowptr)
        self._colind = None if colind is None else np.ascontiguousarray(colind)
        self._val = None if val is None else np.ascontiguousarray(val, dtype=np.float64)
        self._dev = device
        if self._rowptr is None and device is None:
            raise ValueError("CSRMat needs host arrays or a device matrix")

    # ---- construction helpers
    @classmethod
    def from_scipy(cls, S):
        S = S.tocsr()
        if not S.has_sorted_indices:
            S = S.sorted_indices()
        return cls(S.shape, S.indptr, S.indices, S.data)

    def to_scipy(self):
        import scipy.sparse as sp

        return sp.csr_matrix((self.val, self.colind, self.rowptr), shape=self._shape)

    # ---- lazy host / device views
    def _download(self):
        if self._rowptr is None:
            self._rowptr, self._colind, self._val = self._dev.to_csr(np.int32)

    @property
    def rowptr(self):
        self._download()
        return self._rowptr

    @property
    def colind(self):
        self._download()
        return self._colind

    @property
    def val(self):
        self._download()
        return self._val

    def device(self) -> DeviceMat:
        if self._dev is None:
            _ensure_init()
            self._dev = DeviceMat.from_csr(self._shape[0], self._shape[1], self._rowptr, self._colind, self._val)
        return self._dev

    def set_values(self, val):
        """New values on the same pattern (a fresh ``assemble`` in a Newton loop, reference common.py:432-435)."""
        self._download()
        self._val = np.ascontiguousarray(val, dtype=np.float64)
        if self._dev is not None:
            self._dev.update_values(self._val)

    # ---- PETSc.Mat API
    def getSize(self):
        return self._shape

    def getSizes(self):
        return ((self._shape[0], self._shape[0]), (self._shape[1], self._shape[1]))

    def getValuesCSR(self):
        return self.rowptr, self.colind, self.val

    def getComm(self):
        return None

    def assemble(self):
        return None

    def setOption(self, *args, **kwargs):
        return None

    def createVecLeft(self):
        return Vec(np.zeros(self._shape[0]))

    def createVecRight(self):
        return Vec(np.zeros(self._shape[1]))

    def getDiagonal(self, result=None):
        d = self.device().diagonal()
        if result is not None:
            arg2v(result).array[:] = d
            return result
        return Vec(d)

    def mult(self, x, y):
        yv = arg2v(y)
        yv.array[:] = self.device().spmv(arg2v(x).array)

    def multTranspose(self, x, y):
        yv = arg2v(y)
        yv.array[:] = self.device().spmv(arg2v(x).array, trans=True)

    def multAdd(self, x, v, y):
        """y = v + A x (MatMultAdd, reference common.py:364)."""
        yv = arg2v(y)
        yv.array[:] = self.device().spmv(arg2v(x).array, y=arg2v(v).array.copy(), alpha=1.0, beta=1.0)

    def transpose(self, out=None):
        """In-place like petsc4py's ``Mat.transpose()`` with no argument: self becomes its
        transpose and is returned."""
        T = self.device().transpose()
        self._dev = T
        self._shape = (self._shape[1], self._shape[0])
        self._rowptr = self._colind = self._val = None
        return self

    def matMult(self, other):
        raise NotImplementedError("use AT_R_A: the two MatMatMult calls of the reference are fused into one PtAP")

    def _replace_device(self, dev: DeviceMat):
        self._dev = dev
        self._shape = dev.shape
        self._rowptr = self._colind = self._val = None

    def zeroRows(self, rows, diag=1.0):
        """``Mat.zeroRows`` as trimNodes uses it (reference common.py:284,327): in place; a listed row keeps only
        (i, i) = diag (the entry is created if it was not stored, like PETSc without KEEP_NONZERO_PATTERN)."""
        self._replace_device(self.device().zero_rows(np.atleast_1d(np.asarray(rows, dtype=np.int64)), diag))

    def addDiagonal(self, d):
        """``A += A0`` with ``A0.setDiagonal(d)`` (reference common.py:243-249): in place, pattern = union with
        the full diagonal."""
        self._replace_device(self.device().add_diagonal(_vec_array(arg2v(d))))
        return self

    def norm(self):
        return float(np.linalg.norm(self.val))


# --------------------------------------------------------------------------------------------------
# type coercion (reference la_utils.py:28-70)
# --------------------------------------------------------------------------------------------------
def v2p(v):
    """dolfin PETScVector -> petsc4py Vec (reference la_utils.py:28-33)."""
    if HAVE_DOLFIN:
        return _dolfin.as_backend_type(v).vec()
    return arg2v(v)


def m2p(A):
    """dolfin PETScMatrix -> petsc4py Mat (reference la_utils.py:35-40)."""
    if HAVE_DOLFIN and not isinstance(A, (CSRMat,)) and not (HAVE_PETSC and isinstance(A, PETSc.Mat)):
        return _dolfin.as_backend_type(A).mat()
    return arg2m(A)


def arg2v(x):
    """dolfin Function / Vector / PETSc.Vec -> vector object (reference la_utils.py:42-56)."""
    if isinstance(x, Vec):
        return x
    if HAVE_PETSC and isinstance(x, PETSc.Vec):
        return x
    if HAVE_DOLFIN:
        try:
            if isinstance(x, _dolfin.function.function.Function):
                return _dolfin.as_backend_type(x.vector()).vec()
            if isinstance(x, (_dolfin.cpp.la.PETScVector, _dolfin.cpp.la.Vector)):
                return _dolfin.as_backend_type(x).vec()
        except AttributeError:  # pragma: no cover - a dolfin build without these classes
            pass
    if isinstance(x, np.ndarray):
        return Vec(x)
    raise TypeError("Type " + str(type(x)) + " is not supported yet.")


def arg2m(A):
    """dolfin Matrix / PETSc.Mat -> matrix object (reference la_utils.py:58-70)."""
    if isinstance(A, CSRMat):
        return A
    if HAVE_PETSC and isinstance(A, PETSc.Mat):
        return A
    if HAVE_DOLFIN:
        try:
            if isinstance(A, (_dolfin.cpp.la.PETScMatrix, _dolfin.cpp.la.Matrix)):
                return _dolfin.as_backend_type(A).mat()
        except AttributeError:  # pragma: no cover
            pass
    try:
        import scipy.sparse as sp

        if sp.issparse(A):
            return CSRMat.from_scipy(A)
    except ImportError:  # pragma: no cover
        pass
    raise TypeError("Type " + str(type(A)) + " is not supported yet.")


def zero_petsc_vec(num_el, comm=None):
    """New zero vector of global size ``num_el`` (reference la_utils.py:72-91)."""
    if HAVE_PETSC:
        v = PETSc.Vec().create(comm) if comm is not None else PETSc.Vec().create()
        v.setSizes(num_el)
        v.setUp()
        v.assemble()
        return v
    return Vec(np.zeros(int(num_el)))


def zero_petsc_mat(row, col, comm=None, row_loc=None, col_loc=None):
    """New empty AIJ matrix (reference la_utils.py:93-113)."""
    if HAVE_PETSC:
        A = PETSc.Mat(comm)
        A.createAIJ([(row_loc, row), (col_loc, col)], comm=comm)
        A.setUp()
        A.assemble()
        return A
    return CSRMat((row, col), np.zeros(int(row) + 1, dtype=np.int32), np.zeros(0, dtype=np.int32), np.zeros(0))


def updateU(u):
    """Ghost update of a dolfin Function (reference la_utils.py:116-125); a no-op for plain vectors."""
    if HAVE_DOLFIN and hasattr(u, "vector"):
        arg2v(u.vector()).assemble()
        arg2v(u.vector()).ghostUpdate()


# --------------------------------------------------------------------------------------------------
# helpers shared with common.py
# --------------------------------------------------------------------------------------------------
def _as_device(A) -> DeviceMat:
    """Device view of any supported matrix type (uploads PETSc / host matrices)."""
    _ensure_init()
    if isinstance(A, CSRMat):
        return A.device()
    if isinstance(A, DeviceMat):
        return A
    if HAVE_PETSC and isinstance(A, PETSc.Mat):
        rp, ci, v = A.getValuesCSR()
        n, m = A.getSize()
        return DeviceMat.from_csr(n, m, rp, ci, v)
    return arg2m(A).device()


def _torch():
    import torch  # device memory and copies only

    return torch


def _to_device(x):
    """float64 CUDA tensor with the value of a vector (no copy if it already lives on the GPU).  Uploads run
    on torch's stream and are complete on return, so the library's stream can consume them."""
    v = arg2v(x)
    if isinstance(v, Vec) and v.device_tensor() is not None:
        return v.device_tensor()
    _ensure_init()
    torch = _torch()
    dev = torch.device("cuda", max(_iife.current_device(), 0))
    t = torch.from_numpy(np.ascontiguousarray(_vec_array(v), dtype=np.float64)).to(dev)
    torch.cuda.current_stream(dev).synchronize()
    return t


def _vec_array(x) -> np.ndarray:
    v = arg2v(x)
    if isinstance(v, Vec):
        return v.array
    return v.getArray()  # PETSc.Vec: a view of the local array


def _wrap_mat_like(template, dev: DeviceMat):
    """Result matrix of the same kind as the inputs (PETSc.Mat when the inputs were PETSc)."""
    if HAVE_PETSC and isinstance(template, PETSc.Mat):
        n, m = dev.shape
        idt = np.dtype(PETSc.IntType)
        rp, ci, v = dev.to_csr(idt)
        return PETSc.Mat().createAIJ(size=(n, m), csr=(rp, ci, v), comm=template.getComm())
    return CSRMat(dev.shape, device=dev)


# --------------------------------------------------------------------------------------------------
# the three products (reference la_utils.py:129-182)
# --------------------------------------------------------------------------------------------------
def A_x_b(A, x, b):
    """Compute ``b = A x`` (reference la_utils.py:129-141)."""
    Am, xv, bv = arg2m(A), arg2v(x), arg2v(b)
    y = _as_device(Am).spmv(_vec_array(xv))
    _vec_array(bv)[:] = y
    return None


def AT_x(A, x):
    """Compute ``b = A^T x`` into a new vector of size ``ncols`` (reference la_utils.py:143-163)."""
    A_m = arg2m(A)
    x_v = arg2v(x)
    dev = _as_device(A_m)
    row, col = dev.shape
    if not HAVE_PETSC and row > 0 and col > 0:
        # the result stays on the GPU (lazy Vec): the next consumer is solveKSP
        x_d = _to_device(x_v)
        y_d = _torch().empty(col, dtype=_torch().float64, device=x_d.device)
        dev.spmv(x_d, y=y_d, trans=True)
        _iife.sync()
        return Vec(device=y_d)
    y = dev.spmv(_vec_array(x_v), trans=True)
    b = zero_petsc_vec(col, comm=A_m.getComm() if hasattr(A_m, "getComm") else None)
    _vec_array(b)[:] = y
    return b


def AT_R_A(A, R):
    """Compute ``A^T R A`` — called as ``AT_R_A(M, A_f)`` (reference la_utils.py:165-182, common.py:160).

    The reference does transpose / MatMatMult / transpose back / MatMatMult with both symbolic phases
    redone on every call; here it is one PtAP whose symbolic plan is cached by the pattern fingerprints
    of (A, R), so repeated calls with new values (Newton steps, time steps) run the numeric phase only.
    Returns a NEW matrix; ``A`` and ``R`` are unchanged.
    """
    dA, dR = _as_device(arg2m(A)), _as_device(arg2m(R))
    dC, _cached = _iife.ptap(dA, dR)
    return _wrap_mat_like(arg2m(A), dC)
