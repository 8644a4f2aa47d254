"""Fake petsc4py.PETSc on numpy / scipy (test shim, see tests/shims/README.md): the subset of Mat / Vec / KSP / PC /
Options the reference's la_utils.py and common.py touch."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

IntType = np.int32
ScalarType = np.float64
DECIDE = -1


class _Comm:
    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def getRank(self):
        return 0

    def getSize(self):
        return 1


COMM_WORLD = _Comm()
COMM_SELF = _Comm()


class InsertMode:
    INSERT = 1
    ADD = 2


class ScatterMode:
    FORWARD = 0
    REVERSE = 1


class Options(dict):
    def __init__(self, prefix=""):
        super().__init__()
        self.prefix = prefix


class Vec:
    Type = type("Type", (), {"STANDARD": "standard", "MPI": "mpi", "SEQ": "seq"})

    def __init__(self, comm=None):
        self._a = np.zeros(0)
        self._comm = comm or COMM_WORLD

    # construction
    def create(self, comm=None):
        self._comm = comm or COMM_WORLD
        return self

    def createWithArray(self, array, size=None, comm=None):
        self._a = np.asarray(array, dtype=np.float64).reshape(-1)  # a VIEW: the solve writes into the caller's array
        return self

    def createSeq(self, n, comm=None):
        self._a = np.zeros(int(n))
        return self

    def setSizes(self, size, bsize=None):
        n = size[1] if isinstance(size, (tuple, list)) else size
        self._a = np.zeros(int(n))

    def setUp(self):
        return self

    def setType(self, t):
        return None

    def setFromOptions(self):
        return None

    def assemble(self):
        return None

    assemblyBegin = assemblyEnd = assemble

    def ghostUpdate(self, *a, **k):
        return None

    # access
    def getArray(self, readonly=False):
        return self._a

    def setArray(self, a):
        self._a[:] = a

    array = property(getArray, setArray)

    def getSize(self):
        return int(self._a.size)

    def getLocalSize(self):
        return int(self._a.size)

    def getSizes(self):
        return (int(self._a.size), int(self._a.size))

    def getOwnershipRange(self):
        return (0, int(self._a.size))

    def getComm(self):
        return self._comm

    def getValue(self, i):
        return float(self._a[i])

    def setValue(self, i, v, addv=None):
        self._a[i] = v

    def set(self, alpha):
        self._a[:] = alpha

    def zeroEntries(self):
        self._a[:] = 0.0

    def norm(self, norm_type=None):
        return float(np.linalg.norm(self._a))

    def sum(self):
        return float(self._a.sum())

    def dot(self, other):
        return float(self._a @ other._a)

    def copy(self, result=None):
        if result is not None:
            result._a[:] = self._a
            return result
        return Vec().createWithArray(self._a.copy())

    def duplicate(self):
        return Vec().createWithArray(np.zeros_like(self._a))

    def axpy(self, alpha, x):
        self._a += alpha * x._a

    def scale(self, alpha):
        self._a *= alpha

    def __neg__(self):
        return Vec().createWithArray(-self._a)

    def __mul__(self, alpha):
        return Vec().createWithArray(self._a * alpha)

    __rmul__ = __mul__

    def __iadd__(self, other):
        self._a += other._a
        return self

    def __isub__(self, other):
        self._a -= other._a
        return self

    def destroy(self):
        return None


class Mat:
    Type = type("Type", (), {"AIJ": "aij", "MPIAIJ": "mpiaij", "SEQAIJ": "seqaij"})
    Option = type("Option", (), {"NEW_NONZERO_ALLOCATION_ERR": 1, "KEEP_NONZERO_PATTERN": 2, "NEW_NONZERO_LOCATIONS": 3})

    def __init__(self, comm=None):
        self._S = sp.csr_matrix((0, 0))
        self._comm = comm or COMM_WORLD
        self._pending = None

    @staticmethod
    def _gsize(x):
        return int(x[1]) if isinstance(x, (tuple, list)) else int(x)

    def create(self, comm=None):
        self._comm = comm or COMM_WORLD
        return self

    def createAIJ(self, size, bsize=None, nnz=None, csr=None, comm=None):
        n, m = self._gsize(size[0]), self._gsize(size[1])
        if csr is not None:
            rp, ci, v = csr
            self._S = sp.csr_matrix((np.array(v, dtype=np.float64), np.array(ci), np.array(rp)), shape=(n, m))
        else:
            self._S = sp.csr_matrix((n, m))
        self._comm = comm or COMM_WORLD
        return self

    def setSizes(self, size, bsize=None):
        self._S = sp.csr_matrix((self._gsize(size[0]), self._gsize(size[1])))

    def setType(self, t):
        return None

    def setUp(self):
        return self

    def setFromOptions(self):
        return None

    def setOption(self, *a, **k):
        return None

    def setPreallocationNNZ(self, nnz):
        return None

    def assemble(self):
        if self._pending:
            L = self._S.tolil()
            for (i, j), v in self._pending.items():
                L[i, j] = v
            self._S = L.tocsr()
            self._pending = None
        self._S.sort_indices()

    assemblyBegin = assemblyEnd = assemble

    def setValue(self, i, j, v, addv=None):
        if self._pending is None:
            self._pending = {}
        self._pending[(int(i), int(j))] = float(v)  # INSERT: the last value wins

    def getValue(self, i, j):
        return float(self._S[i, j])

    def getValuesCSR(self):
        self.assemble()
        return self._S.indptr.astype(IntType), self._S.indices.astype(IntType), self._S.data

    def getSize(self):
        return self._S.shape

    def getLocalSize(self):
        return self._S.shape

    def getSizes(self):
        n, m = self._S.shape
        return ((n, n), (m, m))

    def getOwnershipRange(self):
        return (0, self._S.shape[0])

    def getComm(self):
        return self._comm

    def createVecLeft(self):
        return Vec().createWithArray(np.zeros(self._S.shape[0]))

    def createVecRight(self):
        return Vec().createWithArray(np.zeros(self._S.shape[1]))

    createVecs = lambda self: (self.createVecRight(), self.createVecLeft())  # noqa: E731

    def getDiagonal(self, result=None):
        d = self._S.diagonal()
        if result is not None:
            result._a[:] = d
            return result
        return Vec().createWithArray(d)

    def setDiagonal(self, diag, addv=None):
        n = min(self._S.shape)
        L = self._S.tolil()
        L.setdiag(diag._a[:n])
        self._S = L.tocsr()

    def zeroRows(self, rows, diag=1.0, x=None, b=None):
        rows = np.asarray(rows, dtype=np.int64)
        L = self._S.tolil()
        for r in rows:
            L.rows[r] = []
            L.data[r] = []
            if diag != 0.0 and r < self._S.shape[1]:
                L[r, r] = diag
        self._S = L.tocsr()

    def mult(self, x, y):
        y._a[:] = self._S @ x._a

    def multTranspose(self, x, y):
        y._a[:] = self._S.T @ x._a

    def multAdd(self, x, v, y):
        y._a[:] = v._a + self._S @ x._a

    def transpose(self, out=None):
        T = self._S.T.tocsr()
        T.sort_indices()
        if out is None:
            self._S = T
            return self
        out._S = T
        return out

    def matMult(self, other, result=None, fill=None):
        C = Mat()
        C._S = (self._S @ other._S).tocsr()
        C._S.sort_indices()
        return C

    def duplicate(self, copy=False):
        C = Mat()
        C._S = self._S.copy()
        return C

    def copy(self, result=None, structure=None):
        return self.duplicate(True)

    def axpy(self, alpha, X, structure=None):
        self._S = (self._S + alpha * X._S).tocsr()

    def __iadd__(self, other):
        self._S = (self._S + other._S).tocsr()
        self._S.sort_indices()
        return self

    def norm(self, norm_type=None):
        return float(np.sqrt((self._S.data ** 2).sum()))

    def destroy(self):
        return None


class PC:
    def __init__(self):
        self.type = "none"
        self.solver = None

    def setType(self, t):
        self.type = str(t)

    def getType(self):
        return self.type

    def setFactorSolverType(self, s):
        self.solver = s

    def setASMOverlap(self, n):
        return None

    def getASMSubKSP(self):
        return [KSP()]

    def setHYPREType(self, t):
        return None

    def setFromOptions(self):
        return None


class KSP:
    Type = type("Type", (), {"FGMRES": "fgmres", "GMRES": "gmres", "GCR": "gcr", "CG": "cg", "PREONLY": "preonly"})

    def __init__(self):
        self.type = "gmres"
        self.pc = PC()
        self.A = None
        self.rtol, self.atol, self.max_it = 1e-5, 1e-50, 10000
        self.its = 0
        self.guess_nonzero = False

    def create(self, comm=None):
        return self

    def setTolerances(self, rtol=None, atol=None, divtol=None, max_it=None):
        if rtol is not None:
            self.rtol = rtol
        if atol is not None:
            self.atol = atol
        if max_it is not None:
            self.max_it = max_it

    def setType(self, t):
        self.type = str(t)

    def getType(self):
        return self.type

    def setOperators(self, A, P=None):
        self.A = A

    def getPC(self):
        return self.pc

    def setUp(self):
        return None

    def setFromOptions(self):
        return None

    def setGMRESRestart(self, m):
        self.restart = m

    def setInitialGuessNonzero(self, flag):
        self.guess_nonzero = bool(flag)

    def setComputeSingularValues(self, flag):
        return None

    def setConvergenceHistory(self, *a, **k):
        return None

    def getConvergenceHistory(self):
        return np.zeros(0)

    def getIterationNumber(self):
        return self.its

    def getConvergedReason(self):
        return 2

    def solve(self, b, x):
        S = self.A._S.tocsc()
        if self.type == "preonly" or self.pc.type == "lu":
            x._a[:] = spla.splu(S).solve(b._a)
            self.its = 1
            return
        # any Krylov type of the shim: a direct solve stands in (the shim tests the PLUMBING of the delegated branches)
        x._a[:] = spla.spsolve(S, b._a)
        self.its = 1
