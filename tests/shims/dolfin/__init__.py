"""Fake dolfin (test shim, see tests/shims/README.md): an EXPLICIT list of the legacy-FEniCS names the reference's
modules and demos use, as inert stubs, plus working stand-ins for the few objects the extraction path touches
(PETScVector / PETScMatrix wrappers, as_backend_type, MPI, parameters, PETScKrylovSolver).  No catch-all
__getattr__: a name that is missing here is reported as missing by tests/test_overlay.py."""
import math as _math
import types as _types

from . import cpp  # noqa: F401
from .cpp.la import PETScMatrix, PETScVector  # noqa: F401

pi = _math.pi
DOLFIN_EPS = 3.0e-16


class _LogLevel:
    INFO, WARNING, ERROR, DEBUG, PROGRESS = 20, 30, 40, 10, 16


LogLevel = _LogLevel
parameters = {"std_out_all_processes": True, "ghost_mode": "none", "form_compiler": {}}


def set_log_level(level):
    return None


def log(level, msg):
    print(msg)


class _MPI:
    comm_world = object()
    comm_self = object()

    @staticmethod
    def rank(comm):
        return 0

    @staticmethod
    def size(comm):
        return 1

    @staticmethod
    def barrier(comm):
        return None


MPI = _MPI


class Function:
    """dolfin.function.function.Function stand-in: holds a PETScVector."""

    def __init__(self, V=None, vec=None):
        self._V = V
        self._vec = vec

    def vector(self):
        return self._vec

    def function_space(self):
        return self._V


function = _types.SimpleNamespace(function=_types.SimpleNamespace(Function=Function))


def as_backend_type(x):
    return x


class PETScKrylovSolver:
    def __init__(self, ksp=None):
        self._ksp = ksp
        self.parameters = {}

    def solve(self, x, b):
        if self.parameters.get("nonzero_initial_guess"):
            self._ksp.setInitialGuessNonzero(True)
        self._ksp.solve(b.vec(), x.vec())
        return self._ksp.getIterationNumber()


def _stub(name):
    def f(*args, **kwargs):
        raise NotImplementedError(f"dolfin.{name} is a stub of the test shim")

    f.__name__ = name
    return f


# inert stubs: mesh / function-space / UFL names used by the reference's common.py and its demos
for _n in ("XDMFFile", "HDF5File", "File", "Mesh", "MeshFunction", "MeshEditor", "RectangleMesh", "BoxMesh", "UnitSquareMesh",
           "UnitCubeMesh", "Point", "facets", "cells", "vertices", "edges", "Measure", "FunctionSpace", "VectorFunctionSpace",
           "TensorFunctionSpace", "FiniteElement", "VectorElement", "MixedElement", "TestFunction", "TrialFunction",
           "TestFunctions", "TrialFunctions", "SpatialCoordinate", "Constant", "Expression", "UserExpression", "CellDiameter",
           "CellVolume", "FacetNormal", "FacetArea", "Circumradius", "inner", "outer", "dot", "cross", "grad", "nabla_grad",
           "div", "curl", "sym", "skew", "tr", "det", "inv", "dev", "sqrt", "sin", "cos", "tan", "atan", "atan_2", "asin", "acos", "sinh", "cosh", "tanh", "exp", "ln", "sign", "erf", "jump", "avg",
           "Identity", "as_vector", "as_tensor", "as_matrix", "conditional", "gt", "lt", "ge", "le", "eq", "ne", "split",
           "assemble", "assemble_system", "derivative", "action", "adjoint", "lhs", "rhs", "system", "replace", "project",
           "interpolate", "errornorm", "norm", "solve", "DirichletBC", "SubDomain", "near", "between", "plot", "info",
           "PETScDMCollection", "PETScOptions", "LUSolver", "KrylovSolver", "NonlinearVariationalProblem",
           "NonlinearVariationalSolver", "LinearVariationalProblem", "LinearVariationalSolver", "TimingType", "timings",
           "Timer", "list_timings", "diff", "variable", "dx", "ds", "dS", "dP", "VectorConstant", "CellType", "refine",
           "BoundingBoxTree", "Cell", "Facet", "Vertex", "vertex_to_dof_map", "dof_to_vertex_map", "as_ufl", "elem_mult",
           "transpose", "perp", "cofac", "variable", "MPI_Comm"):
    if _n not in globals():
        globals()[_n] = _stub(_n)
del _n
