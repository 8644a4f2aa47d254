"""dolfin.cpp.la stand-ins: thin wrappers around (fake or real) petsc4py objects."""


class Vector:
    def __init__(self, v=None):
        self._v = v

    def vec(self):
        return self._v


class PETScVector(Vector):
    pass


class Matrix:
    def __init__(self, m=None):
        self._m = m

    def mat(self):
        return self._m


class PETScMatrix(Matrix):
    pass
