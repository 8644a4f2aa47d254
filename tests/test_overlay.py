"""CPU tests of the drop-in boundary as an OVERLAY of the reference package: with petsc4py / dolfin
importable (here: the duck-typed shims of tests/shims) the mirror re-exports dolfin's names, defines the reference's
module constants, ships profile_utils, keeps the MUMPS / GCR / ASM branches of solveKSP on PETSc, and falls through to
the reference's own common.py for the dolfin-only helpers it does not define.  Runs in a subprocess so that the
shims never leak into the other tests' imports.  No GPU: nothing here calls a compute entry point."""
import ast
import builtins
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "interpolation-based-immersed-fea_b200")
SHIMS = os.path.join(ROOT, "tests", "shims")
REFERENCE = "/root/reference"


def run_py(code, extra_env=None, shims=True):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join(([SHIMS] if shims else []) + [PKG, ROOT])
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    return out.stdout


def test_without_petsc_the_delegated_branches_raise():
    out = run_py("""
        import numpy as np
        from InterpolationBasedImmersedFEA.common import *
        from InterpolationBasedImmersedFEA.profile_utils import profile_separate
        assert not HAVE_PETSC and not HAVE_DOLFIN and mpirank == 0 and mpisize == 1 and worldcomm is None
        A = CSRMat((2, 2), np.array([0, 1, 2]), np.array([0, 1]), np.array([2.0, 4.0]))
        for kw in (dict(method='mumps'), dict(PC='ASM'), dict(PC='ICC'), dict(PC='ILU'), dict(PC='ILUT')):
            try:
                solveKSP(A, Vec(np.ones(2)), Vec(np.zeros(2)), monitor=False, **kw)
            except NotImplementedError as e:
                assert 'petsc4py' in str(e)
            else:
                raise SystemExit('no error for %r' % kw)

        @profile_separate()
        def f(x):
            return x + 1
        assert f(1) == 2
        print('ok')
    """, shims=False)
    assert "ok" in out


def test_delegated_branches_run_on_petsc_when_it_imports():
    """solveKSP(method='mumps' | 'gcr', PC='ASM' | ...) follows the reference's PETSc configuration (reference
    common.py:525-551, 576-616) on whatever petsc4py is importable: here the shim, whose KSP solves with scipy."""
    out = run_py("""
        import numpy as np, scipy.sparse as sp
        from InterpolationBasedImmersedFEA.common import *
        from petsc4py import PETSc
        assert HAVE_PETSC and HAVE_DOLFIN and PETSC4PY_MATRIX is PETSc.Mat and DOLFIN_PETSCMATRIX is cpp.la.PETScMatrix
        rng = np.random.default_rng(0)
        n = 30
        S = sp.random(n, n, density=0.2, random_state=1, format='csr') + sp.eye(n) * 5.0
        S = S.tocsr(); S.sort_indices()
        x_ref = rng.standard_normal(n)
        b = S @ x_ref
        for kw in (dict(method='mumps'), dict(method='gcr', PC='ASM'), dict(method='gmres', PC='ASM'), dict(method='cg', PC='ICC'),
                   dict(PC='ILU'), dict(PC='ILUT')):
            A = PETSc.Mat().createAIJ(size=S.shape, csr=(S.indptr, S.indices, S.data))
            bv = PETSc.Vec().createWithArray(b.copy())
            u = A.createVecLeft()
            assert solveKSP(A, bv, u, monitor=False, **kw) is None
            assert np.allclose(u.getArray(), x_ref, rtol=1e-10), kw
            # the mirror's light objects are accepted by the PETSc branches too
            A2 = CSRMat(S.shape, S.indptr, S.indices, S.data)
            u2 = Vec(np.zeros(n))
            solveKSP(A2, Vec(b.copy()), u2, monitor=False, **kw)
            assert np.allclose(u2.array, x_ref, rtol=1e-10), kw
        # dolfin wrappers are unwrapped as in the reference (la_utils.py:28-70)
        A = PETSc.Mat().createAIJ(size=S.shape, csr=(S.indptr, S.indices, S.data))
        assert m2p(PETScMatrix(A)) is A and v2p(PETScVector(A.createVecLeft())).getSize() == n
        f = Function(None, PETScVector(PETSc.Vec().createWithArray(np.arange(3.0))))
        assert np.array_equal(arg2v(f).getArray(), [0.0, 1.0, 2.0])
        for bad in (3, 'x'):
            try:
                arg2v(bad)
            except TypeError as e:
                assert 'is not supported yet.' in str(e)
        print('ok')
    """)
    assert "ok" in out


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference tree is only present in the build container")
def test_names_of_the_reference_fall_through_and_demo_names_resolve():
    """`from InterpolationBasedImmersedFEA.common import *` + `...profile_utils import profile_separate` +
    `...la_utils import *` (demos/poisson.py:14-16) must deliver every free name the demo uses: dolfin's (re-exported),
    the mirror's own, and the reference-only helpers (generateUnfittedMesh, mixedScalarSpace, cellMetric, L2Norm,
    convertDOFs* ...), which come from the reference's own common.py loaded from IIFE_REFERENCE_PATH."""
    out = run_py("""
        import ast, builtins, sys
        import InterpolationBasedImmersedFEA.common as C
        ns = {}
        exec('from InterpolationBasedImmersedFEA.common import *\\n'
             'from InterpolationBasedImmersedFEA.profile_utils import profile_separate\\n'
             'from InterpolationBasedImmersedFEA.la_utils import *', ns)
        # fall-through: defined by the reference only, executed on top of the mirror's la_utils
        for name in ('generateUnfittedMesh', 'mixedScalarSpace', 'averageCellDiagonal', 'cellMetric', 'L2Norm',
                     'convertDOFsk1', 'convertDOFs2Dk2', 'convertDOFs3Dk2', 'EXTRACTION_DATA_FILE'):
            assert name in ns, name
            assert ns[name].__module__ == '_iife_reference_common' if callable(ns[name]) else True, name
        # the hot-path names are the mirror's, not the reference's
        for name in ('AT_R_A', 'AT_x', 'A_x_b', 'solveKSP', 'assembleLinearSystemBackground', 'transferToForeground',
                     'trimNodes', 'readExOp', 'solveNonlinear', 'solveNewtonsLinear', 'L2Project', 'estimateConditionNumber'):
            assert ns[name].__module__.startswith('InterpolationBasedImmersedFEA.'), (name, ns[name].__module__)
        ref = sys.modules['_iife_reference_common']
        assert ref.AT_R_A is C.AT_R_A  # inside the reference module la_utils resolved to the mirror
        missing = {}
        for demo in ('poisson', 'linear_elasticity', 'biharmonic', 'tg_vortex'):
            tree = ast.parse(open('/root/reference/demos/%s.py' % demo).read())
            defined = set(dir(builtins))
            for node in ast.walk(tree):
                if isinstance(node, (ast.FunctionDef, ast.ClassDef)):
                    defined.add(node.name)
                    defined.update(a.arg for a in node.args.args) if isinstance(node, ast.FunctionDef) else None
                elif isinstance(node, ast.Name) and isinstance(node.ctx, (ast.Store, ast.Del)):
                    defined.add(node.id)
                elif isinstance(node, (ast.Import, ast.ImportFrom)):
                    defined.update((a.asname or a.name).split('.')[0] for a in node.names)
                elif isinstance(node, ast.arg):
                    defined.add(node.arg)
            used = {n.id for n in ast.walk(tree) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load)}
            miss = sorted(n for n in used - defined if n not in ns)
            if miss:
                missing[demo] = miss
        assert not missing, missing
        print('ok')
    """, extra_env={"IIFE_REFERENCE_PATH": REFERENCE})
    assert "ok" in out
