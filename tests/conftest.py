import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "interpolation-based-immersed-fea_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    # the template numeric PtAP is switched off below 32 768 rows by default (small problems are launch bound and the
    # host-side program compilation would dominate their cold time); the parity tests are small, so they lower the bar
    # to keep that kernel on the tested path (tests/test_gpu_parity.py::test_ptap_template_kernel checks the default too)
    os.environ.setdefault("IIFE_TPL_MIN_PROBLEM", "0")


@pytest.fixture(scope="session")
def iife():
    """The product library bound to cuda:0.  Fails (does not skip) when the CUDA path is unavailable:
    a GPU test must never pass on a fallback."""
    import iife_b200

    iife_b200.init(0)
    return iife_b200


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.build()
    return O


def rand_csr(rng, n_rows, n_cols, mean_len, max_len=None, empty_frac=0.0, dtype=np.int32):
    """Random CSR with sorted unique columns; some rows empty; lengths ~ Poisson(mean_len)."""
    max_len = max_len or n_cols
    lens = np.minimum(rng.poisson(mean_len, n_rows), min(max_len, n_cols))
    lens[rng.random(n_rows) < empty_frac] = 0
    rowptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(lens, out=rowptr[1:])
    colind = np.empty(int(rowptr[-1]), dtype=np.int64)
    for i in range(n_rows):
        colind[rowptr[i]:rowptr[i + 1]] = np.sort(rng.choice(n_cols, lens[i], replace=False))
    val = rng.standard_normal(int(rowptr[-1]))
    return rowptr.astype(dtype), colind.astype(dtype), val
