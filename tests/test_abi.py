"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
include/iife.h declares, binds them in iife_b200._lib, and refuses to compute without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "iife.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(iife_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from iife_b200 import _lib

    names = declared_symbols()
    assert len(names) >= 40
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/iife.h but not exported by libiife.so"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype in iife_b200/_lib.py"
    assert set(_lib.PROTOTYPES) <= set(names), set(_lib.PROTOTYPES) - set(names)
    assert _lib.lib.iife_version() == 100


def test_no_cpu_fallback_without_device():
    import iife_b200 as I

    if I.device_count() > 0:
        pytest.skip("a GPU is visible: the no-device behaviour is checked on the CPU box")
    with pytest.raises(I.IifeError) as e:
        I.init(0)
    assert e.value.code == 5  # IIFE_ERR_NO_DEVICE
    # compute entry points refuse to run uninitialised
    h = ctypes.c_void_p(0)
    rp = np.zeros(2, dtype=np.int32)
    rc = I._lib.lib.iife_mat_create_csr(1, 1, rp.ctypes.data_as(ctypes.c_void_p), None, None, 4, 0, ctypes.byref(h))
    assert rc == 5 and b"iife_init" in I._lib.lib.iife_last_error()
    with pytest.raises(I.IifeError):
        I.DeviceMat.from_csr(1, 1, np.array([0, 1]), np.array([0]), np.array([1.0]))


def test_product_does_not_import_the_oracle():
    """the oracle is test infrastructure: nothing under the package may import it"""
    pkg = os.path.join(ROOT, "interpolation-based-immersed-fea_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "iife_oracle" not in txt, f
