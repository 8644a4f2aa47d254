"""ctypes binding of libiife.so (include/iife.h).  There is no fallback: if the shared library is
missing the import fails, and every compute call fails unless ``iife_init`` bound a CUDA device."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# IIFE_LIB selects another build of the same library (tuning variants: `make BUILD=... LIB=../lib/tuned/libiife.so
# EXTRA_NVCCFLAGS=...`); it must still be called libiife.so
LIB_PATH = os.path.normpath(os.environ.get("IIFE_LIB") or os.path.join(_HERE, "..", "lib", "libiife.so"))

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_dbl = ctypes.c_double
c_vp = ctypes.c_void_p
P = ctypes.POINTER


class KspResult(ctypes.Structure):
    _fields_ = [("iterations", c_i64), ("reason", ctypes.c_int32), ("_pad", ctypes.c_int32), ("rnorm", c_dbl),
                ("rnorm0", c_dbl)]


# name -> (restype, argtypes); must list every symbol include/iife.h declares (tests check it)
PROTOTYPES = {
    "iife_version": (c_int, []),
    "iife_last_error": (ctypes.c_char_p, []),
    "iife_init": (c_int, [c_int]),
    "iife_finalize": (c_int, []),
    "iife_device_count": (c_int, [P(c_int)]),
    "iife_set_stream": (c_int, [c_vp]),
    "iife_sync": (c_int, []),
    "iife_device_bytes": (c_int, [P(c_i64)]),
    "iife_launch_count": (c_int, [P(c_i64), c_int]),
    "iife_mat_create_csr": (c_int, [c_i64, c_i64, c_vp, c_vp, c_vp, c_int, c_int, P(c_vp)]),
    "iife_mat_create_csr_ex": (c_int, [c_i64, c_i64, c_vp, c_vp, c_vp, c_int, c_int, c_int, P(c_vp)]),
    "iife_mat_update_values": (c_int, [c_vp, c_vp, c_int]),
    "iife_mat_get_info": (c_int, [c_vp, P(c_i64), P(c_i64), P(c_i64)]),
    "iife_mat_get_csr": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int]),
    "iife_mat_device_ptrs": (c_int, [c_vp, P(c_vp), P(c_vp), P(c_vp)]),
    "iife_mat_touch": (c_int, [c_vp]),
    "iife_mat_fingerprint": (c_int, [c_vp, P(ctypes.c_uint64)]),
    "iife_mat_transpose": (c_int, [c_vp, P(c_vp)]),
    "iife_mat_get_diagonal": (c_int, [c_vp, c_vp, c_int]),
    "iife_mat_zero_rows": (c_int, [c_vp, c_vp, ctypes.c_int64, c_int, c_dbl, c_int, P(c_vp)]),
    "iife_mat_add_diagonal": (c_int, [c_vp, c_vp, c_int, P(c_vp)]),
    "iife_mat_destroy": (c_int, [c_vp]),
    "iife_spmv": (c_int, [c_vp, c_int, c_dbl, c_vp, c_dbl, c_vp, c_int]),
    "iife_ptap_symbolic": (c_int, [c_vp, c_vp, P(c_vp)]),
    "iife_plan_matches": (c_int, [c_vp, c_vp, c_vp, P(c_int)]),
    "iife_plan_get_info": (c_int, [c_vp, P(c_i64), P(c_i64), P(c_i64)]),
    "iife_rap_symbolic": (c_int, [c_vp, c_vp, c_vp, P(c_vp)]),
    "iife_rap_numeric": (c_int, [c_vp, c_vp, c_vp, c_vp, P(c_vp)]),
    "iife_plan_bin_counts": (c_int, [c_vp, P(c_i64)]),
    "iife_plan_bin_counts_n": (c_int, [c_vp, P(c_i64), c_int]),
    "iife_plan_check": (c_int, [c_vp]),
    "iife_plan_tpl_info": (c_int, [c_vp, P(c_i64), P(c_i64), P(c_i64), c_vp]),
    "iife_tpl_emulate_row": (c_int, [c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp]),
    "iife_ptap_numeric": (c_int, [c_vp, c_vp, c_vp, P(c_vp)]),
    "iife_plan_destroy": (c_int, [c_vp]),
    "iife_ptap": (c_int, [c_vp, c_vp, P(c_vp), P(c_int)]),
    "iife_plan_cache_clear": (c_int, []),
    "iife_ksp_solve": (c_int, [c_vp, c_int, c_int, c_dbl, c_dbl, c_dbl, c_i64, c_int, c_vp, c_vp, c_int, c_vp,
                               P(KspResult), c_vp, c_i64]),
    "iife_comm_unique_id": (c_int, [c_vp]),
    "iife_comm_init": (c_int, [c_int, c_int, c_vp]),
    "iife_comm_finalize": (c_int, []),
    "iife_comm_info": (c_int, [P(c_int), P(c_int)]),
    "iife_halo_create": (c_int, [c_i64, c_i64, c_vp, c_vp, c_vp, P(c_vp)]),
    "iife_halo_exchange": (c_int, [c_vp, c_vp]),
    "iife_halo_destroy": (c_int, [c_vp]),
    "iife_halo_p2p_export": (c_int, [c_vp, c_vp]),
    "iife_halo_p2p_attach": (c_int, [c_vp, c_vp, c_vp]),
    "iife_halo_p2p_error": (c_int, [c_vp, P(c_int)]),
    "iife_spmv_dist": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "iife_allreduce_sum": (c_int, [c_vp, c_i64]),
    "iife_alltoallv_bytes": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "iife_ksp_solve_hessenberg": (c_int, [c_vp, c_int, c_dbl, c_dbl, c_dbl, c_i64, c_int, c_vp, c_vp, c_int, P(KspResult),
                                           c_vp, c_i64, P(c_i64)]),
    "iife_ksp_solve_dist": (c_int, [c_vp, c_vp, c_int, c_int, c_dbl, c_dbl, c_dbl, c_i64, c_int, c_vp, c_vp,
                                    P(KspResult), c_vp, c_i64]),
    "iife_synth_cube_counts": (c_int, [c_i64, c_i64, c_i64, P(c_i64), P(c_i64)]),
    "iife_synth_cube_build": (c_int, [c_i64, c_i64, c_i64, c_vp, c_vp, P(c_vp), P(c_vp), c_vp]),
}


class IifeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libiife error {code}: {msg}")
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make -C interpolation-based-immersed-fea_b200/csrc` "
            "(or python -c 'import __graft_entry__ as g; g.build()').  There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int) -> None:
    if rc != 0:
        msg = lib.iife_last_error()
        raise IifeError(rc, msg.decode("utf-8", "replace") if msg else "")
