"""iife_b200 — Python face of libiife.so, the B200-native extraction linear algebra
(A_b = M^T A_f M, b_b = M^T b_f, Jacobi CG / FGMRES) behind the reference's la_utils / solveKSP API."""
from .core import (DeviceMat, PtapPlan, KSPInfo, init, is_initialised, current_device, finalize, device_count, set_stream, sync,
                   device_bytes, launch_count, ptap, plan_cache_clear, ksp_solve, ksp_hessenberg, synth_cube, KSP_CG, KSP_FGMRES, KSP_GCR,
                   PC_NONE, PC_JACOBI, MEM_HOST, MEM_DEVICE, REASONS)
from ._lib import IifeError, LIB_PATH

__all__ = ["DeviceMat", "PtapPlan", "KSPInfo", "init", "is_initialised", "current_device", "finalize", "device_count", "set_stream",
           "sync", "device_bytes", "launch_count", "ptap", "plan_cache_clear", "ksp_solve", "ksp_hessenberg", "synth_cube", "KSP_CG",
           "KSP_FGMRES", "KSP_GCR", "PC_NONE", "PC_JACOBI", "MEM_HOST", "MEM_DEVICE", "REASONS", "IifeError", "LIB_PATH"]
