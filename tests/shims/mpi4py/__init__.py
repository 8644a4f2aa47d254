"""Fake mpi4py (test shim)."""
