"""Fake h5py (test shim): the reference's common.py imports it at module level and never uses it on the hot path."""
