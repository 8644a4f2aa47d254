"""Fake petsc4py (test shim, see tests/shims/README.md)."""
__version__ = "0.0-shim"


def init(*args, **kwargs):
    return None


def get_config():
    return {}
