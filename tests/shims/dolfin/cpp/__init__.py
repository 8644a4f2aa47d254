from . import la, log  # noqa: F401
