from . import LocallyConnected2D  # noqa: F401
