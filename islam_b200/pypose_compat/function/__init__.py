from . import geometry      # noqa: F401
