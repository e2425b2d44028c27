from . import cjsonwriter  # noqa: F401
