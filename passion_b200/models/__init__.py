from . import rfnet  # noqa: F401
