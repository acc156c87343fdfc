from . import FromOriginalModelMixin  # noqa: F401
