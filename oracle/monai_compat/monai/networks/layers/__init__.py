from .factories import Act, Norm, Conv  # noqa: F401
