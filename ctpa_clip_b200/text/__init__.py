from .bert import NativeBert, supports  # noqa: F401
