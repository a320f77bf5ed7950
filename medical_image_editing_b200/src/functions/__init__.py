from .vq_function import VQFunction, vq_lookup  # noqa: F401
