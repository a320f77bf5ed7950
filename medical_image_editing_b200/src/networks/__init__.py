from .vq import VQ  # noqa: F401
