from .vq_module import VQModule as VQ  # noqa: F401  (reference: src/networks/vq/__init__.py:1)
from .vq_module import VQModule  # noqa: F401
