"""The two helpers of the reference's `utils` package that the quantiser uses
(reference: src/utils/__init__.py:109-114)."""
import os


def get_world_size() -> int:
    return int(os.environ.get("WORLD_SIZE", 1))


def is_distributed() -> bool:
    return get_world_size() > 1
