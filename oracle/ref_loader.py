"""Import the UNMODIFIED reference modules from /root/reference -- TEST INFRASTRUCTURE ONLY.

The reference's package `__init__`s pull in kornia / nibabel / pytorch_lightning /
kmeans_pytorch, none of which exist in this image.  We pre-seed
`sys.modules` with:
  * `utils`          -> a stub exposing get_world_size / is_distributed
                        (restating utils/__init__.py:109-114),
  * `networks`       -> an empty package whose __path__ points at the reference
                        directory (so `networks/__init__.py` is skipped but submodules
                        import normally),
  * `kmeans_pytorch` -> a dummy.
No reference file is modified.  `/root/reference` does not exist on the GPU box; `__graft_entry__.build()`
therefore mirrors the reference's `src/` tree into the git-ignored `baseline/_ref/src/` whenever
`/root/reference` is present (the directory travels to the GPU box with the snapshot like the built `.so`,
and never enters the history).  Search order: `$VQ_REF_SRC`, `/root/reference/src`, `baseline/_ref/src`.
Used by `oracle/make_golden*.py`, by tests that skip when no copy of the reference is reachable, and by
`bench.py --impl reference` / the `cpu_baseline` leg (the reference's own module on the host cores).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_MIRROR = os.path.join(_REPO, "baseline", "_ref", "src")
REF_SRC_CANDIDATES = [os.environ.get("VQ_REF_SRC", ""), "/root/reference/src", REF_MIRROR]


def mirror_reference(src: str = "/root/reference/src") -> bool:
    """Copy the reference's python sources into baseline/_ref/src (git-ignored) so that they travel to the GPU box.
    Returns True when the mirror exists afterwards."""
    import shutil
    if os.path.isfile(os.path.join(src, "networks", "vq", "vq_module.py")):
        if os.path.isdir(REF_MIRROR):
            shutil.rmtree(REF_MIRROR)
        shutil.copytree(src, REF_MIRROR, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return os.path.isfile(os.path.join(REF_MIRROR, "networks", "vq", "vq_module.py"))


def reference_src():
    for p in REF_SRC_CANDIDATES:
        if p and os.path.isfile(os.path.join(p, "networks", "vq", "vq_module.py")):
            return p
    return None


def reference_available() -> bool:
    return reference_src() is not None


def _install_stubs(src: str) -> None:
    if "utils" not in sys.modules or not hasattr(sys.modules["utils"], "get_world_size"):
        u = types.ModuleType("utils")
        u.get_world_size = lambda: int(os.environ.get("WORLD_SIZE", 1))
        u.is_distributed = lambda: int(os.environ.get("WORLD_SIZE", 1)) > 1
        sys.modules["utils"] = u
    if "networks" not in sys.modules or getattr(sys.modules["networks"], "__path__", None) != [
            os.path.join(src, "networks")]:
        n = types.ModuleType("networks")
        n.__path__ = [os.path.join(src, "networks")]
        sys.modules["networks"] = n
    if "kmeans_pytorch" not in sys.modules:
        k = types.ModuleType("kmeans_pytorch")
        k.kmeans = lambda *a, **kw: (_ for _ in ()).throw(RuntimeError("kmeans stub"))
        sys.modules["kmeans_pytorch"] = k


def load_reference_vq():
    """Returns the reference `VQModule` class (vq/vq_module.py:139)."""
    src = reference_src()
    if src is None:
        raise RuntimeError("reference sources not present (expected /root/reference/src or baseline/_ref/src)")
    _install_stubs(src)
    return importlib.import_module("networks.vq.vq_module").VQModule


def load_reference_net(name: str):
    """`load_reference_net('vqwnet').VQWNet`, `('unet_encoder').UNetEncoder`, ..."""
    src = reference_src()
    if src is None:
        raise RuntimeError("reference sources not present (expected /root/reference/src or baseline/_ref/src)")
    _install_stubs(src)
    return importlib.import_module("networks." + name)
