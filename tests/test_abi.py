"""CPU: the C-ABI library loads, exports every symbol include/vq_b200.h declares, and the host-side
mirror of the reference interface behaves (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

import medical_image_editing_b200 as pkg
from medical_image_editing_b200 import _native
from util import ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "vq_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(vq_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert {"vq_assign_fwd", "vq_ema_update", "vq_bwd", "vq_lookup", "vq_workspace_bytes",
            "vq_last_error"} <= set(names)
    handle = ctypes.CDLL(pkg.lib_path())
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/vq_b200.h but not exported"
    assert set(names) == set(_native.SIGNATURES), "ctypes table and header disagree"


def test_version_and_sizes():
    L = pkg.lib()
    assert L.vq_version() >= 1000
    assert L.vq_stats_floats(512, 64) == 2 * 512 + 512 * 64
    assert L.vq_stats_sums_offset(5) == 12 and L.vq_stats_floats(5, 4) == 12 + 20
    small = L.vq_workspace_bytes(1024, 512, 64)
    big = L.vq_workspace_bytes(1 << 20, 512, 64)
    assert 0 < small < big
    assert L.vq_workspace_bytes(10, 0, 64) == 0


def test_argument_errors_are_reported_without_a_gpu():
    L = pkg.lib()
    rc = L.vq_assign_fwd(None, 1, 0, 4, 4, None, 8, None, None, None, None, None, None, None, 0, 0, None)
    assert rc == -1 and b"bad shape" in L.vq_last_error()
    rc = L.vq_lookup(None, 4, None, 0, 16, None, 0, 0, 0, 0, None, None)
    assert rc == -1
    rc = L.vq_ema_update(None, None, 1, 8, None, None, 8, 8, 0.99, 1e-5, 1.0, 1.0, None, None)
    assert rc == -1 and b"null" in L.vq_last_error()


def test_module_surface_matches_reference():
    m = pkg.VQ(emb_dim=16, dict_size=10, momentum=0.999, eps=1e-5, knn_backend="torch")
    sd = m.state_dict()
    assert list(sd) == ["embed", "cluster_size", "embed_avg"]
    assert sd["embed"].shape == (10, 16) and sd["cluster_size"].shape == (10,) and sd["embed_avg"].shape == (16, 10)
    assert torch.equal(sd["embed_avg"], sd["embed"].T)
    assert list(m.parameters()) == []
    assert m.get_codebook().shape == (16, 10)
    assert m.get_codebook().data_ptr() == m.embed.data_ptr()      # a view, as in the reference
    m.embed = torch.zeros(10, 16)                                   # assignable (unet_encoder.py:85)
    assert "embed" in m.state_dict()
    for attr in ("emb_dim", "dict_size", "momentum", "eps", "_knn_backend", "forward", "lookup", "_quantize"):
        assert hasattr(m, attr)


def test_no_cpu_fallback():
    m = pkg.VQ(emb_dim=8, dict_size=4, momentum=0.99, eps=1e-5, knn_backend=None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 8, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.lookup(torch.zeros(1, 4, 4, dtype=torch.long))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setenv("VQ_B200_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_native, "_LIB", None)
    with pytest.raises(RuntimeError, match="not found"):
        _native.lib()


def test_product_never_imports_the_oracle():
    pkg_dir = os.path.dirname(pkg.__file__)
    for dp, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle/"
