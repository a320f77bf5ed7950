"""CPU: the VQ-W-Net measurement harness (tools/wnet.py) against the unmodified reference network, and the
`--workload vqwnet` reference arm of bench.py."""
import json
import os
import subprocess
import sys

import pytest
import torch

from util import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle.ref_loader import reference_available, load_reference_net      # noqa: E402
from oracle.vq_oracle import OracleVQ                                       # noqa: E402
from wnet import WNetHarness, reference_key_order                           # noqa: E402

W = [4, 8, 16, 32, 64]


def _harness(k=16):
    return WNetHarness(lambda d, n: OracleVQ(d, n, 0.99, 1e-5, "torch"), 1, W, k)


def test_harness_step_shapes():
    torch.manual_seed(1)
    h = _harness()
    x = torch.randn(2, 1, 32, 32).clamp(-1, 1)
    out = h(x)
    assert out["recon"].shape == x.shape and out["embed"].shape == (2, W[0], 32, 32)
    assert out["ids"].shape == (2, 32, 32) and out["ids"].dtype == torch.int64
    assert int(out["ids"].min()) >= 1 and int(out["ids"].max()) <= 16          # 1-based (vqwnet.py:111)
    (torch.nn.functional.mse_loss(out["recon"], x) + out["commit_loss"]).backward()
    assert all(p.grad is not None for p in h.parameters())


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present")
def test_harness_equals_reference_network():
    """Same parameter shapes in the reference's registration order; same weights -> bit-identical training forward."""
    torch.manual_seed(0)
    ref = load_reference_net("vqwnet").VQWNet(1, 1, filters=W, dict_size=16)
    h = _harness()
    rp, hp = list(ref.parameters()), reference_key_order(h)
    assert len(hp) == len(list(h.parameters()))
    assert [tuple(p.shape) for p in rp] == [tuple(p.shape) for p in hp]
    with torch.no_grad():
        for a, b in zip(hp, rp):
            a.copy_(b)
        for k in ("embed", "cluster_size", "embed_avg"):
            getattr(h.vq, k).copy_(getattr(ref.vq, k))
    x = torch.randn(2, 1, 32, 32).clamp(-1, 1)
    ref.train()
    h.train()
    o_ref, o = ref(x), h(x)
    for k in ("recon", "embed", "commit_loss", "ids"):
        assert torch.equal(o_ref[k], o[k]), k
    for k in ("embed", "cluster_size", "embed_avg"):                           # the EMA update happened identically
        assert torch.equal(getattr(h.vq, k), getattr(ref.vq, k)), k


def test_full_size_parameter_count():
    """Default widths give the reference's 2 x U-Net(64..1024) + head: 62.4 M parameters, no quantiser parameters."""
    h = WNetHarness(lambda d, n: OracleVQ(d, n, 0.99, 1e-5, "torch"))
    n = sum(p.numel() for p in h.parameters())
    assert list(h.vq.parameters()) == []
    assert 60_000_000 < n < 65_000_000, n
