"""EmbeddingLoss (the cross-view cluster loss that consumes the quantiser's outputs).
CPU: the oracle restatement is pinned against golden vectors produced by the UNMODIFIED reference class
(tests/golden/embed_loss_*.npz, oracle/make_golden_embed_loss.py) and, when /root/reference is present, against a live
run of it.  GPU: the CUDA path (module -> autograd Function -> C-ABI) against the golden vectors and the oracle."""
import pytest
import torch

import medical_image_editing_b200 as pkg
from oracle import embed_loss_oracle as elo
from util import load_golden, t, rel_err

TOL = 1e-5
CASES = ["embed_loss_k10_d16", "embed_loss_k6_d8_sparse", "embed_loss_k24_d20_ragged"]
DEV = "cuda:0"


def _inputs(g, device="cpu"):
    return (t(g["embed_1"], device), t(g["labels_1"], device), t(g["embed_2"], device), t(g["labels_2"], device),
            t(g["codebook"], device))


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    e1, l1, e2, l2, cb = _inputs(g)
    e1.requires_grad_(True)
    e2.requires_grad_(True)
    lc, ld, lr = elo.embedding_loss(e1, l1, e2, l2, cb, float(g["margin"]))
    g1, g2 = torch.autograd.grad(lc, (e1, e2))
    assert abs(lc.item() - float(g["l_cross"])) <= 1e-6 * abs(float(g["l_cross"]))
    assert abs(float(ld) - float(g["l_dist"])) <= 1e-6 * abs(float(g["l_dist"]))
    assert abs(float(lr) - float(g["l_reg"])) <= 1e-6 * abs(float(g["l_reg"]))
    assert rel_err(g1, t(g["g_1"])) <= 1e-6 and rel_err(g2, t(g["g_2"])) <= 1e-6


def test_oracle_matches_live_reference():
    Ref = elo.load_reference_embedding_loss()
    if Ref is None:
        pytest.skip("reference sources not present")
    e1, l1, e2, l2, cb = elo.seeded_embed_case(2, 12, 16, 7, seed=77)
    m = Ref(dict_size=7, margin=1.5, use_distance_loss=True, use_regularization_loss=True)
    lc_r, ld_r, lr_r = m(e1, elo.onehot_strip0(l1, 7), e2, elo.onehot_strip0(l2, 7), cb)
    lc, ld, lr = elo.embedding_loss(e1, l1, e2, l2, cb, 1.5)
    assert abs(lc.item() - lc_r.item()) <= 1e-6 * abs(lc_r.item())
    assert abs(float(ld) - float(ld_r)) <= 1e-6 * abs(float(ld_r)) and abs(float(lr) - float(lr_r)) <= 1e-6 * abs(float(lr_r))


def test_no_cpu_fallback():
    m = pkg.EmbeddingLoss(dict_size=4, margin=1.0, use_distance_loss=False, use_regularization_loss=False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 8, 4, 4), torch.zeros(1, 4, 4, dtype=torch.int32), torch.randn(1, 8, 4, 4),
          torch.zeros(1, 4, 4, dtype=torch.int32), torch.randn(8, 4))


@pytest.mark.gpu
@pytest.mark.parametrize("onehot", [False, True], ids=["labels", "onehot"])
@pytest.mark.parametrize("name", CASES)
def test_gpu_matches_reference_golden(name, onehot):
    g = load_golden(name)
    e1, l1, e2, l2, cb = _inputs(g, DEV)
    K = cb.shape[1]
    e1.requires_grad_(True)
    e2.requires_grad_(True)
    m = pkg.EmbeddingLoss(dict_size=K, margin=float(g["margin"]), use_distance_loss=True, use_regularization_loss=True)
    r1 = elo.onehot_strip0(l1.cpu(), K).to(DEV) if onehot else l1
    r2 = elo.onehot_strip0(l2.cpu(), K).to(DEV) if onehot else l2
    lc, ld, lr = m(e1, r1, e2, r2, cb)
    g1, g2 = torch.autograd.grad(lc, (e1, e2))
    assert abs(lc.item() - float(g["l_cross"])) <= TOL * abs(float(g["l_cross"]))
    assert abs(float(ld) - float(g["l_dist"])) <= TOL * abs(float(g["l_dist"]))
    assert abs(float(lr) - float(g["l_reg"])) <= TOL * abs(float(g["l_reg"]))
    assert rel_err(g1, t(g["g_1"])) <= TOL and rel_err(g2, t(g["g_2"])) <= TOL


@pytest.mark.gpu
@pytest.mark.parametrize("B,D,H,K,blocky", [(2, 64, 64, 512, True), (3, 16, 50, 10, True), (1, 30, 17, 9, False),
                                            (2, 256, 32, 64, True)])
def test_gpu_matches_oracle(B, D, H, K, blocky):
    e1, l1, e2, l2, cb = elo.seeded_embed_case(B, D, H, K, seed=500 + D + K, blocky=blocky)
    a1 = e1.clone().requires_grad_(True)
    a2 = e2.clone().requires_grad_(True)
    lc_r = elo.cross_loss(a1, l2, cb) + elo.cross_loss(a2, l1, cb)
    g1_r, g2_r = torch.autograd.grad(3.0 * lc_r, (a1, a2))
    m = pkg.EmbeddingLoss(dict_size=K, margin=1.0, use_distance_loss=False, use_regularization_loss=False)
    b1 = e1.to(DEV).requires_grad_(True)
    b2 = e2.to(DEV).requires_grad_(True)
    lc, ld, lr = m(b1, l1.to(DEV), b2, l2.to(DEV), cb.to(DEV))
    g1, g2 = torch.autograd.grad(3.0 * lc, (b1, b2))
    assert ld == 0.0 and lr == 0.0
    assert abs(lc.item() - lc_r.item()) <= TOL * abs(lc_r.item())
    assert rel_err(g1, g1_r) <= TOL and rel_err(g2, g2_r) <= TOL


@pytest.mark.gpu
def test_gpu_edge_cases():
    K, D, H = 5, 8, 8
    cb = torch.randn(D, K, device=DEV)
    z = torch.randn(2, D, H, H, device=DEV, requires_grad=True)
    m = pkg.EmbeddingLoss(dict_size=K, margin=1.0, use_distance_loss=False, use_regularization_loss=False)
    # no labelled location at all: the reference's mean over nothing is NaN
    none = torch.zeros(2, H, H, dtype=torch.int32, device=DEV)
    lc, _, _ = m(z, none, z, none, cb)
    assert torch.isnan(lc)
    # a single labelled pixel: loss = 2 * |z - c|^2 / (1 + 1e-6) (both views), gradient only there
    one = none.clone()
    one[1, 3, 4] = 2
    lc, _, _ = m(z, one, z, one, cb)
    (gz,) = torch.autograd.grad(lc, z)
    d2 = ((z[1, :, 3, 4] - cb[:, 1]) ** 2).sum().item()
    assert abs(lc.item() - 2 * d2 / (1 + 1e-6)) <= TOL * abs(2 * d2)
    assert int((gz != 0).sum()) == D and bool((gz[1, :, 3, 4] != 0).all())
    # labels beyond K are ignored like class-0
    big = one.clone()
    big[0, 0, 0] = K + 3
    lc2, _, _ = m(z, big, z, big, cb)
    assert abs(lc2.item() - lc.item()) <= TOL * abs(lc.item())
