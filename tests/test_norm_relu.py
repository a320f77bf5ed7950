"""the InstanceNorm2d + ReLU pair in front of the quantiser, fused (`vq_norm_relu_fwd/bwd`).
CPU: the oracle against the stock layers (and the reference's own DoubleConv tail), the in-place module swap.
GPU: parity of the CUDA kernels with the oracle, <= 1e-5 relative to the largest element (forward) / 2e-5 (backward)."""
import pytest
import torch
import torch.nn as nn

from oracle import ref_loader
from oracle.norm_relu_oracle import norm_relu_oracle, norm_relu_oracle_grad
from util import rel_err

DEV = "cuda:0"


def _x(shape, seed, mean=0.0, std=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * std + mean


# ----------------------------------------------------------------------------------------------- CPU
def test_oracle_is_the_stock_layer_pair():
    x = _x((2, 5, 12, 9), 1, mean=0.3)
    ref = nn.Sequential(nn.InstanceNorm2d(5), nn.ReLU(inplace=True))(x.clone())
    assert torch.equal(norm_relu_oracle(x), ref)


@pytest.mark.skipif(not ref_loader.reference_available(), reason="no copy of the reference sources reachable")
def test_oracle_matches_reference_double_conv_tail():
    blocks = ref_loader.load_reference_net("blocks")                  # networks/blocks.py:39-50
    torch.manual_seed(0)
    dc = blocks.DoubleConv(3, 6).double_conv
    x = _x((2, 3, 16, 16), 2)
    with torch.no_grad():
        pre = dc[:4](x)                                               # conv, IN, ReLU, conv
        assert torch.equal(norm_relu_oracle(pre), dc[4:](pre.clone()))


def test_fuse_keeps_state_dict_and_structure():
    from medical_image_editing_b200.src.functions import InstanceNormReLU, fuse_norm_relu_pairs, fuse_vq_tail
    seq = nn.Sequential(nn.Conv2d(3, 4, 3, padding=1), nn.InstanceNorm2d(4), nn.ReLU(inplace=True),
                        nn.Conv2d(4, 4, 3, padding=1), nn.InstanceNorm2d(4), nn.ReLU(inplace=True))
    keys = list(seq.state_dict().keys())
    assert fuse_norm_relu_pairs(seq, only_last=True) == 1
    assert isinstance(seq[4], InstanceNormReLU) and isinstance(seq[5], nn.Identity) and isinstance(seq[1], nn.InstanceNorm2d)
    assert list(seq.state_dict().keys()) == keys
    assert fuse_norm_relu_pairs(seq) == 1 and isinstance(seq[1], InstanceNormReLU)
    aff = nn.Sequential(nn.InstanceNorm2d(4, affine=True), nn.ReLU())
    assert fuse_norm_relu_pairs(aff) == 0                             # affine norms are left alone
    with pytest.raises(RuntimeError):
        InstanceNormReLU(4)(torch.zeros(1, 4, 8, 8))                  # CPU tensor: no fallback
    with pytest.raises(ValueError):
        fuse_vq_tail(nn.Linear(2, 2))


@pytest.mark.skipif(not ref_loader.reference_available(), reason="no copy of the reference sources reachable")
def test_fuse_vq_tail_on_reference_vqwnet_structure():
    from medical_image_editing_b200.src.functions import InstanceNormReLU, fuse_vq_tail
    mod = ref_loader.load_reference_net("vqwnet")
    net = mod.VQWNet(1, 1, filters=[8, 8, 16, 16, 32], dict_size=16)
    keys = list(net.state_dict().keys())
    assert fuse_vq_tail(net) == 1
    assert isinstance(net.up_conv1_1.double_conv.double_conv[4], InstanceNormReLU)
    assert list(net.state_dict().keys()) == keys


# ----------------------------------------------------------------------------------------------- GPU
SHAPES = [((2, 8, 16, 16), 0.0, 1.0), ((3, 5, 7, 9), 0.5, 2.0),            # warp-per-plane path; HW % 4 != 0: scalar path
          ((2, 16, 32, 32), 0.3, 1.5), ((1, 3, 2, 514), 0.0, 1.0),         # largest warp-per-plane plane; just above it
          ((1, 64, 256, 256), 0.2, 1.0),                                   # config-2 planes: one CTA per plane
          ((2, 3, 512, 512), -1.0, 0.5),                                   # clusters of 4 (forward) / 8 (backward)
          ((1, 2, 1024, 1024), 0.0, 1.0),                                  # segments larger than the L2 budget
          ((2, 4, 64, 64), 50.0, 0.1),                                     # mean >> std: conditioning of the shifted sums
          ((1, 300, 12, 12), 0.0, 1.0)]                                    # more planes than persistent CTAs x ...


@pytest.mark.gpu
@pytest.mark.parametrize("shape,mean,std", SHAPES, ids=[str(s[0]) for s in SHAPES])
def test_norm_relu_cuda_matches_oracle(shape, mean, std):
    from medical_image_editing_b200.src.functions import instance_norm_relu
    x = _x(shape, 7, mean, std)
    g_z = _x(shape, 8)
    ref = norm_relu_oracle(x)
    ref_g = norm_relu_oracle_grad(x.double(), g_z.double()).float()       # float64 autograd: the tighter reference
    # An element whose normalised value is within rounding of 0 may fall on either side of the relu in two correct fp32
    # implementations (its gradient is then g_z * rstd or 0): such elements are excluded from the gradient comparison;
    # their effect on the plane means (1 / HW of one element) stays inside the tolerance.
    safe = torch.nn.functional.instance_norm(x.double()).abs() > 1e-5

    def grad_err(a, b, m):
        return float(((a.double() - b.double()).abs() * m).max() / b.double().abs().max().clamp_min(1e-30))

    xd = x.to(DEV).requires_grad_(True)
    z = instance_norm_relu(xd)
    (g_x,) = torch.autograd.grad(z, xd, g_z.to(DEV))
    assert z.shape == x.shape and z.is_contiguous()
    assert rel_err(z.cpu(), ref) <= 1e-5                                   # max |a - b| / max |b|
    assert bool((((z.cpu() > 0) == (ref > 0)) | ~safe).all())              # relu mask: only zero crossings may flip
    assert grad_err(g_x.cpu(), ref_g, safe) <= 2e-5
    # against stock torch-CUDA on the same device (the path the reference takes on a GPU)
    xt = x.to(DEV).requires_grad_(True)
    zt = torch.relu(torch.nn.functional.instance_norm(xt))
    (gt,) = torch.autograd.grad(zt, xt, g_z.to(DEV))
    assert rel_err(z, zt) <= 1e-5 and grad_err(g_x, gt, safe.to(DEV)) <= 2e-5
    print(f"[norm_relu] {shape}: {int((~safe).sum())} elements within 1e-5 of the relu threshold excluded")


@pytest.mark.gpu
def test_norm_relu_edge_cases():
    from medical_image_editing_b200.src.functions import instance_norm_relu, InstanceNormReLU
    const = torch.full((1, 2, 8, 8), 3.0, device=DEV)                      # zero variance: output 0 like torch
    assert float(instance_norm_relu(const).abs().max()) == 0.0
    assert instance_norm_relu(torch.zeros(0, 4, 8, 8, device=DEV)).shape == (0, 4, 8, 8)
    nc = torch.randn(2, 6, 16, 16, device=DEV)[:, ::2]                     # non-contiguous input
    assert rel_err(instance_norm_relu(nc).cpu(), norm_relu_oracle(nc.cpu())) <= 1e-5
    with torch.no_grad():                                                   # inference: no graph, same values
        y = InstanceNormReLU(3)(nc)
    assert not y.requires_grad
    with pytest.raises(ValueError):
        InstanceNormReLU(4)(nc)


@pytest.mark.gpu
@pytest.mark.skipif(not ref_loader.reference_available(), reason="no copy of the reference sources reachable")
def test_fused_tail_inside_reference_vqwnet():
    """The reference VQWNet with this package's VQ, with and without the fused producer pair: same code map, same
    reconstruction / commitment loss / input gradient (<= 1e-5 / 1e-4 through the decoder)."""
    import copy
    import medical_image_editing_b200 as pkg
    from medical_image_editing_b200.src.functions import fuse_vq_tail
    mod = ref_loader.load_reference_net("vqwnet")
    old = mod.VQ
    mod.VQ = pkg.VQ
    try:
        torch.manual_seed(0)
        a = mod.VQWNet(1, 1, filters=[32, 32, 64, 64, 128], dict_size=64).to(DEV)
    finally:
        mod.VQ = old
    b = copy.deepcopy(a)
    assert fuse_vq_tail(b) == 1
    a.train(False)
    b.train(False)
    cudnn = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        x = _x((2, 1, 64, 64), 3).clamp(-1, 1).to(DEV)
        xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        oa, ob = a(xa), b(xb)
        assert rel_err(ob["embed"], oa["embed"]) <= 1e-5
        same = (oa["ids"] == ob["ids"]).float().mean().item()
        assert same >= 0.999, same                                         # z differs in the last bits: near-ties may flip
        if same == 1.0:
            assert rel_err(ob["recon"], oa["recon"]) <= 1e-4
            la = (oa["recon"] - x).pow(2).mean() + oa["commit_loss"]
            lb = (ob["recon"] - x).pow(2).mean() + ob["commit_loss"]
            (ga,) = torch.autograd.grad(la, xa)
            (gb,) = torch.autograd.grad(lb, xb)
            assert rel_err(gb, ga) <= 1e-3
    finally:
        torch.backends.cudnn.allow_tf32 = cudnn


@pytest.mark.gpu
def test_all_pairs_fused_in_wnet_harness_training_step():
    """tools/wnet.py with every InstanceNorm2d + ReLU pair fused (bench.py --fused-norm all) against the stock layers:
    one training step from identical weights -- loss, code-map agreement, parameter gradients."""
    import copy
    import os
    import sys
    import medical_image_editing_b200 as pkg
    from medical_image_editing_b200.src.functions import InstanceNormReLU, fuse_norm_relu_pairs
    from util import ROOT
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from wnet import WNetHarness
    torch.manual_seed(0)
    a = WNetHarness(lambda d, k: pkg.VQ(emb_dim=d, dict_size=k, momentum=0.99, eps=1e-5, knn_backend="torch"), 1,
                    widths=(16, 16, 32, 32, 64), dict_size=32).to(DEV)
    b = copy.deepcopy(a)
    n = sum(fuse_norm_relu_pairs(m) for m in [m for m in b.modules() if isinstance(m, nn.Sequential)])
    assert n == 36 and sum(isinstance(m, InstanceNormReLU) for m in b.modules()) == 36      # 18 conv pairs x 2
    cudnn = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        x = _x((2, 1, 64, 64), 5).clamp(-1, 1).to(DEV)
        oa, ob = a(x), b(x)
        same = (oa["ids"] == ob["ids"]).float().mean().item()
        assert rel_err(ob["embed"], oa["embed"]) <= 1e-4 and same >= 0.995, same
        la = (oa["recon"] - x).pow(2).mean() + oa["commit_loss"]
        lb = (ob["recon"] - x).pow(2).mean() + ob["commit_loss"]
        assert abs(la.item() - lb.item()) <= 1e-3 * abs(la.item())
        la.backward()
        lb.backward()
        if same == 1.0:
            # relative to the largest gradient of the network: a convolution that feeds an InstanceNorm has a gradient of
            # ~0 by scale invariance (1e-6 of the others), i.e. pure rounding noise in both implementations
            gmax = max(float(p.grad.abs().max()) for p in a.parameters())
            for (na, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
                assert float((pb.grad - pa.grad).abs().max()) <= 2e-3 * gmax, na
    finally:
        torch.backends.cudnn.allow_tf32 = cudnn
