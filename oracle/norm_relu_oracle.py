"""CPU oracle of the quantiser's producer layers -- TEST INFRASTRUCTURE ONLY.

The reference feeds `VQ` with the output of `nn.InstanceNorm2d(C)` + `nn.ReLU(inplace=True)` (blocks.py:43-47 inside
`DoubleConv`, the tail of `up_conv1_1`, vqwnet.py:104-109).  Both are stock torch layers with default arguments
(eps = 1e-5, affine = False, track_running_stats = False), so the restatement is the functional form of the very same
torch ops; `tests/test_norm_relu.py` checks it bit-for-bit against the layer objects and, when a copy of the reference is
reachable, against the tail of the reference's own `DoubleConv`.
"""
import torch
import torch.nn.functional as F


def norm_relu_oracle(x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """blocks.py:43-44 / 46-47: InstanceNorm2d (biased variance over H x W per (b, c), eps inside the root) then ReLU."""
    return F.relu(F.instance_norm(x.float(), eps=eps))


def norm_relu_oracle_grad(x: torch.Tensor, g_z: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """Gradient of sum(norm_relu(x) * g_z) with respect to x, through torch autograd on the CPU (float64 available via
    x.double() for a tighter comparison)."""
    xx = x.detach().clone().requires_grad_(True)
    y = F.relu(F.instance_norm(xx, eps=eps))
    (g,) = torch.autograd.grad(y, xx, g_z.to(xx.dtype))
    return g
