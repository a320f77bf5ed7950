"""The two layers that produce the quantiser's input, fused.

In every reference network the tensor handed to `VQ` is the output of `nn.InstanceNorm2d(C)` + `nn.ReLU(inplace=True)`
-- the end of `up_conv1_1.double_conv` (blocks.py:39-50; vqwnet.py:104-109, unet_encoder.py:108-113).  Stock torch runs
that pair as five N*D-sized memory passes forward and seven backward.  `InstanceNormReLU` writes z once from the
convolution output and g_x once from g_z (`vq_norm_relu_fwd` / `vq_norm_relu_bwd`: thread-block clusters per plane,
statistics exchanged through distributed shared memory, second pass out of L2).  z is still materialised: the networks
return it as `embed` (vqwnet.py:108) and the trainers feed it to `EmbeddingLoss`.

`fuse_vq_tail(net)` swaps the pair in a reference-style network in place; neither layer has parameters or buffers
(`InstanceNorm2d` defaults: affine=False, track_running_stats=False), so `state_dict` keys are unchanged.
No CPU / eager fallback: CPU tensors raise.
"""
from __future__ import annotations

import torch
import torch.nn as nn

try:
    from ..._native import lib, check
except ImportError:  # dropped into the reference tree
    from medical_image_editing_b200._native import lib, check


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _NormReLU(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x: torch.Tensor, eps: float) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("B200 InstanceNormReLU: input must be a CUDA tensor; there is no CPU fallback")
        if x.dtype != torch.float32:
            raise TypeError("B200 InstanceNormReLU: input must be float32 (the reference runs fp32)")
        if x.dim() != 4:
            raise ValueError(f"B200 InstanceNormReLU: input must be [B, C, H, W], got {tuple(x.shape)}")
        B, C, H, W = x.shape
        with torch.cuda.device(x.device):
            x = x.contiguous()
            z = torch.empty_like(x)
            stats = torch.empty(B * C, 2, dtype=torch.float32, device=x.device)
            check(lib().vq_norm_relu_fwd(x.data_ptr(), z.data_ptr(), stats.data_ptr(), B, C, H, W, float(eps), _stream()),
                  "vq_norm_relu_fwd")
        ctx.save_for_backward(x, stats)
        return z

    @staticmethod
    def backward(ctx, g_z: torch.Tensor):
        x, stats = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return None, None
        B, C, H, W = x.shape
        with torch.cuda.device(x.device):
            g = g_z.to(torch.float32).contiguous()
            g_x = torch.empty_like(x)
            check(lib().vq_norm_relu_bwd(g.data_ptr(), x.data_ptr(), stats.data_ptr(), g_x.data_ptr(), B, C, H, W, _stream()),
                  "vq_norm_relu_bwd")
        return g_x, None


def instance_norm_relu(x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """relu(instance_norm(x)) for [B, C, H, W] fp32 CUDA tensors (no affine parameters, no running statistics)."""
    return _NormReLU.apply(x, eps)


class InstanceNormReLU(nn.Module):
    """Drop-in for the pair `nn.InstanceNorm2d(num_features)`, `nn.ReLU()` (blocks.py:43-44, 46-47)."""

    def __init__(self, num_features: int, eps: float = 1e-5):
        super().__init__()
        self.num_features = num_features
        self.eps = eps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 4 or x.shape[1] != self.num_features:
            raise ValueError(f"expected [B, {self.num_features}, H, W], got {tuple(x.shape)}")
        return instance_norm_relu(x, self.eps)

    def extra_repr(self) -> str:
        return f"{self.num_features}, eps={self.eps}"


def fuse_norm_relu_pairs(seq: nn.Sequential, only_last: bool = False) -> int:
    """Replace every (`InstanceNorm2d` without affine / running statistics, `ReLU`) pair of an `nn.Sequential` by
    (`InstanceNormReLU`, `Identity`) in place -- indices, and therefore `state_dict` keys, stay what they were.
    Returns the number of pairs replaced."""
    idx = [i for i in range(len(seq) - 1)
           if isinstance(seq[i], nn.InstanceNorm2d) and isinstance(seq[i + 1], nn.ReLU)
           and not seq[i].affine and not seq[i].track_running_stats]
    if only_last:
        idx = idx[-1:]
    for i in idx:
        seq[i] = InstanceNormReLU(seq[i].num_features, seq[i].eps)
        seq[i + 1] = nn.Identity()
    return len(idx)


def fuse_vq_tail(net: nn.Module) -> int:
    """The layers in front of the quantiser of a reference-style network (`VQWNet`, `UNetEncoder`, ...): the last
    InstanceNorm2d + ReLU of `net.up_conv1_1.double_conv.double_conv` (vqwnet.py:104, blocks.py:9-18, 39-50)."""
    block = getattr(net, "up_conv1_1", None)
    dc = getattr(getattr(block, "double_conv", None), "double_conv", None)
    if not isinstance(dc, nn.Sequential):
        raise ValueError("fuse_vq_tail: expected net.up_conv1_1.double_conv.double_conv to be an nn.Sequential")
    return fuse_norm_relu_pairs(dc, only_last=True)
