"""B200 mirror of the reference's `EmbeddingLoss` (src/functions/embed_loss.py:6-107), the cross-view cluster loss
the stage-1 trainers apply to the quantiser's outputs (single_window_trainer.py:91-105).  SURVEY section 8(f), rank 1.

Same constructor and `forward(embed_1, r_ids_1, embed_2, r_ids_2, codebook) -> (l_cross, l_dist, l_reg)`.

`r_ids_*` may be
  * the reference's one-hot map `(b, n_clusters, h, w)` (class 0 already stripped, single_window_trainer.py:98-99), or
  * the integer label map `(b, h, w)` it was encoded from (0 = no class, k + 1 = class k) -- the cheap form: the
    reference's one-hot tensor alone is B*K*H*W floats (2 GB at config 2) and its `_calc_cross_loss` expands to
    B*D*K*H*W; the kernel (`vq_embed_loss_fwd`) reads z once and every pixel meets only the code its label names.
The cross term runs in hand-written CUDA through the C-ABI (no CPU / eager fallback); the two codebook-only terms
(`l_dist`, `l_reg`: K x K x D and K x D elements, no gradient reaches a parameter through them because the codebook is
a buffer) are a few torch ops on the (D, K) view.
"""
from __future__ import annotations

import torch
import torch.nn as nn

try:  # package layout (medical_image_editing_b200.src.functions)
    from ..._native import lib, check
except ImportError:  # dropped into the reference tree (src/functions/embed_loss.py)
    from medical_image_editing_b200._native import lib, check


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def labels_from_onehot(r_ids: torch.Tensor) -> torch.Tensor:
    """(b, n_clusters, h, w) one-hot (rows of zeros = no class) -> int32 (b, h, w), 0 = no class, k + 1 = class k.
    Only HARD one-hot maps are supported (what `OneHotEncoder` produces, single_window_trainer.py:91-99): the reference
    multiplies by `r_ids`, so a soft / weighted map would weight the distances -- that is rejected here, not approximated
    (set VQ_B200_CHECK_IDS=0 to skip the check and its device synchronisation)."""
    import os
    if os.environ.get("VQ_B200_CHECK_IDS", "1") != "0":
        hard = ((r_ids == 0) | (r_ids == 1)).all() & (r_ids.sum(1) <= 1).all()
        if not bool(hard):
            raise ValueError("B200 EmbeddingLoss: r_ids must be a hard one-hot map (values in {0, 1}, at most one class per "
                             "pixel) or an integer label map; soft / weighted maps are not supported")
    present = r_ids.sum(1) > 0
    return ((r_ids.argmax(1) + 1) * present).to(torch.int32)


class _CrossLoss(torch.autograd.Function):
    """loss = mean over present (b, k) of sum_{loc: label = k+1} |z - c_k|^2 / (count + 1e-6)   (embed_loss.py:46-66)"""

    @staticmethod
    def forward(ctx, z: torch.Tensor, labels: torch.Tensor, embed: torch.Tensor) -> torch.Tensor:
        if not (z.is_cuda and labels.is_cuda and embed.is_cuda):
            raise RuntimeError("B200 EmbeddingLoss: tensors must be CUDA tensors; there is no CPU fallback")
        if z.dtype != torch.float32 or embed.dtype != torch.float32:
            raise TypeError("B200 EmbeddingLoss: embed / codebook must be float32 (the reference runs fp32)")
        if z.dim() != 4:
            raise ValueError(f"B200 EmbeddingLoss: embed must be [B, D, H, W], got {tuple(z.shape)}")
        B, D, H, W = z.shape
        K = embed.shape[0]
        if tuple(labels.shape) != (B, H, W):             # a mismatch would read out of bounds in the kernel
            raise ValueError(f"B200 EmbeddingLoss: label map must be {(B, H, W)} for embed {tuple(z.shape)}, got {tuple(labels.shape)}")
        if embed.dim() != 2 or embed.shape[1] != D:
            raise ValueError(f"B200 EmbeddingLoss: codebook must be (n_features={D}, n_clusters), got {tuple(embed.t().shape)}")
        if labels.device != z.device or embed.device != z.device:
            raise RuntimeError("B200 EmbeddingLoss: embed, labels and codebook must be on the same device")
        with torch.cuda.device(z.device):                # launch on z's device and ITS current stream
            z = z.contiguous()
            labels = labels.to(torch.int32).contiguous()
            embed = embed.contiguous()
            L = lib()
            wb = int(L.vq_embed_loss_work_bytes(B, K))
            work = torch.empty(wb + 256, dtype=torch.uint8, device=z.device)
            off = (-work.data_ptr()) % 256
            weights = torch.empty(B * K, dtype=torch.float32, device=z.device)
            loss = torch.empty((), dtype=torch.float32, device=z.device)
            check(L.vq_embed_loss_fwd(z.data_ptr(), labels.data_ptr(), embed.data_ptr(), B, D, H, W, K, loss.data_ptr(),
                                      weights.data_ptr(), work.data_ptr() + off, wb, _stream()), "vq_embed_loss_fwd")
        ctx.save_for_backward(z, labels, embed, weights)
        return loss

    @staticmethod
    def backward(ctx, g_loss: torch.Tensor):
        z, labels, embed, weights = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return None, None, None
        B, D, H, W = z.shape
        with torch.cuda.device(z.device):
            g = g_loss.to(torch.float32).contiguous()
            g_z = torch.empty_like(z)
            check(lib().vq_embed_loss_bwd(g.data_ptr(), z.data_ptr(), labels.data_ptr(), embed.data_ptr(),
                                          weights.data_ptr(), g_z.data_ptr(), B, D, H, W, embed.shape[0], _stream()), "vq_embed_loss_bwd")
        return g_z, None, None


def cross_loss(embed: torch.Tensor, r_ids: torch.Tensor, codebook: torch.Tensor) -> torch.Tensor:
    """`EmbeddingLoss._calc_cross_loss` for one view; `codebook` is the (n_features, n_clusters) view of
    `VQ.get_codebook()`, detached like the reference does (embed_loss.py:51)."""
    labels = labels_from_onehot(r_ids) if r_ids.dim() == 4 else r_ids
    return _CrossLoss.apply(embed, labels, codebook.detach().t())


class EmbeddingLoss(nn.Module):

    epsilon = 1e-6

    def __init__(self, dict_size: int, margin: float, use_distance_loss: bool, use_regularization_loss: bool):
        super().__init__()
        self.margin = margin
        self.use_distance_loss = use_distance_loss
        self.use_regularization_loss = use_regularization_loss

    def forward(self, embed_1, r_ids_1, embed_2, r_ids_2, codebook):
        l_cross = cross_loss(embed_1, r_ids_2, codebook) + cross_loss(embed_2, r_ids_1, codebook)   # embed_loss.py:32-35
        l_dist = self._calc_distance_loss(codebook) if self.use_distance_loss else 0.0
        l_reg = self._calc_regularization_loss(codebook) if self.use_regularization_loss else 0.0
        return l_cross, l_dist, l_reg

    def _calc_distance_loss(self, codebook):
        """embed_loss.py:68-83, row-blocked so that K = 4096 does not materialise K x K x D at once"""
        n_features, n_clusters = codebook.size()
        c = codebook.t()                                            # (K, D)
        total = codebook.new_zeros(())
        step = max(1, min(n_clusters, (1 << 24) // max(1, n_clusters * n_features)))
        for i in range(0, n_clusters, step):
            diff = c[i:i + step, None, :] - c[None, :, :]           # (s, K, D)
            total = total + torch.clamp(2 * self.margin - torch.norm(diff, 2, 2), min=0).pow(2).sum()
        return total / (2 * n_clusters * (n_clusters - 1))

    def _calc_regularization_loss(self, codebook):
        return torch.mean(torch.norm(codebook, 2, 0))               # embed_loss.py:85-87
