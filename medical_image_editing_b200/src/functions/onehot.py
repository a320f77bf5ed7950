"""`OneHotEncoder` on the B200 -- drop-in for the reference class
(src/functions/onehot.py:5-20): `forward(t[B, *spatial]) -> float32 [B, n_classes, *spatial]`.

The reference indexes an identity matrix (`eye(C).index_select`), permutes and makes the result contiguous, then casts:
three passes over B*HW*C elements.  `vq_onehot` (csrc/vq_kernels.cu) writes the channel-major result once with 16-byte
streaming stores.  The stage-1 trainers feed it the code map (`transpose(ids, 1, 2) + 1`, single_window_trainer.py:91-99);
`EmbeddingLoss` in this package also accepts the integer map itself, which skips the B x (K+1) x HW tensor altogether.
A label outside [0, n_classes) raises in the reference (index_select); here it yields an all-zero column unless
`VQ_B200_CHECK_IDS=1`, which checks the range first.  No CPU fallback.
"""
from __future__ import annotations

import os

import torch
from torch import nn

try:
    from ..._native import lib, check
except ImportError:  # dropped into the reference tree
    from medical_image_editing_b200._native import lib, check


class OneHotEncoder(nn.Module):

    def __init__(self, n_classes):
        super().__init__()
        self.n_classes = n_classes

    def forward(self, t: torch.Tensor) -> torch.Tensor:
        if not t.is_cuda:
            raise RuntimeError("B200 OneHotEncoder: input must be a CUDA tensor; there is no CPU fallback")
        if t.dim() < 1:
            raise ValueError("B200 OneHotEncoder: input needs a batch dimension")
        if t.dtype not in (torch.int32, torch.int64):
            t = t.long()                                   # the reference does `t.long()` (:15); truncation toward zero
        t = t.contiguous()
        if os.environ.get("VQ_B200_CHECK_IDS", "0") == "1" and t.numel():
            lo, hi = int(t.min()), int(t.max())
            if lo < 0 or hi >= self.n_classes:
                raise IndexError(f"B200 OneHotEncoder: label range [{lo}, {hi}] outside [0, {self.n_classes})")
        B = t.shape[0]
        hw = t.numel() // B if B else 0
        out = torch.empty((B, self.n_classes) + tuple(t.shape[1:]), dtype=torch.float32, device=t.device)
        with torch.cuda.device(t.device):
            check(lib().vq_onehot(t.data_ptr(), t.element_size(), B, hw, self.n_classes, out.data_ptr(),
                                  torch.cuda.current_stream(t.device).cuda_stream), "vq_onehot")
        return out
