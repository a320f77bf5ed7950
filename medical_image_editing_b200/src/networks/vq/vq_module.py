"""B200-native `VQModule` -- drop-in for the reference quantiser
(reference: src/networks/vq/vq_module.py:139-211).

Same constructor, same three buffers (`embed [K,D]`, `cluster_size [K]`, `embed_avg [D,K]`; no
parameters, so strict state_dict loads keep working: trainers/base.py:85-102, run_recon.py:98-112),
same `forward(input) -> (quantized, commit_loss, ids)`, `lookup(ids)`, `get_codebook()`.
All arithmetic runs in hand-written CUDA (csrc/) behind the C-ABI in include/vq_b200.h, called from
`functions.vq_function.VQFunction`.  CPU tensors raise: there is no fallback path.
"""
from typing import Optional, Tuple

import torch
from torch import nn

try:  # package layout
    from ...functions.vq_function import VQFunction, vq_lookup, wait_pending_update, REDUCE_MODES
except ImportError:  # dropped into the reference tree: src/functions/vq_function.py
    from functions.vq_function import VQFunction, vq_lookup, wait_pending_update, REDUCE_MODES


class VQModule(nn.Module):
    """`VQ(emb_dim, dict_size, momentum, eps, knn_backend)` (reference :139-157).

    Extra, optional keyword (not in the reference): `reduce_mode` selects what WORLD_SIZE > 1 means
    for the EMA statistics: "sum" (default; identical to one process on the concatenated batch),
    "mean" (counts and sums averaged), "reference" (the reference as written, :188-192: rank-local
    counts, averaged sums).  `knn_backend` is accepted and ignored, as the reference does when faiss
    is absent (:120).  `overlap_exchange=True` runs the all-reduce of the statistics and the EMA update on a side
    stream (they are not needed before the next forward), hidden behind the backward / decoder work the caller
    enqueues next; every access through this module (`forward`, `lookup`, `get_codebook`, `state_dict`) joins the
    streams first.  Read the buffer attributes directly only after `sync_codebook()`."""

    def __init__(self,
                 emb_dim: int,
                 dict_size: int,
                 momentum: float,
                 eps: float,
                 knn_backend: Optional[str] = "torch",
                 reduce_mode: str = "sum",
                 overlap_exchange: bool = False,
                 ) -> None:
        super().__init__()
        if reduce_mode not in REDUCE_MODES:
            raise ValueError(f"reduce_mode must be one of {REDUCE_MODES}")
        self.emb_dim = emb_dim
        self.dict_size = dict_size
        self.momentum = momentum
        self.eps = eps
        self._knn_backend = knn_backend
        self.reduce_mode = reduce_mode
        self.overlap_exchange = overlap_exchange
        self.kernel_flags = 0

        embed = torch.randn(self.dict_size, self.emb_dim)
        self.register_buffer('embed', embed)
        self.register_buffer('cluster_size', torch.zeros(self.dict_size))
        self.register_buffer('embed_avg', self.embed.T.clone())
        self._register_state_dict_hook(lambda module, *a: module.sync_codebook())      # checkpoints see the update

    def sync_codebook(self) -> None:
        """Join an overlapped EMA update (no-op otherwise) before the buffers are read outside this module."""
        wait_pending_update(self.embed)

    def forward(self, input: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        return VQFunction.apply(input, self.embed, self.cluster_size, self.embed_avg,
                                self.momentum, self.eps, self.training, self.reduce_mode, self.kernel_flags,
                                self.overlap_exchange)

    @torch.no_grad()
    def _quantize(self, input: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(quantized, ids) as reference :168-201 (including the EMA update when training)."""
        quantized, _, ids = self.forward(input)
        return quantized, ids

    def lookup(self, ids: torch.Tensor) -> torch.Tensor:
        """F.embedding(ids, embed) (reference :203-206)."""
        return vq_lookup(ids, self.embed)

    def get_codebook(self) -> torch.Tensor:
        """[D,K] view of the (post-update) codebook (reference :208-210)."""
        self.sync_codebook()
        return self.embed.transpose(0, 1)
