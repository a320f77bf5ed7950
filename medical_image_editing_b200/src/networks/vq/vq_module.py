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
    from ...functions.vq_function import (VQFunction, vq_lookup, wait_pending_update, REDUCE_MODES, StatsAccumulator,
                                          _exchange_and_update)
except ImportError:  # dropped into the reference tree: src/functions/vq_function.py
    from functions.vq_function import (VQFunction, vq_lookup, wait_pending_update, REDUCE_MODES, StatsAccumulator,
                                       _exchange_and_update)


class VQModule(nn.Module):
    """`VQ(emb_dim, dict_size, momentum, eps, knn_backend)` (reference :139-157).

    Extra, optional keyword (not in the reference): `reduce_mode` selects what WORLD_SIZE > 1 means
    for the EMA statistics: "reference" (default: the reference as written, :188-192 -- rank-local counts,
    sums averaged over the ranks; a drop-in user gets the reference's buffers), "sum" (identical to one process
    on the concatenated batch; what this package's own data-parallel trainer and bench ask for: replicas stay
    bit-identical without buffer broadcasts), "mean" (counts and sums averaged).  `knn_backend` is accepted and ignored, as the reference does when faiss
    is absent (:120).  `overlap_exchange=True` runs the all-reduce of the statistics and the EMA update on a side
    stream (they are not needed before the next forward), hidden behind the backward / decoder work the caller
    enqueues next; every access through this module (`forward`, `lookup`, `get_codebook`, `state_dict`) joins the
    streams first.  Read the buffer attributes directly only after `sync_codebook()`.  `accumulate_steps=n` sums the
    EMA statistics of n training forwards (micro-batches) and applies them in one exchange + one update after the n-th
    (== one forward on the concatenated batch); `flush_ema()` applies a partial accumulation.  `ids_layout="natural"`
    and `ids_base=1` return the code map as the callers build it next (transposed back to (b, h, w), 1-based)."""

    def __init__(self,
                 emb_dim: int,
                 dict_size: int,
                 momentum: float,
                 eps: float,
                 knn_backend: Optional[str] = "torch",
                 reduce_mode: str = "reference",
                 overlap_exchange: bool = False,
                 accumulate_steps: int = 1,
                 ids_layout: str = "reference",
                 ids_base: int = 0,
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
        if ids_layout not in ("reference", "natural") or ids_base not in (0, 1):
            raise ValueError("ids_layout must be 'reference' or 'natural', ids_base 0 or 1")
        # `ids` as the callers want it next: every reference network does `ids = transpose(ids, 1, 2); ids += 1` right
        # after the call (vqwnet.py:110-111, unet_encoder.py:115-116, vqvnet.py:62-63) -- two passes over an N-element
        # int64 map.  ids_layout="natural" / ids_base=1 write that form straight from the search epilogue.
        self.ids_layout = ids_layout
        self.ids_base = ids_base
        self.kernel_flags = 0
        # micro-batching (BASELINE config 5 at 2 / 4 GPUs): the EMA statistics of `accumulate_steps` training forwards
        # are summed and applied in ONE exchange + update after the last one == one forward on the concatenated batch
        self._accumulator = StatsAccumulator(accumulate_steps)

        embed = torch.randn(self.dict_size, self.emb_dim)
        self.register_buffer('embed', embed)
        self.register_buffer('cluster_size', torch.zeros(self.dict_size))
        self.register_buffer('embed_avg', self.embed.T.clone())
        self._register_state_dict_hook(lambda module, *a: module.sync_codebook())      # checkpoints see the update
        # a checkpoint load copies into the buffers in place: an update still in flight on the side stream must land first
        self._register_load_state_dict_pre_hook(lambda *a, **k: self.sync_codebook())

    _BUFFERS = ("embed", "cluster_size", "embed_avg")

    def sync_codebook(self) -> None:
        """Join an overlapped EMA update (no-op otherwise) before the buffers are read outside this module."""
        embed = self._buffers.get("embed") if "_buffers" in self.__dict__ else None
        if embed is not None:
            wait_pending_update(embed)

    def __setattr__(self, name, value):
        # `self.vq.embed = centres` (unet_encoder.py:85): the old tensor may still be written by the side stream; the
        # current stream waits for that update before the tensor is dropped (and its memory possibly reused)
        if name in self._BUFFERS and "_buffers" in self.__dict__ and name in self._buffers:
            self.sync_codebook()
        super().__setattr__(name, value)

    def _apply(self, fn, *args, **kwargs):
        # .to() / .cuda() / .float() replace the buffers: join the side stream first
        self.sync_codebook()
        return super()._apply(fn, *args, **kwargs)

    def forward(self, input: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        flags = self.kernel_flags | (16 if self.ids_layout == "natural" else 0) | (32 if self.ids_base == 1 else 0)
        return VQFunction.apply(input, self.embed, self.cluster_size, self.embed_avg,
                                self.momentum, self.eps, self.training, self.reduce_mode, flags,
                                self.overlap_exchange, self._accumulator)

    @property
    def accumulate_steps(self) -> int:
        return self._accumulator.steps

    @accumulate_steps.setter
    def accumulate_steps(self, n: int) -> None:
        self.flush_ema()
        self._accumulator = StatsAccumulator(n)

    @torch.no_grad()
    def flush_ema(self) -> bool:
        """Apply the statistics accumulated so far (fewer than `accumulate_steps` micro-batches) now.  Returns True
        when an update was made."""
        stats = self._accumulator.take()
        if stats is None:
            return False
        self.sync_codebook()
        with torch.cuda.device(self.embed.device):
            scratch = torch.empty(64, dtype=torch.uint8, device=self.embed.device)
            _exchange_and_update(self.embed, self.cluster_size, self.embed_avg, stats, self.momentum, self.eps,
                                 self.reduce_mode, self.overlap_exchange, scratch)
        return True

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
