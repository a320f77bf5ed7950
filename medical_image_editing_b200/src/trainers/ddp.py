"""Data-parallel training of VQ-bearing networks over plain `torch.distributed` (one process per GPU).

The reference trains through PyTorch-Lightning's DDP strategy (`src/run_vqwnet.py:94-116`,
`src/trainers/single_window_trainer.py:110-145`); Lightning is not part of the hot path, so this module wires the
same step with `torch.distributed` only:

  * once:   parameters AND buffers are broadcast from rank 0 (DDP does the same at construction);
  * step:   forward on this rank's shard of the batch -> `loss = mse(recon, x) + w * commit_loss` (the subset of
            `_train_first_step` whose dependencies are on the path) -> backward -> ONE flat all-reduce of all
            gradients (averaged over ranks, as DDP does) -> optimiser step.
  * the quantiser's EMA statistics are all-reduced inside `VQ.forward` (one packed buffer, see
    `functions/vq_function.py`); every rank applies the identical update, so the codebooks stay bit-identical and no
    per-step buffer broadcast is needed (`broadcast_buffers=False` in DDP terms).

Nothing here touches the CUDA kernels; it runs on `gloo`/CPU for the host-logic tests and on `nccl` on the GPUs.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


@torch.no_grad()
def broadcast_module_state(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank `src`'s parameters and buffers (codebook included)."""
    if _world(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        # the codebook's `embed_avg` is a transposed view-like buffer (strides (1, D)): broadcast needs dense memory
        if t.is_contiguous():
            dist.broadcast(t, src, group=group)
        else:
            tmp = t.contiguous()
            dist.broadcast(tmp, src, group=group)
            t.copy_(tmp)


@torch.no_grad()
def all_reduce_gradients(params: Iterable[torch.nn.Parameter], group=None, average: bool = True) -> int:
    """One flat all-reduce over every existing gradient.  Returns the number of elements reduced."""
    ws = _world(group)
    grads = [p.grad for p in params if p.grad is not None]
    if ws == 1 or not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    if average:
        flat.div_(ws)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return off


class DataParallelVQTrainer:
    """Minimal data-parallel trainer for a model whose forward returns the reference's dict
    (`{'recon', 'commit_loss', 'ids', ...}`, `vqwnet.py:147-152`).

    `training_step(images)` takes THIS RANK's shard; with equal shard sizes the result equals single-process training
    on the concatenated batch (up to fp32 reduction order) when the quantiser uses `reduce_mode="sum"`.
    """

    def __init__(self, model: torch.nn.Module, lr: float = 1e-4, commit_weight: float = 1.0,
                 optimizer: Optional[torch.optim.Optimizer] = None, group=None) -> None:
        self.model = model
        self.group = group
        self.commit_weight = commit_weight
        broadcast_module_state(model, 0, group)
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.optimizer = optimizer if optimizer is not None else torch.optim.Adam(self.params, lr=lr)
        self.world_size = _world(group)

    def training_step(self, images: torch.Tensor) -> Dict[str, torch.Tensor]:
        self.model.train(True)
        self.optimizer.zero_grad(set_to_none=True)
        out = self.model(images)
        recon_loss = F.mse_loss(out["recon"], images)
        loss = recon_loss + self.commit_weight * out["commit_loss"]
        loss.backward()
        all_reduce_gradients(self.params, self.group, average=True)
        self.optimizer.step()
        return {"loss": loss.detach(), "recon_loss": recon_loss.detach(), "commit_loss": out["commit_loss"].detach(),
                "ids": out.get("ids")}

    @torch.no_grad()
    def replicas_in_sync(self) -> bool:
        """True when every rank holds bit-identical parameters and buffers (cheap checksum exchange)."""
        if self.world_size == 1:
            return True
        sums = torch.stack([t.detach().double().sum() for t in list(self.model.parameters()) + list(self.model.buffers())])
        lo, hi = sums.clone(), sums.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
        return bool(torch.equal(lo, hi))
