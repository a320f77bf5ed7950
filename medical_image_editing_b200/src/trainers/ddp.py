"""Data-parallel training of VQ-bearing networks over plain `torch.distributed` (one process per GPU).

The reference trains through PyTorch-Lightning's DDP strategy (`src/run_vqwnet.py:94-116`,
`src/trainers/single_window_trainer.py:110-145`); Lightning is not part of the hot path, so this module wires the
same step with `torch.distributed` only:

  * once:   parameters AND buffers are broadcast from rank 0 (DDP does the same at construction);
  * step:   forward on this rank's shard of the batch -> `loss = mse(recon, x) + w * commit_loss` (the subset of
            `_train_first_step` whose dependencies are on the path) -> backward, during which the gradients are
            all-reduced bucket by bucket (`GradientBuckets`: the `.grad` tensors are views of a few flat buffers,
            a bucket's all-reduce is launched as soon as its last gradient has been accumulated, so the exchange
            overlaps the rest of the backward and nothing is concatenated or copied back) -> optimiser step.
  * the quantiser's EMA statistics are all-reduced inside `VQ.forward` (one packed buffer, see
    `functions/vq_function.py`); every rank applies the identical update, so the codebooks stay bit-identical and no
    per-step buffer broadcast is needed (`broadcast_buffers=False` in DDP terms).

Nothing here touches the CUDA kernels; it runs on `gloo`/CPU for the host-logic tests and on `nccl` on the GPUs.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F


_FUSED_ADAM = os.environ.get("VQ_TRAINER_FUSED_ADAM", "1") == "1"      # default optimiser only; see DESIGN section 5


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


class GradientBuckets:
    """Gradients as views of flat buckets, all-reduced (averaged) while the backward is still running.

    Parameters are taken in reverse registration order (roughly the order in which autograd finishes them) and packed
    into buckets of about `bucket_bytes`; every `p.grad` is a view into its bucket, so autograd accumulates in place
    and a bucket is ready for `all_reduce` the moment its last gradient has landed (post-accumulate hooks).
    `finish()` waits for the outstanding collectives.  With world size 1 nothing is registered."""

    def __init__(self, params, group=None, bucket_bytes: int = 32 << 20) -> None:
        self.group = group
        self.ws = _world(group)
        self.params = [p for p in params if p.requires_grad]
        self.buckets, self._pending, self._works, self._handles = [], [], [], []
        self._of = {}
        self.defer = False            # micro-batching: gradients of all but the last micro-batch only accumulate
        if self.ws == 1 or not self.params:
            return
        cur, cur_bytes = [], 0
        groups = []
        for p in reversed(self.params):
            if cur and (cur_bytes + p.numel() * p.element_size() > bucket_bytes or p.dtype != cur[0].dtype
                        or p.device != cur[0].device):
                groups.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += p.numel() * p.element_size()
        if cur:
            groups.append(cur)
        for bi, ps in enumerate(groups):
            flat = torch.zeros(sum(p.numel() for p in ps), dtype=ps[0].dtype, device=ps[0].device)
            off = 0
            for p in ps:
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
                self._of[p] = bi
                self._handles.append(p.register_post_accumulate_grad_hook(self._ready))
            self.buckets.append((flat, ps))
            self._pending.append(len(ps))
        backend = dist.get_backend(group)
        self._avg = dist.ReduceOp.AVG if backend == "nccl" else None      # gloo has no AVG: divide afterwards

    def zero(self) -> None:
        """Replaces `optimizer.zero_grad(set_to_none=True)`: the views must survive."""
        for bi, (flat, ps) in enumerate(self.buckets):
            flat.zero_()
            self._pending[bi] = len(ps)
            off = 0
            for p in ps:                                   # an optimiser / user may have dropped or replaced a view
                if p.grad is None or p.grad.data_ptr() != flat.data_ptr() + off * flat.element_size():
                    p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
        self._works = []

    def rearm(self) -> None:
        """Next micro-batch of the same step: keep the accumulated gradients, count the hooks again."""
        for bi, (_, ps) in enumerate(self.buckets):
            self._pending[bi] = len(ps)

    def _ready(self, p) -> None:
        if self.defer:
            return
        bi = self._of[p]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            flat = self.buckets[bi][0]
            if self._avg is not None:
                self._works.append((dist.all_reduce(flat, op=self._avg, group=self.group, async_op=True), None))
            else:
                self._works.append((dist.all_reduce(flat, group=self.group, async_op=True), flat))

    def finish(self) -> int:
        """Wait for the bucket all-reduces launched during backward; reduce the buckets whose hooks never all fired
        (parameters without a gradient this step).  Returns the number of elements exchanged."""
        if self.ws == 1:
            return 0
        for bi, (flat, ps) in enumerate(self.buckets):
            if self._pending[bi] > 0:                      # some gradients of the bucket were not produced: still exchange
                self._pending[bi] = 0
                self._works.append((dist.all_reduce(flat, group=self.group, async_op=True), flat))
        n = 0
        for work, flat in self._works:
            work.wait()
            if flat is not None:
                flat.div_(self.ws)
        for flat, _ in self.buckets:
            n += flat.numel()
        self._works = []
        return n


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
        for mod in model.modules():                       # global-batch EMA statistics: replicas stay bit-identical
            if hasattr(mod, "reduce_mode") and hasattr(mod, "embed_avg"):
                mod.reduce_mode = "sum"
        self.params = [p for p in model.parameters() if p.requires_grad]
        if optimizer is None:
            # one multi-tensor kernel per step instead of torch's per-operation `foreach` passes (same update rule)
            fused = _FUSED_ADAM and all(p.is_cuda for p in self.params)
            optimizer = torch.optim.Adam(self.params, lr=lr, fused=True) if fused else torch.optim.Adam(self.params, lr=lr)
        self.optimizer = optimizer
        self.world_size = _world(group)
        self.buckets = GradientBuckets(self.params, group)

    def training_step(self, images: torch.Tensor, micro_batches: int = 1) -> Dict[str, torch.Tensor]:
        """One optimiser step on THIS RANK's shard.  `micro_batches > 1` (BASELINE config 5 at 2 / 4 GPUs: 64 / 32 slices
        of 512 x 512 per GPU do not fit one forward) splits the shard, accumulates the gradients and -- through the
        quantisers' `accumulate_steps` -- the EMA statistics, and exchanges both once: the step equals the one-shot step
        on the whole shard (up to fp32 summation order)."""
        self.model.train(True)
        if self.world_size > 1:
            self.buckets.zero()
        else:
            self.optimizer.zero_grad(set_to_none=True)
        n = max(1, int(micro_batches))
        if images.shape[0] % n:
            raise ValueError(f"batch of {images.shape[0]} slices does not split into {n} equal micro-batches")
        for mod in self.model.modules():
            if hasattr(mod, "accumulate_steps") and hasattr(mod, "embed_avg") and mod.accumulate_steps != n:
                mod.accumulate_steps = n
        mb = images.shape[0] // n
        out, loss, recon_loss = None, None, None
        for i in range(n):
            x = images[i * mb:(i + 1) * mb]
            out = self.model(x)
            recon_loss = F.mse_loss(out["recon"], x)
            loss = recon_loss + self.commit_weight * out["commit_loss"]
            self.buckets.defer = i < n - 1                # only the last micro-batch's backward launches the exchange
            if i == n - 1:
                self.buckets.rearm()
            (loss / n if n > 1 else loss).backward()      # bucket all-reduces are launched from inside the backward
        self.buckets.finish()
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
