"""Custom autograd Function of the B200 quantiser: Python/PyTorch host code over the C-ABI.

Replaces, for the reference's `VQModule.forward` (src/networks/vq/vq_module.py:159-199) and
`_CustomSTE` (src/networks/vq/grad_approximation.py:7-29), the op chain
    transpose+reshape copy -> mm -> 3 elementwise passes -> topk -> one_hot -> embedding ->
    one_hot.sum -> flatten.T @ one_hot -> [all_reduce] -> EMA -> mse_loss -> STE
by `vq_assign_fwd` (+ one packed all-reduce when WORLD_SIZE > 1) + `vq_ema_update`, and its
autograd graph by `vq_bwd`.  PyTorch is used only for device memory, streams and
torch.distributed; there is no CPU or eager fallback.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

try:  # package layout (medical_image_editing_b200.src.functions)
    from ..._native import lib, check, VQ_LAYOUT_ROWS, VQ_LAYOUT_NCHW_T
    from ..utils import get_world_size, is_distributed
except ImportError:  # dropped into the reference tree (src/functions/vq_function.py)
    from medical_image_editing_b200._native import lib, check, VQ_LAYOUT_ROWS, VQ_LAYOUT_NCHW_T
    from utils import get_world_size, is_distributed

REDUCE_MODES = ("sum", "mean", "reference")

_WORKSPACES = {}
_SIDE_STREAMS = {}        # device index -> stream of the overlapped statistics exchange
_PENDING = {}             # embed.data_ptr() -> (event recorded after an overlapped all-reduce + EMA update, recorded under capture?)


def _side_stream(device: torch.device) -> "torch.cuda.Stream":
    st = _SIDE_STREAMS.get(device.index)
    if st is None:
        st = torch.cuda.Stream(device=device)
        _SIDE_STREAMS[device.index] = st
    return st


def wait_pending_update(embed: torch.Tensor) -> None:
    """Make the current stream wait for an overlapped EMA update of this codebook (no-op when none is in flight).
    Called before anything reads or writes the buffers: the next forward, `lookup`, `get_codebook`, `state_dict`,
    `load_state_dict`, `.to()` / `.cuda()` and buffer reassignment (vq_module.py hooks).  The event stays registered
    until it has completed (or the next forward replaces it), so EVERY stream that touches the buffers waits, not
    only the first caller's."""
    if not embed.is_cuda:
        return
    key = embed.data_ptr()
    entry = _PENDING.get(key)
    if entry is None:
        return
    ev, captured = entry
    capturing = torch.cuda.is_current_stream_capturing()
    if captured != capturing:
        # An event recorded inside a CUDA-graph capture only exists inside that graph (waiting on it from eager code is
        # cudaErrorInvalidValue), and eager work of before the capture has long finished when the graph is replayed:
        # either way there is nothing to wait for.  (The capture itself joins the side stream before it ends.)
        _PENDING.pop(key, None)
        return
    torch.cuda.current_stream(embed.device).wait_event(ev)
    if capturing:
        _PENDING.pop(key, None)         # inside a capture the dependency is now an edge of the graph (no event queries there)
    elif ev.query():                    # finished: later readers on any stream need no dependency
        _PENDING.pop(key, None)



def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """Per-(device, stream) scratch buffer, grown on demand (caller-owned memory for the C-ABI)."""
    key = (device.index, _stream())
    buf = _WORKSPACES.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        _WORKSPACES[key] = buf
    return buf


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"B200 VQ: `{name}` must be a CUDA tensor (got {t.device}); there is no CPU fallback")
    if t.dtype != torch.float32:
        raise TypeError(f"B200 VQ: `{name}` must be float32 (got {t.dtype}); the reference runs fp32")


def _all_reduce_stats(stats: torch.Tensor, K: int, reduce_mode: str) -> Tuple[float, float]:
    """One packed all-reduce of [cnt_hi | cnt_lo | sums] instead of the reference's two
    (vq_module.py:188-192; the first of which reduces the unused N x K one-hot).  Returns
    (count_scale, sum_scale) for `vq_ema_update`."""
    import torch.distributed as dist
    ws = get_world_size()
    if reduce_mode == "sum":            # global-batch semantics: == one process on the concatenated batch
        dist.all_reduce(stats)
        return 1.0, 1.0
    if reduce_mode == "mean":           # author-intended: mean counts, mean sums
        dist.all_reduce(stats)
        return 1.0 / ws, 1.0 / ws
    if reduce_mode == "reference":      # as written: counts stay rank-local, sums averaged
        dist.all_reduce(stats[lib().vq_stats_sums_offset(K):])
        return 1.0, 1.0 / ws
    raise ValueError(f"reduce_mode must be one of {REDUCE_MODES}, got {reduce_mode!r}")


class StatsAccumulator:
    """EMA statistics of several forward calls (micro-batches) summed before ONE exchange + ONE EMA update, so that
    n micro-batches equal one forward on their concatenation ("Memory at 2 GPUs in config 5"):
    every micro-batch is assigned with the same, not yet updated codebook, the packed `[cnt_hi | cnt_lo | sums]`
    buffers add up exactly (counts) / in fp32 (sums), and `vq_ema_update` runs once, after the last one."""

    def __init__(self, steps: int) -> None:
        if steps < 1:
            raise ValueError("accumulate_steps must be >= 1")
        self.steps = int(steps)
        self.buf: Optional[torch.Tensor] = None
        self.count = 0

    def add(self, stats: torch.Tensor) -> bool:
        """Adds one call's statistics; True when the update is due."""
        if self.count == 0 or self.buf is None or self.buf.shape != stats.shape or self.buf.device != stats.device:
            self.buf = stats.clone()
            self.count = 1
        else:
            self.buf.add_(stats)
            self.count += 1
        return self.count >= self.steps

    def take(self) -> Optional[torch.Tensor]:
        buf, self.buf, self.count = (self.buf if self.count else None), None, 0
        return buf


def _exchange_and_update(embed, cluster_size, embed_avg, stats, momentum, eps, reduce_mode, overlap_exchange, scratch_main):
    """[all-reduce of the packed statistics] + `vq_ema_update` on the current stream, or -- `overlap_exchange` with
    WORLD_SIZE > 1 -- on the side stream behind whatever the caller enqueues next (vq_module.py:187-199)."""
    L = lib()
    dev = embed.device
    K, D = embed.shape
    if not (cluster_size.is_contiguous() and embed.is_contiguous()):
        raise RuntimeError("B200 VQ: the embed / cluster_size buffers must be contiguous")
    if tuple(embed_avg.shape) != (D, K) or tuple(cluster_size.shape) != (K,):
        raise RuntimeError("B200 VQ: embed_avg must be [emb_dim, dict_size] and cluster_size [dict_size]")
    # embed_avg = embed.T.clone() keeps strides (1, D) (vq_module.py:156): pass them through
    sd, sk = embed_avg.stride()

    def run(scratch, stream_handle):
        cscale, sscale = 1.0, 1.0
        if is_distributed():
            cscale, sscale = _all_reduce_stats(stats, K, reduce_mode)
        check(L.vq_ema_update(cluster_size.data_ptr(), embed_avg.data_ptr(), sd, sk, embed.data_ptr(),
                              stats.data_ptr(), K, D, float(momentum), float(eps), cscale, sscale,
                              scratch.data_ptr(), stream_handle), "vq_ema_update")

    if overlap_exchange and is_distributed():
        # The updated codebook is not needed before the next forward (q was gathered from the old one, the backward
        # uses the snapshot): run the all-reduce and the EMA update on a side stream so they overlap whatever the
        # caller enqueues next (the backward, the decoder); `wait_pending_update` joins the streams before the buffers
        # are touched again.
        cur, side = torch.cuda.current_stream(dev), _side_stream(dev)
        scratch = torch.empty(64, dtype=torch.uint8, device=dev)
        ready = torch.cuda.Event()
        ready.record(cur)
        side.wait_event(ready)
        with torch.cuda.stream(side):
            run(scratch, side.cuda_stream)
            done = torch.cuda.Event()
            done.record(side)
        for t_ in (stats, scratch):
            t_.record_stream(side)
        _PENDING[embed.data_ptr()] = (done, torch.cuda.is_current_stream_capturing())
    else:
        run(scratch_main, _stream())


class VQFunction(torch.autograd.Function):
    """`VQFunction.apply(z, embed, cluster_size, embed_avg, momentum, eps, training, reduce_mode, flags,
    overlap_exchange, accumulator) -> (quantized, commit_loss, ids)`.

    * `quantized` [B,D,H,W] (contiguous NCHW), gathered from the codebook as it was BEFORE this
      call's EMA update (vq_module.py:179 precedes :199); gradient passes straight through to `z`.
    * `commit_loss` = mean((z - quantized)^2), differentiable w.r.t. `z` only.
    * `ids` int64 [B,H,W] in the reference's layout: ids[b,i,j] = code of pixel (h=j, w=i)
      (vq_module.py:171,178); non-differentiable, freshly allocated (callers do `ids += 1`).
    * when `training`: `cluster_size`, `embed_avg`, `embed` are updated in place
      (vq_module.py:194-199); with WORLD_SIZE > 1 the statistics are all-reduced first.
    """

    @staticmethod
    def forward(ctx, z, embed, cluster_size, embed_avg, momentum, eps, training,
                reduce_mode="reference", flags=0, overlap_exchange=False, accumulator=None):
        _require_cuda(z, "input")
        _require_cuda(embed, "embed")
        if z.dim() != 4:
            raise ValueError(f"B200 VQ: input must be [B,D,H,W], got {tuple(z.shape)}")
        B, D, H, W = z.shape
        K = embed.shape[0]
        if embed.shape[1] != D:
            raise ValueError(f"B200 VQ: input has {D} channels but the codebook has emb_dim={embed.shape[1]}")
        if H != W and not (int(flags) & 16):
            # the reference views a (b,w,h)-ordered flatten as (b,h,w) (vq_module.py:171,178): square only
            # (the natural-order code map, VQ_FLAG_IDS_NATURAL, has no such restriction)
            raise ValueError(f"B200 VQ: only square feature maps are supported (got H={H}, W={W})")
        if reduce_mode not in REDUCE_MODES:
            raise ValueError(f"reduce_mode must be one of {REDUCE_MODES}, got {reduce_mode!r}")
        L = lib()
        dev = z.device
        with torch.cuda.device(dev):
            zc = z.contiguous()
            ec = embed.contiguous()
            need_bwd = bool(ctx.needs_input_grad[0])
            q = torch.empty((B, D, H, W), dtype=torch.float32, device=dev)
            ids = torch.empty((B, H, W), dtype=torch.int64, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            ids_nat = torch.empty((B, H, W), dtype=torch.int32, device=dev) if need_bwd else None
            snap = torch.empty((K, D), dtype=torch.float32, device=dev) if need_bwd else None
            stats = torch.empty(L.vq_stats_floats(K, D), dtype=torch.float32, device=dev) if training else None
            nbytes = L.vq_workspace_bytes(B * H * W, K, D)
            ws = _workspace(dev, nbytes)
            wait_pending_update(embed)          # an overlapped update of the previous step must land first
            check(L.vq_assign_fwd(zc.data_ptr(), B, D, H, W, ec.data_ptr(), K, ids.data_ptr(), _ptr(ids_nat),
                                  q.data_ptr(), loss.data_ptr(), _ptr(stats), _ptr(snap),
                                  ws.data_ptr(), ws.numel(), int(flags), _stream()), "vq_assign_fwd")
            if training:
                due = True
                if accumulator is not None and accumulator.steps > 1:
                    due = accumulator.add(stats)          # micro-batch: the update waits for the last one
                    if due:
                        stats = accumulator.take()
                if due:
                    _exchange_and_update(embed, cluster_size, embed_avg, stats, momentum, eps, reduce_mode,
                                         overlap_exchange, ws)
        if need_bwd:
            ctx.save_for_backward(zc, ids_nat, snap)
        ctx.shape = (B, D, H, W, K)
        ctx.mark_non_differentiable(ids)
        ctx.set_materialize_grads(False)
        return q, loss, ids

    @staticmethod
    def backward(ctx, g_q, g_loss, _g_ids):
        zc, ids_nat, snap = ctx.saved_tensors
        B, D, H, W, K = ctx.shape
        with torch.cuda.device(zc.device):
            if g_q is not None:
                g_q = g_q.contiguous()
            if g_loss is not None:
                g_loss = g_loss.contiguous()
            g_z = torch.empty_like(zc)
            check(lib().vq_bwd(_ptr(g_q), _ptr(g_loss), zc.data_ptr(), ids_nat.data_ptr(), snap.data_ptr(),
                               g_z.data_ptr(), B, D, H, W, K, _stream()), "vq_bwd")
        return g_z, None, None, None, None, None, None, None, None, None, None


def vq_lookup(ids: torch.Tensor, embed: torch.Tensor, nchw_friendly: bool = True) -> torch.Tensor:
    """`F.embedding(ids, embed)` (vq_module.py:203-206): returns a fresh tensor of shape ids.shape + (D,).

    For 3-D ids every reference caller immediately does `.transpose(1, -1)` to get NCHW
    (unet_encoder.py:120-123, vqwnet.py:158, styled_vqwnet.py:161, vqgan.py:442-443).  With
    `nchw_friendly` the values are written directly in that layout and returned as a permuted view,
    so the caller's transpose yields a contiguous NCHW tensor with no further copy.  Shapes and
    values are identical to F.embedding either way."""
    if not ids.is_cuda or not embed.is_cuda:
        raise RuntimeError("B200 VQ: lookup needs CUDA tensors; there is no CPU fallback")
    if ids.dtype != torch.int64:
        raise TypeError(f"B200 VQ: ids must be int64 (got {ids.dtype})")
    if embed.dtype != torch.float32:
        raise TypeError("B200 VQ: codebook must be float32")
    K, D = embed.shape
    L = lib()
    dev = embed.device
    wait_pending_update(embed)
    debug = os.environ.get("VQ_B200_CHECK_IDS", "0") == "1"
    with torch.cuda.device(dev):
        idc = ids.contiguous()
        ec = embed.contiguous()
        status = torch.zeros(1, dtype=torch.int32, device=dev) if debug else None
        n = idc.numel()
        if nchw_friendly and idc.dim() == 3:
            Bn, A, C = idc.shape
            mem = torch.empty((Bn, D, C, A), dtype=torch.float32, device=dev)
            check(L.vq_lookup(idc.data_ptr(), n, ec.data_ptr(), K, D, mem.data_ptr(), VQ_LAYOUT_NCHW_T,
                              Bn, A, C, _ptr(status), _stream()), "vq_lookup")
            out = mem.permute(0, 3, 2, 1)
        else:
            out = torch.empty(tuple(idc.shape) + (D,), dtype=torch.float32, device=dev)
            check(L.vq_lookup(idc.data_ptr(), n, ec.data_ptr(), K, D, out.data_ptr(), VQ_LAYOUT_ROWS,
                              0, 0, 0, _ptr(status), _stream()), "vq_lookup")
        if debug and int(status.item()) != 0:
            raise IndexError("B200 VQ: lookup ids out of range")
    return out
