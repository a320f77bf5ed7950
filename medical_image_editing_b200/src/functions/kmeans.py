"""k-means codebook initialisation on the quantiser's own CUDA kernels.

The reference seeds `vq.embed` with `kmeans_pytorch.kmeans` on the gathered encoder output
(src/networks/unet_encoder.py:66-91; kmeans-pytorch==0.3.0, requirements.txt:52): all_gather -> Lloyd iterations on
rank 0 with an N x K distance matrix per iteration -> broadcast.  A Lloyd iteration is exactly the quantiser's training
forward with the EMA switched off -- nearest-centre search (`vq_assign_fwd`: tcgen05 search + exact fp32 re-rank) and the
per-cluster counts / sums it already accumulates -- so nothing is gathered and no distance matrix exists:

    every rank:  assign its own shard  ->  packed all-reduce of [counts | sums]  ->  c[k] = sums[k] / counts[k]

and every rank ends with bit-identical centres (no broadcast).  `kmeans(...)` keeps the package's call signature for the
arguments the reference passes; `initialize_embed(vq, embed)` is the drop-in for `UNetEncoder.initialize_embed`.
No CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

try:
    from ..._native import lib, check
    from ..utils import get_world_size, is_distributed
except ImportError:  # dropped into the reference tree
    from medical_image_editing_b200._native import lib, check
    from utils import get_world_size, is_distributed


def _lloyd(z_dn: torch.Tensor, centers: torch.Tensor, tol: float, iter_limit: int, empty: str):
    """z_dn: [D, N] fp32 CUDA (channel-major, the layout the search kernels read); centres [K, D].  In-place on `centers`."""
    L = lib()
    D, N = z_dn.shape
    K = centers.shape[0]
    dev = z_dn.device
    stream = torch.cuda.current_stream(dev).cuda_stream
    ids = torch.empty(N, dtype=torch.int64, device=dev)
    q = torch.empty((D, N), dtype=torch.float32, device=dev)            # the search kernels always gather; not used here
    loss = torch.empty((), dtype=torch.float32, device=dev)
    stats = torch.empty(L.vq_stats_floats(K, D), dtype=torch.float32, device=dev)
    ws = torch.empty(max(L.vq_workspace_bytes(N, K, D), 1 << 16), dtype=torch.uint8, device=dev)
    soff = L.vq_stats_sums_offset(K)
    it = 0
    with torch.cuda.device(dev):
        while True:
            check(L.vq_assign_fwd(z_dn.data_ptr(), 1, D, 1, N, centers.data_ptr(), K, ids.data_ptr(), None, q.data_ptr(),
                                  loss.data_ptr(), stats.data_ptr(), None, ws.data_ptr(), ws.numel(), 0, stream),
                  "vq_assign_fwd")
            if is_distributed():
                import torch.distributed as dist
                dist.all_reduce(stats)
            counts = stats[:K] * 4096.0 + stats[K:2 * K]                # the histogram travels as two exact fp32 halves (c >> 12, c & 4095)
            sums = stats[soff:soff + K * D].view(K, D)
            new = sums / counts[:, None]                                # 0 / 0 = NaN for an empty cluster, as the package
            if empty == "keep":
                new = torch.where(counts[:, None] > 0, new, centers)
            shift = torch.sum(torch.sqrt(torch.sum((new - centers) ** 2, dim=1)))
            centers.copy_(new)
            it += 1
            s = float(shift)                                            # the package's stopping test reads it on the host too
            if s != s:
                raise RuntimeError("B200 kmeans: a cluster went empty (NaN centre); use empty='keep' or other initial centres")
            if s * s < tol or (iter_limit and it >= iter_limit):
                break
    return ids, centers, it        # `ids`: the assignment that produced the last update, as the package returns it


@torch.no_grad()
def kmeans(X: torch.Tensor, num_clusters: int, distance: str = "euclidean", tol: float = 1e-4,
           device: Optional[torch.device] = None, cluster_centers: Optional[torch.Tensor] = None, iter_limit: int = 0,
           seed: Optional[int] = None, empty: str = "keep") -> Tuple[torch.Tensor, torch.Tensor]:
    """`kmeans_pytorch.kmeans(X=[N, D], num_clusters, distance='euclidean', device=...)` -> (choice_cluster, centres),
    both returned on the CPU like the package does.  `cluster_centers`: explicit initial centres [K, D] (else K distinct
    rows of X, `np.random.choice` seeded with `seed`)."""
    if distance != "euclidean":
        raise NotImplementedError("B200 kmeans: only distance='euclidean' (the one the reference uses)")
    if device is not None:
        X = X.to(device)
    if not X.is_cuda:
        raise RuntimeError("B200 kmeans: X must be a CUDA tensor; there is no CPU fallback")
    X = X.float()
    if cluster_centers is None:
        rng = np.random.RandomState(seed) if seed is not None else np.random
        idx = torch.as_tensor(rng.choice(len(X), num_clusters, replace=False), dtype=torch.long, device=X.device)
        centers = X[idx].clone()
    else:
        centers = cluster_centers.to(X.device).float().clone()
    ids, centers, _ = _lloyd(X.t().contiguous(), centers.contiguous(), tol, iter_limit, empty)
    return ids.cpu(), centers.cpu()


@torch.no_grad()
def kmeans_nchw(embed: torch.Tensor, num_clusters: int, tol: float = 1e-4, iter_limit: int = 0, seed: Optional[int] = 0,
                empty: str = "keep") -> Tuple[torch.Tensor, int]:
    """Data-parallel k-means over THIS rank's [B, D, H, W] encoder output (all ranks call it with equal shapes).  The initial
    centres are K distinct pixels of the global batch drawn with the same seeded generator on every rank; each rank
    contributes the ones it owns and one all-reduce assembles them.  Returns (centres [K, D] on the device, iterations)."""
    if not embed.is_cuda:
        raise RuntimeError("B200 kmeans: embed must be a CUDA tensor; there is no CPU fallback")
    B, D, H, W = embed.shape
    z = embed.detach().float().permute(1, 0, 2, 3).reshape(D, B * H * W).contiguous()      # [D, N]: unet_encoder.py:74-75
    n = z.shape[1]
    ws = get_world_size() if is_distributed() else 1
    rank = 0
    if ws > 1:
        import torch.distributed as dist
        rank = dist.get_rank()
    pick = np.random.RandomState(seed).choice(n * ws, num_clusters, replace=False)
    centers = torch.zeros(num_clusters, D, device=embed.device)
    mine = [(k, int(g) - rank * n) for k, g in enumerate(pick) if rank * n <= g < (rank + 1) * n]
    if mine:
        ks = torch.as_tensor([m[0] for m in mine], device=embed.device)
        cols = torch.as_tensor([m[1] for m in mine], device=embed.device)
        centers[ks] = z[:, cols].t()
    if ws > 1:
        dist.all_reduce(centers)
    _, centers, it = _lloyd(z, centers, tol, iter_limit, empty)
    return centers, it


@torch.no_grad()
def initialize_embed(vq: torch.nn.Module, embed: torch.Tensor, rank: Optional[int] = None, seed: Optional[int] = 0) -> None:
    """Drop-in for `UNetEncoder.initialize_embed(embed, rank)` (unet_encoder.py:66-91): k-means centres of the encoder
    output become the codebook.  `rank` is accepted for signature parity; every rank computes the same centres."""
    centers, _ = kmeans_nchw(embed, vq.dict_size, seed=seed)
    vq.embed = centers.type_as(embed).detach()          # the reference re-assigns the buffer attribute (:85)
