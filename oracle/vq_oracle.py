"""CPU oracle for the VQ bottleneck hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (fp32, CPU) restatement of the reference quantiser
(`/root/reference/src/networks/vq/vq_module.py`, `grad_approximation.py`).  It is
the checker for the CUDA path; it is never the thing shipped or measured.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl
reference` legs may import it.  The product path (`medical_image_editing_b200`)
must never import anything from `oracle/`.

Parity pinning: the reference repo holds no tests, golden vectors or fixtures for
this path, so the oracle is pinned against the reference
itself: `oracle/make_golden.py` imports the unmodified reference `VQModule`
(stub loader in `oracle/ref_loader.py`, works only where `/root/reference`
exists), runs it on seeded inputs and commits the results under `tests/golden/`;
`tests/test_oracle.py` checks this restatement against those vectors bit-for-bit
(ids, quantized, counts) / to 1e-6 (loss, grads, EMA buffers) on every run.

The arithmetic lives in third-party PyTorch (reference pins torch==1.10.2+cu113,
`Dockerfile:22`; executed here on torch 2.11): `torch.mm`, `Tensor.topk`,
`F.one_hot`, `F.embedding`, `F.mse_loss`, `mul_/add_`.  We call the same ops in
the same order so that the oracle's rounding is the reference's rounding.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# utils/__init__.py:109-114
# ----------------------------------------------------------------------------
def get_world_size() -> int:
    return int(os.environ.get("WORLD_SIZE", 1))


def is_distributed() -> bool:
    return get_world_size() > 1


# ----------------------------------------------------------------------------
# vq_module.py:45-62  (_torch_knn, distance == 'l2')
# ----------------------------------------------------------------------------
@torch.no_grad()
def torch_knn_l2(keys: torch.Tensor, queries: torch.Tensor, num_neighbors: int = 1
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
    """scores = 2*E.z^T - |e|^2 - |z|^2 in exactly the reference's op order
    (mm, `*= 2`, `-= e2[:,None]`, `-= z2[None,:]`), then topk over the code axis."""
    scores = keys.mm(queries.t())                               # :54
    scores *= 2                                                 # :55
    scores -= (keys.pow(2)).sum(1, keepdim=True)                # :56
    scores -= (queries.pow(2)).sum(1).unsqueeze_(0)             # :57
    scores, indices = scores.topk(k=num_neighbors, dim=0, largest=True)   # :58
    return scores.t(), indices.t()                              # :59-60


def knn_l2_chunked(keys: torch.Tensor, queries: torch.Tensor, chunk: int = 65536
                   ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Same arithmetic as `torch_knn_l2`, evaluated over row-chunks of the queries
    so that the K x N score matrix fits in host memory (the reference
    needs 64 GB of transients at N=1M, K=4096).  Row-chunking does not change any
    per-element operation order."""
    s_parts, i_parts = [], []
    for lo in range(0, queries.shape[0], chunk):
        s, i = torch_knn_l2(keys, queries[lo:lo + chunk])
        s_parts.append(s)
        i_parts.append(i)
    return torch.cat(s_parts, 0), torch.cat(i_parts, 0)


# ----------------------------------------------------------------------------
# vq_module.py:132-136
# ----------------------------------------------------------------------------
def exponential_moving_average_(base: torch.Tensor, update: torch.Tensor, momentum: float
                                ) -> torch.Tensor:
    return base.mul_(momentum).add_(update, alpha=1 - momentum)


# ----------------------------------------------------------------------------
# grad_approximation.py:7-29
# ----------------------------------------------------------------------------
class _OracleSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input_forward, input_backward):
        ctx.shape = input_backward.shape
        return input_forward

    @staticmethod
    def backward(ctx, grad_in):
        return None, grad_in.sum_to_size(ctx.shape)


# ----------------------------------------------------------------------------
# vq_module.py:139-211
# ----------------------------------------------------------------------------
class OracleVQ(torch.nn.Module):
    """Restatement of `VQModule`.  `reduce_mode` selects the multi-rank semantics
    (SURVEY section 5a):
      * "reference": as written (vq_module.py:188-193) -- counts stay rank-local,
        sums are averaged over ranks;
      * "mean": author-intended -- counts and sums averaged;
      * "sum":  global-batch -- counts and sums summed (== single process on the
        concatenated batch).
    With WORLD_SIZE == 1 all three are identical.  `chunk` only bounds memory.
    """

    def __init__(self, emb_dim: int, dict_size: int, momentum: float, eps: float,
                 knn_backend: Optional[str] = "torch", reduce_mode: str = "reference",
                 chunk: Optional[int] = None) -> None:
        super().__init__()
        self.emb_dim = emb_dim
        self.dict_size = dict_size
        self.momentum = momentum
        self.eps = eps
        self._knn_backend = knn_backend
        self.reduce_mode = reduce_mode
        self.chunk = chunk
        embed = torch.randn(self.dict_size, self.emb_dim)                  # :153
        self.register_buffer("embed", embed)                               # :154
        self.register_buffer("cluster_size", torch.zeros(self.dict_size))  # :155
        self.register_buffer("embed_avg", self.embed.T.clone())            # :156

    def forward(self, input: torch.Tensor):
        quantized, ids = self._quantize(input)                             # :162
        commit_loss = F.mse_loss(input, quantized)                         # :163
        quantized = _OracleSTE.apply(quantized, input)                     # :164
        return quantized, commit_loss, ids

    @torch.no_grad()
    def _quantize(self, input: torch.Tensor):
        flatten = input.transpose(1, -1).reshape(-1, self.emb_dim)         # :171
        if self.chunk is None:
            _, ids = torch_knn_l2(self.embed, flatten)                     # :173
        else:
            _, ids = knn_l2_chunked(self.embed, flatten, self.chunk)
        b, c, h, w = input.size()
        ids = ids.view(b, h, w)                                            # :178
        quantized = self.lookup(ids).transpose(1, -1)                      # :179

        if self.training:
            # :175,183,185 -- one_hot -> sum / flatten.T @ one_hot.  Materialising the
            # N x K one-hot is only feasible for small N; bincount/index_add_ are the
            # same sums in a different association order (counts exact, sums <=1e-6 rel),
            # so use the literal form when it fits and the scatter form otherwise.
            n = flatten.shape[0]
            if n * self.dict_size <= (1 << 26):
                embed_onehot = F.one_hot(ids.view(-1), self.dict_size).to(flatten.dtype)
                embed_onehot_sum = embed_onehot.sum(dim=0)
                embed_sum = flatten.T @ embed_onehot
            else:
                embed_onehot_sum = torch.bincount(ids.view(-1), minlength=self.dict_size
                                                  ).to(flatten.dtype)
                embed_sum = torch.zeros(self.dict_size, self.emb_dim, dtype=flatten.dtype)
                embed_sum.index_add_(0, ids.view(-1), flatten)
                embed_sum = embed_sum.T.contiguous()

            if is_distributed():                                           # :187-192
                import torch.distributed as dist
                ws = get_world_size()
                if self.reduce_mode == "reference":
                    dist.all_reduce(embed_sum)
                    embed_sum /= ws
                elif self.reduce_mode == "mean":
                    dist.all_reduce(embed_onehot_sum)
                    dist.all_reduce(embed_sum)
                    embed_onehot_sum /= ws
                    embed_sum /= ws
                elif self.reduce_mode == "sum":
                    dist.all_reduce(embed_onehot_sum)
                    dist.all_reduce(embed_sum)
                else:
                    raise ValueError(self.reduce_mode)

            exponential_moving_average_(self.cluster_size, embed_onehot_sum, self.momentum)  # :194
            exponential_moving_average_(self.embed_avg, embed_sum, self.momentum)            # :195
            n_tot = self.cluster_size.sum()                                                  # :197
            cluster_size = n_tot * (self.cluster_size + self.eps) / (n_tot + self.dict_size * self.eps)  # :198
            self.embed.copy_(self.embed_avg.T / cluster_size.unsqueeze(1))                   # :199
        return quantized, ids

    def lookup(self, ids: torch.Tensor) -> torch.Tensor:                   # :203-206
        return F.embedding(ids, self.embed)

    def get_codebook(self) -> torch.Tensor:                                # :208-210
        return self.embed.transpose(0, 1)


# ----------------------------------------------------------------------------
# helpers used by tests / bench (not part of the reference)
# ----------------------------------------------------------------------------
def seeded_case(B: int, D: int, H: int, W: int, K: int, seed: int = 1234, kind: str = "gauss"):
    """Deterministic synthetic inputs.  kind:
       gauss     z ~ N(0,1), E ~ N(0,1)
       clustered z = E[randint] + 0.1*N(0,1)
       relu      z = relu(N(0,1)) (the real VQ input follows a ReLU, blocks.py:47-50)"""
    g = torch.Generator().manual_seed(seed)
    embed = torch.randn(K, D, generator=g)
    if kind == "gauss":
        z = torch.randn(B, D, H, W, generator=g)
    elif kind == "clustered":
        pick = torch.randint(0, K, (B, H, W), generator=g)
        z = embed[pick].permute(0, 3, 1, 2).contiguous() + 0.1 * torch.randn(B, D, H, W, generator=g)
    elif kind == "relu":
        z = torch.relu(torch.randn(B, D, H, W, generator=g))
    else:
        raise ValueError(kind)
    return z, embed


def make_oracle(K: int, D: int, embed: torch.Tensor, momentum: float = 0.99, eps: float = 1e-5,
                warmed: bool = False, n_for_warm: int = 0, **kw) -> OracleVQ:
    m = OracleVQ(D, K, momentum, eps, "torch", **kw)
    m.embed.copy_(embed)
    m.embed_avg.copy_(embed.T)
    if warmed:
        g = torch.Generator().manual_seed(99)
        cs = torch.rand(K, generator=g) * (n_for_warm / K) + 1.0
        m.cluster_size.copy_(cs)
        m.embed_avg.copy_(embed.T * cs.unsqueeze(0))
    return m


def score_gap_is_tie(embed: torch.Tensor, z_flat: torch.Tensor, a: torch.Tensor, b: torch.Tensor,
                     ulps: float = 4.0) -> torch.Tensor:
    """For rows where two implementations picked different codes a/b, decide whether the
    oracle's own fp32 scores for a and b are within `ulps` units in the last place of each
    other (=> the row is a numerical tie in the reference's arithmetic, any maximiser is
    valid; SURVEY section 7 'Tie-break is unspecified')."""
    ea, eb = embed[a], embed[b]
    sa = 2 * (ea * z_flat).sum(1) - ea.pow(2).sum(1) - z_flat.pow(2).sum(1)
    sb = 2 * (eb * z_flat).sum(1) - eb.pow(2).sum(1) - z_flat.pow(2).sum(1)
    mag = torch.maximum(torch.maximum(ea.pow(2).sum(1), z_flat.pow(2).sum(1)), sa.abs())
    ulp = torch.finfo(torch.float32).eps * mag
    return (sa - sb).abs() <= ulps * ulp
