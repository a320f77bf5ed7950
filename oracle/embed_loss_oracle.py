"""CPU restatement of the reference's `EmbeddingLoss` (src/functions/embed_loss.py) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.  Plain PyTorch fp32, the same ops
in the same order as the reference, except that `_calc_cross_loss` (embed_loss.py:46-66) is evaluated without the
(b, n_features, n_clusters, n_loc) expansion: with a one-hot `r_ids` the product `norm(embed - centroid)^2 * r_ids` is
non-zero only where the label names the cluster, so the per-(b, k) sums are index_add's of per-pixel squared distances.
Pinned against the unmodified reference class (imported straight from its file: it needs nothing but torch) by
`oracle/make_golden_embed_loss.py` -> tests/golden/embed_loss_*.npz, and live in tests/test_oracle.py.
"""
from __future__ import annotations

import importlib.util
import os

import torch

EPS = 1e-6          # EmbeddingLoss.epsilon, embed_loss.py:8


def onehot_strip0(ids: torch.Tensor, dict_size: int) -> torch.Tensor:
    """OneHotEncoder(dict_size + 1)(ids)[:, 1:] (functions/onehot.py:5-20, single_window_trainer.py:98-99)."""
    oh = torch.nn.functional.one_hot(ids.long(), dict_size + 1).permute(0, 3, 1, 2).float()
    return oh[:, 1:].contiguous()


def cross_loss(embed: torch.Tensor, labels: torch.Tensor, codebook: torch.Tensor) -> torch.Tensor:
    """embed (b, D, h, w); labels (b, h, w) int, 0 = none, k + 1 = class k; codebook (D, K).  embed_loss.py:46-66."""
    b, D, h, w = embed.shape
    K = codebook.shape[1]
    z = embed.reshape(b, D, h * w)
    lab = labels.reshape(b, h * w).long()
    c = codebook.detach().t()                                       # (K, D)
    valid = lab > 0
    idx = (lab - 1).clamp(min=0)
    d2 = ((z.permute(0, 2, 1) - c[idx]) ** 2).sum(2) * valid        # (b, n_loc): |z - c_label|^2
    sums = torch.zeros(b, K, dtype=d2.dtype).scatter_add_(1, idx, d2)
    cnt = torch.zeros(b, K, dtype=d2.dtype).scatter_add_(1, idx, valid.to(d2.dtype))
    present = cnt > 0
    return (sums / (cnt + EPS))[present].mean()


def embedding_loss(embed_1, labels_1, embed_2, labels_2, codebook, margin, use_dist=True, use_reg=True):
    l_cross = cross_loss(embed_1, labels_2, codebook) + cross_loss(embed_2, labels_1, codebook)     # :32-35
    l_dist = l_reg = 0.0
    if use_dist:                                                    # :68-83
        D, K = codebook.shape
        a = codebook.unsqueeze(2).expand(D, K, K)
        diff = a - a.permute(0, 2, 1)
        l_dist = (torch.clamp(2 * margin - torch.norm(diff, 2, 0), min=0) ** 2).sum() / (2 * K * (K - 1))
    if use_reg:                                                     # :85-87
        l_reg = torch.mean(torch.norm(codebook, 2, 0))
    return l_cross, l_dist, l_reg


def load_reference_embedding_loss():
    """The unmodified reference class, loaded from its own file (no package import: functions/__init__.py pulls lpips)."""
    for root in (os.environ.get("VQ_REF_SRC", ""), "/root/reference/src"):
        path = os.path.join(root, "functions", "embed_loss.py") if root else ""
        if path and os.path.isfile(path):
            spec = importlib.util.spec_from_file_location("_ref_embed_loss", path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod.EmbeddingLoss
    return None


def seeded_embed_case(B, D, H, K, seed, frac_none=0.15, blocky=True):
    """Two views (embed_1/2), their integer label maps (0 = none) and a codebook (D, K)."""
    g = torch.Generator().manual_seed(seed)
    codebook = torch.randn(D, K, generator=g)
    out = []
    for _ in range(2):
        if blocky:       # piecewise-constant labels, like a segmentation
            small = torch.randint(0, K + 1, (B, (H + 3) // 4, (H + 3) // 4), generator=g)
            lab = small.repeat_interleave(4, 1).repeat_interleave(4, 2)[:, :H, :H].contiguous()
        else:
            lab = torch.randint(1, K + 1, (B, H, H), generator=g)
        lab = lab * (torch.rand(B, H, H, generator=g) >= frac_none)
        idx = (lab - 1).clamp(min=0)
        emb = codebook.t()[idx].permute(0, 3, 1, 2) + 0.3 * torch.randn(B, D, H, H, generator=g)
        out.append((emb.contiguous(), lab.to(torch.int32)))
    return out[0][0], out[0][1], out[1][0], out[1][1], codebook
