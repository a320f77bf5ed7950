"""Shared helpers for the test-suite."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

SMALL_CASES = [
    "tiny_k10_d16_train", "k512_d64_train_warm", "k512_d64_train_cold", "k64_d256_eval",
    "k512_d64_relu_train", "k512_d64_clustered_eval", "k100_d24_ragged_train", "k10_d16_multistep",
    "k64_d512_vqgan_eval",
]


def load_golden(name):
    d = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: d[k] for k in d.files}


def t(a, device="cpu"):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max(|b|) -- the 'within 1e-5 relative' criterion of BASELINE.json on whole tensors."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = b.abs().max().item()
    if denom == 0:
        return (a - b).abs().max().item()
    return (a - b).abs().max().item() / denom


def set_state(m, embed, cluster_size, embed_avg):
    with torch.no_grad():
        m.embed.copy_(t(embed).to(m.embed.device))
        m.cluster_size.copy_(t(cluster_size).to(m.embed.device))
        m.embed_avg.copy_(t(embed_avg).to(m.embed.device))
