"""Generate tests/golden/embed_loss_*.npz by running the UNMODIFIED reference `EmbeddingLoss`
(/root/reference/src/functions/embed_loss.py) on seeded inputs -- TEST INFRASTRUCTURE.

    python oracle/make_golden_embed_loss.py

Stores inputs (two views, integer label maps, codebook), the three losses and the gradients of
l_cross w.r.t. both views (fp32, CPU, torch as installed)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.embed_loss_oracle import load_reference_embedding_loss, onehot_strip0, seeded_embed_case  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
# name: (B, D, H, K, seed, frac_none, blocky, margin)
CASES = {
    "embed_loss_k10_d16": (2, 16, 16, 10, 31, 0.15, True, 0.5),
    "embed_loss_k6_d8_sparse": (3, 8, 12, 6, 32, 0.6, False, 1.0),       # a class absent from some images
    "embed_loss_k24_d20_ragged": (1, 20, 10, 24, 33, 0.1, False, 2.0),   # H*W not a multiple of 4 per row pattern
}


def main():
    Ref = load_reference_embedding_loss()
    assert Ref is not None, "reference sources not present"
    for name, (B, D, H, K, seed, fn, blocky, margin) in CASES.items():
        e1, l1, e2, l2, cb = seeded_embed_case(B, D, H, K, seed, fn, blocky)
        e1.requires_grad_(True)
        e2.requires_grad_(True)
        m = Ref(dict_size=K, margin=margin, use_distance_loss=True, use_regularization_loss=True)
        lc, ld, lr = m(e1, onehot_strip0(l1, K), e2, onehot_strip0(l2, K), cb)
        g1, g2 = torch.autograd.grad(lc, (e1, e2))
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), embed_1=e1.detach().numpy(), labels_1=l1.numpy(),
                            embed_2=e2.detach().numpy(), labels_2=l2.numpy(), codebook=cb.numpy(),
                            l_cross=np.float32(lc.item()), l_dist=np.float32(float(ld)), l_reg=np.float32(float(lr)),
                            g_1=g1.numpy(), g_2=g2.numpy(), margin=np.float32(margin))
        print(name, float(lc), float(ld), float(lr))


if __name__ == "__main__":
    main()
