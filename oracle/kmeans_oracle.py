"""CPU oracle of the k-means codebook initialisation -- TEST INFRASTRUCTURE ONLY.

The reference initialises the quantiser's codebook with `kmeans_pytorch.kmeans(X, num_clusters, distance='euclidean')`
(src/networks/unet_encoder.py:66-91: all_gather of the encoder output, k-means on rank 0, broadcast of the centres into
`vq.embed`).  The arithmetic lives in the third-party package kmeans-pytorch, pinned to 0.3.0 (requirements.txt:52) and
absent from /root/reference and from this image, so this file restates that release's published algorithm:

    initial centres = num_clusters distinct rows of X (np.random.choice without replacement)
    repeat:  dis[n, k] = sum_d (X[n, d] - c[k, d])^2 ; choice = argmin_k dis ; c[k] = mean of the rows with choice == k
             shift = sum_k ||c_new[k] - c_old[k]||_2 ; stop when shift^2 < tol (1e-4)

**Parity unpinned**: with the package absent there is no golden output to pin against; parity is anchored on the
reference's call site (arguments and what it does with the result).  An empty cluster gives a NaN centre in 0.3.0 (mean of
no rows), after which its stopping test can never succeed; `empty="keep"` keeps the previous centre instead (what later
releases approximate by re-seeding) and is the mode the CUDA implementation is compared with by default.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch


def initial_centers(X: torch.Tensor, num_clusters: int, seed: Optional[int] = None) -> torch.Tensor:
    rng = np.random.RandomState(seed) if seed is not None else np.random
    idx = rng.choice(len(X), num_clusters, replace=False)
    return X[torch.as_tensor(idx, dtype=torch.long)].clone()


def kmeans_oracle(X: torch.Tensor, num_clusters: int, tol: float = 1e-4, centers: Optional[torch.Tensor] = None,
                  seed: Optional[int] = None, iter_limit: int = 0, empty: str = "keep", chunk: int = 16384
                  ) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """Returns (choice_cluster [N] int64, centres [K, D] fp32, iterations)."""
    X = X.float()
    c = (initial_centers(X, num_clusters, seed) if centers is None else centers.clone()).float()
    it = 0
    while True:
        choice = torch.empty(len(X), dtype=torch.long)
        for s in range(0, len(X), chunk):                              # (A - B)^2 summed over d, chunked over N for memory
            d = ((X[s:s + chunk, None, :] - c[None, :, :]) ** 2.0).sum(dim=-1)
            choice[s:s + chunk] = torch.argmin(d, dim=1)
        prev = c.clone()
        for k in range(num_clusters):
            sel = X[choice == k]
            if len(sel) == 0 and empty == "keep":
                continue
            c[k] = sel.mean(dim=0)
        shift = torch.sum(torch.sqrt(torch.sum((c - prev) ** 2, dim=1)))
        it += 1
        if not bool(shift == shift):
            raise RuntimeError("kmeans oracle: NaN centre (empty cluster with empty='nan')")
        if float(shift) ** 2 < tol:
            break
        if iter_limit and it >= iter_limit:
            break
    return choice, c, it
