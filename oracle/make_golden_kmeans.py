"""Golden vectors for the k-means codebook initialisation -- TEST INFRASTRUCTURE ONLY.

The reference calls `kmeans_pytorch.kmeans` (kmeans-pytorch 0.3.0, requirements.txt:52; src/networks/unet_encoder.py:66-91);
the package is absent from /root/reference and from this image, so `oracle/kmeans_oracle.py` restates its published
algorithm.  This script pins that restatement to an INDEPENDENT implementation of the same algorithm that is in the
image: scikit-learn's `KMeans(algorithm="lloyd", init=<given centres>, n_init=1, tol=0, max_iter=n)` -- n plain Lloyd
iterations from identical initial centres.  It writes, per case, the data, the initial centres, and scikit-learn's centres
after n = 1, 2, ... iterations up to the iteration at which the package's stopping rule (shift^2 < 1e-4, restated in
the oracle) fires, plus scikit-learn's labels with respect to the final centres.

    python oracle/make_golden_kmeans.py        # -> tests/golden/kmeans_*.npz (scikit-learn version recorded inside)
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {                       # name: (N, D, K, seed, spread, start offset)
    "kmeans_k10_d16": (8192, 16, 10, 8208, 0.05, 0.3),            # run_recon's codebook shape
    "kmeans_k7_d12_ragged": (5000, 12, 7, 5012, 0.05, 0.3),       # N % 128 != 0
    "kmeans_k32_d64": (4096, 64, 32, 4160, 0.05, 0.3),
    "kmeans_k6_d8_overlap": (3000, 8, 6, 77, 1.2, None),          # overlapping blobs, start = K random rows: many iterations
}


def blobs(n, d, k, seed, spread):
    g = torch.Generator().manual_seed(seed)
    centres = torch.randn(k, d, generator=g) * 2.0
    lab = torch.randint(0, k, (n,), generator=g)
    return centres[lab] + spread * torch.randn(n, d, generator=g), centres


def sklearn_lloyd(X, c0, iters):
    from sklearn.cluster import KMeans
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                           # ConvergenceWarning: max_iter is the point
        km = KMeans(n_clusters=len(c0), init=c0.astype(np.float64), n_init=1, algorithm="lloyd", tol=0.0, max_iter=iters)
        km.fit(X.astype(np.float64))
    return km.cluster_centers_, km.labels_, km.n_iter_


def main():
    import sklearn
    from oracle.kmeans_oracle import kmeans_oracle
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, (n, d, k, seed, spread, off) in CASES.items():
        X, centres = blobs(n, d, k, seed, spread)
        if off is None:                                           # the package's own start: K distinct rows of X
            from oracle.kmeans_oracle import initial_centers
            c0 = initial_centers(X, k, seed=seed)
        else:
            c0 = centres + off * torch.randn(k, d, generator=torch.Generator().manual_seed(3))
        _, _, it = kmeans_oracle(X, k, centers=c0)               # iteration at which the package's stopping rule fires
        per_iter = []
        for i in range(1, it + 1):
            c, lab, n_iter = sklearn_lloyd(X.numpy(), c0.numpy(), i)
            per_iter.append(c)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), X=X.numpy(), c0=c0.numpy(), iters=np.int64(it),
                            centers_per_iter=np.stack(per_iter).astype(np.float64), labels_final=lab.astype(np.int64),
                            sklearn_version=np.array(sklearn.__version__))
        print(name, "iterations", it, "sklearn n_iter_", n_iter)


if __name__ == "__main__":
    main()
