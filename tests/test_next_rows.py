"""k-means codebook initialisation and the one-hot code map.
CPU: the oracles (one-hot pinned to the unmodified reference class through tests/golden/onehot_*.npz; k-means is a
restatement of the absent kmeans-pytorch 0.3.0, pinned to an independent implementation of the same algorithm --
scikit-learn's Lloyd iterations from identical initial centres, tests/golden/kmeans_*.npz from
oracle/make_golden_kmeans.py -- and checked against a plain numpy Lloyd loop).
GPU: the CUDA paths against the oracles and against the scikit-learn golden vectors."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle.kmeans_oracle import kmeans_oracle, initial_centers
from oracle.onehot_oracle import onehot_oracle
from util import ROOT

GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "onehot_*.npz")))
KM_GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "kmeans_*.npz")))
DEV = "cuda:0"


def _blobs(n, d, k, seed, spread=0.05):
    g = torch.Generator().manual_seed(seed)
    centres = torch.randn(k, d, generator=g) * 2.0
    lab = torch.randint(0, k, (n,), generator=g)
    return centres[lab] + spread * torch.randn(n, d, generator=g), centres


# ----------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_onehot_oracle_matches_reference_golden(path):
    d = np.load(path)
    out = onehot_oracle(torch.from_numpy(d["t"]), int(d["n_classes"]))
    assert out.dtype == torch.float32 and torch.equal(out, torch.from_numpy(d["out"]).float())


def test_onehot_golden_present():
    assert len(GOLDEN) == 3


def test_kmeans_oracle_is_lloyd():
    X, _ = _blobs(600, 5, 4, 3)
    c0 = initial_centers(X, 4, seed=11)
    ids, c, it = kmeans_oracle(X, 4, centers=c0)
    x, cc = X.double().numpy(), c0.double().numpy()            # plain numpy Lloyd loop, float64
    for _ in range(it):
        a = ((x[:, None, :] - cc[None]) ** 2).sum(-1).argmin(1)
        cc = np.stack([x[a == k].mean(0) if (a == k).any() else cc[k] for k in range(4)])
    assert np.array_equal(a, ids.numpy())
    assert np.abs(cc - c.double().numpy()).max() < 1e-5
    assert it >= 2


def test_kmeans_golden_present():
    assert len(KM_GOLDEN) == 4


@pytest.mark.parametrize("path", KM_GOLDEN, ids=[os.path.basename(p)[:-4] for p in KM_GOLDEN])
def test_kmeans_oracle_matches_sklearn_golden(path):
    """The restatement of kmeans-pytorch against scikit-learn's Lloyd iterations (independent code, float64): centres after
    every iteration <= 1e-5 relative (max norm), the stopping rule fires at the recorded iteration, labels equal."""
    d = np.load(path)
    X, c0, iters = torch.from_numpy(d["X"]), torch.from_numpy(d["c0"]), int(d["iters"])
    K = c0.shape[0]
    for i in range(1, iters + 1):
        _, c, it = kmeans_oracle(X, K, centers=c0, iter_limit=i)
        ref = d["centers_per_iter"][i - 1]
        assert it == i
        assert np.abs(c.double().numpy() - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max()), (i, path)
    ids, c, it = kmeans_oracle(X, K, centers=c0)
    assert it == iters
    # scikit-learn's labels refer to the final centres; the package's (and the oracle's) to the centres before the last
    # update -- assign once more with the final centres to compare like with like
    lab = torch.cdist(X.double(), c.double()).argmin(1).numpy()
    assert (lab != d["labels_final"]).sum() == 0


def test_kmeans_oracle_matches_live_sklearn():
    """Same pin, live (scikit-learn is in the image): a case that is not in the fixtures."""
    sk = pytest.importorskip("sklearn.cluster")
    import warnings
    X, centres = _blobs(2500, 6, 5, 21, spread=0.8)
    c0 = initial_centers(X, 5, seed=2)
    _, c, it = kmeans_oracle(X, 5, centers=c0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        km = sk.KMeans(n_clusters=5, init=c0.double().numpy(), n_init=1, algorithm="lloyd", tol=0.0, max_iter=it).fit(X.double().numpy())
    assert it >= 3 and np.abs(km.cluster_centers_ - c.double().numpy()).max() <= 1e-5 * max(1.0, np.abs(km.cluster_centers_).max())


def test_kmeans_oracle_recovers_separated_blobs():
    X, centres = _blobs(2000, 8, 6, 5)
    ids, c, _ = kmeans_oracle(X, 6, centers=centres + 0.3)      # start near the truth: converges onto it
    assert torch.cdist(c, centres).min(dim=1).values.max() < 0.05


# ----------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_onehot_cuda_matches_reference_golden(path):
    from medical_image_editing_b200.src.functions import OneHotEncoder
    d = np.load(path)
    t = torch.from_numpy(d["t"]).to(DEV)
    for cast in (lambda x: x, lambda x: x.long(), lambda x: x.float()):
        out = OneHotEncoder(int(d["n_classes"]))(cast(t))
        assert out.dtype == torch.float32 and out.is_contiguous()
        assert torch.equal(out.cpu(), torch.from_numpy(d["out"]).float())


@pytest.mark.gpu
def test_onehot_cuda_large_and_edge_cases():
    from medical_image_editing_b200.src.functions import OneHotEncoder
    g = torch.Generator(device=DEV).manual_seed(1)
    t = torch.randint(0, 11, (16, 512, 512), device=DEV, generator=g)           # run_recon-like code map, +1 shifted
    out = OneHotEncoder(11)(t)
    assert torch.equal(out.argmax(1), t) and float(out.sum()) == t.numel()
    assert torch.equal(out[:, 1:], torch.nn.functional.one_hot(t, 11).permute(0, 3, 1, 2)[:, 1:].float())
    assert OneHotEncoder(5)(torch.zeros((0, 4, 4), dtype=torch.int64, device=DEV)).shape == (0, 5, 4, 4)
    odd = torch.tensor([[0, 3, 7, -1, 2]], device=DEV)                           # out of range -> all-zero columns
    o = OneHotEncoder(4)(odd)
    assert o.shape == (1, 4, 5) and o[0, :, 2].sum() == 0 and o[0, :, 3].sum() == 0 and o[0, 3, 1] == 1
    with pytest.raises(RuntimeError):
        OneHotEncoder(4)(torch.zeros(1, 2, dtype=torch.int64))                  # CPU tensor: no fallback


@pytest.mark.gpu
@pytest.mark.parametrize("n,d,k", [(8192, 16, 10), (4096, 64, 32), (5000, 12, 7), (65536, 8, 4)])   # last: > 4096 points per cluster
def test_kmeans_cuda_matches_oracle(n, d, k):
    """Same initial centres -> same Lloyd iterations: centres within 1e-4, assignments equal.  n = 5000 takes the fp32
    CUDA-core search (N % 128 != 0)."""
    from medical_image_editing_b200.src.functions import kmeans
    X, centres = _blobs(n, d, k, seed=n + d)
    # one initial centre near every blob: a blob split between two centres would put points on a decision boundary, where
    # the package's direct sum (x - c)^2 and the quantiser's 2 x.c - |c|^2 - |x|^2 may legitimately round differently
    c0 = centres + 0.3 * torch.randn(k, d, generator=torch.Generator().manual_seed(3))
    ids_ref, c_ref, it_ref = kmeans_oracle(X, k, centers=c0)
    ids, c = kmeans(X.to(DEV), k, cluster_centers=c0)
    assert not ids.is_cuda and not c.is_cuda                                     # the package returns CPU tensors
    assert torch.equal(ids, ids_ref)
    assert (c - c_ref).abs().max() <= 1e-4 * max(1.0, float(c_ref.abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("path", KM_GOLDEN, ids=[os.path.basename(p)[:-4] for p in KM_GOLDEN])
def test_kmeans_cuda_matches_sklearn_golden(path):
    """The CUDA Lloyd loop against scikit-learn's centres (independent pin of the algorithm): same initial centres, same
    number of iterations (the package's stopping rule), centres <= 1e-4 relative (max norm; fp32 sums against float64),
    labels with respect to the final centres equal except points whose two best distances tie to 1e-5."""
    from medical_image_editing_b200.src.functions import kmeans
    d = np.load(path)
    X, c0, iters = torch.from_numpy(d["X"]), torch.from_numpy(d["c0"]), int(d["iters"])
    K = c0.shape[0]
    ref = d["centers_per_iter"]
    ids, c = kmeans(X.to(DEV), K, cluster_centers=c0, iter_limit=1)
    assert np.abs(c.double().numpy() - ref[0]).max() <= 1e-4 * max(1.0, np.abs(ref[0]).max())
    ids, c = kmeans(X.to(DEV), K, cluster_centers=c0)
    assert np.abs(c.double().numpy() - ref[-1]).max() <= 1e-4 * max(1.0, np.abs(ref[-1]).max()), path
    dist = torch.cdist(X.double(), torch.from_numpy(ref[-1]))
    lab = torch.cdist(X.double(), c.double()).argmin(1).numpy()
    bad = np.nonzero(lab != d["labels_final"])[0]
    top2 = dist.topk(2, dim=1, largest=False).values
    margin = (top2[:, 1] - top2[:, 0]).numpy()
    assert all(margin[i] <= 1e-5 * max(1.0, float(top2[i, 1])) for i in bad), (len(bad), path)
    print(f"[kmeans golden] {os.path.basename(path)}: {iters} iterations, {len(bad)} tie rows")


@pytest.mark.gpu
def test_kmeans_initialize_embed_single_process():
    """`initialize_embed(vq, embed)` (unet_encoder.py:66-91): centres of the encoder output become `vq.embed`."""
    import medical_image_editing_b200 as pkg
    from medical_image_editing_b200.src.functions import initialize_embed, kmeans_nchw
    B, D, H, K = 2, 16, 64, 10
    X, centres = _blobs(B * H * H, D, K, seed=9)
    embed = X.view(B, H, H, D).permute(0, 3, 1, 2).contiguous().to(DEV)
    vq = pkg.VQ(emb_dim=D, dict_size=K, momentum=0.999, eps=1e-5, knn_backend="torch").to(DEV)
    initialize_embed(vq, embed, rank=0, seed=4)
    assert vq.embed.shape == (K, D) and vq.embed.is_cuda
    c2, it = kmeans_nchw(embed, K, seed=4)
    # two runs agree up to the stopping tolerance: the sums are fp32 atomics (order varies), the start is K random pixels, so
    # a point on a decision boundary may flip and the loop may stop one iteration apart (shift^2 < 1e-4)
    assert it >= 1 and (vq.embed - c2).abs().max() < 5e-2
    # every centre is the mean of the pixels assigned to it: the quantiser's own forward reproduces the assignment
    vq.eval()
    q, loss, ids = vq(embed)
    flat = embed.permute(0, 3, 2, 1).reshape(-1, D)                              # (b, w, h) order of `ids`
    for kk in range(K):
        sel = flat[ids.reshape(-1) == kk]
        if len(sel):
            assert (sel.mean(0) - vq.embed[kk]).abs().max() < 2e-2              # fixed point up to the stopping tolerance


def _kmeans_dp_worker(rank, ws, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(ws), RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", rank))
    from medical_image_editing_b200.src.functions import kmeans_nchw
    B, D, H, K = 4, 16, 64, 10
    X, _ = _blobs(B * H * H, D, K, seed=21)
    embed = X.view(B, H, H, D).permute(0, 3, 1, 2).contiguous()
    half = B // ws
    c, it = kmeans_nchw(embed[rank * half:(rank + 1) * half].to(f"cuda:{rank}"), K, seed=5)
    ret[rank] = (c.cpu().numpy(), it)
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_kmeans_two_ranks_match_single_process():
    """Data-parallel Lloyd: each rank assigns its shard, the packed statistics are all-reduced; both ranks end with
    bit-identical centres, equal (up to the stopping tolerance) to one process on the whole batch."""
    import torch.multiprocessing as mp
    from medical_image_editing_b200.src.functions import kmeans_nchw
    ret = mp.Manager().dict()
    mp.spawn(_kmeans_dp_worker, args=(2, 29791, ret), nprocs=2, join=True)
    os.environ.pop("WORLD_SIZE", None)
    os.environ.pop("RANK", None)
    assert np.array_equal(ret[0][0], ret[1][0]) and ret[0][1] == ret[1][1]
    B, D, H, K = 4, 16, 64, 10
    X, _ = _blobs(B * H * H, D, K, seed=21)
    embed = X.view(B, H, H, D).permute(0, 3, 1, 2).contiguous().to(DEV)
    c1, _ = kmeans_nchw(embed, K, seed=5)
    assert np.abs(ret[0][0] - c1.cpu().numpy()).max() < 5e-2
