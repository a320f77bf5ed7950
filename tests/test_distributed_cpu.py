"""World-size-2 `gloo` tests (CPU) of the host-side multi-GPU logic of the product path:

  * `functions.vq_function._all_reduce_stats`: the ONE packed all-reduce of the EMA statistics and the
    (count_scale, sum_scale) it hands to `vq_ema_update`, for the three `reduce_mode`s.  The EMA arithmetic itself
    is CUDA-only, so it is restated here in numpy (vq_module.py:194-199) and checked against the 2-rank golden
    vectors produced by the unmodified reference (`tests/golden/ddp2_reference_semantics.npz`);
  * `trainers.ddp`: broadcast of parameters/buffers, the flat gradient all-reduce, replica-sync check.

No CUDA kernel runs here (the quantiser itself has no CPU path); the GPU side of the N > 1 path is exercised by
`bench.py --gpus N` and `tests/test_parity_gpu.py::test_two_rank_nccl_matches_single_process` on multi-GPU boxes."""
import os

import numpy as np
import torch

from util import load_golden


def _pack_stats(L, ids_flat, flat, K, D):
    """The packed statistics buffer `vq_assign_fwd` produces: [cnt_hi K | cnt_lo K | pad | sums K*D]."""
    counts = np.bincount(ids_flat, minlength=K).astype(np.int64)
    sums = np.zeros((K, D), np.float64)
    np.add.at(sums, ids_flat, flat.astype(np.float64))
    off = L.vq_stats_sums_offset(K)
    st = np.zeros(L.vq_stats_floats(K, D), np.float32)
    st[:K] = counts >> 12
    st[K:2 * K] = counts & 4095
    st[off:] = sums.astype(np.float32).reshape(-1)
    return st


def _ema_numpy(cs, avg, st, K, D, off, momentum, eps, cscale, sscale):
    """vq_module.py:132-136, 194-199 in fp32 numpy, from the packed (already reduced) statistics."""
    f = np.float32
    cnt = (st[:K] * f(4096) + st[K:2 * K]) * f(cscale)
    sums = st[off:].reshape(K, D) * f(sscale)
    cs1 = cs * f(momentum) + f(1 - momentum) * cnt
    avg1 = avg * f(momentum) + f(1 - momentum) * sums.T
    n = cs1.sum(dtype=np.float32)
    smooth = n * (cs1 + f(eps)) / (n + f(K * eps))
    return cs1, avg1, (avg1 / smooth[None, :]).T


def _stats_worker(rank, ws, port, mode, ret):
    import torch.distributed as dist
    os.environ.update(WORLD_SIZE=str(ws), RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from medical_image_editing_b200 import _native
    from medical_image_editing_b200.src.functions.vq_function import _all_reduce_stats
    L = _native.lib()                                    # loads without a GPU; only size helpers are called
    g = load_golden("ddp2_reference_semantics")
    B, D, H, K, _ = [int(x) for x in g["meta"]]
    z = g["z"][rank * B:(rank + 1) * B]
    ids = g[f"r{rank}_ids"]                              # reference layout [b, w, h]
    flat = np.transpose(z, (0, 3, 2, 1)).reshape(-1, D)  # (b, w, h) row order, vq_module.py:171
    st = torch.from_numpy(_pack_stats(L, ids.reshape(-1), flat, K, D))
    cscale, sscale = _all_reduce_stats(st, K, mode)
    cs1, avg1, emb1 = _ema_numpy(g["cluster_size0"], g["embed_avg0"], st.numpy(), K, D, L.vq_stats_sums_offset(K),
                                 0.9, 1e-5, cscale, sscale)
    ret[rank] = dict(cluster_size1=cs1, embed_avg1=avg1, embed1=emb1, scales=(cscale, sscale))
    dist.destroy_process_group()


def _spawn(fn, port, *args):
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(fn, args=(2, port) + args + (ret,), nprocs=2, join=True)
    os.environ.pop("WORLD_SIZE", None)
    os.environ.pop("RANK", None)
    return {r: ret[r] for r in range(2)}


def _close(a, b, tol=2e-6):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


def test_packed_stats_allreduce_reference_mode_matches_reference_on_two_ranks():
    g = load_golden("ddp2_reference_semantics")
    ret = _spawn(_stats_worker, 29751, "reference")
    for r in range(2):
        assert ret[r]["scales"] == (1.0, 0.5)            # rank-local counts, averaged sums (vq_module.py:188-192)
        for k in ("cluster_size1", "embed_avg1", "embed1"):
            assert _close(ret[r][k], g[f"r{r}_{k}"]), k
    assert not np.allclose(ret[0]["cluster_size1"], ret[1]["cluster_size1"])


def test_packed_stats_allreduce_sum_mode_equals_single_process():
    from oracle.vq_oracle import OracleVQ
    from util import set_state, t
    g = load_golden("ddp2_reference_semantics")
    B, D, H, K, _ = [int(x) for x in g["meta"]]
    ret = _spawn(_stats_worker, 29753, "sum")
    m = OracleVQ(D, K, 0.9, 1e-5)
    set_state(m, g["embed0"], g["cluster_size0"], g["embed_avg0"])
    m.train(True)
    m(t(g["z"]))
    for r in range(2):
        assert ret[r]["scales"] == (1.0, 1.0)
        assert _close(ret[r]["cluster_size1"], m.cluster_size.numpy())
        assert _close(ret[r]["embed_avg1"], m.embed_avg.numpy())
        assert _close(ret[r]["embed1"], m.embed.numpy())
    for k in ("cluster_size1", "embed_avg1", "embed1"):  # replicas stay identical without any broadcast
        assert np.array_equal(ret[0][k], ret[1][k])


def test_packed_stats_allreduce_mean_mode():
    ret = _spawn(_stats_worker, 29755, "mean")
    assert ret[0]["scales"] == (0.5, 0.5) and np.array_equal(ret[0]["embed1"], ret[1]["embed1"])


def test_count_halves_stay_exact_beyond_2_pow_24():
    """cnt = hi * 4096 + lo: both halves are exactly representable in fp32 after a sum over many ranks."""
    c = np.array([0, 1, 4095, 4096, 2 ** 24 + 1, 2 ** 31 - 1], np.int64)
    hi, lo = (c >> 12).astype(np.float32), (c & 4095).astype(np.float32)
    ranks = 8
    tot = (hi * ranks).astype(np.float64) * 4096 + (lo * ranks).astype(np.float64)
    assert np.array_equal(tot, (c * ranks).astype(np.float64))


# ---------------------------------------------------------------------------------------------
# trainers.ddp on CPU tensors
# ---------------------------------------------------------------------------------------------
class _ToyModel(torch.nn.Module):
    """Stands in for VQ-W-Net on the CPU: same output dict, a buffer that ranks must share, no quantiser kernels."""

    def __init__(self):
        super().__init__()
        self.enc = torch.nn.Conv2d(1, 4, 3, padding=1)
        self.dec = torch.nn.Conv2d(4, 1, 3, padding=1)
        self.register_buffer("codebook_t", torch.randn(6, 4).T.clone())     # non-contiguous like embed_avg

    def forward(self, x):
        h = torch.relu(self.enc(x))
        return {"recon": torch.tanh(self.dec(h)), "commit_loss": h.pow(2).mean(), "ids": None}


def _trainer_worker(rank, ws, port, ret):
    import torch.distributed as dist
    os.environ.update(WORLD_SIZE=str(ws), RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from medical_image_editing_b200.src.trainers import DataParallelVQTrainer
    torch.manual_seed(100 + rank)                        # different initial weights: the trainer must broadcast rank 0's
    model = _ToyModel()
    tr = DataParallelVQTrainer(model, lr=1e-2, commit_weight=0.5)
    start_sync = tr.replicas_in_sync()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(4, 1, 8, 8, generator=g)
    for _ in range(3):
        out = tr.training_step(x[rank * 2:(rank + 1) * 2])
    ret[rank] = dict(params=[p.detach().numpy().copy() for p in model.parameters()], start_sync=start_sync,
                     end_sync=tr.replicas_in_sync(), loss=float(out["loss"]),
                     buf=model.codebook_t.numpy().copy())
    dist.destroy_process_group()


def test_trainer_two_ranks_equal_single_process_on_concatenated_batch():
    ret = _spawn(_trainer_worker, 29757)
    assert ret[0]["start_sync"] and ret[0]["end_sync"] and ret[1]["end_sync"]
    for a, b in zip(ret[0]["params"], ret[1]["params"]):
        assert np.array_equal(a, b)
    assert np.array_equal(ret[0]["buf"], ret[1]["buf"])
    # single process, same initial weights (rank 0's seed), whole batch
    from medical_image_editing_b200.src.trainers import DataParallelVQTrainer
    torch.manual_seed(100)
    model = _ToyModel()
    tr = DataParallelVQTrainer(model, lr=1e-2, commit_weight=0.5)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(4, 1, 8, 8, generator=g)
    for _ in range(3):
        tr.training_step(x)
    for a, p in zip(ret[0]["params"], model.parameters()):
        assert np.allclose(a, p.detach().numpy(), rtol=1e-4, atol=1e-6)


def _trainer_micro_worker(rank, ws, port, ret):
    import torch.distributed as dist
    os.environ.update(WORLD_SIZE=str(ws), RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from medical_image_editing_b200.src.trainers import DataParallelVQTrainer
    torch.manual_seed(100 + rank)
    model = _ToyModel()
    tr = DataParallelVQTrainer(model, lr=1e-2, commit_weight=0.5)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(8, 1, 8, 8, generator=g)
    for _ in range(3):
        tr.training_step(x[rank * 4:(rank + 1) * 4], micro_batches=2)      # 2 ranks x 2 micro-batches of 2 slices
    ret[rank] = dict(params=[p.detach().numpy().copy() for p in model.parameters()], end_sync=tr.replicas_in_sync())
    dist.destroy_process_group()


def test_trainer_micro_batches_equal_one_shot_step():
    """BASELINE config 5 at 2 / 4 GPUs: the shard is processed in micro-batches, gradients accumulate in the buckets
    and are exchanged once -- same parameters as one process on the whole batch."""
    ret = _spawn(_trainer_micro_worker, 29759)
    assert ret[0]["end_sync"] and ret[1]["end_sync"]
    for a, b in zip(ret[0]["params"], ret[1]["params"]):
        assert np.array_equal(a, b)
    from medical_image_editing_b200.src.trainers import DataParallelVQTrainer
    torch.manual_seed(100)
    model = _ToyModel()
    tr = DataParallelVQTrainer(model, lr=1e-2, commit_weight=0.5)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(8, 1, 8, 8, generator=g)
    for _ in range(3):
        tr.training_step(x)
    for a, p in zip(ret[0]["params"], model.parameters()):
        assert np.allclose(a, p.detach().numpy(), rtol=1e-4, atol=1e-6)
    # single process, micro-batched: also the same
    torch.manual_seed(100)
    model2 = _ToyModel()
    tr2 = DataParallelVQTrainer(model2, lr=1e-2, commit_weight=0.5)
    for _ in range(3):
        tr2.training_step(x, micro_batches=4)
    for p2, p in zip(model2.parameters(), model.parameters()):
        assert np.allclose(p2.detach().numpy(), p.detach().numpy(), rtol=1e-4, atol=1e-6)


def test_trainer_default_optimizer_is_fused_only_on_cuda():
    """The trainer's default Adam is the fused multi-tensor kernel only when every parameter lives on a GPU (DESIGN
    section 5); on the CPU (these tests, the gloo ranks) it stays torch's stock implementation."""
    import torch
    from medical_image_editing_b200.src.trainers.ddp import DataParallelVQTrainer
    net = torch.nn.Sequential(torch.nn.Conv2d(1, 2, 3, padding=1))
    tr = DataParallelVQTrainer(net, lr=1e-3)
    assert isinstance(tr.optimizer, torch.optim.Adam)
    assert not tr.optimizer.defaults.get("fused")
    own = torch.optim.SGD(net.parameters(), lr=0.1)
    assert DataParallelVQTrainer(net, optimizer=own).optimizer is own
