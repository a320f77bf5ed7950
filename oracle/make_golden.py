"""Generate tests/golden/*.npz by running the UNMODIFIED reference VQModule -- TEST INFRASTRUCTURE.

Run in the build container (where /root/reference exists):

    python oracle/make_golden.py

The reference has no golden vectors of its own, so these are the
pin: outputs of `/root/reference/src/networks/vq/vq_module.py:VQModule` (imported via
`oracle/ref_loader.py`, CPU, fp32, torch as installed) on seeded inputs.  Small cases
store inputs and every output; the config-1 sized case regenerates its inputs from the
seed at test time and stores only ids / counts / loss / EMA buffers.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_loader import load_reference_vq  # noqa: E402
from oracle.vq_oracle import seeded_case  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# name: (B, D, H, K, kind, training, warmed, seed, momentum, steps)
SMALL_CASES = {
    "tiny_k10_d16_train":      (2, 16, 8, 10, "gauss", True, False, 11, 0.999, 1),
    "k512_d64_train_warm":     (1, 64, 16, 512, "gauss", True, True, 12, 0.99, 1),
    "k512_d64_train_cold":     (1, 64, 16, 512, "gauss", True, False, 13, 0.99, 1),
    "k64_d256_eval":           (1, 256, 8, 64, "gauss", False, False, 14, 0.99, 1),
    "k512_d64_relu_train":     (2, 64, 16, 512, "relu", True, True, 15, 0.99, 1),
    "k512_d64_clustered_eval": (1, 64, 16, 512, "clustered", False, False, 16, 0.99, 1),
    "k100_d24_ragged_train":   (3, 24, 12, 100, "gauss", True, True, 17, 0.9, 1),
    "k10_d16_multistep":       (2, 16, 8, 10, "gauss", True, False, 18, 0.99, 3),
    "k64_d512_vqgan_eval":     (1, 512, 4, 64, "gauss", False, False, 19, 0.99, 1),
}


def build_ref(K, D, embed, momentum, warmed, n):
    VQ = load_reference_vq()
    m = VQ(emb_dim=D, dict_size=K, momentum=momentum, eps=1e-5, knn_backend="torch")
    with torch.no_grad():
        m.embed.copy_(embed)
        m.embed_avg.copy_(embed.T)
        if warmed:
            g = torch.Generator().manual_seed(99)
            cs = torch.rand(K, generator=g) * (n / K) + 1.0
            m.cluster_size.copy_(cs)
            m.embed_avg.copy_(embed.T * cs.unsqueeze(0))
    return m


def run_case(B, D, H, K, kind, training, warmed, seed, momentum, steps):
    out = {}
    g = torch.Generator().manual_seed(seed + 1000)
    z0, embed = seeded_case(B, D, H, H, K, seed=seed, kind=kind)
    m = build_ref(K, D, embed, momentum, warmed, B * H * H)
    m.train(training)
    out["embed0"] = m.embed.numpy().copy()
    out["cluster_size0"] = m.cluster_size.numpy().copy()
    out["embed_avg0"] = m.embed_avg.numpy().copy()
    out["meta"] = np.array([B, D, H, K, int(training), steps], dtype=np.int64)
    out["momentum"] = np.array([momentum], dtype=np.float64)
    for s in range(steps):
        if s == 0:
            z = z0.clone()
        else:
            z = torch.randn(B, D, H, H, generator=g)
        g_q = torch.randn(B, D, H, H, generator=g)
        w = 0.25 + 0.5 * s
        z.requires_grad_(True)
        q, loss, ids = m(z)
        total = (q * g_q).sum() + w * loss
        (g_z,) = torch.autograd.grad(total, z)
        sfx = "" if s == 0 else f"_s{s}"
        out["z" + sfx] = z.detach().numpy().copy()
        out["g_q" + sfx] = g_q.numpy().copy()
        out["w" + sfx] = np.array([w], dtype=np.float64)
        out["q" + sfx] = q.detach().contiguous().numpy().copy()
        out["loss" + sfx] = np.array([loss.item()], dtype=np.float32)
        out["ids" + sfx] = ids.numpy().copy()
        out["g_z" + sfx] = g_z.numpy().copy()
        out["embed1" + sfx] = m.embed.numpy().copy()
        out["cluster_size1" + sfx] = m.cluster_size.numpy().copy()
        out["embed_avg1" + sfx] = m.embed_avg.numpy().copy()
    return out


def run_config1():
    """BASELINE config 1 quantiser shape: 1x64x256x256, K=512; inputs regenerated from the seed."""
    B, D, H, K = 1, 64, 256, 512
    z, embed = seeded_case(B, D, H, H, K, seed=1234, kind="gauss")
    out = {"meta": np.array([B, D, H, K, 1, 1], dtype=np.int64), "seed": np.array([1234])}
    m = build_ref(K, D, embed, 0.99, True, B * H * H)
    m.train(True)
    q, loss, ids = m(z)
    out["ids_i16"] = ids.numpy().astype(np.int16)
    out["loss"] = np.array([loss.item()], dtype=np.float32)
    out["counts"] = np.bincount(ids.numpy().ravel(), minlength=K).astype(np.int32)
    out["embed1"] = m.embed.numpy().copy()
    out["cluster_size1"] = m.cluster_size.numpy().copy()
    out["embed_avg1"] = m.embed_avg.numpy().copy()
    out["q_checksum"] = np.array([q.double().sum().item(), q.double().pow(2).sum().item()])
    return out


def _ddp_worker(rank, ws, port, ret):
    import torch.distributed as dist
    os.environ.update(WORLD_SIZE=str(ws), RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    B, D, H, K = 2, 16, 8, 32
    z, embed = seeded_case(B * ws, D, H, H, K, seed=77, kind="gauss")
    m = build_ref(K, D, embed, 0.9, True, B * ws * H * H)
    m.train(True)
    zr = z[rank * B:(rank + 1) * B].clone()
    q, loss, ids = m(zr)
    ret[rank] = dict(ids=ids.numpy().copy(), loss=float(loss), embed1=m.embed.numpy().copy(),
                     cluster_size1=m.cluster_size.numpy().copy(), embed_avg1=m.embed_avg.numpy().copy())
    dist.destroy_process_group()


def run_ddp2():
    """2-rank gloo run of the reference: pins the 'as-written' multi-rank semantics (A)."""
    import torch.multiprocessing as mp
    ws = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ddp_worker, args=(ws, 29731, ret), nprocs=ws, join=True)
    os.environ.pop("WORLD_SIZE", None)
    B, D, H, K = 2, 16, 8, 32
    z, embed = seeded_case(B * ws, D, H, H, K, seed=77, kind="gauss")
    out = {"meta": np.array([B, D, H, K, ws], dtype=np.int64), "z": z.numpy(), "embed0": embed.numpy()}
    g = torch.Generator().manual_seed(99)
    cs = torch.rand(K, generator=g) * (B * ws * H * H / K) + 1.0
    out["cluster_size0"] = cs.numpy()
    out["embed_avg0"] = (embed.T * cs.unsqueeze(0)).numpy()
    for r in range(ws):
        for k, v in ret[r].items():
            out[f"r{r}_{k}"] = np.asarray(v)
    return out


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for name, spec in SMALL_CASES.items():
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **run_case(*spec))
        print("wrote", name)
    np.savez_compressed(os.path.join(GOLD, "config1_k512_d64_256.npz"), **run_config1())
    print("wrote config1")
    np.savez_compressed(os.path.join(GOLD, "ddp2_reference_semantics.npz"), **run_ddp2())
    print("wrote ddp2")


if __name__ == "__main__":
    main()
