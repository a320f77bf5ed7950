"""Timing of the SURVEY 8(f) rows 2-3 on the GPU box: one-hot write bandwidth and k-means (Lloyd) iterations, with the CPU
oracles beside them (bounded samples).   python tools/next_rows_bench.py"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from medical_image_editing_b200.src.functions import OneHotEncoder, kmeans_nchw
from oracle.kmeans_oracle import kmeans_oracle, initial_centers
from oracle.onehot_oracle import onehot_oracle

dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)


def gpu_ms(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


# one-hot: run_recon code map (K = 10 -> 11 classes, 16 x 512 x 512) and the config-2 map (513 classes, 16 x 256 x 256: 2.2 GB)
for B, H, C in ((16, 512, 11), (16, 256, 513)):
    t = torch.randint(0, C, (B, H, H), device=dev, generator=g).int()
    enc = OneHotEncoder(C)
    ms = gpu_ms(lambda: enc(t))
    by = B * H * H * (C * 4 + 4)
    tc = t[:1].cpu()
    t0 = time.perf_counter(); onehot_oracle(tc, C); cpu = time.perf_counter() - t0
    print(f"onehot B={B} H={H} C={C}: {ms:.3f} ms, {by / ms / 1e6:.0f} GB/s written+read; CPU oracle {cpu * 1e3 * B:.0f} ms (scaled from 1 slice)")

# k-means: encoder output of run_recon (D = 16, K = 10) on 4 slices of 512 x 512, fixed number of Lloyd iterations
B, D, H, K = 4, 16, 512, 10
centres = torch.randn(K, D, device=dev, generator=g) * 2
lab = torch.randint(0, K, (B, H, H), device=dev, generator=g)
embed = (centres[lab] + 0.3 * torch.randn(B, H, H, D, device=dev, generator=g)).permute(0, 3, 1, 2).contiguous()
torch.cuda.synchronize()
t0 = time.perf_counter(); c, it = kmeans_nchw(embed, K, seed=1, iter_limit=10); torch.cuda.synchronize(); t_gpu = time.perf_counter() - t0
t0 = time.perf_counter(); c, it = kmeans_nchw(embed, K, seed=1, iter_limit=10); torch.cuda.synchronize(); t_gpu = time.perf_counter() - t0
X = embed[:1].permute(0, 2, 3, 1).reshape(-1, D).cpu()
t0 = time.perf_counter(); kmeans_oracle(X, K, centers=initial_centers(X, K, seed=1), iter_limit=3); t_cpu = (time.perf_counter() - t0) / 3 * B
print(f"kmeans N={B * H * H} D={D} K={K}: {t_gpu / it * 1e3:.2f} ms per Lloyd iteration ({it} iterations, host-synchronised stopping test); "
      f"CPU oracle {t_cpu * 1e3:.0f} ms per iteration (scaled from 1 slice, {torch.get_num_threads()} threads)")
