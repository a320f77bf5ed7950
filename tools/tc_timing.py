"""Phase timing of the tensor-core epilogue.  Build first with
   VQ_EXTRA_FLAGS=-DVQ_TC_TIMING FORCE=1 bash medical_image_editing_b200/csrc/build.sh
then run on the GPU box:  python tools/tc_timing.py [D K B train noise|clustered]"""
import ctypes
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import medical_image_editing_b200 as pkg

D = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = int(sys.argv[2]) if len(sys.argv) > 2 else 512
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16
train = (sys.argv[4] != "0") if len(sys.argv) > 4 else True
data = sys.argv[5] if len(sys.argv) > 5 else "noise"        # noise | clustered (z = code + 0.1 noise)
H = 256
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)
L = pkg.lib()
m = pkg.VQ(emb_dim=D, dict_size=K, momentum=0.99, eps=1e-5, knn_backend="torch").to(dev)
with torch.no_grad():
    m.cluster_size.fill_(2048.0)
    m.embed_avg.copy_(m.embed.T * 2048.0)       # consistent EMA state: embed == embed_avg / cluster_size
m.train(train)
if data == "clustered":
    z = [(m.embed.detach()[torch.randint(0, K, (B, H, H), device=dev, generator=g)].permute(0, 3, 1, 2)
          + 0.1 * torch.randn(B, D, H, H, device=dev, generator=g)).contiguous() for _ in range(3)]
else:
    z = [torch.randn(B, D, H, H, device=dev, generator=g) for _ in range(3)]
with torch.no_grad():
    for i in range(3):
        m(z[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    m(z[0])
    e1.record()
    torch.cuda.synchronize()
print("forward ms", e0.elapsed_time(e1))
buf = np.zeros(148 * 16 * 8, dtype=np.int64)
n = L.vq_debug_tc_timing(buf.ctypes.data_as(ctypes.c_void_p), buf.size)
print("timing entries", n)
if n > 0:
    t = buf.reshape(148, 16, 8).astype(np.float64)
    tiles = t[:, :, 6].mean()
    print(f"tiles per CTA {tiles:.1f}")
    streamed = K * D * 4 > 140 * 1024
    if streamed:      # streamed-codebook kernel: 16 epilogue warps in two teams, every warp does scan + outputs for its tiles
        names = ["wait |z|^2", "wait tmem", "scan work", "team barrier+merge", "re-rank", "outputs+loop"]
        r = t[:, :, :]
        tiles_w = r[:, :, 6] / 2.0          # each team takes every other tile
        print(f"epilogue warps: cycles per (own) tile {(r[:, :, :6].sum(axis=2) / tiles_w).mean():.0f}")
        for i, nm in enumerate(names):
            per = r[:, :, i] / tiles_w
            print(f"  {nm:20s} mean {per.mean():8.0f}  min {per.min():8.0f}  max {per.max():8.0f}")
        sys.exit(0)
    for role, sl, names in (("scan warps", slice(0, 8), ["wait |z|^2", "wait tmem", "scan work", "wait pub slot", "publish+loop"]),
                            ("output warps", slice(8, 16), ["wait z/|z|^2", "wait scan", "merge", "pairs+rerank", "outputs"])):
        r = t[:, sl, :]
        tot = r[:, :, :5].sum(axis=2).mean() / tiles
        print(f"{role}: cycles per tile {tot:.0f}")
        for i, nm in enumerate(names):
            per = r[:, :, i] / r[:, :, 6]
            print(f"  {nm:14s} mean {per.mean():8.0f}  min {per.min():8.0f}  max {per.max():8.0f}")
    print(f"re-rank pairs per output warp and tile: {(t[:, 8:, 7] / t[:, 8:, 6]).mean():.2f}")
