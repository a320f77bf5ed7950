"""Fallback rows and search-kernel time per training step, bench.py's config (warmed EMA state, noise input).
    python tools/fb_steps.py [D K steps]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import medical_image_editing_b200 as pkg
from medical_image_editing_b200.src.functions import vq_function as vf
D = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = int(sys.argv[2]) if len(sys.argv) > 2 else 512
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 24
B, H, dev = 16, 256, "cuda:0"
N = B * H * H
L = pkg.lib()
gen = torch.Generator(device=dev).manual_seed(1234)
zb = [torch.randn(B, D, H, H, device=dev, generator=gen) for _ in range(4)]
vq = pkg.VQ(emb_dim=D, dict_size=K, momentum=0.99, eps=1e-5, knn_backend="torch").to(dev)
vq.train(True)
with torch.no_grad():
    cs = torch.rand(K, generator=torch.Generator().manual_seed(1234)) * (N / K) + 1.0
    vq.cluster_size.copy_(cs.to(dev))
    vq.embed_avg.copy_((vq.embed * vq.cluster_size[:, None]).T)
for i in range(steps):
    L.vq_profile_enable(1); L.vq_profile_read(None, None)
    with torch.no_grad():
        q, loss, ids = vq(zb[i % 4])
    torch.cuda.synchronize()
    tot, nl = ctypes.c_double(0), ctypes.c_int(0)
    L.vq_profile_read(ctypes.byref(tot), ctypes.byref(nl)); L.vq_profile_enable(0)
    wsb = vf._WORKSPACES.get((0, torch.cuda.current_stream().cuda_stream))
    fb = int(L.vq_debug_fallback_rows(wsb.data_ptr(), N, K, D, torch.cuda.current_stream().cuda_stream))
    nrm = vq.embed.norm(dim=1)
    print(f"step {i:2d}: kernel {tot.value:.3f} ms, fallback rows {fb:6d}, loss {loss.item():.4f}, |e| min/med/max {nrm.min().item():.3f} {nrm.median().item():.3f} {nrm.max().item():.3f}, used codes {int((torch.bincount(ids.flatten(), minlength=K) > 0).sum())}")
