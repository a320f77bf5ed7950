"""Quick A/B timing of one build of libvq_b200.so (select with VQ_B200_LIB=...): search-kernel time (library events)
and whole-forward time, training and eval mode.   python tools/ab.py [D K B]"""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import medical_image_editing_b200 as pkg

D = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = int(sys.argv[2]) if len(sys.argv) > 2 else 512
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16
data = sys.argv[4] if len(sys.argv) > 4 else "noise"      # noise | clustered (z = code + 0.1 noise) | relu
H = int(os.environ.get("ABH", 256))      # ABH=512: 512 x 512 slices
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)
L = pkg.lib()
m = pkg.VQ(emb_dim=D, dict_size=K, momentum=0.99, eps=1e-5, knn_backend="torch").to(dev)
if os.environ.get("ABSIMT") == "1":       # force the fp32 CUDA-core search
    m.kernel_flags = 1
N = B * H * H
with torch.no_grad():
    cs = torch.rand(K, generator=torch.Generator().manual_seed(1234)) * (N / K) + 1.0
    m.cluster_size.copy_(cs.to(dev))
    m.embed_avg.copy_((m.embed * m.cluster_size[:, None]).T)
if data == "clustered":
    def mk():
        idx = torch.randint(0, K, (B, H, H), device=dev, generator=g)
        return (m.embed.detach()[idx].permute(0, 3, 1, 2) + 0.1 * torch.randn(B, D, H, H, device=dev, generator=g)).contiguous()
    z = [mk() for _ in range(4)]
elif data == "relu":
    z = [torch.relu(torch.randn(B, D, H, H, device=dev, generator=g)) for _ in range(4)]
else:
    z = [torch.randn(B, D, H, H, device=dev, generator=g) for _ in range(4)]
out = []
for train in (True, False):
    m.train(train)
    with torch.no_grad():
        for i in range(3):
            m(z[i % 4])
        torch.cuda.synchronize()
        L.vq_profile_enable(1)
        L.vq_profile_read(None, None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 10
        for i in range(n):
            m(z[i % 4])
        e1.record()
        torch.cuda.synchronize()
        tot, nl = ctypes.c_double(0), ctypes.c_int(0)
        L.vq_profile_read(ctypes.byref(tot), ctypes.byref(nl))
        L.vq_profile_enable(0)
        out.append(f"{'train' if train else 'eval '} kernel {tot.value / max(nl.value, 1):.4f} ms fwd {e0.elapsed_time(e1) / n:.4f} ms")
print(os.environ.get("VQ_B200_LIB", "default"), f"D={D} K={K} {data}:", " | ".join(out))
