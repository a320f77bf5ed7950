"""Small launches of every kernel family (smoke run on the GPU box; compute-sanitizer is closed on this pool):
   python tools/sanitize_smoke.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import medical_image_editing_b200 as pkg
from medical_image_editing_b200 import _native
from medical_image_editing_b200.src.functions import OneHotEncoder, kmeans
from medical_image_editing_b200.src.functions.embed_loss import cross_loss

dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
# resident / streamed (ragged D: the q boxes of the last chunk are clipped by the TMA) / pair / small / CUDA-core search
for (B, H, K, D, flags) in ((2, 32, 512, 64, 0), (1, 32, 600, 132, 0), (1, 32, 512, 256, _native.VQ_FLAG_PAIR),
                            (1, 32, 4096, 68, 0), (2, 24, 10, 16, 0), (1, 20, 77, 10, 0), (1, 32, 64, 256, 0)):
    m = pkg.VQ(emb_dim=D, dict_size=K, momentum=0.99, eps=1e-5, knn_backend="torch").to(dev)
    m.kernel_flags = flags
    z = torch.randn(B, D, H, H, device=dev, generator=g).requires_grad_(True)
    for train in (True, False):
        m.train(train)
        q, loss, ids = m(z)
        torch.autograd.grad(q.sum() + loss, z)
    lab = torch.randint(0, K, (B, H, H), device=dev, generator=g)
    m.lookup(lab)
    l = cross_loss(z, lab.int(), m.get_codebook())
    torch.autograd.grad(l, z)
    OneHotEncoder(K + 1)(lab)
X = torch.randn(1000, 12, device=dev, generator=g)
kmeans(X, 5, seed=1, iter_limit=3)
kmeans(torch.randn(1024, 16, device=dev, generator=g), 8, seed=1, iter_limit=3)
torch.cuda.synchronize()
print("sanitize smoke done")
