"""Fallback-row rate of the tensor-core search for a shape: python tools/fb_rate.py B D H K [kind]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import medical_image_editing_b200 as pkg
from medical_image_editing_b200.src.functions import vq_function as vf
B, D, H, K = (int(x) for x in sys.argv[1:5])
kind = sys.argv[5] if len(sys.argv) > 5 else "gauss"
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(3)
L = pkg.lib()
z = torch.randn(B, D, H, H, device=dev, generator=g)
E = torch.randn(K, D, device=dev, generator=g)
if kind == "relu":
    z = torch.relu(z)
if kind == "clustered":
    idx = torch.randint(0, K, (B, H, H), device=dev, generator=g)
    z = E[idx].permute(0, 3, 1, 2).contiguous() + 0.1 * torch.randn(B, D, H, H, device=dev, generator=g)
m = pkg.VQ(emb_dim=D, dict_size=K, momentum=0.99, eps=1e-5, knn_backend="torch").to(dev)
with torch.no_grad():
    m.embed.copy_(E)
m.eval()
with torch.no_grad():
    q, loss, ids = m(z)
torch.cuda.synchronize()
wsb = list(vf._WORKSPACES.values())[0]
N = B * H * H
fb = L.vq_debug_fallback_rows(wsb.data_ptr(), N, K, D, torch.cuda.current_stream().cuda_stream)
m2 = pkg.VQ(emb_dim=D, dict_size=K, momentum=0.99, eps=1e-5, knn_backend="torch").to(dev)
m2.kernel_flags = 1
with torch.no_grad():
    m2.embed.copy_(E)
m2.eval()
with torch.no_grad():
    q2, loss2, ids2 = m2(z)
print(f"B={B} D={D} H={H} K={K} {kind}: path {L.vq_assign_path(B, D, H, H, K, 0)} fallback rows {fb} of {N} ({100.0 * fb / N:.3f} %), "
      f"ids equal to CUDA-core path: {bool(torch.equal(ids, ids2))}")
