"""EmbeddingLoss forward + backward once at config 2 (for an `ncu --metrics gpu__time_duration.sum,dram__bytes...` pass)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import medical_image_editing_b200 as pkg
from medical_image_editing_b200._native import check

L = pkg.lib()
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
g = torch.Generator(device=dev).manual_seed(11)
S = torch.cuda.current_stream().cuda_stream
p = lambda t: t.data_ptr()
B, D, H, K = 16, 64, 256, 512
z = torch.randn(B, D, H, H, device=dev, generator=g)
gz = torch.empty_like(z)
E = torch.randn(K, D, device=dev, generator=g)
gl = torch.ones((), device=dev)
small = torch.randint(0, K + 1, (B, H // 8, H // 8), device=dev, generator=g)
lab = small.repeat_interleave(8, 1).repeat_interleave(8, 2).to(torch.int32).contiguous()
loss = torch.empty((), device=dev)
w = torch.empty(B * K, device=dev)
work = torch.empty(max(L.vq_embed_loss_work_bytes(B, K), 256), dtype=torch.uint8, device=dev)
for _ in range(2):
    check(L.vq_embed_loss_fwd(p(z), p(lab), p(E), B, D, H, H, K, p(loss), p(w), p(work), work.numel(), S), "el_fwd")
    check(L.vq_embed_loss_bwd(p(gl), p(z), p(lab), p(E), p(w), p(gz), B, D, H, H, K, S), "el_bwd")
torch.cuda.synchronize()
print("done", float(loss.item()))
