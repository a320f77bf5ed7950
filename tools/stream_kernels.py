"""Run the memory-bound kernels once each at benchmark sizes (for an `ncu --metrics dram__bytes...` pass):
backward at config 2, lookup and small-codebook inference at the run_recon shape, EmbeddingLoss forward/backward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import medical_image_editing_b200 as pkg
from medical_image_editing_b200.src.functions.embed_loss import cross_loss

dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(5)
for rep in range(3):
    # config 2: quantiser forward + backward (vq_bwd_vec_kernel)
    m = pkg.VQ(emb_dim=64, dict_size=512, momentum=0.99, eps=1e-5, knn_backend="torch").to(dev)
    z = torch.randn(16, 64, 256, 256, device=dev, generator=g).requires_grad_(True)
    q, loss, ids = m(z)
    torch.autograd.grad(q.sum() + loss, z)
    # run_recon shape: lookup (vq_lookup_nchw_t_kernel) and inference forward (vq_assign_small_kernel)
    r = pkg.VQ(emb_dim=16, dict_size=10, momentum=0.999, eps=1e-5, knn_backend="torch").to(dev)
    lab = torch.randint(0, 10, (16, 512, 512), device=dev, generator=g)
    e = r.lookup(lab)
    r.eval()
    with torch.no_grad():
        r(torch.randn(16, 16, 512, 512, device=dev, generator=g))
    # EmbeddingLoss cross term at config 2 (vq_el_accum_vec_kernel, vq_el_bwd_vec_kernel)
    labels = torch.randint(0, 513, (16, 32, 32), device=dev, generator=g).repeat_interleave(8, 1).repeat_interleave(8, 2).to(torch.int32)
    l = cross_loss(z, labels.contiguous(), m.get_codebook())
    torch.autograd.grad(l, z)
torch.cuda.synchronize()
print("done")
