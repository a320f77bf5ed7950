"""Self-consistency check of the quantiser forward at full size (no oracle): q == embed[ids] bit for bit, loss == mean((z - q)^2),
counts == bincount(ids); repeated to catch races.   python tools/r3_check.py [kind] [reps] [train]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import medical_image_editing_b200 as pkg

kind = sys.argv[1] if len(sys.argv) > 1 else "relu"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
train = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
D, K, B, H = 64, 512, 16, 256
dev = "cuda:0"
torch.manual_seed(11)
g = torch.Generator(device=dev).manual_seed(7)
m = pkg.VQ(emb_dim=D, dict_size=K, momentum=0.99, eps=1e-5, knn_backend="torch").to(dev)
with torch.no_grad():
    m.cluster_size.fill_(2048.0)
    m.embed_avg.copy_(m.embed.T * 2048.0)
m.train(train)
bad_total = 0
for r in range(reps):
    if kind == "relu":
        z = torch.relu(torch.randn(B, D, H, H, device=dev, generator=g))
    elif kind == "clustered":
        z = (m.embed.detach()[torch.randint(0, K, (B, H, H), device=dev, generator=g)].permute(0, 3, 1, 2)
             + 0.1 * torch.randn(B, D, H, H, device=dev, generator=g)).contiguous()
    else:
        z = torch.randn(B, D, H, H, device=dev, generator=g)
    e0 = m.embed.detach().clone()
    cs0 = m.cluster_size.detach().clone()
    with torch.no_grad():
        q, loss, ids = m(z)
    torch.cuda.synchronize()
    ids_nat = ids.transpose(1, 2)                         # reference layout is (b, w, h)
    q_chk = e0[ids_nat].permute(0, 3, 1, 2)
    nbad = int((q != q_chk).any(dim=1).sum().item())
    loss_chk = ((z.double() - q_chk.double()) ** 2).mean().item()
    counts = torch.bincount(ids.reshape(-1), minlength=K).float()
    msg = f"{kind} train={train} rep {r}: pixels with q != embed[ids]: {nbad}   loss {loss.item():.8f} vs {loss_chk:.8f} rel {abs(loss.item() - loss_chk) / loss_chk:.2e}"
    if train:
        cs_chk = 0.99 * cs0 + 0.01 * counts
        msg += f"   cluster_size max rel err {((m.cluster_size - cs_chk).abs().max() / cs_chk.abs().max()).item():.2e}"
    from medical_image_editing_b200.src.functions import vq_function as _vf
    wsb = _vf._WORKSPACES.get((0, torch.cuda.current_stream().cuda_stream))
    if wsb is not None:
        nfb = int(pkg.lib().vq_debug_fallback_rows(wsb.data_ptr(), B * H * H, K, D, torch.cuda.current_stream().cuda_stream))
        excess = (loss.item() - loss_chk) * z.numel()
        msg += f"   misc {wsb.view(torch.int32)[33856:33864].tolist()}"
        msg += f"   fallback rows {nfb}  excess loss sum {excess:.1f} ({excess / max(nfb, 1):.2f} per fb row)  max|e|^2 {e0.pow(2).sum(1).max().item():.3g}"
    print(msg)
    bad_total += nbad
print("TOTAL bad pixels", bad_total)
