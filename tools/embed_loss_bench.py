"""EmbeddingLoss cross term on the GPU: time of forward and backward against the HBM roofline, with the CPU oracle
beside it on a bounded sample.   python tools/embed_loss_bench.py [B D H K]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import medical_image_editing_b200 as pkg
from medical_image_editing_b200.src.functions.embed_loss import cross_loss
from oracle import embed_loss_oracle as elo

B, D, H, K = (int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (16, 64, 256, 512)
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(3)
cb = torch.randn(D, K, device=dev, generator=g)
small = torch.randint(0, K + 1, (B, H // 8, H // 8), device=dev, generator=g)
lab = small.repeat_interleave(8, 1).repeat_interleave(8, 2).to(torch.int32).contiguous()
zs = [torch.randn(B, D, H, H, device=dev, generator=g).requires_grad_(True) for _ in range(4)]
for i in range(3):
    l = cross_loss(zs[i % 4], lab, cb)
    torch.autograd.grad(l, zs[i % 4])
torch.cuda.synchronize()
n = 10
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
tf = tb = 0.0
for i in range(n):
    ev[0].record()
    l = cross_loss(zs[i % 4], lab, cb)
    ev[1].record()
    torch.autograd.grad(l, zs[i % 4])
    ev[2].record()
    torch.cuda.synchronize()
    tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2])
tf /= n; tb /= n
N = B * H * H
bf, bb = N * (4 * D + 4), N * (8 * D + 4)
print(f"B={B} D={D} H={H} K={K}: forward {tf:.3f} ms ({bf / tf * 1e-6:.0f} GB/s of {bf / 1e6:.0f} MB algorithmic), "
      f"backward {tb:.3f} ms ({bb / tb * 1e-6:.0f} GB/s), {N / (tf + tb) * 1e-6:.1f} G locations/s fwd+bwd")
# CPU oracle on one slice
e1 = zs[0][:1].detach().cpu().requires_grad_(True)
t0 = time.perf_counter()
l = elo.cross_loss(e1, lab[:1].cpu(), cb.cpu())
torch.autograd.grad(l, e1)
dt = time.perf_counter() - t0
print(f"CPU oracle (restatement without the B*D*K*HW expansion, {torch.get_num_threads()} threads), 1 slice fwd+bwd: {dt * 1e3:.1f} ms "
      f"-> {H * H / dt * 1e-6:.2f} M locations/s")
