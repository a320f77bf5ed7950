"""Warm per-kernel durations and launch gaps of the bench step (config 2) from the CUPTI trace of torch.profiler
(no ncu: caches stay warm, launches stay asynchronous).   python tools/step_profile.py [steps]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import medical_image_editing_b200 as pkg

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = "cuda:0"
B, D, H, K = 16, 64, 256, 512
g = torch.Generator(device=dev).manual_seed(1234)
zs = [torch.randn(B, D, H, H, device=dev, generator=g).requires_grad_(True) for _ in range(4)]
g_q = torch.randn(B, D, H, H, device=dev, generator=g)
one = torch.ones((), device=dev)
vq = pkg.VQ(emb_dim=D, dict_size=K, momentum=0.99, eps=1e-5, knn_backend="torch").to(dev)
with torch.no_grad():
    cs = torch.rand(K, generator=torch.Generator().manual_seed(1234)) * (B * H * H / K) + 1.0
    vq.cluster_size.copy_(cs.to(dev))
    vq.embed_avg.copy_((vq.embed * vq.cluster_size[:, None]).T)
vq.train(True)


def step(i):
    z = zs[i % 4]
    q, loss, ids = vq(z)
    torch.autograd.grad((q, loss), z, (g_q, one))


for i in range(5):
    step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(steps):
        step(i)
    torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
print(f"{'kernel':60s} {'start_us':>10s} {'dur_us':>8s} {'gap_us':>8s}")
t0 = ev[0].time_range.start
last_end = None
per = {}
for e in ev[-int(len(ev) / steps) * 2:]:                     # the last two steps
    gap = (e.time_range.start - last_end) if last_end is not None else 0.0
    print(f"{e.name[:60]:60s} {e.time_range.start - t0:10.1f} {e.time_range.elapsed_us():8.1f} {gap:8.1f}")
    last_end = e.time_range.end
span = (ev[-1].time_range.end - ev[0].time_range.start) / steps
busy = sum(e.time_range.elapsed_us() for e in ev) / steps
print(f"per step: span {span:.1f} us, kernels {busy:.1f} us, gaps {span - busy:.1f} us")
