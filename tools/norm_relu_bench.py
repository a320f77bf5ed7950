"""Timing of the fused InstanceNorm2d + ReLU in front of the quantiser (`vq_norm_relu_fwd/bwd`) against
the stock torch-CUDA pair on the same tensors.   python tools/norm_relu_bench.py [B C H]
Algorithmic bytes per element: forward 8 (read x, write z), backward 12 (read x, g_z, write g_x)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from medical_image_editing_b200.src.functions import instance_norm_relu

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
C = int(sys.argv[2]) if len(sys.argv) > 2 else 64
H = int(sys.argv[3]) if len(sys.argv) > 3 else 256
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)
NB = 3                                                   # rotate inputs larger than L2 between iterations
xs = [torch.randn(B, C, H, H, device=dev, generator=g).requires_grad_(True) for _ in range(NB)]
gz = torch.randn(B, C, H, H, device=dev, generator=g)
try:
    hbm = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    hbm = 6650.0


def stock(x):
    return F.relu(F.instance_norm(x), inplace=True)


def timed(fn, n=12):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {"shape": [B, C, H, H], "hbm_peak_gbs": hbm}
n_el = B * C * H * H
for name, f in (("fused", instance_norm_relu), ("torch", stock)):
    def fwd(i):
        with torch.no_grad():
            f(xs[i % NB])
    ms_f = timed(fwd)
    ys = [f(x) for x in xs]

    def bwd(i):
        torch.autograd.grad(ys[i % NB], xs[i % NB], gz, retain_graph=True)
    ms_b = timed(bwd)
    out[name] = {"fwd_ms": ms_f, "bwd_ms": ms_b, "fwd_GBps": n_el * 8 / ms_f / 1e6, "bwd_GBps": n_el * 12 / ms_b / 1e6,
                 "fwd_frac_of_hbm": n_el * 8 / ms_f / 1e6 / hbm, "bwd_frac_of_hbm": n_el * 12 / ms_b / 1e6 / hbm}
    del ys
print(json.dumps(out))
