"""One warm forward + backward of the fused InstanceNorm2d + ReLU at the config-2 plane shape (for ncu --set full captures)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from medical_image_editing_b200.src.functions import instance_norm_relu

H = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)
xs = [torch.randn(16, 64, H, H, device=dev, generator=g).requires_grad_(True) for _ in range(3)]
gz = torch.randn(16, 64, H, H, device=dev, generator=g)
for i in range(3):            # the last iteration is the one profiled (-s 4 -c 2: skip two forward/backward pairs)
    y = instance_norm_relu(xs[i])
    torch.autograd.grad(y, xs[i], gz)
torch.cuda.synchronize()
