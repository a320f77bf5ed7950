import os, sys
sys.path.insert(0, "/root/repo")
import torch
import medical_image_editing_b200 as pkg
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(5)
r = pkg.VQ(emb_dim=16, dict_size=10, momentum=0.999, eps=1e-5, knn_backend="torch").to(dev)
labs = [torch.randint(0, 10, (16, 512, 512), device=dev, generator=g) for _ in range(4)]
for i in range(4): e = r.lookup(labs[i])
ref = torch.nn.functional.embedding(labs[0], r.embed)
assert torch.equal(r.lookup(labs[0]), ref)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20): r.lookup(labs[i % 4])
e1.record(); torch.cuda.synchronize()
print(os.environ.get("VQ_B200_LIB", "default"), "lookup 16x512x512 K=10 D=16:", e0.elapsed_time(e1) / 20, "ms")
