"""Event trace of the third-generation resident kernel (CTA 0, its first 64 tiles).  Build with
   tools/build_variant.sh r3trace -DVQ_R3_TRACE ; VQ_B200_LIB=build_variants/lib_r3trace.so python tools/r3_trace.py [D K B train data]"""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import medical_image_editing_b200 as pkg

D = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = int(sys.argv[2]) if len(sys.argv) > 2 else 512
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16
train = (sys.argv[4] != "0") if len(sys.argv) > 4 else True
data = sys.argv[5] if len(sys.argv) > 5 else "noise"
H = 256
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)
L = pkg.lib()
m = pkg.VQ(emb_dim=D, dict_size=K, momentum=0.99, eps=1e-5, knn_backend="torch").to(dev)
with torch.no_grad():
    m.cluster_size.fill_(2048.0)
    m.embed_avg.copy_(m.embed.T * 2048.0)
m.train(train)
if data == "clustered":
    z = [(m.embed.detach()[torch.randint(0, K, (B, H, H), device=dev, generator=g)].permute(0, 3, 1, 2)
          + 0.1 * torch.randn(B, D, H, H, device=dev, generator=g)).contiguous() for _ in range(3)]
else:
    z = [torch.randn(B, D, H, H, device=dev, generator=g) for _ in range(3)]
with torch.no_grad():
    for i in range(3):
        m(z[i])
    torch.cuda.synchronize()
    m(z[0])
    torch.cuda.synchronize()
buf = np.zeros(32 * 64, dtype=np.int64)
n = L.vq_debug_tc_timing(buf.ctypes.data_as(ctypes.c_void_p), buf.size)
if n <= 0:
    raise SystemExit("this build has no trace (compile with -DVQ_R3_TRACE)")
NAMES = {0: "P load issue", 1: "M z_full", 2: "M tmemfree b0", 3: "M tmemfree b1", 4: "Z |z|^2 done",
         5: "S b0 wait", 6: "S b0 tmemfull", 7: "S b0 cols read", 8: "S b1 wait", 9: "S b1 tmemfull", 10: "S b1 cols read",
         11: "O merge start", 12: "O pub seen", 13: "O merge done", 14: "S published", 15: "O tile start", 16: "O z_full",
         17: "O z copied", 18: "O win seen", 19: "O outputs done", 20: "G batch start", 21: "G rerank done", 22: "G z loaded"}
ev = buf.reshape(32, 64).astype(np.float64)
ntiles = int((ev[0] > 0).sum())
print(f"tiles traced {ntiles}  (D={D} K={K} train={train} {data})")
lo, hi = 12, min(ntiles, 44)
print("steady state (tiles %d..%d): period %.0f cycles" % (lo, hi - 1, (ev[0, hi - 1] - ev[0, lo]) / (hi - 1 - lo)))
for e, nm in NAMES.items():
    ok = ev[e, lo:hi] > 0
    if not ok.any():
        continue
    rel = (ev[e, lo:hi] - ev[0, lo:hi])[ok]
    gap = np.diff(ev[e, lo:hi][ok]) if ok.sum() > 1 else np.zeros(1)
    print(f"  {e:2d} {nm:16s} rel {np.median(rel):8.0f} [{rel.min():8.0f} {rel.max():8.0f}]   period {np.median(gap):7.0f}  n {int(ok.sum())}")
print("straggler batches: iterations", ev[23, lo:hi].astype(int).tolist(), "entries", ev[24, lo:hi].astype(int).tolist())
print("tiles 20..23 (cycles since tile 20's load issue):")
for it in range(20, min(ntiles, 24)):
    print(f"  tile {it}: " + " ".join(f"{(ev[e, it] - ev[0, 20]):7.0f}" if ev[e, it] > 0 else "      -" for e in NAMES))
