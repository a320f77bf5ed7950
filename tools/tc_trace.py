"""Event trace of the resident tensor-core kernel (CTA 0, its first 64 tiles).  Build with
   tools/build_variant.sh trace -DVQ_TC_TRACE ; VQ_B200_LIB=build_variants/lib_trace.so python tools/tc_trace.py [D K B train data]
Prints, per tile, every event relative to the tile's z-load issue, and the steady-state medians."""
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
buf = np.zeros(148 * 16 * 8, dtype=np.int64)
n = L.vq_debug_tc_timing(buf.ctypes.data_as(ctypes.c_void_p), buf.size)
if n <= 0:
    raise SystemExit("this build has no trace (compile with -DVQ_TC_TRACE)")
NAMES = ["P load issue", "M z_full", "M tmemfree b0", "M commit b0", "M tmemfree b1", "M commit b1", "Z z_full", "Z |z|^2 done",
         "S tile start", "S tmemfull b0", "S done b0", "S tmemfull b1", "S done b1", "S published",
         "O tile start", "O z_full", "O zcopy+|z|^2", "O scan seen", "O merge done", "O rerank done", "O outputs done"]
ev = buf[:32 * 64].reshape(32, 64).astype(np.float64)
ntiles = int((ev[0] > 0).sum())
print(f"tiles traced {ntiles}  (D={D} K={K} train={train} {data})")
t0 = ev[0, 0]
lo, hi = 12, min(ntiles, 44)
print("steady state (tiles %d..%d): period %.0f cycles" % (lo, hi - 1, (ev[0, hi - 1] - ev[0, lo]) / (hi - 1 - lo)))
print("event times relative to the tile's own load issue (median / min / max), and absolute gaps between consecutive tiles")
for e, nm in enumerate(NAMES):
    rel = ev[e, lo:hi] - ev[0, lo:hi]
    gap = np.diff(ev[e, lo:hi])
    print(f"  {e:2d} {nm:16s} rel {np.median(rel):8.0f} [{rel.min():8.0f} {rel.max():8.0f}]   period {np.median(gap):7.0f}")
print("first tiles, absolute cycles since the first load issue:")
for it in range(0, min(ntiles, 8)):
    print(f"  tile {it}: " + " ".join(f"{(ev[e, it] - t0):7.0f}" for e in range(len(NAMES))))
print("tiles 20..23:")
for it in range(20, min(ntiles, 24)):
    print(f"  tile {it}: " + " ".join(f"{(ev[e, it] - ev[0, 20]):7.0f}" for e in range(len(NAMES))))
