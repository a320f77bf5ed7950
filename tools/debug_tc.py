"""GPU debug helper for the tensor-core search kernel: dumps the raw TMEM accumulators and compares them
with z.e - |e|^2/2 computed in fp64; then runs the full forward on both paths and compares ids.
    python tools/debug_tc.py [B D H K]
"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import medical_image_editing_b200 as pkg

B, D, H, K = (int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (1, 64, 16, 512)
dev = "cuda:0"
L = pkg.lib()
g = torch.Generator(device=dev).manual_seed(5)
z = torch.randn(B, D, H, H, device=dev, generator=g)
E = torch.randn(K, D, device=dev, generator=g)
N = B * H * H
print("path", L.vq_assign_path(B, D, H, H, K, 0), "ncols", L.vq_debug_tc_ncols(D, K))
ncols = L.vq_debug_tc_ncols(D, K)
if ncols:
    ws = torch.empty(L.vq_workspace_bytes(N, K, D), dtype=torch.uint8, device=dev)
    out_all = torch.full((N * ncols + N * 8,), float("nan"), device=dev)
    out = out_all[:N * ncols].view(N, ncols)
    dbg2 = out_all[N * ncols:].view(N, 8)
    rc = L.vq_debug_tc_scores(z.data_ptr(), B, D, H, H, E.data_ptr(), K, out.data_ptr(), ws.data_ptr(), ws.numel(),
                              torch.cuda.current_stream().cuda_stream)
    print("rc", rc, L.vq_last_error())
    torch.cuda.synchronize()
    flat = z.permute(0, 2, 3, 1).reshape(N, D)
    print("dbg2 rows 0..2:", dbg2[:3].tolist())
    print("z2 ok:", float((dbg2[:, 0] - flat.pow(2).sum(1)).abs().max()), "z[.,0] ok:", float((dbg2[:, 1] - flat[:, 0]).abs().max()),
          "z[.,1] ok:", float((dbg2[:, 2] - flat[:, 1]).abs().max()))
    pp = torch.arange(N, device=dev) % 128
    print("E[p][0] ok:", float((dbg2[:, 3] - E[pp, 0]).abs().max()), "E[p][5] ok:", float((dbg2[:, 4] - E[pp, 5]).abs().max()))
    print("aug[p][0] vs -e2/2:", float((dbg2[:, 5] + 0.5 * E[pp].pow(2).sum(1)).abs().max()), "A_aug:", dbg2[:4, 6].tolist(),
          "tmem_base bits:", dbg2[0, 7].view(torch.int32).item())
    flat = z.permute(0, 2, 3, 1).reshape(N, D).double()
    ref = flat @ E.double().T - 0.5 * E.double().pow(2).sum(1)[None]
    # columns are in ascending-norm order of the codes: undo the permutation
    order = torch.argsort(E.pow(2).sum(1).sqrt(), stable=True)
    got = torch.empty_like(ref)
    got[:, order] = out[:, :K].double()
    err = (got - ref).abs()
    print("nan count", int(torch.isnan(out[:, :K]).sum()), "max abs err", float(err.nan_to_num(1e9).max()),
          "mean abs err", float(err.nan_to_num(0).mean()))
    print("pad cols min/max", float(out[:, K:].min()) if ncols > K else None)
    print("ref[0,:6]", ref[0, :6].tolist())
    print("got[0,:6]", got[0, :6].tolist())
    print("ref[5,250:258]", ref[5, 250:258].tolist())
    print("got[5,250:258]", got[5, 250:258].tolist())
    # error by row block / column block
    e2d = err.nan_to_num(1e9)
    for r0 in range(0, min(N, 256), 32):
        print("rows", r0, "colblocks", [round(float(e2d[r0:r0 + 32, c0:c0 + 64].max()), 3) for c0 in range(0, K, 64)][:8])
# full forward on both paths
outs = []
for flags in (1, 0):
    m = pkg.VQ(emb_dim=D, dict_size=K, momentum=0.99, eps=1e-5, knn_backend="torch").to(dev)
    m.kernel_flags = flags
    with torch.no_grad():
        m.embed.copy_(E); m.embed_avg.copy_(E.T); m.cluster_size.fill_(1.0)
    m.train(True)
    zz = z.clone().requires_grad_(True)
    q, loss, ids = m(zz)
    torch.cuda.synchronize()
    outs.append((ids, q.detach(), loss.item(), m.cluster_size.clone(), m.embed.clone()))
    if flags == 0:
        from medical_image_editing_b200.src.functions import vq_function as vf
        wsb = list(vf._WORKSPACES.values())[0]
        print("fallback rows:", L.vq_debug_fallback_rows(wsb.data_ptr(), N, K, D, torch.cuda.current_stream().cuda_stream), "of", N)
print("ids equal", torch.equal(outs[0][0], outs[1][0]), "mismatch", int((outs[0][0] != outs[1][0]).sum()))
print("q equal", torch.equal(outs[0][1], outs[1][1]))
print("loss", outs[0][2], outs[1][2])
print("cluster_size equal", torch.equal(outs[0][3], outs[1][3]), "embed maxdiff", float((outs[0][4] - outs[1][4]).abs().max()))

# several training steps: fallback rows and time per step as the codebook evolves
from medical_image_editing_b200.src.functions import vq_function as vf
m = pkg.VQ(emb_dim=D, dict_size=K, momentum=0.99, eps=1e-5, knn_backend="torch").to(dev)
with torch.no_grad():
    m.embed.copy_(E); m.embed_avg.copy_(E.T)
m.train(True)
for step in range(6):
    zz = torch.randn(B, D, H, H, device=dev, generator=g)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    norms = m.embed.norm(dim=1)
    e0.record()
    with torch.no_grad():
        q, loss, ids = m(zz)
    e1.record()
    torch.cuda.synchronize()
    wsb = list(vf._WORKSPACES.values())[0]
    fbr = L.vq_debug_fallback_rows(wsb.data_ptr(), N, K, D, torch.cuda.current_stream().cuda_stream)
    print(f"step {step}: fwd {e0.elapsed_time(e1):.3f} ms, fallback rows {fbr}, code norms min/med/max "
          f"{norms.min().item():.3g}/{norms.median().item():.3g}/{norms.max().item():.3g}, loss {loss.item():.4f}")
