"""Does tcgen05 kind::tf32 truncate (round toward zero) the fp32 operands it reads from shared memory?
Dumps the raw TMEM accumulators of the search kernel and compares them, in fp64, with three operand models.
    python tools/trunc_check.py [B D H K]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import medical_image_editing_b200 as pkg

B, D, H, K = (int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (2, 64, 32, 512)
dev = "cuda:0"
L = pkg.lib()
g = torch.Generator(device=dev).manual_seed(11)
z = torch.randn(B, D, H, H, device=dev, generator=g) * 3.0
E = torch.randn(K, D, device=dev, generator=g)
N = B * H * H
ncols = L.vq_debug_tc_ncols(D, K)
ws = torch.empty(L.vq_workspace_bytes(N, K, D), dtype=torch.uint8, device=dev)
out_all = torch.full((N * ncols + N * 8,), float("nan"), device=dev)
out = out_all[:N * ncols].view(N, ncols)
rc = L.vq_debug_tc_scores(z.data_ptr(), B, D, H, H, E.data_ptr(), K, out.data_ptr(), ws.data_ptr(), ws.numel(),
                          torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
assert rc == 0, L.vq_last_error()
flat = z.permute(0, 2, 3, 1).reshape(N, D)
order = torch.argsort(E.pow(2).sum(1).sqrt(), stable=True)
got = torch.empty(N, K, dtype=torch.float64, device=dev)
got[:, order] = out[:, :K].double()


def trunc(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


def rn(x):      # round to nearest even on 13 dropped bits
    i = x.view(torch.int32)
    r = i + 0xFFF + ((i >> 13) & 1)
    return (r & ~0x1FFF).view(torch.float32)


e2h = 0.5 * E.double().pow(2).sum(1)[None]
S = flat.double().abs() @ E.double().abs().T                       # sum_d |z_d e_d|
models = {"exact": (flat, E), "trunc": (trunc(flat), trunc(E)), "rn": (rn(flat), rn(E)),
          "trunc z only": (trunc(flat), E), "trunc e only": (flat, trunc(E))}
for name, (a, b) in models.items():
    ref = a.double() @ b.double().T - e2h
    err = got - ref
    print(f"{name:14s} max |got - model| = {float(err.abs().max()):.3e}   max |.|/S = {float((err.abs() / S).max()):.3e}")
ref = flat.double() @ E.double().T - e2h
rel = (got - ref) / S
print(f"(got - exact)/S: min {float(rel.min()):.3e} (2^-9 = {2**-9:.3e})  max {float(rel.max()):.3e}")
cen = (got + (got + e2h) * 2.0 ** -10 - ref) / S
print(f"centred (dot * (1 + 2^-10)): min {float(cen.min()):.3e}  max {float(cen.max()):.3e}   (2^-10 = {2**-10:.3e})")
nrm = flat.double().norm(dim=1)[:, None] * E.double().norm(dim=1)[None]
print(f"|got - exact| / (|z||e|): max {float(((got - ref).abs() / nrm).max()):.3e};  centred: {float(((got + (got + e2h) * 2.0 ** -10 - ref).abs() / nrm).max()):.3e}")
