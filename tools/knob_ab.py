"""A/B of the launch-shape knobs of the memory-bound kernels (read from the environment at every call, vq_kernels.cu
`tuning_knob`): channel split of the backward kernels / the EmbeddingLoss accumulation, register cap of the transposed
lookup.  CUDA events over 20 launches on rotating > L2 inputs; algorithmic bytes / time against the measured HBM peak.
    python tools/knob_ab.py            # prints one JSON line per (kernel, knob value)
"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import medical_image_editing_b200 as pkg
from medical_image_editing_b200._native import check

L = pkg.lib()
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
g = torch.Generator(device=dev).manual_seed(11)
S = torch.cuda.current_stream().cuda_stream
try:
    HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    HBM = float(HBM["hbm_gbs"])
except Exception:
    HBM = 6546.2


def timed(fn, n=20, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def report(kernel, knob, val, ms, nbytes, extra=None):
    d = {"kernel": kernel, "knob": knob, "value": val, "ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1),
         "GBps": round(nbytes / ms * 1e-6, 0), "frac_of_hbm": round(nbytes / ms * 1e-6 / HBM, 3)}
    if extra:
        d.update(extra)
    print(json.dumps(d), flush=True)


def p(t):
    return None if t is None else t.data_ptr()


# ---- config 2 shapes ------------------------------------------------------------------------------------------------
B, D, H, K = 16, 64, 256, 512
N = B * H * H
zs = [torch.randn(B, D, H, H, device=dev, generator=g) for _ in range(3)]
gq = [torch.randn(B, D, H, H, device=dev, generator=g) for _ in range(2)]
gz = torch.empty(B, D, H, H, device=dev)
E = torch.randn(K, D, device=dev, generator=g)
ids_nat = torch.randint(0, K, (N,), device=dev, generator=g, dtype=torch.int32)
gl = torch.ones((), device=dev)

ref = None
for v, mb in ((1, 0), (2, 0), (4, 0), (1, 4), (2, 4), (4, 4)):
    os.environ["VQ_BWD_DSPLIT"] = str(v)
    os.environ["VQ_BWD_MINB"] = str(mb)
    ms = timed(lambda i: check(L.vq_bwd(p(gq[i % 2]), p(gl), p(zs[i % 3]), p(ids_nat), p(E), p(gz), B, D, H, H, K, S), "vq_bwd"))
    check(L.vq_bwd(p(gq[0]), p(gl), p(zs[0]), p(ids_nat), p(E), p(gz), B, D, H, H, K, S), "vq_bwd")
    out = gz.clone()
    same = True if ref is None else bool(torch.equal(out, ref))
    ref = out if ref is None else ref
    report("vq_bwd_vec", "VQ_BWD_DSPLIT,VQ_BWD_MINB", [v, mb], ms, N * (12 * D + 8), {"bit_identical_to_first": same})
os.environ.pop("VQ_BWD_MINB", None)
os.environ.pop("VQ_BWD_DSPLIT", None)
del gq, ref, out

# ---- EmbeddingLoss at config 2 (labels piecewise constant, 8 x 8 blocks, 0 = no class) -----------------------------
small = torch.randint(0, K + 1, (B, H // 8, H // 8), device=dev, generator=g)
lab = small.repeat_interleave(8, 1).repeat_interleave(8, 2).to(torch.int32).contiguous()
loss = torch.empty((), device=dev)
w = torch.empty(B * K, device=dev)
wb = L.vq_embed_loss_work_bytes(B, K)
work = torch.empty(max(wb, 256), dtype=torch.uint8, device=dev)
ref_l = None
for v, mb in ((1, 0), (2, 0), (1, 5), (1, 6)):
    os.environ["VQ_EL_ACC_DSPLIT"] = str(v)
    os.environ["VQ_EL_ACC_MINB"] = str(mb)
    ms = timed(lambda i: check(L.vq_embed_loss_fwd(p(zs[i % 3]), p(lab), p(E), B, D, H, H, K, p(loss), p(w), p(work), work.numel(), S), "el_fwd"))
    check(L.vq_embed_loss_fwd(p(zs[0]), p(lab), p(E), B, D, H, H, K, p(loss), p(w), p(work), work.numel(), S), "el_fwd")
    lv = float(loss.item())
    ref_l = lv if ref_l is None else ref_l
    report("vq_embed_loss_fwd (memset + accum + finish)", "VQ_EL_ACC_DSPLIT,VQ_EL_ACC_MINB", [v, mb], ms, N * (4 * D + 4), {"loss": lv, "rel_diff_to_first": abs(lv - ref_l) / abs(ref_l)})
os.environ.pop("VQ_EL_ACC_MINB", None)
os.environ.pop("VQ_EL_ACC_DSPLIT", None)
for v in (1, 2, 4):
    os.environ["VQ_EL_BWD_DSPLIT"] = str(v)
    ms = timed(lambda i: check(L.vq_embed_loss_bwd(p(gl), p(zs[i % 3]), p(lab), p(E), p(w), p(gz), B, D, H, H, K, S), "el_bwd"))
    report("vq_el_bwd_vec", "VQ_EL_BWD_DSPLIT", v, ms, N * (8 * D + 4))
os.environ.pop("VQ_EL_BWD_DSPLIT", None)
del zs, gz, lab, small
torch.cuda.empty_cache()

# ---- run_recon shape: transposed lookup ------------------------------------------------------------------------------
B2, D2, H2, K2 = 16, 16, 512, 10
E2 = torch.randn(K2, D2, device=dev, generator=g)
idl = [torch.randint(0, K2, (B2, H2, H2), device=dev, generator=g) for _ in range(3)]
outs = [torch.empty(B2, D2, H2, H2, device=dev) for _ in range(2)]
status = torch.zeros(1, dtype=torch.int32, device=dev)
ref = None
for v in (0, 4, 5):
    os.environ["VQ_LOOKUP_MINB"] = str(v)
    ms = timed(lambda i: check(L.vq_lookup(p(idl[i % 3]), B2 * H2 * H2, p(E2), K2, D2, p(outs[i % 2]), 1, B2, H2, H2, p(status), S), "vq_lookup"))
    check(L.vq_lookup(p(idl[0]), B2 * H2 * H2, p(E2), K2, D2, p(outs[0]), 1, B2, H2, H2, p(status), S), "vq_lookup")
    o = outs[0].clone()
    same = True if ref is None else bool(torch.equal(o, ref))
    ref = o if ref is None else ref
    report("vq_lookup_nchw_tw<64,16>", "VQ_LOOKUP_MINB", v, ms, B2 * H2 * H2 * (8 + 4 * D2), {"bit_identical_to_first": same})
os.environ.pop("VQ_LOOKUP_MINB", None)
print("done")
