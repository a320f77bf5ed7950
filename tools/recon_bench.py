"""BASELINE config 4 (run_recon.py:115-192): code map -> embedding map (`UNetEncoder.get_embed_from_ids` = VQ.lookup + the
transposes) -> styled U-Net decoder, on one B200.  The networks are the UNMODIFIED reference classes (baseline/_ref mirror
or /root/reference, through oracle/ref_loader.py -- measurement infrastructure, not the product path); `networks.vq.VQ` is
rebound to this package for the `b200` rows and left alone for the `reference` rows, same weights, same GPU.

    python tools/recon_bench.py [B]        # one JSON line: lookup-only, inference (ids -> recon) and decoder train step
"""
import contextlib
import copy
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import medical_image_editing_b200 as pkg
from oracle import ref_loader

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S, K, D = 512, 10, 16
dev = "cuda:0"


@contextlib.contextmanager
def rebound(mod):
    old = mod.VQ
    mod.VQ = pkg.VQ
    try:
        yield
    finally:
        mod.VQ = old


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    if not ref_loader.reference_available():
        print(json.dumps({"unavailable": "no copy of the reference sources reachable"}))
        return
    enc_mod = ref_loader.load_reference_net("unet_encoder")
    dec_mod = ref_loader.load_reference_net("unet_decoder")
    enc_args = (1, [16, 32, 64, 128, 256], K, 0.999, "torch", False, 1, False)     # run_recon.py:116-125
    torch.manual_seed(0)
    enc_ref = enc_mod.UNetEncoder(*enc_args)
    with rebound(enc_mod):
        enc_new = enc_mod.UNetEncoder(*enc_args)
    enc_new.load_state_dict(copy.deepcopy(enc_ref.state_dict()), strict=True)
    dec = dec_mod.UNetDecoder(in_channels=16, out_channels=1, filters=[32, 64, 128, 256, 512], dropped_skip_layers=[],
                              use_styled_up_block=True, use_pixel_shuffle=False)           # run_recon.py:127-139
    enc_ref, enc_new, dec = enc_ref.to(dev).eval(), enc_new.to(dev).eval(), dec.to(dev)
    g = torch.Generator().manual_seed(1234)
    ids = torch.randint(0, K, (B, S, S), generator=g).to(dev)
    mask = (torch.rand(B, S, S, generator=g) > 0.1).to(dev).long()
    target = torch.randn(B, 1, S, S, generator=g).clamp(-1, 1).to(dev)
    out = {"workload": f"recon_k10d16 (BASELINE config 4): ids {B}x{S}x{S} -> embed {B}x{D}x{S}x{S} -> UNetDecoder "
                       "(reference classes, stock cuDNN)", "B": B}
    with torch.no_grad():
        assert torch.equal(enc_new.get_embed_from_ids(ids), enc_ref.get_embed_from_ids(ids))
    out["get_embed_from_ids_bit_exact"] = True

    def infer(enc):
        def f():
            with torch.no_grad():
                e = enc.get_embed_from_ids(ids)
                e = e * mask[:, None, :, :]
                e = e * mask.numel() / mask.sum()
                return dec(e)
        return f

    def lookup(enc):
        def f():
            with torch.no_grad():
                return enc.get_embed_from_ids(ids)
        return f

    dec.eval()
    for name, enc in (("b200", enc_new), ("reference", enc_ref)):
        ms_l = timed(lookup(enc), 20)
        ms_i = timed(infer(enc), 5, 2)
        out[name] = {"lookup_ms": ms_l, "lookup_ids_per_s": B * S * S / (ms_l * 1e-3),
                     "lookup_GBps_algorithmic": B * S * S * (8 + 4 * D) / (ms_l * 1e-3) / 1e9,
                     "inference_ms": ms_i, "inference_slices_per_s": B / (ms_i * 1e-3)}
    # decoder training step of the reconstruction model (encoder frozen: its code map is the input)
    dec.train(True)
    opt = torch.optim.Adam(dec.parameters(), lr=1e-4)
    mb = min(B, 4)

    def train(enc):
        def f():
            with torch.no_grad():
                e = enc.get_embed_from_ids(ids[:mb])
            opt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.mse_loss(dec(e), target[:mb])
            loss.backward()
            opt.step()
        return f

    for name, enc in (("b200", enc_new), ("reference", enc_ref)):
        ms_t = timed(train(enc), 5, 2)
        out[name]["decoder_train_ms"] = ms_t
        out[name]["decoder_train_slices_per_s"] = mb / (ms_t * 1e-3)
        out[name]["decoder_train_batch"] = mb
    print(json.dumps(out))


if __name__ == "__main__":
    main()
