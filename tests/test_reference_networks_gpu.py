"""The UNMODIFIED reference networks with `networks.vq.VQ` rebound to this package, against the same networks
with the reference's own `VQModule`, on the same GPU and the same weights (SURVEY section 4
"integration").  The reference sources come from /root/reference/src or, on the GPU box, from the git-ignored mirror
baseline/_ref/src that `__graft_entry__.build()` makes (oracle/ref_loader.py); without either the tests skip.

Checked: strict `load_state_dict` of a reference checkpoint into the rebound network; code maps bit-exact (a differing
pixel only when the reference's own fp32 scores tie); reconstruction / commitment loss / gradients <= 1e-5 relative;
`generate_images_from_ids` / `get_embed_from_ids` (the run_recon.py path) bit-exact."""
import contextlib
import copy

import pytest
import torch

import medical_image_editing_b200 as pkg
from oracle import ref_loader
from oracle.vq_oracle import score_gap_is_tie
from util import rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.reference_available(),
                                                  reason="no copy of the reference sources reachable")]
DEV = "cuda:0"
TOL = 1e-5


@contextlib.contextmanager
def rebound_vq(mod):
    """`from .vq import VQ` made `VQ` a module-level name of the reference network file: point it at this package."""
    ref_vq = mod.VQ
    mod.VQ = pkg.VQ
    try:
        yield ref_vq
    finally:
        mod.VQ = ref_vq


def build_pair(mod, cls_name, *args, **kw):
    """(reference network with the reference VQModule, same network with the B200 VQ) with identical weights."""
    torch.manual_seed(0)
    ref_net = getattr(mod, cls_name)(*args, **kw)
    with rebound_vq(mod):
        torch.manual_seed(1)                              # different init on purpose: the checkpoint must carry everything
        new_net = getattr(mod, cls_name)(*args, **kw)
    assert type(new_net.vq).__module__.startswith("medical_image_editing_b200")
    assert not type(ref_net.vq).__module__.startswith("medical_image_editing_b200")
    missing = new_net.load_state_dict(copy.deepcopy(ref_net.state_dict()), strict=True)     # trainers/base.py:85-102
    assert not missing.missing_keys and not missing.unexpected_keys
    return ref_net.to(DEV), new_net.to(DEV)


def ids_equal_or_tied(ids_new, ids_ref, embed, z):
    a, b = ids_new.reshape(-1).cpu(), ids_ref.reshape(-1).cpu()
    bad = (a != b).nonzero().reshape(-1)
    if bad.numel():
        flat = z.detach().cpu().permute(0, 2, 3, 1).reshape(-1, z.shape[1])
        tie = score_gap_is_tie(embed.detach().cpu(), flat[bad], a[bad], b[bad])
        assert bool(tie.all()), f"{int((~tie).sum())} code-map mismatches that are not fp32 ties"
    print(f"[reference networks] {bad.numel()} tolerated tie pixels of {a.numel()}")
    return int(bad.numel())


@pytest.fixture(autouse=True)
def _deterministic():
    old = torch.backends.cudnn.deterministic, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.deterministic, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("training", [False, True])
def test_vqwnet_with_rebound_vq_matches_reference(training):
    mod = ref_loader.load_reference_net("vqwnet")                      # networks/vqwnet.py:13-152
    ref_net, new_net = build_pair(mod, "VQWNet", 1, 1, filters=[32, 32, 64, 64, 128], dict_size=512)
    ref_net.train(training)
    new_net.train(training)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 1, 64, 64, generator=g).clamp(-1, 1).to(DEV)
    xr, xn = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    embed0 = ref_net.vq.embed.detach().clone()                        # the codebook both searches see (pre-update)
    out_r, out_n = ref_net(xr), new_net(xn)
    assert out_n["ids"].dtype == out_r["ids"].dtype and out_n["ids"].shape == out_r["ids"].shape
    assert rel_err(out_n["embed"], out_r["embed"]) == 0.0             # identical convolutions feed both quantisers
    # the networks return transpose(ids, 1, 2) + 1 (vqwnet.py:110-111): natural (b, h, w) order, 1-based
    nties = ids_equal_or_tied(out_n["ids"] - 1, out_r["ids"] - 1, embed0, out_r["embed"])
    if nties == 0:
        assert rel_err(out_n["recon"], out_r["recon"]) <= TOL
        assert abs(out_n["commit_loss"].item() - out_r["commit_loss"].item()) <= TOL * abs(out_r["commit_loss"].item())
        loss_r = (out_r["recon"] - x).pow(2).mean() + out_r["commit_loss"]
        loss_n = (out_n["recon"] - x).pow(2).mean() + out_n["commit_loss"]
        (gr,) = torch.autograd.grad(loss_r, xr)
        (gn,) = torch.autograd.grad(loss_n, xn)
        assert rel_err(gn, gr) <= 1e-4                                 # through ~40 cuDNN layers
        if training:                                                   # EMA buffers after one step (vq_module.py:194-199)
            for name in ("embed", "cluster_size", "embed_avg"):
                assert rel_err(getattr(new_net.vq, name), getattr(ref_net.vq, name)) <= TOL, name
    else:
        assert nties <= 2
    # editing path: ids -> image (vqwnet.py:154-176)
    ids = out_r["ids"].clone()
    with torch.no_grad():
        new_net.vq.embed.copy_(ref_net.vq.embed)
    gr_, gn_ = ref_net.generate_images_from_ids(ids - 1), new_net.generate_images_from_ids(ids - 1)
    assert torch.equal(gn_["ids"], gr_["ids"])
    assert rel_err(gn_["recon"], gr_["recon"]) <= TOL


def test_unet_encoder_get_embed_from_ids_matches_reference():
    mod = ref_loader.load_reference_net("unet_encoder")                # networks/unet_encoder.py:15-123, run_recon.py:115-139
    ref_net, new_net = build_pair(mod, "UNetEncoder", 1, [16, 32, 64, 128, 256], 10, 0.999, "torch", False, 1, True)
    ref_net.eval()
    new_net.eval()
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(0, 10, (2, 128, 128), generator=g).to(DEV)
    er, en = ref_net.get_embed_from_ids(ids), new_net.get_embed_from_ids(ids)     # run_recon.py:189
    assert en.shape == er.shape == (2, 16, 128, 128)
    assert torch.equal(en, er)
    x = torch.randn(2, 1, 128, 128, generator=g).to(DEV)
    with torch.no_grad():
        qr, lr, ir = ref_net(x)
        qn, ln, inn = new_net(x)
    assert torch.equal(inn, ir), "code maps differ"
    assert torch.equal(qn, qr) and abs(ln.item() - lr.item()) <= TOL * abs(lr.item())


def test_vqgan_with_rebound_vq_matches_reference():
    mod = ref_loader.load_reference_net("vqgan")                       # networks/vqgan.py:380-446 (emb_dim 512, K 64 defaults)
    kw = dict(in_channels=1, mid_channels=32, out_channels=3, emb_dim=512, dict_size=64, enc_ch_multiplier=(1, 2),
              dec_ch_multiplier=(1, 2), num_res_blocks=1, enc_attn_resolutions=[], dec_attn_resolutions=[], resolution=64)
    ref_net, new_net = build_pair(mod, "VQGAN", **kw)
    ref_net.eval()
    new_net.eval()
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 1, 64, 64, generator=g).to(DEV)
    with torch.no_grad():
        rr, lr, ir, er = ref_net(x)
        rn, ln, inn, en = new_net(x)
    assert torch.equal(inn, ir), "code maps differ"
    assert torch.equal(en, er)
    assert rel_err(rn, rr) <= TOL and abs(ln.item() - lr.item()) <= TOL * abs(lr.item())
    with torch.no_grad():
        assert rel_err(new_net.generate_image_from_ids(ir), ref_net.generate_image_from_ids(ir)) <= TOL
