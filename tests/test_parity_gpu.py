"""GPU parity tests: the CUDA path (through the C-ABI, via the module / autograd Function) against
 (1) golden vectors produced by the unmodified reference (tests/golden),
 (2) the CPU oracle on seeded inputs at sizes it finishes in seconds,
 (3) size-independent properties at BASELINE.json's full sizes.
Bars (BASELINE.json north_star): ids and counts bit-exact; quantized bit-exact (pure gather);
loss / gradients / EMA buffers within 1e-5 relative in fp32."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import medical_image_editing_b200 as pkg
from medical_image_editing_b200 import _native
from oracle.vq_oracle import OracleVQ, seeded_case, make_oracle, score_gap_is_tie
from util import SMALL_CASES, load_golden, t, rel_err, set_state

pytestmark = pytest.mark.gpu

TOL = 1e-5
DEV = "cuda:0"
PATHS = [pytest.param(_native.VQ_FLAG_FORCE_SIMT, id="simt"), pytest.param(0, id="auto")]


def new_vq(K, D, momentum=0.99, flags=0, **kw):
    m = pkg.VQ(emb_dim=D, dict_size=K, momentum=momentum, eps=1e-5, knn_backend="torch", **kw).to(DEV)
    m.kernel_flags = flags
    return m


def assert_ids_match(ids_gpu, ids_ref, embed, z, allow_ties=True):
    """Bit-exact ids; a differing row is tolerated only if the oracle's own fp32 scores of the two
    codes are a numerical tie (the reference's tie-break is unspecified)."""
    a = ids_gpu.cpu().reshape(-1)
    b = ids_ref.reshape(-1)
    bad = (a != b).nonzero().reshape(-1)
    if bad.numel() == 0:
        return 0
    assert allow_ties, f"{bad.numel()} id mismatches"
    flat = z.detach().cpu().transpose(1, -1).reshape(-1, z.shape[1])
    tie = score_gap_is_tie(embed.cpu(), flat[bad], a[bad], b[bad])
    assert bool(tie.all()), f"{int((~tie).sum())} id mismatches that are not fp32 ties (of {bad.numel()} differing rows)"
    assert bad.numel() <= max(2, a.numel() // 100000), f"too many tie rows: {bad.numel()}"
    report_ties(int(bad.numel()), a.numel())
    return int(bad.numel())


def report_ties(n, total, where=""):
    """A differing row is tolerated only when the oracle's own fp32 scores of the two codes tie; say how many there
    were (expected: 0), so that a regression from 0 is visible (`pytest -s`, or gpurun_out/parity_ties.log)."""
    import os
    msg = f"[parity] {n} tolerated tie rows of {total} {where}".rstrip()
    print(msg)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if n and os.path.isdir(out):
        with open(os.path.join(out, "parity_ties.log"), "a") as f:
            f.write(msg + "\n")


# ---------------------------------------------------------------------------------------------
# (1) golden vectors from the reference
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("flags", PATHS)
@pytest.mark.parametrize("name", SMALL_CASES)
def test_golden_forward_backward_ema(name, flags):
    g = load_golden(name)
    B, D, H, K, training, steps = [int(x) for x in g["meta"]]
    m = new_vq(K, D, float(g["momentum"][0]), flags)
    set_state(m, g["embed0"], g["cluster_size0"], g["embed_avg0"])
    m.train(bool(training))
    for s in range(steps):
        sfx = "" if s == 0 else f"_s{s}"
        if s > 0:   # re-sync state from the reference each step (1-ulp drift may flip later near-ties)
            p = "" if s == 1 else f"_s{s - 1}"
            set_state(m, g["embed1" + p], g["cluster_size1" + p], g["embed_avg1" + p])
        z = t(g["z" + sfx], DEV).requires_grad_(True)
        q, loss, ids = m(z)
        assert q.shape == z.shape and ids.shape == (B, H, H) and ids.dtype == torch.int64 and loss.dim() == 0
        total = (q * t(g["g_q" + sfx], DEV)).sum() + float(g["w" + sfx][0]) * loss
        (g_z,) = torch.autograd.grad(total, z)
        assert torch.equal(ids.cpu(), t(g["ids" + sfx])), "ids must be bit-exact"
        assert torch.equal(q.detach().cpu(), t(g["q" + sfx])), "quantized must be bit-exact"
        assert abs(loss.item() - float(g["loss" + sfx][0])) <= TOL * abs(float(g["loss" + sfx][0]))
        assert rel_err(g_z, t(g["g_z" + sfx])) <= TOL
        assert rel_err(m.embed, t(g["embed1" + sfx])) <= TOL
        assert rel_err(m.cluster_size, t(g["cluster_size1" + sfx])) <= TOL
        assert rel_err(m.embed_avg, t(g["embed_avg1" + sfx])) <= TOL


@pytest.mark.parametrize("flags", PATHS)
def test_golden_config1_quantiser(flags):
    """BASELINE config 1 quantiser shape: 1x64x256x256, K=512, against the reference's own run."""
    g = load_golden("config1_k512_d64_256")
    B, D, H, K = [int(x) for x in g["meta"][:4]]
    z, embed = seeded_case(B, D, H, H, K, seed=int(g["seed"][0]), kind="gauss")
    ora = make_oracle(K, D, embed, warmed=True, n_for_warm=B * H * H)
    m = new_vq(K, D, 0.99, flags)
    set_state(m, ora.embed.numpy(), ora.cluster_size.numpy(), ora.embed_avg.numpy())
    m.train(True)
    q, loss, ids = m(z.to(DEV))
    assert torch.equal(ids.cpu(), t(g["ids_i16"]).long())
    counts = torch.bincount(ids.reshape(-1), minlength=K).cpu().numpy()
    assert np.array_equal(counts, g["counts"])
    assert abs(loss.item() - float(g["loss"][0])) <= TOL * float(g["loss"][0])
    assert rel_err(m.embed, t(g["embed1"])) <= TOL
    assert rel_err(m.embed_avg, t(g["embed_avg1"])) <= TOL
    assert rel_err(m.cluster_size, t(g["cluster_size1"])) <= TOL
    qs = q.double()
    assert abs(qs.sum().item() - g["q_checksum"][0]) <= 1e-9 * max(1.0, abs(g["q_checksum"][0])) + 1e-6
    assert abs(qs.pow(2).sum().item() - g["q_checksum"][1]) <= 1e-9 * g["q_checksum"][1]


# ---------------------------------------------------------------------------------------------
# (2) seeded inputs vs the CPU oracle
# ---------------------------------------------------------------------------------------------
SWEEP = [(K, D, kind) for K in (64, 512, 4096) for D in (64, 256) for kind in ("gauss",)] + [
    (512, 64, "clustered"), (512, 64, "relu"), (512, 256, "relu"), (10, 16, "gauss"), (100, 24, "gauss"),
    (64, 512, "gauss"), (300, 40, "clustered"), (512, 128, "gauss"), (1000, 96, "relu")]


@pytest.mark.parametrize("flags", PATHS)
@pytest.mark.parametrize("K,D,kind", SWEEP)
def test_seeded_vs_oracle(K, D, kind, flags):
    B, H = 2, 64                                   # N = 8192 vectors
    z, embed = seeded_case(B, D, H, H, K, seed=1234 + K + D, kind=kind)
    ora = make_oracle(K, D, embed, warmed=True, n_for_warm=B * H * H, chunk=8192)
    m = new_vq(K, D, 0.99, flags)
    set_state(m, ora.embed.numpy(), ora.cluster_size.numpy(), ora.embed_avg.numpy())
    ora.train(True)
    m.train(True)
    g = torch.Generator().manual_seed(7)
    g_q = torch.randn(B, D, H, H, generator=g)

    z_ref = z.clone().requires_grad_(True)
    q_ref, loss_ref, ids_ref = ora(z_ref)
    (gz_ref,) = torch.autograd.grad((q_ref * g_q).sum() + 0.7 * loss_ref, z_ref)

    z_gpu = z.to(DEV).requires_grad_(True)
    q, loss, ids = m(z_gpu)
    (gz,) = torch.autograd.grad((q * g_q.to(DEV)).sum() + 0.7 * loss, z_gpu)

    nties = assert_ids_match(ids, ids_ref, embed, z)
    if nties == 0:
        assert torch.equal(q.detach().cpu(), q_ref.detach().contiguous())
        assert rel_err(m.cluster_size, ora.cluster_size) <= TOL
        assert rel_err(m.embed_avg, ora.embed_avg) <= TOL
        assert rel_err(m.embed, ora.embed) <= TOL
        assert rel_err(gz, gz_ref) <= TOL
    assert abs(loss.item() - loss_ref.item()) <= TOL * abs(loss_ref.item())


# tile counts that are not a multiple of the grid, single-tile and two-tile launches, odd quad counts: the streamed-codebook
# kernel splits the tiles of a CTA between two teams of epilogue warps (a team may get no tile at all)
MULTI_TILE = [(5, 96, 512, 256, "gauss"), (3, 128, 600, 132, "clustered"), (1, 16, 512, 256, "gauss"),
              (7, 64, 4096, 68, "gauss"), (9, 48, 512, 128, "relu"), (4, 80, 512, 64, "gauss"), (11, 32, 64, 64, "clustered"),
              # single-block codebooks in the streamed kernel (the ring may not run a whole tile ahead of the tensor core)
              (6, 64, 64, 256, "gauss"), (5, 64, 256, 256, "clustered"), (13, 32, 64, 192, "gauss"), (3, 96, 40, 224, "relu")]


@pytest.mark.parametrize("B,H,K,D,kind", MULTI_TILE)
def test_multi_tile_vs_oracle(B, H, K, D, kind):
    z, embed = seeded_case(B, D, H, H, K, seed=4242 + K + D + B, kind=kind)
    ora = make_oracle(K, D, embed, warmed=True, n_for_warm=B * H * H, chunk=8192)
    m = new_vq(K, D, 0.99, 0)
    assert pkg.lib().vq_assign_path(B, D, H, H, K, 0) == 1, "this shape should take the tensor-core path"
    set_state(m, ora.embed.numpy(), ora.cluster_size.numpy(), ora.embed_avg.numpy())
    ora.train(True)
    m.train(True)
    z_ref = z.clone().requires_grad_(True)
    q_ref, loss_ref, ids_ref = ora(z_ref)
    (gz_ref,) = torch.autograd.grad(q_ref.sum() + 0.3 * loss_ref, z_ref)
    z_gpu = z.to(DEV).requires_grad_(True)
    q, loss, ids = m(z_gpu)
    (gz,) = torch.autograd.grad(q.sum() + 0.3 * loss, z_gpu)
    nties = assert_ids_match(ids, ids_ref, embed, z)
    if nties == 0:
        assert torch.equal(q.detach().cpu(), q_ref.detach().contiguous())
        assert rel_err(m.cluster_size, ora.cluster_size) <= TOL
        assert rel_err(m.embed_avg, ora.embed_avg) <= TOL
        assert rel_err(m.embed, ora.embed) <= TOL
        assert rel_err(gz, gz_ref) <= TOL
    assert abs(loss.item() - loss_ref.item()) <= TOL * abs(loss_ref.item())


# inference calls on small codebooks (K <= 16, K*D <= 256: run_recon's K = 10, D = 16) take the thread-per-pixel kernel
SMALL_CODEBOOK = [(2, 64, 10, 16, "gauss"), (3, 40, 16, 16, "clustered"), (1, 50, 7, 12, "relu"), (5, 16, 3, 5, "gauss"),
                  (2, 128, 10, 16, "clustered"), (1, 24, 1, 8, "gauss")]


@pytest.mark.parametrize("B,H,K,D,kind", SMALL_CODEBOOK)
def test_small_codebook_inference_vs_oracle_and_simt(B, H, K, D, kind):
    z, embed = seeded_case(B, D, H, H, K, seed=977 + K + D + B, kind=kind)
    assert pkg.lib().vq_assign_path(B, D, H, H, K, _native.VQ_FLAG_NO_STATS) == 2
    ora = make_oracle(K, D, embed, warmed=True, n_for_warm=B * H * H, chunk=8192)
    ora.eval()
    q_ref, loss_ref, ids_ref = ora(z)
    outs = []
    for flags in (0, _native.VQ_FLAG_FORCE_SIMT):
        m = new_vq(K, D, 0.99, flags)
        set_state(m, ora.embed.numpy(), ora.cluster_size.numpy(), ora.embed_avg.numpy())
        m.eval()
        zg = z.to(DEV).requires_grad_(True)
        q, loss, ids = m(zg)
        (gz,) = torch.autograd.grad(q.sum() + 0.5 * loss, zg)
        outs.append((ids, q.detach(), loss.item(), gz))
        assert torch.equal(m.embed.cpu(), ora.embed), "eval must not touch the buffers"
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]), "small-codebook kernel != CUDA-core search"
    assert torch.equal(outs[0][3], outs[1][3])
    nties = assert_ids_match(outs[0][0], ids_ref, embed, z)
    if nties == 0:
        assert torch.equal(outs[0][1].cpu(), q_ref.detach().contiguous())
    assert abs(outs[0][2] - loss_ref.item()) <= TOL * abs(loss_ref.item())


@pytest.mark.parametrize("flags", PATHS)
def test_cold_start_exploded_codes(flags):
    """First EMA step with cluster_size == 0 blows unused codes up to ~1e5 x;
    the next forward must still pick the reference's codes."""
    B, D, H, K = 1, 64, 32, 512                   # N = 1024 < K*... many dead codes
    z, embed = seeded_case(B, D, H, H, K, seed=31)
    ora = make_oracle(K, D, embed)
    m = new_vq(K, D, 0.99, flags)
    set_state(m, ora.embed.numpy(), ora.cluster_size.numpy(), ora.embed_avg.numpy())
    ora.train(True)
    m.train(True)
    for step in range(3):
        zz, _ = seeded_case(B, D, H, H, K, seed=100 + step)
        _, loss_ref, ids_ref = ora(zz)
        _, loss, ids = m(zz.to(DEV))
        assert torch.equal(ids.cpu(), ids_ref)
        assert abs(loss.item() - loss_ref.item()) <= TOL * abs(loss_ref.item())
        assert rel_err(m.embed, ora.embed) <= TOL
        assert ora.embed.abs().max() > 1e4         # the explosion really happened
        set_state(m, ora.embed.numpy(), ora.cluster_size.numpy(), ora.embed_avg.numpy())


@pytest.mark.parametrize("flags", PATHS)
@pytest.mark.parametrize("B,D,H,K", [(1, 3, 5, 7), (2, 1, 1, 1), (1, 17, 9, 33), (3, 8, 2, 2), (1, 260, 6, 5)])
def test_ragged_shapes(B, D, H, K, flags):
    z, embed = seeded_case(B, D, H, H, K, seed=B * 1000 + D)
    ora = make_oracle(K, D, embed, momentum=0.9, warmed=True, n_for_warm=B * H * H)
    m = new_vq(K, D, 0.9, flags)
    set_state(m, ora.embed.numpy(), ora.cluster_size.numpy(), ora.embed_avg.numpy())
    ora.train(True)
    m.train(True)
    z_ref = z.clone().requires_grad_(True)
    q_ref, loss_ref, ids_ref = ora(z_ref)
    (gz_ref,) = torch.autograd.grad(q_ref.sum() * 0.5 + loss_ref, z_ref)
    z_gpu = z.to(DEV).requires_grad_(True)
    q, loss, ids = m(z_gpu)
    (gz,) = torch.autograd.grad(q.sum() * 0.5 + loss, z_gpu)
    assert torch.equal(ids.cpu(), ids_ref)
    assert torch.equal(q.detach().cpu(), q_ref.detach().contiguous())
    assert abs(loss.item() - loss_ref.item()) <= TOL * abs(loss_ref.item()) + 1e-12
    assert rel_err(gz, gz_ref) <= TOL
    assert rel_err(m.embed, ora.embed) <= TOL


def test_empty_batch():
    m = new_vq(8, 4)
    m.train(True)
    with torch.no_grad():
        m.cluster_size.fill_(2.0)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    q, loss, ids = m(torch.empty(0, 4, 6, 6, device=DEV))
    assert q.shape == (0, 4, 6, 6) and ids.shape == (0, 6, 6)
    # EMA with zero counts: cluster_size decays, embed_avg decays (reference arithmetic on an empty batch)
    assert torch.allclose(m.cluster_size, before["cluster_size"] * 0.99)
    assert m.lookup(torch.empty(0, 6, 6, dtype=torch.long, device=DEV)).shape == (0, 6, 6, 4)


def test_rejects_non_square_and_bad_dtype():
    m = new_vq(8, 4)
    with pytest.raises(ValueError, match="square"):
        m(torch.randn(1, 4, 6, 8, device=DEV))
    with pytest.raises(TypeError):
        m(torch.randn(1, 4, 6, 6, device=DEV, dtype=torch.float16))
    with pytest.raises(ValueError, match="channels"):
        m(torch.randn(1, 5, 6, 6, device=DEV))


def test_duplicate_codes_pick_a_maximiser_lowest_index():
    """Exact ties: the reference's choice is implementation-defined; ours is the lowest index."""
    D, K, H = 8, 6, 4
    g = torch.Generator().manual_seed(5)
    embed = torch.randn(K, D, generator=g)
    embed[4] = embed[1]
    z = embed[torch.tensor([1, 4, 1, 4] * 4)].T.reshape(1, D, H, H).contiguous()
    m = new_vq(K, D)
    set_state(m, embed.numpy(), np.zeros(K, np.float32), embed.T.numpy())
    m.eval()
    q, loss, ids = m(z.to(DEV))
    assert bool((ids == 1).all())
    assert loss.item() == 0.0


def test_eval_and_no_grad_semantics():
    K, D, H = 32, 16, 8
    z, embed = seeded_case(2, D, H, H, K, seed=9)
    m = new_vq(K, D)
    set_state(m, embed.numpy(), np.ones(K, np.float32), embed.T.numpy())
    before = {k: v.clone() for k, v in m.state_dict().items()}
    m.eval()
    zg = z.to(DEV).requires_grad_(True)
    q, loss, ids = m(zg)
    for k, v in m.state_dict().items():
        assert torch.equal(v, before[k]), f"eval() must not touch {k}"
    assert q.requires_grad and loss.requires_grad and not ids.requires_grad
    ids += 1                                        # callers mutate ids in place (vqwnet.py:111)
    with torch.no_grad():
        q2, loss2, ids2 = m(z.to(DEV))
    assert not q2.requires_grad and not loss2.requires_grad
    assert torch.equal(ids2 + 1, ids)
    m.train()
    with torch.no_grad():
        m(z.to(DEV))
    assert not torch.equal(m.cluster_size, before["cluster_size"])   # training updates even under no_grad


def test_backward_partial_grads():
    K, D, H = 32, 16, 8
    z, embed = seeded_case(2, D, H, H, K, seed=10)
    ora = make_oracle(K, D, embed)
    ora.eval()
    m = new_vq(K, D)
    set_state(m, embed.numpy(), np.zeros(K, np.float32), embed.T.numpy())
    m.eval()
    for use_q, use_loss in ((True, False), (False, True)):
        zr = z.clone().requires_grad_(True)
        qr, lr, _ = ora(zr)
        zg = z.to(DEV).requires_grad_(True)
        qg, lg, _ = m(zg)
        obj_r = (qr.pow(2).sum() if use_q else 0) + (3.0 * lr if use_loss else 0)
        obj_g = (qg.pow(2).sum() if use_q else 0) + (3.0 * lg if use_loss else 0)
        (gr,) = torch.autograd.grad(obj_r, zr)
        (gg,) = torch.autograd.grad(obj_g, zg)
        assert rel_err(gg, gr) <= TOL


def test_non_contiguous_input_and_reassigned_codebook():
    K, D, H = 40, 12, 10
    z, embed = seeded_case(2, D, H, H, K, seed=12)
    ora = make_oracle(K, D, embed)
    ora.eval()
    m = new_vq(K, D)
    m.embed = embed.to(DEV)                          # reassignment as in unet_encoder.py:85
    m.eval()
    z_nc = z.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)   # channels-last strides
    assert not z_nc.is_contiguous()
    q, loss, ids = m(z_nc.to(DEV))
    q_ref, loss_ref, ids_ref = ora(z)
    assert torch.equal(ids.cpu(), ids_ref)
    assert torch.equal(q.cpu(), q_ref.contiguous())


# ---------------------------------------------------------------------------------------------
# lookup (vq_module.py:203-206) -- pure gather: bit-exact vs F.embedding
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("K,D", [(10, 16), (512, 64), (64, 512), (7, 3), (4096, 256)])
def test_lookup_matches_embedding(K, D):
    g = torch.Generator().manual_seed(K + D)
    embed = torch.randn(K, D, generator=g)
    m = new_vq(K, D)
    set_state(m, embed.numpy(), np.zeros(K, np.float32), embed.T.numpy())
    for shape in ((2, 16, 16), (1, 33, 17), (5,), (3, 7), (2, 3, 4, 5)):
        ids = torch.randint(0, K, shape, generator=g)
        out = m.lookup(ids.to(DEV))
        ref = F.embedding(ids, embed)
        assert out.shape == ref.shape
        assert torch.equal(out.cpu(), ref)
        if len(shape) == 3:
            # the callers' transpose(1,-1) (unet_encoder.py:120-123) lands on contiguous NCHW memory
            assert out.transpose(1, -1).is_contiguous()
            assert torch.equal(out.transpose(1, -1).cpu(), ref.transpose(1, -1))


def test_recon_config4_lookup_path():
    """BASELINE config 4 (run_recon.py:170-192): K=10, D=16, 512x512 label map -> embedding map."""
    K, D, S = 10, 16, 512
    g = torch.Generator().manual_seed(4)
    embed = torch.randn(K, D, generator=g)
    ids = torch.randint(0, K, (1, S, S), generator=g)
    m = new_vq(K, D, momentum=0.999)
    set_state(m, embed.numpy(), np.zeros(K, np.float32), embed.T.numpy())
    x = m.lookup(torch.transpose(ids.to(DEV), 1, 2)).transpose(1, -1)       # get_embed_from_ids
    ref = F.embedding(torch.transpose(ids, 1, 2), embed).transpose(1, -1)
    assert x.shape == (1, D, S, S) and torch.equal(x.cpu(), ref)


# ---------------------------------------------------------------------------------------------
# (3) properties at BASELINE's full sizes (config 2 quantiser shape: 16x64x256x256, K=512; N = 1M)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("flags", PATHS)
@pytest.mark.parametrize("K,D", [(512, 64), (512, 256), (64, 64)])
def test_full_size_properties(K, D, flags):
    B, H = 16, 256
    N = B * H * H
    gen = torch.Generator(device=DEV).manual_seed(1234)
    z = torch.randn(B, D, H, H, device=DEV, generator=gen)
    embed = torch.randn(K, D, device=DEV, generator=gen)
    m = new_vq(K, D, 0.99, flags)
    with torch.no_grad():
        m.embed.copy_(embed)
        m.embed_avg.copy_(embed.T)
        m.cluster_size.fill_(1.0)
    m.train(True)
    zg = z.clone().requires_grad_(True)
    q, loss, ids = m(zg)
    ids_nat = ids.transpose(1, 2)                                   # [b,h,w]
    # round trip: q == codebook[ids] (pre-update codebook)
    q_chk = F.embedding(ids_nat, embed).permute(0, 3, 1, 2)
    assert torch.equal(q.detach(), q_chk)
    # every pixel got a code; histogram is exact
    counts = torch.bincount(ids.reshape(-1), minlength=K)
    assert int(counts.sum()) == N and int(ids.min()) >= 0 and int(ids.max()) < K
    # the chosen code is a nearest code: no other code is closer than fp32 noise allows (sampled rows)
    samp = torch.randint(0, N, (4096,), device=DEV, generator=gen)
    flat = z.permute(0, 2, 3, 1).reshape(N, D)[samp].double()
    d = torch.cdist(flat, embed.double())
    chosen = ids_nat.reshape(-1)[samp]
    gap = d.gather(1, chosen[:, None]).squeeze(1) - d.min(1).values
    assert float(gap.max()) <= 1e-4
    # loss == mean((z-q)^2)
    ref_loss = (z.double() - q.detach().double()).pow(2).mean().item()
    assert abs(loss.item() - ref_loss) <= TOL * ref_loss
    # EMA statistics: cluster_size = 0.99*1 + 0.01*counts ; sum over codes of embed_avg is linear in z
    cs_ref = 0.99 * 1.0 + (1 - 0.99) * counts.double()
    assert rel_err(m.cluster_size, cs_ref) <= TOL
    sums_ref = torch.zeros(K, D, device=DEV, dtype=torch.float64).index_add_(
        0, ids_nat.reshape(-1), z.permute(0, 2, 3, 1).reshape(N, D).double())
    avg_ref = 0.99 * embed.double().T + (1 - 0.99) * sums_ref.T
    assert rel_err(m.embed_avg, avg_ref) <= TOL
    # backward: g_z = g_q + 2 w (z-q)/numel
    g_q = torch.randn(B, D, H, H, device=DEV, generator=gen)
    (gz,) = torch.autograd.grad((q * g_q).sum() + 2.5 * loss, zg)
    gz_ref = g_q.double() + 2.5 * 2.0 * (z.double() - q.detach().double()) / z.numel()
    assert rel_err(gz, gz_ref) <= TOL
    # idempotence: quantising the quantised map (same codebook) returns the same codes and zero loss
    with torch.no_grad():
        m.embed.copy_(embed)
    m.eval()
    q2, loss2, ids2 = m(q.detach())
    assert torch.equal(ids2, ids) and loss2.item() == 0.0


FULL_SIZE = [(512, 64, "gauss"), (512, 64, "clustered"), (512, 64, "relu"),
             (512, 256, "gauss"), (512, 256, "clustered"), (512, 256, "relu")]


@pytest.mark.parametrize("K,D,kind", FULL_SIZE)
def test_full_size_vs_oracle(K, D, kind):
    """BASELINE config 2 (N = 1 048 576, K = 512, D = 64) and the north-star point (K = 512, D = 256) against the CPU
    oracle itself (chunked over N, same per-element arithmetic as the reference, vq_module.py:45-62): ids and counts
    bit-exact, q bit-exact, loss and EMA buffers <= 1e-5 (max-norm relative, tests/util.py:rel_err).  The rare candidate
    misses of a low-precision search only show up at this scale."""
    B, H = 16, 256
    N = B * H * H
    z, embed = seeded_case(B, D, H, H, K, seed=4242 + D, kind=kind)
    ora = make_oracle(K, D, embed, warmed=True, n_for_warm=N, chunk=65536)
    m = new_vq(K, D, 0.99, 0)
    set_state(m, ora.embed.numpy(), ora.cluster_size.numpy(), ora.embed_avg.numpy())
    ora.train(True)
    m.train(True)
    embed0 = ora.embed.clone()
    with torch.no_grad():
        q_ref, loss_ref, ids_ref = ora(z)
        q, loss, ids = m(z.to(DEV))
    torch.cuda.synchronize()
    nties = assert_ids_match(ids, ids_ref, embed0, z)
    report_ties(nties, N, f"(full size K={K} D={D} {kind})")
    assert abs(loss.item() - loss_ref.item()) <= TOL * abs(loss_ref.item())
    if nties == 0:
        assert torch.equal(q.cpu(), q_ref.contiguous()), "quantized must be bit-exact"
        counts = torch.bincount(ids.reshape(-1), minlength=K).cpu()
        assert torch.equal(counts, torch.bincount(ids_ref.reshape(-1), minlength=K)), "histogram must be bit-exact"
        assert rel_err(m.cluster_size, ora.cluster_size) <= TOL
        assert rel_err(m.embed_avg, ora.embed_avg) <= TOL
        assert rel_err(m.embed, ora.embed) <= TOL


def test_simt_and_auto_paths_agree_at_full_size():
    B, D, H, K = 16, 64, 256, 512
    gen = torch.Generator(device=DEV).manual_seed(99)
    z = torch.randn(B, D, H, H, device=DEV, generator=gen)
    embed = torch.randn(K, D, device=DEV, generator=gen)
    outs = []
    for flags in (_native.VQ_FLAG_FORCE_SIMT, 0):
        m = new_vq(K, D, 0.99, flags)
        with torch.no_grad():
            m.embed.copy_(embed)
            m.embed_avg.copy_(embed.T)
        m.train(True)
        q, loss, ids = m(z)
        outs.append((ids, q, loss, m.cluster_size.clone(), m.embed.clone()))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])
    assert abs(outs[0][2].item() - outs[1][2].item()) <= TOL * outs[0][2].item()
    assert torch.equal(outs[0][3], outs[1][3])
    assert rel_err(outs[1][4], outs[0][4]) <= TOL


# The streamed-codebook kernel can run as 2-CTA clusters (VQ_FLAG_PAIR; cta_group::2: each CTA holds half of every 256-code
# slice); the default is one CTA per tile.  Both must produce the same bits.
PAIR_SHAPES = [(16, 256, 512, 256, "clustered"), (4, 128, 512, 128, "gauss"), (2, 64, 4096, 64, "gauss"),
               (1, 16, 512, 256, "relu"), (3, 48, 1000, 132, "gauss"), (2, 96, 4096, 256, "clustered")]


@pytest.mark.parametrize("B,H,K,D,kind", PAIR_SHAPES)
def test_streamed_pair_and_single_cta_kernels_agree(B, H, K, D, kind):
    gen = torch.Generator(device=DEV).manual_seed(77 + K + D)
    embed = torch.randn(K, D, device=DEV, generator=gen)
    if kind == "clustered":
        z = embed[torch.randint(0, K, (B, H, H), device=DEV, generator=gen)].permute(0, 3, 1, 2).contiguous()
        z = z + 0.1 * torch.randn(B, D, H, H, device=DEV, generator=gen)
    else:
        z = torch.randn(B, D, H, H, device=DEV, generator=gen)
        if kind == "relu":
            z = torch.relu(z)
    outs = []
    for flags in (0, _native.VQ_FLAG_PAIR, _native.VQ_FLAG_FORCE_SIMT):
        m = new_vq(K, D, 0.99, flags)
        with torch.no_grad():
            m.embed.copy_(embed)
            m.cluster_size.fill_(float(B * H * H) / K)
            m.embed_avg.copy_((embed * m.cluster_size[:, None]).T)
        m.eval()
        with torch.no_grad():
            q2, loss2, ids2 = m(z)                                 # the inference instantiation (no statistics)
        m.train(True)
        zz = z.clone().requires_grad_(True)
        q, loss, ids = m(zz)
        (gz,) = torch.autograd.grad(q.sum() + loss, zz)
        outs.append((ids, q.detach(), loss.detach(), m.cluster_size.clone(), m.embed.clone(), gz, ids2, q2, loss2))
    for o in outs[1:]:
        assert torch.equal(outs[0][0], o[0]) and torch.equal(outs[0][1], o[1])
        assert abs(outs[0][2].item() - o[2].item()) <= TOL * abs(outs[0][2].item())
        assert torch.equal(outs[0][3], o[3])
        assert rel_err(o[4], outs[0][4]) <= TOL and rel_err(o[5], outs[0][5]) <= TOL
        assert torch.equal(outs[0][6], o[6]) and torch.equal(outs[0][7], o[7])
        assert abs(outs[0][8].item() - o[8].item()) <= TOL * abs(outs[0][8].item())


@pytest.mark.parametrize("B,H,W", [(3, 8, 16), (1, 16, 24), (5, 8, 48)])
def test_streamed_pair_kernel_odd_tile_count(B, H, W):
    """An odd number of 128-pixel tiles (only reachable through the C-ABI: the module takes square maps, whose tile
    count is even): the odd CTA of the last pair gets a tile past the end and must not write anything."""
    K, D = 512, 128
    L = pkg.lib()
    gen = torch.Generator(device=DEV).manual_seed(5)
    z = torch.randn(B, D, H, W, device=DEV, generator=gen)
    embed = torch.randn(K, D, device=DEV, generator=gen)
    assert (B * H * W // 128) % 2 == 1
    res = []
    for flags in (_native.VQ_FLAG_FORCE_TC, _native.VQ_FLAG_PAIR | _native.VQ_FLAG_FORCE_TC, _native.VQ_FLAG_FORCE_SIMT):
        q = torch.full((B, D, H, W), -7.0, device=DEV)
        guard = torch.full((4096,), -7.0, device=DEV)              # memory right behind q must stay untouched
        ids = torch.full((B, W, H), -1, dtype=torch.int64, device=DEV)
        loss = torch.zeros((), device=DEV)
        stats = torch.zeros(L.vq_stats_floats(K, D), device=DEV)
        ws = torch.empty(L.vq_workspace_bytes(B * H * W, K, D), dtype=torch.uint8, device=DEV)
        _native.check(L.vq_assign_fwd(z.data_ptr(), B, D, H, W, embed.data_ptr(), K, ids.data_ptr(), None, q.data_ptr(),
                                      loss.data_ptr(), stats.data_ptr(), None, ws.data_ptr(), ws.numel(), flags,
                                      torch.cuda.current_stream().cuda_stream), "vq_assign_fwd")
        torch.cuda.synchronize()
        assert bool((guard == -7.0).all())
        res.append((ids, q, loss, stats))
    for r in res[1:]:
        assert torch.equal(res[0][0], r[0]) and torch.equal(res[0][1], r[1])
        assert abs(res[0][2].item() - r[2].item()) <= TOL * abs(res[0][2].item())
        assert rel_err(r[3], res[0][3]) <= TOL
    assert int(res[1][0].min()) >= 0


def test_multistep_cold_training_paths_agree():
    """Several training steps from the cold state (cluster_size == 0): dead codes explode, many pixels
    collapse onto a few live codes, candidate lists overflow -- the tensor-core path (with its exhaustive
    fallback) must keep choosing exactly what the fp32 CUDA-core search chooses."""
    B, D, H, K = 4, 64, 256, 512
    gen = torch.Generator(device=DEV).manual_seed(2024)
    embed = torch.randn(K, D, device=DEV, generator=gen)
    ms = []
    for flags in (_native.VQ_FLAG_FORCE_SIMT, 0):
        m = new_vq(K, D, 0.99, flags)
        with torch.no_grad():
            m.embed.copy_(embed)
            m.embed_avg.copy_(embed.T)
        m.train(True)
        ms.append(m)
    assert _native.lib().vq_assign_path(B, D, H, H, K, 0) == 1
    for step in range(5):
        z = torch.randn(B, D, H, H, device=DEV, generator=gen)
        if step == 3:
            z = torch.relu(z)                       # post-ReLU statistics: many exact zeros
        with torch.no_grad():
            outs = [m(z) for m in ms]
        assert torch.equal(outs[0][2], outs[1][2]), f"step {step}: ids differ"
        assert torch.equal(outs[0][0], outs[1][0]), f"step {step}: quantized differs"
        assert abs(outs[0][1].item() - outs[1][1].item()) <= TOL * abs(outs[0][1].item())
        assert torch.equal(ms[0].cluster_size, ms[1].cluster_size)
        assert rel_err(ms[1].embed, ms[0].embed) <= TOL
        with torch.no_grad():                       # keep the two replicas bit-identical for the next step
            for a, b in zip(ms[1].buffers(), ms[0].buffers()):
                a.copy_(b)


@pytest.mark.parametrize("flags", PATHS)
@pytest.mark.parametrize("n_micro", [2, 4])
def test_micro_batch_accumulation_equals_one_big_batch(n_micro, flags):
    """`accumulate_steps=n` (BASELINE config 5 at 2 / 4 GPUs: 64 / 32 slices of 512^2 per GPU do not fit one forward of
    the full network): n training forwards on the n micro-batches == ONE forward of the oracle on the concatenated
    batch -- ids / counts bit-exact, buffers <= 1e-5; the codebook must not move before the last micro-batch."""
    B, D, H, K = 8, 64, 32, 512
    z, embed = seeded_case(B, D, H, H, K, seed=515 + n_micro, kind="clustered")
    ora = make_oracle(K, D, embed, warmed=True, n_for_warm=B * H * H)
    ora.train(True)
    m = new_vq(K, D, 0.99, flags, accumulate_steps=n_micro)
    set_state(m, ora.embed.numpy(), ora.cluster_size.numpy(), ora.embed_avg.numpy())
    m.train(True)
    embed0 = m.embed.clone()
    with torch.no_grad():
        _, loss_ref, ids_ref = ora(z)
    mb = B // n_micro
    ids_parts, losses = [], []
    for i in range(n_micro):
        zi = z[i * mb:(i + 1) * mb].to(DEV).requires_grad_(True)
        q, loss, ids = m(zi)
        (g,) = torch.autograd.grad(q.sum() + loss, zi)
        ids_parts.append(ids.cpu())
        losses.append(loss.item())
        if i < n_micro - 1:
            assert torch.equal(m.embed, embed0), "the codebook moved before the last micro-batch"
    assert torch.equal(torch.cat(ids_parts), ids_ref)
    assert abs(sum(losses) / n_micro - loss_ref.item()) <= TOL * abs(loss_ref.item())
    assert rel_err(m.cluster_size, ora.cluster_size) <= TOL
    assert rel_err(m.embed_avg, ora.embed_avg) <= TOL
    assert rel_err(m.embed, ora.embed) <= TOL
    # a partial accumulation is applied by flush_ema(); nothing is pending afterwards
    m(z[:mb].to(DEV))
    before = m.cluster_size.clone()
    assert m.flush_ema() and not torch.equal(m.cluster_size, before)
    assert not m.flush_ema()


@pytest.mark.parametrize("flags", PATHS)
@pytest.mark.parametrize("K,D,H", [(512, 64, 32), (10, 16, 64), (64, 256, 16), (100, 24, 8)])
def test_natural_one_based_ids_from_the_epilogue(K, D, H, flags):
    """`ids_layout="natural", ids_base=1` == the reference module followed by what every caller does
    next, `ids = transpose(ids, 1, 2); ids += 1` (vqwnet.py:110-111, unet_encoder.py:115-116) -- all search kernels."""
    B = 3
    z, embed = seeded_case(B, D, H, H, K, seed=77 + K)
    ora = make_oracle(K, D, embed)
    ora.eval()
    with torch.no_grad():
        _, _, ids_ref = ora(z)
    want = torch.transpose(ids_ref, 1, 2) + 1
    for training in (False, True):
        m = new_vq(K, D, 0.99, flags, ids_layout="natural", ids_base=1)
        set_state(m, ora.embed.numpy(), ora.cluster_size.numpy(), ora.embed_avg.numpy())
        m.train(training)
        zz = z.to(DEV).requires_grad_(True)
        q, loss, ids = m(zz)
        (g,) = torch.autograd.grad(q.sum() + loss, zz)        # the backward keeps using the 0-based natural map
        assert ids.dtype == torch.int64 and ids.shape == (B, H, H)
        assert torch.equal(ids.cpu(), want)
        assert torch.equal(q.detach().cpu(), F.embedding(ids_ref, ora.embed).transpose(1, -1).contiguous())
        assert torch.isfinite(g).all()
    # natural order lifts the reference's square-only restriction
    zr = torch.randn(2, D, 8, 16, generator=torch.Generator().manual_seed(5))
    m = new_vq(K, D, 0.99, flags, ids_layout="natural")
    set_state(m, ora.embed.numpy(), ora.cluster_size.numpy(), ora.embed_avg.numpy())
    m.eval()
    with torch.no_grad():
        _, _, ids = m(zr.to(DEV))
        flat = zr.permute(0, 2, 3, 1).reshape(-1, D)
        from oracle.vq_oracle import torch_knn_l2
        _, ids_flat = torch_knn_l2(ora.embed, flat)
    assert torch.equal(ids.cpu().reshape(-1), ids_flat.reshape(-1))


def test_embed_avg_layouts():
    """`embed.T.clone()` (vq_module.py:156) keeps strides (1, D); a checkpoint round trip can make the
    buffer contiguous.  Both must give the reference's update."""
    K, D, H = 48, 20, 8
    z, embed = seeded_case(2, D, H, H, K, seed=44)
    ora = make_oracle(K, D, embed, warmed=True, n_for_warm=2 * H * H)
    ora.train(True)
    ms = [new_vq(K, D), new_vq(K, D)]
    assert ms[0].embed_avg.stride() == (1, D)
    ms[1].embed_avg = ms[1].embed_avg.contiguous()
    assert ms[1].embed_avg.stride() == (K, 1)
    for m in ms:
        set_state(m, ora.embed.numpy(), ora.cluster_size.numpy(), ora.embed_avg.numpy())
        m.train(True)
    ora(z)
    for m in ms:
        m(z.to(DEV))
        assert rel_err(m.embed_avg, ora.embed_avg) <= TOL
        assert rel_err(m.embed, ora.embed) <= TOL
        assert rel_err(m.cluster_size, ora.cluster_size) <= TOL


# ---------------------------------------------------------------------------------------------
# multi-GPU: 2 NCCL ranks == one process on the concatenated batch (skipped on single-GPU boxes)
# ---------------------------------------------------------------------------------------------
def _nccl_worker(rank, ws, port, ret, overlap):
    import os
    import torch.distributed as dist
    os.environ.update(WORLD_SIZE=str(ws), RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", rank))
    dev = f"cuda:{rank}"
    B, D, H, K = 4, 64, 128, 512
    gen = torch.Generator().manual_seed(77)
    zs = [torch.randn(B, D, H, H, generator=gen) for _ in range(2)]
    embed = torch.randn(K, D, generator=gen)
    m = pkg.VQ(emb_dim=D, dict_size=K, momentum=0.99, eps=1e-5, knn_backend="torch", reduce_mode="sum",
               overlap_exchange=overlap).to(dev)
    with torch.no_grad():
        m.embed.copy_(embed)
        m.embed_avg.copy_(embed.T)
        m.cluster_size.fill_(1.0)
    m.train(True)
    half = B // ws
    out = {}
    for s_, z in enumerate(zs):                    # two steps: the second forward must see the first step's update
        zr = z[rank * half:(rank + 1) * half].to(dev).requires_grad_(True)
        q, loss, ids = m(zr)
        (gz,) = torch.autograd.grad(q.sum() + loss, zr)      # backward overlaps the exchange when `overlap`
        out[f"ids{s_}"] = ids.cpu().numpy()
    out["embed"] = m.get_codebook().t().contiguous().cpu().numpy()      # joins the side stream
    out["cs"] = m.state_dict()["cluster_size"].cpu().numpy()
    torch.cuda.synchronize()
    ret[rank] = out
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("overlap", [False, True], ids=["inline", "overlapped"])
def test_two_rank_nccl_matches_single_process(overlap):
    import os
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_nccl_worker, args=(2, 29761 + int(overlap), ret, overlap), nprocs=2, join=True)
    os.environ.pop("WORLD_SIZE", None)
    os.environ.pop("RANK", None)
    B, D, H, K = 4, 64, 128, 512
    gen = torch.Generator().manual_seed(77)
    zs = [torch.randn(B, D, H, H, generator=gen) for _ in range(2)]
    embed = torch.randn(K, D, generator=gen)
    m = new_vq(K, D, 0.99, 0)
    with torch.no_grad():
        m.embed.copy_(embed.to(DEV))
        m.embed_avg.copy_(embed.T.to(DEV))
        m.cluster_size.fill_(1.0)
    m.train(True)
    for s_, z in enumerate(zs):
        _, _, ids = m(z.to(DEV))
        assert np.array_equal(np.concatenate([ret[0][f"ids{s_}"], ret[1][f"ids{s_}"]]), ids.cpu().numpy())
    for r in range(2):
        assert np.array_equal(ret[r]["cs"], m.cluster_size.cpu().numpy())
        assert rel_err(t(ret[r]["embed"]), m.embed) <= TOL
    assert np.array_equal(ret[0]["embed"], ret[1]["embed"])          # replicas bit-identical without broadcast
