#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 VQ bottleneck (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own module on the host cores
    python bench.py --workload vqwnet512 --gpus N ...        # BASELINE config 5: 128 slices of 512 x 512, data parallel

A "step" is one pass of the hot path over one batch of synthetic input: the quantiser's training step
(nearest-code search + gather + commitment loss + EMA statistics/update [+ packed all-reduce when
N > 1] + straight-through/commitment backward) on the quantiser input of BASELINE config 2
(VQ-W-Net training step, batch 16 of 256x256 slices -> z = 16x64x256x256, K = 512): 1,048,576
lookups per GPU per step (weak scaling: 16 slices per GPU).

One JSON line on stdout (rank 0).  `value` = whole-job lookups/s with inputs resident in HBM (the timed region replays
CUDA graphs of the step; `eager` = the same loop launched from Python, `--no-graphs` makes that the timed region);
`e2e` = the same through the module's public API with HOST (pinned) input and the result (loss +
code map) read back every step; `roofline` = the dominant kernel (nearest-code search) against the
measured HBM peak; `cpu_baseline` = the reference's own `VQModule` (git-ignored mirror baseline/_ref/src made by
`__graft_entry__.build()`; the oracle port when no copy is reachable) on the host cores, bounded sample, with
`cpu_baseline.reference_on_gpu` = the same unmodified module under stock torch-CUDA on this GPU (sanity / noise floor:
its step time and the number of code ids that differ from this library's in an eval forward); `north_star` =
the K = 512, D = 256 point (eval and train, clustered / ReLU / Gaussian input) with its own roofline fractions;
`parity_check` = TC-vs-CUDA-core ids on one buffer and, for N > 1, the all-reduced statistics against the gathered
local ones.  `vqwnet_train` = the other half of BASELINE's metric: VQ-W-Net training slices/s
(BASELINE config 2: batch 16 of 256x256 slices per GPU, tools/wnet.py around this repo's quantiser, stock cuDNN
convolutions, Adam), device-resident and end-to-end, measured after the headline region; `--workload vqwnet`
makes it the line's own metric (and `--impl reference --workload vqwnet` times the same network on the host cores).
`--fused-norm all|tail|none` (default all): the harness's InstanceNorm2d + ReLU pairs run through vq_norm_relu_fwd/bwd
(`tail` = only the pair that produces the quantiser's input, `none` = stock torch layers).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

# BASELINE config 2 quantiser shape
CFG = dict(B=16, D=64, H=256, K=512, momentum=0.99, eps=1e-5)
WORKLOADS = {
    "config2": dict(B=16, D=64, H=256, K=512),
    "k512d256": dict(B=16, D=256, H=256, K=512),
    "k64d64": dict(B=16, D=64, H=256, K=64),
    "k4096d64": dict(B=16, D=64, H=256, K=4096),
    "k64d256": dict(B=16, D=256, H=256, K=64),
    "k4096d256": dict(B=16, D=256, H=256, K=4096),
    "recon_k10d16": dict(B=16, D=16, H=512, K=10),          # run_recon.py:27-48 (LungConfig): K = 10, D = 16, 512 x 512 slices
}
METRIC = "vq_lookups_per_s"
UNIT = "lookups/s"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1590.0)), "measured"
    except Exception:
        return 6650.0, 1590.0, "fallback"


# ---------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe) during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.lines = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ts, line in self.lines:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(self.NAMES, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference op chain on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_quantiser(D, K):
    """(module, kind): the UNMODIFIED reference `VQModule` (vq_module.py:139-211) when a copy of the reference is
    reachable (/root/reference/src here, the git-ignored mirror baseline/_ref/src on the GPU box), else the oracle port."""
    try:
        from oracle import ref_loader
        if ref_loader.reference_available():
            VQ = ref_loader.load_reference_vq()
            return VQ(emb_dim=D, dict_size=K, momentum=CFG["momentum"], eps=CFG["eps"], knn_backend="torch"), "reference"
    except Exception:
        pass
    from oracle.vq_oracle import OracleVQ
    return OracleVQ(D, K, CFG["momentum"], CFG["eps"], "torch", chunk=65536), "port"


def cpu_step_factory(wl, slices):
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(1234)
    D, H, K = wl["D"], wl["H"], wl["K"]
    m, kind = cpu_quantiser(D, K)
    m.train(True)
    with torch.no_grad():       # same warmed EMA state as the B200 arm
        cs = torch.rand(K, generator=torch.Generator().manual_seed(1234)) * (slices * H * H / K) + 1.0
        m.cluster_size.copy_(cs)
        m.embed_avg.copy_((m.embed * cs[:, None]).T)
    zs = [torch.randn(slices, D, H, H, generator=g) for _ in range(2)]
    g_q = torch.randn(slices, D, H, H, generator=g)
    one = torch.ones(())
    state = {"i": 0}

    def step():
        z = zs[state["i"] % 2].requires_grad_(True)
        state["i"] += 1
        q, loss, ids = m(z)
        torch.autograd.grad((q, loss), z, (g_q, one))
        return loss

    step.kind = kind
    return step, slices * H * H


def run_cpu_baseline(wl, budget_s=12.0, slices=1):
    step, n = cpu_step_factory(wl, slices)
    step()
    times = []
    t_all = time.perf_counter()
    while len(times) < 3 or (time.perf_counter() - t_all < budget_s and len(times) < 50):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return {"value": n / med, "unit": UNIT, "cores": torch.get_num_threads(), "kind": step.kind,
            "sample": f"{slices} slice(s) of the workload ({n} lookups) per step, quantiser train step "
                      f"(fwd+EMA+bwd) of {'the unmodified reference VQModule' if step.kind == 'reference' else 'the oracle port'} "
                      f"on the host cores, median of {len(times)} steps, {med * 1e3:.1f} ms/step"}


def run_reference_on_gpu(wl, dev, vq_b200, steps=3):
    """Sanity / noise-floor leg: the UNMODIFIED reference `VQModule` run by stock torch-CUDA on the same
    B200, same seeded input and EMA state as the B200 arm -- training step (forward + EMA + backward) time, and how many
    code ids of an eval forward differ between the reference's fp32 cuBLAS GEMM + topk and this library (ties /
    rounding of the reference's own GEMM: the noise floor of 'bit-exact ids').  None when no copy of the reference is
    reachable.  Part of the baseline legs: nothing of it runs inside the timed regions of the B200 arm."""
    try:
        from oracle import ref_loader
        if not ref_loader.reference_available():
            return None
        RefVQ = ref_loader.load_reference_vq()
    except Exception as exc:
        return {"error": f"{type(exc).__name__}: {exc}"[:200]}
    B, D, H, K = wl["B"], wl["D"], wl["H"], wl["K"]
    n = B * H * H
    if n * K * 16 > 60e9:                                          # N x K one-hot (int64 + fp32) + K x N scores must fit
        return {"skipped": f"reference scratch for N = {n}, K = {K} would need {n * K * 16 / 1e9:.0f} GB"}
    os.environ["WORLD_SIZE"] = os.environ.get("WORLD_SIZE", "1")
    try:
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats(dev)
        gen = torch.Generator(device=dev).manual_seed(4242)
        ref = RefVQ(emb_dim=D, dict_size=K, momentum=CFG["momentum"], eps=CFG["eps"], knn_backend="torch").to(dev)
        with torch.no_grad():
            ref.embed.copy_(vq_b200.embed)
            ref.cluster_size.copy_(vq_b200.cluster_size)
            ref.embed_avg.copy_(vq_b200.embed_avg)
        z = torch.randn(B, D, H, H, device=dev, generator=gen)
        g_q = torch.randn(B, D, H, H, device=dev, generator=gen)
        one = torch.ones((), device=dev)
        # eval forward of both on the same input: differing ids
        ref.eval()
        was_training = vq_b200.training
        vq_b200.eval()
        with torch.no_grad():
            _, loss_r, ids_r = ref(z)
            _, loss_b, ids_b = vq_b200(z)
        vq_b200.train(was_training)
        differ = int((ids_r != ids_b).sum().item())
        loss_rel = abs(float(loss_r) - float(loss_b)) / max(abs(float(loss_r)), 1e-30)
        del ids_r, ids_b
        ref.train(True)

        def step():
            zz = z.detach().requires_grad_(True)
            q, loss, ids = ref(zz)
            torch.autograd.grad((q, loss), zz, (g_q, one))

        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        peak_gb = torch.cuda.max_memory_allocated(dev) / 1e9
        del ref, z, g_q
        torch.cuda.empty_cache()
        return {"value": n / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
                "what": "the unmodified reference VQModule (vq_module.py:139-211) under stock torch-CUDA (fp32 cuBLAS, "
                        "allow_tf32 off) on this GPU: quantiser train step, same shape / EMA state as the B200 arm",
                "eval_ids_differing_from_b200": differ, "eval_loss_rel_diff": loss_rel, "lookups": n,
                "peak_memory_gb": peak_gb, "torch": torch.__version__}
    except Exception as exc:
        torch.cuda.empty_cache()
        return {"error": f"{type(exc).__name__}: {exc}"[:300]}


def run_reference_arm(args, wl, wl_name):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    # rank 0 alone times the CPU path: the reference gates its all_reduce on $WORLD_SIZE (utils/__init__.py:109-114),
    # which torchrun sets for the N > 1 launches -- this leg is a single process without a process group
    os.environ["WORLD_SIZE"] = "1"
    # bounded sample: as many of the workload's slices per step as keep the whole run near two minutes (the reference
    # materialises K x N scores and two N x K one-hot matrices per call: ~9 GB and seconds per step at the full 16 slices)
    slices = 2
    step, n = cpu_step_factory(wl, slices)
    step()
    t0 = time.perf_counter()
    step()
    per_slice = (time.perf_counter() - t0) / slices
    budget = float(os.environ.get("VQ_REF_BUDGET_S", 120.0))
    want = slices
    for cand in (4, 8, wl["B"]):
        if cand <= wl["B"] and per_slice * cand * (args.steps + max(1, min(args.warmup, 2))) <= budget:
            want = cand
    if want != slices:
        slices = want
        step, n = cpu_step_factory(wl, slices)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl, wl_name, args.gpus, extra={"reference_sample_slices_per_step": slices}),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": step.kind,
                         "sample": f"{slices} of the workload's {wl['B']} slices ({n} lookups) per step; "
                                   + ("the unmodified reference VQModule (baseline/_ref mirror) on the host cores"
                                      if step.kind == "reference" else "oracle port of vq_module.py on the host cores "
                                      "(no copy of the reference reachable)")},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def workload_config(wl, wl_name, gpus, extra=None):
    c = {"workload": f"{wl_name}: VQ-W-Net quantiser train step (search+gather+loss+EMA+backward), "
                     f"z = {wl['B']}x{wl['D']}x{wl['H']}x{wl['H']} fp32 per GPU, K = {wl['K']} codes",
         "slices_per_gpu": wl["B"], "emb_dim": wl["D"], "dict_size": wl["K"], "resolution": wl["H"],
         "lookups_per_gpu_per_step": wl["B"] * wl["H"] * wl["H"], "parallelism": f"dp{gpus}",
         "l2": "inputs rotate over 4 buffers of >=268 MB each (> 126 MB L2)"}
    if extra:
        c.update(extra)
    return c


def load_traffic(key):
    """(bytes, source) of a pre-recorded `ncu --set full` capture of this workload's search kernel, or (None, None).
    The figure is NOT measured in this run: it is copied from profiles/traffic.json, which names the capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f).get(key)
        if tj:
            return tj["bytes"], f"pre-recorded ncu capture {tj.get('source', '?')} via profiles/traffic.json (not measured in this run)"
    except Exception:
        pass
    return None, None


def kernel_time_ms(L, fn, launches, warm=2):
    """Average CUDA-event time of the search kernel (library events on the launching stream) over `launches` calls."""
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    tot, nl = ctypes.c_double(0), ctypes.c_int(0)
    L.vq_profile_enable(1)
    L.vq_profile_read(None, None)
    for i in range(launches):
        fn(warm + i)
    torch.cuda.synchronize()
    L.vq_profile_read(ctypes.byref(tot), ctypes.byref(nl))
    L.vq_profile_enable(0)
    return (tot.value / nl.value) if nl.value else None


def measure_north_star(dev, L, pkg, launches=6):
    """The north-star point of BASELINE.json: fused quantiser at K = 512, D = 256, N = 1 048 576 on one GPU, eval and
    train instantiations of the search kernel, on the three synthetic inputs of SURVEY 8(d)."""
    B, D, H, K = 16, 256, 256, 512
    N = B * H * H
    hbm, bf16, which = measured_peaks()
    alg_bytes = N * (8 * D + 8)
    alg_flops = 2.0 * K * D * N
    t_hbm, t_tc = alg_bytes / (hbm * 1e9), alg_flops / (bf16 * 1e12)
    out = {"workload": f"k512d256: z = {B}x{D}x{H}x{H} fp32, K = {K}", "algorithmic_bytes_per_launch": alg_bytes,
           "algorithmic_flops_per_launch": alg_flops, "roofline_ms": max(t_hbm, t_tc) * 1e3,
           "bound": "hbm" if t_hbm >= t_tc else "tensor", "peak_source": f"MEASURED_PEAKS.json ({which})", "cases": {}}
    gen = torch.Generator(device=dev).manual_seed(4321)
    vq = pkg.VQ(emb_dim=D, dict_size=K, momentum=CFG["momentum"], eps=CFG["eps"], knn_backend="torch").to(dev)
    embed0 = vq.embed.detach().clone()
    cs0 = (torch.rand(K, generator=torch.Generator().manual_seed(1234)) * (N / K) + 1.0).to(dev)

    def reset():
        with torch.no_grad():
            vq.embed.copy_(embed0)
            vq.cluster_size.copy_(cs0)
            vq.embed_avg.copy_((embed0 * cs0[:, None]).T)

    for kind in ("clustered", "relu", "gauss"):
        if kind == "clustered":
            mk = lambda: (embed0[torch.randint(0, K, (B, H, H), device=dev, generator=gen)].permute(0, 3, 1, 2)
                          + 0.1 * torch.randn(B, D, H, H, device=dev, generator=gen)).contiguous()
        elif kind == "relu":
            mk = lambda: torch.relu(torch.randn(B, D, H, H, device=dev, generator=gen))
        else:
            mk = lambda: torch.randn(B, D, H, H, device=dev, generator=gen)
        zs = [mk() for _ in range(2)]                       # 2 x 1 GB: every launch reads a buffer >> L2
        case = {}
        for mode in ("eval", "train"):
            reset()
            vq.train(mode == "train")
            with torch.no_grad():
                ms = kernel_time_ms(L, lambda i: vq(zs[i % 2]), launches)
            traffic, src = load_traffic(f"k512d256_{mode}_{kind}")
            ach = alg_bytes / (ms * 1e-3) / 1e9 if ms else None
            case[mode] = {"kernel_ms": ms, "achieved_gbs": ach, "frac": (max(t_hbm, t_tc) * 1e3 / ms) if ms else None,
                          "traffic": traffic, "traffic_source": src}
        out["cases"][kind] = case
        del zs
        torch.cuda.empty_cache()
    return out


def parity_self_check(dev, L, vq, z, world):
    """Outside the timed region: (a) the tensor-core search and the CUDA-core search give the same code map on one
    buffer; (b) N > 1: the all-reduced packed statistics equal the sum of the all-gathered local ones (counts bit for
    bit; the fp32 sums to 1e-6 -- NCCL's reduction order is not the gather order)."""
    import torch.distributed as dist
    res = {}
    was_training, flags = vq.training, vq.kernel_flags
    vq.eval()
    with torch.no_grad():
        vq.kernel_flags = 0
        ids_a = vq(z)[2].clone()
        vq.kernel_flags = 1
        ids_b = vq(z)[2].clone()
    vq.kernel_flags = flags
    vq.train(was_training)
    res["tc_vs_simt_ids_equal"] = bool(torch.equal(ids_a, ids_b))
    ok = res["tc_vs_simt_ids_equal"]
    if world > 1:
        B, D, H, W = z.shape
        K = vq.embed.shape[0]
        n = B * H * W
        stats = torch.empty(L.vq_stats_floats(K, D), dtype=torch.float32, device=dev)
        q = torch.empty_like(z)
        ids = torch.empty((B, H, W), dtype=torch.int64, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        ws = torch.empty(L.vq_workspace_bytes(n, K, D), dtype=torch.uint8, device=dev)
        rc = L.vq_assign_fwd(z.detach().data_ptr(), B, D, H, W, vq.embed.data_ptr(), K, ids.data_ptr(), None, q.data_ptr(),
                             loss.data_ptr(), stats.data_ptr(), None, ws.data_ptr(), ws.numel(), 0,
                             torch.cuda.current_stream().cuda_stream)
        assert rc == 0, L.vq_last_error()
        gathered = [torch.empty_like(stats) for _ in range(world)]
        dist.all_gather(gathered, stats)
        total = gathered[0].clone()
        for g_ in gathered[1:]:
            total += g_
        reduced = stats.clone()
        dist.all_reduce(reduced)
        off = L.vq_stats_sums_offset(K)
        res["counts_bit_exact"] = bool(torch.equal(reduced[:off], total[:off]))
        denom = total[off:].abs().max().clamp_min(1e-30)
        res["sums_max_rel"] = float(((reduced[off:] - total[off:]).abs().max() / denom).item())
        ok = ok and res["counts_bit_exact"] and res["sums_max_rel"] <= 1e-6
        okt = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        ok = bool(okt.item() > 0.5)
    res["ok"] = ok
    return res


# ---------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------
def run_b200_arm(args, wl, wl_name):
    import torch.distributed as dist
    import medical_image_editing_b200 as pkg

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the quantiser has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    L = pkg.lib()

    B, D, H, K = wl["B"], wl["D"], wl["H"], wl["K"]
    n_per_gpu = B * H * H
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    NBUF = 4
    zbufs = [torch.randn(B, D, H, H, device=dev, generator=gen).requires_grad_(True) for _ in range(NBUF)]
    g_q = torch.randn(B, D, H, H, device=dev, generator=gen)
    one = torch.ones((), device=dev)
    # N > 1: the all-reduce of the EMA statistics and the EMA update run on a side stream, behind the backward
    vq = pkg.VQ(emb_dim=D, dict_size=K, momentum=CFG["momentum"], eps=CFG["eps"], knn_backend="torch", reduce_mode="sum",
                overlap_exchange=(world > 1 and not args.inline_exchange)).to(dev)
    if args.simt:
        vq.kernel_flags = 1
    vq.train(True)
    if not args.cold:
        # steady-state ("warmed") EMA state of SURVEY 8(d): cluster_size = rand(K) * N/K + 1 and a consistent
        # embed_avg = embed * cluster_size, instead of the first-step state (cluster_size = 0) whose update blows
        # the unused codes up by ~1e5 -- that transient is covered by the parity tests
        with torch.no_grad():
            gcpu = torch.Generator().manual_seed(1234)
            cs = torch.rand(K, generator=gcpu) * (world * n_per_gpu / K) + 1.0
            vq.cluster_size.copy_(cs.to(dev))
            vq.embed_avg.copy_((vq.embed * vq.cluster_size[:, None]).T)
    if world > 1:      # identical codebooks on every rank (DDP would broadcast rank 0's buffers once)
        from medical_image_editing_b200.src.trainers import broadcast_module_state
        broadcast_module_state(vq, 0)
    path = L.vq_assign_path(B, D, H, H, K, vq.kernel_flags)

    def step(i):
        z = zbufs[i % NBUF]
        q, loss, ids = vq(z)
        (g_z,) = torch.autograd.grad((q, loss), z, (g_q, one))
        return loss, ids, g_z

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()

    # A step is ~0.4 ms of GPU work behind ~25 host-side enqueues (allocations, three C-ABI calls, for N > 1 the NCCL
    # launch and a stream fork/join), and the ranks of one box share its host cores: the eager loop is launch-bound
    # (N = 1: 0.43 ms eager against 0.40 ms replayed; N = 8: 0.51 to 0.59 ms eager from run to run).  The timed region
    # therefore replays CUDA graphs -- one per input buffer: forward + overlapped statistics exchange + backward, the
    # exchange joined inside the step (`--no-graphs`: eager loop).  Events cannot bracket a kernel inside a graph, so
    # the search kernel's own time (roofline) is measured with the library's events over the same number of EAGER steps
    # right before; that eager loop is itself timed and reported as `eager`.
    use_graphs = not args.no_graphs
    sampler = ClockSampler(local) if rank == 0 else None
    tot_ms, nl = ctypes.c_double(0), ctypes.c_int(0)
    L.vq_profile_enable(1)
    L.vq_profile_read(None, None)
    launches0 = L.vq_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    launches = L.vq_launch_count() - launches0               # per-step launches are the same when the graphs replay them
    L.vq_profile_read(ctypes.byref(tot_ms), ctypes.byref(nl))
    L.vq_profile_enable(0)
    eager_ms = e0.elapsed_time(e1)
    elapsed_ms = eager_ms
    graph_note = None
    if use_graphs:
        vq.sync_codebook()
        graphs = []
        try:
            pool = torch.cuda.graph_pool_handle()
            cap = torch.cuda.Stream(device=dev)
            for bidx in range(NBUF):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool, stream=cap):
                    out_b = step(bidx)
                    vq.sync_codebook()                         # join the side stream inside the captured step
                graphs.append((g, out_b))
            ok = 1.0
        except Exception as exc:                               # keep the line: fall back to the eager loop's time
            graph_note = f"graph capture failed ({type(exc).__name__}: {exc})"[:200]
            ok = 0.0
        okt = torch.tensor([ok], device=dev)
        if world > 1:                                          # capture is rank-local (NCCL only talks at replay): agree first
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        use_graphs = bool(okt.item() > 0.5)
        if not use_graphs:
            graphs = []
            graph_note = graph_note or "graph capture failed on another rank"
            torch.cuda.synchronize()
    if use_graphs:
        for bidx in range(NBUF):                               # one untimed replay of every graph
            graphs[bidx][0].replay()
        barrier()
        e0.record()
        for i in range(args.steps):
            graphs[i % NBUF][0].replay()
        e1.record()
        barrier()
        elapsed_ms = e0.elapsed_time(e1)
    t_wall1 = time.time()
    t = torch.tensor([elapsed_ms, eager_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, eager_ms = float(t[0].item()), float(t[1].item())

    # ---- e2e: host (pinned) input -> module -> loss + code map back on the host, every step --------------
    # The way a training loop feeds the module: a copy stream prefetches step i+1's batch (pinned host -> HBM, double
    # buffered) while step i computes; every step's input still crosses PCIe inside the timed region, and the host
    # reads every step's result (loss + code map) before it starts the next one.
    z_host = [torch.randn(B, D, H, H).pin_memory() for _ in range(2)]
    z_dev = [torch.empty(B, D, H, H, device=dev).requires_grad_(True) for _ in range(2)]
    ids_host = torch.empty(B, H, H, dtype=torch.int64).pin_memory()
    loss_host = torch.empty(()).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def e2e_prefetch(i):
        j = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[j])                  # the step that last used this buffer is done with it
            with torch.no_grad():
                z_dev[j].copy_(z_host[j], non_blocking=True)
            copied[j].record(copy_stream)

    def e2e_step(i):
        j = i % 2
        cur = torch.cuda.current_stream()
        cur.wait_event(copied[j])
        e2e_prefetch(i + 1)                                      # next step's H2D overlaps this step's kernels
        q, loss, ids = vq(z_dev[j])
        torch.autograd.grad((q, loss), z_dev[j], (g_q, one))
        consumed[j].record(cur)
        ids_host.copy_(ids, non_blocking=True)
        loss_host.copy_(loss, non_blocking=True)
        cur.synchronize()                                        # the caller reads the result of every step

    e2e_steps = max(1, min(args.steps, 20))
    for j in range(2):
        consumed[j].record(torch.cuda.current_stream())
    e2e_prefetch(0)
    e2e_step(0)
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(1, e2e_steps + 1):
        e2e_step(i)
    # the prefetch issued by the last step stands in for the first step's copy, which was started before the timed
    # region: wait for it, so that e2e_steps full host->device copies lie inside the region
    torch.cuda.current_stream().wait_event(copied[(e2e_steps + 1) % 2])
    s1.record()
    barrier()
    t_wall2 = time.time()
    e2e_ms = s0.elapsed_time(s1)
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    clocks = sampler.stop(t_wall0, t_wall2) if sampler else None

    # ---- eval-forward only (the reference's inference path) for the report -------------------
    vq.eval()
    with torch.no_grad():
        for i in range(2):
            vq(zbufs[i % NBUF])
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        fsteps = max(3, min(args.steps, 20))
        for i in range(fsteps):
            vq(zbufs[i % NBUF])
        f1.record()
        barrier()
        eval_ms = f0.elapsed_time(f1) / fsteps
        # rows of the last forward that the tensor-core search handed to the exhaustive fp32 fallback
        from medical_image_editing_b200.src.functions import vq_function as _vf
        fb_rows = None
        if path == 1:
            wsb = _vf._WORKSPACES.get((dev.index, torch.cuda.current_stream().cuda_stream))
            if wsb is not None:
                fb_rows = int(L.vq_debug_fallback_rows(wsb.data_ptr(), n_per_gpu, K, D, torch.cuda.current_stream().cuda_stream))
    vq.train(True)
    parity = parity_self_check(dev, L, vq, zbufs[0].detach(), world)

    # ---- the north-star point (K = 512, D = 256) on one GPU, outside the headline regions -------------------------
    north = None
    if world == 1 and wl_name == "config2" and not args.no_north_star:
        for t_ in zbufs[1:]:
            t_.grad = None
        try:
            north = measure_north_star(dev, L, pkg)
        except Exception as exc:
            north = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    # ---- the other half of BASELINE's metric: VQ-W-Net train slices/s (after, and outside, the headline regions) ----
    wnet = None
    if wl_name == "config2" and not args.no_model:
        try:
            wnet = measure_wnet_b200(dev, rank, world, steps=max(3, min(args.steps, 10)), warmup=3,
                                     inline_exchange=args.inline_exchange, fused_norm=args.fused_norm)
        except Exception as exc:                                 # never lose the headline line to the model leg
            if world > 1:
                raise
            wnet = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    if rank == 0:
        hbm, bf16, which = measured_peaks()
        total_lookups = world * n_per_gpu * args.steps
        value = total_lookups / (elapsed_ms * 1e-3)
        alg_bytes = n_per_gpu * (8 * D + 8)                       # read z + write q + int64 id
        alg_flops = 2.0 * K * D * n_per_gpu
        avg_kernel_ms = (tot_ms.value / nl.value) if nl.value else None
        roof = None
        if avg_kernel_ms:
            t_hbm = alg_bytes / (hbm * 1e9)
            t_tc = alg_flops / (bf16 * 1e12)
            if t_hbm >= t_tc:
                ach = alg_bytes / (avg_kernel_ms * 1e-3) / 1e9
                roof = {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm}
            else:
                ach = alg_flops / (avg_kernel_ms * 1e-3) / 1e12
                roof = {"bound": "tensor", "achieved": ach, "peak": bf16, "unit": "TFLOP/s", "frac": ach / bf16}
            traffic, traffic_src = load_traffic(wl_name) if path == 1 else (None, None)
            roof.update({"traffic": traffic, "traffic_source": traffic_src, "kernel": {1: "vq_assign_tc", 2: "vq_assign_small"}.get(path, "vq_assign_simt"),
                         "kernel_ms": avg_kernel_ms, "launches_timed": nl.value,
                         "algorithmic_bytes_per_launch": alg_bytes, "algorithmic_flops_per_launch": alg_flops,
                         "peak_source": f"MEASURED_PEAKS.json ({which})"})
        cpu = run_cpu_baseline(wl) if (world == 1 and not args.no_cpu) else None
        if cpu is not None:
            cpu["reference_on_gpu"] = run_reference_on_gpu(wl, dev, vq)
        if isinstance(wnet, dict) and "error" not in wnet and world == 1 and not args.no_cpu:
            try:
                wnet["cpu_baseline"] = run_cpu_wnet_baseline(budget_s=12.0)
            except Exception as exc:
                wnet["cpu_baseline"] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(wl, wl_name, world, extra={"search_path": {1: "tcgen05+fp32-rerank", 2: "fp32-small-codebook"}.get(path, "fp32-cuda-core"),
                                                          "ema_state": "cold" if args.cold else "warmed",
                                                          "stats_exchange": ("none" if world == 1 else "inline" if args.inline_exchange
                                                                             else "packed all-reduce + EMA on a side stream"),
                                                          "launch": ("CUDA-graph replay (one graph per input buffer: whole step incl. the NCCL exchange); "
                                                                     "roofline kernel_ms from the eager loop timed right before (`eager`)"
                                                                     if use_graphs else (graph_note or "eager loop (--no-graphs)"))}),
            "e2e": {"value": world * n_per_gpu * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": z_host[0].numel() * 4, "d2h_bytes_per_step": ids_host.numel() * 8 + 4,
                    "returned": "code map (int64) + loss; quantized and the input gradient stay on the device",
                    "overlap": "step i+1 H2D (copy stream, double buffer) overlaps step i kernels; result read every step",
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "h2d_GBps_per_gpu": z_host[0].numel() * 4 / (e2e_ms / e2e_steps * 1e-3) / 1e9,
                    "h2d_ceiling": "measured on this pool's 8-GPU box (profiles/r02_h2d_concurrent_8gpu.json, tools/h2d_concurrent.py): "
                                   "55.5 GB/s per GPU copying alone, 23.7 (GPUs 0-3) / 35.9 (GPUs 4-7) GB/s per GPU with all eight "
                                   "copying at once (238 GB/s in total): the N = 8 e2e figure is bound by the host's PCIe fabric"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "eval_forward": {"value": world * n_per_gpu / (eval_ms * 1e-3), "unit": UNIT, "ms_per_step": eval_ms,
                             "fallback_rows": fb_rows},
            "eager": {"value": world * n_per_gpu * args.steps / (eager_ms * 1e-3), "unit": UNIT, "ms_per_step": eager_ms / args.steps},
            "parity_check": parity,
            "north_star": north,
            "vqwnet_train": wnet,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        if not args.no_graphs:                                  # also when the capture was attempted and abandoned
            # NCCL kernels captured into CUDA graphs: tearing the process group down with the graphs around hung on the
            # GPU box (after the line had been printed).  Make sure every rank is done, then leave without the teardown.
            barrier()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# VQ-W-Net training step (BASELINE config 2 / 5): the hot path inside its caller
# ---------------------------------------------------------------------------------------------
WNET = dict(B=16, H=256, K=512, D=64, lr=1e-4, commit_weight=1.0)
WNET_METRIC = "vqwnet_train_slices_per_s"
WNET_UNIT = "slices/s"


WNET512_GLOBAL = 128          # BASELINE config 5: batch 128 of 512 x 512 slices over the GPUs of one box
WNET512_MICRO = 16            # slices per forward (one micro-batch); 16 x 512^2 through the full-resolution double U-Net


def wnet_shape(name, world):
    """(slices per GPU per step, resolution, micro-batches per step)"""
    if name == "vqwnet512":
        per_gpu = WNET512_GLOBAL // world
        return per_gpu, 512, max(1, per_gpu // WNET512_MICRO)
    return WNET["B"], WNET["H"], 1


def wnet_config(gpus, extra=None, name="vqwnet"):
    B, H, micro = wnet_shape(name, gpus)
    c = {"workload": f"{name}: VQ-W-Net training step (tools/wnet.py = vqwnet.py topology, filters 64..1024, K = {WNET['K']}; "
                     f"loss = mse(recon, x) + commit_loss, Adam lr 1e-4), batch {B} of 1x{H}x{H} "
                     "synthetic slices per GPU, fp32 (stock cuDNN convolutions, TF32 allowed as torch defaults; cudnn.benchmark per the cudnn_benchmark key; Adam fused unless VQ_TRAINER_FUSED_ADAM=0)"
                     + (f"; global batch {WNET512_GLOBAL} (run_vqwnet.py:112-121), {micro} micro-batch(es) of {B // micro} slices per "
                        "step with gradients and EMA statistics accumulated, exchanged once" if name == "vqwnet512" else ""),
         "slices_per_gpu": B, "resolution": H, "dict_size": WNET["K"], "emb_dim": WNET["D"], "micro_batches": micro,
         "parallelism": f"dp{gpus}", "l2": "a step streams > 10 GB of activations (>> 126 MB L2)"}
    if extra:
        c.update(extra)
    return c


def wnet_images(n, B, H, seed, pin=False):
    g = torch.Generator().manual_seed(seed)
    out = [torch.randn(B, 1, H, H, generator=g).clamp_(-1, 1) for _ in range(n)]
    return [t.pin_memory() for t in out] if pin else out


def fuse_wnet_norms(model, mode):
    """--fused-norm: swap InstanceNorm2d + ReLU pairs of the harness for the fused CUDA pair:
    'tail' = the pair that produces the quantiser's input (end of the first U-Net's last up stage), 'all' = every pair."""
    from medical_image_editing_b200.src.functions import fuse_norm_relu_pairs
    if mode == "tail":
        return fuse_norm_relu_pairs(model.first.up[-1].body, only_last=True)
    if mode == "all":
        return sum(fuse_norm_relu_pairs(m) for m in [m for m in model.modules() if isinstance(m, torch.nn.Sequential)])
    return 0


CUDNN_BENCHMARK = int(os.environ.get("VQ_BENCH_CUDNN_BENCHMARK", "1"))


def measure_wnet_b200(dev, rank, world, steps, warmup, inline_exchange=False, name="vqwnet", fused_norm="none"):
    """VQ-W-Net train slices/s on this rank's GPU (data parallel over `world` ranks).  Returns a dict on rank 0."""
    import torch.distributed as dist
    import medical_image_editing_b200 as pkg
    from medical_image_editing_b200.src.trainers.ddp import DataParallelVQTrainer
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from wnet import WNetHarness

    L = pkg.lib()
    B, H, micro = wnet_shape(name, world)
    # the convolutions around the quantiser are stock cuDNN (out of scope); the only knob the harness
    # touches is cuDNN's own algorithm search (same fp32 / TF32 arithmetic, the fastest algorithm per shape)
    torch.backends.cudnn.benchmark = bool(CUDNN_BENCHMARK)
    torch.manual_seed(0)
    model = WNetHarness(lambda d, k: pkg.VQ(emb_dim=d, dict_size=k, momentum=0.99, eps=1e-5, knn_backend="torch",
                                            reduce_mode="sum", overlap_exchange=(world > 1 and not inline_exchange)),
                        1, dict_size=WNET["K"]).to(dev)
    n_fused = fuse_wnet_norms(model, fused_norm)
    trainer = DataParallelVQTrainer(model, lr=WNET["lr"], commit_weight=WNET["commit_weight"])
    host = wnet_images(4, B, H, 4321 + rank, pin=True)
    resident = [t.to(dev) for t in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxed(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for i in range(warmup):
        trainer.training_step(resident[i % 4], micro_batches=micro)
    barrier()
    L.vq_profile_enable(1)
    L.vq_profile_read(None, None)
    launches0 = L.vq_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        trainer.training_step(resident[i % 4], micro_batches=micro)
    e1.record()
    barrier()
    ms = maxed(e0.elapsed_time(e1))
    launches = L.vq_launch_count() - launches0
    tot_ms, nl = ctypes.c_double(0), ctypes.c_int(0)
    L.vq_profile_read(ctypes.byref(tot_ms), ctypes.byref(nl))
    L.vq_profile_enable(0)

    # end to end: each step's images come from pinned host memory, the loss is read back by the host
    dev_in = torch.empty(B, 1, H, H, device=dev)
    losses = []
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(steps):
        dev_in.copy_(host[i % 4], non_blocking=True)
        out = trainer.training_step(dev_in, micro_batches=micro)
        losses.append(float(out["loss"].item()))
    s1.record()
    barrier()
    e2e_ms = maxed(s0.elapsed_time(s1))
    in_sync = trainer.replicas_in_sync()
    del trainer, model
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    hbm, _, which = measured_peaks()
    n_vec = (B // micro) * H * H                             # vectors per quantiser call
    alg_bytes = n_vec * (8 * WNET["D"] + 8)
    k_ms = (tot_ms.value / nl.value) if nl.value else None
    return {
        "metric": WNET_METRIC, "value": world * B * steps / (ms * 1e-3), "unit": WNET_UNIT, "ms_per_step": ms / steps,
        "steps": steps, "warmup": warmup, "n_gpus": world, "slices_per_gpu": B, "resolution": H, "micro_batches": micro,
        "e2e": {"value": world * B * steps / (e2e_ms * 1e-3), "unit": WNET_UNIT, "h2d_bytes_per_step": B * H * H * 4,
                "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / steps},
        "gpu_launches": int(launches),
        "quantiser_search_kernel_ms": k_ms,
        "roofline": (None if not k_ms else {"bound": "hbm", "achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                            "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm, "traffic": None, "kernel": "vq_assign_tc",
                                            "peak_source": f"MEASURED_PEAKS.json ({which})"}),
        "replicas_in_sync": bool(in_sync), "final_loss": losses[-1] if losses else None,
        "fused_norm": fused_norm, "fused_norm_pairs": n_fused, "cudnn_benchmark": bool(CUDNN_BENCHMARK),
    }


def cpu_wnet_step_factory(slices, H=None):
    """One Adam step of VQ-W-Net on the host cores: the UNMODIFIED reference `VQWNet` (vqwnet.py:13-152, with its own
    `VQModule`) when a copy of the reference is reachable, else tools/wnet.py around the oracle quantiser."""
    H = H or WNET["H"]
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    model, kind = None, "port"
    try:
        from oracle import ref_loader
        if ref_loader.reference_available():
            model = ref_loader.load_reference_net("vqwnet").VQWNet(1, 1, dict_size=WNET["K"])
            kind = "reference"
    except Exception:
        model = None
    if model is None:
        from oracle.vq_oracle import OracleVQ
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from wnet import WNetHarness
        model = WNetHarness(lambda d, k: OracleVQ(d, k, 0.99, 1e-5, "torch", chunk=65536), 1, dict_size=WNET["K"])
    model.train(True)
    opt = torch.optim.Adam(model.parameters(), lr=WNET["lr"])
    imgs = wnet_images(2, slices, H, 4321)
    state = {"i": 0}

    def step():
        x = imgs[state["i"] % 2]
        state["i"] += 1
        opt.zero_grad(set_to_none=True)
        out = model(x)
        loss = torch.nn.functional.mse_loss(out["recon"], x) + WNET["commit_weight"] * out["commit_loss"]
        loss.backward()
        opt.step()
        return loss

    step.kind = kind
    return step


def run_cpu_wnet_baseline(budget_s=15.0, slices=1, H=None):
    step = cpu_wnet_step_factory(slices, H)
    step()
    times = []
    t_all = time.perf_counter()
    while len(times) < 2 or (time.perf_counter() - t_all < budget_s and len(times) < 20):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return {"value": slices / med, "unit": WNET_UNIT, "cores": torch.get_num_threads(), "kind": step.kind,
            "sample": f"{slices} slice(s) of {H or WNET['H']}^2 per step: "
                      + ("the unmodified reference VQWNet (baseline/_ref mirror) on the host cores" if step.kind == "reference" else
                         "tools/wnet.py with the CPU oracle quantiser (bit-identical to the reference VQWNet on CPU, tests/test_wnet_harness.py)")
                      + f", median of {len(times)} steps, {med * 1e3:.0f} ms/step"}


def run_wnet_reference_arm(args):
    if env_int("RANK", 0) != 0:
        return
    os.environ["WORLD_SIZE"] = "1"
    slices = 1
    _, H, _ = wnet_shape(args.workload, max(1, args.gpus))
    step = cpu_wnet_step_factory(slices, H)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = slices * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": WNET_METRIC, "value": val, "unit": WNET_UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.workload == "vqwnet512" else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": wnet_config(args.gpus, extra={"reference_sample_slices_per_step": slices}, name=args.workload),
        "cpu_baseline": {"value": val, "unit": WNET_UNIT, "cores": torch.get_num_threads(), "kind": step.kind,
                         "sample": f"{slices} slice of {H}^2 per step; " + ("the unmodified reference VQWNet" if step.kind == "reference"
                                                                          else "VQ-W-Net topology with the oracle quantiser") + " on the host cores"},
        "e2e": {"value": val, "unit": WNET_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def run_wnet_b200_arm(args):
    import torch.distributed as dist
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the quantiser has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sampler = ClockSampler(local) if rank == 0 else None
    t0 = time.time()
    r = measure_wnet_b200(dev, rank, world, args.steps, args.warmup, args.inline_exchange, name=args.workload,
                          fused_norm=args.fused_norm)
    t1 = time.time()
    if rank == 0:
        clocks = sampler.stop(t0, t1)
        _, H, _ = wnet_shape(args.workload, world)
        cpu = run_cpu_wnet_baseline(H=H) if (world == 1 and not args.no_cpu) else None
        out = {"metric": WNET_METRIC, "value": r["value"], "unit": WNET_UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
               "scaling": "strong" if args.workload == "vqwnet512" else "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": wnet_config(world, name=args.workload),
               "e2e": r["e2e"], "gpu_launches": r["gpu_launches"], "clocks": clocks, "roofline": r["roofline"],
               "cpu_baseline": cpu, "replicas_in_sync": r["replicas_in_sync"], "final_loss": r["final_loss"],
               "fused_norm": r["fused_norm"], "fused_norm_pairs": r["fused_norm_pairs"],
               "cudnn_benchmark": r["cudnn_benchmark"]}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS) + ["vqwnet", "vqwnet512"])
    ap.add_argument("--no-model", action="store_true", help="skip the VQ-W-Net train slices/s block of the default line")
    ap.add_argument("--no-north-star", action="store_true", help="skip the K = 512, D = 256 block of the default line")
    ap.add_argument("--simt", action="store_true", help="force the fp32 CUDA-core search")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cold", action="store_true", help="start from the first-step EMA state (cluster_size = 0)")
    ap.add_argument("--inline-exchange", action="store_true", help="N > 1: all-reduce + EMA update on the compute stream")
    ap.add_argument("--fused-norm", default="all", choices=["none", "tail", "all"],
                    help="VQ-W-Net legs: InstanceNorm2d + ReLU pairs run by vq_norm_relu_fwd/bwd (tail = the quantiser's producer)")
    ap.add_argument("--no-graphs", action="store_true", help="eager step loop in the timed region instead of CUDA-graph replay")
    ap.add_argument("--cudnn-benchmark", type=int, default=CUDNN_BENCHMARK, choices=[0, 1],
                    help="VQ-W-Net legs: torch.backends.cudnn.benchmark for the stock convolutions around the quantiser")
    args = ap.parse_args()
    globals()["CUDNN_BENCHMARK"] = args.cudnn_benchmark
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.workload in ("vqwnet", "vqwnet512"):
        return run_wnet_reference_arm(args) if args.impl == "reference" else run_wnet_b200_arm(args)
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, wl, args.workload)
    else:
        run_b200_arm(args, wl, args.workload)


if __name__ == "__main__":
    main()
