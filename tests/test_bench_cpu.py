"""CPU: the reference arm of bench.py (the CPU port of the reference's op chain) prints its JSON line, also when it is
launched the way the driver launches the N > 1 arms (torchrun sets WORLD_SIZE / RANK; ranks > 0 exit without work)."""
import json
import os
import subprocess
import sys

from util import ROOT


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus",
                        env_extra.get("WORLD_SIZE", "1"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip()


def test_reference_arm_single_process():
    d = json.loads(_run({}).splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "vq_lookups_per_s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


def test_reference_arm_under_torchrun_env():
    out0 = _run({"WORLD_SIZE": "2", "RANK": "0", "LOCAL_RANK": "0"})
    d = json.loads(out0.splitlines()[-1])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    assert _run({"WORLD_SIZE": "2", "RANK": "1", "LOCAL_RANK": "1"}) == ""       # other ranks: no work, exit 0


def test_reference_arm_vqwnet_workload():
    """`--workload vqwnet --impl reference`: the VQ-W-Net harness with the oracle quantiser on the host cores."""
    env = dict(os.environ)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "vqwnet",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "vqwnet_train_slices_per_s" and d["unit"] == "slices/s"
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["d2h_bytes_per_step"] == 0
