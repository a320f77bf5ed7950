"""CPU: the reference arm of bench.py (the reference's own module when a copy of the reference is reachable, else the
CPU port of its op chain) prints its JSON line, also when it is launched the way the driver launches the N > 1 arms
(torchrun sets WORLD_SIZE / RANK; ranks > 0 exit without work)."""
import json
import os
import subprocess
import sys

from util import ROOT
from oracle import ref_loader

KIND = "reference" if ref_loader.reference_available() else "port"


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
    assert d["cpu_baseline"]["kind"] == KIND and d["e2e"]["h2d_bytes_per_step"] == 0


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
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == KIND and d["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_arm_falls_back_to_the_port_without_a_reference_copy():
    """With no copy of the reference reachable ($VQ_REF_SRC pointing nowhere does not help: the loader also looks at
    /root/reference and the mirror), the port must still print the line: force it by hiding both locations."""
    code = ("import sys, json; sys.argv=['bench.py','--impl','reference','--steps','1','--warmup','1'];"
            "from oracle import ref_loader; ref_loader.REF_SRC_CANDIDATES[:] = ['/nonexistent'];"
            "import runpy; runpy.run_path('bench.py', run_name='__main__')")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=600,
                       env=dict(os.environ, VQ_REF_BUDGET_S="5"))
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["kind"] == "port" and d["value"] > 0
