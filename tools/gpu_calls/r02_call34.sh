#!/usr/bin/env bash
# one GPU: BASELINE config 5 at N = 1 (128 slices of 512^2, 8 micro-batches), then the config-3 sweep + run_recon lines
O=gpurun_out/r02c34
mkdir -p $O
timeout 400 python bench.py --workload vqwnet512 --gpus 1 --steps 2 --warmup 3 --no-cpu > $O/wnet512_n1.log 2> $O/wnet512_n1.err; echo "wnet512 rc $?"; tail -c 600 $O/wnet512_n1.log; tail -3 $O/wnet512_n1.err
nvidia-smi --query-gpu=memory.used --format=csv,noheader
: > $O/sweep.jsonl
for w in k64d64 k64d256 k512d256 k4096d64 k4096d256 recon_k10d16; do
  timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-model --no-north-star >> $O/sweep.jsonl 2> $O/sweep_$w.err; echo "$w rc $?"
done
python - <<'PY'
import json
for l in open("gpurun_out/r02c34/sweep.jsonl"):
    try: d = json.loads(l)
    except Exception: continue
    r = d.get("roofline") or {}
    print(d["config"]["workload"][:40], "ms/step %.3f" % d["ms_per_step"], "kernel_ms", r.get("kernel_ms"), "frac", r.get("frac"), "bound", r.get("bound"), "eval_ms", d["eval_forward"]["ms_per_step"])
PY
