#!/usr/bin/env bash
# N GPUs (argument): BASELINE config 5 with the final code (fused norms on by default)
N=${1:-1}
O=gpurun_out/r02c46
mkdir -p $O
if [ "$N" = "1" ]; then
  timeout 500 python bench.py --workload vqwnet512 --gpus 1 --steps 2 --warmup 3 > $O/wnet512_n1.log 2> $O/wnet512_n1.err; echo "wnet512 N=1 rc $?"
else
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 bench.py --workload vqwnet512 --gpus $N --steps 3 --warmup 3 --no-cpu > $O/wnet512_n$N.log 2> $O/wnet512_n$N.err; echo "wnet512 N=$N rc $?"
fi
python - $O/wnet512_n$N.log <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("N", d["n_gpus"], "slices/s %.1f ms/step %.1f micro %d fused %s in_sync %s cpu %s" % (d["value"], d["ms_per_step"], d["config"]["micro_batches"], d.get("fused_norm"), d.get("replicas_in_sync"), (d.get("cpu_baseline") or {}).get("value")))
PY
