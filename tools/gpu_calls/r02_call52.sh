#!/usr/bin/env bash
# two GPUs: the default line exactly as the driver launches it at N = 2, final code (fused Adam on bucket views, cuDNN
# algorithm search per rank, backward kernels in channel slices)
O=gpurun_out/r02c52
mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_n2.log 2> $O/bench_n2.err; echo "bench rc $?"; tail -3 $O/bench_n2.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02c52/bench_n2.log") if l.startswith("{")][-1])
w = d["vqwnet_train"]
print("N", d["n_gpus"], "value %.4g ms/step %.4f parity %s | wnet %.1f slices/s in_sync %s loss %.4f" % (d["value"], d["ms_per_step"], d.get("parity_check"), w["value"], w["replicas_in_sync"], w["final_loss"]))
PY
