#!/usr/bin/env bash
# N GPUs (argument): BASELINE config 5 (global batch 128 of 512^2 slices, data parallel, micro-batched)
N=${1:-2}
O=gpurun_out/r02c35
mkdir -p $O
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --workload vqwnet512 --gpus $N --steps 3 --warmup 3 --no-cpu > $O/wnet512_n$N.log 2> $O/wnet512_n$N.err; echo "wnet512 N=$N rc $?"; tail -c 900 $O/wnet512_n$N.log; tail -3 $O/wnet512_n$N.err
