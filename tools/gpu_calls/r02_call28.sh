#!/usr/bin/env bash
# 2-GPU bench line after the pending-event fix (graph replay, then eager end-to-end steps)
mkdir -p gpurun_out/r02c28
O=gpurun_out/r02c28
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --no-model > $O/bench_n2.log 2> $O/bench_n2.err; echo "bench rc $?"; tail -c 2500 $O/bench_n2.log; grep -v "^$" $O/bench_n2.err | grep -i "error" | head -5
