#!/usr/bin/env bash
# eight GPUs: the default bench line exactly as the driver launches it (final code of the round)
O=gpurun_out/r02c42
mkdir -p $O
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 20 --warmup 3 > $O/bench_n8.log 2> $O/bench_n8.err; echo "bench rc $?"; tail -c 1200 $O/bench_n8.log; tail -3 $O/bench_n8.err
