#!/usr/bin/env bash
# 8-GPU box: concurrent H2D ceiling, then the quantiser bench line at N = 8
mkdir -p gpurun_out/r02c33
O=gpurun_out/r02c33
nvidia-smi topo -m > $O/topo.log 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 tools/h2d_concurrent.py > $O/h2d8.log 2> $O/h2d8.err; echo "h2d rc $?"; tail -2 $O/h2d8.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 8 --steps 20 --warmup 3 --no-model > $O/bench_n8.log 2> $O/bench_n8.err; echo "bench rc $?"; tail -c 1800 $O/bench_n8.log
