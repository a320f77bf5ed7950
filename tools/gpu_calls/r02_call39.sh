#!/usr/bin/env bash
# two GPUs: the default bench line under torchrun (fused norms in the VQ-W-Net block under data parallelism), two-rank tests
O=gpurun_out/r02c39
mkdir -p $O
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_n2.log 2> $O/bench_n2.err; echo "bench rc $?"; tail -c 1500 $O/bench_n2.log; tail -3 $O/bench_n2.err
timeout 300 python -m pytest tests -q -m gpu -k "two_rank or two_ranks" > $O/pytest_two_rank.log 2>&1; echo "pytest rc $?"; tail -3 $O/pytest_two_rank.log
