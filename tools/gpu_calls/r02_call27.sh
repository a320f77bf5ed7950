#!/usr/bin/env bash
# two-rank NCCL parity tests on real hardware + a 2-GPU bench line
mkdir -p gpurun_out/r02c27
O=gpurun_out/r02c27
nvidia-smi -L > $O/gpus.log 2>&1
timeout 600 python -m pytest tests -m gpu -v --timeout 300 -k "two_rank or two_ranks" > $O/two_rank_tests.log 2>&1; echo "pytest rc $?"; tail -8 $O/two_rank_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --no-model > $O/bench_n2.log 2> $O/bench_n2.err; echo "bench rc $?"; tail -c 3000 $O/bench_n2.log; tail -3 $O/bench_n2.err
