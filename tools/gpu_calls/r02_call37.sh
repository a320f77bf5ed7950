#!/usr/bin/env bash
# one GPU: GPU suite again (cluster plan of the fused norm + relu changed), its timing at 256^2 / 512^2 planes
O=gpurun_out/r02c37
mkdir -p $O
timeout 600 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -5 $O/pytest_gpu.log
timeout 200 python tools/norm_relu_bench.py 16 64 256 > $O/norm_relu_256.json 2> $O/norm_relu.err; cat $O/norm_relu_256.json
timeout 200 python tools/norm_relu_bench.py 16 64 512 > $O/norm_relu_512.json 2>> $O/norm_relu.err; cat $O/norm_relu_512.json
timeout 200 python tools/norm_relu_bench.py 16 1024 16 > $O/norm_relu_16.json 2>> $O/norm_relu.err; cat $O/norm_relu_16.json
tail -3 $O/norm_relu.err
