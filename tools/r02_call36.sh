#!/usr/bin/env bash
# one GPU: full GPU suite (streamed kernel now takes emb_dim 512), VQGAN shape (K = 64, D = 512) tensor-core vs CUDA-core search
O=gpurun_out/r02c36
mkdir -p $O
timeout 600 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 $O/pytest_gpu.log
for data in noise clustered; do
  timeout 120 python tools/ab.py 512 64 4 $data >> $O/ab_d512.log 2>&1
  ABSIMT=1 timeout 120 python tools/ab.py 512 64 4 $data >> $O/ab_d512.log 2>&1
done
cat $O/ab_d512.log
