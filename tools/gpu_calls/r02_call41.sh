#!/usr/bin/env bash
# one GPU: measured DRAM traffic of the streamed search kernel at the north-star point (K = 512, D = 256), three input kinds
O=gpurun_out/r02c41
mkdir -p $O
for kind in clustered relu noise; do
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:vq_assign_tcs -c 40 --csv --log-file $O/ncu_k512d256_$kind.csv python tools/ab.py 256 512 16 $kind > $O/ncu_$kind.log 2>&1; echo "ncu $kind rc $?"
done
