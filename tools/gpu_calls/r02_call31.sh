#!/usr/bin/env bash
mkdir -p gpurun_out/r02c31
O=gpurun_out/r02c31
timeout 200 python tools/stream_kernels.py > $O/plain.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/stream.csv python tools/stream_kernels.py > $O/ncu.log 2>&1
echo "rc $?"
