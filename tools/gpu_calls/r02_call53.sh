#!/usr/bin/env bash
# one GPU: DRAM bytes and duration of the memory-bound kernels at benchmark sizes with the final launch shapes
O=gpurun_out/r02c53
mkdir -p $O
timeout 120 python tools/stream_kernels.py > $O/plain.log 2>&1; echo "plain rc $?"
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"vq_bwd_vec|vq_lookup|vq_assign_small|vq_el_" --csv --log-file $O/stream_kernels.csv python tools/stream_kernels.py > $O/ncu.log 2>&1; echo "ncu rc $?"; wc -l $O/stream_kernels.csv
