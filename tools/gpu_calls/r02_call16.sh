#!/usr/bin/env bash
mkdir -p gpurun_out/r02c16
O=gpurun_out/r02c16
timeout 400 ncu --set full --clock-control none --import-source on -k regex:vq_assign_r3_kernel -s 12 -c 2 -o $O/prof_r3b_noise python tools/ab.py 64 512 16 noise > $O/ncu_f.log 2>&1
echo "ncu full rc $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:vq_assign_r3_kernel -s 13 -c 1 -o $O/prof_r3b_clustered python tools/ab.py 64 512 16 clustered > $O/ncu_f2.log 2>&1
echo "ncu full rc $?"
