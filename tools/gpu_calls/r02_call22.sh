#!/usr/bin/env bash
mkdir -p gpurun_out/r02c22
O=gpurun_out/r02c22
timeout 400 ncu --set full --clock-control none --import-source on -k regex:vq_assign_r3_kernel -s 13 -c 1 -o $O/prof_r3c_noise python tools/ab.py 64 512 16 noise > $O/ncu_f.log 2>&1
echo "ncu full rc $?"
