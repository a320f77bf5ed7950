#!/usr/bin/env bash
# one GPU: ncu --set full of the fused norm + relu forward / backward (config-2 planes)
O=gpurun_out/r02c43
mkdir -p $O
timeout 60 python tools/norm_relu_once.py > $O/once.log 2>&1; echo "plain run rc $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:vq_norm_relu -s 4 -c 2 -o $O/norm_relu_256 -f python tools/norm_relu_once.py > $O/ncu_full.log 2>&1; echo "ncu rc $?"; ls -la $O
