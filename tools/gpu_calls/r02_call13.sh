#!/usr/bin/env bash
mkdir -p gpurun_out/r02c13
O=gpurun_out/r02c13
{
timeout 120 python tools/r3_check.py relu 4 1
VQ_B200_LIB=build_variants/lib_r3zc.so timeout 120 python tools/r3_check.py relu 4 1
timeout 120 python tools/r3_check.py relu 4 1
VQ_B200_LIB=build_variants/lib_r3zc.so timeout 120 python tools/r3_check.py relu 4 1
} > $O/check.log 2>&1
cut -c1-60,100-260 $O/check.log
