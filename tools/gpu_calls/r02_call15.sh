#!/usr/bin/env bash
mkdir -p gpurun_out/r02c15
O=gpurun_out/r02c15
{
timeout 100 python tools/ab.py 64 512 16 noise
timeout 100 python tools/ab.py 64 512 16 relu
} > $O/ab.log 2>&1
cat $O/ab.log
{
VQ_B200_LIB=build_variants/lib_r3trace.so timeout 100 python tools/r3_trace.py 64 512 16 0 noise
VQ_B200_LIB=build_variants/lib_r3trace.so timeout 100 python tools/r3_trace.py 64 512 16 1 relu
} > $O/trace.log 2>&1
grep -v "^  tile\|^tiles 20" $O/trace.log
