#!/usr/bin/env bash
mkdir -p gpurun_out/r02c11
O=gpurun_out/r02c11
{
VQ_B200_LIB=build_variants/lib_r3trace.so timeout 100 python tools/r3_trace.py 64 512 16 0 clustered
VQ_B200_LIB=build_variants/lib_r3trace.so timeout 100 python tools/r3_trace.py 64 512 16 0 noise
VQ_B200_LIB=build_variants/lib_r3trace.so timeout 100 python tools/r3_trace.py 64 512 16 1 noise
} > $O/trace.log 2>&1
cat $O/trace.log
