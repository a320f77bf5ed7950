#!/usr/bin/env bash
# round 2, call 1: parity of the new scan / q-through-TMA resident kernel + A/B against the round-1 library
mkdir -p gpurun_out
O=gpurun_out/r02c1
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
ab() { lib=$1; shift; if [ "$lib" = default ]; then timeout 200 python tools/ab.py "$@"; else VQ_B200_LIB=build_variants/lib_$lib.so timeout 200 python tools/ab.py "$@"; fi; }
{
for lib in r01 default q0 pf0 q0pf0; do ab $lib 64 512 16 noise; done
for lib in r01 default q0 pf0; do ab $lib 64 512 16 clustered; done
for lib in r01 default; do ab $lib 64 512 16 relu; done
for lib in r01 default; do ab $lib 256 512 16 clustered; done
for lib in r01 default; do ab $lib 256 512 16 noise; done
for lib in r01 default q0; do ab $lib 64 64 16 noise; done
for lib in r01 default; do ab $lib 128 512 16 clustered; done
for lib in r01 default; do ab $lib 64 4096 16 clustered; done
} > $O/ab.log 2>&1
{
VQ_B200_LIB=build_variants/lib_timing.so timeout 200 python tools/tc_timing.py 64 512 16 1 noise
VQ_B200_LIB=build_variants/lib_timing.so timeout 200 python tools/tc_timing.py 64 512 16 0 noise
VQ_B200_LIB=build_variants/lib_timing_q0.so timeout 200 python tools/tc_timing.py 64 512 16 1 noise
VQ_B200_LIB=build_variants/lib_timing_q0.so timeout 200 python tools/tc_timing.py 64 512 16 0 noise
} > $O/timing.log 2>&1
tail -5 $O/pytest.log; cat $O/ab.log; cat $O/timing.log
