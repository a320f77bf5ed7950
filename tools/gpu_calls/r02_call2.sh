#!/usr/bin/env bash
mkdir -p gpurun_out/r02c2
O=gpurun_out/r02c2
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
ab() { lib=$1; shift; if [ "$lib" = default ]; then timeout 200 python tools/ab.py "$@"; else VQ_B200_LIB=build_variants/lib_$lib.so timeout 200 python tools/ab.py "$@"; fi; }
{
for lib in r01 default aux0 coop0 pf0 abl_noq abl_noout abl_norr; do ab $lib 64 512 16 noise; done
for lib in r01 default aux0 coop0; do ab $lib 64 512 16 clustered; done
for lib in r01 default; do ab $lib 64 64 16 noise; done
for lib in r01 default; do ab $lib 16 10 16 noise; done
} > $O/ab.log 2>&1
{
VQ_B200_LIB=build_variants/lib_timing.so timeout 200 python tools/tc_timing.py 64 512 16 1 noise
VQ_B200_LIB=build_variants/lib_timing.so timeout 200 python tools/tc_timing.py 64 512 16 0 noise
} > $O/timing.log 2>&1
{
VQ_B200_LIB=build_variants/lib_trace.so timeout 200 python tools/tc_trace.py 64 512 16 0 noise
VQ_B200_LIB=build_variants/lib_trace.so timeout 200 python tools/tc_trace.py 64 512 16 1 noise
} > $O/trace.log 2>&1
tail -4 $O/pytest.log; cat $O/ab.log; cat $O/timing.log; cat $O/trace.log
