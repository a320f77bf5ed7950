#!/usr/bin/env bash
# fail fast: a hung kernel must not eat the GPU budget
mkdir -p gpurun_out/r02c3
O=gpurun_out/r02c3
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -20 $O/smoke.log; exit 1; }
timeout 100 python tools/ab.py 64 512 16 noise > $O/ab_first.log 2>&1 || { echo "AB FAILED/HUNG"; tail -5 $O/ab_first.log; exit 1; }
cat $O/ab_first.log
timeout 600 python -m pytest tests -m gpu -x -q --timeout 150 > $O/pytest.log 2>&1; rc=$?; echo "pytest rc $rc" >> $O/pytest.log
tail -4 $O/pytest.log
[ $rc -ne 0 ] && { grep -E "FAILED|Error|Timeout" $O/pytest.log | head -20; }
ab() { lib=$1; shift; if [ "$lib" = default ]; then timeout 100 python tools/ab.py "$@"; else VQ_B200_LIB=build_variants/lib_$lib.so timeout 100 python tools/ab.py "$@"; fi; }
{
for lib in r01 default zldg0 coop0 pf0 abl_norr; do ab $lib 64 512 16 noise; done
for lib in r01 default zldg0 coop0; do ab $lib 64 512 16 clustered; done
for lib in r01 default; do ab $lib 64 512 16 relu; done
for lib in r01 default; do ab $lib 64 64 16 noise; done
} > $O/ab.log 2>&1
cat $O/ab.log
{
VQ_B200_LIB=build_variants/lib_timing.so timeout 100 python tools/tc_timing.py 64 512 16 1 noise
VQ_B200_LIB=build_variants/lib_timing.so timeout 100 python tools/tc_timing.py 64 512 16 0 noise
VQ_B200_LIB=build_variants/lib_timing.so timeout 100 python tools/tc_timing.py 64 512 16 0 clustered
} > $O/timing.log 2>&1
{
VQ_B200_LIB=build_variants/lib_trace.so timeout 100 python tools/tc_trace.py 64 512 16 0 noise
VQ_B200_LIB=build_variants/lib_trace.so timeout 100 python tools/tc_trace.py 64 512 16 0 clustered
VQ_B200_LIB=build_variants/lib_trace_pf0.so timeout 100 python tools/tc_trace.py 64 512 16 0 clustered
VQ_B200_LIB=build_variants/lib_trace.so timeout 100 python tools/tc_trace.py 64 512 16 1 noise
} > $O/trace.log 2>&1
cat $O/timing.log; cat $O/trace.log
