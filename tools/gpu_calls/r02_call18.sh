#!/usr/bin/env bash
mkdir -p gpurun_out/r02c18
O=gpurun_out/r02c18
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -20 $O/smoke.log; exit 1; }
{
timeout 120 python tools/r3_check.py relu 4 1
timeout 120 python tools/r3_check.py noise 3 1
} > $O/check.log 2>&1
cut -c1-60,100-260 $O/check.log
{
timeout 100 python tools/ab.py 64 512 16 noise
timeout 100 python tools/ab.py 64 512 16 clustered
timeout 100 python tools/ab.py 64 512 16 relu
timeout 100 python tools/ab.py 64 64 16 noise
} > $O/ab.log 2>&1
cat $O/ab.log
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q --timeout 150 > $O/pytest.log 2>&1; rc=$?; tail -3 $O/pytest.log
{
VQ_B200_LIB=build_variants/lib_r3trace.so timeout 100 python tools/r3_trace.py 64 512 16 0 noise
VQ_B200_LIB=build_variants/lib_r3trace.so timeout 100 python tools/r3_trace.py 64 512 16 1 noise
} > $O/trace.log 2>&1
grep -v "^  tile\|^tiles 20" $O/trace.log
