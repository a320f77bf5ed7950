#!/usr/bin/env bash
mkdir -p gpurun_out/r02c17
O=gpurun_out/r02c17
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -20 $O/smoke.log; exit 1; }
{
timeout 100 python tools/ab.py 64 512 16 noise
timeout 100 python tools/ab.py 64 512 16 clustered
VQ_B200_LIB=build_variants/lib_r3sleep256.so timeout 100 python tools/ab.py 64 512 16 noise
VQ_B200_LIB=build_variants/lib_r3sleep256.so timeout 100 python tools/ab.py 64 512 16 clustered
} > $O/ab.log 2>&1
cat $O/ab.log
