#!/usr/bin/env bash
mkdir -p gpurun_out/r02c21
O=gpurun_out/r02c21
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -20 $O/smoke.log; exit 1; }
timeout 60 python tools/r3_check.py noise 2 1 > $O/check0.log 2>&1 || { echo "CHECK FAILED/HUNG"; tail -5 $O/check0.log | cut -c1-300; exit 1; }
cut -c1-60,100-200 $O/check0.log
{
timeout 100 python tools/ab.py 64 512 16 noise
timeout 100 python tools/ab.py 64 512 16 clustered
timeout 100 python tools/ab.py 64 512 16 relu
} > $O/ab.log 2>&1
cat $O/ab.log
