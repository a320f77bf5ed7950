#!/usr/bin/env bash
mkdir -p gpurun_out/r02c25
O=gpurun_out/r02c25
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -20 $O/smoke.log; exit 1; }
{
timeout 100 python tools/ab.py 64 512 16 noise
timeout 100 python tools/ab.py 64 512 16 clustered
timeout 100 python tools/ab.py 64 512 16 relu
VQ_B200_R3=1 timeout 100 python tools/ab.py 64 512 16 clustered
} > $O/ab.log 2>&1
cat $O/ab.log
timeout 600 python -m pytest tests -m gpu -x -q --timeout 100 > $O/pytest.log 2>&1; rc=$?; tail -3 $O/pytest.log
