#!/usr/bin/env bash
# third-generation resident kernel: smoke, A/B against the second generation, GPU test suite
mkdir -p gpurun_out/r02c9
O=gpurun_out/r02c9
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -20 $O/smoke.log; exit 1; }
{
timeout 100 python tools/ab.py 64 512 16 noise
VQ_B200_R3=0 timeout 100 python tools/ab.py 64 512 16 noise
timeout 100 python tools/ab.py 64 512 16 clustered
timeout 100 python tools/ab.py 64 512 16 relu
timeout 100 python tools/ab.py 64 64 16 noise
} > $O/ab.log 2>&1
cat $O/ab.log
timeout 900 python -m pytest tests -m gpu -x -q --timeout 150 > $O/pytest.log 2>&1; rc=$?; echo "pytest rc $rc" >> $O/pytest.log
tail -5 $O/pytest.log
[ $rc -ne 0 ] && { grep -E "FAILED|Error|Timeout|assert" $O/pytest.log | head -30; }
true
