#!/usr/bin/env bash
mkdir -p gpurun_out/r02c26
O=gpurun_out/r02c26
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -20 $O/smoke.log; exit 1; }
{
timeout 100 python tools/ab.py 256 512 16 noise
timeout 100 python tools/ab.py 256 512 16 clustered
timeout 100 python tools/ab.py 256 512 16 relu
timeout 100 python tools/ab.py 64 4096 16 noise
} > $O/ab.log 2>&1
cat $O/ab.log
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q --timeout 100 -k "full_size or seeded or multi_tile or micro" > $O/pytest.log 2>&1; tail -2 $O/pytest.log
