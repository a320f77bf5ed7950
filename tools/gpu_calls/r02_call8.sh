#!/usr/bin/env bash
# baseline of the session: smoke, A/B timing of the default build, full ncu capture (with source) of one training and one eval launch
mkdir -p gpurun_out/r02c8
O=gpurun_out/r02c8
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -20 $O/smoke.log; exit 1; }
{
timeout 100 python tools/ab.py 64 512 16 noise
timeout 100 python tools/ab.py 64 512 16 clustered
timeout 100 python tools/ab.py 256 512 16 noise
timeout 100 python tools/ab.py 256 512 16 clustered
} > $O/ab.log 2>&1
cat $O/ab.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:vq_assign_tc_kernel -s 12 -c 2 -o $O/prof_r02b python tools/ab.py 64 512 16 noise > $O/ncu_f.log 2>&1
echo "ncu full rc $?"; ls -la $O
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv
