#!/usr/bin/env bash
mkdir -p gpurun_out/r02c10
O=gpurun_out/r02c10
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -20 $O/smoke.log; exit 1; }
{
timeout 100 python tools/ab.py 64 512 16 noise
timeout 100 python tools/ab.py 64 512 16 clustered
timeout 100 python tools/ab.py 64 512 16 relu
} > $O/ab.log 2>&1
cat $O/ab.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:vq_assign_r3_kernel -s 12 -c 2 -o $O/prof_r3_noise python tools/ab.py 64 512 16 noise > $O/ncu_f.log 2>&1
echo "ncu full rc $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:vq_assign_r3_kernel -s 13 -c 1 -o $O/prof_r3_clustered python tools/ab.py 64 512 16 clustered > $O/ncu_f2.log 2>&1
echo "ncu full rc $?"; ls -la $O
