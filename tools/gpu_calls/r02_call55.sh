#!/usr/bin/env bash
# one GPU: launch list of the final code (backward in four channel slices)
O=gpurun_out/r02c55
mkdir -p $O
CMD="python bench.py --steps 2 --warmup 3 --no-model --no-north-star --no-cpu --no-graphs"
timeout 100 $CMD > $O/plain.log 2>&1 && timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu_l.log 2>&1
echo "ncu launches rc $?"
