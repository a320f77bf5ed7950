#!/usr/bin/env bash
# default bench line (N = 1), reference arm, launch list and full ncu capture of the dominant kernel
mkdir -p gpurun_out/r02c29
O=gpurun_out/r02c29
timeout 700 python bench.py > $O/bench.log 2> $O/bench.err; echo "bench rc $?"; tail -c 7000 $O/bench.log; tail -3 $O/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.log 2> $O/bench_ref.err; echo "ref rc $?"; tail -c 1500 $O/bench_ref.log
CMD="python bench.py --steps 2 --warmup 3 --no-model --no-north-star --no-cpu --no-graphs"
timeout 200 $CMD > $O/plain.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu_l.log 2>&1
echo "ncu launches rc $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:vq_assign_tc_kernel -s 4 -c 2 -o $O/prof_r02c $CMD > $O/ncu_f.log 2>&1
echo "ncu full rc $?"; ls -la $O
