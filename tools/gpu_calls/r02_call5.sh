#!/usr/bin/env bash
mkdir -p gpurun_out/r02c5
O=gpurun_out/r02c5
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -20 $O/smoke.log; exit 1; }
timeout 120 python -m pytest tests -m gpu -x -q --timeout 100 -k "lookup or natural or micro_batch" > $O/pytest_sub.log 2>&1; tail -3 $O/pytest_sub.log
timeout 500 python bench.py > $O/bench.log 2> $O/bench.err; echo "bench rc $?"; tail -c 6000 $O/bench.log; tail -5 $O/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.log 2> $O/bench_ref.err; echo "ref rc $?"; cat $O/bench_ref.log
CMD="python bench.py --steps 2 --warmup 3 --no-model --no-north-star --no-cpu --no-graphs"
timeout 200 $CMD > $O/plain.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu_l.log 2>&1
echo "ncu launches rc $?"
timeout 200 $CMD > $O/plain2.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:vq_assign_tc -s 4 -c 2 -o $O/prof_r02a $CMD > $O/ncu_f.log 2>&1
echo "ncu full rc $?"; ls -la $O
