#!/usr/bin/env bash
# one GPU: norm + relu tests and timing (streaming g_x stores), ncu DRAM traffic of the two kernels, then the default bench line
# (new: reference_on_gpu leg, fused norms in the VQ-W-Net block) and the reference arm
O=gpurun_out/r02c38
mkdir -p $O
timeout 300 python -m pytest tests/test_norm_relu.py -q -m gpu -s > $O/pytest_norm.log 2>&1; echo "pytest rc $?"; grep "norm_relu\]" $O/pytest_norm.log | head -8; tail -2 $O/pytest_norm.log
timeout 200 python tools/norm_relu_bench.py 16 64 256 > $O/norm_relu_256.json 2> $O/norm_relu.err; cat $O/norm_relu_256.json
timeout 200 python tools/norm_relu_bench.py 16 64 512 > $O/norm_relu_512.json 2>> $O/norm_relu.err; cat $O/norm_relu_512.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,launch__cluster_dim_x,launch__grid_size --clock-control none -k regex:norm_relu -c 12 --csv --log-file $O/ncu_norm_relu_256.csv python tools/norm_relu_bench.py 16 64 256 > $O/ncu1.log 2>&1; echo "ncu256 rc $?"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,launch__cluster_dim_x,launch__grid_size --clock-control none -k regex:norm_relu -c 12 --csv --log-file $O/ncu_norm_relu_512.csv python tools/norm_relu_bench.py 16 64 512 > $O/ncu2.log 2>&1; echo "ncu512 rc $?"
tail -4 $O/ncu_norm_relu_512.csv
timeout 600 python bench.py > $O/bench_default.log 2> $O/bench_default.err; echo "bench rc $?"; tail -c 2500 $O/bench_default.log; tail -3 $O/bench_default.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.log 2> $O/bench_ref.err; echo "ref rc $?"; tail -c 600 $O/bench_ref.log
