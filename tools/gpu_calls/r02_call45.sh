#!/usr/bin/env bash
# one GPU: warp-per-plane norm + relu with 2 / 8 / 32 float4 per lane: tests, timings of the deep U-Net levels, VQ-W-Net step,
# cluster sizes of the 512^2 backward
O=gpurun_out/r02c45
mkdir -p $O
timeout 300 python -m pytest tests/test_norm_relu.py -q -m gpu > $O/pytest_norm.log 2>&1; echo "pytest rc $?"; tail -2 $O/pytest_norm.log
for cfg in "16 1024 16" "16 512 32" "16 256 64" "16 128 128"; do
  timeout 200 python tools/norm_relu_bench.py $cfg >> $O/norm_relu_small.jsonl 2>> $O/norm_relu.err
done
python - <<'PY'
import json
for l in open("gpurun_out/r02c45/norm_relu_small.jsonl"):
    d = json.loads(l); f, t = d["fused"], d["torch"]
    print(d["shape"], "fused fwd %.3f bwd %.3f ms (%.0f%% / %.0f%% of HBM) | torch %.3f / %.3f" % (f["fwd_ms"], f["bwd_ms"], 100 * f["fwd_frac_of_hbm"], 100 * f["bwd_frac_of_hbm"], t["fwd_ms"], t["bwd_ms"]))
PY
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__cluster_dim_x,launch__grid_size --clock-control none -k regex:vq_norm_relu_bwd -c 3 --csv --log-file $O/ncu_bwd_512.csv python tools/norm_relu_once.py 512 > $O/ncu_bwd.log 2>&1; echo "ncu rc $?"; grep -E "cluster_dim_x|grid_size|duration|dram" $O/ncu_bwd_512.csv | tail -5 | cut -d, -f13-
timeout 300 python bench.py --workload vqwnet --steps 10 --warmup 3 --no-cpu > $O/wnet.log 2> $O/wnet.err; echo "wnet rc $?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02c45/wnet.log") if l.startswith("{")][-1])
print("wnet slices/s %.1f ms/step %.2f" % (d["value"], d["ms_per_step"]))
PY
