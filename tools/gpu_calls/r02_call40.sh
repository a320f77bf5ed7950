#!/usr/bin/env bash
# one GPU: full GPU suite (prep2 with eight threads per code; warp-per-plane norm + relu), small-kernel timeline, norm + relu small
# planes, default bench without the cpu legs
O=gpurun_out/r02c40
mkdir -p $O
timeout 600 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -4 $O/pytest_gpu.log
timeout 200 python tools/step_profile.py > $O/step_profile.log 2>&1; tail -25 $O/step_profile.log
timeout 200 python tools/norm_relu_bench.py 16 1024 16 > $O/norm_relu_16.json 2> $O/norm_relu.err; cat $O/norm_relu_16.json
timeout 200 python tools/norm_relu_bench.py 16 512 32 > $O/norm_relu_32.json 2>> $O/norm_relu.err; cat $O/norm_relu_32.json
timeout 200 python tools/norm_relu_bench.py 16 256 64 > $O/norm_relu_64.json 2>> $O/norm_relu.err; cat $O/norm_relu_64.json
timeout 600 python bench.py --no-cpu > $O/bench_nocpu.log 2> $O/bench_nocpu.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02c40/bench_nocpu.log") if l.startswith("{")][-1])
print("value %.4g ms %.4f eager %.4f kernel_ms %.4f frac %.3f wnet %.1f" % (d["value"], d["ms_per_step"], d["eager"]["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["vqwnet_train"]["value"]))
PY
