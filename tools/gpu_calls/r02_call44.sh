#!/usr/bin/env bash
# one GPU: what the driver runs at round end -- pytest -m gpu, smoke(), the default bench line, the reference arm
O=gpurun_out/r02c44
mkdir -p $O
timeout 600 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc $?"; tail -2 $O/smoke.log
timeout 900 python bench.py > $O/bench_default.log 2> $O/bench_default.err; echo "bench rc $?"; tail -2 $O/bench_default.err
timeout 600 python bench.py --impl reference > $O/bench_ref.log 2> $O/bench_ref.err; echo "ref rc $?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02c44/bench_default.log") if l.startswith("{")][-1])
print("value %.4g ms %.4f e2e %.4g kernel_ms %.4f frac %.3f wnet %.1f cpu %.3g refgpu_ms %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["vqwnet_train"]["value"], d["cpu_baseline"]["value"], d["cpu_baseline"]["reference_on_gpu"]["ms_per_step"]))
r = json.loads([l for l in open("gpurun_out/r02c44/bench_ref.log") if l.startswith("{")][-1])
print("reference arm: %.4g lookups/s, %s" % (r["value"], r["cpu_baseline"]["kind"]))
PY
