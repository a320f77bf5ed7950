#!/usr/bin/env bash
# one GPU: the driver's sequence on the final code of the round (backward kernels with four channel slices and the
# 64-register build, transposed lookup at four CTAs per SM, EmbeddingLoss accumulation folding equal keys per thread,
# VQ-W-Net leg with cuDNN algorithm search + fused Adam)
O=gpurun_out/r02c51
mkdir -p $O
timeout 400 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 $O/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc $?"; tail -2 $O/smoke.log
timeout 200 python tools/knob_ab.py > $O/knob_ab.jsonl 2> $O/knob_ab.err; echo "knob_ab rc $?"; grep -c kernel $O/knob_ab.jsonl
timeout 600 python bench.py > $O/bench.log 2> $O/bench.err; echo "bench rc $?"; tail -3 $O/bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02c51/bench.log") if l.startswith("{")][-1])
print("value %.4g ms/step %.4f eager %.4f kernel_ms %.4f frac %.3f e2e %.4g wnet %.1f" % (d["value"], d["ms_per_step"], d["eager"]["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["vqwnet_train"]["value"]))
PY
