#!/usr/bin/env bash
# one GPU: second pass of the launch-shape knobs (register cap of the backward / EmbeddingLoss accumulation, the accumulation
# with equal keys folded inside the thread), the parity suite with the new defaults, the quantiser line with them
O=gpurun_out/r02c50
mkdir -p $O
timeout 200 python tools/knob_ab.py > $O/knob_ab.jsonl 2> $O/knob_ab.err; echo "knob_ab rc $?"; cat $O/knob_ab.jsonl | cut -c1-260; tail -3 $O/knob_ab.err
timeout 400 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 $O/pytest_gpu.log
VQ_BWD_MINB=4 timeout 400 python -m pytest tests/test_parity_gpu.py -q -m gpu > $O/pytest_gpu_bwdminb4.log 2>&1; echo "pytest(bwd minb 4) rc $?"; tail -2 $O/pytest_gpu_bwdminb4.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-model --no-north-star --no-cpu > $O/bench_q.log 2> $O/bench_q.err; echo "bench rc $?"
VQ_BWD_MINB=4 timeout 200 python bench.py --steps 20 --warmup 3 --no-model --no-north-star --no-cpu > $O/bench_q_minb4.log 2> $O/bench_q_minb4.err; echo "bench(minb4) rc $?"
python - <<'PY'
import json
for n in ("bench_q", "bench_q_minb4"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r02c50/{n}.log") if l.startswith("{")][-1])
        print(n, "value %.4g ms/step %.4f eager %.4f kernel_ms %.4f e2e %.4g" % (d["value"], d["ms_per_step"], d["eager"]["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"]))
    except Exception as e:
        print(n, "no line", e)
PY
