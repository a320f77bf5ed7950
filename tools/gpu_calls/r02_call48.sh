#!/usr/bin/env bash
# one GPU, last call of the round: harness knobs of the VQ-W-Net leg (cuDNN algorithm search, fused Adam) A/B, launch list
# of the final code, then the parity suite + smoke + the default line on the rebuilt library
O=gpurun_out/r02c48
mkdir -p $O
ab() {   # name, env, extra flags
  env $2 timeout 150 python bench.py --workload vqwnet --steps 10 --warmup 3 --no-cpu $3 > $O/wnet_$1.log 2> $O/wnet_$1.err; echo "wnet $1 rc $?"
  python - $O/wnet_$1.log $1 <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print(sys.argv[2], "slices/s %.1f ms/step %.2f e2e %.1f loss %.5f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["final_loss"]))
except Exception as e:
    print(sys.argv[2], "no line", e)
PY
}
ab base "A=1" ""
ab cudnnbench "A=1" "--cudnn-benchmark 1"
ab fusedadam "VQ_TRAINER_FUSED_ADAM=1" ""
CMD="python bench.py --steps 2 --warmup 3 --no-model --no-north-star --no-cpu --no-graphs"
timeout 200 $CMD > $O/plain.log 2>&1 && timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu_l.log 2>&1
echo "ncu launches rc $?"
timeout 400 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 $O/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc $?"; tail -2 $O/smoke.log
timeout 600 python bench.py > $O/bench.log 2> $O/bench.err; echo "bench rc $?"; tail -c 600 $O/bench.log; tail -3 $O/bench.err
