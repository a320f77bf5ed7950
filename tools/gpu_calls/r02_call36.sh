#!/usr/bin/env bash
# one GPU: full GPU suite (streamed kernel now takes emb_dim 512; fused norm + relu; k-means golden), VQGAN shape timing,
# fused norm+relu timing, run_recon pipeline, VQ-W-Net step with fused norms
O=gpurun_out/r02c36
mkdir -p $O
timeout 600 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 $O/pytest_gpu.log
for data in noise clustered; do
  timeout 120 python tools/ab.py 512 64 4 $data >> $O/ab_d512.log 2>&1
  ABSIMT=1 timeout 120 python tools/ab.py 512 64 4 $data >> $O/ab_d512.log 2>&1
done
cat $O/ab_d512.log
timeout 200 python tools/norm_relu_bench.py 16 64 256 > $O/norm_relu_256.json 2> $O/norm_relu.err; cat $O/norm_relu_256.json
timeout 200 python tools/norm_relu_bench.py 16 64 512 > $O/norm_relu_512.json 2>> $O/norm_relu.err; cat $O/norm_relu_512.json
timeout 300 python tools/recon_bench.py 16 > $O/recon.json 2> $O/recon.err; cat $O/recon.json; tail -2 $O/recon.err
for f in none tail all; do
  timeout 300 python bench.py --workload vqwnet --steps 5 --warmup 3 --no-cpu --fused-norm $f > $O/wnet_$f.log 2> $O/wnet_$f.err; echo "wnet $f rc $?"
  python - $O/wnet_$f.log <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["fused_norm"], d["fused_norm_pairs"], "slices/s %.1f ms/step %.2f loss %.5f" % (d["value"], d["ms_per_step"], d["final_loss"]))
except Exception as e:
    print("parse failed", e)
PY
done
