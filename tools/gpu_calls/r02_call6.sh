#!/usr/bin/env bash
mkdir -p gpurun_out/r02c6
O=gpurun_out/r02c6
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -20 $O/smoke.log; exit 1; }
run() { lib=$1; if [ "$lib" = default ]; then timeout 150 python bench.py --no-model --no-north-star --no-cpu; else VQ_B200_LIB=build_variants/lib_$lib.so timeout 150 python bench.py --no-model --no-north-star --no-cpu; fi; }
for lib in r01 default aux0 pf0 ilp1 plain r01 default; do
  run $lib 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', 'kernel_ms %.4f' % d['roofline']['kernel_ms'], 'step %.4f' % d['ms_per_step'], 'eager %.4f' % d['eager']['ms_per_step'], 'eval %.4f' % d['eval_forward']['ms_per_step'], 'fb', d['eval_forward']['fallback_rows'])
"
done | tee $O/bench_variants.log
