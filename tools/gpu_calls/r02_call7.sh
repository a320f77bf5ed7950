#!/usr/bin/env bash
mkdir -p gpurun_out/r02c7
O=gpurun_out/r02c7
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -20 $O/smoke.log; exit 1; }
timeout 600 python -m pytest tests -m gpu -x -q --timeout 150 > $O/pytest.log 2>&1; rc=$?; echo "pytest rc $rc" >> $O/pytest.log
tail -4 $O/pytest.log
[ $rc -ne 0 ] && { grep -E "FAILED|Error|Timeout" $O/pytest.log | head -20; }
timeout 200 python bench.py --no-model --no-north-star --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('default', 'kernel_ms %.4f' % d['roofline']['kernel_ms'], 'step %.4f' % d['ms_per_step'], 'eager %.4f' % d['eager']['ms_per_step'], 'eval %.4f' % d['eval_forward']['ms_per_step'], d['parity_check'])
"
timeout 100 python tools/lookup_bench.py > $O/lookup.log 2>&1; tail -8 $O/lookup.log
