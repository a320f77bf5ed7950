#!/usr/bin/env bash
# one GPU: launch-shape knobs of the memory-bound kernels (tools/knob_ab.py), the parity suite with every candidate value
# switched on, and the VQ-W-Net leg with the harness defaults of call 48 (cuDNN algorithm search + fused Adam)
O=gpurun_out/r02c49
mkdir -p $O
timeout 200 python tools/knob_ab.py > $O/knob_ab.jsonl 2> $O/knob_ab.err; echo "knob_ab rc $?"; cat $O/knob_ab.jsonl | cut -c1-250; tail -3 $O/knob_ab.err
VQ_BWD_DSPLIT=2 VQ_EL_ACC_DSPLIT=4 VQ_EL_BWD_DSPLIT=2 VQ_LOOKUP_MINB=4 timeout 400 python -m pytest tests -q -m gpu > $O/pytest_gpu_knobs.log 2>&1; echo "pytest(knobs) rc $?"; tail -3 $O/pytest_gpu_knobs.log
timeout 150 python bench.py --workload vqwnet --steps 10 --warmup 3 --no-cpu > $O/wnet_defaults.log 2> $O/wnet_defaults.err; echo "wnet rc $?"; cut -c1-300 $O/wnet_defaults.log; tail -2 $O/wnet_defaults.err
