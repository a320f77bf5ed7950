#!/usr/bin/env bash
# one GPU: the parity suite with the opt-in third-generation resident kernel (VQ_B200_R3=1)
O=gpurun_out/r02c47
mkdir -p $O
VQ_B200_R3=1 timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_reference_networks_gpu.py -q -m gpu > $O/pytest_r3.log 2>&1; echo "pytest rc $?"; tail -5 $O/pytest_r3.log
