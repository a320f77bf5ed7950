#!/usr/bin/env bash
mkdir -p gpurun_out/r02c32
O=gpurun_out/r02c32
{
for v in 1 2 3 0 1 2; do echo "VQ_LOOKUP_T64=$v"; VQ_LOOKUP_T64=$v timeout 100 python tools/lookup_bench.py; done
} > $O/lookup.log 2>&1; grep -v "^$" $O/lookup.log | tail -14
