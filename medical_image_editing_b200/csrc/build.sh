#!/usr/bin/env bash
# Build libvq_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr"
OUT=libvq_b200.so
SRCS="vq_kernels.cu vq_assign_tc.cu vq_assign_small.cu vq_embed_loss.cu vq_norm_relu.cu vq_capi.cu"
# rebuild only when a source is newer than the library
if [ -f "$OUT" ] && [ -z "$(find $SRCS vq_common.cuh ../../include/vq_b200.h -newer $OUT 2>/dev/null)" ] && [ "${FORCE:-0}" != "1" ]; then
  echo "libvq_b200.so up to date"; exit 0
fi
OBJS=""
PIDS=""
for s in $SRCS; do
  o="${s%.cu}.o"
  $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} ${VQ_EXTRA_FLAGS:-} -c "$s" -o "$o" &
  PIDS="$PIDS $!"
  OBJS="$OBJS $o"
done
for p in $PIDS; do wait "$p"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" $OBJS -cudart static
echo "built $(pwd)/$OUT"
