#!/usr/bin/env bash
# Build an experimental variant of libvq_b200.so into build_variants/lib_<name>.so (select it with VQ_B200_LIB=...).
#   tools/build_variant.sh <name> [extra nvcc flags, e.g. -DVQ_TC_TIMING]
set -euo pipefail
name=$1; shift
cd "$(dirname "$0")/../medical_image_editing_b200/csrc"
out=../../build_variants
mkdir -p $out/obj_$name
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
for s in vq_kernels.cu vq_assign_tc.cu vq_assign_small.cu vq_embed_loss.cu vq_norm_relu.cu vq_capi.cu; do
  nvcc $FLAGS "$@" -c $s -o $out/obj_$name/${s%.cu}.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/lib_$name.so $out/obj_$name/*.o -cudart static
rm -rf $out/obj_$name
echo built build_variants/lib_$name.so
