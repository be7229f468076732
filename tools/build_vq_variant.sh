#!/usr/bin/env bash
# Builds world_modelz_b200/_C/libwm_vq_<name>.so from a variant source of vq_tc.cu (tuning experiments / timeline):
#   tools/build_vq_variant.sh <name> <source.cu> [extra nvcc flags, e.g. -DWM_VQ_EXP=32]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="${HERE}/../world_modelz_b200/csrc"; OUT="${HERE}/../world_modelz_b200/_C"
name=$1; src=$2; shift 2
cp "$src" "${SRC}/_vq_variant_${name}.cu"
trap 'rm -f "${SRC}/_vq_variant_${name}.cu"' EXIT
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr "$@" \
     -c "${SRC}/_vq_variant_${name}.cu" -o "${OUT}/vq_tc_${name}.o"
nvcc -arch=sm_100a -shared -o "${OUT}/libwm_vq_${name}.so" "${OUT}"/{api,attn_simt,attn_tc,attn_tc_bwd_ws,vq_exact,optim,layer_ops,train_ops}.o "${OUT}/vq_tc_${name}.o" \
     -cudart static -Xlinker --exclude-libs=ALL
echo "built ${OUT}/libwm_vq_${name}.so"
