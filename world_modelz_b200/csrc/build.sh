#!/usr/bin/env bash
# Build libwm_b200.so for sm_100a, in-tree (the .so is git-ignored but travels with gpurun).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../_C"
mkdir -p "${OUT}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17
       -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xptxas -v --expt-relaxed-constexpr)
OBJS=()
for src in api attn_simt attn_tc attn_tc_bwd_ws vq_exact vq_tc optim layer_ops train_ops; do
  if [[ ! -f "${OUT}/${src}.o" || "${HERE}/${src}.cu" -nt "${OUT}/${src}.o" || "${HERE}/wm_common.cuh" -nt "${OUT}/${src}.o" || "${HERE}/tc_common.cuh" -nt "${OUT}/${src}.o" || "${HERE}/attn_tc.cuh" -nt "${OUT}/${src}.o" \
        || "${HERE}/../../include/wm_b200.h" -nt "${OUT}/${src}.o" ]]; then
    "${NVCC}" "${FLAGS[@]}" -c "${HERE}/${src}.cu" -o "${OUT}/${src}.o" 2> "${OUT}/${src}.ptxas.log" || { cat "${OUT}/${src}.ptxas.log"; exit 1; }
  fi
  OBJS+=("${OUT}/${src}.o")
done
"${NVCC}" -arch=sm_100a -shared -o "${OUT}/libwm_b200.so" "${OBJS[@]}" -cudart static -Xlinker --exclude-libs=ALL
echo "built ${OUT}/libwm_b200.so"
