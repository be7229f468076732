#!/usr/bin/env bash
# Builds world_modelz_b200/_C/libwm_exp7.so: the library with the clock64 timeline instrumentation of the
# attention kernels compiled in (-DWM_EXPERIMENT=7).  Use with  WM_B200_LIB=.../libwm_exp7.so python tools/dbg_timeline.py
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="${HERE}/../world_modelz_b200/csrc"; OUT="${HERE}/../world_modelz_b200/_C"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr)
OBJS=()
for src in api attn_simt attn_tc attn_tc_bwd_ws vq_exact vq_tc optim layer_ops train_ops; do
  case $src in
    attn_tc|attn_tc_bwd_ws) nvcc "${FLAGS[@]}" -DWM_EXPERIMENT=7 ${WM_EXTRA_DEFS:-} -c "${SRC}/${src}.cu" -o "${OUT}/${src}_exp7.o"; OBJS+=("${OUT}/${src}_exp7.o");;
    *) OBJS+=("${OUT}/${src}.o");;
  esac
done
nvcc -arch=sm_100a -shared -o "${OUT}/libwm_exp7.so" "${OBJS[@]}" -cudart static -Xlinker --exclude-libs=ALL
echo "built ${OUT}/libwm_exp7.so"
