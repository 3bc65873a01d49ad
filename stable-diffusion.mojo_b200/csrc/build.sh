#!/usr/bin/env bash
# Builds libtsd_b200.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -Wno-deprecated-gpu-targets -std=c++17 -Xcompiler -fPIC
       -Xcompiler -fvisibility=hidden -DTSD_BUILD)
# TSD_LAB_TRACE=1: in-kernel cycle stamps for tools/lab/*trace*.py (slower hot loops; never for measurements)
[ "${TSD_LAB_TRACE:-0}" = "1" ] && FLAGS+=(-DTSD_LAB_TRACE)
[ "${TSD_LAB_NOSTORE:-0}" = "1" ] && FLAGS+=(-DTSD_LAB_NOSTORE)
BUILD=${TSD_BUILD_DIR:-build}
OUT=${TSD_OUT:-libtsd_b200.so}
mkdir -p "$BUILD"
SRCS=(gemm_tcgen05.cu attention_tcgen05.cu elementwise.cu runtime.cu models.cu models_clip.cu c_api.cu c_api_models.cu host_io.cu dist_nccl.cu)
pids=()
for s in "${SRCS[@]}"; do
  [ -f "$s" ] || continue
  o="$BUILD/${s%.cu}.o"
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ -n "$(find . -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer "$o" -print -quit)" ] \
     || [ ../../include/tsd_b200.h -nt "$o" ]; then
    "$NVCC" "${FLAGS[@]}" -c "$s" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
OBJS=()
for s in "${SRCS[@]}"; do [ -f "$BUILD/${s%.cu}.o" ] && OBJS+=("$BUILD/${s%.cu}.o"); done
"$NVCC" -Wno-deprecated-gpu-targets -shared -o "$OUT" "${OBJS[@]}" -lcudart_static -lpthread -ldl -lrt
echo "built $(pwd)/$OUT"
