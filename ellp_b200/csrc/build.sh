#!/usr/bin/env bash
# Builds ellp_b200/libellp_b200.so for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libellp_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false
       -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function -shared)
if [[ "${ELLP_PTXAS_V:-0}" == "1" ]]; then FLAGS+=(-Xptxas -v); fi
"${NVCC}" "${FLAGS[@]}" -o "${OUT}" "${HERE}/engine.cu" "${HERE}/host_model.cpp"
echo "built ${OUT}"
