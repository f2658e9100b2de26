#!/bin/bash
# Build alternate libpsd_b200.so variants of the warp-specialised score kernel (different compile-time
# shapes) next to the default one, for A/B runs on a GPU box:
#   tools/ws_alt_builds.sh "26 3" "28 3" "24 4"      ->  pyscenedetect_b200/csrc/build/alt_w26s3.so ...
# then on the box:  tools/gpu_bench_alts.sh   (swaps each alt in, runs bench.py, restores the default)
set -e
cd "$(dirname "$0")/../pyscenedetect_b200/csrc"
make -j8 >/dev/null
mkdir -p build_alt
for cfg in "$@"; do
  set -- $cfg; w=$1; s=$2; tag=w${w}s${s}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 -cudart static \
       -DPSD_WS_WARPS=$w -DPSD_WS_STAGES=$s -diag-suppress 128 -c score_kernel.cu -o build_alt/score_kernel_$tag.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o build/alt_$tag.so build/engine.o \
       build_alt/score_kernel_$tag.o build/edge_kernels.o build/resize_kernel.o build/scan_kernels.o \
       build/cut_kernels.o build/synth_kernel.o
  echo "built build/alt_$tag.so"
done
