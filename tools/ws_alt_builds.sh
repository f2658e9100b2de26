#!/bin/bash
# Build alternate libpsd_b200.so variants of the fused score kernel (different compile-time switches)
# next to the default one, for A/B runs on a GPU box:
#   tools/ws_alt_builds.sh old:"-DPSD_WS_LOOP=0 -DPSD_WS_STAGES=3" u4:"-DPSD_WS_UNROLL=4"
#     ->  pyscenedetect_b200/csrc/build/alt_old.so, alt_u4.so
# then on the box:  tools/gpu_bench_alts.sh   (swaps each alt in, runs bench.py, restores the default)
set -e
cd "$(dirname "$0")/../pyscenedetect_b200/csrc"
make -j8 >/dev/null
mkdir -p build_alt
rm -f build/alt_*.so
for cfg in "$@"; do
  tag=${cfg%%:*}; flags=${cfg#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 -cudart static \
       $flags -diag-suppress 128 -c score_kernel.cu -o build_alt/score_kernel_$tag.o &
done
wait
for cfg in "$@"; do
  tag=${cfg%%:*}
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o build/alt_$tag.so build/engine.o \
       build_alt/score_kernel_$tag.o build/edge_kernels.o build/resize_kernel.o build/scan_kernels.o \
       build/cut_kernels.o build/hash_kernels.o build/synth_kernel.o
  echo "built build/alt_$tag.so: $(python ../../tools/sass_loop_stats.py build/alt_$tag.so | head -1 | sed 's/.*consumer loop//')"
done
