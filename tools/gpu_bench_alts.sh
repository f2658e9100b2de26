#!/bin/bash
# Run on a GPU box (gpurun -- 'bash tools/gpu_bench_alts.sh'): bench the default library and every
# pyscenedetect_b200/csrc/build/alt_*.so built by tools/ws_alt_builds.sh; prints fps and roofline fraction.
O=gpurun_out/alts; mkdir -p $O
L=pyscenedetect_b200/libpsd_b200.so; cp $L /tmp/orig.so
run() { timeout 200 python bench.py --frames 4096 --steps 5 --warmup 3 --no-cpu --no-e2e > $O/bench_$1.json 2> $O/bench_$1.err; }
run default
for a in pyscenedetect_b200/csrc/build/alt_*.so; do [ -f "$a" ] || continue; t=$(basename $a .so); cp $a $L; run $t; done
cp /tmp/orig.so $L
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d['value']), round(d['roofline']['frac'],4))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done
