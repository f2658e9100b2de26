#!/bin/bash
# round 2, call B: persistent fused pass - correctness, A/B (unroll, chunking), ncu
O=gpurun_out/r02b; mkdir -p $O
timeout 900 python -m pytest tests -q -x -m gpu > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt
L=pyscenedetect_b200/libpsd_b200.so; cp $L /tmp/orig.so
run() { timeout 200 python bench.py --frames 4096 --steps 6 --warmup 3 --no-cpu --no-e2e > $O/bench_$1.json 2> $O/bench_$1.err; }
run default
PSD_CHUNKS=32 run c32
PSD_CHUNKS=74 run c74
PSD_CHUNKS=148 run c148
for a in pyscenedetect_b200/csrc/build/alt_*.so; do [ -f "$a" ] || continue; t=$(basename $a .so); cp $a $L; run $t; done
cp /tmp/orig.so $L
timeout 300 python bench.py --no-cpu --no-e2e > $O/bench_full10k.json 2> $O/bench_full10k.err
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d['value']), round(d['roofline']['frac'],4), d.get('clocks',{}).get('sm_mhz'))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done | tee $O/summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:psd_score_ws_kernel -s 2 -c 1 -f -o $O/ws_persist python bench.py --frames 1024 --steps 2 --warmup 1 --no-cpu --no-e2e > $O/ncu_ws.log 2>&1
ls -la $O | tail -30
