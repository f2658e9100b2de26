#!/bin/bash
# round 2, call A: correctness of the new consumer loop, A/B of loop shapes, pipes3 microbench, one ncu capture
O=gpurun_out/r02a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt
timeout 900 python -m pytest tests -q -x -m gpu > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt
timeout 120 tools/microbench/build/pipes3 > $O/pipes3.txt 2>&1
L=pyscenedetect_b200/libpsd_b200.so; cp $L /tmp/orig.so
run() { timeout 200 python bench.py --frames 4096 --steps 6 --warmup 3 --no-cpu --no-e2e > $O/bench_$1.json 2> $O/bench_$1.err; }
run default
for a in pyscenedetect_b200/csrc/build/alt_*.so; do [ -f "$a" ] || continue; t=$(basename $a .so); cp $a $L; run $t; done
cp /tmp/orig.so $L
run default2
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d['value']), round(d['roofline']['frac'],4), d.get('clocks',{}).get('sm_mhz'))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done | tee $O/summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:psd_score_ws_kernel -s 2 -c 1 -f -o $O/ws_loop1 python bench.py --frames 1024 --steps 2 --warmup 1 --no-cpu --no-e2e > $O/ncu_ws.log 2>&1
ls -la $O | tail -30
