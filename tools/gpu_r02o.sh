#!/bin/bash
# round 2, call O: hysteresis back on per-warp runs (32 tiles, round-robin), A/B against one contiguous run per warp
O=gpurun_out/r02o; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.txt 2>&1; tail -4 $O/pytest_gpu.txt
L=pyscenedetect_b200/libpsd_b200.so; cp $L /tmp/orig.so
edges() { timeout 300 python bench.py --detector content_edges --frames 4096 --steps 5 --warmup 3 --no-cpu --no-e2e $2 > $O/bench_content_edges_$1.json 2> $O/bench_content_edges_$1.err; }
edges default
for a in pyscenedetect_b200/csrc/build/edgealt_*.so; do t=$(basename $a .so); t=${t#edgealt_}; cp $a $L; edges $t; done; cp /tmp/orig.so $L
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d['value']), round(d['roofline']['frac'],4), d.get('clocks',{}).get('sm_mhz'), d.get('parity_check',{}).get('bit_equal'))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done | tee $O/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_edges.csv python bench.py --detector content_edges --frames 2048 --steps 2 --warmup 1 --no-cpu --no-e2e --parity-frames 0 > $O/ncu_launches.log 2>&1
ls -la $O | tail -8
