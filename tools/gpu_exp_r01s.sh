#!/bin/bash
# one-off experiment driver: strided hysteresis scans (parity + bench + launch list) + other detectors
O=gpurun_out/r01t; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "edge or golden or smoke or full_size" > $O/pytest_edges.txt 2>&1; tail -3 $O/pytest_edges.txt
timeout 300 python bench.py --detector content_edges --frames 2048 --steps 3 --warmup 2 --no-cpu --no-e2e > $O/bench_edges.json 2> $O/bench_edges.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_edges.csv python bench.py --detector content_edges --frames 512 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/ncu_edges.log 2>&1
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d['value']), round(d['roofline']['frac'],4), d.get('gpu_launches'))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done
