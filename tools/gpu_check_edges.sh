#!/bin/bash
O=gpurun_out/edges_check; mkdir -p $O
timeout 900 python -m pytest tests -q -x -m gpu > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt
timeout 300 python bench.py --detector content_edges --frames 4096 --steps 5 --warmup 3 --no-cpu --no-e2e > $O/bench_content_edges.json 2> $O/bench_content_edges.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_edges.csv python bench.py --detector content_edges --frames 512 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/ncu_edges.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/edges_check/bench_content_edges.json').read().strip().splitlines()[-1]); print('edges', round(d['value']), d['gpu_launches'])
PY
