#!/bin/bash
# round 2, call T: GPU suite + edge / adaptive lines with the final defaults (2048-frame sub-batches, dilation bands of 64 rows)
O=gpurun_out/r02t; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
B="--frames 4096 --steps 5 --warmup 3 --no-cpu --no-e2e"
timeout 300 python bench.py --detector content_edges $B > $O/bench_content_edges.json 2> $O/bench_content_edges.err
timeout 300 python bench.py --detector adaptive $B > $O/bench_adaptive.json 2> $O/bench_adaptive.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_content_edges.csv python bench.py --detector content_edges --frames 4096 --steps 2 --warmup 1 --no-cpu --no-e2e --parity-frames 0 > $O/ncu_launches_edges.log 2>&1
for f in $O/bench*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d['value']), d.get('roofline',{}).get('frac'), d.get('gpu_launches'), (d.get('parity_check') or {}).get('bit_equal'), (d.get('clocks') or {}).get('reasons'))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done | tee $O/summary.txt
