#!/bin/bash
# round 2, call P: stage-by-stage waits for the bandwidth-bound passes (histogram, byte sum), auto-downscale parity in bench
O=gpurun_out/r02p; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt
B="--frames 4096 --steps 5 --warmup 3 --no-cpu --no-e2e"
timeout 300 python bench.py --detector histogram --no-cpu --no-e2e > $O/bench_histogram.json 2> $O/bench_histogram.err
timeout 300 python bench.py --detector threshold --no-cpu --no-e2e > $O/bench_threshold.json 2> $O/bench_threshold.err
timeout 300 python bench.py --no-cpu --no-e2e > $O/bench_content.json 2> $O/bench_content.err
timeout 300 python bench.py --auto-downscale $B > $O/bench_autodownscale.json 2> $O/bench_autodownscale.err
timeout 300 python bench.py --detector histogram --width 3840 --height 2160 --frames 2048 --steps 5 --warmup 3 --no-cpu --no-e2e > $O/bench_histogram_4k.json 2> $O/bench_histogram_4k.err
for f in $O/bench*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d['value']), d.get('roofline',{}).get('frac'), d.get('gpu_launches'), (d.get('parity_check') or {}).get('within_1e-4'), (d.get('clocks') or {}).get('reasons'))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done | tee $O/summary.txt
tail -3 $O/*.err
