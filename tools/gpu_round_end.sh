#!/bin/bash
# round-end validation + evidence run (one gpurun call): GPU suite, smoke, every bench line, ncu captures
O=gpurun_out/round_end; mkdir -p $O
timeout 900 python -m pytest tests -q -x -m gpu > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
timeout 400 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 python bench.py --detector content_edges --frames 4096 --steps 5 --warmup 3 --no-cpu --no-e2e > $O/bench_content_edges.json 2> $O/bench_content_edges.err
timeout 300 python bench.py --detector threshold --no-cpu --no-e2e > $O/bench_threshold.json 2> $O/bench_threshold.err
timeout 300 python bench.py --detector histogram --no-cpu --no-e2e > $O/bench_histogram.json 2> $O/bench_histogram.err
timeout 300 python bench.py --auto-downscale --frames 4096 --steps 5 --warmup 3 --no-cpu --no-e2e > $O/bench_autodownscale.json 2> $O/bench_autodownscale.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --frames 2048 --steps 2 --warmup 1 --no-cpu --no-e2e > $O/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:psd_score_ws_kernel -s 2 -c 1 -f -o $O/ws_v7_final python bench.py --frames 1024 --steps 2 --warmup 1 --no-cpu --no-e2e > $O/ncu_ws.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:psd_score_ws_kernel -s 2 -c 1 -f -o $O/ws_hist_final python bench.py --detector histogram --frames 1024 --steps 2 --warmup 1 --no-cpu --no-e2e > $O/ncu_hist.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:psd_canny_classify_bits_kernel -s 1 -c 1 -f -o $O/classify_bits python bench.py --detector content_edges --frames 512 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/ncu_tile.log 2>&1
for f in $O/bench*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d['value']), d.get('roofline',{}).get('frac'), (d.get('e2e') or {}).get('value'), d.get('gpu_launches'))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done
ls -la $O | head -40
