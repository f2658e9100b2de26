#!/bin/bash
# round 2 final evidence run (one gpurun call, one GPU): GPU suite, smoke, every bench line, launch lists, ncu captures
O=gpurun_out/r02_final; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
timeout 400 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
B="--frames 4096 --steps 5 --warmup 3 --no-cpu --no-e2e"
timeout 300 python bench.py --detector content_edges $B > $O/bench_content_edges.json 2> $O/bench_content_edges.err
timeout 300 python bench.py --detector adaptive $B > $O/bench_adaptive.json 2> $O/bench_adaptive.err
timeout 300 python bench.py --detector hash $B > $O/bench_hash.json 2> $O/bench_hash.err
timeout 300 python bench.py --detector threshold --no-cpu --no-e2e > $O/bench_threshold.json 2> $O/bench_threshold.err
timeout 300 python bench.py --detector histogram --no-cpu --no-e2e > $O/bench_histogram.json 2> $O/bench_histogram.err
timeout 300 python bench.py --auto-downscale $B > $O/bench_autodownscale.json 2> $O/bench_autodownscale.err
timeout 600 python bench.py --sweep --steps 5 > $O/sweep_n1.jsonl 2> $O/sweep_n1.err
NC="ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv"
timeout 300 $NC --log-file $O/launches.csv python bench.py --frames 2048 --steps 2 --warmup 1 --no-cpu --no-e2e > $O/ncu_launches.log 2>&1
timeout 300 $NC --log-file $O/launches_content_edges.csv python bench.py --detector content_edges --frames 2048 --steps 2 --warmup 1 --no-cpu --no-e2e --parity-frames 0 > $O/ncu_launches_edges.log 2>&1
timeout 300 $NC --log-file $O/launches_hash.csv python bench.py --detector hash --frames 2048 --steps 2 --warmup 1 --no-cpu --no-e2e --parity-frames 0 > $O/ncu_launches_hash.log 2>&1
FULL="ncu --set full --clock-control none --import-source on -c 1 -f"
timeout 400 $FULL -k regex:psd_score_ws_kernel -s 2 -o $O/ws_hsv python bench.py --frames 1024 --steps 2 --warmup 1 --no-cpu --no-e2e --parity-frames 0 > $O/ncu_ws.log 2>&1
timeout 400 $FULL -k regex:psd_canny_classify_pairs_kernel -s 1 -o $O/classify_pairs python bench.py --detector content_edges --frames 512 --steps 1 --warmup 1 --no-cpu --no-e2e --parity-frames 0 --edge-batch 256 > $O/ncu_classify.log 2>&1
timeout 400 $FULL -k regex:psd_hyst_bits_kernel -s 1 -o $O/hyst_bits python bench.py --detector content_edges --frames 512 --steps 1 --warmup 1 --no-cpu --no-e2e --parity-frames 0 --edge-batch 256 > $O/ncu_hyst.log 2>&1
timeout 400 $FULL -k regex:psd_hash_rows_kernel -s 1 -o $O/hash_rows python bench.py --detector hash --frames 512 --steps 1 --warmup 1 --no-cpu --no-e2e --parity-frames 0 > $O/ncu_hash.log 2>&1
timeout 400 $FULL -k regex:psd_score_ws_kernel -s 2 -o $O/ws_hist python bench.py --detector histogram --frames 1024 --steps 2 --warmup 1 --no-cpu --no-e2e --parity-frames 0 > $O/ncu_hist.log 2>&1
for f in $O/bench*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d['value']), d.get('roofline',{}).get('frac'), (d.get('e2e') or {}).get('value'), d.get('gpu_launches'), (d.get('parity_check') or {}).get('within_1e-4'))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done | tee $O/summary.txt
ls -la $O | head -60
