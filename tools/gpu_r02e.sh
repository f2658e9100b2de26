#!/bin/bash
# round 2, call E: branch-free classify, HashDetector: correctness, edge benches, launch list, ncu
O=gpurun_out/r02e; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.txt 2>&1; tail -8 $O/pytest_gpu.txt
timeout 300 python bench.py --detector content_edges --frames 4096 --steps 5 --warmup 3 --no-cpu --no-e2e > $O/bench_content_edges.json 2> $O/bench_content_edges.err
timeout 300 python bench.py --detector adaptive --frames 4096 --steps 5 --warmup 3 --no-cpu --no-e2e > $O/bench_adaptive.json 2> $O/bench_adaptive.err
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d['value']), round(d['roofline']['frac'],4), d.get('clocks',{}).get('sm_mhz'), d.get('parity_check'))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done | tee $O/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_edges.csv python bench.py --detector content_edges --frames 512 --steps 2 --warmup 1 --no-cpu --no-e2e --parity-frames 0 > $O/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:psd_canny_classify_bits_kernel -s 1 -c 1 -f -o $O/classify_bits python bench.py --detector content_edges --frames 512 --steps 1 --warmup 1 --no-cpu --no-e2e --parity-frames 0 > $O/ncu_classify.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:psd_hyst_bits_kernel -s 1 -c 1 -f -o $O/hyst_bits python bench.py --detector content_edges --frames 512 --steps 1 --warmup 1 --no-cpu --no-e2e --parity-frames 0 > $O/ncu_hyst.log 2>&1
ls -la $O | tail -12
