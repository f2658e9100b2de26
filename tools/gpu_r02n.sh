#!/bin/bash
# round 2, call N: V-histogram bins addressed by IDP4A (and what the histogram costs at all), hysteresis phase times
O=gpurun_out/r02n; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.txt 2>&1; tail -8 $O/pytest_gpu.txt
L=pyscenedetect_b200/libpsd_b200.so; cp $L /tmp/orig.so
edges() { timeout 300 python bench.py --detector content_edges --frames 4096 --steps 5 --warmup 3 --no-cpu --no-e2e $2 > $O/bench_content_edges_$1.json 2> $O/bench_content_edges_$1.err; }
edges default
cp pyscenedetect_b200/csrc/build/alt_novhist.so $L; edges novhist "--parity-frames 0"; cp /tmp/orig.so $L
cp pyscenedetect_b200/csrc/build/edgealt_hs.so $L
timeout 300 python bench.py --detector content_edges --frames 1024 --steps 2 --warmup 1 --no-cpu --no-e2e --parity-frames 0 --edge-batch 256 > $O/hyst_stats.log 2>&1
cp /tmp/orig.so $L
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d['value']), round(d['roofline']['frac'],4), d.get('clocks',{}).get('sm_mhz'), d.get('parity_check',{}).get('bit_equal'))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done | tee $O/summary.txt
grep -A16 "hyst launch 4" $O/hyst_stats.log | head -40
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_edges.csv python bench.py --detector content_edges --frames 2048 --steps 2 --warmup 1 --no-cpu --no-e2e --parity-frames 0 > $O/ncu_launches.log 2>&1
ls -la $O | tail -12
