#!/bin/bash
# round 2, call G: pixel-pair classify kernel vs the 32-bit one, weak-candidate tile flags, I2F lift A/B, edge launch list
O=gpurun_out/r02g; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.txt 2>&1; tail -8 $O/pytest_gpu.txt
L=pyscenedetect_b200/libpsd_b200.so; cp $L /tmp/orig.so
run() { timeout 200 python bench.py --frames 4096 --steps 6 --warmup 3 --no-cpu --no-e2e > $O/bench_$1.json 2> $O/bench_$1.err; }
run default
for a in pyscenedetect_b200/csrc/build/alt_*.so; do [ -f "$a" ] || continue; t=$(basename $a .so); cp $a $L; run $t; done
cp /tmp/orig.so $L
run default2
edges() { timeout 300 python bench.py --detector content_edges --frames 4096 --steps 5 --warmup 3 --no-cpu --no-e2e > $O/bench_content_edges_$1.json 2> $O/bench_content_edges_$1.err; }
edges pairs
cp pyscenedetect_b200/csrc/build/edgealt_old.so $L; edges oldclassify; cp /tmp/orig.so $L
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d['value']), round(d['roofline']['frac'],4), d.get('clocks',{}).get('sm_mhz'), d.get('parity_check',{}).get('bit_equal'))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done | tee $O/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_edges.csv python bench.py --detector content_edges --frames 512 --steps 2 --warmup 1 --no-cpu --no-e2e --parity-frames 0 > $O/ncu_launches.log 2>&1
ls -la $O | tail -8
