#!/bin/bash
# one-off experiment driver: bench of the restructured ws loop at several consumer-warp counts, then the full GPU suite
O=gpurun_out/r01o; mkdir -p $O
L=pyscenedetect_b200/libpsd_b200.so; cp $L /tmp/orig.so
run() { PSD_HSV_VARIANT=$2 timeout 200 python bench.py --frames 4096 --steps 5 --warmup 3 --no-cpu --no-e2e > $O/bench_$1_v$2.json 2> $O/bench_$1_v$2.err; }
run w24 7; run w24 8; run w24 5
for a in w26 w28 w30; do cp pyscenedetect_b200/csrc/build/alt_$a.so $L; run $a 7; done
cp pyscenedetect_b200/csrc/build/alt_w28.so $L; run w28 8
cp /tmp/orig.so $L
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d['value']), round(d['roofline']['frac'],4))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done
timeout 1200 python -m pytest tests -q -x -m gpu > $O/pytest.txt 2>&1; tail -5 $O/pytest.txt
