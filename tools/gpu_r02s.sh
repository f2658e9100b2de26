#!/bin/bash
# round 2, call S: edge sub-batch 2048, dilation bands of 64 rows
O=gpurun_out/r02s; mkdir -p $O
L=pyscenedetect_b200/libpsd_b200.so; cp $L /tmp/orig.so
edges() { timeout 300 python bench.py --detector content_edges --frames 4096 --steps 5 --warmup 3 --no-cpu --no-e2e $2 > $O/bench_content_edges_$1.json 2> $O/bench_content_edges_$1.err; }
edges default
edges b2048 "--edge-batch 2048"
cp pyscenedetect_b200/csrc/build/edgealt_dil64.so $L; edges dil64; edges dil64_b2048 "--edge-batch 2048"; cp /tmp/orig.so $L
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d['value']), round(d['roofline']['frac'],4), d.get('clocks',{}).get('sm_mhz'), d.get('parity_check',{}).get('bit_equal'))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done | tee $O/summary.txt
