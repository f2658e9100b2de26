#!/bin/bash
# one-off experiment driver for a gpurun call: ncu capture of the variant-7 kernel + bench of alternate builds
O=gpurun_out/r01n; mkdir -p $O
L=pyscenedetect_b200/libpsd_b200.so; cp $L /tmp/orig.so
PSD_HSV_VARIANT=7 timeout 400 ncu --set full --clock-control none --import-source on -k regex:psd_score_ws_kernel -s 2 -c 1 -f -o $O/ws_v7 python bench.py --frames 1024 --steps 2 --warmup 1 --no-cpu --no-e2e > $O/ncu.log 2>&1
run() { PSD_HSV_VARIANT=$2 PSD_CHUNK_FRAMES=$3 timeout 200 python bench.py --frames 4096 --steps 5 --warmup 3 --no-cpu --no-e2e > $O/bench_$1_v$2_c$3.json 2> $O/bench_$1_v$2_c$3.err; }
run base 7 128; run base 7 256; run base 8 256
for a in w28 w31 s4; do cp pyscenedetect_b200/csrc/build/alt_$a.so $L; run $a 7 128; run $a 8 128; done
cp /tmp/orig.so $L
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d['value']), round(d['roofline']['frac'],4))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done
ls -la $O
