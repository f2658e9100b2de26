#!/bin/bash
# BASELINE.json configs[4]: ContentDetector batch sweep (resolution x frame count) on one GPU,
# HBM-roofline report.  Usage: tools/sweep.sh > profiles/rNN_sweep.jsonl
for res in "640 360" "1280 720" "1920 1080" "3840 2160"; do
  set -- $res
  for n in 1000 10000 100000; do
    bytes=$(( $1 * $2 * 3 * n ))
    if [ $bytes -gt 150000000000 ]; then continue; fi   # must fit in HBM next to the results
    timeout 600 python bench.py --width $1 --height $2 --frames $n --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'width':$1,'height':$2,'frames':$n,'fps':round(d['value'],1),'ms_per_step':round(d['ms_per_step'],3),'roofline_frac':round(d['roofline']['frac'],4),'achieved_gbs':round(d['roofline']['achieved'],1),'launches_per_step':d['gpu_launches']/d['steps'],'sm_mhz':d['clocks']['sm_mhz']}))"
  done
done
