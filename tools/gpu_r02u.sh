#!/bin/bash
# round 2, call U: the driver's own command on the final tree
O=gpurun_out/r02u; mkdir -p $O
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
python - $O/bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d['value']), d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'], d['parity_check'], d['clocks'], d['cpu_baseline']['value'])
PY
tail -n 3 $O/bench.err
