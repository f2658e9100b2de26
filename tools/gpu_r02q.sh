#!/bin/bash
# round 2, call Q (2 GPUs): the new paths sharded over two ranks (edge component, hash), auto-downscale parity line
O=gpurun_out/r02q; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519"
B="--steps 5 --warmup 3 --no-cpu --e2e-steps 1"
timeout 300 python bench.py --auto-downscale --frames 4096 --steps 5 --warmup 3 --no-cpu --no-e2e > $O/bench_autodownscale.json 2> $O/bench_autodownscale.err
timeout 400 $TR bench.py --gpus 2 --detector content_edges --frames 4096 $B > $O/n2_content_edges.json 2> $O/n2_content_edges.err
timeout 400 $TR bench.py --gpus 2 --detector hash --frames 4096 $B > $O/n2_hash.json 2> $O/n2_hash.err
timeout 400 $TR bench.py --gpus 2 --detector adaptive --frames 4096 $B > $O/n2_adaptive.json 2> $O/n2_adaptive.err
for f in $O/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e=d.get('e2e') or {}; p=d.get('parity_check') or {}
    print(sys.argv[1].split('/')[-1], 'N', d['n_gpus'], round(d['value']), round(d['roofline']['frac'],4), 'e2e', e.get('value') and round(e['value']), e.get('breakdown'), p.get('within_1e-4'), p.get('shard_boundaries_equal'))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done | tee $O/summary.txt
for f in $O/*.err; do echo == $f; tail -n 3 $f; done
