#!/bin/bash
# multi-GPU evidence (gpurun --gpus N -- 'bash tools/gpu_multi.sh N [what]'): config 4 (8 GPUs), strong + weak scaling
# with the end-to-end leg, and the config-5 sweep at N ranks.
N=${1:-2}; WHAT=${2:-all}
O=gpurun_out/multi_n$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
if [ $N -eq 1 ]; then TR="python"; fi
if [ "$WHAT" = all ] || [ "$WHAT" = cfg4 ]; then
  timeout 600 $TR bench.py --gpus $N --detector histogram --width 3840 --height 2160 --frames 6250 --steps 3 --warmup 3 --no-e2e --no-cpu --parity-frames 4 > $O/cfg4_hist4k.json 2> $O/cfg4_hist4k.err
fi
if [ "$WHAT" = all ] || [ "$WHAT" = scale ]; then
  timeout 600 $TR bench.py --gpus $N --scaling strong --frames 10000 --steps 10 --warmup 3 --no-cpu --e2e-steps 2 > $O/strong_10k.json 2> $O/strong_10k.err
  timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --e2e-steps 2 > $O/weak_10k.json 2> $O/weak_10k.err
fi
if [ "$WHAT" = all ] || [ "$WHAT" = sweep ]; then
  timeout 900 $TR bench.py --gpus $N --sweep --steps 5 > $O/sweep.jsonl 2> $O/sweep.err
fi
for f in $O/*.json $O/*.jsonl; do python - "$f" <<'PY'
import json,sys
for line in open(sys.argv[1]).read().strip().splitlines():
    try:
        d=json.loads(line); e=d.get('e2e') or {}
        print(sys.argv[1].split('/')[-1], d['config'].get('workload','')[:60], 'N',d['n_gpus'], d['scaling'], round(d['value']), 'frac',round(d['roofline']['frac'],4), 'e2e', e.get('value') and round(e['value']), e.get('breakdown'), d.get('parity_check',{}).get('within_1e-4'), d.get('parity_check',{}).get('shard_boundaries_equal'))
    except Exception as ex: print(sys.argv[1], 'FAILED', ex)
PY
done | tee $O/summary.txt
tail -3 $O/*.err | tail -20
