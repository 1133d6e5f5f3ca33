#!/bin/bash
O=gpurun_out/r2p36; mkdir -p $O
for rep in 1 2; do
for v in split nosplit; do
  if [ $v = nosplit ]; then export ZOE_CUDA_NO_SPLIT_UPLOAD=1; else unset ZOE_CUDA_NO_SPLIT_UPLOAD; fi
  timeout 100 python bench.py --config 3 --n 125000 --steps 30 --warmup 5 --no-cpu-baseline > $O/cfg3s_${v}_$rep.json 2> $O/cfg3s_${v}_$rep.err
done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2p36/*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(j['ms_per_step'],4), 'e2e ms', round(j['e2e']['ms_per_step'],4))
    except Exception as e: print(f, 'ERR', e)
PY
