#!/bin/bash
O=gpurun_out/r2p26; mkdir -p $O
export ZOE_CUDA_DEBUG_GRID=1
for v in bal nobal; do
  if [ $v = nobal ]; then export ZOE_CUDA_NO_BALANCE=1; fi
  timeout 300 python bench.py --config 3 --n 125000 --steps 20 --warmup 5 --no-cpu-baseline > $O/cfg3_125k_$v.json 2> $O/cfg3_125k_$v.err
  timeout 300 python bench.py --config 1 --steps 30 --warmup 5 --no-cpu-baseline > $O/cfg1_$v.json 2> $O/cfg1_$v.err
  timeout 300 python bench.py --config 2 --n 20000 --steps 30 --warmup 5 --legs none --no-cpu-baseline > $O/cfg2_20k_$v.json 2> $O/cfg2_20k_$v.err
done
for f in $O/*.err; do echo $f; sort $f | uniq -c | sort -rn | head -4; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2p26/*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(j['value'],1), round(j['ms_per_step'],4), 'e2e', round(j['e2e']['value'],1), round(j['e2e']['ms_per_step'],4))
    except Exception as e: print(f, 'ERR', e)
PY
