#!/bin/bash
O=gpurun_out/r2p24; mkdir -p $O
for v in staged bytetab; do
  if [ $v = bytetab ]; then export ZOE_CUDA_LIB=$PWD/zoe_b200/libzoe_cuda_bytetab.so; fi
  timeout 300 python bench.py --config 2 --steps 6 --warmup 3 --legs none --no-cpu-baseline > $O/cfg2_$v.json 2> $O/cfg2_$v.err
  timeout 300 python bench.py --config 3 --steps 8 --warmup 3 --no-cpu-baseline > $O/cfg3_$v.json 2> $O/cfg3_$v.err
  timeout 300 python bench.py --config 3 --n 125000 --steps 20 --warmup 5 --no-cpu-baseline > $O/cfg3_125k_$v.json 2> $O/cfg3_125k_$v.err
  timeout 300 python bench.py --config 4 --n 20000 --steps 4 --warmup 3 --no-cpu-baseline > $O/cfg4_$v.json 2> $O/cfg4_$v.err
  timeout 300 python bench.py --config 1 --steps 30 --warmup 5 --no-cpu-baseline > $O/cfg1_$v.json 2> $O/cfg1_$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2p24/*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], j['value'], j['ms_per_step'], 'e2e', j['e2e']['value'], j['e2e'].get('checksum_matches_n1'))
    except Exception as e: print(f, 'ERR', e)
PY
