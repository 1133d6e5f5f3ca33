#!/bin/bash
O=gpurun_out/r2p30; mkdir -p $O
for v in split vhead vscan vstaged vhead2; do
  case $v in
    split) unset ZOE_CUDA_LIB;;
    vhead2) export ZOE_CUDA_LIB=$PWD/zoe_b200/libzoe_cuda_vhead.so;;
    *) export ZOE_CUDA_LIB=$PWD/zoe_b200/libzoe_cuda_$v.so;;
  esac
  timeout 300 python bench.py --config 2 --steps 4 --warmup 3 --legs none --no-cpu-baseline > $O/cfg2_$v.json 2> $O/cfg2_$v.err
  timeout 300 python bench.py --config 3 --steps 8 --warmup 3 --no-cpu-baseline > $O/cfg3_$v.json 2> $O/cfg3_$v.err
  timeout 300 python bench.py --config 3 --mode ranges --steps 6 --warmup 3 --no-cpu-baseline > $O/cfg3r_$v.json 2> $O/cfg3r_$v.err
  timeout 300 python bench.py --config 4 --n 20000 --steps 3 --warmup 3 --no-cpu-baseline > $O/cfg4_$v.json 2> $O/cfg4_$v.err
  timeout 300 python bench.py --config 5 --steps 6 --warmup 3 --no-cpu-baseline > $O/cfg5_$v.json 2> $O/cfg5_$v.err
  timeout 300 python bench.py --config 1 --steps 30 --warmup 5 --no-cpu-baseline > $O/cfg1_$v.json 2> $O/cfg1_$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2p30/*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(j['value'],1), round(j['ms_per_step'],4), 'e2e', round(j['e2e']['value'],1), j['e2e'].get('checksum_matches_n1'), 'peak', round(j['roofline']['peak'],1))
    except Exception as e: print(f, 'ERR', e)
PY
