#!/bin/bash
O=gpurun_out/r2p32; mkdir -p $O
for rep in 1 2; do
for v in staged vbyte; do
  if [ $v = vbyte ]; then export ZOE_CUDA_LIB=$PWD/zoe_b200/libzoe_cuda_vbyte.so; else unset ZOE_CUDA_LIB; fi
  timeout 300 python bench.py --config 3 --steps 8 --warmup 3 --no-cpu-baseline > $O/cfg3_${v}_$rep.json 2> $O/cfg3_${v}_$rep.err
  timeout 300 python bench.py --config 3 --mode ranges --steps 6 --warmup 3 --no-cpu-baseline > $O/cfg3r_${v}_$rep.json 2> $O/cfg3r_${v}_$rep.err
  timeout 300 python bench.py --config 3 --n 125000 --steps 20 --warmup 5 --no-cpu-baseline > $O/cfg3s_${v}_$rep.json 2> $O/cfg3s_${v}_$rep.err
done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2p32/*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(j['value'],1), round(j['ms_per_step'],4), 'e2e', round(j['e2e']['value'],1), j['e2e'].get('checksum_matches_n1'), 'peak', round(j['roofline']['peak'],1))
    except Exception as e: print(f, 'ERR', e)
PY
