#!/bin/bash
O=gpurun_out/r2p25; mkdir -p $O
for v in bal nobal; do
  if [ $v = nobal ]; then export ZOE_CUDA_NO_BALANCE=1; fi
  timeout 300 python bench.py --config 3 --n 125000 --steps 20 --warmup 5 --no-cpu-baseline > $O/cfg3_125k_$v.json 2> $O/cfg3_125k_$v.err
  timeout 300 python bench.py --config 2 --n 125000 --steps 10 --warmup 3 --legs none --no-cpu-baseline > $O/cfg2_125k_$v.json 2> $O/cfg2_125k_$v.err
  timeout 300 python bench.py --config 1 --steps 30 --warmup 5 --no-cpu-baseline > $O/cfg1_$v.json 2> $O/cfg1_$v.err
  timeout 300 python bench.py --config 3 --n 125000 --mode ranges --steps 20 --warmup 5 --no-cpu-baseline > $O/cfg3r_125k_$v.json 2> $O/cfg3r_125k_$v.err
  timeout 300 python bench.py --config 3 --n 250000 --steps 10 --warmup 5 --no-cpu-baseline > $O/cfg3_250k_$v.json 2> $O/cfg3_250k_$v.err
done
unset ZOE_CUDA_NO_BALANCE
timeout 300 python bench.py --config 3 --steps 8 --warmup 3 --no-cpu-baseline > $O/cfg3.json 2> $O/cfg3.err
timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1; tail -2 $O/pytest_all.log
timeout 200 python scripts/soak.py 100 606 > $O/soak_seed606.txt 2>&1; tail -1 $O/soak_seed606.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2p25/*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(j['value'],1), round(j['ms_per_step'],4), 'e2e', round(j['e2e']['value'],1), round(j['e2e']['ms_per_step'],4))
    except Exception as e: print(f, 'ERR', e)
PY
