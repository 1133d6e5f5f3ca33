#!/bin/bash
O=gpurun_out/r2p22; mkdir -p $O
timeout 300 python bench.py --config 2 --steps 6 --warmup 3 --legs none > $O/cfg2.json 2> $O/cfg2.err
timeout 300 python bench.py --config 3 --steps 8 --warmup 3 --no-cpu-baseline > $O/cfg3.json 2> $O/cfg3.err
timeout 300 python bench.py --config 3 --n 125000 --steps 20 --warmup 5 --no-cpu-baseline > $O/cfg3_125k.json 2> $O/cfg3_125k.err
ZOE_CUDA_NO_SPLIT_UPLOAD=1 timeout 300 python bench.py --config 3 --n 125000 --steps 20 --warmup 5 --no-cpu-baseline > $O/cfg3_125k_nosplit.json 2> $O/cfg3_125k_nosplit.err
ZOE_CUDA_NO_SPLIT_UPLOAD=1 timeout 300 python bench.py --config 3 --steps 8 --warmup 3 --no-cpu-baseline > $O/cfg3_nosplit.json 2> $O/cfg3_nosplit.err
ZOE_CUDA_NO_SPLIT_UPLOAD=1 timeout 300 python bench.py --config 2 --steps 6 --warmup 3 --legs none --no-cpu-baseline > $O/cfg2_nosplit.json 2> $O/cfg2_nosplit.err
timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1; tail -3 $O/pytest_all.log
timeout 200 python scripts/soak.py 100 505 > $O/soak_seed505.txt 2>&1; tail -1 $O/soak_seed505.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2p22/*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], j['value'], j['ms_per_step'], 'e2e', j['e2e']['value'], j.get('parity'))
    except Exception as e: print(f, 'ERR', e)
PY
