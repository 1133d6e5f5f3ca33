#!/bin/bash
O=gpurun_out/r2p23; mkdir -p $O
timeout 300 python bench.py --config 2 --steps 6 --warmup 3 --legs none --no-cpu-baseline > $O/cfg2.json 2> $O/cfg2.err
timeout 300 python bench.py --config 3 --steps 8 --warmup 3 --no-cpu-baseline > $O/cfg3.json 2> $O/cfg3.err
timeout 300 python bench.py --config 3 --n 125000 --steps 20 --warmup 5 --no-cpu-baseline > $O/cfg3_125k.json 2> $O/cfg3_125k.err
ZOE_CUDA_NO_SPLIT_UPLOAD=1 timeout 300 python bench.py --config 3 --n 125000 --steps 20 --warmup 5 --no-cpu-baseline > $O/cfg3_125k_nosplit.json 2> $O/cfg3_125k_nosplit.err
ZOE_CUDA_NO_SPLIT_UPLOAD=1 timeout 300 python bench.py --config 3 --steps 8 --warmup 3 --no-cpu-baseline > $O/cfg3_nosplit.json 2> $O/cfg3_nosplit.err
timeout 300 python bench.py --config 3 --mode ranges --steps 8 --warmup 3 --no-cpu-baseline > $O/cfg3_ranges.json 2> $O/cfg3_ranges.err
timeout 600 python -m pytest tests -x -q -m gpu -k "align or ranges or score" > $O/pytest_sel.log 2>&1; tail -2 $O/pytest_sel.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2p23/*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], j['value'], j['ms_per_step'], 'e2e', j['e2e']['value'], j.get('checksum'))
    except Exception as e: print(f, 'ERR', e)
PY
