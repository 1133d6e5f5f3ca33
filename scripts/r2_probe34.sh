#!/bin/bash
O=gpurun_out/r2p34; mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1; tail -2 $O/pytest_all.log
timeout 300 python bench.py --config 3 --steps 8 --warmup 3 > $O/cfg3.json 2> $O/cfg3.err
timeout 300 python bench.py --config 3 --mode ranges --steps 6 --warmup 3 --no-cpu-baseline > $O/cfg3r.json 2> $O/cfg3r.err
timeout 300 python bench.py --config 3 --mode 3pass --steps 6 --warmup 3 --no-cpu-baseline > $O/cfg3_3pass.json 2> $O/cfg3_3pass.err
timeout 300 python bench.py --config 3 --n 125000 --steps 20 --warmup 5 --no-cpu-baseline > $O/cfg3_125k.json 2> $O/cfg3_125k.err
timeout 400 python bench.py --config 4 --n 20000 --mode 3pass --steps 2 --warmup 1 --no-cpu-baseline > $O/cfg4_3pass.json 2> $O/cfg4_3pass.err
timeout 300 python bench.py --config 3 --n 200000 --scoring 2,-3,-4,-1 --steps 5 --warmup 2 --no-cpu-baseline > $O/cfg3_tie.json 2> $O/cfg3_tie.err
timeout 200 python scripts/soak.py 100 1010 > $O/soak_seed1010.txt 2>&1; tail -1 $O/soak_seed1010.txt
timeout 200 python scripts/soak_long.py 10 > $O/soak_long_seed10.txt 2>&1; tail -1 $O/soak_long_seed10.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2p34/*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(j['value'],1), round(j['ms_per_step'],4), 'e2e', round(j['e2e']['value'],1), j['e2e'].get('checksum_matches_n1'), 'peak', round(j['roofline']['peak'],1), (j.get('parity') or {}).get('mismatches'))
    except Exception as e: print(f, 'ERR', e)
PY
