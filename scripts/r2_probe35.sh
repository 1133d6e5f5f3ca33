#!/bin/bash
O=gpurun_out/r2p35; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu -k "long or align or mixed" > $O/pytest_sel.log 2>&1; tail -2 $O/pytest_sel.log
timeout 300 python bench.py --config 3 --steps 8 --warmup 3 > $O/cfg3.json 2> $O/cfg3.err
timeout 300 python bench.py --config 3 --n 125000 --steps 20 --warmup 5 --no-cpu-baseline > $O/cfg3_125k.json 2> $O/cfg3_125k.err
timeout 300 python bench.py --config 3 --mode 3pass --steps 6 --warmup 3 --no-cpu-baseline > $O/cfg3_3pass.json 2> $O/cfg3_3pass.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2p35/*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(j['value'],1), round(j['ms_per_step'],4), 'e2e', round(j['e2e']['value'],1), round(j['e2e']['ms_per_step'],3), j['e2e'].get('checksum_matches_n1'), (j.get('parity') or {}).get('mismatches'))
    except Exception as e: print(f, 'ERR', e)
PY
