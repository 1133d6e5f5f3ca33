#!/bin/bash
O=gpurun_out/r2p37; mkdir -p $O
timeout 40 python -m pytest tests/test_split_upload_gpu.py -x -q -m gpu 2>&1 | tail -1
timeout 40 python bench.py --config 3 --n 125000 --steps 20 --warmup 3 --no-cpu-baseline > $O/cfg3s.json 2> $O/cfg3s.err
timeout 40 python bench.py --config 3 --steps 4 --warmup 2 --no-cpu-baseline > $O/cfg3.json 2> $O/cfg3.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2p37/*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(j['ms_per_step'],4), 'e2e ms', round(j['e2e']['ms_per_step'],4), j['e2e'].get('checksum_matches_n1'), j['gpu_launches'])
    except Exception as e: print(f, 'ERR', e)
PY
