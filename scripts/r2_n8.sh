#!/bin/bash
O=gpurun_out/r2n8; mkdir -p $O
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > $O/n8.json 2> $O/n8.err; echo "rc=$?"
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2n8/n8.json').read().strip().splitlines()[-1])
print('cfg2', j['value'], j['ms_per_step'], 'e2e', j['e2e']['value'], j['e2e'].get('checksum_matches_n1'), j['n_gpus'])
a=j.get('align') or j.get('legs',{}).get('align')
print('align', a['value'], a['ms_per_step'], 'e2e', a['e2e']['value'], a['e2e'].get('checksum_matches_n1'))
PY
