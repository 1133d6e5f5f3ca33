#!/bin/bash
# one-GPU evidence run of HEAD: default bench line, reference arm, GPU test log, smoke, ncu launch list of the headline command
O=gpurun_out/r2fin; mkdir -p $O
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 1200 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
timeout 300 python bench.py --steps 2 --warmup 1 --legs none --no-cpu-baseline > $O/headline_plain.json 2> $O/headline_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_headline.csv \
   python bench.py --steps 2 --warmup 1 --legs none --no-cpu-baseline > $O/ncu_headline.log 2>&1; echo "ncu rc=$?"
timeout 300 python bench.py --config 3 --steps 2 --warmup 1 --no-cpu-baseline > $O/cfg3_plain.json 2> $O/cfg3_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg3.csv \
   python bench.py --config 3 --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_cfg3.log 2>&1; echo "ncu cfg3 rc=$?"
timeout 300 python scripts/soak.py 150 707 > $O/soak_seed707.txt 2>&1; tail -1 $O/soak_seed707.txt
timeout 300 python scripts/soak_long.py 8 > $O/soak_long_seed8.txt 2>&1; tail -1 $O/soak_long_seed8.txt
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2fin/bench_default.json').read().strip().splitlines()[-1])
print('cfg2', j['value'], j['ms_per_step'], 'e2e', j['e2e']['value'], 'frac', j['roofline']['frac'], 'parity', j.get('parity'))
for k,v in j.get('legs',{}).items():
    print(k, v.get('value'), v.get('ms_per_step'), 'e2e', (v.get('e2e') or {}).get('value'), 'parity', (v.get('parity') or {}).get('mismatches'))
r=json.loads(open('gpurun_out/r2fin/bench_reference.json').read().strip().splitlines()[-1])
print('reference', r.get('value'), r.get('cpu_baseline'))
PY
