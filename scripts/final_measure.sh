#!/bin/bash
# Round-end measurement batch (run on the GPU box through gpurun): GPU tests, every bench config, launch lists.
# Outputs go to gpurun_out/<tag>_*; copy what should be judged into profiles/.
TAG=${1:-final3}
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; tail -3 $O/${TAG}_pytest_gpu.log
python bench.py > $O/${TAG}_cfg2.json 2> $O/${TAG}_cfg2.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_cfg2_ref.json 2> $O/${TAG}_cfg2_ref.err
python bench.py --config 3 > $O/${TAG}_cfg3.json 2> $O/${TAG}_cfg3.err
python bench.py --config 3 --mode ranges > $O/${TAG}_cfg3_ranges.json 2> $O/${TAG}_cfg3_ranges.err
python bench.py --config 3 --mode 3pass > $O/${TAG}_cfg3_3pass.json 2> $O/${TAG}_cfg3_3pass.err
python bench.py --config 5 > $O/${TAG}_cfg5.json 2> $O/${TAG}_cfg5.err
python bench.py --config 4 --steps 3 --warmup 3 > $O/${TAG}_cfg4.json 2> $O/${TAG}_cfg4.err
python bench.py --config 1 > $O/${TAG}_cfg1.json 2> $O/${TAG}_cfg1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_cfg2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_cfg2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_cfg3_3pass.csv \
    python bench.py --config 3 --mode 3pass --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_cfg3_3pass.log 2>&1
for f in $O/${TAG}_cfg*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print(sys.argv[1], "value", round(d.get("value", 0)), "ms", round(d.get("ms_per_step", 0), 2), "e2e", round(e.get("value", 0)),
          "frac", round((d.get("roofline") or {}).get("frac", 0), 3), "parity", d.get("parity"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as ex:
    print(sys.argv[1], "unreadable", ex)
PY
done
