#!/bin/bash
mkdir -p gpurun_out/r2p13
timeout 1200 python -m pytest tests/test_align_gpu.py tests/test_ranges_gpu.py tests/test_width_policy_gpu.py -x -q -m gpu > gpurun_out/r2p13/pytest.log 2>&1
tail -3 gpurun_out/r2p13/pytest.log
python bench.py --config 3 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2p13/cfg3_1M.json 2> gpurun_out/r2p13/cfg3_1M.err
python bench.py --config 3 --mode ranges --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2p13/cfg3_ranges.json 2> gpurun_out/r2p13/cfg3_ranges.err
python bench.py --config 3 --n 125000 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2p13/cfg3_n125000.json 2> gpurun_out/r2p13/cfg3_n125000.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2p13/launches_cfg3_1M.csv python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2p13/ncu_cfg3.log 2>&1
