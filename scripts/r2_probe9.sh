#!/bin/bash
mkdir -p gpurun_out/r2p9
timeout 1500 python -m pytest tests/test_long_ranges_gpu.py tests/test_align_gpu.py -x -q -m gpu > gpurun_out/r2p9/pytest.log 2>&1
tail -15 gpurun_out/r2p9/pytest.log
python bench.py --config 3 --n 125000 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2p9/cfg3_n125000.json 2> gpurun_out/r2p9/cfg3_n125000.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p9/launches_cfg3_125k.csv python bench.py --config 3 --n 125000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2p9/ncu_cfg3.log 2>&1
ZOE_CUDA_DEBUG=1 python bench.py --config 3 --n 125000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2p9/cfg3_dbg.json 2> gpurun_out/r2p9/cfg3_dbg.err
