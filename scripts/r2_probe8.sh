#!/bin/bash
mkdir -p gpurun_out/r2p8
timeout 1500 python -m pytest tests/test_long_ranges_gpu.py tests/test_3pass_gpu.py tests/test_exact_fast_gpu.py tests/test_align_gpu.py -x -q -m gpu > gpurun_out/r2p8/pytest.log 2>&1
tail -15 gpurun_out/r2p8/pytest.log
python bench.py --config 4 --n 20000 --mode 3pass --steps 2 --warmup 1 > gpurun_out/r2p8/cfg4_3pass_20k.json 2> gpurun_out/r2p8/cfg4_3pass_20k.err
python bench.py --config 3 --n 200000 --scoring 4,-2,-3,-1 --steps 3 --warmup 1 > gpurun_out/r2p8/hz_n200000.json 2> gpurun_out/r2p8/hz_n200000.err
python bench.py --config 3 --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/r2p8/cfg3.json 2> gpurun_out/r2p8/cfg3.err
python bench.py --config 4 --n 20000 --mode 3pass --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2p8/launches_cfg4_3pass.csv python bench.py --config 4 --n 20000 --mode 3pass --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2p8/ncu_cfg4.log 2>&1
