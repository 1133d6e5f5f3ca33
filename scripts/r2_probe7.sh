#!/bin/bash
mkdir -p gpurun_out/r2p7
timeout 1500 python -m pytest tests/test_long_ranges_gpu.py tests/test_3pass_gpu.py -x -q -m gpu > gpurun_out/r2p7/pytest.log 2>&1
tail -15 gpurun_out/r2p7/pytest.log
python bench.py --config 4 --n 20000 --mode 3pass --steps 2 --warmup 1 > gpurun_out/r2p7/cfg4_3pass_20k.json 2> gpurun_out/r2p7/cfg4_3pass_20k.err
python bench.py --config 3 --mode 3pass --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/r2p7/cfg3_3pass.json 2> gpurun_out/r2p7/cfg3_3pass.err
python bench.py --config 4 --n 20000 --mode 3pass --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2p7/launches_cfg4_3pass.csv python bench.py --config 4 --n 20000 --mode 3pass --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2p7/ncu_cfg4.log 2>&1
python bench.py --config 3 --n 50000 --scoring 4,-2,-3,-1 --steps 1 --warmup 1 --no-cpu-baseline --align-opts 0,6,128 > /dev/null 2>&1 && ncu --set full --import-source on --clock-control none -k regex:sw_exact_fast_kernel -s 2 -c 1 -o gpurun_out/r2p7/exact_fast python bench.py --config 3 --n 50000 --scoring 4,-2,-3,-1 --steps 1 --warmup 1 --no-cpu-baseline --align-opts 0,6,128 > gpurun_out/r2p7/ncu_exact.log 2>&1
