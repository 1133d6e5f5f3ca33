#!/bin/bash
mkdir -p gpurun_out/r2p16
timeout 900 python -m pytest tests/test_score_gpu.py tests/test_long_ranges_gpu.py -x -q -m gpu > gpurun_out/r2p16/pytest.log 2>&1; tail -3 gpurun_out/r2p16/pytest.log
timeout 200 python scripts/soak_long.py 90 7 > gpurun_out/r2p16/soak_long_seed7.txt 2>&1; tail -2 gpurun_out/r2p16/soak_long_seed7.txt
python bench.py --config 4 --n 20000 --steps 3 --warmup 2 > gpurun_out/r2p16/cfg4_20k.json 2> gpurun_out/r2p16/cfg4_20k.err
python bench.py --config 4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2p16/cfg4_100k.json 2> gpurun_out/r2p16/cfg4_100k.err
python bench.py --config 4 --n 20000 --mode 3pass --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2p16/cfg4_3pass_20k.json 2> gpurun_out/r2p16/cfg4_3pass_20k.err
