#!/bin/bash
mkdir -p gpurun_out/r2p5
timeout 1500 python -m pytest tests/test_long_ranges_gpu.py tests/test_exact_fast_gpu.py tests/test_ranges_gpu.py tests/test_3pass_gpu.py -x -q -m gpu > gpurun_out/r2p5/pytest.log 2>&1
tail -15 gpurun_out/r2p5/pytest.log
python bench.py --config 3 --n 200000 --scoring 4,-2,-3,-1 --steps 3 --warmup 1 > gpurun_out/r2p5/hz_new_n200000.json 2> gpurun_out/r2p5/hz_new_n200000.err
python bench.py --config 3 --scoring 4,-2,-3,-1 --steps 3 --warmup 1 > gpurun_out/r2p5/hz_new_1M.json 2> gpurun_out/r2p5/hz_new_1M.err
python bench.py --config 3 --n 125000 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2p5/cfg3_n125000.json 2> gpurun_out/r2p5/cfg3_n125000.err
python bench.py --config 3 --mode 3pass --steps 5 --warmup 2 > gpurun_out/r2p5/cfg3_3pass.json 2> gpurun_out/r2p5/cfg3_3pass.err
python bench.py --config 4 --n 20000 --mode ranges --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2p5/cfg4_ranges_20k.json 2> gpurun_out/r2p5/cfg4_ranges_20k.err
python bench.py --config 4 --n 20000 --mode 3pass --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2p5/cfg4_3pass_20k.json 2> gpurun_out/r2p5/cfg4_3pass_20k.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2p5/launches_hz_200k.csv python bench.py --config 3 --n 200000 --scoring 4,-2,-3,-1 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2p5/ncu_hz.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2p5/launches_cfg4_3pass.csv python bench.py --config 4 --n 20000 --mode 3pass --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2p5/ncu_cfg4.log 2>&1
